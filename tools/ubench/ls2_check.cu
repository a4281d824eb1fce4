// ls2_check.cu -- stand-alone check + timing of the pipelined lane-serial backward (fm_scan_bwd_ls2.cuh) against the first
// lane-serial kernel (fm_scan_bwd_ls.cuh, parity-tested) on the same random inputs at BASELINE configs[1].
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o ls2_check tools/ubench/ls2_check.cu
//   ./ls2_check [batch] [L]
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../fusionmamba_b200/csrc/fm_scan_bwd_ls2.cuh"
namespace fm {
void count_launch() {}
int env_int(const char* name, int dflt) { const char* v = getenv(name); return v ? atoi(v) : dflt; }
}
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
static double cmp(const char* name, const float* a, const float* b, size_t n) {
    std::vector<float> ha(n), hb(n);
    cudaMemcpy(ha.data(), a, n * 4, cudaMemcpyDeviceToHost); cudaMemcpy(hb.data(), b, n * 4, cudaMemcpyDeviceToHost);
    double mx = 0, md = 0; size_t bad = 0;
    for (size_t i = 0; i < n; ++i) { mx = fmax(mx, fabs(ha[i])); double d = fabs((double)ha[i] - hb[i]); if (!(d == d)) ++bad; else md = fmax(md, d); }
    printf("  %-8s max|ref| %.4g  max|diff| %.3g  rel %.3g  nan %zu\n", name, mx, md, md / (mx + 1e-30), bad);
    return bad ? 1e9 : md / (mx + 1e-30);
}
int main(int argc, char** argv) {
    const int Bn = argc > 1 ? atoi(argv[1]) : 8, dim = 768, L = argc > 2 ? atoi(argv[2]) : 4096, N = 16, G = 4;
    const size_t E = (size_t)Bn * dim * L, Gx = (size_t)Bn * G * N * L;
    const int n_hck = (L + 7) / 8 - 1;
    float *u, *dl, *g, *Bm, *Cm, *A, *D, *bias, *hck;
    float *du[2], *dd[2], *dB[2], *dC[2], *dA[2], *dD[2], *dbias[2];
    CK(cudaMalloc(&u, E * 4)); CK(cudaMalloc(&dl, E * 4)); CK(cudaMalloc(&g, E * 4));
    CK(cudaMalloc(&Bm, Gx * 4)); CK(cudaMalloc(&Cm, Gx * 4));
    CK(cudaMalloc(&A, dim * N * 4)); CK(cudaMalloc(&D, dim * 4)); CK(cudaMalloc(&bias, dim * 4));
    const size_t nh = (size_t)Bn * dim * (n_hck > 0 ? n_hck : 1) * N;
    CK(cudaMalloc(&hck, nh * 4));
    for (int i = 0; i < 2; ++i) {
        CK(cudaMalloc(&du[i], E * 4)); CK(cudaMalloc(&dd[i], E * 4)); CK(cudaMalloc(&dB[i], Gx * 4)); CK(cudaMalloc(&dC[i], Gx * 4));
        CK(cudaMalloc(&dA[i], dim * N * 4)); CK(cudaMalloc(&dD[i], dim * 4)); CK(cudaMalloc(&dbias[i], dim * 4));
    }
    std::vector<float> h(E > nh ? E : nh);
    auto fill = [&](float* d, size_t n, float lo, float hi) {
        for (size_t i = 0; i < n; ++i) h[i] = lo + (hi - lo) * (rand() / (float)RAND_MAX);
        cudaMemcpy(d, h.data(), n * 4, cudaMemcpyHostToDevice);
    };
    fill(u, E, -1, 1); fill(dl, E, -3.f, 1.5f); fill(g, E, -1, 1); fill(Bm, Gx, -1, 1); fill(Cm, Gx, -1, 1);
    fill(A, dim * N, -0.5f, 0); fill(D, dim, -1, 1); fill(bias, dim, 0, 0.5f); fill(hck, nh, -1, 1);
    FmScanBwdParams q = {};
    FmScanFwdParams& p = q.f;
    p.dtype = FM_F32; p.batch = Bn; p.dim = dim; p.seqlen = L; p.dstate = N; p.n_groups = G; p.delta_softplus = 1;
    p.hck_len = 8; p.n_hck = n_hck; p.hck = hck;
    p.u_batch_stride = p.delta_batch_stride = (int64_t)dim * L; p.u_d_stride = p.delta_d_stride = L;
    p.A_d_stride = N; p.A_dstate_stride = 1;
    p.B_batch_stride = p.C_batch_stride = (int64_t)G * N * L; p.B_group_stride = p.C_group_stride = (int64_t)N * L;
    p.B_dstate_stride = p.C_dstate_stride = L;
    p.u = u; p.delta = dl; p.A = A; p.B = Bm; p.C = Cm; p.D = D; p.delta_bias = bias;
    q.dout_batch_stride = q.du_batch_stride = q.ddelta_batch_stride = (int64_t)dim * L;
    q.dout_d_stride = q.du_d_stride = q.ddelta_d_stride = L;
    q.dB_batch_stride = q.dC_batch_stride = (int64_t)G * N * L; q.dB_group_stride = q.dC_group_stride = (int64_t)N * L;
    q.dB_dstate_stride = q.dC_dstate_stride = L;
    q.dout = g;
    const int vec = (L % 4 == 0);
    auto bind = [&](int i) {
        q.du = du[i]; q.ddelta = dd[i]; q.dA = dA[i]; q.dB = dB[i]; q.dC = dC[i]; q.dD = dD[i]; q.ddelta_bias = dbias[i];
    };
    auto zero = [&](int i) {
        cudaMemset(dB[i], 0, Gx * 4); cudaMemset(dC[i], 0, Gx * 4); cudaMemset(dA[i], 0, dim * N * 4);
        cudaMemset(dD[i], 0, dim * 4); cudaMemset(dbias[i], 0, dim * 4); cudaMemset(du[i], 0, E * 4); cudaMemset(dd[i], 0, E * 4);
    };
    zero(0); zero(1);
    bind(0); CK(fm::launch_scan_bwd_ls_T<float>(q, 0, vec, vec));
    bind(1); CK(fm::launch_scan_bwd_ls2_T<float>(q, 0, vec, vec));
    CK(cudaDeviceSynchronize());
    double worst = 0;
    worst = fmax(worst, cmp("du", du[0], du[1], E)); worst = fmax(worst, cmp("ddelta", dd[0], dd[1], E));
    worst = fmax(worst, cmp("dB", dB[0], dB[1], Gx)); worst = fmax(worst, cmp("dC", dC[0], dC[1], Gx));
    worst = fmax(worst, cmp("dA", dA[0], dA[1], dim * N)); worst = fmax(worst, cmp("dD", dD[0], dD[1], dim));
    worst = fmax(worst, cmp("dbias", dbias[0], dbias[1], dim));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int it = 10;
    float ms[2];
    for (int which = 0; which < 2; ++which) {
        bind(which);
        for (int i = 0; i < 3; ++i) CK(which ? fm::launch_scan_bwd_ls2_T<float>(q, 0, vec, vec) : fm::launch_scan_bwd_ls_T<float>(q, 0, vec, vec));
        CK(cudaDeviceSynchronize());
        cudaEventRecord(e0);
        for (int i = 0; i < it; ++i) CK(which ? fm::launch_scan_bwd_ls2_T<float>(q, 0, vec, vec) : fm::launch_scan_bwd_ls_T<float>(q, 0, vec, vec));
        cudaEventRecord(e1); CK(cudaDeviceSynchronize());
        cudaEventElapsedTime(&ms[which], e0, e1);
    }
    printf("{\"batch\": %d, \"L\": %d, \"ls_us\": %.1f, \"ls2_us\": %.1f, \"worst_rel\": %.3g, \"ok\": %s}\n", Bn, L, ms[0] / it * 1e3,
           ms[1] / it * 1e3, worst, worst < 2e-4 ? "true" : "false");
    return 0;
}
