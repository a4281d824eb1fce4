// ls_diag.cu -- stand-alone timing harness for the lane-serial backward kernel (fm_scan_bwd_ls.cuh) at BASELINE configs[1],
// used with -DFM_LS_DIAG=<mask> to measure what each part of the kernel costs (results are wrong for mask != 0).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -DFM_LS_DIAG=0 -o ls_diag0 tools/ubench/ls_diag.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../fusionmamba_b200/csrc/fm_scan_bwd_ls.cuh"
namespace fm {
void count_launch() {}
int env_int(const char* name, int dflt) { const char* v = getenv(name); return v ? atoi(v) : dflt; }
}
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
int main(int argc, char** argv) {
    const int Bn = argc > 1 ? atoi(argv[1]) : 8, dim = 768, L = argc > 2 ? atoi(argv[2]) : 4096, N = 16, G = 4;
    const size_t E = (size_t)Bn * dim * L, Gx = (size_t)Bn * G * N * L;
    const int n_hck = (L + 7) / 8 - 1;
    float *u, *dl, *g, *du, *dd, *Bm, *Cm, *dB, *dC, *A, *D, *bias, *dA, *dD, *dbias, *hck;
    CK(cudaMalloc(&u, E * 4)); CK(cudaMalloc(&dl, E * 4)); CK(cudaMalloc(&g, E * 4)); CK(cudaMalloc(&du, E * 4)); CK(cudaMalloc(&dd, E * 4));
    CK(cudaMalloc(&Bm, Gx * 4)); CK(cudaMalloc(&Cm, Gx * 4)); CK(cudaMalloc(&dB, Gx * 4)); CK(cudaMalloc(&dC, Gx * 4));
    CK(cudaMalloc(&A, dim * N * 4)); CK(cudaMalloc(&dA, dim * N * 4)); CK(cudaMalloc(&D, dim * 4)); CK(cudaMalloc(&bias, dim * 4));
    CK(cudaMalloc(&dD, dim * 4)); CK(cudaMalloc(&dbias, dim * 4)); CK(cudaMalloc(&hck, (size_t)Bn * dim * n_hck * N * 4));
    std::vector<float> h(E);
    auto fill = [&](float* d, size_t n, float lo, float hi) {
        for (size_t i = 0; i < n; ++i) h[i] = lo + (hi - lo) * (rand() / (float)RAND_MAX);
        cudaMemcpy(d, h.data(), n * 4, cudaMemcpyHostToDevice);
    };
    fill(u, E, -1, 1); fill(dl, E, 0, 0.5f); fill(g, E, -1, 1); fill(Bm, Gx, -1, 1); fill(Cm, Gx, -1, 1);
    fill(A, dim * N, -0.5f, 0); fill(D, dim, -1, 1); fill(bias, dim, 0, 0.5f);
    CK(cudaMemset(hck, 0, (size_t)Bn * dim * n_hck * N * 4)); CK(cudaMemset(dB, 0, Gx * 4)); CK(cudaMemset(dC, 0, Gx * 4));
    CK(cudaMemset(dA, 0, dim * N * 4)); CK(cudaMemset(dD, 0, dim * 4)); CK(cudaMemset(dbias, 0, dim * 4));
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
    q.dout = g; q.du = du; q.ddelta = dd; q.dA = dA; q.dB = dB; q.dC = dC; q.dD = dD; q.ddelta_bias = dbias;
    for (int i = 0; i < 3; ++i) CK(fm::launch_scan_bwd_ls_T<float>(q, 0, 1, 1));
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int it = 10;
    cudaEventRecord(e0);
    for (int i = 0; i < it; ++i) CK(fm::launch_scan_bwd_ls_T<float>(q, 0, 1, 1));
    cudaEventRecord(e1); CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("{\"diag\": %d, \"batch\": %d, \"L\": %d, \"us\": %.1f}\n", FM_LS_DIAG, Bn, L, ms / it * 1e3);
    return 0;
}
