/*
 * scan_oracle.c -- plain-C CPU restatement of the selective scan (TEST INFRASTRUCTURE ONLY).
 *
 * Checker / CPU baseline for the CUDA path; never linked into or called from fusionmamba_b200/.
 * Follows the reference's pure-PyTorch definition and its autograd, not the CUDA kernels:
 *   forward   selective_scan_ref                 mamba_ssm/ops/selective_scan_interface.py:92-158
 *   backward  closed form of that function's autograd (the reference kernel states
 *             the same products at selective_scan/selective_scan_bwd_kernel.cuh:171-207, 277-296, 439-452)
 * All inputs are contiguous fp32 (half-precision test inputs are widened exactly by the caller); every
 * accumulation is done in double.  Rows (batch, channel) are independent -> OpenMP over (batch, group).
 * Pinned against the reference-generated fixtures in tests/golden by tests/test_oracle_golden.py.
 *
 * Build: see oracle/Makefile  (gcc -O2 -fopenmp -shared -fPIC).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

static inline double softplus_d(double x) { return x > 20.0 ? x : log1p(exp(x)); }  /* F.softplus, threshold 20 */
static inline double sigmoid_d(double x) { return 1.0 / (1.0 + exp(-x)); }

/* Launchers such as torch.distributed.run export OMP_NUM_THREADS=1; the timed CPU baseline asks for the host's cores. */
void fm_oracle_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int fm_oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* u, delta, z: (batch, dim, L); A: (dim, N); B, C: (batch, G, N, L); D, bias: (dim) or NULL.
 * out: (batch, dim, L) double = y (z == NULL) or y*silu(z); y_pre (optional): pre-gate y;
 * last_state (optional): (batch, dim, N) double. */
int fm_oracle_scan_fwd(int batch, int dim, int L, int N, int G,
                       const float *u, const float *delta, const float *A, const float *B, const float *C,
                       const float *D, const float *z, const float *bias, int delta_softplus,
                       double *out, double *y_pre, double *last_state) {
    if (batch <= 0 || dim <= 0 || L <= 0 || N <= 0 || G <= 0 || dim % G) return 1;
    const int H = dim / G;
    const int64_t rows = (int64_t)batch * dim;
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t r = 0; r < rows; ++r) {
        const int b = (int)(r / dim), d = (int)(r % dim), g = d / H;
        const float *ur = u + r * L, *dr = delta + r * L;
        const float *Bg = B + ((int64_t)b * G + g) * N * L, *Cg = C + ((int64_t)b * G + g) * N * L;
        double h[256];
        for (int n = 0; n < N; ++n) h[n] = 0.0;
        const double bi = bias ? bias[d] : 0.0, Dv = D ? D[d] : 0.0;
        for (int t = 0; t < L; ++t) {
            double x = (double)dr[t] + bi;                                   /* :111 */
            double dt = delta_softplus ? softplus_d(x) : x;                  /* :113 */
            double dtu = dt * ur[t];
            double y = 0.0;
            for (int n = 0; n < N; ++n) {
                double a = exp(dt * A[d * N + n]);                           /* :127 */
                h[n] = a * h[n] + dtu * Bg[(int64_t)n * L + t];              /* :135, :140 */
                y += h[n] * Cg[(int64_t)n * L + t];                          /* :147 */
            }
            y += Dv * ur[t];                                                 /* :154 */
            if (y_pre) y_pre[r * L + t] = y;
            if (z) { double zz = z[r * L + t]; y = y * zz * sigmoid_d(zz); } /* :155-156 */
            out[r * L + t] = y;
        }
        if (last_state) for (int n = 0; n < N; ++n) last_state[r * N + n] = h[n];   /* :148-149 */
    }
    return 0;
}

/* Gradients for upstream gradient dout (batch, dim, L) fp32.  Outputs (double, caller-allocated):
 * du, ddelta (batch, dim, L); dA (dim, N); dB, dC (batch, G, N, L); dD, dbias (dim) or NULL; dz or NULL. */
int fm_oracle_scan_bwd(int batch, int dim, int L, int N, int G,
                       const float *u, const float *delta, const float *A, const float *B, const float *C,
                       const float *D, const float *z, const float *bias, int delta_softplus, const float *dout,
                       double *du, double *ddelta, double *dA, double *dB, double *dC,
                       double *dD, double *dbias, double *dz) {
    if (batch <= 0 || dim <= 0 || L <= 0 || N <= 0 || G <= 0 || dim % G) return 1;
    const int H = dim / G;
    memset(dA, 0, sizeof(double) * (size_t)dim * N);
    memset(dB, 0, sizeof(double) * (size_t)batch * G * N * L);
    memset(dC, 0, sizeof(double) * (size_t)batch * G * N * L);
    if (dD) memset(dD, 0, sizeof(double) * (size_t)dim);
    if (dbias) memset(dbias, 0, sizeof(double) * (size_t)dim);
    int fail = 0;
    /* one task per (batch, group): its dB/dC tile is private to the task; dA/dD/dbias need atomics over batch */
#pragma omp parallel for schedule(dynamic, 1) collapse(2)
    for (int b = 0; b < batch; ++b) {
        for (int g = 0; g < G; ++g) {
            double *hs = (double *)malloc(sizeof(double) * (size_t)L * N);   /* h_t */
            double *as = (double *)malloc(sizeof(double) * (size_t)L * N);   /* a_t */
            double *dts = (double *)malloc(sizeof(double) * (size_t)L * 3);  /* dt, x, dy */
            if (!hs || !as || !dts) { fail = 1; free(hs); free(as); free(dts); continue; }
            const float *Bg = B + ((int64_t)b * G + g) * N * L, *Cg = C + ((int64_t)b * G + g) * N * L;
            double *dBg = dB + ((int64_t)b * G + g) * N * L, *dCg = dC + ((int64_t)b * G + g) * N * L;
            for (int dd = 0; dd < H; ++dd) {
                const int d = g * H + dd;
                const int64_t r = (int64_t)b * dim + d;
                const float *ur = u + r * L, *dr = delta + r * L, *gr = dout + r * L;
                const double bi = bias ? bias[d] : 0.0, Dv = D ? D[d] : 0.0;
                double h[256], dh[256], dA_loc[256];
                for (int n = 0; n < N; ++n) { h[n] = 0.0; dh[n] = 0.0; dA_loc[n] = 0.0; }
                double dD_loc = 0.0, dbias_loc = 0.0;
                for (int t = 0; t < L; ++t) {
                    double x = (double)dr[t] + bi;
                    double dt = delta_softplus ? softplus_d(x) : x;
                    double dtu = dt * ur[t], y = 0.0;
                    for (int n = 0; n < N; ++n) {
                        double a = exp(dt * A[d * N + n]);
                        h[n] = a * h[n] + dtu * Bg[(int64_t)n * L + t];
                        hs[(size_t)t * N + n] = h[n];
                        as[(size_t)t * N + n] = a;
                        y += h[n] * Cg[(int64_t)n * L + t];
                    }
                    y += Dv * ur[t];
                    double gy = gr[t];
                    if (z) {
                        double zz = z[r * L + t], sg = sigmoid_d(zz);
                        dz[r * L + t] = gy * y * sg * (1.0 + zz * (1.0 - sg));
                        gy = gy * zz * sg;
                    }
                    dts[3 * t] = dt; dts[3 * t + 1] = x; dts[3 * t + 2] = gy;
                }
                for (int t = L - 1; t >= 0; --t) {
                    const double dt = dts[3 * t], x = dts[3 * t + 1], dy = dts[3 * t + 2];
                    double s1 = 0.0, sw = 0.0;
                    for (int n = 0; n < N; ++n) {
                        const double an = (t + 1 < L) ? as[(size_t)(t + 1) * N + n] : 0.0;
                        dh[n] = Cg[(int64_t)n * L + t] * dy + an * dh[n];
                        const double hp = t > 0 ? hs[(size_t)(t - 1) * N + n] : 0.0;
                        const double w = dh[n] * as[(size_t)t * N + n] * hp;        /* dh * (h_t - b_t) */
                        s1 += dh[n] * Bg[(int64_t)n * L + t];
                        sw += w * A[d * N + n];
                        dA_loc[n] += w * dt;
                        dBg[(int64_t)n * L + t] += dh[n] * dt * ur[t];
                        dCg[(int64_t)n * L + t] += dy * hs[(size_t)t * N + n];
                    }
                    du[r * L + t] = dt * s1 + Dv * dy;
                    double ddt = ur[t] * s1 + sw;
                    double dd_ = (delta_softplus && x <= 20.0) ? ddt * sigmoid_d(x) : ddt;
                    ddelta[r * L + t] = dd_;
                    dbias_loc += dd_;
                    dD_loc += dy * ur[t];
                }
                for (int n = 0; n < N; ++n) {
#pragma omp atomic
                    dA[d * N + n] += dA_loc[n];
                }
                if (dD) {
#pragma omp atomic
                    dD[d] += dD_loc;
                }
                if (dbias) {
#pragma omp atomic
                    dbias[d] += dbias_loc;
                }
            }
            free(hs); free(as); free(dts);
        }
    }
    return fail;
}
