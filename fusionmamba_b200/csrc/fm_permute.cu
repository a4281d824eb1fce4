// fm_permute.cu -- stand-alone scan unfold / merge (bit-exact data movement) for sm_100a.
//
// Index maps restated from the reference (not copied):
//   EFFICIENT_V2  EfficientScan.forward / EfficientMerge.forward   models/cross.py:139-169, 34-58
//       Hp = ceil(H/2), Wp = ceil(W/2), Lk = Hp*Wp; sub-grid k = (h&1) | ((w&1)<<1);
//       k in {0,2}: l = (h/2)*Wp + (w/2) (row-major);  k in {1,3}: l = (w/2)*Hp + (h/2) (column-major);
//       positions that fall outside HxW (odd sizes) read 0 on unfold and are dropped on merge.
//   CROSS_V0      classic CrossScan / CrossMerge                     models/cross.py:610-612, 639-642
//       k=0: l = h*W+w; k=1: l = w*H+h; k=2,3: the same reversed (L-1-l); merge is the 4-way sum
//       ((o0 + o2') + o1') + o3' evaluated in that order, each add rounded to the tensor dtype like torch.
// Layout: x / y (batch, dim, H*W); xs / ys (batch, 4, dim, Lk).  One thread per destination element: stores are
// fully coalesced, loads are gathers that hit L2 (each source element is read exactly once).
#include <type_traits>
#include "fm_common.cuh"
#include "fm_launch.h"

namespace fm {

// source pixel (flat h*W+w, or -1 for padding) of element l in direction k
__device__ __forceinline__ int map_pixel(int map, int k, int l, int H, int W) {
    if (map == FM_MAP_CROSS_V0) {
        const int L = H * W;
        if (k >= 2) l = L - 1 - l;
        return (k & 1) ? (l % H) * W + (l / H) : l;
    }
    const int Hp = (H + 1) >> 1, Wp = (W + 1) >> 1;
    int i, j;
    if (k & 1) { j = l / Hp; i = l % Hp; } else { i = l / Wp; j = l % Wp; }
    const int h = 2 * i + (k & 1), w = 2 * j + (k >> 1);
    return (h < H && w < W) ? h * W + w : -1;
}

template <typename T>
__global__ void unfold_kernel(const T* __restrict__ x, T* __restrict__ xs, int map, int batch, int dim, int H, int W, int Lk) {
    const int64_t total = (int64_t)batch * 4 * dim * Lk;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int l = (int)(idx % Lk);
        const int64_t r = idx / Lk;
        const int d = (int)(r % dim);
        const int k = (int)((r / dim) % 4);
        const int b = (int)(r / ((int64_t)4 * dim));
        const int px = map_pixel(map, k, l, H, W);
        xs[idx] = px >= 0 ? x[((int64_t)b * dim + d) * H * W + px] : Cvt<T>::from_f(0.f);
    }
}

template <typename T>
__global__ void merge_kernel(const T* __restrict__ ys, T* __restrict__ y, int map, int batch, int dim, int H, int W, int Lk) {
    const int HW = H * W;
    const int64_t total = (int64_t)batch * dim * HW;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int px = (int)(idx % HW);
        const int64_t r = idx / HW;
        const int d = (int)(r % dim);
        const int b = (int)(r / dim);
        const int h = px / W, w = px % W;
        const T* base = ys + ((int64_t)b * 4 * dim + d) * Lk;
        const int64_t ks = (int64_t)dim * Lk;   // direction stride
        if (map == FM_MAP_EFFICIENT_V2) {
            const int Hp = (H + 1) >> 1, Wp = (W + 1) >> 1;
            const int k = (h & 1) | ((w & 1) << 1);
            const int l = (k & 1) ? (w >> 1) * Hp + (h >> 1) : (h >> 1) * Wp + (w >> 1);
            y[idx] = base[k * ks + l];
        } else {
            const int l0 = px, l1 = w * H + h;
            // y = out_y[:,0] + flip(out_y[:,2]) + wh(out_y[:,1]) + wh(flip(out_y[:,3])), left to right (models/cross.py:642)
            float acc = Cvt<T>::to_f(base[l0]);
            acc = Cvt<T>::to_f(Cvt<T>::from_f(acc + Cvt<T>::to_f(base[2 * ks + (HW - 1 - l0)])));
            acc = Cvt<T>::to_f(Cvt<T>::from_f(acc + Cvt<T>::to_f(base[1 * ks + l1])));
            acc = acc + Cvt<T>::to_f(base[3 * ks + (HW - 1 - l1)]);
            y[idx] = Cvt<T>::from_f(acc);
        }
    }
}

// EFFICIENT_V2, tiled: one CTA moves a 32x32 pixel tile (even origin) of CHP channel images through shared memory so that BOTH
// sides are coalesced -- image rows on the (batch, dim, H, W) side, runs of 16 consecutive l on the (batch, 4, dim, L) side
// (row-major sub-grids k = 0, 2 run along w, column-major sub-grids k = 1, 3 run along h).  kUnfold: x -> xs (zero fill of the
// padded positions of odd sizes); otherwise ys -> y.  The one-thread-per-element kernels above remain for CROSS_V0.
template <typename T, bool kUnfold, int TT>
__global__ void __launch_bounds__(256)
permute_v2_tiled_kernel(const T* __restrict__ src, T* __restrict__ dst, int dim, int H, int W) {
    constexpr int CHP = 4096 / (TT * TT), OP = TT + 2, CHS = TT * OP + 2;   // TT = 8 / 16 / 32 by image size; ~4096 elements per CTA
    __shared__ T tile[CHP * CHS];
    const int tiles_w = (W + TT - 1) / TT;
    const int h0 = (blockIdx.x / tiles_w) * TT, w0 = (blockIdx.x % tiles_w) * TT;
    const int c0 = blockIdx.y * CHP;
    const int b = blockIdx.z;
    const int tid = threadIdx.x;
    const int Hp = (H + 1) >> 1, Wp = (W + 1) >> 1;
    const int64_t L = static_cast<int64_t>(Hp) * Wp, HW = static_cast<int64_t>(H) * W;
    constexpr int HT = TT / 2;
    const T* xsb_c = nullptr;
    T* xsb = nullptr;
    if (kUnfold) xsb = dst + static_cast<int64_t>(b) * 4 * dim * L; else xsb_c = src + static_cast<int64_t>(b) * 4 * dim * L;

    auto image_side = [&](bool load) {   // lanes along w: image rows are contiguous
        for (int e = tid; e < CHP * TT * TT; e += 256) {
            const int col = e % TT, row = (e / TT) % TT, c = e / (TT * TT);
            const int h = h0 + row, w = w0 + col;
            if (c0 + c >= dim) continue;
            const int64_t g = (static_cast<int64_t>(b) * dim + c0 + c) * HW + static_cast<int64_t>(h) * W + w;
            if (load) tile[c * CHS + row * OP + col] = (h < H && w < W) ? src[g] : Cvt<T>::from_f(0.f);
            else if (h < H && w < W) dst[g] = tile[c * CHS + row * OP + col];
        }
    };
    auto scan_side = [&](bool store) {   // consecutive threads walk 16 consecutive l of one sub-grid line
        for (int e = tid; e < 4 * CHP * HT * HT; e += 256) {
            const int ln = e % HT, line = (e / HT) % HT, c = (e / (HT * HT)) % CHP, k = e / (HT * HT * CHP);
            if (c0 + c >= dim) continue;
            int row, col;
            int64_t l;
            if (k & 1) {
                row = 2 * ln + 1; col = 2 * line + (k >> 1);
                const int i = (h0 >> 1) + ln, j = (w0 >> 1) + line;
                if (i >= Hp || j >= Wp) continue;
                l = static_cast<int64_t>(j) * Hp + i;
            } else {
                row = 2 * line; col = 2 * ln + (k >> 1);
                const int i = (h0 >> 1) + line, j = (w0 >> 1) + ln;
                if (i >= Hp || j >= Wp) continue;
                l = static_cast<int64_t>(i) * Wp + j;
            }
            const int64_t g = (static_cast<int64_t>(k) * dim + c0 + c) * L + l;
            if (store) xsb[g] = tile[c * CHS + row * OP + col];
            else tile[c * CHS + row * OP + col] = xsb_c[g];
        }
    };
    if (kUnfold) { image_side(true); __syncthreads(); scan_side(true); }
    else { scan_side(false); __syncthreads(); image_side(false); }
}

// EFFICIENT_V2, tiled + vectorised (default): same tiling, but the shared tile is kept in UNFOLDED order
// [channel][sub-grid k][line][l within line] (line pitch padded by one 16-byte vector), so both global sides move 16-byte
// vectors: VE consecutive pixels of an image row, VS consecutive l of a sub-grid line.  The scalar tiled kernel above is
// element-rate bound at ~2 TB/s (fp32) / ~1 TB/s (bf16) on B200 (profiles/r01_permute_bench.jsonl); ragged widths and
// unaligned bases fall back to element accesses per thread, inside the same kernel.
template <typename T, bool kUnfold, int TT>
__global__ void __launch_bounds__(256)
permute_v2_vec_kernel(const T* __restrict__ src, T* __restrict__ dst, int dim, int H, int W, int vec_img, int vec_scan) {
    constexpr int VE = 16 / sizeof(T);                     // elements per 16-byte vector
    constexpr int HT = TT / 2;
    constexpr int VS = VE < HT ? VE : HT;                  // scan-side vector (HT = 4 with 16-bit data: 8 bytes)
    constexpr int LPT = HT + VE;                           // line pitch: keeps vectors aligned, spreads the transposed writes
    constexpr int PP = HT * LPT;                           // one sub-grid plane
    constexpr int CHS = 4 * PP + VE;                       // channel pitch
    constexpr int CHP = 4096 / (TT * TT);                  // channels per CTA (~4096 elements)
    __shared__ __align__(16) T tile[CHP * CHS];
    const int tiles_w = (W + TT - 1) / TT;
    const int h0 = (blockIdx.x / tiles_w) * TT, w0 = (blockIdx.x % tiles_w) * TT;
    const int c0 = blockIdx.y * CHP;
    const int b = blockIdx.z;
    const int tid = threadIdx.x;
    const int Hp = (H + 1) >> 1, Wp = (W + 1) >> 1;
    const int64_t L = static_cast<int64_t>(Hp) * Wp, HW = static_cast<int64_t>(H) * W;
    const T* img_c = kUnfold ? src + static_cast<int64_t>(b) * dim * HW : nullptr;
    T* img = kUnfold ? nullptr : dst + static_cast<int64_t>(b) * dim * HW;
    const T* seq_c = kUnfold ? nullptr : src + static_cast<int64_t>(b) * 4 * dim * L;
    T* seq = kUnfold ? dst + static_cast<int64_t>(b) * 4 * dim * L : nullptr;

    auto toff = [](int row, int col) {                     // (row, col) of the pixel tile -> offset inside a channel's unfolded tile
        const int k = (row & 1) | ((col & 1) << 1), a = row >> 1, c2 = col >> 1;
        return k * PP + ((k & 1) ? c2 * LPT + a : a * LPT + c2);
    };
    auto image_side = [&]() {                              // lanes along w: image rows are contiguous
        constexpr int VPR = (TT + VE - 1) / VE;            // vectors per tile row
        for (int e = tid; e < CHP * TT * VPR; e += 256) {
            const int col = (e % VPR) * VE, row = (e / VPR) % TT, c = e / (VPR * TT);
            const int h = h0 + row, w = w0 + col;
            if (c0 + c >= dim) continue;
            T* tc = tile + c * CHS;
            const int64_t g = (static_cast<int64_t>(c0 + c)) * HW + static_cast<int64_t>(h) * W + w;
            T v[VE];
            if (kUnfold) {
                if (vec_img && h < H && w < W) {
                    *reinterpret_cast<uint4*>(v) = __ldg(reinterpret_cast<const uint4*>(img_c + g));
                } else {
#pragma unroll
                    for (int j = 0; j < VE; ++j) v[j] = (h < H && w + j < W && col + j < TT) ? img_c[g + j] : Cvt<T>::from_f(0.f);
                }
#pragma unroll
                for (int j = 0; j < VE; ++j)
                    if (col + j < TT) tc[toff(row, col + j)] = v[j];
            } else {
                if (h >= H || w >= W) continue;
#pragma unroll
                for (int j = 0; j < VE; ++j) v[j] = (col + j < TT) ? tc[toff(row, col + j)] : Cvt<T>::from_f(0.f);
                if (vec_img) {
                    *reinterpret_cast<uint4*>(img + g) = *reinterpret_cast<const uint4*>(v);
                } else {
#pragma unroll
                    for (int j = 0; j < VE; ++j)
                        if (w + j < W && col + j < TT) img[g + j] = v[j];
                }
            }
        }
    };
    auto scan_side = [&]() {                               // one thread moves VS consecutive l of one sub-grid line
        constexpr int VPL = HT / VS;
        for (int e = tid; e < 4 * CHP * HT * VPL; e += 256) {
            const int lv = e % VPL, line = (e / VPL) % HT, c = (e / (VPL * HT)) % CHP, k = e / (VPL * HT * CHP);
            if (c0 + c >= dim) continue;
            T* sp = tile + c * CHS + k * PP + line * LPT + lv * VS;
            int64_t l;
            int lim;
            if (k & 1) {   // column-major sub-grid: line = j, elements along i
                const int i = (h0 >> 1) + lv * VS, j = (w0 >> 1) + line;
                if (i >= Hp || j >= Wp) continue;
                l = static_cast<int64_t>(j) * Hp + i;
                lim = Hp - i;
            } else {       // row-major sub-grid: line = i, elements along j
                const int i = (h0 >> 1) + line, j = (w0 >> 1) + lv * VS;
                if (i >= Hp || j >= Wp) continue;
                l = static_cast<int64_t>(i) * Wp + j;
                lim = Wp - j;
            }
            const int64_t g = (static_cast<int64_t>(k) * dim + c0 + c) * L + l;
            using VT = typename std::conditional<VS * sizeof(T) == 16, uint4, uint2>::type;
            if (vec_scan && lim >= VS) {
                if (kUnfold) *reinterpret_cast<VT*>(seq + g) = *reinterpret_cast<const VT*>(sp);
                else *reinterpret_cast<VT*>(sp) = __ldg(reinterpret_cast<const VT*>(seq_c + g));
            } else {
                for (int q = 0; q < VS && q < lim; ++q) {
                    if (kUnfold) seq[g + q] = sp[q]; else sp[q] = seq_c[g + q];
                }
            }
        }
    };
    if (kUnfold) { image_side(); __syncthreads(); scan_side(); }
    else { scan_side(); __syncthreads(); image_side(); }
}

// CROSS_V0, tiled + vectorised: one CTA moves a TTxTT pixel tile of CHP channel images.  Directions 0 / 2 (row-major and
// its reversal) never touch shared memory: the thread that holds VE consecutive pixels of an image row reads / writes the
// same run of l (reversed element order and address for direction 2).  Directions 1 / 3 (column-major) go through a
// [row][col] tile with an odd pitch: the image side scatters / gathers scalars, the scan side moves VE consecutive h of
// one image column as a vector (conflict-free: lanes differ in 4*hvec + w banks).  The merge keeps the reference's add order
// ((o0 + o2') + o1') + o3' with a rounding to T after each of the first two adds (models/cross.py:642), so it stays bit-exact.
// Vector paths need W % VE == 0 (row side) and H % VE == 0 (column side) and 16-byte aligned bases; otherwise the same kernel
// runs element accesses.
template <typename T, bool kUnfold, int TT>
__global__ void __launch_bounds__(256)
permute_v0_tiled_kernel(const T* __restrict__ src, T* __restrict__ dst, int dim, int H, int W, int vec_row, int vec_col) {
    constexpr int VE = 16 / sizeof(T);
    constexpr int PT = TT + (sizeof(T) == 4 ? 1 : 2);      // pitch: an odd number of 32-bit words
    constexpr int CHP = 4096 / (TT * TT);
    __shared__ T t1[CHP * TT * PT];
    __shared__ T t3[kUnfold ? 1 : CHP * TT * PT];
    const int tiles_w = (W + TT - 1) / TT;
    const int h0 = (blockIdx.x / tiles_w) * TT, w0 = (blockIdx.x % tiles_w) * TT;
    const int c0 = blockIdx.y * CHP;
    const int b = blockIdx.z;
    const int tid = threadIdx.x;
    const int64_t L = static_cast<int64_t>(H) * W;
    const T* img_c = kUnfold ? src + static_cast<int64_t>(b) * dim * L : nullptr;
    T* img = kUnfold ? nullptr : dst + static_cast<int64_t>(b) * dim * L;
    const T* seq_c = kUnfold ? nullptr : src + static_cast<int64_t>(b) * 4 * dim * L;
    T* seq = kUnfold ? dst + static_cast<int64_t>(b) * 4 * dim * L : nullptr;
    const int64_t ks = static_cast<int64_t>(dim) * L;      // direction stride
    constexpr int VPR = TT / VE > 0 ? TT / VE : 1;         // vectors per tile row / column
    constexpr int VW = TT < VE ? TT : VE;                  // elements per thread item

    auto load_vec = [](const T* p, bool vec, int n, T (&v)[VW]) {       // n valid elements (1..VW)
        if (vec && n == VW && VW == VE) { *reinterpret_cast<uint4*>(v) = __ldg(reinterpret_cast<const uint4*>(p)); return; }
#pragma unroll
        for (int j = 0; j < VW; ++j) v[j] = j < n ? p[j] : Cvt<T>::from_f(0.f);
    };
    auto store_vec = [](T* p, bool vec, int n, const T (&v)[VW]) {
        if (vec && n == VW && VW == VE) { *reinterpret_cast<uint4*>(p) = *reinterpret_cast<const uint4*>(v); return; }
#pragma unroll
        for (int j = 0; j < VW; ++j) if (j < n) p[j] = v[j];
    };
    // reversed run: elements v[0..n) belong to l, l+1, ... and go to / come from rl = L-1-l, rl-1, ...  (ascending address
    // rl-n+1 holds v[n-1]); with n == VW the vector is written at rl-VW+1 in reversed element order
    auto load_rev = [&](const T* base, int64_t rl, bool vec, int n, T (&v)[VW]) {
        if (vec && n == VW && VW == VE) {
            T r[VW];
            *reinterpret_cast<uint4*>(r) = __ldg(reinterpret_cast<const uint4*>(base + rl - (VW - 1)));
#pragma unroll
            for (int j = 0; j < VW; ++j) v[j] = r[VW - 1 - j];
            return;
        }
#pragma unroll
        for (int j = 0; j < VW; ++j) v[j] = j < n ? base[rl - j] : Cvt<T>::from_f(0.f);
    };
    auto store_rev = [&](T* base, int64_t rl, bool vec, int n, const T (&v)[VW]) {
        if (vec && n == VW && VW == VE) {
            T r[VW];
#pragma unroll
            for (int j = 0; j < VW; ++j) r[j] = v[VW - 1 - j];
            *reinterpret_cast<uint4*>(base + rl - (VW - 1)) = *reinterpret_cast<const uint4*>(r);
            return;
        }
#pragma unroll
        for (int j = 0; j < VW; ++j) if (j < n) base[rl - j] = v[j];
    };

    if (kUnfold) {
        // image rows -> directions 0 and 2 straight from registers, and into the tile for the column-major directions
        for (int e = tid; e < CHP * TT * VPR; e += 256) {
            const int col = (e % VPR) * VW, row = (e / VPR) % TT, c = e / (VPR * TT);
            const int h = h0 + row, w = w0 + col;
            if (c0 + c >= dim || h >= H || w >= W) continue;
            const int n = (W - w) < VW ? (W - w) : VW;
            const int64_t l0 = static_cast<int64_t>(h) * W + w;
            const int64_t ch = static_cast<int64_t>(c0 + c) * L;
            T v[VW];
            load_vec(img_c + ch + l0, vec_row, n, v);
            store_vec(seq + ch + l0, vec_row, n, v);
            store_rev(seq + 2 * ks + ch, L - 1 - l0, vec_row, n, v);
#pragma unroll
            for (int j = 0; j < VW; ++j) if (j < n) t1[(c * TT + row) * PT + col + j] = v[j];
        }
        __syncthreads();
        // tile columns -> directions 1 and 3: VW consecutive h of one image column
        for (int e = tid; e < CHP * TT * VPR; e += 256) {
            const int row = (e % VPR) * VW, col = (e / VPR) % TT, c = e / (VPR * TT);
            const int h = h0 + row, w = w0 + col;
            if (c0 + c >= dim || h >= H || w >= W) continue;
            const int n = (H - h) < VW ? (H - h) : VW;
            const int64_t l1 = static_cast<int64_t>(w) * H + h;
            const int64_t ch = static_cast<int64_t>(c0 + c) * L;
            T v[VW];
#pragma unroll
            for (int j = 0; j < VW; ++j) v[j] = j < n ? t1[(c * TT + row + j) * PT + col] : Cvt<T>::from_f(0.f);
            store_vec(seq + ks + ch + l1, vec_col, n, v);
            store_rev(seq + 3 * ks + ch, L - 1 - l1, vec_col, n, v);
        }
    } else {
        // column-major directions into the tiles
        for (int e = tid; e < CHP * TT * VPR; e += 256) {
            const int row = (e % VPR) * VW, col = (e / VPR) % TT, c = e / (VPR * TT);
            const int h = h0 + row, w = w0 + col;
            if (c0 + c >= dim || h >= H || w >= W) continue;
            const int n = (H - h) < VW ? (H - h) : VW;
            const int64_t l1 = static_cast<int64_t>(w) * H + h;
            const int64_t ch = static_cast<int64_t>(c0 + c) * L;
            T a[VW], r[VW];
            load_vec(seq_c + ks + ch + l1, vec_col, n, a);
            load_rev(seq_c + 3 * ks + ch, L - 1 - l1, vec_col, n, r);
#pragma unroll
            for (int j = 0; j < VW; ++j)
                if (j < n) { t1[(c * TT + row + j) * PT + col] = a[j]; t3[(c * TT + row + j) * PT + col] = r[j]; }
        }
        __syncthreads();
        for (int e = tid; e < CHP * TT * VPR; e += 256) {
            const int col = (e % VPR) * VW, row = (e / VPR) % TT, c = e / (VPR * TT);
            const int h = h0 + row, w = w0 + col;
            if (c0 + c >= dim || h >= H || w >= W) continue;
            const int n = (W - w) < VW ? (W - w) : VW;
            const int64_t l0 = static_cast<int64_t>(h) * W + w;
            const int64_t ch = static_cast<int64_t>(c0 + c) * L;
            T a[VW], r[VW], o[VW];
            load_vec(seq_c + ch + l0, vec_row, n, a);
            load_rev(seq_c + 2 * ks + ch, L - 1 - l0, vec_row, n, r);
#pragma unroll
            for (int j = 0; j < VW; ++j) {
                float acc = Cvt<T>::to_f(a[j]);
                acc = Cvt<T>::to_f(Cvt<T>::from_f(acc + Cvt<T>::to_f(r[j])));
                acc = Cvt<T>::to_f(Cvt<T>::from_f(acc + Cvt<T>::to_f(t1[(c * TT + row) * PT + col + j])));
                acc = acc + Cvt<T>::to_f(t3[(c * TT + row) * PT + col + j]);
                o[j] = Cvt<T>::from_f(acc);
            }
            store_vec(img + ch + l0, vec_row, n, o);
        }
    }
}

static int seq_len(const FmPermuteParams& p) {
    return p.map == FM_MAP_CROSS_V0 ? p.h * p.w : ((p.h + 1) / 2) * ((p.w + 1) / 2);
}

template <typename T>
static cudaError_t launch_perm(const FmPermuteParams& p, cudaStream_t st, bool unfold) {
    // the tiled kernels index batch and channel tiles through gridDim.z / .y; beyond those limits the grid-stride
    // one-thread-per-element kernels at the bottom take over
    const bool tiled_ok = p.batch <= 65535 && p.dim <= 65535;
    if (tiled_ok && p.map == FM_MAP_EFFICIENT_V2) {
        const int m = p.h > p.w ? p.h : p.w;
        if (env_int("FM_PERMUTE_VEC", 1)) {
            constexpr int VE = 16 / (int)sizeof(T);
            const int Hp = (p.h + 1) / 2, Wp = (p.w + 1) / 2;
            const void* img = unfold ? p.src : p.dst;
            const void* sq = unfold ? p.dst : p.src;
            const int vec_img = (p.w % VE == 0) && aligned16(img);
#define FM_PERM_VEC(tt)                                                                                                      \
    {                                                                                                                        \
        constexpr int chp = 4096 / (tt * tt);                                                                                \
        constexpr int vs = VE < tt / 2 ? VE : tt / 2;                                                                        \
        const int vec_scan = (Hp % vs == 0) && (Wp % vs == 0) && aligned16(sq);                                              \
        dim3 grid(((p.h + tt - 1) / tt) * ((p.w + tt - 1) / tt), (p.dim + chp - 1) / chp, p.batch);                         \
        if (unfold) permute_v2_vec_kernel<T, true, tt><<<grid, 256, 0, st>>>(static_cast<const T*>(p.src), static_cast<T*>(p.dst), p.dim, p.h, p.w, vec_img, vec_scan); \
        else permute_v2_vec_kernel<T, false, tt><<<grid, 256, 0, st>>>(static_cast<const T*>(p.src), static_cast<T*>(p.dst), p.dim, p.h, p.w, vec_img, vec_scan);       \
    }
            if (m <= 8) FM_PERM_VEC(8)
            else if (m <= 16) FM_PERM_VEC(16)
            else FM_PERM_VEC(32)
#undef FM_PERM_VEC
            count_launch();
            return cudaGetLastError();
        }
#define FM_PERM_TILED(tt)                                                                                                    \
    {                                                                                                                        \
        constexpr int chp = 4096 / (tt * tt);                                                                                \
        dim3 grid(((p.h + tt - 1) / tt) * ((p.w + tt - 1) / tt), (p.dim + chp - 1) / chp, p.batch);                         \
        if (unfold) permute_v2_tiled_kernel<T, true, tt><<<grid, 256, 0, st>>>(static_cast<const T*>(p.src), static_cast<T*>(p.dst), p.dim, p.h, p.w); \
        else permute_v2_tiled_kernel<T, false, tt><<<grid, 256, 0, st>>>(static_cast<const T*>(p.src), static_cast<T*>(p.dst), p.dim, p.h, p.w);       \
    }
        if (m <= 8) FM_PERM_TILED(8)
        else if (m <= 16) FM_PERM_TILED(16)
        else FM_PERM_TILED(32)
#undef FM_PERM_TILED
        count_launch();
        return cudaGetLastError();
    }
    if (tiled_ok && p.map == FM_MAP_CROSS_V0 && env_int("FM_PERMUTE_VEC", 1)) {
        constexpr int VE = 16 / (int)sizeof(T);
        const int m = p.h > p.w ? p.h : p.w;
        const int vec_row = (p.w % VE == 0) && aligned16(p.src) && aligned16(p.dst);
        const int vec_col = (p.h % VE == 0) && aligned16(p.src) && aligned16(p.dst);
#define FM_PERM_V0(tt)                                                                                                       \
    {                                                                                                                        \
        constexpr int chp = 4096 / (tt * tt);                                                    \
        dim3 grid(((p.h + tt - 1) / tt) * ((p.w + tt - 1) / tt), (p.dim + chp - 1) / chp, p.batch);                         \
        if (unfold) permute_v0_tiled_kernel<T, true, tt><<<grid, 256, 0, st>>>(static_cast<const T*>(p.src), static_cast<T*>(p.dst), p.dim, p.h, p.w, vec_row, vec_col); \
        else permute_v0_tiled_kernel<T, false, tt><<<grid, 256, 0, st>>>(static_cast<const T*>(p.src), static_cast<T*>(p.dst), p.dim, p.h, p.w, vec_row, vec_col);       \
    }
        if (m <= 8) FM_PERM_V0(8)
        else if (m <= 16) FM_PERM_V0(16)
        else FM_PERM_V0(32)
#undef FM_PERM_V0
        count_launch();
        return cudaGetLastError();
    }
    const int Lk = seq_len(p);
    const int64_t total = unfold ? (int64_t)p.batch * 4 * p.dim * Lk : (int64_t)p.batch * p.dim * p.h * p.w;
    const int threads = 256;
    int64_t blocks = (total + threads - 1) / threads;
    const int64_t cap = 148LL * 32;
    if (blocks > cap) blocks = cap;
    if (unfold)
        unfold_kernel<T><<<(unsigned)blocks, threads, 0, st>>>(static_cast<const T*>(p.src), static_cast<T*>(p.dst), p.map,
                                                                p.batch, p.dim, p.h, p.w, Lk);
    else
        merge_kernel<T><<<(unsigned)blocks, threads, 0, st>>>(static_cast<const T*>(p.src), static_cast<T*>(p.dst), p.map,
                                                               p.batch, p.dim, p.h, p.w, Lk);
    count_launch();
    return cudaGetLastError();
}

static cudaError_t dispatch(const FmPermuteParams& p, cudaStream_t st, bool unfold) {
    switch (p.dtype) {
        case FM_F32: return launch_perm<float>(p, st, unfold);
        case FM_F16: return launch_perm<__half>(p, st, unfold);
        default: return launch_perm<__nv_bfloat16>(p, st, unfold);
    }
}

cudaError_t launch_unfold(const FmPermuteParams& p, cudaStream_t st) { return dispatch(p, st, true); }
cudaError_t launch_merge(const FmPermuteParams& p, cudaStream_t st) { return dispatch(p, st, false); }

}  // namespace fm
