// fm_api.cu -- the extern "C" boundary of libfm_scan.so (see include/fm_scan.h).
// Validation mirrors the reference's TORCH_CHECKs (selective_scan/selective_scan.cpp:233-294, 350-456);
// allocation stays with the caller (the Python shim), as ownership rules in SURVEY.md section 8b require.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "fm_launch.h"

namespace fm {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int env_int(const char* name, int dflt) {
    const char* v = std::getenv(name);
    if (!v || !*v) return dflt;
    return std::atoi(v);
}

int scan_lanes_per_row(int64_t rows, int seqlen, int seg_len, const char* env_name) {
    int forced = env_int(env_name, 0);
    int g;
    if (forced > 0) {
        g = forced;
    } else {
        // target ~ 148 SMs x 16 warps of lanes
        const int64_t want = 148LL * 16 * 32;
        g = 1;
        while (g < 32 && rows * g < want) g <<= 1;
    }
    int gmax = 1;
    while (gmax < 32 && gmax * seg_len < seqlen) gmax <<= 1;
    if (g > gmax) g = gmax;
    if (g < 1) g = 1;
    int p2 = 1;
    while (p2 * 2 <= g) p2 <<= 1;
    return p2;
}

static int fail(FmStatus s, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return (int)s;
}

static int check_fwd(const FmScanFwdParams& p, const char* who) {
    if (p.abi_version != FM_SCAN_ABI_VERSION)
        return fail(FM_ERR_INVALID_ARG, "%s: abi_version %d != %d", who, p.abi_version, FM_SCAN_ABI_VERSION);
    if (p.dtype != FM_F32 && p.dtype != FM_F16 && p.dtype != FM_BF16)
        return fail(FM_ERR_INVALID_ARG, "%s: dtype must be fp32, fp16 or bf16", who);
    if (p.reserved0 != 0) return fail(FM_ERR_INVALID_ARG, "%s: reserved0 must be 0", who);
    if (p.out_dtype != p.dtype) {
        const bool fwd_call = std::strcmp(who, "fm_selective_scan_fwd") == 0;
        if (!(fwd_call && p.out_dtype == FM_F32 && p.dtype != FM_F32 && !p.z && p.dstate == 16))
            return fail(FM_ERR_UNSUPPORTED, "%s: out_dtype may differ from dtype only as fp32 output of a 16-bit forward "
                                             "without z at dstate 16", who);
    }
    if (p.batch <= 0 || p.dim <= 0 || p.seqlen <= 0 || p.dstate <= 0 || p.n_groups <= 0)
        return fail(FM_ERR_INVALID_ARG, "%s: batch/dim/seqlen/dstate/n_groups must be positive (got %d %d %d %d %d)", who,
                    p.batch, p.dim, p.seqlen, p.dstate, p.n_groups);
    if (p.dstate > 256) return fail(FM_ERR_INVALID_ARG, "selective_scan only supports state dimension <= 256");
    if (p.dim % p.n_groups != 0)
        return fail(FM_ERR_INVALID_ARG, "%s: dim (%d) must be a multiple of n_groups (%d)", who, p.dim, p.n_groups);
    if (p.batch > 65535) return fail(FM_ERR_INVALID_ARG, "%s: batch > 65535 not supported", who);
    if (p.chunk_len <= 0 || p.n_chunks != (p.seqlen + p.chunk_len - 1) / p.chunk_len)
        return fail(FM_ERR_INVALID_ARG, "%s: n_chunks must equal ceil(seqlen / chunk_len)", who);
    if (p.chunk_len % 512 != 0)
        return fail(FM_ERR_INVALID_ARG, "%s: chunk_len must be a multiple of 512", who);
    const bool is_fwd = std::strcmp(who, "fm_selective_scan_fwd") == 0;
    if (!p.u || !p.delta || !p.A || !p.B || !p.C || (is_fwd && (!p.out || !p.x)))
        return fail(FM_ERR_INVALID_ARG, "%s: u, delta, A, B, C (and out, x for fwd) must be non-null device pointers", who);
    if (p.hck) {
        if (!(p.hck_len == 8 && p.dstate == 16) && (p.hck_len <= 0 || p.hck_len % 16 != 0 || 512 % p.hck_len != 0))
            return fail(FM_ERR_INVALID_ARG, "%s: hck_len must be 16, 32, 64, 128, 256 or 512 (or 8 with dstate == 16)", who);
        if (p.n_hck != (p.seqlen + p.hck_len - 1) / p.hck_len - 1)
            return fail(FM_ERR_INVALID_ARG, "%s: n_hck must equal ceil(seqlen / hck_len) - 1", who);
    }
    if (p.z && !p.out_z && std::strcmp(who, "fm_selective_scan_fwd") == 0)
        return fail(FM_ERR_INVALID_ARG, "%s: out_z is required when z is given", who);
    if (p.u_map != FM_MAP_LINEAR || p.out_map != FM_MAP_LINEAR) {
        if (p.u_map < 0 || p.u_map > FM_MAP_EFFICIENT_V2 || p.out_map < 0 || p.out_map > FM_MAP_EFFICIENT_V2_CL)
            return fail(FM_ERR_INVALID_ARG, "%s: unknown index map", who);
        // Fused merge-on-store: the forward writes y (batch, dim/4, H*W) directly (EfficientMerge, a pure permutation).
        // The interface has no unfold-on-load (u_map is LINEAR by contract, include/fm_scan.h) and no V0 merge-on-store (a 4-way
        // sum across scans); fm_scan_unfold / fm_scan_merge / fm_conv_unfold serve those.
        if (!is_fwd || p.u_map != FM_MAP_LINEAR || (p.out_map != FM_MAP_EFFICIENT_V2 && p.out_map != FM_MAP_EFFICIENT_V2_CL))
            return fail(FM_ERR_INVALID_ARG, "%s: u_map must be LINEAR; out_map must be LINEAR, EFFICIENT_V2 or EFFICIENT_V2_CL (forward only)", who);
        if (p.n_groups != 4 || p.dstate != 16 || p.z || p.hck)
            return fail(FM_ERR_UNSUPPORTED, "%s: fused merge needs n_groups == 4, dstate == 16, no z and no checkpoints", who);
        if (p.map_h <= 0 || p.map_w <= 0 || p.seqlen != ((p.map_h + 1) / 2) * ((p.map_w + 1) / 2))
            return fail(FM_ERR_INVALID_ARG, "%s: seqlen must equal ceil(H/2)*ceil(W/2) for the fused merge", who);
    }
    return FM_OK;
}

}  // namespace fm

using namespace fm;

extern "C" {

int fm_selective_scan_fwd(const FmScanFwdParams* params, void* stream) {
    if (!params) return fail(FM_ERR_INVALID_ARG, "fm_selective_scan_fwd: params is null");
    int rc = check_fwd(*params, "fm_selective_scan_fwd");
    if (rc) return rc;
    cudaError_t e = launch_scan_fwd(*params, static_cast<cudaStream_t>(stream));
    if (e == cudaErrorInvalidConfiguration)
        return fail(FM_ERR_UNSUPPORTED, "fm_selective_scan_fwd: no kernel configuration fits this shape (batch %d, dim %d, seqlen %d, "
                                        "dstate %d, groups %d, hck_len %d)", params->batch, params->dim, params->seqlen, params->dstate,
                    params->n_groups, params->hck_len);
    if (e != cudaSuccess) return fail(FM_ERR_CUDA, "fm_selective_scan_fwd: %s", cudaGetErrorString(e));
    return FM_OK;
}

int64_t fm_scan_fwd_workspace_bytes(const FmScanFwdParams* params) {
    if (!params || params->abi_version != FM_SCAN_ABI_VERSION || params->n_groups <= 0 || params->dim <= 0) return 0;
    return fwd16_split_plan(*params).ws_bytes;
}

int fm_selective_scan_bwd(const FmScanBwdParams* params, void* stream) {
    if (!params) return fail(FM_ERR_INVALID_ARG, "fm_selective_scan_bwd: params is null");
    const FmScanBwdParams& p = *params;
    int rc = check_fwd(p.f, "fm_selective_scan_bwd");
    if (rc) return rc;
    if (!p.dout || !p.du || !p.ddelta || !p.dA || !p.dB || !p.dC)
        return fail(FM_ERR_INVALID_ARG, "fm_selective_scan_bwd: dout, du, ddelta, dA, dB, dC must be non-null");
    if (!p.f.hck && p.f.seqlen > 256)
        return fail(FM_ERR_INVALID_ARG, "fm_selective_scan_bwd: hck (dense state checkpoints written by the forward) is "
                                         "required when seqlen > 256");
    if (p.f.z && !p.f.out) return fail(FM_ERR_INVALID_ARG, "fm_selective_scan_bwd: out (saved y) is required when z is given");
    if (p.f.z && !p.dz) return fail(FM_ERR_INVALID_ARG, "fm_selective_scan_bwd: dz is required when z is given");
    if ((p.f.D != nullptr) != (p.dD != nullptr))
        return fail(FM_ERR_INVALID_ARG, "fm_selective_scan_bwd: dD must be given iff D is given");
    if ((p.f.delta_bias != nullptr) != (p.ddelta_bias != nullptr))
        return fail(FM_ERR_INVALID_ARG, "fm_selective_scan_bwd: ddelta_bias must be given iff delta_bias is given");
    cudaError_t e = launch_scan_bwd(p, static_cast<cudaStream_t>(stream));
    if (e == cudaErrorInvalidConfiguration)
        return fail(FM_ERR_UNSUPPORTED, "fm_selective_scan_bwd: no kernel configuration fits this shape (batch %d, dim %d, seqlen %d, "
                                        "dstate %d, groups %d, hck_len %d)", p.f.batch, p.f.dim, p.f.seqlen, p.f.dstate, p.f.n_groups, p.f.hck_len);
    if (e != cudaSuccess) return fail(FM_ERR_CUDA, "fm_selective_scan_bwd: %s", cudaGetErrorString(e));
    return FM_OK;
}

static int check_perm(const FmPermuteParams* p, const char* who) {
    if (!p) return fail(FM_ERR_INVALID_ARG, "%s: params is null", who);
    if (p->abi_version != FM_SCAN_ABI_VERSION) return fail(FM_ERR_INVALID_ARG, "%s: abi_version mismatch", who);
    if (p->dtype != FM_F32 && p->dtype != FM_F16 && p->dtype != FM_BF16)
        return fail(FM_ERR_INVALID_ARG, "%s: dtype must be fp32, fp16 or bf16", who);
    if (p->map != FM_MAP_CROSS_V0 && p->map != FM_MAP_EFFICIENT_V2)
        return fail(FM_ERR_INVALID_ARG, "%s: map must be CROSS_V0 or EFFICIENT_V2", who);
    if (p->batch <= 0 || p->dim <= 0 || p->h <= 0 || p->w <= 0 || !p->src || !p->dst)
        return fail(FM_ERR_INVALID_ARG, "%s: bad shape or null pointer", who);
    return FM_OK;
}

int fm_scan_unfold(const FmPermuteParams* params, void* stream) {
    int rc = check_perm(params, "fm_scan_unfold");
    if (rc) return rc;
    cudaError_t e = launch_unfold(*params, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return fail(FM_ERR_CUDA, "fm_scan_unfold: %s", cudaGetErrorString(e));
    return FM_OK;
}

int fm_scan_merge(const FmPermuteParams* params, void* stream) {
    int rc = check_perm(params, "fm_scan_merge");
    if (rc) return rc;
    cudaError_t e = launch_merge(*params, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return fail(FM_ERR_CUDA, "fm_scan_merge: %s", cudaGetErrorString(e));
    return FM_OK;
}

int fm_merge_norm(const FmNormParams* p, void* stream) {
    if (!p) return fail(FM_ERR_INVALID_ARG, "fm_merge_norm: params is null");
    if (p->abi_version != FM_SCAN_ABI_VERSION) return fail(FM_ERR_INVALID_ARG, "fm_merge_norm: abi_version mismatch");
    if (p->out_dtype != FM_F32 && p->out_dtype != FM_F16 && p->out_dtype != FM_BF16)
        return fail(FM_ERR_INVALID_ARG, "fm_merge_norm: out_dtype must be fp32, fp16 or bf16");
    if (p->batch <= 0 || p->batch > 65535 || p->dim <= 0 || p->positions <= 0 || !p->src || !p->dst || !(p->eps >= 0.f))
        return fail(FM_ERR_INVALID_ARG, "fm_merge_norm: bad shape, eps or null pointer");
    if ((p->src_channels_last != 0 && p->src_channels_last != 1) || (p->gate && (p->gate_channel_offset < 0 || p->gate_channel_stride < p->gate_channel_offset + p->dim)))
        return fail(FM_ERR_INVALID_ARG, "fm_merge_norm: bad gate stride / offset");
    cudaError_t e = launch_merge_norm(*p, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return fail(FM_ERR_CUDA, "fm_merge_norm: %s", cudaGetErrorString(e));
    return FM_OK;
}

int64_t fm_layer_norm_bwd_workspace_bytes(int32_t dim, int64_t rows) {
    if (dim <= 0 || dim % 4 != 0 || dim > 1024 || rows <= 0) return 0;
    return static_cast<int64_t>(layer_norm_bwd_ctas(dim, rows)) * 2 * dim * static_cast<int64_t>(sizeof(float));
}

int fm_layer_norm_bwd(const FmNormBwdParams* p, void* stream) {
    if (!p) return fail(FM_ERR_INVALID_ARG, "fm_layer_norm_bwd: params is null");
    if (p->abi_version != FM_SCAN_ABI_VERSION) return fail(FM_ERR_INVALID_ARG, "fm_layer_norm_bwd: abi_version mismatch");
    if (p->dim <= 0 || p->dim % 4 != 0 || p->dim > 1024 || p->rows <= 0 || p->reserved0 != 0 || !(p->eps >= 0.f))
        return fail(FM_ERR_INVALID_ARG, "fm_layer_norm_bwd: dim must be a multiple of 4 up to 1024, rows positive, eps >= 0");
    if (!p->x || !p->dy || !p->dx || !p->workspace)
        return fail(FM_ERR_INVALID_ARG, "fm_layer_norm_bwd: x, dy, dx and workspace must be non-null device pointers");
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
    if (!al16(p->x) || !al16(p->dy) || !al16(p->dx) || (p->weight && !al16(p->weight)) || !al16(p->workspace))
        return fail(FM_ERR_INVALID_ARG, "fm_layer_norm_bwd: pointers must be 16-byte aligned");
    if (p->workspace_bytes < fm_layer_norm_bwd_workspace_bytes(p->dim, p->rows))
        return fail(FM_ERR_INVALID_ARG, "fm_layer_norm_bwd: workspace smaller than fm_layer_norm_bwd_workspace_bytes()");
    cudaError_t e = launch_layer_norm_bwd(*p, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return fail(FM_ERR_CUDA, "fm_layer_norm_bwd: %s", cudaGetErrorString(e));
    return FM_OK;
}

static bool blk_shape_ok(int batch, int positions, int dim) {
    return batch > 0 && batch <= 65535 && positions > 0 && dim > 0 && dim % 4 == 0 && dim <= 1024;
}
static bool blk_al16(const void* q) { return q != nullptr && (reinterpret_cast<uintptr_t>(q) & 15u) == 0; }
static bool blk_dtype_ok(int d) { return d == FM_F32 || d == FM_F16 || d == FM_BF16; }

int64_t fm_block_gates_workspace_bytes(int32_t batch, int32_t positions, int32_t dim) {
    if (!blk_shape_ok(batch, positions, dim)) return 0;
    return static_cast<int64_t>(batch) * block_gates_slabs(batch, positions) * 2 * dim * static_cast<int64_t>(sizeof(float));
}

int fm_block_gates(const FmBlockGatesParams* p, void* stream) {
    if (!p) return fail(FM_ERR_INVALID_ARG, "fm_block_gates: params is null");
    if (p->abi_version != FM_SCAN_ABI_VERSION) return fail(FM_ERR_INVALID_ARG, "fm_block_gates: abi_version mismatch");
    if (!blk_dtype_ok(p->dtype) || !blk_shape_ok(p->batch, p->positions, p->dim) || p->reserved0 != 0 || !(p->eps >= 0.f))
        return fail(FM_ERR_INVALID_ARG, "fm_block_gates: dtype / shape (dim %% 4 == 0, dim <= 1024) / eps");
    if (!blk_al16(p->x) || !blk_al16(p->workspace) || p->workspace_bytes < fm_block_gates_workspace_bytes(p->batch, p->positions, p->dim))
        return fail(FM_ERR_INVALID_ARG, "fm_block_gates: x / workspace must be 16-byte aligned, workspace >= fm_block_gates_workspace_bytes()");
    if (!p->eca_scale && !p->se_gate) return fail(FM_ERR_INVALID_ARG, "fm_block_gates: no output requested");
    if (p->eca_scale && !p->eca_weight) return fail(FM_ERR_INVALID_ARG, "fm_block_gates: eca_scale needs eca_weight");
    if (p->se_gate && (!p->w1 || !p->w2 || p->reduce_dim <= 0 || p->reduce_dim > 1024))
        return fail(FM_ERR_INVALID_ARG, "fm_block_gates: se_gate needs w1, w2 and 0 < reduce_dim <= 1024");
    cudaError_t e = launch_block_gates(*p, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return fail(FM_ERR_CUDA, "fm_block_gates: %s", cudaGetErrorString(e));
    return FM_OK;
}

int fm_block_scale(const FmBlockScaleParams* p, void* stream) {
    if (!p) return fail(FM_ERR_INVALID_ARG, "fm_block_scale: params is null");
    if (p->abi_version != FM_SCAN_ABI_VERSION) return fail(FM_ERR_INVALID_ARG, "fm_block_scale: abi_version mismatch");
    if (!blk_dtype_ok(p->dtype) || !blk_shape_ok(p->batch, p->positions, p->dim) || p->reserved0 != 0)
        return fail(FM_ERR_INVALID_ARG, "fm_block_scale: dtype / shape (dim %% 4 == 0, dim <= 1024)");
    if (!blk_al16(p->x) || !blk_al16(p->gate) || !blk_al16(p->y))
        return fail(FM_ERR_INVALID_ARG, "fm_block_scale: x, gate, y must be non-null and 16-byte aligned");
    cudaError_t e = launch_block_scale(*p, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return fail(FM_ERR_CUDA, "fm_block_scale: %s", cudaGetErrorString(e));
    return FM_OK;
}

int fm_block_combine_norm(const FmBlockCombineParams* p, void* stream) {
    if (!p) return fail(FM_ERR_INVALID_ARG, "fm_block_combine_norm: params is null");
    if (p->abi_version != FM_SCAN_ABI_VERSION) return fail(FM_ERR_INVALID_ARG, "fm_block_combine_norm: abi_version mismatch");
    if (!blk_dtype_ok(p->dtype) || !blk_shape_ok(p->batch, p->positions, p->dim) || !(p->eps >= 0.f) || p->reserved0 != 0)
        return fail(FM_ERR_INVALID_ARG, "fm_block_combine_norm: dtype / shape (dim %% 4 == 0, dim <= 1024) / eps");
    if (p->input_dtype != FM_F32 && p->input_dtype != p->dtype)
        return fail(FM_ERR_INVALID_ARG, "fm_block_combine_norm: input_dtype must be fp32 or equal to dtype");
    // norm-only form: x_ssm == x_conv == gates == x_out == NULL -> y_out = LayerNorm(input) in `dtype`
    const bool norm_only = !p->x_ssm && !p->x_conv && !p->gate_ssm && !p->gate_conv && !p->x_out;
    if (!blk_al16(p->input) || !blk_al16(p->y_out) || (p->ln_weight && !blk_al16(p->ln_weight)) || (p->ln_bias && !blk_al16(p->ln_bias)) ||
        (!norm_only && (!blk_al16(p->x_ssm) || !blk_al16(p->x_conv) || !blk_al16(p->gate_ssm) || !blk_al16(p->gate_conv) || !blk_al16(p->x_out))))
        return fail(FM_ERR_INVALID_ARG, "fm_block_combine_norm: pointers must be non-null and 16-byte aligned (the branch inputs, gates and x_out may all be NULL together)");
    cudaError_t e = launch_block_combine(*p, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return fail(FM_ERR_CUDA, "fm_block_combine_norm: %s", cudaGetErrorString(e));
    return FM_OK;
}

int fm_conv_unfold(const FmConvUnfoldParams* p, void* stream) {
    if (!p) return fail(FM_ERR_INVALID_ARG, "fm_conv_unfold: params is null");
    if (p->abi_version != FM_SCAN_ABI_VERSION) return fail(FM_ERR_INVALID_ARG, "fm_conv_unfold: abi_version mismatch");
    if (p->dtype != FM_F32 && p->dtype != FM_F16 && p->dtype != FM_BF16)
        return fail(FM_ERR_INVALID_ARG, "fm_conv_unfold: dtype must be fp32, fp16 or bf16");
    if (p->batch <= 0 || p->batch > 65535 || p->dim <= 0 || p->h <= 0 || p->w <= 0 || !p->src || !p->dst || !p->weight ||
        p->src_channel_offset < 0 || p->src_channel_stride < p->src_channel_offset + p->dim || p->reserved0 != 0)
        return fail(FM_ERR_INVALID_ARG, "fm_conv_unfold: bad shape, stride or null pointer");
    cudaError_t e = launch_conv_unfold(*p, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return fail(FM_ERR_CUDA, "fm_conv_unfold: %s", cudaGetErrorString(e));
    return FM_OK;
}

int fm_conv_unfold_bwd(const FmConvUnfoldBwdParams* p, void* stream) {
    if (!p) return fail(FM_ERR_INVALID_ARG, "fm_conv_unfold_bwd: params is null");
    if (p->abi_version != FM_SCAN_ABI_VERSION) return fail(FM_ERR_INVALID_ARG, "fm_conv_unfold_bwd: abi_version mismatch");
    if (p->dtype != FM_F32 && p->dtype != FM_F16 && p->dtype != FM_BF16)
        return fail(FM_ERR_INVALID_ARG, "fm_conv_unfold_bwd: dtype must be fp32, fp16 or bf16");
    if (p->batch <= 0 || p->batch > 65535 || p->dim <= 0 || p->h <= 0 || p->w <= 0 || !p->src || !p->dxs || !p->dsrc || !p->weight ||
        !p->dweight || p->src_channel_offset < 0 || p->src_channel_stride < p->src_channel_offset + p->dim ||
        p->dsrc_channel_offset < 0 || p->dsrc_channel_stride < p->dsrc_channel_offset + p->dim)
        return fail(FM_ERR_INVALID_ARG, "fm_conv_unfold_bwd: bad shape, stride or null pointer");
    cudaError_t e = launch_conv_unfold_bwd(*p, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return fail(FM_ERR_CUDA, "fm_conv_unfold_bwd: %s", cudaGetErrorString(e));
    return FM_OK;
}

int fm_dt_proj(const FmDtProjParams* p, void* stream) {
    if (!p) return fail(FM_ERR_INVALID_ARG, "fm_dt_proj: params is null");
    if (p->abi_version != FM_SCAN_ABI_VERSION) return fail(FM_ERR_INVALID_ARG, "fm_dt_proj: abi_version mismatch");
    if (p->dtype != FM_F32 && p->dtype != FM_F16 && p->dtype != FM_BF16)
        return fail(FM_ERR_INVALID_ARG, "fm_dt_proj: dtype must be fp32, fp16 or bf16");
    if (p->weight_dtype != p->dtype && p->weight_dtype != FM_F32)
        return fail(FM_ERR_INVALID_ARG, "fm_dt_proj: weight_dtype must be dtype or fp32");
    if (p->batch <= 0 || p->n_groups <= 0 || p->dim <= 0 || p->seqlen <= 0 || !p->src || !p->dst || !p->weight ||
        (int64_t)p->batch * p->n_groups > 65535 || (p->dim + 31) / 32 > 65535)
        return fail(FM_ERR_INVALID_ARG, "fm_dt_proj: bad shape or null pointer");
    if (p->rank < 1 || p->rank > 12) return fail(FM_ERR_INVALID_ARG, "fm_dt_proj: rank must be 1..12 (use a GEMM beyond)");
    cudaError_t e = launch_dt_proj(*p, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return fail(FM_ERR_CUDA, "fm_dt_proj: %s", cudaGetErrorString(e));
    return FM_OK;
}

const char* fm_last_error(void) { return g_err; }
int fm_abi_version(void) { return FM_SCAN_ABI_VERSION; }
int fm_target_sm(void) { return 100; }
int64_t fm_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

}  // extern "C"
