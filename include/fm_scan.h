/*
 * fm_scan.h -- C ABI of libfm_scan.so: the B200 (sm_100a) selective-scan / SS2D hot path.
 *
 * Drop-in boundary.  Every entry point replaces one interface of the reference (paths relative to the
 * upstream FusionMamba repository):
 *
 *   fm_selective_scan_fwd   <- selective_scan_cuda.fwd   selective_scan/selective_scan.cpp:226-336
 *                              (param record: SSMParamsBase, selective_scan/selective_scan.h:26-66)
 *   fm_selective_scan_bwd   <- selective_scan_cuda.bwd   selective_scan/selective_scan.cpp:338-492
 *                              (param record: SSMParamsBwd,  selective_scan/selective_scan.h:68-101)
 *   fm_scan_unfold / fm_scan_merge
 *                           <- EfficientScan / EfficientMerge fwd+bwd   models/cross.py:34-88, 139-190
 *                              and the classic CrossScan / CrossMerge    models/cross.py:610-612, 639-642
 *   fm_conv_unfold          <- permute + depthwise conv2d + SiLU + EfficientScan           models/cross.py:727-731, 297
 *   fm_merge_norm           <- y.transpose(1, 2).contiguous(); out_norm(y); .to(x.dtype)   models/cross.py:334-337
 *   fm_layer_norm_bwd       <- autograd backward of out_norm (nn.LayerNorm)               models/cross.py:334-335
 *   fm_block_gates / fm_block_scale / fm_block_combine_norm
 *                           <- ECA, BiAttn, residual adds and norm2 of VSSBlock_new._forward      models/cross.py:1362-1377
 *   FmScanFwdParams.out_map (EfficientMerge fused into the forward kernel's store)
 *                           <- EfficientMerge inside cross_selective_scan (models/cross.py:328): ys (B, 4, D, L) is never
 *                              materialised.  The unfold is NOT fused into the scan's load: xs (B, 4, D, L) is also the operand of
 *                              the x_proj GEMM (models/cross.py:305) and has to exist anyway -- fm_conv_unfold produces it in
 *                              one pass instead.  u_map must be FM_MAP_LINEAR.
 *
 * Conventions: plain C, raw DEVICE pointers, sizes in elements, strides in ELEMENTS (int64, unlike the
 * reference's uint32 strides), last (sequence) dimension contiguous.  Calls are asynchronous on `stream`
 * (a cudaStream_t passed as void*), stateless and thread-safe.  Return 0 on success; on failure a non-zero
 * FmStatus is returned and fm_last_error() (thread-local) describes it -- the Python shim raises
 * RuntimeError(fm_last_error()), mirroring the reference's TORCH_CHECK failures.
 */
#ifndef FM_SCAN_H_
#define FM_SCAN_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FM_SCAN_ABI_VERSION 3

typedef enum FmStatus {
    FM_OK = 0,
    FM_ERR_INVALID_ARG = 1,   /* shape / dtype / stride precondition violated (TORCH_CHECK analogue) */
    FM_ERR_UNSUPPORTED = 2,   /* valid for the reference, outside this library's scope (complex A, constant B/C) */
    FM_ERR_CUDA = 3           /* CUDA runtime error at launch */
} FmStatus;

/* dtype of u, delta, B, C, z, out, dout, du, ddelta, dz (A, D, delta_bias, x, dA, dD, ddelta_bias: fp32) */
typedef enum FmDtype { FM_F32 = 0, FM_F16 = 1, FM_BF16 = 2 } FmDtype;

/* Index maps.  FmPermuteParams.map (fm_scan_unfold / fm_scan_merge): CROSS_V0 = classic 4-direction CrossScan, seqlen == H*W,
 * merge = 4-way sum; EFFICIENT_V2 = the 4 stride-2 sub-grids of EfficientScan, seqlen == ceil(H/2)*ceil(W/2), merge = permutation.
 * FmScanFwdParams.out_map (forward only): LINEAR = out is (batch, dim, seqlen) as in the reference op; EFFICIENT_V2 /
 * EFFICIENT_V2_CL = EfficientMerge fused into the store.  FmScanFwdParams.u_map: LINEAR only (see the header comment). */
typedef enum FmIndexMap { FM_MAP_LINEAR = 0, FM_MAP_CROSS_V0 = 1, FM_MAP_EFFICIENT_V2 = 2,
                          /* out_map only: EfficientMerge fused into the store AND channels-last output y (batch, H*W, dim/4):
                             out_batch_stride = stride of batch, out_d_stride = stride of a pixel (>= dim/4), channel stride 1.
                             The 16 rows of a CTA at one pixel are 64 contiguous bytes, and LayerNorm needs no transpose. */
                          FM_MAP_EFFICIENT_V2_CL = 3 } FmIndexMap;

typedef struct FmScanFwdParams {
    int32_t abi_version;       /* FM_SCAN_ABI_VERSION */
    int32_t dtype;             /* FmDtype */
    int32_t batch, dim, seqlen, dstate, n_groups;
    int32_t n_chunks;          /* checkpoint slots in x: ceil(seqlen / chunk_len) */
    int32_t chunk_len;         /* timesteps per checkpoint slot; the reference uses 2048 (selective_scan.cpp:307) */
    int32_t delta_softplus;    /* bool */
    int32_t u_map, out_map;    /* FmIndexMap.  u_map: must be FM_MAP_LINEAR.  out_map: LINEAR, or EFFICIENT_V2 / EFFICIENT_V2_CL (merge on store) */
    int32_t map_h, map_w;      /* image H, W for the non-linear maps */
    int32_t hck_len;           /* spacing (timesteps, multiple of 16) of the dense state checkpoints in hck; 0 if hck == NULL */
    int32_t n_hck;             /* ceil(seqlen / hck_len) - 1 interior boundaries */
    int32_t out_dtype;         /* FmDtype of `out`.  Equal to dtype, except that a forward without z may write fp32 `out` from
                                  16-bit inputs: the reference upcasts bf16/fp16 activations before the scan and keeps y in
                                  fp32 (models/cross.py:312-318); reading the 16-bit tensors directly is bit-identical
                                  (the upcast is exact) and skips the casts */
    int32_t reserved0;         /* must be 0 (keeps the 64-bit members aligned) */
    /* element strides; sequence stride is 1 */
    int64_t u_batch_stride, u_d_stride;
    int64_t delta_batch_stride, delta_d_stride;
    int64_t z_batch_stride, z_d_stride;
    int64_t out_batch_stride, out_d_stride;
    int64_t out_z_batch_stride, out_z_d_stride;
    int64_t A_d_stride, A_dstate_stride;
    int64_t B_batch_stride, B_group_stride, B_dstate_stride;
    int64_t C_batch_stride, C_group_stride, C_dstate_stride;
    const void *u, *delta, *A, *B, *C;
    const void *D;             /* (dim) fp32 or NULL */
    const void *z;             /* (batch, dim, seqlen) or NULL */
    const void *delta_bias;    /* (dim) fp32 or NULL */
    void *out;                 /* y (pre-gate) */
    void *out_z;               /* y * silu(z); required iff z != NULL */
    void *x;                   /* (batch, dim, n_chunks, 2*dstate) fp32 contiguous: [2n]=running decay product,
                                  [2n+1]=state h at the end of the slot; x[:, :, -1, 1::2] is last_state
                                  (selective_scan_interface.py:46) */
    void *hck;                 /* optional (batch, dim, n_hck, dstate) fp32: state h after timestep (j+1)*hck_len-1.
                                  Written by fwd when non-NULL; READ by bwd (required there when seqlen > hck_len):
                                  lets the backward start any chunk without re-running the forward recurrence.
                                  Implementation detail of this library (the reference keeps only x). */
    void *workspace;           /* optional scratch for the forward (device memory, 16-byte aligned, contents undefined before and
                                  after the call) of at least fm_scan_fwd_workspace_bytes() bytes, or NULL.  With it, a forward
                                  over few rows and a long sequence (one 1024x1024 pair: 768 rows x 65536 steps) is split in time
                                  over several CTAs per row; without it the single-pass kernel runs.  The library never allocates. */
    int64_t workspace_bytes;
} FmScanFwdParams;

typedef struct FmScanBwdParams {
    FmScanFwdParams f;         /* same inputs as forward; f.out = saved y (needed iff z != NULL), f.x = saved checkpoints,
                                  f.out_z = optional recomputed out_z (or NULL) */
    int64_t dout_batch_stride, dout_d_stride;
    int64_t du_batch_stride, du_d_stride;
    int64_t ddelta_batch_stride, ddelta_d_stride;
    int64_t dz_batch_stride, dz_d_stride;
    int64_t dB_batch_stride, dB_group_stride, dB_dstate_stride;   /* fp32 accumulators, zero-initialised by caller */
    int64_t dC_batch_stride, dC_group_stride, dC_dstate_stride;
    const void *dout;          /* (batch, dim, seqlen) */
    void *du, *ddelta;         /* (batch, dim, seqlen), dtype */
    void *dz;                  /* (batch, dim, seqlen) dtype; required iff z != NULL */
    float *dA;                 /* (dim, dstate) fp32, zero-initialised by caller, row stride dstate */
    float *dB, *dC;            /* (batch, n_groups, dstate, seqlen) fp32, zero-initialised by caller */
    float *dD;                 /* (dim) fp32 zero-initialised, or NULL */
    float *ddelta_bias;        /* (dim) fp32 zero-initialised, or NULL */
} FmScanBwdParams;

/* Stand-alone scan unfold / merge (both directions of both permutations; bit-exact data movement).
 *   unfold: src x (batch, dim, H, W)  -> dst xs (batch, 4, dim, Lk)      [EfficientScan.forward / CrossScan,
 *                                                                          == EfficientMerge.backward]
 *   merge : src ys (batch, 4, dim, Lk) -> dst y (batch, dim, H*W)         [EfficientMerge.forward == EfficientScan.backward;
 *                                                                          CROSS_V0: 4-way fp32 sum in the reference's order] */
typedef struct FmPermuteParams {
    int32_t abi_version;
    int32_t dtype;             /* FmDtype */
    int32_t map;               /* FM_MAP_CROSS_V0 or FM_MAP_EFFICIENT_V2 */
    int32_t batch, dim, h, w;
    const void *src;
    void *dst;
} FmPermuteParams;

/* SS2D epilogue: transpose + LayerNorm over the channel dimension + cast, one pass (inference path).
 *   src y (batch, dim, positions) fp32 contiguous  ->  dst (batch, positions, dim) out_dtype contiguous
 * replaces  y.transpose(1, 2).contiguous(); out_norm(y); .to(x.dtype)   models/cross.py:334-337
 * (nn.LayerNorm semantics: biased variance, eps inside the square root, fp32 affine weight/bias, either may be NULL). */
typedef struct FmNormParams {
    int32_t abi_version;
    int32_t out_dtype;         /* FmDtype of dst */
    int32_t batch, dim, positions;
    float eps;
    const void *src;           /* fp32; (batch, dim, positions), or (batch, positions, dim) if src_channels_last */
    const void *weight, *bias; /* fp32 (dim) or NULL */
    void *dst;
    /* optional gate (SS2D.forward, models/cross.py:728-729, 740): dst = LayerNorm(y) * SiLU(gate), gate read from a channels-last
       tensor (batch, positions, gate_channel_stride) of dtype out_dtype at channels [gate_channel_offset, +dim).  For 16-bit
       outputs the LayerNorm result and SiLU(gate) are each rounded to out_dtype before the product, like the reference's
       separate ops.  NULL: no gate. */
    const void *gate;
    int64_t gate_channel_stride;
    int32_t gate_channel_offset;
    int32_t src_channels_last; /* 0: src is (batch, dim, positions); 1: src is (batch, positions, dim) (FM_MAP_EFFICIENT_V2_CL output) */
} FmNormParams;

/* LayerNorm(dim) BACKWARD over channels-last rows (training side of the epilogue; nn.LayerNorm semantics as above):
 *   x, dy (rows, dim) fp32 contiguous  ->  dx (rows, dim) fp32,  dweight / dbias (dim) fp32 (written, not accumulated; either may
 *   be NULL).  Row statistics are recomputed from x; the column sums are reduced deterministically through `workspace`
 *   (caller-allocated device scratch of at least fm_layer_norm_bwd_workspace_bytes(dim, rows) bytes).  dim % 4 == 0, dim <= 1024.
 * replaces the autograd backward of  out_norm(y)  models/cross.py:334-335  and of the LayerNorms around the SS2D path. */
typedef struct FmNormBwdParams {
    int32_t abi_version;
    int32_t dim;
    int64_t rows;
    float eps;
    int32_t reserved0;         /* must be 0 */
    const void *x, *dy;        /* fp32 (rows, dim) */
    const void *weight;        /* fp32 (dim) or NULL (= ones) */
    void *dx;                  /* fp32 (rows, dim) */
    void *dweight, *dbias;     /* fp32 (dim) or NULL */
    void *workspace;
    int64_t workspace_bytes;
} FmNormBwdParams;

/* The inference tail of VSSBlock_new around the SS2D op (models/cross.py:1362-1377): three entry points (fm_block.cu).
 * All activations are channels-last (batch, positions, dim), dim % 4 == 0, dim <= 1024, contiguous, 16-byte aligned.
 *
 * fm_block_gates: one pass over x -> the two per-(batch, channel) gates of the block
 *   eca_scale = sigmoid(conv1d_k3(mean_pos(x)))                              eca_layer, models/cross.py:1236-1259
 *   se_gate   = sigmoid(W2 gelu(W1 mean_pos(LayerNorm(x)) + b1) + b2)        BiAttn,    models/cross.py:744-768
 * either output may be NULL (with its weights). */
typedef struct FmBlockGatesParams {
    int32_t abi_version;
    int32_t dtype;             /* FmDtype of x */
    int32_t batch, positions, dim;
    int32_t reduce_dim;        /* hidden width of the BiAttn bottleneck (dim / 8); 0 if se_gate == NULL */
    float eps;                 /* of the BiAttn LayerNorm */
    int32_t reserved0;         /* must be 0 */
    const void *x;
    const void *ln_weight, *ln_bias;      /* fp32 (dim) or NULL */
    const void *eca_weight;               /* fp32 (3): Conv1d(1, 1, 3, padding 1, bias=False) over the channel axis, or NULL */
    const void *w1, *b1, *w2, *b2;        /* fp32 (reduce_dim, dim), (reduce_dim), (dim, reduce_dim), (dim); biases may be NULL */
    void *eca_scale, *se_gate;            /* fp32 (batch, dim) or NULL */
    void *workspace;                      /* >= fm_block_gates_workspace_bytes(batch, positions, dim) bytes of device scratch */
    int64_t workspace_bytes;
} FmBlockGatesParams;

/* fm_block_scale:  y = x + x * gate[b, c]   (ECA apply and the add feeding the LDC conv, models/cross.py:1365-1369; products and
 * sum rounded to dtype like the reference's separate ops) */
typedef struct FmBlockScaleParams {
    int32_t abi_version;
    int32_t dtype;
    int32_t batch, positions, dim;
    int32_t reserved0;
    const void *x;             /* dtype */
    const void *gate;          /* fp32 (batch, dim) */
    void *y;                   /* dtype */
} FmBlockScaleParams;

/* fm_block_combine_norm:  x_out = input + (x_ssm * gate_ssm + x_conv * gate_conv)   (fp32 residual stream, models/cross.py:1370-1373)
 *                         y_out = LayerNorm(x_out) in dtype                          (norm2, the input of mlp.fc1, :1375) */
typedef struct FmBlockCombineParams {
    int32_t abi_version;
    int32_t dtype;             /* FmDtype of x_ssm, x_conv, y_out */
    int32_t batch, positions, dim;
    float eps;                 /* of norm2 */
    int32_t input_dtype;       /* FmDtype of input and x_out: FM_F32 or `dtype` (the residual stream has the activation dtype wherever an
                                  autocast Linear produced it, e.g. behind PatchMerging2D) */
    int32_t reserved0;         /* must be 0 */
    const void *input;         /* input_dtype */
    const void *x_ssm, *x_conv;
    const void *gate_ssm, *gate_conv;     /* fp32 (batch, dim) */
    const void *ln_weight, *ln_bias;      /* fp32 (dim) or NULL */
    void *x_out;               /* input_dtype */
    void *y_out;               /* dtype */
    /* norm-only form: x_ssm, x_conv, gate_ssm, gate_conv and x_out all NULL -> y_out = LayerNorm(input) rounded to `dtype`
     * (the block's first norm under autocast: the fp32 result of models/cross.py:1363 is only ever consumed by an autocast Linear) */
} FmBlockCombineParams;

/* SS2D prologue: depthwise 3x3 conv (padding 1) + bias + SiLU + EfficientScan unfold, one pass (inference path).
 *   src xz (batch, H, W, src_channel_stride) channels-last; the conv input is channels [src_channel_offset, +dim)
 *   ->  dst xs (batch, 4, dim, ceil(H/2)*ceil(W/2)), same dtype
 * replaces  x.permute(0,3,1,2).contiguous(); act(conv2d(x)); EfficientScan   models/cross.py:727-731, 297
 * weight (dim, 1, 3, 3) and bias (dim, or NULL) are fp32. */
typedef struct FmConvUnfoldParams {
    int32_t abi_version;
    int32_t dtype;             /* FmDtype of src and dst */
    int32_t batch, dim, h, w;
    int32_t src_channel_offset;
    int32_t reserved0;
    int64_t src_channel_stride;   /* elements between consecutive pixels of src (>= offset + dim) */
    const void *src;
    const void *weight, *bias;
    void *dst;
} FmConvUnfoldParams;

/* Backward of fm_conv_unfold (training path of SS2D.forward): one pass instead of EfficientScan.backward, the SiLU backward, the two
 * depthwise-conv backward kernels (data, weight) and the permute backward of   models/cross.py:171-190, 727-731
 *   src   xz  (batch, H, W, src_channel_stride) channels-last, conv input = channels [src_channel_offset, +dim): the forward's input
 *   dxs       (batch, 4, dim, ceil(H/2)*ceil(W/2)): gradient of the unfolded output, same dtype as src
 *   ->  dsrc  (batch, H, W, dsrc_channel_stride) channels-last, gradient of the conv input written at channels
 *             [dsrc_channel_offset, +dim) (so it can land inside the gradient of the in_proj output), same dtype
 *   ->  dweight (dim, 1, 3, 3), dbias (dim, or NULL): fp32, ACCUMULATED (atomics): the caller zeroes them
 * z = conv(x) + bias is recomputed from src; weight / bias are fp32. */
typedef struct FmConvUnfoldBwdParams {
    int32_t abi_version;
    int32_t dtype;             /* FmDtype of src, dxs and dsrc */
    int32_t batch, dim, h, w;
    int32_t src_channel_offset, dsrc_channel_offset;
    int64_t src_channel_stride, dsrc_channel_stride;   /* elements between consecutive pixels */
    const void *src;
    const void *weight, *bias;
    const void *dxs;
    void *dsrc;
    void *dweight, *dbias;     /* fp32 accumulators */
} FmConvUnfoldBwdParams;

/* Rank-R dt projection of the SS2D core (inference):  delta[b,k,d,l] = sum_r weight[k,d,r] * dts[b,k,r,l]
 * replaces  torch.einsum("b k r l, k d r -> b k d l", dts, dt_projs_weight)   models/cross.py:309-310
 * src is a strided view (last dim contiguous) of x_dbl, dst is contiguous (batch, n_groups, dim, seqlen); rank <= 12
 * (longer contractions belong to the GEMM library); fp32 accumulation, dst rounded to `dtype`. */
typedef struct FmDtProjParams {
    int32_t abi_version;
    int32_t dtype;             /* FmDtype of src and dst */
    int32_t weight_dtype;      /* FmDtype of weight: `dtype` or FM_F32 */
    int32_t batch, n_groups, dim, rank, seqlen;   /* dim = channels per group (d_inner) */
    int64_t src_batch_stride, src_group_stride, src_rank_stride;   /* elements */
    const void *src;
    const void *weight;        /* (n_groups, dim, rank) contiguous */
    void *dst;
} FmDtProjParams;

int fm_selective_scan_fwd(const FmScanFwdParams *params, void *stream);
/* Bytes of FmScanFwdParams.workspace this forward can make use of (0: none; host-only query, no CUDA call). */
int64_t fm_scan_fwd_workspace_bytes(const FmScanFwdParams *params);
int fm_selective_scan_bwd(const FmScanBwdParams *params, void *stream);
int fm_scan_unfold(const FmPermuteParams *params, void *stream);
int fm_scan_merge(const FmPermuteParams *params, void *stream);
int fm_merge_norm(const FmNormParams *params, void *stream);
int fm_layer_norm_bwd(const FmNormBwdParams *params, void *stream);
int64_t fm_layer_norm_bwd_workspace_bytes(int32_t dim, int64_t rows);   /* host-only query */
int fm_block_gates(const FmBlockGatesParams *params, void *stream);
int64_t fm_block_gates_workspace_bytes(int32_t batch, int32_t positions, int32_t dim);   /* host-only query */
int fm_block_scale(const FmBlockScaleParams *params, void *stream);
int fm_block_combine_norm(const FmBlockCombineParams *params, void *stream);
int fm_conv_unfold(const FmConvUnfoldParams *params, void *stream);
int fm_conv_unfold_bwd(const FmConvUnfoldBwdParams *params, void *stream);
int fm_dt_proj(const FmDtProjParams *params, void *stream);

/* Thread-local description of the last failure on the calling thread ("" if none). */
const char *fm_last_error(void);
/* ABI version the library was built with, and the SM architecture it targets (100 for sm_100a). */
int fm_abi_version(void);
int fm_target_sm(void);
/* Number of kernel launches issued by this library since load (all threads); bench.py reports the delta. */
int64_t fm_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* FM_SCAN_H_ */
