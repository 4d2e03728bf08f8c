/*
 * dvt_b200 -- C ABI of the B200-native (sm_100a) hot paths of Denoising-ViT (DVT).
 *
 * The reference (Jiawei-Yang/Denoising-ViT) has no FFI layer of its own: its hot paths are reached through the
 * Python API of `dvt.models`, which in turn calls timm (ViT forward), tiny-cuda-nn (hash-grid encoding) and
 * torch (Linear / grid_sample / Adam).  This header is the boundary a maintainer binds instead of those
 * libraries; every entry point cites the reference interface it replaces.  See INTEGRATION.md for the ctypes
 * stubs on the reference side.
 *
 * Conventions
 *   - every function returns 0 on success; non-zero codes are DVT_ERR_*; dvt_last_error() gives the message
 *     (thread-local).  Nothing aborts the process.
 *   - all tensor pointers are DEVICE pointers owned by the caller unless a parameter name ends in `_host`.
 *   - `stream` is a cudaStream_t passed as void*; work is enqueued, not synchronised.
 *   - hot calls do not allocate: handles own their workspaces, sized at create time.
 *   - bf16 tensors are passed as void* (uint16 storage).
 */
#ifndef DVT_B200_H_
#define DVT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DVT_OK 0
#define DVT_ERR_INVALID 1
#define DVT_ERR_CUDA 2
#define DVT_ERR_DEVICE 3

#define DVT_DTYPE_BF16 0
#define DVT_DTYPE_F32 1

/* ---------------------------------------------------------------------------------------------------------
 * library
 * ------------------------------------------------------------------------------------------------------- */
int dvt_version(void);
const char* dvt_last_error(void);
/* Reads (and clears) the device-side error word written by a kernel watchdog; 0 = none. */
int dvt_device_error(unsigned int* code_out);
/* Process-wide kernel implementation switch for debugging: 0 = tcgen05 tensor-core kernels (default),
 * 1 = plain SIMT reference kernels (same semantics, slow).  Also settable with DVT_GEMM_IMPL=simt. */
int dvt_set_debug_impl(int impl);

/* ---------------------------------------------------------------------------------------------------------
 * unit operators (kernel-level parity tests; each is also a building block of the two paths below)
 * ------------------------------------------------------------------------------------------------------- */

/* C[M,N] = act(A[M,K] . B[N,K]^T + bias) with A, B row-major, K contiguous (torch nn.Linear convention:
 * B is the Linear weight).  dtype: DVT_DTYPE_BF16 (bf16 operands) or DVT_DTYPE_F32 (fp32 operands, TF32
 * tensor-core math).  act: 0 none, 1 GELU(erf), 2 ReLU.  out_dtype: bf16 or f32.  splits > 1 accumulates
 * split-K partial sums atomically into a zero-initialised fp32 `out`.
 * Replaces: torch.nn.Linear / cuBLAS calls made by timm Block (qkv, proj, fc1, fc2) and by
 * dvt/models/neural_feature_field.py:40-44, dvt/models/offline_denoiser.py:40-46. */
int dvt_gemm_tn(const void* A, int lda, const void* B, int ldb, int dtype, int M, int N, int K, const float* bias,
                int act, void* out, int ldo, int out_dtype, int splits, void* stream);

/* out[m,n] += gamma[n] * (A.B^T + bias)[m,n] on an fp32 residual stream (LayerScale + residual add).
 * Replaces: `x = x + ls(attn(...))` / `x = x + ls(mlp(...))` in timm Block.forward
 * (restated in the reference at evaluation/vitdet/vision_transformer.py:98-117). */
int dvt_gemm_tn_residual(const void* A, int lda, const void* B, int ldb, int dtype, int M, int N, int K,
                         const float* bias, const float* gamma, float* x_inout, int ldx, void* stream);

/* y = LayerNorm(x) over rows of C fp32 (eps inside the sqrt, affine gamma/beta), output bf16 or fp32.
 * in_group/skip: rows are grouped in runs of `in_group`; the first `skip` rows of every group are dropped and the
 * output is compacted (used to strip prefix tokens); pass in_group=1, skip=0 for a plain LayerNorm.
 * Replaces: nn.LayerNorm(eps=1e-6) in timm Block / VisionTransformer.norm. */
int dvt_layernorm(const float* x, int ldx, const float* gamma, const float* beta, void* y, int ldy, int y_dtype,
                  int rows, int C, float eps, int in_group, int skip, void* stream);

/* Multi-head attention, head_dim 64, no mask, scale 1/8: qkv bf16 [B, N, 3*heads*64] -> out bf16 [B, N, heads*64].
 * Replaces: timm Attention.forward -> F.scaled_dot_product_attention
 * (reference restatement: evaluation/vitdet/vision_transformer.py:73-91). */
int dvt_attention_fwd(const void* qkv_bf16, void* out_bf16, int B, int N, int heads, void* stream);

/* Patch extraction for Conv2d(3->C, kernel P, stride S): x [B,3,H,W] (f32 or bf16) -> bf16 [B*h*w, Kp],
 * Kp = round_up(3*P*P, 8), column = c*P*P + i*P + j.  h = (H-P)/S+1, w = (W-P)/S+1
 * (dvt/models/vit_wrapper.py:78-91: stride override + dynamic_feat_size). */
int dvt_im2col(const void* x, int x_dtype, void* out_bf16, int B, int H, int W, int P, int S, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * HP-1: frozen ViT forward (replaces PretrainedViTWrapper.get_intermediate_layers,
 * dvt/models/vit_wrapper.py:122-143, i.e. timm VisionTransformer.forward_intermediates)
 * ------------------------------------------------------------------------------------------------------- */
typedef struct dvt_vit dvt_vit_t;

/* prefix_tokens = 1 (cls) + number of register tokens.  Only head_dim 64 (embed == 64*heads). */
int dvt_vit_create(dvt_vit_t** out, int embed, int depth, int heads, int patch, int mlp_hidden, int swiglu,
                   int layerscale, int prefix_tokens, float ln_eps);
void dvt_vit_destroy(dvt_vit_t* h);
/* Loads one fp32 tensor by its timm state-dict key (without the wrapper's "model." prefix), e.g.
 * "blocks.3.attn.qkv.weight".  `src` may be host or device memory.  cls_token / reg_token / pos_embed are not
 * loaded here: the caller passes the (resampled) position table and the prefix rows to dvt_vit_forward. */
int dvt_vit_load(dvt_vit_t* h, const char* timm_key, const float* src, size_t numel);
/* Pre-sizes the activation workspaces (otherwise grown on first use). */
int dvt_vit_reserve(dvt_vit_t* h, int max_batch, int H, int W, int stride);
/* x: [B,3,H,W] f32 or bf16.  pos_patch: f32 [h*w, C] position embedding of the patch tokens for this grid
 * (already resampled).  prefix_rows: f32 [prefix_tokens, C] rows written in front of the patches (cls + its
 * position, register tokens).  Runs blocks 0..layer_index, applies the final LayerNorm if `norm`.
 * out (f32): all_tokens == 0 -> [B, h, w, C] (prefix stripped, NHWC); all_tokens == 1 -> [B, prefix + h*w, C]. */
int dvt_vit_forward(dvt_vit_t* h, const void* x, int x_dtype, int B, int H, int W, int stride,
                    const float* pos_patch, const float* prefix_rows, int layer_index, int norm, float* out,
                    int all_tokens, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DVT_B200_H_ */
