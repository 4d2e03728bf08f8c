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

#ifdef __cplusplus
}
#endif
#endif /* DVT_B200_H_ */
