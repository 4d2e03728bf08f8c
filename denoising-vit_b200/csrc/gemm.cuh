// Shared declarations for the TN GEMM (C[M,N] = A[M,K] * B[N,K]^T, both operands K-contiguous) and its
// fused epilogues.  Used by the ViT forward (HP-1) and by the neural-field fit (HP-2).
#pragma once
#include "common.cuh"

namespace dvt {

enum GemmAct { ACT_NONE = 0, ACT_GELU = 1, ACT_RELU = 2 };
enum GemmOut {
  OUT_BF16 = 0,        // out[m, n] = v                     (bf16)
  OUT_F32 = 1,         // out[m, n] = v                     (fp32)
  OUT_F32_ATOMIC = 2,  // out[m, n] += v                    (fp32 atomics; split-K partial sums)
  OUT_F32_RESID = 3,   // out[m, n] += gamma[n] * v         (fp32 residual stream, in place; gamma may be null)
  OUT_F32_REMAP = 4,   // out[remap(m), n] = v + addend[m % rows_per_group, n]   (patch-embed -> token rows)
};

struct GemmEpi {
  // pre-stage (per element, thread-per-row registers)
  const float* bias = nullptr;        // [N] added to the accumulator
  int act = ACT_NONE;                 // activation applied after bias
  const __nv_bfloat16* mask = nullptr;  // optional [M, ldmask]: v *= (mask[m, n] > 0)   (ReLU backward)
  int ldmask = 0;
  float alpha = 1.0f;                 // v *= alpha (after activation / mask)
  __nv_bfloat16* out_t = nullptr;     // optional transposed bf16 copy: out_t[n, m], leading dim ldt
  int ldt = 0;
  // post-stage (coalesced)
  int out_mode = OUT_BF16;
  void* out = nullptr;
  int ldo = 0;
  const float* gamma = nullptr;       // LayerScale (OUT_F32_RESID)
  const float* addend = nullptr;      // [rows_per_group, N] (OUT_F32_REMAP)
  int rows_per_group = 0;             // patches per image
  int group_stride = 0;               // tokens per image
  int row_offset = 0;                 // prefix tokens
};

struct GemmShape {
  int M, N, K;
  int splits;  // split-K factor (>1 requires OUT_F32_ATOMIC)
};

__device__ __forceinline__ float epi_pre(const GemmEpi& e, int m, int n, float acc) {
  float v = acc;
  if (e.bias) v += __ldg(e.bias + n);
  if (e.act == ACT_GELU) v = gelu_erf(v);
  else if (e.act == ACT_RELU) v = fmaxf(v, 0.0f);
  if (e.mask) {
    float h = __bfloat162float(e.mask[(size_t)m * e.ldmask + n]);
    v = h > 0.0f ? v : 0.0f;
  }
  return v * e.alpha;
}

__device__ __forceinline__ size_t epi_out_row(const GemmEpi& e, int m) {
  if (e.out_mode == OUT_F32_REMAP) {
    int g = m / e.rows_per_group;
    int p = m - g * e.rows_per_group;
    return (size_t)g * e.group_stride + e.row_offset + p;
  }
  return (size_t)m;
}

// scalar post-stage (SIMT debug path and ragged tails)
__device__ __forceinline__ void epi_post1(const GemmEpi& e, int m, int n, float v) {
  size_t row = epi_out_row(e, m);
  switch (e.out_mode) {
    case OUT_BF16:
      reinterpret_cast<__nv_bfloat16*>(e.out)[row * e.ldo + n] = __float2bfloat16_rn(v);
      break;
    case OUT_F32:
      reinterpret_cast<float*>(e.out)[row * e.ldo + n] = v;
      break;
    case OUT_F32_ATOMIC:
      atomicAdd(reinterpret_cast<float*>(e.out) + row * e.ldo + n, v);
      break;
    case OUT_F32_RESID: {
      float* o = reinterpret_cast<float*>(e.out) + row * e.ldo + n;
      float g = e.gamma ? __ldg(e.gamma + n) : 1.0f;
      *o = fmaf(g, v, *o);
    } break;
    case OUT_F32_REMAP: {
      int p = m % e.rows_per_group;
      float a = e.addend ? __ldg(e.addend + (size_t)p * e.ldo + n) : 0.0f;
      reinterpret_cast<float*>(e.out)[row * e.ldo + n] = v + a;
    } break;
  }
}

// vector post-stage: 4 consecutive columns, n % 4 == 0, ldo % 4 == 0, all in range
__device__ __forceinline__ void epi_post4(const GemmEpi& e, int m, int n, float4 v) {
  size_t row = epi_out_row(e, m);
  switch (e.out_mode) {
    case OUT_BF16: {
      uint2 p;
      p.x = pack_bf16x2(v.x, v.y);
      p.y = pack_bf16x2(v.z, v.w);
      *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(e.out) + row * e.ldo + n) = p;
    } break;
    case OUT_F32:
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(e.out) + row * e.ldo + n) = v;
      break;
    case OUT_F32_ATOMIC: {
      float* o = reinterpret_cast<float*>(e.out) + row * e.ldo + n;
      atomicAdd(o + 0, v.x);
      atomicAdd(o + 1, v.y);
      atomicAdd(o + 2, v.z);
      atomicAdd(o + 3, v.w);
    } break;
    case OUT_F32_RESID: {
      float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(e.out) + row * e.ldo + n);
      float4 x = *o;
      float4 g = e.gamma ? __ldg(reinterpret_cast<const float4*>(e.gamma + n)) : make_float4(1.f, 1.f, 1.f, 1.f);
      x.x = fmaf(g.x, v.x, x.x);
      x.y = fmaf(g.y, v.y, x.y);
      x.z = fmaf(g.z, v.z, x.z);
      x.w = fmaf(g.w, v.w, x.w);
      *o = x;
    } break;
    case OUT_F32_REMAP: {
      int p = m % e.rows_per_group;
      float4 a = e.addend ? __ldg(reinterpret_cast<const float4*>(e.addend + (size_t)p * e.ldo + n))
                          : make_float4(0.f, 0.f, 0.f, 0.f);
      v.x += a.x;
      v.y += a.y;
      v.z += a.z;
      v.w += a.w;
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(e.out) + row * e.ldo + n) = v;
    } break;
  }
}

enum GemmImpl { GEMM_TCGEN05 = 0, GEMM_SIMT_DEBUG = 1 };

// dtype: TMAP_BF16 (kind::f16, bf16 operands) or TMAP_F32 (kind::tf32, fp32 operands).
// lda / ldb in elements.  Both A and B are row-major with K contiguous.
int launch_gemm_tn(const void* A, int lda, const void* B, int ldb, TmapDtype dtype, const GemmShape& shape,
                   const GemmEpi& epi, cudaStream_t stream, int impl = -1 /* -1: process default */);

int default_gemm_impl();

}  // namespace dvt
