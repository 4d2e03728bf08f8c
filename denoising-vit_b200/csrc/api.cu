// extern "C" entry points declared in include/dvt_b200.h.
#include "../../include/dvt_b200.h"

#include "common.cuh"
#include "gemm.cuh"

namespace dvt {
const char* last_error();
extern int g_debug_impl_override;
int g_debug_impl_override = -1;
}  // namespace dvt

using namespace dvt;

static inline int eff_impl() { return dvt::g_debug_impl_override >= 0 ? dvt::g_debug_impl_override : -1; }

extern "C" {

int dvt_version(void) { return 100; }

const char* dvt_last_error(void) { return dvt::last_error(); }

int dvt_device_error(unsigned int* code_out) {
  unsigned int v = 0, z = 0;
  cudaError_t e = cudaMemcpyFromSymbol(&v, dvt::g_dvt_dev_error, sizeof(v));
  if (e != cudaSuccess) return dvt::cuda_fail(e, "read device error word", __FILE__, __LINE__);
  if (v) cudaMemcpyToSymbol(dvt::g_dvt_dev_error, &z, sizeof(z));
  if (code_out) *code_out = v;
  return DVT_OK;
}

int dvt_set_debug_impl(int impl) {
  if (impl != 0 && impl != 1 && impl != -1) {
    dvt::set_last_error("dvt_set_debug_impl: impl must be -1, 0 or 1");
    return DVT_ERR_INVALID;
  }
  dvt::g_debug_impl_override = impl;
  return DVT_OK;
}

int dvt_gemm_tn(const void* A, int lda, const void* B, int ldb, int dtype, int M, int N, int K, const float* bias,
                int act, void* out, int ldo, int out_dtype, int splits, void* stream) {
  DVT_REQUIRE(dtype == DVT_DTYPE_BF16 || dtype == DVT_DTYPE_F32, "dvt_gemm_tn: bad dtype %d", dtype);
  DVT_REQUIRE(out_dtype == DVT_DTYPE_BF16 || out_dtype == DVT_DTYPE_F32, "dvt_gemm_tn: bad out_dtype %d", out_dtype);
  DVT_REQUIRE(A && B && out, "dvt_gemm_tn: null pointer");
  GemmEpi e;
  e.bias = bias;
  e.act = act;
  e.out = out;
  e.ldo = ldo;
  if (splits > 1) {
    DVT_REQUIRE(out_dtype == DVT_DTYPE_F32 && act == 0, "dvt_gemm_tn: split-K needs fp32 output and no activation");
    e.out_mode = OUT_F32_ATOMIC;
  } else {
    e.out_mode = out_dtype == DVT_DTYPE_BF16 ? OUT_BF16 : OUT_F32;
  }
  GemmShape s{M, N, K, splits < 1 ? 1 : splits};
  return launch_gemm_tn(A, lda, B, ldb, dtype == DVT_DTYPE_BF16 ? TMAP_BF16 : TMAP_F32, s, e,
                        reinterpret_cast<cudaStream_t>(stream), eff_impl());
}

int dvt_gemm_tn_residual(const void* A, int lda, const void* B, int ldb, int dtype, int M, int N, int K,
                         const float* bias, const float* gamma, float* x_inout, int ldx, void* stream) {
  DVT_REQUIRE(dtype == DVT_DTYPE_BF16 || dtype == DVT_DTYPE_F32, "dvt_gemm_tn_residual: bad dtype %d", dtype);
  DVT_REQUIRE(A && B && x_inout, "dvt_gemm_tn_residual: null pointer");
  GemmEpi e;
  e.bias = bias;
  e.gamma = gamma;
  e.out = x_inout;
  e.ldo = ldx;
  e.out_mode = OUT_F32_RESID;
  GemmShape s{M, N, K, 1};
  return launch_gemm_tn(A, lda, B, ldb, dtype == DVT_DTYPE_BF16 ? TMAP_BF16 : TMAP_F32, s, e,
                        reinterpret_cast<cudaStream_t>(stream), eff_impl());
}

}  // extern "C"
