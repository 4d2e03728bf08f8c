// Where do the ~2000 clk of the attention kernel's per-tile exp phase go?  One warp per scheduler (4 warps per CTA, one CTA
// per SM) runs the phase on 128 values held in registers, in variants that add one ingredient at a time.  clock64 per variant.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o softmax_phase tools/micro/softmax_phase.cu && ./softmax_phase
#include <cstdio>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ unsigned pack(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<unsigned*>(&v);
}

template <int VAR>
__global__ void __launch_bounds__(128) k(const float* in, unsigned* out, long long* clk, float scale, float m) {
  float s[128];
#pragma unroll
  for (int i = 0; i < 128; ++i) s[i] = in[(i * 128 + threadIdx.x) & 4095];
  unsigned pk[64];
  float ls[4] = {0.f, 0.f, 0.f, 0.f};
  float mx[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) mx[i] = -1e30f;
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int rep = 0; rep < 8; ++rep) {
    if (VAR >= 4) {
#pragma unroll
      for (int i = 0; i < 128; ++i) mx[i & 7] = fmaxf(mx[i & 7], s[i]);
      m += 1e-30f * (((mx[0] + mx[1]) + (mx[2] + mx[3])) + ((mx[4] + mx[5]) + (mx[6] + mx[7])));
    }
#pragma unroll
    for (int i = 0; i < 64; ++i) {
      float x0 = s[2 * i], x1 = s[2 * i + 1];
      if (VAR >= 1) { x0 = fmaf(x0, scale, -m); x1 = fmaf(x1, scale, -m); }
      const float p0 = ex2(x0), p1 = ex2(x1);
      if (VAR >= 2) ls[i & 3] += p0 + p1;
      if (VAR >= 3) pk[i] = pack(p0, p1);
      else pk[i] = __float_as_uint(p0) ^ __float_as_uint(p1);
    }
#pragma unroll
    for (int i = 0; i < 64; ++i) s[2 * i] += __uint_as_float(pk[i] & 1u);  // keep the values live across repetitions
  }
  const long long t1 = clock64();
  unsigned acc = 0;
#pragma unroll
  for (int i = 0; i < 64; ++i) acc ^= pk[i];
  out[blockIdx.x * 128 + threadIdx.x] = acc ^ __float_as_uint(ls[0] + ls[1] + ls[2] + ls[3]);
  if (blockIdx.x == 0 && threadIdx.x == 0) clk[VAR] = (t1 - t0) / 8;
}

int main() {
  float* in; unsigned* out; long long* clk;
  cudaMalloc(&in, 4096 * 4); cudaMalloc(&out, 148 * 2 * 128 * 4); cudaMalloc(&clk, 8 * 8);
  cudaMemset(in, 0, 4096 * 4);
  const char* names[] = {"128 ex2", "+ 128 FFMA", "+ row sums (128 FADD)", "+ 64 bf16x2 packs", "+ row max (128 FMNMX)"};
  for (int ctas = 148; ctas <= 296; ctas += 148) {
    k<0><<<ctas, 128>>>(in, out, clk, 0.18f, 1.f); k<1><<<ctas, 128>>>(in, out, clk, 0.18f, 1.f);
    k<2><<<ctas, 128>>>(in, out, clk, 0.18f, 1.f); k<3><<<ctas, 128>>>(in, out, clk, 0.18f, 1.f);
    k<4><<<ctas, 128>>>(in, out, clk, 0.18f, 1.f);
    cudaDeviceSynchronize();
    long long h[8]; cudaMemcpy(h, clk, 64, cudaMemcpyDeviceToHost);
    printf("%d CTAs of 4 warps (%d warp(s) per scheduler): clk per 128-value phase\n", ctas, ctas / 148);
    for (int v = 0; v < 5; ++v) printf("  %-28s %6lld\n", names[v], h[v]);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
