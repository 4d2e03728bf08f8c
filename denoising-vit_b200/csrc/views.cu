// GPU view generation, the step in front of HP-1.
// Replaces the reference's CPU path -- `RandomResizedCropFlip.forward` (dvt/dataset/transform.py:39-76) run by 8
// DataLoader workers (main_img_denoising.py:277-310): per view `F.resized_crop(img, top, left, h, w, size, BICUBIC,
// antialias=True)` (+ `F.hflip`) and the [h_patches, w_patches, 2] grid of (x, y) patch coordinates of the crop inside
// the image.  One launch produces all views of an image from the device-resident normalised image; the crop boxes and
// flip decisions stay on the host (same torch / numpy RNG calls as the reference), 5 ints per view.
//
// Resampling = ATen's separable anti-aliased bicubic (`_upsample_bicubic2d_aa`, what torchvision calls for tensors):
//   scale = in / out, support = 2 * max(scale, 1), center = scale * (i + 0.5),
//   taps xmin = max(int(center - support + 0.5), 0) .. min(int(center + support + 0.5), in), weights = cubic(a = -0.5)
//   of (tap - center + 0.5) / max(scale, 1), normalised to sum 1; horizontal pass first, then vertical.
// One CTA computes a 64 x 16 output tile of one view: the per-column / per-row taps are built once in shared memory
// (80 threads), then the two passes run per channel (the image is 3.2 MB: L2 / L1 resident).
#include "common.cuh"

#include <algorithm>

namespace dvt {

constexpr int VW_TX = 64, VW_TY = 16, VW_ROWS_PER_THREAD = 4;  // block = (64, 4)

__device__ __forceinline__ float vw_cubic(float x) {  // ATen HelperInterpCubic::aa_filter, A = -0.5
  x = fabsf(x);
  if (x < 1.0f) return ((1.5f * x - 2.5f) * x) * x + 1.0f;       // ((A + 2) x - (A + 3)) x x + 1
  if (x < 2.0f) return ((-0.5f * x + 2.5f) * x - 4.0f) * x + 2.0f;  // ((A x - 5 A) x + 8 A) x - 4 A
  return 0.0f;
}

// taps of output index `o` along an axis of `in` source pixels resampled to `out`
template <int KMAX>
__device__ __forceinline__ void vw_axis_taps(int o, int in, int out, float* w, int* mn, int* sz) {
  const float scale = (float)in / (float)out;
  const float support = scale >= 1.0f ? 2.0f * scale : 2.0f;
  const float invscale = scale >= 1.0f ? 1.0f / scale : 1.0f;
  const float center = scale * ((float)o + 0.5f);
  const int xmin = max((int)(center - support + 0.5f), 0);
  const int xsize = min(min((int)(center + support + 0.5f), in) - xmin, KMAX);
  float total = 0.f;
  for (int j = 0; j < xsize; ++j) {
    const float wj = vw_cubic(((float)(j + xmin) - center + 0.5f) * invscale);
    w[j] = wj;
    total += wj;
  }
  for (int j = 0; j < xsize; ++j) w[j] = total != 0.f ? w[j] / total : w[j];
  *mn = xmin;
  *sz = xsize;
}

// Separable: per channel the CTA first runs the horizontal pass of the source rows its 16 output rows need (all taps of
// a row-column pair are read once) into shared memory, then the vertical pass out of shared memory.  Same products
// and the same summation order as the two ATen passes: hsum_j = sum_i wx[i] * src[j][i], out = sum_j wy[j] * hsum_j.
// (Round 1 evaluated the 2-D tap loop per output pixel: 25 instead of ~9 multiply-adds per pixel at the stage-1 crop scales.)
template <int KMAX, typename OutT>
__global__ void __launch_bounds__(VW_TX * VW_TY / VW_ROWS_PER_THREAD)
view_crops_kernel(const float* __restrict__ img, int H, int W, const int* __restrict__ boxes, const int* __restrict__ flips,
                  OutT* __restrict__ out, int OH, int OW, int span_cap) {
  extern __shared__ float s_h[];  // [span_cap][VW_TX]: horizontal pass of the source rows of this tile
  __shared__ float s_wx[VW_TX][KMAX], s_wy[VW_TY][KMAX];
  __shared__ int s_xmn[VW_TX], s_xsz[VW_TX], s_ymn[VW_TY], s_ysz[VW_TY];
  const int v = blockIdx.z;
  const int top = boxes[4 * v], left = boxes[4 * v + 1], ch = boxes[4 * v + 2], cw = boxes[4 * v + 3];
  const bool flip = flips[v] != 0;
  const int tid = threadIdx.y * VW_TX + threadIdx.x;
  if (tid < VW_TX) {
    const int ox = blockIdx.x * VW_TX + tid;
    s_xsz[tid] = 0;
    if (ox < OW) vw_axis_taps<KMAX>(flip ? OW - 1 - ox : ox, cw, OW, s_wx[tid], &s_xmn[tid], &s_xsz[tid]);
  } else if (tid < VW_TX + VW_TY) {
    const int r = tid - VW_TX;
    const int oy = blockIdx.y * VW_TY + r;
    s_ysz[r] = 0;
    s_ymn[r] = 0;
    if (oy < OH) vw_axis_taps<KMAX>(oy, ch, OH, s_wy[r], &s_ymn[r], &s_ysz[r]);
  }
  __syncthreads();
  // source rows [row0, row1) of the crop that the tile's output rows read (taps start at non-decreasing rows)
  const int rows_valid = min(VW_TY, OH - blockIdx.y * VW_TY);
  const int row0 = s_ymn[0];
  int row1 = row0;
  for (int r = 0; r < rows_valid; ++r) row1 = max(row1, s_ymn[r] + s_ysz[r]);
  const int span = min(row1 - row0, span_cap);  // (span_cap is sized by the host for the widest crop: never binding)
  const int ox = blockIdx.x * VW_TX + threadIdx.x;
  const int xmn = left + s_xmn[threadIdx.x], xsz = s_xsz[threadIdx.x];
  const float* wx = s_wx[threadIdx.x];
#pragma unroll 1
  for (int c = 0; c < 3; ++c) {
    const float* src = img + ((size_t)c * H + top + row0) * W + xmn;
    for (int j = threadIdx.y; j < span; j += VW_TY / VW_ROWS_PER_THREAD) {
      float hsum = 0.f;
      for (int i = 0; i < xsz; ++i) hsum += wx[i] * __ldg(src + (size_t)j * W + i);
      s_h[j * VW_TX + threadIdx.x] = hsum;
    }
    __syncthreads();
    if (ox < OW) {
#pragma unroll
      for (int rr = 0; rr < VW_ROWS_PER_THREAD; ++rr) {
        const int r = threadIdx.y + rr * (VW_TY / VW_ROWS_PER_THREAD);
        const int oy = blockIdx.y * VW_TY + r;
        if (oy >= OH) continue;
        const float* wy = s_wy[r];
        const float* hcol = s_h + (s_ymn[r] - row0) * VW_TX + threadIdx.x;
        const int ysz = min(s_ysz[r], span - (s_ymn[r] - row0));
        float acc = 0.f;
        for (int j = 0; j < ysz; ++j) acc += wy[j] * hcol[j * VW_TX];
        const size_t o = (((size_t)v * 3 + c) * OH + oy) * OW + ox;
        if constexpr (sizeof(OutT) == 2) out[o] = __float2bfloat16_rn(acc);
        else out[o] = acc;
      }
    }
    __syncthreads();
  }
}

// torch.linspace(start, end, steps)[idx] in fp32 (symmetric form of ATen's RangeFactories linspace kernel)
__device__ __forceinline__ float vw_linspace(float start, float end, int steps, int idx) {
  if (steps <= 1) return start;
  const float step = (end - start) / (float)(steps - 1);
  return idx < steps / 2 ? start + step * (float)idx : end - step * (float)(steps - idx - 1);
}

// coords[v, y, x] = (x coordinate, y coordinate) of patch (y, x) of view v inside the image, in [0, 1]
// (transform.py:55-73: linspace over the crop extent; on a flip x -> (x_max - x) + x_min)
__global__ void view_coords_kernel(const int* __restrict__ boxes, const int* __restrict__ flips, int H, int W, int V, int hp,
                                   int wp, float* __restrict__ coords) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= V * hp * wp) return;
  const int x = t % wp, y = (t / wp) % hp, v = t / (wp * hp);
  const int top = boxes[4 * v], left = boxes[4 * v + 1], ch = boxes[4 * v + 2], cw = boxes[4 * v + 3];
  const double ni = (double)top / (double)H, nj = (double)left / (double)W;
  const double nh = (double)ch / (double)H, nw = (double)cw / (double)W;
  const float ys = (float)ni, ye = (float)(ni + nh), xs = (float)nj, xe = (float)(nj + nw);
  float cx = vw_linspace(xs, xe, wp, x);
  const float cy = vw_linspace(ys, ye, hp, y);
  if (flips[v]) cx = (vw_linspace(xs, xe, wp, wp - 1) - cx) + vw_linspace(xs, xe, wp, 0);
  coords[2 * (size_t)t] = cx;
  coords[2 * (size_t)t + 1] = cy;
}

template <int KMAX>
static int launch_view_crops_k(const float* img, int H, int W, const int* boxes, const int* flips, int V, void* out,
                               bool out_bf16, int OH, int OW, int span_cap, cudaStream_t st) {
  const dim3 grid((OW + VW_TX - 1) / VW_TX, (OH + VW_TY - 1) / VW_TY, V), block(VW_TX, VW_TY / VW_ROWS_PER_THREAD);
  const size_t smem = (size_t)span_cap * VW_TX * sizeof(float);
  if (out_bf16) {
    auto k = view_crops_kernel<KMAX, __nv_bfloat16>;
    if (smem > 32 * 1024) DVT_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, block, smem, st>>>(img, H, W, boxes, flips, (__nv_bfloat16*)out, OH, OW, span_cap);
  } else {
    auto k = view_crops_kernel<KMAX, float>;
    if (smem > 32 * 1024) DVT_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, block, smem, st>>>(img, H, W, boxes, flips, (float*)out, OH, OW, span_cap);
  }
  return DVT_OK;
}

int view_crops(const float* image, int H, int W, const int* boxes_host, const int* flips_host, int V, void* out, bool out_bf16,
               int OH, int OW, float* coords_out, int hp, int wp, cudaStream_t st) {
  DVT_REQUIRE(image && boxes_host && flips_host && out && V > 0 && H > 0 && W > 0 && OH > 0 && OW > 0,
              "view_crops: bad arguments");
  int need = 0;  // taps per axis the widest crop needs
  int span_cap = 8;
  for (int v = 0; v < V; ++v) {
    const int top = boxes_host[4 * v], left = boxes_host[4 * v + 1], ch = boxes_host[4 * v + 2], cw = boxes_host[4 * v + 3];
    DVT_REQUIRE(ch > 0 && cw > 0 && top >= 0 && left >= 0 && top + ch <= H && left + cw <= W,
                "view_crops: box %d (top %d left %d h %d w %d) outside the %dx%d image", v, top, left, ch, cw, H, W);
    const float sy = (float)ch / (float)OH, sx = (float)cw / (float)OW;
    const float s = std::max(std::max(sx, sy), 1.0f);
    need = std::max(need, (int)(2.0f * 2.0f * s) + 2);
    // source rows under one 16-row output tile: its rows' centres span 15 * sy, plus the filter support on both sides
    span_cap = std::max(span_cap, (int)(15.0f * sy + 2.0f * 2.0f * std::max(sy, 1.0f)) + 4);
  }
  DVT_REQUIRE(need <= 32, "view_crops: down-scaling factor too large (%d taps per axis, at most 32)", need);
  int* dev = nullptr;  // boxes then flips, stream-ordered allocation: safe with calls in flight on other streams
  DVT_CUDA_OK(cudaMallocAsync((void**)&dev, (size_t)V * 5 * sizeof(int), st));
  DVT_CUDA_OK(cudaMemcpyAsync(dev, boxes_host, (size_t)V * 4 * sizeof(int), cudaMemcpyHostToDevice, st));
  DVT_CUDA_OK(cudaMemcpyAsync(dev + 4 * V, flips_host, (size_t)V * sizeof(int), cudaMemcpyHostToDevice, st));
  DVT_REQUIRE((size_t)span_cap * VW_TX * sizeof(float) <= 160 * 1024, "view_crops: %d source rows per tile exceed shared memory", span_cap);
  int lrc;
  if (need <= 8) lrc = launch_view_crops_k<8>(image, H, W, dev, dev + 4 * V, V, out, out_bf16, OH, OW, span_cap, st);
  else lrc = launch_view_crops_k<32>(image, H, W, dev, dev + 4 * V, V, out, out_bf16, OH, OW, span_cap, st);
  if (lrc) return lrc;
  DVT_CUDA_OK(cudaGetLastError());
  count_launch();
  if (coords_out) {
    DVT_REQUIRE(hp > 0 && wp > 0, "view_crops: bad coordinate grid %dx%d", hp, wp);
    const int n = V * hp * wp;
    view_coords_kernel<<<(n + 255) / 256, 256, 0, st>>>(dev, dev + 4 * V, H, W, V, hp, wp, coords_out);
    DVT_CUDA_OK(cudaGetLastError());
    count_launch();
  }
  DVT_CUDA_OK(cudaFreeAsync(dev, st));
  return DVT_OK;
}

}  // namespace dvt
