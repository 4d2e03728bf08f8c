// TN GEMM on 5th-gen tensor cores: C[M,N] = A[M,K] * B[N,K]^T with fused epilogues.
//
// One persistent CTA per SM, 10 warps:
//   warp 0      TMA producer   (A: 128 x 128B box, B: BN x 128B box, SWIZZLE_128B, 3-5 stage mbarrier ring)
//   warp 1      MMA issuer     (lane 0 issues tcgen05.mma M=128 N=BN K=16/8; accumulators in TMEM, 2 stages)
//   warps 2..9  epilogue       (tcgen05.ld 32x32b -> registers -> per-warp smem transpose -> coalesced stores;
//               two warps per TMEM lane quadrant take alternating 32-column chunks; 4 warps in the x3 kernels)
// K tails, M tails and N tails are handled by TMA zero-fill plus masking in the epilogue.
//
// Used for: patch-embed, QKV, attention out-proj, MLP fc1/fc2 (reference: timm VisionTransformer reached from
// dvt/models/vit_wrapper.py:136-143) and for the field / residual MLP forward+backward of the stage-1 fit
// (reference: dvt/models/neural_feature_field.py:40-44, dvt/models/offline_denoiser.py:40-46).
#include "gemm.cuh"

#include <algorithm>
#include <cstdlib>

namespace dvt {

namespace {

constexpr int BM = 128;
constexpr int KB_BYTES = 128;  // bytes of K per pipeline stage row (= one 128B swizzle atom)
// 8 epilogue warps per CTA (two per TMEM lane quadrant, alternating 32-column chunks: twice the ALU throughput and
// memory-level parallelism of one warp per scheduler).  The x3 (fit) kernels use 128x64 tiles: their problems are small
// (M = 2048), so narrower tiles mean more CTAs, half the MMA time per k-block and a one-chunk-per-warp epilogue.
// The 128x256 bf16 tiles (HP-1) get 12 epilogue warps: their GELU / residual epilogues are ALU- and latency-bound.
// The 128x128 x3 tiles keep three 64 KB stages: only four epilogue warps' transpose scratch fits beside them.
// (so does the experimental four-stage variant of the 128x64 x3 tile: DVT_GEMM_X3_STAGES=4)
constexpr int epi_warps(int bn, bool x3, int stages) { return (bn == 256 && !x3) ? 12 : (x3 && (bn == 128 || stages == 4)) ? 4 : 8; }
constexpr int SCR_PITCH = 36;  // floats; 16B-aligned rows, conflict-free for the access pattern below

template <int BN, int STAGES, bool X3 = false>
struct GemmSmem {
  static constexpr int EW = epi_warps(BN, X3, STAGES);
  static constexpr int THREADS = 32 * (2 + EW);
  static constexpr int A_BYTES = BM * KB_BYTES * (X3 ? 2 : 1);  // x3: hi plane then lo plane
  static constexpr int B_BYTES = BN * KB_BYTES * (X3 ? 2 : 1);
  static constexpr int SCR_BYTES = EW * 32 * SCR_PITCH * 4;
  static constexpr int OFF_A = 0;
  static constexpr int OFF_B = OFF_A + STAGES * A_BYTES;
  static constexpr int OFF_SCR = OFF_B + STAGES * B_BYTES;
  static constexpr int OFF_BAR = OFF_SCR + SCR_BYTES;
  static constexpr int NUM_BARS = 2 * STAGES + 4;
  static constexpr int OFF_TMEM = OFF_BAR + NUM_BARS * 8;
  static constexpr int TOTAL = OFF_TMEM + 16 + 1024;  // + alignment slack
};

// Ragged-N tail of one 32x32 chunk (last chunk of N = 385, 129, ...): rare, so its loops stay rolled (unrolled it
// made every kernel 130 KB; an out-of-line call was tried too and cost 30 % on the big GEMMs through ABI spills).
template <bool FIT>
__device__ __forceinline__ void epi_scalar_tail(const GemmEpi& e, const GemmShape& s, const float* scr, int lane, int m0,
                                             int n, bool has_k) {
#pragma unroll 1
  for (int jj = 0; jj < 8; ++jj) {
    const int i = (lane >> 3) + 4 * jj;
    const int m = m0 + 4 * jj;
    if (m >= s.M) continue;
#pragma unroll 1
    for (int q = 0; q < 4; ++q) {
      if (n + q >= s.N) break;
      const float x = scr[i * SCR_PITCH + (lane & 7) * 4 + q];
      epi_post1<FIT>(e, m, n + q, epi_pre<FIT>(e, m, n + q, has_k ? x : 0.0f));
    }
  }
}

// One 32-row x 32-column chunk of an accumulator tile, as read by tcgen05.ld 32x32b (thread = row): transpose through the
// warp's shared-memory scratch, then the fused epilogue in the coalesced layout (each lane: 4 consecutive columns of 8
// rows).  m0: global row of this lane's first row (tile row base + quadrant * 32 + lane / 8); n_base: global column of the
// chunk.  Shared by the single-CTA and the CTA-pair kernels.
template <bool FIT = false>
__device__ __forceinline__ void epi_chunk(const GemmEpi& e, const GemmShape& s, float* scr, int lane, const uint32_t (&r)[32],
                                          int n_base, int m0, bool has_k) {
  // ---- transpose through smem: afterwards each lane holds 4 consecutive columns of 8 rows.  No math on the
  // thread-per-row registers: everything that depends on the column (bias, LayerScale, ...) is loaded once per
  // chunk as a float4 in the coalesced layout below.
#pragma unroll
  for (int j = 0; j < 8; ++j)
    *reinterpret_cast<uint4*>(scr + lane * SCR_PITCH + 4 * j) = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
  __syncwarp();
  const int col4 = (lane & 7) * 4;
  const int n = n_base + col4;
  if (n + 3 < s.N) {
    // ---------------- vector path ----------------
    float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (e.bias) b4 = __ldg(reinterpret_cast<const float4*>(e.bias + n));
    float4 g4 = make_float4(1.f, 1.f, 1.f, 1.f);
    if (e.out_mode == OUT_F32_RESID && e.gamma) g4 = __ldg(reinterpret_cast<const float4*>(e.gamma + n));
    float4 xin[8];
    if (!FIT && e.out_mode == OUT_F32_RESID) {  // issue all residual loads before any store (memory-level parallelism)
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const int m = m0 + 4 * jj;
        if (m < s.M) xin[jj] = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(e.out) + (size_t)m * e.ldo + n);
      }
    }
    const bool plain = e.mask == nullptr && e.mask_f32 == nullptr && e.alpha == 1.0f && has_k;
    if (!FIT && plain && e.out_mode == OUT_BF16) {
      // ---- fast path: bias (+ GELU / ReLU) -> bf16.  The activation is chosen ONCE per chunk, the row loop is
      // branch-free (QKV and fc1+GELU, the two largest epilogues of the ViT forward).
      __nv_bfloat16* outp = reinterpret_cast<__nv_bfloat16*>(e.out);
      if (e.act == ACT_GELU) {
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          const int m = m0 + 4 * jj;
          float4 x = *reinterpret_cast<const float4*>(scr + ((lane >> 3) + 4 * jj) * SCR_PITCH + col4);
          uint2 pk;
          const float2 g01 = gelu_erf2(fadd2(make_float2(x.x, x.y), make_float2(b4.x, b4.y)));
          const float2 g23 = gelu_erf2(fadd2(make_float2(x.z, x.w), make_float2(b4.z, b4.w)));
          pk.x = pack_bf16x2(g01.x, g01.y);
          pk.y = pack_bf16x2(g23.x, g23.y);
          if (m < s.M) *reinterpret_cast<uint2*>(outp + (size_t)m * e.ldo + n) = pk;
        }
      } else if (e.act == ACT_RELU) {
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          const int m = m0 + 4 * jj;
          float4 x = *reinterpret_cast<const float4*>(scr + ((lane >> 3) + 4 * jj) * SCR_PITCH + col4);
          uint2 pk;
          pk.x = pack_bf16x2(fmaxf(x.x + b4.x, 0.0f), fmaxf(x.y + b4.y, 0.0f));
          pk.y = pack_bf16x2(fmaxf(x.z + b4.z, 0.0f), fmaxf(x.w + b4.w, 0.0f));
          if (m < s.M) *reinterpret_cast<uint2*>(outp + (size_t)m * e.ldo + n) = pk;
        }
      } else {   // bias only (QKV): two FADD2 and two packs per four outputs
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          const int m = m0 + 4 * jj;
          float4 x = *reinterpret_cast<const float4*>(scr + ((lane >> 3) + 4 * jj) * SCR_PITCH + col4);
          const float2 v01 = fadd2(make_float2(x.x, x.y), make_float2(b4.x, b4.y));
          const float2 v23 = fadd2(make_float2(x.z, x.w), make_float2(b4.z, b4.w));
          uint2 pk;
          pk.x = pack_bf16x2(v01.x, v01.y);
          pk.y = pack_bf16x2(v23.x, v23.y);
          if (m < s.M) *reinterpret_cast<uint2*>(outp + (size_t)m * e.ldo + n) = pk;
        }
      }
    } else if (!FIT && plain && e.out_mode == OUT_F32_RESID && e.act == ACT_NONE) {
      // ---- fast path: x += gamma * (acc + bias)  (attention out-proj, fc2) ----
      float* outp = reinterpret_cast<float*>(e.out);
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const int m = m0 + 4 * jj;
        const float4 x = *reinterpret_cast<const float4*>(scr + ((lane >> 3) + 4 * jj) * SCR_PITCH + col4);
        const float2 o01 = ffma2(make_float2(g4.x, g4.y), fadd2(make_float2(x.x, x.y), make_float2(b4.x, b4.y)),
                                 make_float2(xin[jj].x, xin[jj].y));
        const float2 o23 = ffma2(make_float2(g4.z, g4.w), fadd2(make_float2(x.z, x.w), make_float2(b4.z, b4.w)),
                                 make_float2(xin[jj].z, xin[jj].w));
        if (m < s.M) *reinterpret_cast<float4*>(outp + (size_t)m * e.ldo + n) = make_float4(o01.x, o01.y, o23.x, o23.y);
      }
    } else {
      // ---- generic path (fit epilogues: masks, split planes, atomics, remap) ----
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
      const int i = (lane >> 3) + 4 * jj;
      const int m = m0 + 4 * jj;
      float4 x = *reinterpret_cast<const float4*>(scr + i * SCR_PITCH + col4);
      if (m >= s.M) continue;
      if (!has_k) x = make_float4(0.f, 0.f, 0.f, 0.f);
      x.x += b4.x; x.y += b4.y; x.z += b4.z; x.w += b4.w;
      if (!FIT && e.act == ACT_GELU) {
        x.x = gelu_erf(x.x); x.y = gelu_erf(x.y); x.z = gelu_erf(x.z); x.w = gelu_erf(x.w);
      } else if (e.act == ACT_RELU) {
        x.x = fmaxf(x.x, 0.f); x.y = fmaxf(x.y, 0.f); x.z = fmaxf(x.z, 0.f); x.w = fmaxf(x.w, 0.f);
      }
      if (e.mask_f32) {
        const float4 h = *reinterpret_cast<const float4*>(e.mask_f32 + (size_t)m * e.ldmask + n);
        x.x = h.x > 0.f ? x.x : 0.f; x.y = h.y > 0.f ? x.y : 0.f; x.z = h.z > 0.f ? x.z : 0.f; x.w = h.w > 0.f ? x.w : 0.f;
      } else if (!FIT && e.mask) {
        const uint2 hb = *reinterpret_cast<const uint2*>(e.mask + (size_t)m * e.ldmask + n);
        const __nv_bfloat162 h01 = *reinterpret_cast<const __nv_bfloat162*>(&hb.x);
        const __nv_bfloat162 h23 = *reinterpret_cast<const __nv_bfloat162*>(&hb.y);
        if (e.mask_mode == 1) {  // GELU backward: the mask tensor is the forward pre-activation
          x.x *= gelu_grad(__low2float(h01)); x.y *= gelu_grad(__high2float(h01));
          x.z *= gelu_grad(__low2float(h23)); x.w *= gelu_grad(__high2float(h23));
        } else {
          x.x = __low2float(h01) > 0.f ? x.x : 0.f; x.y = __high2float(h01) > 0.f ? x.y : 0.f;
          x.z = __low2float(h23) > 0.f ? x.z : 0.f; x.w = __high2float(h23) > 0.f ? x.w : 0.f;
        }
      }
      if (e.alpha != 1.0f) { x.x *= e.alpha; x.y *= e.alpha; x.z *= e.alpha; x.w *= e.alpha; }
      if (!FIT && e.out_mode == OUT_F32_RESID) {
        float4 o = xin[jj];
        o.x = fmaf(g4.x, x.x, o.x); o.y = fmaf(g4.y, x.y, o.y); o.z = fmaf(g4.z, x.z, o.z); o.w = fmaf(g4.w, x.w, o.w);
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(e.out) + (size_t)m * e.ldo + n) = o;
      } else {
        epi_post4<FIT>(e, m, n, x);
      }
    }
    }
  } else {
    // ---------------- ragged N tail: scalar path ----------------
    epi_scalar_tail<FIT>(e, s, scr, lane, m0, n, has_k);
  }
  }

template <int BN, int STAGES, bool TF32, bool A_MN, bool B_MN, bool X3 = false>
__global__ void __launch_bounds__(GemmSmem<BN, STAGES, X3>::THREADS, 1)
gemm_tn_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, GemmShape s,
                  GemmEpi e) {
  static_assert(!X3 || TF32, "x3 is a TF32 mode");
  static_assert(X3 || !(TF32 && (A_MN || B_MN)), "plain TF32 is K-major only");
  using L = GemmSmem<BN, STAGES, X3>;
  constexpr int ELEM = TF32 ? 4 : 2;
  constexpr int BK = KB_BYTES / ELEM;         // elements of K per stage
  constexpr int UMMA_K_BYTES = 32;            // K=16 bf16 or K=8 tf32 per instruction
  constexpr int MMAS_PER_STAGE = KB_BYTES / UMMA_K_BYTES;
  // x3 with a K-major B: the hi and lo planes of a B tile are adjacent [BN x 128 B] blocks, i.e. ONE K-major tile of 2 BN
  // rows.  A_hi . [B_hi | B_lo] is therefore a single MMA of width 2 BN whose accumulator holds hi.hi in columns [0, BN) and
  // hi.lo in [BN, 2 BN); A_lo . B_hi accumulates into the first half and the epilogue adds the halves.  Two instructions
  // per k-step instead of three: a 128 x N x 8 TF32 MMA costs ~78 clk for any N <= 128 (tools/gemm_timeline.py), so the
  // k-block gets a quarter cheaper.  An MN-major B tile is made of 32-column atoms ([32 k-rows][32 mn] = 4 KB each); for the
  // same trick its atoms are loaded plane by plane -- [plane][atom] instead of [atom][plane] -- so that the lo atoms
  // continue the hi atoms at the same 4 KB stride (one more TMA instruction per atom, same bytes).
  constexpr bool CAT = X3 && 4 * BN <= 512;
  constexpr int ACC_W = CAT ? 2 * BN : BN;  // TMEM columns of one accumulator stage
  constexpr uint32_t TMEM_COLS = (2 * ACC_W <= 32) ? 32 : (2 * ACC_W <= 64) ? 64 : (2 * ACC_W <= 128) ? 128
                                 : (2 * ACC_W <= 256) ? 256 : 512;
  static_assert(2 * ACC_W <= 512, "two accumulator stages must fit TMEM");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem + L::OFF_A;
  uint8_t* sB = smem + L::OFF_B;
  float* scr_all = reinterpret_cast<float*>(smem + L::OFF_SCR);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* tfull = bars + 2 * STAGES;
  uint64_t* tempty = bars + 2 * STAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::OFF_TMEM);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  auto stampt = [&](int slot) {
    if (e.debug_ts && blockIdx.x == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      e.debug_ts[slot] = t;
    }
  };
  if (threadIdx.x == 0) stampt(0);  // kernel entry
  if (threadIdx.x == 0 && e.debug_ts && blockIdx.x == 0) e.debug_ts[14] = (unsigned long long)clock64();

  const int tiles_m = (s.M + BM - 1) / BM;
  const int tiles_n = (s.N + BN - 1) / BN;
  const int num_tiles = tiles_m * tiles_n * s.splits;
  const int kb_total = (s.K + BK - 1) / BK;
  const int kb_per_split = (kb_total + s.splits - 1) / s.splits;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], L::EW * 32);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) stampt(1);  // setup done (barriers, TMEM)
  // PDL: everything above overlapped the tail of the previous kernel; its results are needed from here on
  pdl_wait();
  pdl_trigger();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int split = t % s.splits;
        const int mn = t / s.splits;
        const int tn = mn % tiles_n;
        const int tm = mn / tiles_n;
        const int kb0 = split * kb_per_split;
        const int kb1 = min(kb_total, kb0 + kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1, 1);
          mbar_expect_tx(&full[stage], L::A_BYTES + L::B_BYTES);
          if (X3) {
            // fp32 hi/lo planes: 3-D maps (inner, rows, plane); one box brings both planes of a tile / atom
            if (A_MN) {
#pragma unroll
              for (int a = 0; a < BM / 32; ++a)
                tma_load_3d(sA + stage * L::A_BYTES + a * 8192, &tmA, &full[stage], tm * BM + a * 32, kb * BK, 0);
            } else {
              tma_load_3d(sA + stage * L::A_BYTES, &tmA, &full[stage], kb * BK, tm * BM, 0);
            }
            if (B_MN) {
              // (tensor map box = ONE plane of an atom: launch_gemm_tn)
#pragma unroll
              for (int pl = 0; pl < 2; ++pl)
#pragma unroll
                for (int a = 0; a < BN / 32; ++a)
                  tma_load_3d(sB + stage * L::B_BYTES + (pl * (BN / 32) + a) * 4096, &tmB, &full[stage], tn * BN + a * 32, kb * BK, pl);
            } else {
              tma_load_3d(sB + stage * L::B_BYTES, &tmB, &full[stage], kb * BK, tn * BN, 0);
            }
          } else {
            if (A_MN) {
              // stored [K, M]: boxes of 64 (M, contiguous) x BK (K rows) = one MN-major swizzle-atom column each
#pragma unroll
              for (int a = 0; a < BM / 64; ++a)
                tma_load_2d(sA + stage * L::A_BYTES + a * (BK * 128), &tmA, &full[stage], tm * BM + a * 64, kb * BK);
            } else {
              tma_load_2d(sA + stage * L::A_BYTES, &tmA, &full[stage], kb * BK, tm * BM);
            }
            if (B_MN) {
#pragma unroll
              for (int a = 0; a < BN / 64; ++a)
                tma_load_2d(sB + stage * L::B_BYTES + a * (BK * 128), &tmB, &full[stage], tn * BN + a * 64, kb * BK);
            } else {
              tma_load_2d(sB + stage * L::B_BYTES, &tmB, &full[stage], kb * BK, tn * BN);
            }
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(TF32 ? 2u : 1u, BM, BN, A_MN ? 1u : 0u, B_MN ? 1u : 0u);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int split = t % s.splits;
        const int kb0 = split * kb_per_split;
        const int kb1 = min(kb_total, kb0 + kb_per_split);
        mbar_wait(&tempty[as], aphase ^ 1, 2);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * ACC_W;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full[stage], phase, 3);
          tc_fence_after();
          if (kb == kb0) stampt(2);  // first operands landed
          // K-major: rows of 128 B (one swizzle atom of K), 8-row groups 1024 B apart; K advances 32 B per MMA.
          // MN-major (bf16): tile = [MN/64 atoms][BK k-rows][64 mn]; atoms BK*128 B apart (LBO), 8-k groups 1024 B
          // apart (SBO); K advances 16 rows = 2048 B per MMA.
          // x3 (fp32 hi/lo planes): K-major tile = [plane][rows][128 B]; MN-major A tile = [MN/32 atoms][plane][32 k-rows]
          // [32 mn] (atoms 8192 B apart, lo plane +4096 B), MN-major B tile = [plane][MN/32 atoms][32 k-rows][32 mn] (atoms
          // 4096 B apart, lo plane after the hi atoms); K advances 8 rows = 1024 B per MMA.  MN-major TF32 operands
          // must use the "128B swizzle with 32B atoms" layout (descriptor layout type 1, TMA SWIZZLE_128B_ATOM_32B):
          // the swizzle pattern repeats every 4 K-rows, so the stride between K groups (SBO) is 512 B.
          const uint32_t a_base = smem_u32(sA + stage * L::A_BYTES), b_base = smem_u32(sB + stage * L::B_BYTES);
          if (X3) {
            const uint32_t a_lo = a_base + (A_MN ? 4096 : BM * KB_BYTES), b_lo = b_base + (B_MN ? (BN / 32) * 4096 : BN * KB_BYTES);
            const uint32_t lbo_a = A_MN ? 8192 : 0, lbo_b = B_MN ? 4096 : 0;
#pragma unroll
            for (int k = 0; k < MMAS_PER_STAGE; ++k) {
              const uint32_t ka = A_MN ? k * 1024 : k * UMMA_K_BYTES, kbb = B_MN ? k * 1024 : k * UMMA_K_BYTES;
              constexpr uint32_t sbo_a = A_MN ? 512 : 1024, sbo_b = B_MN ? 512 : 1024;
              constexpr uint32_t lt_a = A_MN ? 1 : 2, lt_b = B_MN ? 1 : 2;
              const uint64_t dah = make_smem_desc(a_base + ka, lbo_a, sbo_a, lt_a), dal = make_smem_desc(a_lo + ka, lbo_a, sbo_a, lt_a);
              const uint64_t dbh = make_smem_desc(b_base + kbb, lbo_b, sbo_b, lt_b), dbl = make_smem_desc(b_lo + kbb, lbo_b, sbo_b, lt_b);
              const uint32_t acc0 = (kb > kb0 || k > 0) ? 1u : 0u;
              if (CAT && s.x3 == 1) {
                constexpr uint32_t idesc_cat = make_idesc(2u, BM, 2 * BN, A_MN ? 1u : 0u, B_MN ? 1u : 0u);
                umma_tf32(d_tmem, dah, dbh, idesc_cat, acc0);  // [hi.hi | hi.lo]: the descriptor at B_hi spans both planes
                umma_tf32(d_tmem, dal, dbh, idesc, 1u);        // + lo.hi
              } else if (s.x3 == 1) {
                umma_tf32(d_tmem, dal, dbh, idesc, acc0);  // small terms first
                umma_tf32(d_tmem, dah, dbl, idesc, 1u);
                umma_tf32(d_tmem, dah, dbh, idesc, 1u);
              } else {  // x3 == 2: plain TF32 product of the hi parts (same operand layout, a third of the MMA work)
                umma_tf32(d_tmem, dah, dbh, idesc, acc0);
              }
            }
          } else {
            const uint64_t da = make_smem_desc(a_base, A_MN ? BK * 128 : 0, 1024, 2);
            const uint64_t db = make_smem_desc(b_base, B_MN ? BK * 128 : 0, 1024, 2);
#pragma unroll
            for (int k = 0; k < MMAS_PER_STAGE; ++k) {
              const uint64_t adv_a = (uint64_t)((A_MN ? k * 2048 : k * UMMA_K_BYTES) >> 4);
              const uint64_t adv_b = (uint64_t)((B_MN ? k * 2048 : k * UMMA_K_BYTES) >> 4);
              if (TF32) umma_tf32(d_tmem, da + adv_a, db + adv_b, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
              else umma_f16(d_tmem, da + adv_a, db + adv_b, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            }
          }
          umma_commit(&empty[stage]);  // frees the smem slot once these MMAs have read it
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tfull[as]);  // accumulator complete
        stampt(3);                // all MMAs issued
        if (++as == 2) {
          as = 0;
          aphase ^= 1;
        }
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int ew = warp - 2;
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access
    constexpr int CSTEP = L::EW / 4;       // warps sharing a quadrant
    const int c_first = ew >> 2;           // ... take alternating 32-column chunks
    float* scr = scr_all + ew * 32 * SCR_PITCH;
    int as = 0;
    uint32_t aphase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int split = t % s.splits;
      const int mn = t / s.splits;
      const int tn = mn % tiles_n;
      const int tm = mn / tiles_n;
      const int kb0 = split * kb_per_split;
      const bool has_k = kb0 < kb_total;
      mbar_wait(&tfull[as], aphase, 4);
      tc_fence_after();
      if (ew == 0 && lane == 0) stampt(4);  // accumulator ready, epilogue starts
      const uint32_t taddr_row = tmem_base + ((uint32_t)(quad * 32) << 16) + as * ACC_W;
#pragma unroll 1
      for (int c = c_first; c < BN / 32; c += CSTEP) {
        uint32_t r[32];
        tmem_ld_32x32(taddr_row + c * 32, r);
        if (CAT && s.x3 == 1) {
          uint32_t r2[32];
          tmem_ld_32x32(taddr_row + BN + c * 32, r2);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) + __uint_as_float(r2[i]));
        } else {
          tmem_ld_wait();
        }
        if (c + CSTEP >= BN / 32) {  // this warp's last read of the accumulator stage
          tc_fence_before();
          mbar_arrive(&tempty[as]);
        }
        const int n_base = tn * BN + c * 32;
        if (n_base >= s.N) continue;  // whole chunk out of range (warp-uniform)
        epi_chunk<X3>(e, s, scr, lane, r, n_base, tm * BM + quad * 32 + (lane >> 3), has_k);
        __syncwarp();
        if (ew == 0 && lane == 0) stampt(8 + (c & 7));  // chunk done (profiling aid)
      }
      if (++as == 2) {
        as = 0;
        aphase ^= 1;
      }
    }
  }

  if (warp == 2 && lane == 0) stampt(5);  // first epilogue warp done
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
  if (threadIdx.x == 0) stampt(6);  // kernel exit
  if (threadIdx.x == 0 && e.debug_ts && blockIdx.x == 0) e.debug_ts[15] = (unsigned long long)clock64();
}

// ----------------------------------------------------------------------------------------------------
// SIMT debug GEMM (fp32 FMA, 16x16 tiles).  Same epilogue semantics; selected with DVT_GEMM_IMPL=simt or the
// `impl` argument.  It exists so that a broken tensor-core path can be told apart from a broken caller in
// one GPU session; it is never the default.
// ----------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ float ld_as_float(const T* p);
template <>
__device__ __forceinline__ float ld_as_float<float>(const float* p) {
  return *p;
}
template <>
__device__ __forceinline__ float ld_as_float<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __bfloat162float(*p);
}

template <typename T>
__global__ void gemm_tn_simt_kernel(const T* __restrict__ A, int lda, const T* __restrict__ B, int ldb, GemmShape s,
                                    GemmEpi e) {
  __shared__ float sa[16][17];
  __shared__ float sb[16][17];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int m = blockIdx.y * 16 + ty;
  const int n = blockIdx.x * 16 + tx;
  float acc = 0.0f;
  for (int k0 = 0; k0 < s.K; k0 += 16) {
    int ka = k0 + tx;
    const T* ap = s.a_mn ? A + (size_t)ka * lda + m : A + (size_t)m * lda + ka;
    sa[ty][tx] = (m < s.M && ka < s.K) ? ld_as_float(ap) + (s.x3 ? ld_as_float(ap + s.plane_a) : 0.0f) : 0.0f;
    int nb = blockIdx.x * 16 + ty;
    const T* bp = s.b_mn ? B + (size_t)ka * ldb + nb : B + (size_t)nb * ldb + ka;
    sb[ty][tx] = (nb < s.N && ka < s.K) ? ld_as_float(bp) + (s.x3 ? ld_as_float(bp + s.plane_b) : 0.0f) : 0.0f;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) acc = fmaf(sa[ty][k], sb[tx][k], acc);
    __syncthreads();
  }
  if (m < s.M && n < s.N) {
    epi_post1(e, m, n, epi_pre(e, m, n, acc));
  }
}

template <int BN, int STAGES, bool TF32, bool A_MN, bool B_MN, bool X3 = false>
int launch_tc(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmShape& s, const GemmEpi& e,
              cudaStream_t stream) {
  using L = GemmSmem<BN, STAGES, X3>;
  auto kern = gemm_tn_tc_kernel<BN, STAGES, TF32, A_MN, B_MN, X3>;
  const int tiles = ((s.M + BM - 1) / BM) * ((s.N + BN - 1) / BN) * s.splits;
  int grid = tiles < num_sms() ? tiles : num_sms();
  // DVT_GEMM_TILES_PER_CTA=n (n > 0) bounds the tiles one CTA works through, i.e. launches more, shorter-lived CTAs than
  // SMs: a persistent CTA keeps its SM for the whole kernel, which starves concurrent high-priority streams (the fit
  // running beside the next image's ViT forwards); short-lived CTAs hand SMs over every few microseconds instead.
  static int tiles_per_cta = -1;
  if (tiles_per_cta < 0) {
    const char* v = getenv("DVT_GEMM_TILES_PER_CTA");
    tiles_per_cta = v ? atoi(v) : 0;
  }
  if (!X3 && tiles_per_cta > 0) grid = std::max(grid, (tiles + tiles_per_cta - 1) / tiles_per_cta);
  DVT_CUDA_OK(launch_kx(LaunchOpt{s.pdl != 0, s.prio_drop}, kern, dim3(grid), dim3(L::THREADS), (size_t)L::TOTAL, stream, tmA, tmB, s, e));
  count_launch();
  DVT_CUDA_OK(cudaGetLastError());
  return DVT_OK;
}

template <int BN, int STAGES, bool TF32, bool A_MN, bool B_MN, bool X3 = false>
int prep_one() {
  DVT_CUDA_OK(cudaFuncSetAttribute(gemm_tn_tc_kernel<BN, STAGES, TF32, A_MN, B_MN, X3>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, GemmSmem<BN, STAGES, X3>::TOTAL));
  return DVT_OK;
}

}  // namespace

// Opts every instantiation into its dynamic shared memory size.  Called once, outside any stream capture.
int gemm_prepare() {
  static bool done = false;
  if (done) return DVT_OK;
  int rc;
  if ((rc = prep_one<256, 3, true, false, false>())) return rc;
  if ((rc = prep_one<128, 5, true, false, false>())) return rc;
  if ((rc = prep_one<256, 3, false, true, true>())) return rc;
  if ((rc = prep_one<128, 5, false, true, true>())) return rc;
  if ((rc = prep_one<256, 3, false, false, true>())) return rc;
  if ((rc = prep_one<128, 5, false, false, true>())) return rc;
  if ((rc = prep_one<256, 3, false, false, false>())) return rc;
  if ((rc = prep_one<128, 5, false, false, false>())) return rc;
  if ((rc = prep_one<64, 3, true, false, false, true>())) return rc;
  if ((rc = prep_one<64, 3, true, false, true, true>())) return rc;
  if ((rc = prep_one<64, 3, true, true, true, true>())) return rc;
  if ((rc = prep_one<64, 4, true, false, false, true>())) return rc;
  if ((rc = prep_one<64, 4, true, false, true, true>())) return rc;
  if ((rc = prep_one<64, 4, true, true, true, true>())) return rc;
  if ((rc = prep_one<128, 3, true, false, false, true>())) return rc;
  if ((rc = prep_one<128, 3, true, false, true, true>())) return rc;
  if ((rc = prep_one<128, 3, true, true, true, true>())) return rc;
  done = true;
  return DVT_OK;
}

// Tile width of the 3xTF32 kernels.  Measured on the fit's shapes (tools/gemm_timeline.py): a k-block costs ~0.66 us with
// 12 MMAs whether the tile is 128 x 64 or 128 x 128 (each 128 x N x 8 TF32 instruction takes ~78 clk for N <= 128), so the
// wide tile halves the tensor time and moves a third less operand data through L2 per flop; the narrow tile gives twice
// the CTAs and the shorter epilogue.  The caller chooses per call (GemmShape::x3_wide_min_n; 0 = always 128 x 64).
int gemm_x3_tile_n(int N, int wide_min_n) { return (wide_min_n > 0 && N >= wide_min_n) ? 128 : 64; }

int default_gemm_impl() {
  static int impl = -1;
  if (impl < 0) {
    const char* v = getenv("DVT_GEMM_IMPL");
    impl = (v && v[0] == 's') ? GEMM_SIMT_DEBUG : GEMM_TCGEN05;
  }
  return impl;
}

int launch_gemm_tn(const void* A, int lda, const void* B, int ldb, TmapDtype dtype, const GemmShape& shape,
                   const GemmEpi& epi_in, cudaStream_t stream, int impl) {
  if (impl < 0) impl = default_gemm_impl();
  GemmShape s = shape;
  GemmEpi epi = epi_in;
  if (s.splits < 1) s.splits = 1;
  DVT_REQUIRE(s.M > 0 && s.N > 0 && s.K > 0, "gemm: empty shape M=%d N=%d K=%d", s.M, s.N, s.K);
  DVT_REQUIRE(s.splits == 1 || epi.out_mode == OUT_F32_ATOMIC, "gemm: split-K needs OUT_F32_ATOMIC");
  DVT_REQUIRE(epi.out == nullptr || epi.ldo % 4 == 0, "gemm: ldo must be a multiple of 4 (got %d)", epi.ldo);
  DVT_REQUIRE(epi.last_col_out == nullptr || epi.out_mode == OUT_F32_ATOMIC, "gemm: last_col_out needs OUT_F32_ATOMIC");
  epi.last_col_n = epi.last_col_out ? s.N - 1 : -1;
  // the vector post-stage must not straddle the redirected column: keep N-1 in a scalar tail
  DVT_REQUIRE(epi.last_col_out == nullptr || (s.N - 1) % 4 == 0, "gemm: last_col_out needs (N-1) %% 4 == 0 (N=%d)", s.N);

  if (impl == GEMM_SIMT_DEBUG) {
    dim3 grid((s.N + 15) / 16, (s.M + 15) / 16), block(16, 16);
    if (dtype == TMAP_BF16)
      gemm_tn_simt_kernel<__nv_bfloat16><<<grid, block, 0, stream>>>(
          reinterpret_cast<const __nv_bfloat16*>(A), lda, reinterpret_cast<const __nv_bfloat16*>(B), ldb, s, epi);
    else
      gemm_tn_simt_kernel<float><<<grid, block, 0, stream>>>(reinterpret_cast<const float*>(A), lda,
                                                             reinterpret_cast<const float*>(B), ldb, s, epi);
    DVT_CUDA_OK(cudaGetLastError());
    count_launch();
    return DVT_OK;
  }

  {
    int prc = gemm_prepare();
    if (prc) return prc;
  }
  const int elem = dtype == TMAP_BF16 ? 2 : 4;
  const int bk = KB_BYTES / elem;
  if (s.x3) {
    DVT_REQUIRE(dtype == TMAP_F32, "gemm: x3 needs fp32 operands");
    DVT_REQUIRE(!(s.a_mn && !s.b_mn), "gemm: A MN-major with B K-major is not instantiated");
    DVT_REQUIRE((lda * 4) % 16 == 0 && (ldb * 4) % 16 == 0 && (s.plane_a * 4) % 16 == 0 && (s.plane_b * 4) % 16 == 0,
                "gemm: x3 pitches must be multiples of 16 bytes");
    CUtensorMap tA, tB;
    int rc3;
    const int bn3 = gemm_x3_tile_n(s.N, s.x3_wide_min_n);
    if (s.a_mn) rc3 = make_tmap_3d(&tA, A, TMAP_F32, (uint64_t)s.M, (uint64_t)s.K, 2, (uint64_t)lda * 4, s.plane_a * 4, 32, 32, 2, true);
    else rc3 = make_tmap_3d(&tA, A, TMAP_F32, (uint64_t)s.K, (uint64_t)s.M, 2, (uint64_t)lda * 4, s.plane_a * 4, 32, BM, 2);
    if (rc3) return rc3;
    if (s.b_mn) rc3 = make_tmap_3d(&tB, B, TMAP_F32, (uint64_t)s.N, (uint64_t)s.K, 2, (uint64_t)ldb * 4, s.plane_b * 4, 32, 32, 1, true);
    else rc3 = make_tmap_3d(&tB, B, TMAP_F32, (uint64_t)s.K, (uint64_t)s.N, 2, (uint64_t)ldb * 4, s.plane_b * 4, 32, bn3, 2);
    if (rc3) return rc3;
    if (bn3 == 128) {
      if (s.a_mn) return launch_tc<128, 3, true, true, true, true>(tA, tB, s, epi, stream);
      if (s.b_mn) return launch_tc<128, 3, true, false, true, true>(tA, tB, s, epi, stream);
      return launch_tc<128, 3, true, false, false, true>(tA, tB, s, epi, stream);
    }
    static const bool four = [] { const char* e = getenv("DVT_GEMM_X3_STAGES"); return e && e[0] == '4'; }();
    if (four) {  // experiment: a fourth 48 KB stage (and four epilogue warps) for the 128 x 64 tile
      if (s.a_mn) return launch_tc<64, 4, true, true, true, true>(tA, tB, s, epi, stream);
      if (s.b_mn) return launch_tc<64, 4, true, false, true, true>(tA, tB, s, epi, stream);
      return launch_tc<64, 4, true, false, false, true>(tA, tB, s, epi, stream);
    }
    if (s.a_mn) return launch_tc<64, 3, true, true, true, true>(tA, tB, s, epi, stream);
    if (s.b_mn) return launch_tc<64, 3, true, false, true, true>(tA, tB, s, epi, stream);
    return launch_tc<64, 3, true, false, false, true>(tA, tB, s, epi, stream);
  }
  DVT_REQUIRE(dtype == TMAP_BF16 || (!s.a_mn && !s.b_mn), "gemm: MN-major operands are implemented for bf16 only");
  DVT_REQUIRE(!(s.a_mn && !s.b_mn), "gemm: A MN-major with B K-major is not instantiated");
  DVT_REQUIRE((lda * elem) % 16 == 0 && (ldb * elem) % 16 == 0,
              "gemm: row pitch must be a multiple of 16 bytes (lda=%d ldb=%d elem=%d)", lda, ldb, elem);
  DVT_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0,
              "gemm: operands must be 16-byte aligned");
  // wide tiles when N is large enough to fill them; narrow ones for the small fit GEMMs
  const bool wide = s.N >= 256 && (s.N % 256 == 0 || s.N > 1024);
  if (dtype == TMAP_BF16 && !s.a_mn && !s.b_mn && wide && s.splits == 1 && s.M >= 256 && epi.out_mode != OUT_F32_ATOMIC &&
      impl != GEMM_TCGEN05_1CTA && gemm_cg2_enabled())
    return launch_gemm_cg2(A, lda, B, ldb, s, epi, stream);  // 256 x 256 tiles on CTA pairs (tcgen05 cta_group::2)
  CUtensorMap tmA, tmB;
  int rc;
  if (s.a_mn) rc = make_tmap_2d(&tmA, A, dtype, (uint64_t)s.K, (uint64_t)s.M, (uint64_t)lda * elem, bk, 64);
  else rc = make_tmap_2d(&tmA, A, dtype, (uint64_t)s.M, (uint64_t)s.K, (uint64_t)lda * elem, BM, bk);
  if (rc) return rc;
  if (s.b_mn) rc = make_tmap_2d(&tmB, B, dtype, (uint64_t)s.K, (uint64_t)s.N, (uint64_t)ldb * elem, bk, 64);
  else rc = make_tmap_2d(&tmB, B, dtype, (uint64_t)s.N, (uint64_t)s.K, (uint64_t)ldb * elem, wide ? 256 : 128, bk);
  if (rc) return rc;
  if (dtype == TMAP_F32)
    return wide ? launch_tc<256, 3, true, false, false>(tmA, tmB, s, epi, stream)
                : launch_tc<128, 5, true, false, false>(tmA, tmB, s, epi, stream);
  if (s.a_mn)
    return wide ? launch_tc<256, 3, false, true, true>(tmA, tmB, s, epi, stream)
                : launch_tc<128, 5, false, true, true>(tmA, tmB, s, epi, stream);
  if (s.b_mn)
    return wide ? launch_tc<256, 3, false, false, true>(tmA, tmB, s, epi, stream)
                : launch_tc<128, 5, false, false, true>(tmA, tmB, s, epi, stream);
  return wide ? launch_tc<256, 3, false, false, false>(tmA, tmB, s, epi, stream)
              : launch_tc<128, 5, false, false, false>(tmA, tmB, s, epi, stream);
}

}  // namespace dvt
