// TN GEMM on 5th-gen tensor cores: C[M,N] = A[M,K] * B[N,K]^T with fused epilogues.
//
// One persistent CTA per SM, 6 warps:
//   warp 0      TMA producer   (A: 128 x 128B box, B: BN x 128B box, SWIZZLE_128B, 4-stage mbarrier ring)
//   warp 1      MMA issuer     (lane 0 issues tcgen05.mma M=128 N=BN K=16/8; accumulators in TMEM, 2 stages)
//   warps 2..5  epilogue       (tcgen05.ld 32x32b -> registers -> per-warp smem transpose -> coalesced stores)
// K tails, M tails and N tails are handled by TMA zero-fill plus masking in the epilogue.
//
// Used for: patch-embed, QKV, attention out-proj, MLP fc1/fc2 (reference: timm VisionTransformer reached from
// dvt/models/vit_wrapper.py:136-143) and for the field / residual MLP forward+backward of the stage-1 fit
// (reference: dvt/models/neural_feature_field.py:40-44, dvt/models/offline_denoiser.py:40-46).
#include "gemm.cuh"

#include <cstdlib>

namespace dvt {

namespace {

constexpr int BM = 128;
constexpr int KB_BYTES = 128;  // bytes of K per pipeline stage row (= one 128B swizzle atom)
constexpr int NUM_EPI_WARPS = 4;
constexpr int GEMM_THREADS = 32 * (2 + NUM_EPI_WARPS);
constexpr int SCR_PITCH = 36;  // floats; 16B-aligned rows, conflict-free for the access pattern below

template <int BN, int STAGES>
struct GemmSmem {
  static constexpr int A_BYTES = BM * KB_BYTES;
  static constexpr int B_BYTES = BN * KB_BYTES;
  static constexpr int SCR_BYTES = NUM_EPI_WARPS * 32 * SCR_PITCH * 4;
  static constexpr int OFF_A = 0;
  static constexpr int OFF_B = OFF_A + STAGES * A_BYTES;
  static constexpr int OFF_SCR = OFF_B + STAGES * B_BYTES;
  static constexpr int OFF_BAR = OFF_SCR + SCR_BYTES;
  static constexpr int NUM_BARS = 2 * STAGES + 4;
  static constexpr int OFF_TMEM = OFF_BAR + NUM_BARS * 8;
  static constexpr int TOTAL = OFF_TMEM + 16 + 1024;  // + alignment slack
};

template <int BN, int STAGES, bool TF32>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tn_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, GemmShape s,
                  GemmEpi e) {
  using L = GemmSmem<BN, STAGES>;
  constexpr int ELEM = TF32 ? 4 : 2;
  constexpr int BK = KB_BYTES / ELEM;         // elements of K per stage
  constexpr int UMMA_K_BYTES = 32;            // K=16 bf16 or K=8 tf32 per instruction
  constexpr int MMAS_PER_STAGE = KB_BYTES / UMMA_K_BYTES;
  constexpr uint32_t TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128
                                 : (2 * BN <= 256) ? 256 : 512;
  static_assert(2 * BN <= 512, "two accumulator stages must fit TMEM");

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
      mbar_init(&tempty[i], NUM_EPI_WARPS * 32);
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
          tma_load_2d(sA + stage * L::A_BYTES, &tmA, &full[stage], kb * BK, tm * BM);
          tma_load_2d(sB + stage * L::B_BYTES, &tmB, &full[stage], kb * BK, tn * BN);
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
      constexpr uint32_t idesc = make_idesc(TF32 ? 2u : 1u, BM, BN, 0, 0);
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
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full[stage], phase, 3);
          tc_fence_after();
          const uint64_t da = make_smem_desc(smem_u32(sA + stage * L::A_BYTES), 0, 1024, 2);
          const uint64_t db = make_smem_desc(smem_u32(sB + stage * L::B_BYTES), 0, 1024, 2);
#pragma unroll
          for (int k = 0; k < MMAS_PER_STAGE; ++k) {
            const uint64_t adv = (uint64_t)((k * UMMA_K_BYTES) >> 4);
            if (TF32) umma_tf32(d_tmem, da + adv, db + adv, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            else umma_f16(d_tmem, da + adv, db + adv, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty[stage]);  // frees the smem slot once these MMAs have read it
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tfull[as]);  // accumulator complete
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
      const int m_thread = tm * BM + quad * 32 + lane;  // row owned in the TMEM layout
      const uint32_t taddr_row = tmem_base + ((uint32_t)(quad * 32) << 16) + as * BN;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(taddr_row + c * 32, r);
        tmem_ld_wait();
        if (c == BN / 32 - 1) {
          tc_fence_before();
          mbar_arrive(&tempty[as]);
        }
        const int n_base = tn * BN + c * 32;
        if (n_base >= s.N) continue;  // whole chunk out of range (warp-uniform)
        // ---- pre-stage: thread-per-row ----
        float v[32];
        const bool row_ok = m_thread < s.M;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float acc = has_k ? __uint_as_float(r[j]) : 0.0f;
          int n = n_base + j;
          v[j] = (row_ok && n < s.N) ? epi_pre(e, m_thread, n, acc) : 0.0f;
        }
        if (e.out_t) {
          if (row_ok) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              int n = n_base + j;
              if (n < s.N) e.out_t[(size_t)n * e.ldt + m_thread] = __float2bfloat16_rn(v[j]);
            }
          }
        }
        if (e.out == nullptr) continue;
        // ---- transpose through smem: afterwards each lane holds 4 consecutive columns of 8 rows ----
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<float4*>(scr + lane * SCR_PITCH + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        __syncwarp();
        const int col4 = (lane & 7) * 4;
        const int n = n_base + col4;
        const bool vec_ok = (n + 3 < s.N);
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          const int i = (lane >> 3) + 4 * jj;
          const int m = tm * BM + quad * 32 + i;
          float4 x = *reinterpret_cast<const float4*>(scr + i * SCR_PITCH + col4);
          if (m < s.M) {
            if (vec_ok) {
              epi_post4(e, m, n, x);
            } else {
              if (n + 0 < s.N) epi_post1(e, m, n + 0, x.x);
              if (n + 1 < s.N) epi_post1(e, m, n + 1, x.y);
              if (n + 2 < s.N) epi_post1(e, m, n + 2, x.z);
              if (n + 3 < s.N) epi_post1(e, m, n + 3, x.w);
            }
          }
        }
        __syncwarp();
      }
      if (++as == 2) {
        as = 0;
        aphase ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
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
    sa[ty][tx] = (m < s.M && ka < s.K) ? ld_as_float(A + (size_t)m * lda + ka) : 0.0f;
    int nb = blockIdx.x * 16 + ty;
    sb[ty][tx] = (nb < s.N && ka < s.K) ? ld_as_float(B + (size_t)nb * ldb + ka) : 0.0f;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) acc = fmaf(sa[ty][k], sb[tx][k], acc);
    __syncthreads();
  }
  if (m < s.M && n < s.N) {
    float v = epi_pre(e, m, n, acc);
    if (e.out_t) e.out_t[(size_t)n * e.ldt + m] = __float2bfloat16_rn(v);
    if (e.out) epi_post1(e, m, n, v);
  }
}

template <int BN, int STAGES, bool TF32>
int launch_tc(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmShape& s, const GemmEpi& e,
              cudaStream_t stream) {
  using L = GemmSmem<BN, STAGES>;
  auto kern = gemm_tn_tc_kernel<BN, STAGES, TF32>;
  static bool attr_set = false;
  if (!attr_set) {
    DVT_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    attr_set = true;
  }
  const int tiles = ((s.M + BM - 1) / BM) * ((s.N + BN - 1) / BN) * s.splits;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  kern<<<grid, GEMM_THREADS, L::TOTAL, stream>>>(tmA, tmB, s, e);
  DVT_CUDA_OK(cudaGetLastError());
  return DVT_OK;
}

}  // namespace

int default_gemm_impl() {
  static int impl = -1;
  if (impl < 0) {
    const char* v = getenv("DVT_GEMM_IMPL");
    impl = (v && v[0] == 's') ? GEMM_SIMT_DEBUG : GEMM_TCGEN05;
  }
  return impl;
}

int launch_gemm_tn(const void* A, int lda, const void* B, int ldb, TmapDtype dtype, const GemmShape& shape,
                   const GemmEpi& epi, cudaStream_t stream, int impl) {
  if (impl < 0) impl = default_gemm_impl();
  GemmShape s = shape;
  if (s.splits < 1) s.splits = 1;
  DVT_REQUIRE(s.M > 0 && s.N > 0 && s.K > 0, "gemm: empty shape M=%d N=%d K=%d", s.M, s.N, s.K);
  DVT_REQUIRE(s.splits == 1 || epi.out_mode == OUT_F32_ATOMIC, "gemm: split-K needs OUT_F32_ATOMIC");
  DVT_REQUIRE(epi.out == nullptr || epi.ldo % 4 == 0, "gemm: ldo must be a multiple of 4 (got %d)", epi.ldo);

  if (impl == GEMM_SIMT_DEBUG) {
    dim3 grid((s.N + 15) / 16, (s.M + 15) / 16), block(16, 16);
    if (dtype == TMAP_BF16)
      gemm_tn_simt_kernel<__nv_bfloat16><<<grid, block, 0, stream>>>(
          reinterpret_cast<const __nv_bfloat16*>(A), lda, reinterpret_cast<const __nv_bfloat16*>(B), ldb, s, epi);
    else
      gemm_tn_simt_kernel<float><<<grid, block, 0, stream>>>(reinterpret_cast<const float*>(A), lda,
                                                             reinterpret_cast<const float*>(B), ldb, s, epi);
    DVT_CUDA_OK(cudaGetLastError());
    return DVT_OK;
  }

  const int elem = dtype == TMAP_BF16 ? 2 : 4;
  const int bk = KB_BYTES / elem;
  DVT_REQUIRE((lda * elem) % 16 == 0 && (ldb * elem) % 16 == 0,
              "gemm: row pitch must be a multiple of 16 bytes (lda=%d ldb=%d elem=%d)", lda, ldb, elem);
  DVT_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0,
              "gemm: operands must be 16-byte aligned");
  // wide tiles when N is large enough to fill them; narrow ones for the small fit GEMMs
  const bool wide = s.N >= 256 && (s.N % 256 == 0 || s.N > 1024);
  CUtensorMap tmA, tmB;
  int rc = make_tmap_2d(&tmA, A, dtype, (uint64_t)s.M, (uint64_t)s.K, (uint64_t)lda * elem, BM, bk);
  if (rc) return rc;
  rc = make_tmap_2d(&tmB, B, dtype, (uint64_t)s.N, (uint64_t)s.K, (uint64_t)ldb * elem, wide ? 256 : 128, bk);
  if (rc) return rc;
  if (dtype == TMAP_BF16) {
    return wide ? launch_tc<256, 4, false>(tmA, tmB, s, epi, stream) : launch_tc<128, 6, false>(tmA, tmB, s, epi, stream);
  } else {
    return wide ? launch_tc<256, 4, true>(tmA, tmB, s, epi, stream) : launch_tc<128, 6, true>(tmA, tmB, s, epi, stream);
  }
}

}  // namespace dvt
