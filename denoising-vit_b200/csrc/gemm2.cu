// TN GEMM on CTA PAIRS (tcgen05 cta_group::2): C[M,N] = A[M,K] * B[N,K]^T, bf16 operands, both K-major -- the large GEMMs
// of the ViT forward (QKV, out-proj, fc1, fc2; reference: timm VisionTransformer reached from
// dvt/models/vit_wrapper.py:136-143) and of the stage-2 denoiser block.
//
// Why: the single-CTA 128x256 kernel of gemm.cu is shared-memory-bandwidth bound (per 64-wide k-block 48 KB of TMA writes
// + 48 KB of tensor-core operand reads against 512 clk of MMA).  A pair of CTAs on the two SMs of a TPC computes a 256x256
// tile with ONE instruction stream: each CTA stages its 128 rows of A and only HALF of the B tile (128 of the 256 rows);
// the tensor cores of both SMs read B from both shared memories.  Per SM and k-block: 32 KB of TMA writes instead of 48.
//
// Cluster of 2 CTAs, per CTA 14 warps:
//   warp 0      TMA producer (both CTAs): its A half and its B half, `cp.async.bulk.tensor ... cta_group::2` completing on
//               the LEADER's (rank 0) full barrier; 5-stage ring of 32 KB
//   warp 1      MMA issuer (leader only): tcgen05.mma.cta_group::2 M256 N256 K16, accumulators in the TMEM of both CTAs
//               (two stages of 256 columns); tcgen05.commit multicast frees the smem slot / publishes the accumulator in
//               both CTAs
//   warps 2-13  epilogue (both CTAs, each on its own 128 rows): tcgen05.ld -> epi_chunk() of gemm.cu (same fused
//               epilogues); "accumulator drained" arrives on the leader's barrier from both CTAs
#include "gemm.cuh"

namespace dvt {

namespace {

constexpr int G2_BM = 128;          // rows per CTA (256 per pair)
constexpr int G2_BN = 256;
constexpr int G2_BK = 64;           // bf16 elements per stage row = one 128-byte swizzle atom
constexpr int G2_STAGES = 5;
constexpr int G2_EW = 12;
constexpr int G2_THREADS = 32 * (2 + G2_EW);
constexpr int G2_A_BYTES = G2_BM * 128;            // 16 KB
constexpr int G2_B_BYTES = (G2_BN / 2) * 128;      // 16 KB: this CTA's half of the B tile
constexpr int G2_STAGE_BYTES = G2_A_BYTES + G2_B_BYTES;
constexpr int G2_OFF_SCR = G2_STAGES * G2_STAGE_BYTES;
constexpr int G2_SCR_BYTES = G2_EW * 32 * SCR_PITCH * 4;
constexpr int G2_OFF_BAR = G2_OFF_SCR + G2_SCR_BYTES;
constexpr int G2_NUM_BARS = 2 * G2_STAGES + 4;
constexpr int G2_OFF_TMEM = G2_OFF_BAR + G2_NUM_BARS * 8;
constexpr int G2_SMEM_TOTAL = G2_OFF_TMEM + 16 + 1024;
static_assert(G2_SMEM_TOTAL <= 227 * 1024, "pair GEMM exceeds shared memory");

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `smem_addr` (a shared::cta address of this CTA) in the CTA of rank `rank`
__device__ __forceinline__ uint32_t mapa_shared(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_2d_cg2(void* smem_dst, const CUtensorMap* m, uint32_t mbar_cluster_addr, int c0,
                                                int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(mbar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_f16_cg2(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once the MMAs issued so far have completed) on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_cg2(uint64_t* bar) {
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(mask)
               : "memory");
}
__device__ __forceinline__ void tmem_alloc_cg2(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_cg2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_cg2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(G2_THREADS, 1)
gemm_tn_cg2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, GemmShape s, GemmEpi e) {
  extern __shared__ uint8_t smem_raw[];
  // (dynamic shared memory starts at the same CTA-relative offset in both CTAs, so the aligned base does too)
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* scr_all = reinterpret_cast<float*>(smem + G2_OFF_SCR);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + G2_OFF_BAR);
  uint64_t* full = bars;                      // used in the leader only
  uint64_t* empty = bars + G2_STAGES;         // per CTA
  uint64_t* tfull = bars + 2 * G2_STAGES;     // per CTA
  uint64_t* tempty = bars + 2 * G2_STAGES + 2;  // used in the leader only (both CTAs' epilogue warps arrive)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + G2_OFF_TMEM);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

  const int tiles_m = (s.M + 2 * G2_BM - 1) / (2 * G2_BM);
  const int tiles_n = (s.N + G2_BN - 1) / G2_BN;
  const int num_tiles = tiles_m * tiles_n;
  const int kb_total = (s.K + G2_BK - 1) / G2_BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < G2_STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 2 * G2_EW * 32);
    }
    fence_mbar_init();
  }
  if (warp == 1) {  // the same warp in both CTAs: a pair-wide allocation of all 512 columns
    tmem_alloc_cg2(tmem_slot, 512);
    tmem_relinquish_cg2();
  }
  tc_fence_before();
  cluster_sync_all();  // barriers of BOTH CTAs are initialised before anything is signalled across the pair
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_trigger();

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = cluster_id; t < num_tiles; t += num_clusters) {
        const int tn = t % tiles_n, tm = t / tiles_n;
        const int row_a = tm * 2 * G2_BM + (int)rank * G2_BM;
        const int row_b = tn * G2_BN + (int)rank * (G2_BN / 2);
        for (int kb = 0; kb < kb_total; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1, 0x71);
          uint8_t* sA = smem + stage * G2_STAGE_BYTES;
          const uint32_t full_leader = mapa_shared(smem_u32(&full[stage]), 0);
          if (leader) mbar_expect_tx(&full[stage], 2 * G2_STAGE_BYTES);  // the bytes of both CTAs land on this barrier
          tma_load_2d_cg2(sA, &tmA, full_leader, kb * G2_BK, row_a);
          tma_load_2d_cg2(sA + G2_A_BYTES, &tmB, full_leader, kb * G2_BK, row_b);
          if (++stage == G2_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader && lane == 0) {
      constexpr uint32_t idesc = make_idesc(1u, 2 * G2_BM, G2_BN, 0u, 0u);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int t = cluster_id; t < num_tiles; t += num_clusters) {
        mbar_wait(&tempty[as], aphase ^ 1, 0x72);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * G2_BN;
        for (int kb = 0; kb < kb_total; ++kb) {
          mbar_wait(&full[stage], phase, 0x73);
          tc_fence_after();
          const uint32_t a_base = smem_u32(smem + stage * G2_STAGE_BYTES);
          const uint64_t da = make_smem_desc(a_base, 0, 1024, 2);
          const uint64_t db = make_smem_desc(a_base + G2_A_BYTES, 0, 1024, 2);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16_cg2(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit_cg2(&empty[stage]);  // frees this smem slot in both CTAs
          if (++stage == G2_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit_cg2(&tfull[as]);  // accumulator complete, in both CTAs
        if (++as == 2) {
          as = 0;
          aphase ^= 1;
        }
      }
    }
  } else {
    // ===================== epilogue warps (both CTAs) =====================
    const int ew = warp - 2;
    const int quad = warp & 3;
    constexpr int CSTEP = G2_EW / 4;
    const int c_first = ew >> 2;
    float* scr = scr_all + ew * 32 * SCR_PITCH;
    int as = 0;
    uint32_t aphase = 0;
    for (int t = cluster_id; t < num_tiles; t += num_clusters) {
      const int tn = t % tiles_n, tm = t / tiles_n;
      mbar_wait(&tfull[as], aphase, 0x74);
      tc_fence_after();
      const uint32_t taddr_row = tmem_base + ((uint32_t)(quad * 32) << 16) + as * G2_BN;
      const uint32_t tempty_leader = mapa_shared(smem_u32(&tempty[as]), 0);
#pragma unroll 1
      for (int c = c_first; c < G2_BN / 32; c += CSTEP) {
        uint32_t r[32];
        tmem_ld_32x32(taddr_row + c * 32, r);
        tmem_ld_wait();
        if (c + CSTEP >= G2_BN / 32) {  // this warp's last read of the accumulator stage
          tc_fence_before();
          mbar_arrive_cluster(tempty_leader);
        }
        const int n_base = tn * G2_BN + c * 32;
        if (n_base >= s.N) continue;
        epi_chunk(e, s, scr, lane, r, n_base, tm * 2 * G2_BM + (int)rank * G2_BM + quad * 32 + (lane >> 3), true);
        __syncwarp();
      }
      if (++as == 2) {
        as = 0;
        aphase ^= 1;
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();  // the leader's MMAs read the peer's shared memory / write its TMEM: nobody leaves early
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_cg2(tmem_base, 512);
  }
}

}  // namespace

bool gemm_cg2_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* v = getenv("DVT_GEMM_CG2");
    on = (v && v[0] == '0') ? 0 : 1;
  }
  return on != 0;
}

// bf16, K-major operands, no split-K: the shapes where the pair kernel is used (launch_gemm_tn decides).
int launch_gemm_cg2(const void* A, int lda, const void* B, int ldb, const GemmShape& s, const GemmEpi& e, cudaStream_t stream) {
  static bool prepared = false;
  if (!prepared) {
    DVT_CUDA_OK(cudaFuncSetAttribute(gemm_tn_cg2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, G2_SMEM_TOTAL));
    prepared = true;
  }
  CUtensorMap tmA, tmB;
  int rc = make_tmap_2d(&tmA, A, TMAP_BF16, (uint64_t)s.M, (uint64_t)s.K, (uint64_t)lda * 2, G2_BM, G2_BK);
  if (rc) return rc;
  rc = make_tmap_2d(&tmB, B, TMAP_BF16, (uint64_t)s.N, (uint64_t)s.K, (uint64_t)ldb * 2, G2_BN / 2, G2_BK);
  if (rc) return rc;
  const int tiles = ((s.M + 2 * G2_BM - 1) / (2 * G2_BM)) * ((s.N + G2_BN - 1) / G2_BN);
  const int clusters = std::max(1, std::min(tiles, num_sms() / 2));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * clusters);
  cfg.blockDim = dim3(G2_THREADS);
  cfg.dynamicSmemBytes = G2_SMEM_TOTAL;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];   // (the cluster shape is a compile-time property of the kernel: __cluster_dims__(2, 1, 1))
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = s.pdl ? 1 : 0;
  DVT_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_tn_cg2_kernel, tmA, tmB, s, e));
  count_launch();
  DVT_CUDA_OK(cudaGetLastError());
  return DVT_OK;
}

}  // namespace dvt
