// CTA-pair variant of the tcgen05 GEMM (umma_gemm.cuh): tcgen05.mma.cta_group::2, UMMA 256 x BN x 8.
//
// Why: with 32-bit (tf32) operands a single CTA's MMA stream reads A (128 x 8) and B (BN x 8) from
// shared memory for every instruction while TMA refills the same stages; at 128 x 256 that is ~190
// B/clk against a 128 B/clk shared-memory port, which caps the tensor pipe near 2/3.  In a CTA pair
// (two SMs of a TPC, cluster dims 2x1x1) each CTA stages its own 128 rows of A but only HALF of the
// B tile; the instruction issued by the leader CTA consumes both halves, so the per-SM shared-memory
// and L2 traffic for B halves and the pipe can run close to full rate.
//
//   per CTA:  warp 0 TMA producer (own A rows + own half of B, signalling the LEADER's full barrier),
//             warp 1 TMEM allocator (both CTAs) / MMA issuer (leader only),
//             warps 2..9 epilogue of the CTA's own 128 accumulator rows (same functors as umma_gemm.cuh).
//   barriers: full[s]   in the leader  -- 2 arrivals (one per CTA) + the bytes of both CTAs' TMA loads
//             empty[s]  in each CTA    -- tcgen05.commit multicast from the leader frees the stage in both
//             tfull[a]  in each CTA    -- tcgen05.commit multicast: accumulator a complete
//             tempty[a] in the leader  -- 2 x 8 epilogue warps (the peer's arrive remotely)
#pragma once
#include "umma_gemm.cuh"

namespace c2v {
namespace umma {

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local` (an address in this CTA's shared window) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_commit_pair(uint64_t* bar) {      // arrive on `bar` in both CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void mma_tf32_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

template <int BN, int STAGES>
struct SmemLayout2 {
  static constexpr int kABytes = BM * BK * 4;              // this CTA's 128 rows of A
  static constexpr int kBBytes = (BN / 2) * BK * 4;        // this CTA's half of the B tile
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kEpiStageOffset = STAGES * kStageBytes;
  static constexpr int kBarOffset = kEpiStageOffset + kEpiWarps * kEpiStageBytes;
  static constexpr int kTotal = kBarOffset + 256 + 1024;
  static_assert(kTotal <= 232448, "exceeds the 227 KB of shared memory a CTA can opt into");
};

template <int BN, int STAGES, bool A_MN, bool B_MN, class Epi>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
umma_gemm2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmAlo, const __grid_constant__ CUtensorMap tmBlo, GemmShape gs, Epi epi) {
  if (epi_gate_closed(epi, 0)) return;      // gated fallback pass: both CTAs of every pair leave together
  using L = SmemLayout2<BN, STAGES>;
  static_assert(BN % 64 == 0, "each CTA stages BN/2 columns of B in 32-element chunks");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;      // [2]
  uint64_t* tempty_bar = tfull_bar + 2;          // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta = cluster_ctarank();        // 0 = leader (issues the MMAs), 1 = peer
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  constexpr uint32_t kTmemCols = (2 * BN <= 256) ? 256 : 512;
  const int total_items = gs.m_tiles * gs.n_tiles * gs.splits;      // m_tiles counts 256-row tiles here
  const int total_kblocks = (gs.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 2); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 2 * kEpiWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc_pair(tmem_ptr, kTmemCols);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                            // the peer's barriers are initialised before anything targets them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  auto decode = [&](int item, int& mt, int& nt, int& sp) {
    if (gs.n_fastest) {
      nt = item % gs.n_tiles;
      const int r = item / gs.n_tiles;
      mt = r % gs.m_tiles;
      sp = r / gs.m_tiles;
    } else {
      mt = item % gs.m_tiles;
      const int r = item / gs.m_tiles;
      nt = r % gs.n_tiles;
      sp = r / gs.n_tiles;
    }
  };

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = pair; item < total_items; item += num_pairs) {
        int mt, nt, sp;
        decode(item, mt, nt, sp);
        const int kb0 = sp * gs.kblocks_per_split;
        const int kb1 = min(total_kblocks, kb0 + gs.kblocks_per_split);
        const int m0 = mt * (2 * BM) + (int)cta * BM;             // this CTA's rows of A
        const int n0 = nt * BN + (int)cta * (BN / 2);             // this CTA's columns of B
        auto load_stage = [&](const CUtensorMap* mA, const CUtensorMap* mB, int kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * L::kStageBytes;
          uint8_t* sb = sa + L::kABytes;
          const uint32_t leader_full = mapa_u32(smem_u32(&full_bar[stage]), 0);
          if (A_MN) {
#pragma unroll
            for (int c = 0; c < BM / 32; ++c) tma_load_2d_pair(sa + c * (BK * 128), mA, leader_full, m0 + c * 32, kb * BK);
          } else {
            tma_load_2d_pair(sa, mA, leader_full, kb * BK, m0);
          }
          if (B_MN) {
#pragma unroll
            for (int c = 0; c < BN / 64; ++c) tma_load_2d_pair(sb + c * (BK * 128), mB, leader_full, n0 + c * 32, kb * BK);
          } else {
            tma_load_2d_pair(sb, mB, leader_full, kb * BK, n0);
          }
          if (cta == 0) mbar_expect_tx(&full_bar[stage], 2 * L::kStageBytes);     // bytes of both CTAs land on this barrier
          else mbar_arrive_cluster(leader_full);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        };
        if (gs.terms == 3) {                                                     // 3xTF32: cross terms first (umma_gemm.cuh)
          for (int kb = kb0; kb < kb1; ++kb) load_stage(&tmAlo, &tmB, kb);
          for (int kb = kb0; kb < kb1; ++kb) load_stage(&tmA, &tmBlo, kb);
        }
        for (int kb = kb0; kb < kb1; ++kb) load_stage(&tmA, &tmB, kb);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (cta == 0 && lane == 0) {
      constexpr uint32_t idesc = make_idesc_tf32(2 * BM, BN, A_MN, B_MN);
      constexpr uint32_t a_lbo = A_MN ? BK * 128 : 0, b_lbo = B_MN ? BK * 128 : 0;
      constexpr uint32_t a_sbo = A_MN ? 512 : 1024, b_sbo = B_MN ? 512 : 1024;
      constexpr uint32_t a_lt = A_MN ? kLayoutSw128Base32 : kLayoutSw128, b_lt = B_MN ? kLayoutSw128Base32 : kLayoutSw128;
      constexpr uint32_t a_kstep = A_MN ? 1024 : UMMA_K * 4, b_kstep = B_MN ? 1024 : UMMA_K * 4;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int item = pair; item < total_items; item += num_pairs) {
        int mt, nt, sp;
        decode(item, mt, nt, sp);
        const int kb0 = sp * gs.kblocks_per_split;
        const int kb1 = min(total_kblocks, kb0 + gs.kblocks_per_split);
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);          // both CTAs' epilogues have drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        const int n_steps = (kb1 - kb0) * gs.terms;
        for (int it = 0; it < n_steps; ++it) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * L::kStageBytes);
          const uint32_t sb = sa + L::kABytes;
          const uint64_t adesc = make_smem_desc(sa, a_lbo, a_sbo, a_lt);
          const uint64_t bdesc = make_smem_desc(sb, b_lbo, b_sbo, b_lt);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            mma_tf32_pair(tmem_d, adesc + (uint64_t)((k * a_kstep) >> 4), bdesc + (uint64_t)((k * b_kstep) >> 4), idesc,
                          (it > 0 || k > 0) ? 1u : 0u);
          }
          tc_commit_pair(&empty_bar[stage]);                  // frees this stage in both CTAs
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        tc_commit_pair(&tfull_bar[acc]);                      // accumulator complete in both CTAs
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue warps (own 128 accumulator rows) =====================
    const int q = warp & 3;
    const int half = (warp - kEpiWarp0) >> 2;
    constexpr int kChunksPerHalf = BN / 64;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int item = pair; item < total_items; item += num_pairs) {
      int mt, nt, sp;
      decode(item, mt, nt, sp);
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const int m = mt * (2 * BM) + (int)cta * BM + q * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
      typename Epi::State est;
      epi.begin(est);
      drain_accumulator(epi, est, taddr, half * kChunksPerHalf * 32, kChunksPerHalf * 32,
                        mt * (2 * BM) + (int)cta * BM + q * 32, lane, nt * BN, gs.M, gs.N, sp,
                        reinterpret_cast<float*>(smem + L::kEpiStageOffset + (warp - kEpiWarp0) * kEpiStageBytes));
      epi.end(m, 2 * nt + half, sp, m < gs.M, est);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (cta == 0) mbar_arrive(&tempty_bar[acc]);
        else mbar_arrive_cluster(mapa_u32(smem_u32(&tempty_bar[acc]), 0));
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                            // neither CTA frees TMEM / exits while the other may still use it
  if (warp == 1) tmem_dealloc_pair(tmem_base, kTmemCols);
}

template <int BN, int STAGES, bool A_MN, bool B_MN, class Epi>
inline cudaError_t launch2_cfg(cudaStream_t st, int M, int N, int K, int splits, const Operand& A, const Operand& B, const Epi& epi,
                               int num_sms) {
  using L = SmemLayout2<BN, STAGES>;
  CUtensorMap tmA, tmB, tmAlo, tmBlo;
  const bool three = A.lo != nullptr && B.lo != nullptr;
  if ((A.lo != nullptr) != (B.lo != nullptr)) return cudaErrorInvalidValue;
  auto mapA = [&](CUtensorMap* m, const float* base) {
    return A_MN ? make_tensor_map(m, base, (uint64_t)K, (uint64_t)M, A.ld, BK, true)
                : make_tensor_map(m, base, (uint64_t)M, (uint64_t)K, A.ld, BM, false);
  };
  auto mapB = [&](CUtensorMap* m, const float* base) {
    return B_MN ? make_tensor_map(m, base, (uint64_t)K, (uint64_t)N, B.ld, BK, true)
                : make_tensor_map(m, base, (uint64_t)N, (uint64_t)K, B.ld, BN / 2, false);
  };
  if (!mapA(&tmA, A.base) || !mapB(&tmB, B.base)) return cudaErrorInvalidValue;
  if (three) { if (!mapA(&tmAlo, A.lo) || !mapB(&tmBlo, B.lo)) return cudaErrorInvalidValue; }
  else { tmAlo = tmA; tmBlo = tmB; }
  GemmShape gs;
  gs.M = M; gs.N = N; gs.K = K;
  gs.terms = three ? 3 : 1;
  gs.m_tiles = (M + 2 * BM - 1) / (2 * BM);
  gs.n_tiles = (N + BN - 1) / BN;
  const int total_kblocks = (K + BK - 1) / BK;
  if (splits < 1) splits = 1;
  if (splits > total_kblocks) splits = total_kblocks;
  gs.kblocks_per_split = (total_kblocks + splits - 1) / splits;
  gs.splits = (total_kblocks + gs.kblocks_per_split - 1) / gs.kblocks_per_split;
  gs.n_fastest = (gs.n_tiles < gs.m_tiles) ? 1 : 0;
  auto kern = umma_gemm2_kernel<BN, STAGES, A_MN, B_MN, Epi>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal);
  if (e != cudaSuccess) return e;
  int pairs = gs.m_tiles * gs.n_tiles * gs.splits;
  if (pairs > num_sms / 2) pairs = num_sms / 2;
  kern<<<2 * pairs, kThreads, L::kTotal, st>>>(tmA, tmB, tmAlo, tmBlo, gs, epi);
  return cudaGetLastError();
}

// BN = 256 -> 6 stages of 32 KB, BN = 192 -> 6 stages of 28 KB (per CTA), plus 32 KB of epilogue blocks.
template <int BN, int STAGES, class Epi>
inline cudaError_t launch2(cudaStream_t st, int M, int N, int K, int splits, const Operand& A, const Operand& B, const Epi& epi,
                           int num_sms) {
  if (!A.major_mn && !B.major_mn) return launch2_cfg<BN, STAGES, false, false, Epi>(st, M, N, K, splits, A, B, epi, num_sms);
  if (!A.major_mn && B.major_mn) return launch2_cfg<BN, STAGES, false, true, Epi>(st, M, N, K, splits, A, B, epi, num_sms);
  if (A.major_mn && !B.major_mn) return launch2_cfg<BN, STAGES, true, false, Epi>(st, M, N, K, splits, A, B, epi, num_sms);
  return launch2_cfg<BN, STAGES, true, true, Epi>(st, M, N, K, splits, A, B, epi, num_sms);
}

}  // namespace umma
}  // namespace c2v
