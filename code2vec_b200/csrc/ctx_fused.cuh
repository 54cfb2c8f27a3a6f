// Fused context forward for sm_100a:  H = tanh( dropout([ tok[src] | path[pth] | tok[tgt] ]) . W )
// (tensorflow_model.py:238-252) in ONE kernel -- the embedding gather feeds the tensor core directly, the
// gathered context matrix X' is never read back from HBM.
//
// Persistent, warp-specialised, one CTA per SM (576 threads):
//   warp 0        : TMA producer of the B operand (W tiles, cp.async.bulk.tensor, 128-byte swizzle, MN-major);
//                   also stages each tile's index triples (src, path, tgt) into shared memory with 1-D TMA
//                   bulk copies (cp.async.bulk.shared.global) one tile ahead
//   warp 1        : MMA issuer -- tcgen05.mma.cta_group::1.kind::tf32 (UMMA 128 x BN x 8), TMEM allocation
//   warps 2..9    : epilogue -- tcgen05.ld of their TMEM lane quadrant, tanh, coalesced stores of H
//   warps 10..17  : GATHER producers of the A operand, one PAIR of warps per shared-memory stage.  A stage is 128
//                   contexts x 32 floats (one 128-byte swizzle row per context): 8 lanes read the 128 contiguous bytes of
//                   one table row segment with 128-bit loads (4 rows per warp instruction, 16 per thread), apply the
//                   Philox dropout multipliers in registers and store 16-byte chunks at the SWIZZLE_128B position
//                   (chunk ^ (row & 7)) the tensor core expects -- exactly the image TMA would have written from a
//                   materialised X'.  The four pairs work a quarter period apart: 64 KB of rows in flight per SM.
//                   When training, the dropped-out rows are also written out once (X', for the dW = X'^T.dU GEMM
//                   of the backward pass): a write the unfused path made too, without its read-back.
// Barriers: full[s] collects the TMA transaction of the W tile plus one arrival per gather warp (after a
// fence.proxy.async: the tensor core reads shared memory through the async proxy); empty[s] is the MMA's
// tcgen05.commit and releases both producers.
#pragma once
#include "umma_gemm.cuh"

namespace c2v {
namespace umma {

constexpr int kGatherWarps = 8;
constexpr int kGatherWarp0 = kEpiWarp0 + kEpiWarps;              // 10
constexpr int kFusedThreads = 32 * (kGatherWarp0 + kGatherWarps);   // 576
constexpr int kGatherRowsPerThread = BM * 8 / (32 * kGatherWarps);  // 4 chunks of 16 B per thread per stage
constexpr int kIdxBufs = 3;          // index triples are staged two tiles ahead of the gather warps

__device__ __forceinline__ void bulk_copy_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <int BN, int STAGES>
struct FusedSmem {
  static constexpr int kABytes = BM * BK * 4;
  static constexpr int kBBytes = BN * BK * 4;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kEpiStageOffset = STAGES * kStageBytes;
  static constexpr int kIdxOffset = kEpiStageOffset + kEpiWarps * kEpiStageBytes;     // kIdxBufs buffers x 3 x BM int32
  static constexpr int kBarOffset = kIdxOffset + kIdxBufs * 3 * BM * 4;
  static constexpr int kTotal = kBarOffset + 256 + 1024;
  static_assert(kTotal <= 232448, "exceeds the 227 KB of shared memory a CTA can opt into");
};

template <int BN, int STAGES, class Epi>
// 576 threads: the register file is handed out to a CTA in units of 4 warps, so 20 x 32 x 96 registers is the most that fits
__global__ void __launch_bounds__(kFusedThreads, 1)
ctx_fused_kernel(const __grid_constant__ CUtensorMap tmB, GemmShape gs, const __grid_constant__ ContextSource cs,
                 const __grid_constant__ Dropout dp, float* __restrict__ Xout, Epi epi) {
  using L = FusedSmem<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  int32_t* sidx = reinterpret_cast<int32_t*>(smem + L::kIdxOffset);        // [kIdxBufs][3][BM]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;      // [2]
  uint64_t* tempty_bar = tfull_bar + 2;          // [2]
  uint64_t* ifull_bar = tempty_bar + 2;          // [kIdxBufs] index triples of a tile have landed
  uint64_t* iempty_bar = ifull_bar + kIdxBufs;   // [kIdxBufs] ... and have been consumed by the gather warps
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(iempty_bar + kIdxBufs);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr uint32_t kTmemCols = (2 * BN <= 256) ? 256 : 512;
  const int total_items = gs.m_tiles * gs.n_tiles;
  const int total_kblocks = gs.K / BK;           // the launcher guarantees d % 32 == 0, so K = 3d has no tail
  const int kps = cs.d / BK;                     // k-blocks per segment (source token | path | target token)

  if (warp == 0 && lane == 0) {
    // full: the TMA transaction of the W tile + the stage's pair of gather warps
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1 + 2); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], kEpiWarps); }
    for (int a = 0; a < kIdxBufs; ++a) { mbar_init(&ifull_bar[a], 1); mbar_init(&iempty_bar[a], kGatherWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_ptr, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  // work item -> (m tile, n tile): n fastest, so the CTAs that run together share gathered rows through L2
  auto decode = [&](int item, int& mt, int& nt) {
    nt = item % gs.n_tiles;
    mt = item / gs.n_tiles;
  };

  if (warp == 0) {
    // ===================== TMA producer: W tiles + index triples =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int ib = 0;
      uint32_t iphase = 0;
      // a tile's index triples: three 1-D bulk copies of <= 512 B, issued ONE TILE AHEAD of its W loads (rows past M are
      // never read by the gather warps)
      auto stage_indices = [&](int item) {
        int mt, nt;
        decode(item, mt, nt);
        mbar_wait(&iempty_bar[ib], iphase ^ 1);
        const int m0 = mt * BM;
        const int rows = min(BM, gs.M - m0);
        const uint32_t bytes = (uint32_t)((rows * 4 + 15) & ~15);
        int32_t* dst = sidx + ib * 3 * BM;
        mbar_expect_tx(&ifull_bar[ib], 3 * bytes);
        bulk_copy_g2s(dst, cs.src + m0, bytes, &ifull_bar[ib]);
        bulk_copy_g2s(dst + BM, cs.pth + m0, bytes, &ifull_bar[ib]);
        bulk_copy_g2s(dst + 2 * BM, cs.tgt + m0, bytes, &ifull_bar[ib]);
        if (++ib == kIdxBufs) { ib = 0; iphase ^= 1; }
      };
      if ((int)blockIdx.x < total_items) stage_indices(blockIdx.x);
      for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
        int mt, nt;
        decode(item, mt, nt);
        if (item + (int)gridDim.x < total_items) stage_indices(item + gridDim.x);
        for (int kb = 0; kb < total_kblocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sb = smem + stage * L::kStageBytes + L::kABytes;
          mbar_expect_tx(&full_bar[stage], L::kBBytes);
#pragma unroll
          for (int c = 0; c < BN / 32; ++c) tma_load_2d(sb + c * (BK * 128), &tmB, &full_bar[stage], nt * BN + c * 32, kb * BK);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_tf32(BM, BN, false, true);
      constexpr uint32_t b_lbo = BK * 128, a_sbo = 1024, b_sbo = 512;
      constexpr uint32_t a_kstep = UMMA_K * 4, b_kstep = 1024;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = 0; kb < total_kblocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * L::kStageBytes);
          const uint32_t sb = sa + L::kABytes;
          const uint64_t adesc = make_smem_desc(sa, 0, a_sbo, kLayoutSw128);
          const uint64_t bdesc = make_smem_desc(sb, b_lbo, b_sbo, kLayoutSw128Base32);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            mma_tf32(tmem_d, adesc + (uint64_t)((k * a_kstep) >> 4), bdesc + (uint64_t)((k * b_kstep) >> 4), idesc,
                     (kb > 0 || k > 0) ? 1u : 0u);
          }
          tc_commit(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        tc_commit(&tfull_bar[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp < kGatherWarp0) {
    // ===================== epilogue warps (as in umma_gemm_kernel) =====================
    const int q = warp & 3;
    const int half = (warp - kEpiWarp0) >> 2;
    constexpr int kChunksPerHalf = BN / 64;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      int mt, nt;
      decode(item, mt, nt);
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const int m = mt * BM + q * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
      typename Epi::State est;
      epi.begin(est);
      drain_accumulator(epi, est, taddr, half * kChunksPerHalf * 32, kChunksPerHalf * 32, mt * BM + q * 32, lane, nt * BN,
                        gs.M, gs.N, 0, reinterpret_cast<float*>(smem + L::kEpiStageOffset + (warp - kEpiWarp0) * kEpiStageBytes));
      epi.end(m, 2 * nt + half, 0, m < gs.M, est);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // ===================== gather producers of the A operand =====================
    // The (tile, k-block) steps of this CTA form one stream g = 0, 1, 2, ...; step g lands in smem stage g % STAGES.
    // A PAIR of warps owns every STAGES-th step (pair p <-> stage p): its 64 threads issue the 16 row-segment loads
    // (16 B each) of their step, wait for the stage to be free, apply dropout, store the swizzled chunks, fence and
    // arrive.  The four pairs run a quarter period apart, so four stages' worth of rows (64 KB per SM) are in flight
    // while no thread ever executes its proxy fence with loads of a LATER step outstanding (the fence would wait
    // for them and serialise the stream).
    static_assert(kGatherWarps == 2 * STAGES, "one pair of gather warps per shared-memory stage");
    const int gw = warp - kGatherWarp0;
    const int pair = gw >> 1;                                // == the smem stage this pair fills
    const int t64 = (gw & 1) * 32 + lane;                    // 0 .. 63 within the pair
    const int c = t64 & 7;                                   // 16-byte chunk of the 128-byte row segment
    const int rg = t64 >> 3;                                 // rows rg, rg + 8, ..., rg + 120
    constexpr int R = BM / 8;                                // 16 rows per thread
    const uint32_t soff = (uint32_t)(rg * 128 + ((c ^ rg) << 4));     // (r & 7) == rg for every row of this thread
    const bool drop = dp.enabled != 0;
    const int items_here = (total_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int total_steps = items_here * total_kblocks;
    uint8_t* sa = smem + pair * L::kStageBytes + soff;
    int released = 0;                                        // tiles whose index buffers this warp has released
    auto release_to = [&](int k_end) {                       // this warp will not read the indices of tiles < k_end again
      for (; released < k_end; ++released) {
        __syncwarp();
        if (lane == 0) mbar_arrive(&iempty_bar[released % kIdxBufs]);
      }
    };
    uint32_t visit = 0;
    for (int g = pair; g < total_steps; g += STAGES, ++visit) {
      const int k = g / total_kblocks, kb = g - k * total_kblocks;
      const int item = blockIdx.x + k * gridDim.x;
      const int nt = item % gs.n_tiles;
      const int m0 = (item / gs.n_tiles) * BM;
      release_to(k);
      mbar_wait(&ifull_bar[k % kIdxBufs], (uint32_t)(k / kIdxBufs) & 1u);       // the tile's index triples have landed
      const int seg = kb / kps;
      const int col = (kb - seg * kps) * BK + c * 4;
      const int32_t* ids = sidx + (k % kIdxBufs) * 3 * BM + seg * BM;
      float4 x[R];
#pragma unroll
      for (int i = 0; i < R; ++i) {
        const int r = rg + 8 * i;
        if (m0 + r < gs.M) {
          const float* rowp = (seg == 1) ? table_row(cs.path, ids[r], cs.d) : table_row(cs.tok, ids[r], cs.d);
          x[i] = __ldg(reinterpret_cast<const float4*>(rowp + col));
        } else {
          x[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      mbar_wait(&empty_bar[pair], (visit & 1u) ^ 1u);        // the MMAs that read this stage last time have retired
      const bool wx = (Xout != nullptr) && nt == 0;
      float* xo = Xout + (size_t)(m0 + rg) * gs.K + kb * BK + c * 4;
#pragma unroll
      for (int i = 0; i < R; ++i) {
        float4 v = x[i];
        if (drop) {
          const float4 mlt = dropout_mult4(dp, m0 + rg + 8 * i, kb * (BK / 4) + c);
          v.x *= mlt.x; v.y *= mlt.y; v.z *= mlt.z; v.w *= mlt.w;
        }
        *reinterpret_cast<float4*>(sa + i * (8 * 128)) = v;
        if (wx && m0 + rg + 8 * i < gs.M) *reinterpret_cast<float4*>(xo + (size_t)i * 8 * gs.K) = v;
      }
      fence_proxy_async_smem();                              // generic-proxy stores -> visible to the tensor core
      __syncwarp();
      if (lane == 0) mbar_arrive(&full_bar[pair]);
    }
    release_to(items_here);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

// H[M, N] = epi( gather(cs)[M, K = 3d] . W[K, N] ).  W row-major [K, N] (N contiguous).  Xout: nullptr or [M, K].
// Preconditions (checked by the caller): d % 32 == 0; the three index arrays are 16-byte aligned and readable up to the
// next multiple of 4 entries past M (the engine's staging buffers are whole [max_batch, C] arrays in a 256-byte aligned
// workspace; caller-owned arrays are checked for M % 4 == 0).
template <int BN, int STAGES, class Epi>
inline cudaError_t launch_ctx_fused(cudaStream_t st, int M, int N, const float* W, size_t ldw, const ContextSource& cs,
                                    const Dropout& dp, float* Xout, const Epi& epi, int num_sms) {
  using L = FusedSmem<BN, STAGES>;
  const int K = 3 * cs.d;
  CUtensorMap tmB;
  if (!make_tensor_map(&tmB, W, (uint64_t)K, (uint64_t)N, ldw, BK, true)) return cudaErrorInvalidValue;
  GemmShape gs;
  gs.M = M; gs.N = N; gs.K = K;
  gs.terms = 1;
  gs.m_tiles = (M + BM - 1) / BM;
  gs.n_tiles = (N + BN - 1) / BN;
  gs.splits = 1;
  gs.kblocks_per_split = K / BK;
  gs.n_fastest = 1;
  auto kern = ctx_fused_kernel<BN, STAGES, Epi>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal);
  if (e != cudaSuccess) return e;
  int grid = gs.m_tiles * gs.n_tiles;
  if (grid > num_sms) grid = num_sms;
  kern<<<grid, kFusedThreads, L::kTotal, st>>>(tmB, gs, cs, dp, Xout, epi);
  return cudaGetLastError();
}

}  // namespace umma
}  // namespace c2v
