// tcgen05 (5th-gen tensor core) GEMM for sm_100a:  C[M,N] = A[M,K] . B[K,N],  fp32 storage,
// kind::tf32 operands (10-bit mantissa), fp32 accumulation in TMEM  (C2V_MATH_TF32).
//
// Persistent, warp-specialised, one CTA per SM:
//   warp 0      : TMA producer   -- cp.async.bulk.tensor (128-byte swizzle) into a 4-stage smem ring
//   warp 1      : MMA issuer     -- one elected lane issues tcgen05.mma.cta_group::1.kind::tf32
//                                   (UMMA 128 x BN x 8), tcgen05.commit frees smem stages / publishes
//                                   the accumulator; also owns the TMEM allocation
//   warps 2..9  : epilogue       -- tcgen05.ld (32x32b) of their TMEM lane quadrant (two warps per
//                                   quadrant, half of the columns each), fused epilogue functor
//                                   (store / tanh / log-sum-exp partials / split-K slice)
// The accumulator is double-buffered in TMEM (2 x BN columns) so the epilogue of tile i overlaps
// the MMAs of tile i+1.  Operands may be K-major (K contiguous) or MN-major (M resp. N
// contiguous) in global memory; both are staged as 128-byte swizzled rows and described to the
// tensor core through shared-memory matrix descriptors.  Out-of-range rows / K-tail are
// zero-filled by TMA, so any M, N, K work; split-K over blockIdx-independent work items.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace c2v {
namespace umma {

constexpr int BM = 128;          // UMMA_M (cta_group::1)
constexpr int BK = 32;           // fp32 elements per stage along K = one 128-byte swizzle row
constexpr int UMMA_K = 8;        // K per tcgen05.mma for 32-bit operands (32 bytes)
constexpr int kEpiWarps = 8;     // two warps per TMEM lane quadrant, each draining half of the tile's columns
constexpr int kThreads = 32 * (2 + kEpiWarps);
constexpr int kEpiWarp0 = 2;

// fast transcendental forms for the tensor-core path (operands are already tf32-rounded):
// exp via ex2.approx (rel. error 2^-22), tanh(x) = 1 - 2 / (exp(2x) + 1) (abs. error ~1e-7).
__device__ __forceinline__ float fast_tanh(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 2.885390081777927f));      // exp(2x) = 2^(2 log2(e) x); +-inf / 0 at the ends are fine
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.f));
  return fmaf(-2.f, r, 1.f);
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier ---------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
// Bounded wait: a pipeline bug traps (reported as a CUDA error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 20000000000LL) __trap();     // ~10 s at 2 GHz
  }
}

// ---- TMA ----------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}

// ---- tcgen05 ------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// 32 lanes x 32 consecutive 32-bit columns: thread i gets row (lane_base + i), columns [c, c+32)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
// 64 consecutive columns in one round trip: r0 = columns [c, c+32), r1 = [c+32, c+64)
__device__ __forceinline__ void tmem_ld64(uint32_t taddr, uint32_t (&a)[32], uint32_t (&b)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
      "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
      "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
      : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]), "=r"(a[4]), "=r"(a[5]), "=r"(a[6]), "=r"(a[7]), "=r"(a[8]),
        "=r"(a[9]), "=r"(a[10]), "=r"(a[11]), "=r"(a[12]), "=r"(a[13]), "=r"(a[14]), "=r"(a[15]), "=r"(a[16]),
        "=r"(a[17]), "=r"(a[18]), "=r"(a[19]), "=r"(a[20]), "=r"(a[21]), "=r"(a[22]), "=r"(a[23]), "=r"(a[24]),
        "=r"(a[25]), "=r"(a[26]), "=r"(a[27]), "=r"(a[28]), "=r"(a[29]), "=r"(a[30]), "=r"(a[31]),
        "=r"(b[0]), "=r"(b[1]), "=r"(b[2]), "=r"(b[3]), "=r"(b[4]), "=r"(b[5]), "=r"(b[6]), "=r"(b[7]), "=r"(b[8]),
        "=r"(b[9]), "=r"(b[10]), "=r"(b[11]), "=r"(b[12]), "=r"(b[13]), "=r"(b[14]), "=r"(b[15]), "=r"(b[16]),
        "=r"(b[17]), "=r"(b[18]), "=r"(b[19]), "=r"(b[20]), "=r"(b[21]), "=r"(b[22]), "=r"(b[23]), "=r"(b[24]),
        "=r"(b[25]), "=r"(b[26]), "=r"(b[27]), "=r"(b[28]), "=r"(b[29]), "=r"(b[30]), "=r"(b[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- descriptors ----------------------------------------------------------------------------------
// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout): start address >> 4 in
// [0,14), leading byte offset >> 4 in [16,30), stride byte offset >> 4 in [32,46), version = 1 in
// [46,48), layout type in [61,64): SWIZZLE_128B = 2 (16-byte chunks swizzled over 8 rows; K-major
// operands) or SWIZZLE_128B_BASE32B = 1 (32-byte chunks swizzled over 4 rows) -- the only layout the
// tensor core accepts for MN-major 32-bit (tf32) operands; TMA writes it with
// CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.
constexpr uint32_t kLayoutSw128 = 2, kLayoutSw128Base32 = 1;
__host__ __device__ constexpr uint64_t make_smem_desc_hi(uint32_t sbo_bytes, uint32_t layout_type) {
  return (uint64_t)((sbo_bytes >> 4) & 0x3FFF) | (1ull << 14) | ((uint64_t)layout_type << 29);     // upper 32 bits
}
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                   uint32_t layout_type) {
  const uint32_t lo = ((smem_addr >> 4) & 0x3FFF) | (((lbo_bytes >> 4) & 0x3FFF) << 16);
  return ((uint64_t)make_smem_desc_hi(sbo_bytes, layout_type) << 32) | lo;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, majors, N >> 3, M >> 4.
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N, bool a_mn, bool b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---- epilogues ---------------------------------------------------------------------------------------
// Each epilogue warp owns 32 accumulator rows (its TMEM lane quadrant) and a range of columns.  A
// functor supplies:  State / begin / end  -- per-(row, tile) state and its publication;
//                    observe(m, n0, 32 raw accumulators, nvalid, state) -- row-wise math (log-sum-exp);
//                    map(x)    -- the element-wise transform applied on the way out (identity, tanh);
//                    out(split), ldc -- where the tile goes;
//                    Pre / prefetch / store4 / store1(base, offset, value) -- what "storing" an element
//                    means (a plain write, or the optimizer update of the parameter the element is
//                    the gradient of; prefetch issues that update's loads ahead of the arithmetic).
// The store itself is done by drain_accumulator: TMEM -> registers (thread = row) -> a 4 KB
// XOR-swizzled shared-memory transpose per warp -> global stores in which every instruction writes
// four complete 128-byte row segments.  (Storing straight from the TMEM register layout makes each
// store instruction touch 32 different lines: 8x the LSU wavefronts, and that -- not the tensor
// pipe -- was what bounded the short-K GEMMs.)
struct EpiNoState {};

struct EpiStore {
  using State = EpiNoState;
  float* C;
  size_t ldc;
  size_t split_stride;
  __device__ __forceinline__ void begin(State&) const {}
  __device__ __forceinline__ void end(int, int, int, bool, State&) const {}
  __device__ __forceinline__ void observe(int, int, const uint32_t (&)[32], int, State&) const {}
  __device__ __forceinline__ float map(float x) const { return x; }
  __device__ __forceinline__ float* out(int split) const { return C + (size_t)split * split_stride; }
  using Pre = EpiNoState;
  static constexpr int kRowBatch = 8;
  static constexpr bool kWideDrain = true;
  __device__ __forceinline__ void prefetch(const float*, size_t, Pre&) const {}
  __device__ __forceinline__ void store4(float* c, size_t off, float4 v, const Pre&) const { *reinterpret_cast<float4*>(c + off) = v; }
  __device__ __forceinline__ void store1(float* c, size_t off, float x) const { c[off] = x; }
};
template <bool PRECISE>      // PRECISE: tanhf (the fp32-faithful 3xTF32 mode); else the ex2.approx form
struct EpiTanhStoreT {
  using State = EpiNoState;
  float* C;
  size_t ldc;
  __device__ __forceinline__ void begin(State&) const {}
  __device__ __forceinline__ void end(int, int, int, bool, State&) const {}
  __device__ __forceinline__ void observe(int, int, const uint32_t (&)[32], int, State&) const {}
  __device__ __forceinline__ float map(float x) const { return PRECISE ? tanhf(x) : fast_tanh(x); }
  __device__ __forceinline__ float* out(int) const { return C; }
  using Pre = EpiNoState;
  static constexpr int kRowBatch = 8;
  static constexpr bool kWideDrain = true;
  __device__ __forceinline__ void prefetch(const float*, size_t, Pre&) const {}
  __device__ __forceinline__ void store4(float* c, size_t off, float4 v, const Pre&) const { *reinterpret_cast<float4*>(c + off) = v; }
  __device__ __forceinline__ void store1(float* c, size_t off, float x) const { c[off] = x; }
};

// Logits epilogue: stores the tile of S and folds the row-wise (max, sum exp) of this warp's columns
// into a per-(row, half tile) partial, so the cross entropy needs no extra pass over S for its
// log-sum-exp (tensorflow_model.py:227-230).
template <bool PRECISE>      // PRECISE: expf (3xTF32 mode); else ex2.approx
struct EpiStoreLseT {
  struct State { float mx, sum; };
  float* C;
  size_t ldc;
  float2* partial;       // [M, slots]; slot = 2 * n_tile + column half
  int slots;
  __device__ __forceinline__ float ex(float x) const { return PRECISE ? expf(x) : __expf(x); }
  __device__ __forceinline__ void begin(State& st) const { st.mx = -INFINITY; st.sum = 0.f; }
  __device__ __forceinline__ void end(int m, int slot, int, bool row_ok, State& st) const {
    if (row_ok) partial[(size_t)m * slots + slot] = make_float2(st.mx, st.sum);
  }
  __device__ __forceinline__ float map(float x) const { return x; }
  __device__ __forceinline__ float* out(int) const { return C; }
  using Pre = EpiNoState;
  static constexpr int kRowBatch = 8;
  static constexpr bool kWideDrain = true;
  __device__ __forceinline__ void prefetch(const float*, size_t, Pre&) const {}
  __device__ __forceinline__ void store4(float* c, size_t off, float4 v, const Pre&) const { *reinterpret_cast<float4*>(c + off) = v; }
  __device__ __forceinline__ void store1(float* c, size_t off, float x) const { c[off] = x; }
  __device__ __forceinline__ void observe(int, int, const uint32_t (&r)[32], int nvalid, State& st) const {
    float cm = -INFINITY;
    if (nvalid >= 32) {              // a full chunk (all but the last tile of a row): no per-element bounds
#pragma unroll
      for (int j = 0; j < 32; ++j) cm = fmaxf(cm, __uint_as_float(r[j]));
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < nvalid) cm = fmaxf(cm, __uint_as_float(r[j]));
    }
    if (cm > st.mx) { st.sum *= ex(st.mx - cm); st.mx = cm; }
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    if (!PRECISE && nvalid >= 32) {
      // exp(x - mx) as one FFMA + one ex2.approx: 2^(x log2e - mx log2e)   (same 2^-22 relative accuracy as __expf)
      constexpr float L2E = 1.4426950408889634f;
      const float c = -st.mx * L2E;
      auto e2 = [](float t) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(t)); return y; };
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        a0 += e2(fmaf(__uint_as_float(r[j + 0]), L2E, c));
        a1 += e2(fmaf(__uint_as_float(r[j + 1]), L2E, c));
        a2 += e2(fmaf(__uint_as_float(r[j + 2]), L2E, c));
        a3 += e2(fmaf(__uint_as_float(r[j + 3]), L2E, c));
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        if (j + 0 < nvalid) a0 += ex(__uint_as_float(r[j + 0]) - st.mx);
        if (j + 1 < nvalid) a1 += ex(__uint_as_float(r[j + 1]) - st.mx);
        if (j + 2 < nvalid) a2 += ex(__uint_as_float(r[j + 2]) - st.mx);
        if (j + 3 < nvalid) a3 += ex(__uint_as_float(r[j + 3]) - st.mx);
      }
    }
    st.sum += (a0 + a1) + (a2 + a3);
  }
};

// Logits pass 1 of the recomputing schedule: ONLY the per-(row, half tile) (max, sum exp) partials -- nothing is stored, so the
// epilogue neither transposes through shared memory nor writes the 1.07 GB slab (tensorflow_model.py:227-230).
template <bool PRECISE>
struct EpiLseOnlyT : EpiStoreLseT<PRECISE> {
  static constexpr bool kStores = false;
};
// Logits pass 2: the same product again, and the epilogue writes dL/dlogits = (softmax - onehot) / B straight from the
// accumulator, given each row's log-sum-exp from pass 1: the slab is written once, as the gradient, and never read back
// by a softmax pass.  SPLIT (3xTF32): written as its tf32 split (high parts to C, residuals to C_lo).
template <bool PRECISE, bool SPLIT>
struct EpiSoftmaxGradT {
  struct State { float l; int tgt; };      // l: the row's log-sum-exp (PRECISE) or its exponent offset -lse log2e + log2(1/B)
  float* C;
  float* C_lo;
  size_t ldc;
  const float* lse;        // [M]
  const int32_t* target;   // [M] global class ids
  int row0;                // first class of this slab (row-sharded target table)
  float inv_batch;
  int M;
  __device__ __forceinline__ void begin(State& st) const { st.l = 0.f; st.tgt = -1; }
  __device__ __forceinline__ void end(int, int, int, bool, State&) const {}
  // called with this thread's row before any of its elements is mapped
  __device__ __forceinline__ void observe(int m, int, const uint32_t (&)[32], int, State& st) const {
    st.l = PRECISE ? lse[m] : fmaf(-lse[m], 1.4426950408889634f, log2f(inv_batch));
    st.tgt = target[m] - row0;
  }
  __device__ __forceinline__ float map(float x) const { return x; }
  __device__ __forceinline__ float map_at(float x, int col, const State& st) const {
    float p;
    if (PRECISE) {
      p = expf(x - st.l) * inv_batch;
    } else {          // exp(x - lse) / B as one FFMA + one ex2.approx
      asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p) : "f"(fmaf(x, 1.4426950408889634f, st.l)));
    }
    return col == st.tgt ? p - inv_batch : p;
  }
  __device__ __forceinline__ float* out(int) const { return C; }
  using Pre = EpiNoState;
  static constexpr int kRowBatch = 8;
  static constexpr bool kWideDrain = true;
  __device__ __forceinline__ void prefetch(const float*, size_t, Pre&) const {}
  __device__ __forceinline__ void store4(float* c, size_t off, float4 v, const Pre&) const {
    if (SPLIT) {
      float4 hi, lo;
      split_tf32(v, hi, lo);
      *reinterpret_cast<float4*>(c + off) = hi;
      *reinterpret_cast<float4*>(C_lo + off) = lo;
    } else {
      *reinterpret_cast<float4*>(c + off) = v;
    }
  }
  __device__ __forceinline__ void store1(float* c, size_t off, float x) const {
    if (SPLIT) {
      float hi, lo;
      split_tf32(x, hi, lo);
      c[off] = hi; C_lo[off] = lo;
    } else {
      c[off] = x;
    }
  }
};

// Logits epilogue of the deferred-normalisation schedule (option "exp_slab"): the slab receives U = exp(s - c_row) instead
// of the logits, c_row a per-example constant known before the product (the example's true-class logit), and the per-(row,
// half tile) partial holds (max U, sum U).  softmax - onehot is then U with one patched element per row times a per-row
// factor 1 / sum U, which the two gradient GEMMs apply to their small operand / result: no pass ever reads the slab back
// to normalise it (tensorflow_model.py:227-230).  max U is the range guard: the caller falls back to the two-pass
// schedule when some row's largest U leaves [kExpSlabMin, kExpSlabMax] (common.cuh).  SPLIT (3xTF32): U is written as its tf32 split.
template <bool PRECISE, bool SPLIT>
struct EpiExpSumT {
  struct State { float mx, sum; };
  float* C;
  float* C_lo;
  size_t ldc;
  const float* offset;   // [M] c_row
  float2* partial;       // [M, slots]; slot = 2 * n_tile + column half
  int slots;
  __device__ __forceinline__ void begin(State& st) const { st.mx = 0.f; st.sum = 0.f; }
  __device__ __forceinline__ void end(int m, int slot, int, bool row_ok, State& st) const {
    if (row_ok) partial[(size_t)m * slots + slot] = make_float2(st.mx, st.sum);
  }
  __device__ __forceinline__ float map(float x) const { return x; }
  __device__ __forceinline__ float* out(int) const { return C; }
  using Pre = EpiNoState;
  static constexpr int kRowBatch = 8;
  static constexpr bool kWideDrain = true;
  __device__ __forceinline__ void prefetch(const float*, size_t, Pre&) const {}
  __device__ __forceinline__ void store4(float* c, size_t off, float4 v, const Pre&) const {
    if (SPLIT) {
      float4 hi, lo;
      split_tf32(v, hi, lo);
      *reinterpret_cast<float4*>(c + off) = hi;
      *reinterpret_cast<float4*>(C_lo + off) = lo;
    } else {
      *reinterpret_cast<float4*>(c + off) = v;
    }
  }
  __device__ __forceinline__ void store1(float* c, size_t off, float x) const {
    if (SPLIT) {
      float hi, lo;
      split_tf32(x, hi, lo);
      c[off] = hi; C_lo[off] = lo;
    } else {
      c[off] = x;
    }
  }
  // rewrites the chunk's accumulators IN PLACE (the store that follows maps with the identity)
  __device__ __forceinline__ void observe(int m, int, uint32_t (&r)[32], int nvalid, State& st) const {
    const float c = offset[m];
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, x0 = 0.f, x1 = 0.f, x2 = 0.f, x3 = 0.f;
    if (!PRECISE && nvalid >= 32) {
      constexpr float L2E = 1.4426950408889634f;
      const float c2 = -c * L2E;           // exp(x - c) as one FFMA + one ex2.approx
      auto e2 = [](float t) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(t)); return y; };
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float u0 = e2(fmaf(__uint_as_float(r[j + 0]), L2E, c2)), u1 = e2(fmaf(__uint_as_float(r[j + 1]), L2E, c2));
        const float u2 = e2(fmaf(__uint_as_float(r[j + 2]), L2E, c2)), u3 = e2(fmaf(__uint_as_float(r[j + 3]), L2E, c2));
        r[j + 0] = __float_as_uint(u0); r[j + 1] = __float_as_uint(u1); r[j + 2] = __float_as_uint(u2); r[j + 3] = __float_as_uint(u3);
        a0 += u0; a1 += u1; a2 += u2; a3 += u3;
        x0 = fmaxf(x0, u0); x1 = fmaxf(x1, u1); x2 = fmaxf(x2, u2); x3 = fmaxf(x3, u3);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float u = 0.f;
        if (j < nvalid) u = PRECISE ? expf(__uint_as_float(r[j]) - c) : __expf(__uint_as_float(r[j]) - c);
        r[j] = __float_as_uint(u);
        if ((j & 3) == 0) { a0 += u; x0 = fmaxf(x0, u); }
        else if ((j & 3) == 1) { a1 += u; x1 = fmaxf(x1, u); }
        else if ((j & 3) == 2) { a2 += u; x2 = fmaxf(x2, u); }
        else { a3 += u; x3 = fmaxf(x3, u); }
      }
    }
    st.sum += (a0 + a1) + (a2 + a3);
    st.mx = fmaxf(st.mx, fmaxf(fmaxf(x0, x1), fmaxf(x2, x3)));
  }
};
// The two-pass fallback of that schedule is launched every step behind a device-side gate: a logits epilogue that carries
// `gate` makes the whole kernel return at once while *gate == 0.
template <bool PRECISE>
struct EpiStoreLseGatedT : EpiStoreLseT<PRECISE> {
  const int* gate;
};

using EpiTanhStore = EpiTanhStoreT<false>;
using EpiTanhStorePrecise = EpiTanhStoreT<true>;
using EpiStoreLse = EpiStoreLseT<false>;
using EpiStoreLsePrecise = EpiStoreLseT<true>;

// Target-table gradient epilogue with the optimizer folded in (option "fuse_target_adam"): the
// accumulator element is dYtab[y, j]; instead of writing it out for adam_kernel to read back, the
// epilogue applies TF1 Adam (tensorflow_model.py:232) to (theta, m, v) in place -- the
// same correctly rounded fp32 operations in the same order as adam_kernel, so the result is
// bit-identical; the gradient itself is never stored.
struct EpiAdam {
  using State = EpiNoState;
  float* P;          // theta [M, ldc]; m and v have the same layout
  float* Mo;
  float* Vo;
  size_t ldc;
  float lr_t, b1, b2, eps, omb1, omb2;
  int l2_prefetch;   // engine option "adam_epilogue_prefetch"
  __device__ __forceinline__ void begin(State&) const {}
  __device__ __forceinline__ void end(int, int, int, bool, State&) const {}
  __device__ __forceinline__ void observe(int, int, const uint32_t (&)[32], int, State&) const {}
  __device__ __forceinline__ float map(float x) const { return x; }
  __device__ __forceinline__ float* out(int) const { return P; }
  __device__ __forceinline__ void upd(float& pp, float gg, float& mm, float& vv) const {
    mm = __fadd_rn(__fmul_rn(mm, b1), __fmul_rn(omb1, gg));
    vv = __fadd_rn(__fmul_rn(vv, b2), __fmul_rn(omb2, __fmul_rn(gg, gg)));
    pp = adam_move_dense(pp, lr_t, mm, vv, eps);
  }
  // Called by an epilogue warp one tile AHEAD of the tile it is about to drain (row m, columns [n0, n0 + ncols)): pulls the
  // (theta, m, v) lines of that region into L2, so that the update's loads -- only kRowBatch rows of them in flight per
  // thread -- see L2 latency rather than DRAM latency.
  __device__ __forceinline__ void prefetch_tile(int m, int n0, int ncols, int M, int N) const {
    if (m >= M || !l2_prefetch) return;
    const size_t row = (size_t)m * ldc;
    for (int n = n0; n < n0 + ncols && n < N; n += 32) {
      asm volatile("prefetch.global.L2 [%0];" ::"l"(P + row + n));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(Mo + row + n));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(Vo + row + n));
    }
  }
  struct Pre { float4 p, m, v; };
  static constexpr int kRowBatch = 4;            // 8 rows in flight spill and measured slower (dY 0.77 vs 0.68 ms)
  static constexpr bool kWideDrain = false;      // 32 columns per tcgen05.ld: the update needs the registers
  __device__ __forceinline__ void prefetch(const float* c, size_t off, Pre& q) const {
    q.p = *reinterpret_cast<const float4*>(c + off);
    q.m = *reinterpret_cast<const float4*>(Mo + off);
    q.v = *reinterpret_cast<const float4*>(Vo + off);
  }
  __device__ __forceinline__ void store4(float* c, size_t off, float4 g, const Pre& q) const {
    float4 p = q.p, m = q.m, v = q.v;
    upd(p.x, g.x, m.x, v.x); upd(p.y, g.y, m.y, v.y); upd(p.z, g.z, m.z, v.z); upd(p.w, g.w, m.w, v.w);
    *reinterpret_cast<float4*>(c + off) = p;
    *reinterpret_cast<float4*>(Mo + off) = m;
    *reinterpret_cast<float4*>(Vo + off) = v;
  }
  __device__ __forceinline__ void store1(float* c, size_t off, float g) const {
    float p = c[off], m = Mo[off], v = Vo[off];
    upd(p, g, m, v);
    c[off] = p; Mo[off] = m; Vo[off] = v;
  }
};

constexpr int kEpiStageBytes = 32 * 32 * 4;      // one 32 x 32 fp32 block per epilogue warp

// One 32-column chunk: registers (thread = row) -> swizzled smem -> coalesced global stores.
// element-wise transform on the way out: functors with a (row, column)-dependent transform define map_at(x, column, state)
template <class Epi>
__device__ __forceinline__ auto epi_map(const Epi& epi, float x, int col, const typename Epi::State& st, int)
    -> decltype(epi.map_at(x, col, st)) {
  return epi.map_at(x, col, st);
}
template <class Epi>
__device__ __forceinline__ float epi_map(const Epi& epi, float x, int, const typename Epi::State&, long) {
  return epi.map(x);
}
template <class Epi>
__device__ __forceinline__ auto epi_gate_closed(const Epi& epi, int) -> decltype(*epi.gate == 0) { return *epi.gate == 0; }
template <class Epi>
__device__ __forceinline__ bool epi_gate_closed(const Epi&, long) { return false; }
template <class Epi>
__device__ __forceinline__ auto epi_prefetch_tile(const Epi& epi, int m, int n0, int ncols, int M, int N, int)
    -> decltype(epi.prefetch_tile(m, n0, ncols, M, N)) {
  epi.prefetch_tile(m, n0, ncols, M, N);
}
template <class Epi>
__device__ __forceinline__ void epi_prefetch_tile(const Epi&, int, int, int, int, int, long) {}
template <class Epi>
constexpr bool epi_stores(...) { return true; }
template <class Epi, bool V = Epi::kStores>
constexpr bool epi_stores(int) { return V; }

template <class Epi>
__device__ __forceinline__ void store_chunk(const Epi& epi, const typename Epi::State& est, const uint32_t (&r)[32], float* stage, int lane,
                                            int m_base, int n, int M, int N, float* cbase, size_t ldc) {
#pragma unroll
  for (int j4 = 0; j4 < 8; ++j4) {
    const int pos = j4 ^ (lane & 7);
    *reinterpret_cast<float4*>(stage + lane * 32 + pos * 4) =
        make_float4(epi_map(epi, __uint_as_float(r[4 * j4]), n + 4 * j4, est, 0), epi_map(epi, __uint_as_float(r[4 * j4 + 1]), n + 4 * j4 + 1, est, 0),
                    epi_map(epi, __uint_as_float(r[4 * j4 + 2]), n + 4 * j4 + 2, est, 0), epi_map(epi, __uint_as_float(r[4 * j4 + 3]), n + 4 * j4 + 3, est, 0));
  }
  __syncwarp();
  const int col4 = lane & 7;
  const int gn = n + col4 * 4;
  // two passes per batch of rows so that a read-modify-write epilogue has kBatch rows' loads in flight
  constexpr int kBatch = Epi::kRowBatch;
#pragma unroll
  for (int it0 = 0; it0 < 8; it0 += kBatch) {
    float4 v[kBatch];
    typename Epi::Pre pre[kBatch];
#pragma unroll
    for (int i = 0; i < kBatch; ++i) {
      const int row = (it0 + i) * 4 + (lane >> 3);
      v[i] = *reinterpret_cast<const float4*>(stage + row * 32 + ((col4 ^ (row & 7)) * 4));
      const int gm = m_base + row;
      if (gm < M && gn + 3 < N) epi.prefetch(cbase, (size_t)gm * ldc + gn, pre[i]);
    }
#pragma unroll
    for (int i = 0; i < kBatch; ++i) {
      const int gm = m_base + (it0 + i) * 4 + (lane >> 3);
      if (gm < M && gn < N) {
        const size_t off = (size_t)gm * ldc + gn;
        if (gn + 3 < N) {
          epi.store4(cbase, off, v[i], pre[i]);
        } else {
          epi.store1(cbase, off, v[i].x);
          if (gn + 1 < N) epi.store1(cbase, off + 1, v[i].y);
          if (gn + 2 < N) epi.store1(cbase, off + 2, v[i].z);
        }
      }
    }
  }
  __syncwarp();
}

// Drains this warp's share of one accumulator: columns [col0, col0 + ncols) of the tile, rows
// m_base .. m_base + 31 (this thread's row is m_base + lane); 64 columns per tcgen05.ld round trip.
template <class Epi>
__device__ __forceinline__ void drain_accumulator(const Epi& epi, typename Epi::State& est, uint32_t taddr, int col0, int ncols,
                                                  int m_base, int lane, int n_tile0, int M, int N, int sp, float* stage) {
  const int m = m_base + lane;
  float* cbase = epi.out(sp);
  int c = col0;
#pragma unroll 1
  for (; Epi::kWideDrain && c + 64 <= col0 + ncols; c += 64) {
    uint32_t r0[32], r1[32];
    tmem_ld64(taddr + c, r0, r1);
    tmem_ld_wait();
    const int n = n_tile0 + c;
    if (n < N) {
      if (m < M) epi.observe(m, n, r0, N - n, est);
      if (epi_stores<Epi>(0)) store_chunk(epi, est, r0, stage, lane, m_base, n, M, N, cbase, epi.ldc);
    }
    if (n + 32 < N) {
      if (m < M) epi.observe(m, n + 32, r1, N - n - 32, est);
      if (epi_stores<Epi>(0)) store_chunk(epi, est, r1, stage, lane, m_base, n + 32, M, N, cbase, epi.ldc);
    }
  }
#pragma unroll 1
  for (; c < col0 + ncols; c += 32) {
    uint32_t r[32];
    tmem_ld32(taddr + c, r);
    tmem_ld_wait();
    const int n = n_tile0 + c;
    if (n < N) {
      if (m < M) epi.observe(m, n, r, N - n, est);
      if (epi_stores<Epi>(0)) store_chunk(epi, est, r, stage, lane, m_base, n, M, N, cbase, epi.ldc);
    }
  }
}

// ---- A-operand transforms ---------------------------------------------------------------------------------
// Optional warps between TMA and the tensor core: they wait for a stage to land, rewrite the A tile IN PLACE in
// shared memory (element-wise, so the swizzled layout is untouched), fence the generic-proxy writes for the async
// proxy and only then release the stage to the MMA issuer.  Used to feed the two target-side gradient GEMMs with
// dL/dlogits = (softmax(S) - onehot) / B computed on the fly from the LOGITS slab S and the per-row log-sum-exp, so
// that the separate pass that rewrote the 1.07 GB slab (read + write) disappears:
//   dv = P . Ytab     A = S, K-major   (rows = examples, K = classes)      -> AXSoftmaxGradK
//   dY = P^T . v      A = S^T, MN-major (rows = classes, K = examples)     -> AXSoftmaxGradMN
struct AXNone {
  static constexpr int kWarps = 0;
  struct Pre {};
  __device__ __forceinline__ void load(int, int, int, int, int, Pre&) const {}
  __device__ __forceinline__ void apply(uint8_t*, int, int, int, int, int, const Pre&) const {}
};

struct SoftmaxGradArgs {
  const float* lse;          // [examples] log-sum-exp of each row of S (all classes, all ranks)
  const int32_t* target;     // [examples] true class (global id)
  int row0;                  // first class held in S (row-sharded target table), 0 otherwise
  float inv_batch;
};
// p = exp(x - lse) / B as ONE fused multiply-add and one ex2: 2^(x log2(e) + c), c = -lse log2(e) + log2(1/B).  No branches, so
// the 32 values of a row are independent instruction streams (a single warp transforms a whole stage: it needs the ILP).
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
constexpr float kLog2e = 1.4426950408889634f;
__device__ __forceinline__ float4 softmax_grad4(float4 x, float c) {
  return make_float4(ex2_approx(fmaf(x.x, kLog2e, c)), ex2_approx(fmaf(x.y, kLog2e, c)), ex2_approx(fmaf(x.z, kLog2e, c)),
                     ex2_approx(fmaf(x.w, kLog2e, c)));
}
__device__ __forceinline__ float4 softmax_grad4_tail(float4 x, float c, int y0, int ymax) {     // classes >= ymax are zero fill: keep 0
  float4 p = softmax_grad4(x, c);
  p.x = (y0 + 0 < ymax) ? p.x : 0.f; p.y = (y0 + 1 < ymax) ? p.y : 0.f;
  p.z = (y0 + 2 < ymax) ? p.z : 0.f; p.w = (y0 + 3 < ymax) ? p.w : 0.f;
  return p;
}
// ONE warp rewrites a whole stage (warp w owns every kWarps-th step of the CTA's k-block stream), so kWarps stages are
// being transformed at any time.  load() runs BEFORE the wait for the stage: the per-example exponent offset and target are in
// registers by the time the tile has landed.  The "- onehot / B" term touches at most one element per example row and is
// applied to that element after the row has been rewritten.
//
// A tile = BM example rows x BK classes, K-major SWIZZLE_128B: row r at r*128, 16-byte chunk c at (c ^ (r & 7)) * 16.
// Lane l rewrites rows l, l + 32, l + 64, l + 96.
template <int NW>
struct AXSoftmaxGradK {
  static constexpr int kWarps = NW;
  SoftmaxGradArgs a;
  struct Pre { float c[BM / 32]; int tgt[BM / 32]; };
  __device__ __forceinline__ void load(int lane, int m0, int, int M, int, Pre& p) const {
    const float lb = log2f(a.inv_batch);
#pragma unroll
    for (int j = 0; j < BM / 32; ++j) {
      const int b = m0 + lane + 32 * j;
      p.c[j] = (b < M) ? fmaf(-a.lse[b], kLog2e, lb) : 0.f;
      p.tgt[j] = (b < M) ? a.target[b] - a.row0 : -1;
    }
  }
  __device__ __forceinline__ void apply(uint8_t* sa, int lane, int m0, int k0, int M, int K, const Pre& p) const {
    const bool full = k0 + BK <= K;
#pragma unroll
    for (int j = 0; j < BM / 32; ++j) {
      const int r = lane + 32 * j;
      if (m0 + r >= M) continue;                             // rows past the batch: TMA zero fill, results are discarded
      uint8_t* row = sa + r * 128;
      float4 v[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) v[c] = *reinterpret_cast<const float4*>(row + ((c ^ (r & 7)) << 4));
      if (full) {
#pragma unroll
        for (int c = 0; c < 8; ++c) v[c] = softmax_grad4(v[c], p.c[j]);
      } else {
#pragma unroll
        for (int c = 0; c < 8; ++c) v[c] = softmax_grad4_tail(v[c], p.c[j], k0 + c * 4, K);
      }
#pragma unroll
      for (int c = 0; c < 8; ++c) *reinterpret_cast<float4*>(row + ((c ^ (r & 7)) << 4)) = v[c];
      const int e = p.tgt[j] - k0;                           // the example's own class, if it is in this tile: - 1/B
      if (e >= 0 && e < BK) *(reinterpret_cast<float*>(row + (((e >> 2) ^ (r & 7)) << 4)) + (e & 3)) -= a.inv_batch;
    }
  }
};
// A tile = BM classes x BK examples, MN-major SWIZZLE_128B_BASE32B: BM/32 blocks of [BK example rows x 128 B]; in a row the
// 32-byte unit u (8 classes) sits at (u ^ (row & 3)) * 32.  Lane l rewrites example row l of every block.
template <int NW>
struct AXSoftmaxGradMN {
  static constexpr int kWarps = NW;
  SoftmaxGradArgs a;
  struct Pre { float c; int tgt; };
  __device__ __forceinline__ void load(int lane, int, int k0, int, int K, Pre& p) const {
    const int b = k0 + lane;
    p.c = (b < K) ? fmaf(-a.lse[b], kLog2e, log2f(a.inv_batch)) : 0.f;
    p.tgt = (b < K) ? a.target[b] - a.row0 : -1;
  }
  __device__ __forceinline__ void apply(uint8_t* sa, int lane, int m0, int k0, int M, int K, const Pre& p) const {
    static_assert(BK == 32, "one example row per lane");
    if (k0 + lane >= K) return;                              // examples past the batch: zero fill stays zero
    const bool full = m0 + BM <= M;
#pragma unroll
    for (int ch = 0; ch < BM / 32; ++ch) {
      uint8_t* row = sa + ch * (BK * 128) + lane * 128;
      float4 v[8];                                           // v[2u + h] = classes m0 + 32 ch + 8 u + 4 h ..+3
#pragma unroll
      for (int q = 0; q < 8; ++q) v[q] = *reinterpret_cast<const float4*>(row + (((q >> 1) ^ (lane & 3)) << 5) + ((q & 1) << 4));
      if (full) {
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = softmax_grad4(v[q], p.c);
      } else {
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = softmax_grad4_tail(v[q], p.c, m0 + ch * 32 + q * 4, M);
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) *reinterpret_cast<float4*>(row + (((q >> 1) ^ (lane & 3)) << 5) + ((q & 1) << 4)) = v[q];
      const int e = p.tgt - (m0 + ch * 32);
      if (e >= 0 && e < 32)
        *(reinterpret_cast<float*>(row + (((e >> 3) ^ (lane & 3)) << 5) + (((e >> 2) & 1) << 4)) + (e & 3)) -= a.inv_batch;
    }
  }
};

// ---- kernel -------------------------------------------------------------------------------------------
struct GemmShape {
  int M, N, K;
  int m_tiles, n_tiles, splits;
  int kblocks_per_split;      // K blocks (of BK) per split-K slice
  int n_fastest;              // raster order of work items: 1 = consecutive items share the A tile
  int terms;                  // 1 = plain tf32;  3 = 3xTF32: every K block is issued as A_lo.B_hi + A_hi.B_lo + A_hi.B_hi
};

template <int BN, int STAGES>
struct SmemLayout {
  static constexpr int kABytes = BM * BK * 4;          // 16 KB
  static constexpr int kBBytes = BN * BK * 4;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kEpiStageOffset = STAGES * kStageBytes;                 // 8 x 4 KB store-transpose blocks
  static constexpr int kBarOffset = kEpiStageOffset + kEpiWarps * kEpiStageBytes;
  static constexpr int kTotal = kBarOffset + 256 + 1024;   // barriers + tmem ptr, + slack for 1024-B alignment
  static_assert(kTotal <= 232448, "exceeds the 227 KB of shared memory a CTA can opt into");
};

template <int BN, int STAGES, bool A_MN, bool B_MN, class Epi, class AX = AXNone>
__global__ void __launch_bounds__(kThreads + 32 * AX::kWarps, 1)
umma_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmAlo, const __grid_constant__ CUtensorMap tmBlo, GemmShape gs, Epi epi,
                 AX ax = AX{}) {
  if (epi_gate_closed(epi, 0)) return;      // gated fallback pass: nothing to do (uniform over the grid)
  using L = SmemLayout<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;      // [2]
  uint64_t* tempty_bar = tfull_bar + 2;          // [2]
  uint64_t* xf_bar = tempty_bar + 2;             // [STAGES] A tile transformed (only with an A transform)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(xf_bar + STAGES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr uint32_t kTmemCols = (2 * BN <= 256) ? 256 : 512;
  constexpr bool kXform = AX::kWarps > 0;
  const int total_items = gs.m_tiles * gs.n_tiles * gs.splits;
  const int total_kblocks = (gs.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1);
      if (kXform) mbar_init(&xf_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], kEpiWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_ptr, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  // work item -> (m tile, n tile, split): m fastest so CTAs running together share B tiles in L2
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
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
        int mt, nt, sp;
        decode(item, mt, nt, sp);
        const int kb0 = sp * gs.kblocks_per_split;
        const int kb1 = min(total_kblocks, kb0 + gs.kblocks_per_split);
        // 3xTF32: the two small cross terms (A_lo.B_hi, A_hi.B_lo) over the whole K range first, then A_hi.B_hi.
        // The tensor core adds into the fp32 accumulator with truncation, ~half an ulp of the accumulator per
        // MMA; while only the 2^-11-sized cross terms have been added that loss is negligible, so the
        // accumulation error is that of ONE pass over K instead of three.
        auto load_stage = [&](const CUtensorMap* mA, const CUtensorMap* mB, int kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * L::kStageBytes;
          uint8_t* sb = sa + L::kABytes;
          mbar_expect_tx(&full_bar[stage], L::kStageBytes);
          if (A_MN) {
            // global A^T is [K rows, M contiguous]: boxes of {32 m, BK k-rows} (4 KB each)
#pragma unroll
            for (int c = 0; c < BM / 32; ++c) tma_load_2d(sa + c * (BK * 128), mA, &full_bar[stage], mt * BM + c * 32, kb * BK);
          } else {
            // global A is [M rows, K contiguous]: one box {BK k, BM rows}
            tma_load_2d(sa, mA, &full_bar[stage], kb * BK, mt * BM);
          }
          if (B_MN) {
#pragma unroll
            for (int c = 0; c < BN / 32; ++c) tma_load_2d(sb + c * (BK * 128), mB, &full_bar[stage], nt * BN + c * 32, kb * BK);
          } else {
            tma_load_2d(sb, mB, &full_bar[stage], kb * BK, nt * BN);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        };
        if (gs.terms == 3) {
          for (int kb = kb0; kb < kb1; ++kb) load_stage(&tmAlo, &tmB, kb);
          for (int kb = kb0; kb < kb1; ++kb) load_stage(&tmA, &tmBlo, kb);
        }
        for (int kb = kb0; kb < kb1; ++kb) load_stage(&tmA, &tmB, kb);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_tf32(BM, BN, A_MN, B_MN);
      // K-major (SWIZZLE_128B): rows of 128 B, 8-row groups 1024 B apart (SBO); one MMA consumes
      //   32 B of each row, so the K step is +32 B inside the swizzle row.
      // MN-major (SWIZZLE_128B_BASE32B): chunks of 32 M/N-elements x BK k-rows (128 B per k-row);
      //   chunks are BK*128 B apart (LBO), 4-k-row swizzle atoms 512 B apart (SBO); one MMA consumes
      //   8 k-rows, so the K step is +1024 B.
      constexpr uint32_t a_lbo = A_MN ? BK * 128 : 0, b_lbo = B_MN ? BK * 128 : 0;
      constexpr uint32_t a_sbo = A_MN ? 512 : 1024, b_sbo = B_MN ? 512 : 1024;
      constexpr uint32_t a_lt = A_MN ? kLayoutSw128Base32 : kLayoutSw128, b_lt = B_MN ? kLayoutSw128Base32 : kLayoutSw128;
      constexpr uint32_t a_kstep = A_MN ? 1024 : UMMA_K * 4, b_kstep = B_MN ? 1024 : UMMA_K * 4;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
        int mt, nt, sp;
        decode(item, mt, nt, sp);
        const int kb0 = sp * gs.kblocks_per_split;
        const int kb1 = min(total_kblocks, kb0 + gs.kblocks_per_split);
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);          // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        const int n_steps = (kb1 - kb0) * gs.terms;          // 3xTF32: three passes over the K range (see the producer)
        for (int it = 0; it < n_steps; ++it) {
          mbar_wait(kXform ? &xf_bar[stage] : &full_bar[stage], phase);    // with a transform: once the A tile has been rewritten
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * L::kStageBytes);
          const uint32_t sb = sa + L::kABytes;
          const uint64_t adesc = make_smem_desc(sa, a_lbo, a_sbo, a_lt);
          const uint64_t bdesc = make_smem_desc(sb, b_lbo, b_sbo, b_lt);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            mma_tf32(tmem_d, adesc + (uint64_t)((k * a_kstep) >> 4), bdesc + (uint64_t)((k * b_kstep) >> 4), idesc,
                     (it > 0 || k > 0) ? 1u : 0u);
          }
          tc_commit(&empty_bar[stage]);                       // frees the smem stage when the MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        tc_commit(&tfull_bar[acc]);                           // accumulator complete -> epilogue
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (kXform && warp >= kEpiWarp0 + kEpiWarps) {
    // ===================== A-transform warps =====================
    // warp w takes every kWarps-th step of this CTA's (item, k-block) stream; step g lives in stage g % STAGES
    constexpr int kXW = kXform ? AX::kWarps : 1;
    static_assert(STAGES % kXW == 0, "a transform warp must always meet the same stages");
    const int xw = warp - kEpiWarp0 - kEpiWarps;
    int g = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      int mt, nt, sp;
      decode(item, mt, nt, sp);
      const int kb0 = sp * gs.kblocks_per_split;
      const int kb1 = min(total_kblocks, kb0 + gs.kblocks_per_split);
      for (int kb = kb0; kb < kb1; ++kb, ++g) {
        if (g % kXW != xw) continue;
        const int stage = g % STAGES;
        const uint32_t phase = (uint32_t)(g / STAGES) & 1u;
        typename AX::Pre pre;
        ax.load(lane, mt * BM, kb * BK, gs.M, gs.K, pre);     // per-example scalars: in flight while the tile lands
        mbar_wait(&full_bar[stage], phase);                   // the TMA tiles of this stage have landed
        ax.apply(smem + stage * L::kStageBytes, lane, mt * BM, kb * BK, gs.M, gs.K, pre);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&xf_bar[stage]);
      }
    }
  } else {
    // ===================== epilogue warps (TMEM lane quadrant = warp % 4) =====================
    const int q = warp & 3;                       // TMEM lane quadrant this warp may access
    const int half = (warp - kEpiWarp0) >> 2;     // which half of the tile's columns it drains
    constexpr int kChunksPerHalf = BN / 64;
    int acc = 0;
    uint32_t acc_phase = 0;
    auto prefetch_item = [&](int it) {            // read-modify-write epilogues: warm L2 with the tile's destination lines
      if (it >= total_items) return;
      int mt2, nt2, sp2;
      decode(it, mt2, nt2, sp2);
      epi_prefetch_tile(epi, mt2 * BM + q * 32 + lane, nt2 * BN + half * kChunksPerHalf * 32, kChunksPerHalf * 32, gs.M, gs.N, 0);
    };
    prefetch_item(blockIdx.x);
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      int mt, nt, sp;
      decode(item, mt, nt, sp);
      prefetch_item(item + gridDim.x);
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const int m = mt * BM + q * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
      typename Epi::State est;
      epi.begin(est);
      drain_accumulator(epi, est, taddr, half * kChunksPerHalf * 32, kChunksPerHalf * 32, mt * BM + q * 32, lane, nt * BN,
                        gs.M, gs.N, sp, reinterpret_cast<float*>(smem + L::kEpiStageOffset + (warp - kEpiWarp0) * kEpiStageBytes));
      epi.end(m, 2 * nt + half, sp, m < gs.M, est);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

// ---- host side ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && p)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D fp32 tensor map over a row-major [rows, cols] matrix with row pitch ld (floats); box = {32 cols, box_rows}.
// Encoded maps are kept in a small per-thread cache: a training step issues the same dozen (buffer, shape)
// combinations every time, so after the first step no launch calls into the driver for a descriptor.
struct TensorMapKey {
  const float* base;
  uint64_t rows, cols, ld;
  uint32_t box_rows;
  bool atom32;
  bool operator==(const TensorMapKey& o) const {
    return base == o.base && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows && atom32 == o.atom32;
  }
};
struct TensorMapCache {
  static constexpr int kCap = 96;
  TensorMapKey key[kCap];
  CUtensorMap map[kCap];
  int n = 0, next = 0;
};
inline bool make_tensor_map(CUtensorMap* map, const float* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                            bool atom32) {
  static thread_local TensorMapCache cache;
  const TensorMapKey k{base, rows, cols, ld, box_rows, atom32};
  for (int i = 0; i < cache.n; ++i)
    if (cache.key[i] == k) { *map = cache.map[i]; return true; }
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * sizeof(float)};
  cuuint32_t box[2] = {32, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return false;
  const int slot = (cache.n < TensorMapCache::kCap) ? cache.n++ : (cache.next++ % TensorMapCache::kCap);
  cache.key[slot] = k;
  cache.map[slot] = *map;
  return true;
}

// Operand description: `major_mn == false`: element (x, k) at base[x*ld + k] (K contiguous);
//                      `major_mn == true` : element (x, k) at base[k*ld + x] (M / N contiguous).
// 3xTF32 (C2V_MATH_3XTF32): `base` holds the tf32-rounded high parts and `lo` (same layout) the tf32-rounded
// residuals x - hi; both operands of a product must carry one, or neither.
struct Operand {
  const float* base;
  size_t ld;
  bool major_mn;
  const float* lo = nullptr;
};

template <int BN, int STAGES, bool A_MN, bool B_MN, class Epi, class AX = AXNone>
inline cudaError_t launch_cfg(cudaStream_t st, int M, int N, int K, int splits, const Operand& A, const Operand& B, const Epi& epi,
                              int num_sms, const AX& ax = AX{}) {
  using L = SmemLayout<BN, STAGES>;
  CUtensorMap tmA, tmB, tmAlo, tmBlo;
  const bool three = A.lo != nullptr && B.lo != nullptr;
  if ((A.lo != nullptr) != (B.lo != nullptr)) return cudaErrorInvalidValue;
  auto mapA = [&](CUtensorMap* m, const float* base) {
    return A_MN ? make_tensor_map(m, base, (uint64_t)K, (uint64_t)M, A.ld, BK, true)
                : make_tensor_map(m, base, (uint64_t)M, (uint64_t)K, A.ld, BM, false);
  };
  auto mapB = [&](CUtensorMap* m, const float* base) {
    return B_MN ? make_tensor_map(m, base, (uint64_t)K, (uint64_t)N, B.ld, BK, true)
                : make_tensor_map(m, base, (uint64_t)N, (uint64_t)K, B.ld, BN, false);
  };
  if (!mapA(&tmA, A.base) || !mapB(&tmB, B.base)) return cudaErrorInvalidValue;
  if (three) { if (!mapA(&tmAlo, A.lo) || !mapB(&tmBlo, B.lo)) return cudaErrorInvalidValue; }
  else { tmAlo = tmA; tmBlo = tmB; }
  GemmShape gs;
  gs.M = M; gs.N = N; gs.K = K;
  gs.terms = three ? 3 : 1;
  gs.m_tiles = (M + BM - 1) / BM;
  gs.n_tiles = (N + BN - 1) / BN;
  const int total_kblocks = (K + BK - 1) / BK;
  if (splits < 1) splits = 1;
  if (splits > total_kblocks) splits = total_kblocks;
  gs.kblocks_per_split = (total_kblocks + splits - 1) / splits;
  gs.splits = (total_kblocks + gs.kblocks_per_split - 1) / gs.kblocks_per_split;
  // few, wide n-tiles under many m-tiles: walk n fastest so the (large) A tile is fetched from HBM once
  gs.n_fastest = (gs.n_tiles < gs.m_tiles) ? 1 : 0;
  if (AX::kWarps > 0 && three) return cudaErrorInvalidValue;      // the transforms rewrite a single fp32 tile
  auto kern = umma_gemm_kernel<BN, STAGES, A_MN, B_MN, Epi, AX>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal);
  if (e != cudaSuccess) return e;
  int grid = gs.m_tiles * gs.n_tiles * gs.splits;
  if (grid > num_sms) grid = num_sms;
  kern<<<grid, kThreads + 32 * AX::kWarps, L::kTotal, st>>>(tmA, tmB, tmAlo, tmBlo, gs, epi, ax);
  return cudaGetLastError();
}

// number of split-K slices launch_cfg will actually produce
inline int effective_splits(int K, int splits) {
  const int total_kblocks = (K + BK - 1) / BK;
  if (splits < 1) splits = 1;
  if (splits > total_kblocks) splits = total_kblocks;
  const int per = (total_kblocks + splits - 1) / splits;
  return (total_kblocks + per - 1) / per;
}

// Runtime dispatch over operand majors.  BN = 256 and BN = 192 both run 4 stages (192 / 160 KB) next to
// the 32 KB of epilogue transpose blocks.
template <int BN, int STAGES, class Epi>
inline cudaError_t launch(cudaStream_t st, int M, int N, int K, int splits, const Operand& A, const Operand& B, const Epi& epi,
                          int num_sms) {
  if (!A.major_mn && !B.major_mn) return launch_cfg<BN, STAGES, false, false, Epi>(st, M, N, K, splits, A, B, epi, num_sms);
  if (!A.major_mn && B.major_mn) return launch_cfg<BN, STAGES, false, true, Epi>(st, M, N, K, splits, A, B, epi, num_sms);
  if (A.major_mn && !B.major_mn) return launch_cfg<BN, STAGES, true, false, Epi>(st, M, N, K, splits, A, B, epi, num_sms);
  return launch_cfg<BN, STAGES, true, true, Epi>(st, M, N, K, splits, A, B, epi, num_sms);
}

// TMA constraints on an operand: 16-byte aligned base, row pitch a multiple of 16 bytes.
inline bool operand_ok(const Operand& o) {
  return (reinterpret_cast<uintptr_t>(o.base) % 16 == 0) && (o.ld % 4 == 0);
}

}  // namespace umma
}  // namespace c2v
