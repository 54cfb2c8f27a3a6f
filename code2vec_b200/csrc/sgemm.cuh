// fp32 SIMT GEMM (C2V_MATH_FP32): the reference's own arithmetic class (fp32 FMA, as cuBLAS /
// Eigen SGEMM run tensorflow_model.py:226,252,297), used for bit-level top-k parity and as the
// shape-general path.  128x128x8 CTA tile, 8x8 register micro-tile, register-staged double
// buffering.  Operands come through small loader functors so that the gathered context matrix
// (three embedding lookups + concat + dropout, tensorflow_model.py:238-246) is never
// materialised, and results leave through epilogue functors (tanh, scatter-add, split-K).
#pragma once
#include "common.cuh"

namespace c2v {
namespace simt {

constexpr int BM = 128, BN = 128, BK = 8, NT = 256;
constexpr int LDT = 128 + 4;   // padded smem row: keeps float4 alignment, breaks store conflicts

__device__ __forceinline__ float4 ld4_guard(const float* p, int nvalid) {
  if (nvalid >= 4) return *reinterpret_cast<const float4*>(p);
  float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
  if (nvalid > 0) r.x = p[0];
  if (nvalid > 1) r.y = p[1];
  if (nvalid > 2) r.z = p[2];
  return r;
}

// ----- operand loaders: each fills a [BK][128] slice T[kk][x] of the operand tile --------------
// fetch(): global -> one float4 per thread;  stash(): that float4 -> shared memory.

// element (x, k) at p[x*ld + k]  (K contiguous: row-major A[M,K] or B^T stored [N,K])
struct RowsK {
  const float* p;
  size_t ld;
  __device__ __forceinline__ void fetch(float4& r, int x0, int k, int tid, int X, int kend) const {
    const int x = x0 + (tid >> 1), kc = k + ((tid & 1) << 2);
    r = (x < X) ? ld4_guard(p + (size_t)x * ld + kc, kend - kc) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __device__ __forceinline__ void stash(float (*T)[LDT], const float4& r, int tid) const {
    const int x = tid >> 1, kq = (tid & 1) << 2;
    T[kq + 0][x] = r.x; T[kq + 1][x] = r.y; T[kq + 2][x] = r.z; T[kq + 3][x] = r.w;
  }
};

// element (x, k) at p[k*ld + x]  (X contiguous: B[K,N] row-major, or A^T stored [K,M])
struct ColsX {
  const float* p;
  size_t ld;
  __device__ __forceinline__ void fetch(float4& r, int x0, int k, int tid, int X, int kend) const {
    const int kr = k + (tid >> 5), x = x0 + ((tid & 31) << 2);
    r = (kr < kend) ? ld4_guard(p + (size_t)kr * ld + x, X - x) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __device__ __forceinline__ void stash(float (*T)[LDT], const float4& r, int tid) const {
    *reinterpret_cast<float4*>(&T[tid >> 5][(tid & 31) << 2]) = r;
  }
};

// A = gathered context matrix X'[n, j] (after dropout), rows n along M, K = 3d contiguous.
struct GatherRowsK {
  ContextSource cs;
  Dropout dp;
  __device__ __forceinline__ void fetch(float4& r, int x0, int k, int tid, int X, int kend) const {
    const int n = x0 + (tid >> 1), j = k + ((tid & 1) << 2);
    if (n < X && j < kend) {
      r = *reinterpret_cast<const float4*>(ctx_ptr(cs, n, j));
      const float4 m = dropout_mult4(dp, n, j >> 2);
      r.x *= m.x; r.y *= m.y; r.z *= m.z; r.w *= m.w;
    } else {
      r = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  __device__ __forceinline__ void stash(float (*T)[LDT], const float4& r, int tid) const {
    const int x = tid >> 1, kq = (tid & 1) << 2;
    T[kq + 0][x] = r.x; T[kq + 1][x] = r.y; T[kq + 2][x] = r.z; T[kq + 3][x] = r.w;
  }
};

// A = X'^T: element (x = j, k = n) = X'[n, j]; the 3d columns are the contiguous extent.
struct GatherColsX {
  ContextSource cs;
  Dropout dp;
  __device__ __forceinline__ void fetch(float4& r, int x0, int k, int tid, int X, int kend) const {
    const int n = k + (tid >> 5), j = x0 + ((tid & 31) << 2);
    if (n < kend && j < X) {
      r = *reinterpret_cast<const float4*>(ctx_ptr(cs, n, j));
      const float4 m = dropout_mult4(dp, n, j >> 2);
      r.x *= m.x; r.y *= m.y; r.z *= m.z; r.w *= m.w;
    } else {
      r = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  __device__ __forceinline__ void stash(float (*T)[LDT], const float4& r, int tid) const {
    *reinterpret_cast<float4*>(&T[tid >> 5][(tid & 31) << 2]) = r;
  }
};

// ----- epilogues: called once per (row m, 4 consecutive columns n..n+3), nvalid = N - n ---------

__device__ __forceinline__ void st4_guard(float* p, const float4& v, int nvalid) {
  if (nvalid >= 4) { *reinterpret_cast<float4*>(p) = v; return; }
  if (nvalid > 0) p[0] = v.x;
  if (nvalid > 1) p[1] = v.y;
  if (nvalid > 2) p[2] = v.z;
}

// C[m, n] = acc ; with split-K, slice blockIdx.z goes to C + z*split_stride (reduced afterwards)
struct StoreC {
  float* C;
  size_t ldc;
  size_t split_stride;
  __device__ __forceinline__ void operator()(int m, int n, const float4& v, int nvalid) const {
    st4_guard(C + (size_t)blockIdx.z * split_stride + (size_t)m * ldc + n, v, nvalid);
  }
};

// H[m, n] = tanh(acc)      (tensorflow_model.py:252)
struct TanhStore {
  float* H;
  size_t ldh;
  __device__ __forceinline__ void operator()(int m, int n, const float4& v, int nvalid) const {
    st4_guard(H + (size_t)m * ldh + n, make_float4(tanhf(v.x), tanhf(v.y), tanhf(v.z), tanhf(v.w)), nvalid);
  }
};

// dX'[n, j] -> dropout backward -> scatter-add into the token / path gradient tables.
// (autodiff of the three tf.nn.embedding_lookup calls: IndexedSlices summed densely.)
// Masked contexts carry exact zeros (alpha == 0) and are skipped.
struct ScatterDx {
  ContextSource cs;      // indices (the parameter tables in it are not used here)
  ShardedTable g_tok;    // gradient tables (possibly peer shards)
  ShardedTable g_path;
  const float* mask;     // [rows]
  Dropout dp;
  float grad_scale;      // 1/world under data parallelism with sharded tables, else 1
  __device__ __forceinline__ void operator()(int m, int n, const float4& v, int nvalid) const {
    if (nvalid < 4) return;                 // 3d % 4 == 0: never partial
    if (mask[m] == 0.f) return;
    const float4 mu = dropout_mult4(dp, m, n >> 2);
    const float4 g = make_float4(v.x * mu.x * grad_scale, v.y * mu.y * grad_scale, v.z * mu.z * grad_scale,
                                 v.w * mu.w * grad_scale);
    const int seg = n / cs.d, off = n - seg * cs.d;
    float* dst;
    if (seg == 0) dst = table_row(g_tok, cs.src[m], cs.d) + off;
    else if (seg == 1) dst = table_row(g_path, cs.pth[m], cs.d) + off;
    else dst = table_row(g_tok, cs.tgt[m], cs.d) + off;
    atomicAdd(reinterpret_cast<float4*>(dst), g);     // red.global.add.v4.f32 (sm_90+)
  }
};

// ----- the kernel ------------------------------------------------------------------------------
// C[M,N] (+)= A[M,K] . B[K,N] over k in [z*kchunk, min(K,(z+1)*kchunk)), z = blockIdx.z.
template <class AL, class BL, class EP>
__global__ void __launch_bounds__(NT, 2)
sgemm_kernel(int M, int N, int K, int kchunk, const __grid_constant__ AL al, const __grid_constant__ BL bl,
             const __grid_constant__ EP ep) {
  __shared__ __align__(16) float As[2][BK][LDT];
  __shared__ __align__(16) float Bs[2][BK][LDT];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int kbeg = blockIdx.z * kchunk;
  const int kend = min(K, kbeg + kchunk);
  const int ty = tid >> 4, tx = tid & 15;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float4 ra, rb;
  if (kbeg < kend) {
    al.fetch(ra, m0, kbeg, tid, M, kend);
    bl.fetch(rb, n0, kbeg, tid, N, kend);
    al.stash(As[0], ra, tid);
    bl.stash(Bs[0], rb, tid);
  }
  __syncthreads();
  int buf = 0;
  for (int k = kbeg; k < kend; k += BK) {
    const bool more = (k + BK) < kend;
    if (more) {
      al.fetch(ra, m0, k + BK, tid, M, kend);
      bl.fetch(rb, n0, k + BK, tid, N, kend);
    }
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (more) {
      al.stash(As[buf ^ 1], ra, tid);
      bl.stash(Bs[buf ^ 1], rb, tid);
    }
    __syncthreads();
    buf ^= 1;
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= M) continue;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int n = n0 + h * 64 + tx * 4;
      if (n >= N) continue;
      ep(m, n, make_float4(acc[i][h * 4 + 0], acc[i][h * 4 + 1], acc[i][h * 4 + 2], acc[i][h * 4 + 3]), N - n);
    }
  }
}

template <class AL, class BL, class EP>
inline cudaError_t launch(cudaStream_t st, int M, int N, int K, int ksplit, const AL& al, const BL& bl, const EP& ep) {
  if (M <= 0 || N <= 0) return cudaSuccess;
  if (ksplit < 1) ksplit = 1;
  int kchunk = (K + ksplit - 1) / ksplit;
  kchunk = (kchunk + BK - 1) / BK * BK;
  if (kchunk < BK) kchunk = BK;
  ksplit = (K + kchunk - 1) / kchunk;
  if (ksplit < 1) ksplit = 1;
  dim3 grid((M + BM - 1) / BM, (N + BN - 1) / BN, ksplit);
  sgemm_kernel<AL, BL, EP><<<grid, NT, 0, st>>>(M, N, K, kchunk, al, bl, ep);
  return cudaGetLastError();
}

// number of k-slices launch() will actually use for (K, requested ksplit)
inline int effective_ksplit(int K, int ksplit) {
  if (ksplit < 1) ksplit = 1;
  int kchunk = (K + ksplit - 1) / ksplit;
  kchunk = (kchunk + BK - 1) / BK * BK;
  if (kchunk < BK) kchunk = BK;
  int s = (K + kchunk - 1) / kchunk;
  return s < 1 ? 1 : s;
}

}  // namespace simt
}  // namespace c2v
