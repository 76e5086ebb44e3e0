// Fused fp32 attention, tensor-core generation (default): same three kernels and the same argument block as
// mrg_attention.cu (flash-style forward, dQ, dK/dV; functional mask), but every 64x64xd product runs on the tensor
// cores as 3xTF32 — A.B ~ A_lo.B_hi + A_hi.B_lo + A_hi.B_hi with hi = tf32(x), lo = tf32(x - hi), fp32 accumulate —
// the split that keeps the projection GEMMs inside the 1e-5 parity budget.  The CUDA-core kernels of mrg_attention.cu
// reach 22 TFLOP/s (29 % of the fp32 FMA peak, the same as torch's sm_80 kernels); the warp-level mma path has ~10x
// the FMA rate, so the 3x work of the split still leaves a large margin.
//
// Structure per 64x64 tile: warp-level fragments (nvcuda::wmma m16n16k8, precision::tf32) for the products, the
// score tile staged in shared memory for the element-wise part (mask, online softmax, dS), operand tiles in their
// NATURAL [row][d] layout (cp.async, double-buffered, no transposed copies): Q.K^T and dO.V^T read K / V as
// col_major B fragments, P.V and dS.K read them as row_major B, P^T.dO and dS^T.Q read the score tile as col_major A.
// Accumulator fragments have an opaque element order; the row of each element is obtained once by loading a
// matrix whose entries are their own row index, which makes the per-row rescale of the online softmax legal.
#include <mma.h>

#include <cstddef>
#include <cstdlib>

#include "mrg_common.cuh"

namespace mrg {

using namespace nvcuda;

constexpr int TC_T = 64;         // queries / keys per tile
constexpr int TC_LDS = 72;       // row stride of the score tiles (floats; multiple of 8: 32-byte aligned fragments)
constexpr int TC_THREADS = 256;  // 8 warps: (row strip = warp & 3, column half = warp >> 2)

typedef wmma::fragment<wmma::matrix_a, 16, 16, 8, wmma::precision::tf32, wmma::row_major> FragA;
typedef wmma::fragment<wmma::matrix_a, 16, 16, 8, wmma::precision::tf32, wmma::col_major> FragAT;
typedef wmma::fragment<wmma::matrix_b, 16, 16, 8, wmma::precision::tf32, wmma::row_major> FragB;
typedef wmma::fragment<wmma::matrix_b, 16, 16, 8, wmma::precision::tf32, wmma::col_major> FragBT;
typedef wmma::fragment<wmma::accumulator, 16, 16, 8, float> FragC;

template <class F>
__device__ __forceinline__ void tc_split(F& hi, F& lo) {  // hi holds raw fp32 on entry
#pragma unroll
  for (int e = 0; e < hi.num_elements; ++e) {
    const float v = hi.x[e];
    const float h = wmma::__float_to_tf32(v);
    hi.x[e] = h;
    lo.x[e] = wmma::__float_to_tf32(v - h);
  }
}
template <class FA, class FB>
__device__ __forceinline__ void tc_mma3(FragC& c, const FA& ahi, const FA& alo, const FB& bhi, const FB& blo) {
  wmma::mma_sync(c, alo, bhi, c);
  wmma::mma_sync(c, ahi, blo, c);
  wmma::mma_sync(c, ahi, bhi, c);
}

__device__ __forceinline__ void tc_cp16(float* dst, const float* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void tc_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tc_wait0() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void tc_wait1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

// dst[r][0..HD) (row stride HD + 8) <- src[(r0 + r) * ld + ..]; rows past nrows are zero
template <int HD>
__device__ __forceinline__ void tc_load_rows(float* dst, const float* __restrict__ src, int ld, int r0, int nrows) {
  constexpr int C4 = HD / 4, LDH = HD + 8;
  for (int f = threadIdx.x; f < TC_T * C4; f += TC_THREADS) {
    const int r = f / C4, c = f % C4;
    float* d = dst + r * LDH + c * 4;
    if (r0 + r < nrows) tc_cp16(d, src + (size_t)(r0 + r) * ld + c * 4);
    else *reinterpret_cast<float4*>(d) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// row index of every element of an accumulator fragment (scratch: >= 16 x TC_LDS floats, clobbered)
__device__ __forceinline__ void tc_rows_of(int (&row_of)[8], float* scratch) {
  for (int i = threadIdx.x; i < 16 * 16; i += TC_THREADS) scratch[(i >> 4) * TC_LDS + (i & 15)] = (float)(i >> 4);
  __syncthreads();
  FragC rid;
  wmma::load_matrix_sync(rid, scratch, TC_LDS, wmma::mem_row_major);
#pragma unroll
  for (int e = 0; e < 8; ++e) row_of[e] = (int)rid.x[e];
  __syncthreads();
}

// C[16 x 32] (two fragments) = A[rows a_row0.., :HD] . B[rows b_row0 .. b_row0+31, :HD]^T   (both [row][d], stride LDH)
template <int HD>
__device__ __forceinline__ void tc_nt(FragC (&c)[2], const float* A, int a_row0, const float* Bm, int b_row0) {
  constexpr int LDH = HD + 8;
  wmma::fill_fragment(c[0], 0.f);
  wmma::fill_fragment(c[1], 0.f);
#pragma unroll
  for (int ks = 0; ks < HD / 8; ++ks) {
    FragA ahi, alo;
    wmma::load_matrix_sync(ahi, A + a_row0 * LDH + ks * 8, LDH);
    tc_split(ahi, alo);
#pragma unroll
    for (int nf = 0; nf < 2; ++nf) {
      FragBT bhi, blo;
      wmma::load_matrix_sync(bhi, Bm + (b_row0 + nf * 16) * LDH + ks * 8, LDH);
      tc_split(bhi, blo);
      tc_mma3(c[nf], ahi, alo, bhi, blo);
    }
  }
}

// C[16 x (NF*16)] += S[s_row0 .. +15, 0..63] . Bm[0..63, b_col0 ..]     (S: score tile stride TC_LDS, Bm stride LDH)
template <int HD, int NF>
__device__ __forceinline__ void tc_nn(FragC (&c)[NF], const float* S, int s_row0, const float* Bm, int b_col0) {
  constexpr int LDH = HD + 8;
#pragma unroll
  for (int ks = 0; ks < TC_T / 8; ++ks) {
    FragA ahi, alo;
    wmma::load_matrix_sync(ahi, S + s_row0 * TC_LDS + ks * 8, TC_LDS);
    tc_split(ahi, alo);
#pragma unroll
    for (int nf = 0; nf < NF; ++nf) {
      FragB bhi, blo;
      wmma::load_matrix_sync(bhi, Bm + ks * 8 * LDH + b_col0 + nf * 16, LDH);
      tc_split(bhi, blo);
      tc_mma3(c[nf], ahi, alo, bhi, blo);
    }
  }
}

// C[16 x (NF*16)] += S[0..63, s_col0 .. +15]^T . Bm[0..63, b_col0 ..]
template <int HD, int NF>
__device__ __forceinline__ void tc_tn(FragC (&c)[NF], const float* S, int s_col0, const float* Bm, int b_col0) {
  constexpr int LDH = HD + 8;
#pragma unroll
  for (int ks = 0; ks < TC_T / 8; ++ks) {
    FragAT ahi, alo;
    wmma::load_matrix_sync(ahi, S + ks * 8 * TC_LDS + s_col0, TC_LDS);
    tc_split(ahi, alo);
#pragma unroll
    for (int nf = 0; nf < NF; ++nf) {
      FragB bhi, blo;
      wmma::load_matrix_sync(bhi, Bm + ks * 8 * LDH + b_col0 + nf * 16, LDH);
      tc_split(bhi, blo);
      tc_mma3(c[nf], ahi, alo, bhi, blo);
    }
  }
}

__device__ __forceinline__ bool tc_masked(int mode, int rate, int i, int j) {
  if (mode == 1) return (j / rate) > i;
  if (mode == 2) return j > (i / rate);
  return false;
}
__device__ __forceinline__ int tc_key_tiles(const AttnArgs& a, int i1) {
  int n = (a.Tk + TC_T - 1) / TC_T;
  if (a.mask_mode == 1) n = min(n, (int)((((long long)i1 + 1) * a.rate - 1) / TC_T) + 1);
  else if (a.mask_mode == 2) n = min(n, (i1 / a.rate) / TC_T + 1);
  return n;
}
__device__ __forceinline__ int tc_first_query_tile(const AttnArgs& a, int j0) {
  if (a.mask_mode == 1) return (j0 / a.rate) / TC_T;
  if (a.mask_mode == 2) return (int)(((long long)j0 * a.rate) / TC_T);
  return 0;
}

// element-wise thread mapping over a 64 x 64 score tile: row = tid / 4, 16 columns starting at (tid % 4) * 16.
// Returns the log2-domain scores of this thread's 16 entries, -inf where masked or out of range.
__device__ __forceinline__ void tc_scores(float (&s)[16], const float* S, const AttnArgs& a, int b, int qi, int j0, int row,
                                          int part) {
  const float4* sp = reinterpret_cast<const float4*>(S + row * TC_LDS + part * 16);
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const float4 v = sp[c];
    s[4 * c] = v.x; s[4 * c + 1] = v.y; s[4 * c + 2] = v.z; s[4 * c + 3] = v.w;
  }
  const bool rowok = qi < a.Tq;
  const unsigned pq = (a.pad_q && rowok) ? a.pad_q[(size_t)b * a.Tq + qi] : 0u;
#pragma unroll
  for (int c = 0; c < 16; ++c) {
    const int kj = j0 + part * 16 + c;
    bool ok = rowok && kj < a.Tk && !tc_masked(a.mask_mode, a.rate, qi, kj);
    if (ok && pq) ok = a.pad_k[(size_t)b * a.Tk + kj] == 0;
    s[c] = ok ? s[c] * a.scale_log2 : -INFINITY;
  }
}
__device__ __forceinline__ void tc_store16(float* S, int row, int part, const float (&v)[16]) {
  float4* sp = reinterpret_cast<float4*>(S + row * TC_LDS + part * 16);
#pragma unroll
  for (int c = 0; c < 4; ++c) sp[c] = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
}

// ---------------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(TC_THREADS) attn_tc_fwd_kernel(AttnArgs a) {
  constexpr int LDH = HD + 8, TILE = TC_T * LDH, NOF = HD / 32;
  extern __shared__ __align__(128) float tc_sm[];
  float* Qs = tc_sm;                  // [64][LDH]
  float* Kb = Qs + TILE;              // 2 x [64][LDH]
  float* Vb = Kb + 2 * TILE;          // 2 x [64][LDH]
  float* Ss = Vb + 2 * TILE;          // [64][TC_LDS]
  float* alpha = Ss + TC_T * TC_LDS;  // [64]
  const int b = blockIdx.y / a.nh, h = blockIdx.y % a.nh;
  const int i0 = blockIdx.x * TC_T;
  const int warp = threadIdx.x >> 5;
  const int rs = warp & 3, ch = warp >> 2;
  const int row = threadIdx.x >> 2, part = threadIdx.x & 3;
  const float* qb = a.q + (size_t)b * a.Tq * a.ldq + h * HD;
  const float* kb = a.k + (size_t)b * a.Tk * a.ldk + h * HD;
  const float* vb = a.v + (size_t)b * a.Tk * a.ldv + h * HD;

  const int njt = tc_key_tiles(a, min(i0 + TC_T, a.Tq) - 1);
  tc_load_rows<HD>(Qs, qb, a.ldq, i0, a.Tq);
  tc_load_rows<HD>(Kb, kb, a.ldk, 0, a.Tk);
  tc_load_rows<HD>(Vb, vb, a.ldv, 0, a.Tk);
  tc_commit();
  int row_of[8];
  tc_rows_of(row_of, Ss);

  FragC o[NOF];
#pragma unroll
  for (int nf = 0; nf < NOF; ++nf) wmma::fill_fragment(o[nf], 0.f);
  float m_run = -INFINITY, l_run = 0.f;
  for (int jt = 0; jt < njt; ++jt) {
    const float* Ks = Kb + (jt & 1) * TILE;
    const float* Vs = Vb + (jt & 1) * TILE;
    if (jt + 1 < njt) {
      tc_load_rows<HD>(Kb + ((jt + 1) & 1) * TILE, kb, a.ldk, (jt + 1) * TC_T, a.Tk);
      tc_load_rows<HD>(Vb + ((jt + 1) & 1) * TILE, vb, a.ldv, (jt + 1) * TC_T, a.Tk);
      tc_commit();
      tc_wait1();
    } else {
      tc_wait0();
    }
    __syncthreads();
    {
      FragC s[2];
      tc_nt<HD>(s, Qs, rs * 16, Ks, ch * 32);
      wmma::store_matrix_sync(Ss + rs * 16 * TC_LDS + ch * 32, s[0], TC_LDS, wmma::mem_row_major);
      wmma::store_matrix_sync(Ss + rs * 16 * TC_LDS + ch * 32 + 16, s[1], TC_LDS, wmma::mem_row_major);
    }
    __syncthreads();
    {
      float s[16];
      tc_scores(s, Ss, a, b, i0 + row, jt * TC_T, row, part);
      float mx = s[0];
#pragma unroll
      for (int c = 1; c < 16; ++c) mx = fmaxf(mx, s[c]);
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      const float mn = fmaxf(m_run, mx);
      const float al = mn == -INFINITY ? 1.f : ex2_ftz(m_run - mn);
      float rsum = 0.f;
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        s[c] = mn == -INFINITY ? 0.f : ex2_ftz(s[c] - mn);
        rsum += s[c];
      }
      rsum += __shfl_xor_sync(0xffffffffu, rsum, 1);
      rsum += __shfl_xor_sync(0xffffffffu, rsum, 2);
      l_run = l_run * al + rsum;
      m_run = mn;
      tc_store16(Ss, row, part, s);
      if (part == 0) alpha[row] = al;
    }
    __syncthreads();
#pragma unroll
    for (int nf = 0; nf < NOF; ++nf)
#pragma unroll
      for (int e = 0; e < 8; ++e) o[nf].x[e] *= alpha[rs * 16 + row_of[e]];
    tc_nn<HD, NOF>(o, Ss, rs * 16, Vs, ch * (HD / 2));
    __syncthreads();
  }
  // normalise, stage through shared memory, write rows < Tq
  if (part == 0) alpha[row] = l_run > 0.f ? 1.f / l_run : 0.f;  // a query with no visible key gives 0 (torch: NaN)
  __syncthreads();
#pragma unroll
  for (int nf = 0; nf < NOF; ++nf) {
#pragma unroll
    for (int e = 0; e < 8; ++e) o[nf].x[e] *= alpha[rs * 16 + row_of[e]];
    wmma::store_matrix_sync(Ss + rs * 16 * TC_LDS + ch * (HD / 2) + nf * 16, o[nf], TC_LDS, wmma::mem_row_major);
  }
  __syncthreads();
  const int qi = i0 + row;
  if (qi < a.Tq) {
    float* orow = a.o + (size_t)b * a.Tq * a.ldo + (size_t)qi * a.ldo + h * HD + part * (HD / 4);
    const float* srow = Ss + row * TC_LDS + part * (HD / 4);
#pragma unroll
    for (int c = 0; c < HD / 16; ++c) reinterpret_cast<float4*>(orow)[c] = reinterpret_cast<const float4*>(srow)[c];
    if (part == 0 && a.lse) a.lse[(size_t)blockIdx.y * a.Tq + qi] = l_run > 0.f ? m_run + log2f(l_run) : 0.f;
  }
}

// ---------------------------------------------------------------------------------------------------------
// backward 1: dQ (and D = dO . O), query tile resident
// ---------------------------------------------------------------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(TC_THREADS) attn_tc_dq_kernel(AttnArgs a) {
  constexpr int LDH = HD + 8, TILE = TC_T * LDH, NOF = HD / 32;
  extern __shared__ __align__(128) float tc_sm[];
  float* Qs = tc_sm;
  float* dOs = Qs + TILE;
  float* Kb = dOs + TILE;            // 2 x
  float* Vb = Kb + 2 * TILE;         // 2 x
  float* Ss = Vb + 2 * TILE;         // [64][TC_LDS]  scores -> dS
  float* Ps = Ss + TC_T * TC_LDS;    // [64][TC_LDS]  dP
  const int b = blockIdx.y / a.nh, h = blockIdx.y % a.nh;
  const int i0 = blockIdx.x * TC_T;
  const int warp = threadIdx.x >> 5;
  const int rs = warp & 3, ch = warp >> 2;
  const int row = threadIdx.x >> 2, part = threadIdx.x & 3;
  const float* qb = a.q + (size_t)b * a.Tq * a.ldq + h * HD;
  const float* kb = a.k + (size_t)b * a.Tk * a.ldk + h * HD;
  const float* vb = a.v + (size_t)b * a.Tk * a.ldv + h * HD;
  const float* dob = a.dout + (size_t)b * a.Tq * a.lddo + h * HD;

  const int njt = tc_key_tiles(a, min(i0 + TC_T, a.Tq) - 1);
  tc_load_rows<HD>(Qs, qb, a.ldq, i0, a.Tq);
  tc_load_rows<HD>(dOs, dob, a.lddo, i0, a.Tq);
  tc_load_rows<HD>(Kb, kb, a.ldk, 0, a.Tk);
  tc_load_rows<HD>(Vb, vb, a.ldv, 0, a.Tk);
  tc_commit();
  // D = dO . O and the saved log-sum-exp of this thread's row
  const int qi = i0 + row;
  float dvec = 0.f, lse = 0.f;
  if (qi < a.Tq) {
    const float4* dp4 = reinterpret_cast<const float4*>(dob + (size_t)qi * a.lddo + part * (HD / 4));
    const float4* op4 = reinterpret_cast<const float4*>(a.o + (size_t)b * a.Tq * a.ldo + (size_t)qi * a.ldo + h * HD + part * (HD / 4));
#pragma unroll
    for (int c = 0; c < HD / 16; ++c) {
      const float4 x = __ldg(dp4 + c), y = __ldg(op4 + c);
      dvec = fmaf(x.x, y.x, fmaf(x.y, y.y, fmaf(x.z, y.z, fmaf(x.w, y.w, dvec))));
    }
    lse = a.lse[(size_t)blockIdx.y * a.Tq + qi];
  }
  dvec += __shfl_xor_sync(0xffffffffu, dvec, 1);
  dvec += __shfl_xor_sync(0xffffffffu, dvec, 2);
  if (part == 0 && qi < a.Tq) a.dvec[(size_t)blockIdx.y * a.Tq + qi] = dvec;

  FragC dq[NOF];
#pragma unroll
  for (int nf = 0; nf < NOF; ++nf) wmma::fill_fragment(dq[nf], 0.f);
  for (int jt = 0; jt < njt; ++jt) {
    const float* Ks = Kb + (jt & 1) * TILE;
    const float* Vs = Vb + (jt & 1) * TILE;
    if (jt + 1 < njt) {
      tc_load_rows<HD>(Kb + ((jt + 1) & 1) * TILE, kb, a.ldk, (jt + 1) * TC_T, a.Tk);
      tc_load_rows<HD>(Vb + ((jt + 1) & 1) * TILE, vb, a.ldv, (jt + 1) * TC_T, a.Tk);
      tc_commit();
      tc_wait1();
    } else {
      tc_wait0();
    }
    __syncthreads();
    {
      FragC s[2];
      tc_nt<HD>(s, Qs, rs * 16, Ks, ch * 32);
      wmma::store_matrix_sync(Ss + rs * 16 * TC_LDS + ch * 32, s[0], TC_LDS, wmma::mem_row_major);
      wmma::store_matrix_sync(Ss + rs * 16 * TC_LDS + ch * 32 + 16, s[1], TC_LDS, wmma::mem_row_major);
      tc_nt<HD>(s, dOs, rs * 16, Vs, ch * 32);
      wmma::store_matrix_sync(Ps + rs * 16 * TC_LDS + ch * 32, s[0], TC_LDS, wmma::mem_row_major);
      wmma::store_matrix_sync(Ps + rs * 16 * TC_LDS + ch * 32 + 16, s[1], TC_LDS, wmma::mem_row_major);
    }
    __syncthreads();
    {
      float s[16];
      tc_scores(s, Ss, a, b, qi, jt * TC_T, row, part);
      const float4* dp4 = reinterpret_cast<const float4*>(Ps + row * TC_LDS + part * 16);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float4 d = dp4[c];
        const float dd[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) s[4 * c + e] = ex2_ftz(s[4 * c + e] - lse) * (dd[e] - dvec) * a.scale;  // masked: 0
      }
      tc_store16(Ss, row, part, s);
    }
    __syncthreads();
    tc_nn<HD, NOF>(dq, Ss, rs * 16, Ks, ch * (HD / 2));
    __syncthreads();
  }
#pragma unroll
  for (int nf = 0; nf < NOF; ++nf)
    wmma::store_matrix_sync(Ss + rs * 16 * TC_LDS + ch * (HD / 2) + nf * 16, dq[nf], TC_LDS, wmma::mem_row_major);
  __syncthreads();
  if (qi < a.Tq) {
    float* drow = a.dq + (size_t)b * a.Tq * a.lddq + (size_t)qi * a.lddq + h * HD + part * (HD / 4);
    const float* srow = Ss + row * TC_LDS + part * (HD / 4);
#pragma unroll
    for (int c = 0; c < HD / 16; ++c) reinterpret_cast<float4*>(drow)[c] = reinterpret_cast<const float4*>(srow)[c];
  }
}

// ---------------------------------------------------------------------------------------------------------
// backward 2: dK, dV, key tile resident
// ---------------------------------------------------------------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(TC_THREADS) attn_tc_dkv_kernel(AttnArgs a) {
  constexpr int LDH = HD + 8, TILE = TC_T * LDH, NOF = HD / 32;
  extern __shared__ __align__(128) float tc_sm[];
  float* Ks = tc_sm;
  float* Vs = Ks + TILE;
  float* Qb = Vs + TILE;             // 2 x
  float* dOb = Qb + 2 * TILE;        // 2 x
  float* Ss = dOb + 2 * TILE;        // [64 queries][TC_LDS keys]  scores -> P
  float* Ps = Ss + TC_T * TC_LDS;    //                            dP -> dS
  const int b = blockIdx.y / a.nh, h = blockIdx.y % a.nh;
  const int j0 = blockIdx.x * TC_T;
  const int warp = threadIdx.x >> 5;
  const int rs = warp & 3, ch = warp >> 2;
  const int row = threadIdx.x >> 2, part = threadIdx.x & 3;
  const float* qb = a.q + (size_t)b * a.Tq * a.ldq + h * HD;
  const float* kb = a.k + (size_t)b * a.Tk * a.ldk + h * HD;
  const float* vb = a.v + (size_t)b * a.Tk * a.ldv + h * HD;
  const float* dob = a.dout + (size_t)b * a.Tq * a.lddo + h * HD;

  const int it0 = tc_first_query_tile(a, j0), nit = (a.Tq + TC_T - 1) / TC_T;
  tc_load_rows<HD>(Ks, kb, a.ldk, j0, a.Tk);
  tc_load_rows<HD>(Vs, vb, a.ldv, j0, a.Tk);
  if (it0 < nit) {
    tc_load_rows<HD>(Qb + (it0 & 1) * TILE, qb, a.ldq, it0 * TC_T, a.Tq);
    tc_load_rows<HD>(dOb + (it0 & 1) * TILE, dob, a.lddo, it0 * TC_T, a.Tq);
  }
  tc_commit();
  FragC dk[NOF], dv[NOF];
#pragma unroll
  for (int nf = 0; nf < NOF; ++nf) {
    wmma::fill_fragment(dk[nf], 0.f);
    wmma::fill_fragment(dv[nf], 0.f);
  }
  if (it0 >= nit) tc_wait0();
  for (int it = it0; it < nit; ++it) {
    const float* Qs = Qb + (it & 1) * TILE;
    const float* dOs = dOb + (it & 1) * TILE;
    if (it + 1 < nit) {
      tc_load_rows<HD>(Qb + ((it + 1) & 1) * TILE, qb, a.ldq, (it + 1) * TC_T, a.Tq);
      tc_load_rows<HD>(dOb + ((it + 1) & 1) * TILE, dob, a.lddo, (it + 1) * TC_T, a.Tq);
      tc_commit();
      tc_wait1();
    } else {
      tc_wait0();
    }
    __syncthreads();
    {
      FragC s[2];
      tc_nt<HD>(s, Qs, rs * 16, Ks, ch * 32);
      wmma::store_matrix_sync(Ss + rs * 16 * TC_LDS + ch * 32, s[0], TC_LDS, wmma::mem_row_major);
      wmma::store_matrix_sync(Ss + rs * 16 * TC_LDS + ch * 32 + 16, s[1], TC_LDS, wmma::mem_row_major);
      tc_nt<HD>(s, dOs, rs * 16, Vs, ch * 32);
      wmma::store_matrix_sync(Ps + rs * 16 * TC_LDS + ch * 32, s[0], TC_LDS, wmma::mem_row_major);
      wmma::store_matrix_sync(Ps + rs * 16 * TC_LDS + ch * 32 + 16, s[1], TC_LDS, wmma::mem_row_major);
    }
    __syncthreads();
    {
      const int qi = it * TC_T + row;
      float lse = 0.f, dvec = 0.f;
      if (qi < a.Tq) {
        lse = a.lse[(size_t)blockIdx.y * a.Tq + qi];
        dvec = a.dvec[(size_t)blockIdx.y * a.Tq + qi];
      }
      float s[16], ds[16];
      tc_scores(s, Ss, a, b, qi, j0, row, part);
      const float4* dp4 = reinterpret_cast<const float4*>(Ps + row * TC_LDS + part * 16);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float4 d = dp4[c];
        const float dd[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          s[4 * c + e] = ex2_ftz(s[4 * c + e] - lse);  // rows past Tq and masked entries: ex2(-inf) = 0
          ds[4 * c + e] = s[4 * c + e] * (dd[e] - dvec) * a.scale;
        }
      }
      tc_store16(Ss, row, part, s);
      tc_store16(Ps, row, part, ds);
    }
    __syncthreads();
    tc_tn<HD, NOF>(dv, Ss, rs * 16, dOs, ch * (HD / 2));  // dV[j] += sum_i P[i][j] dO[i]
    tc_tn<HD, NOF>(dk, Ps, rs * 16, Qs, ch * (HD / 2));   // dK[j] += sum_i dS[i][j] Q[i]
    __syncthreads();
  }
#pragma unroll
  for (int nf = 0; nf < NOF; ++nf) {
    wmma::store_matrix_sync(Ss + rs * 16 * TC_LDS + ch * (HD / 2) + nf * 16, dk[nf], TC_LDS, wmma::mem_row_major);
    wmma::store_matrix_sync(Ps + rs * 16 * TC_LDS + ch * (HD / 2) + nf * 16, dv[nf], TC_LDS, wmma::mem_row_major);
  }
  __syncthreads();
  const int kj = j0 + row;
  if (kj < a.Tk) {
    float* rk = a.dk + (size_t)b * a.Tk * a.lddk + (size_t)kj * a.lddk + h * HD + part * (HD / 4);
    float* rv = a.dv + (size_t)b * a.Tk * a.lddv + (size_t)kj * a.lddv + h * HD + part * (HD / 4);
    const float* sk = Ss + row * TC_LDS + part * (HD / 4);
    const float* sv = Ps + row * TC_LDS + part * (HD / 4);
#pragma unroll
    for (int c = 0; c < HD / 16; ++c) {
      reinterpret_cast<float4*>(rk)[c] = reinterpret_cast<const float4*>(sk)[c];
      reinterpret_cast<float4*>(rv)[c] = reinterpret_cast<const float4*>(sv)[c];
    }
  }
}

template <int HD>
constexpr size_t tc_fwd_smem() { return sizeof(float) * (5 * TC_T * (HD + 8) + TC_T * TC_LDS + TC_T) + 128; }
template <int HD>
constexpr size_t tc_bwd_smem() { return sizeof(float) * (6 * TC_T * (HD + 8) + 2 * TC_T * TC_LDS) + 128; }

template <int HD>
static int attn_tc_launch_hd(const AttnArgs& a, int backward, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    MRG_CUDA_CHECK(cudaFuncSetAttribute(attn_tc_fwd_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_fwd_smem<HD>()));
    MRG_CUDA_CHECK(cudaFuncSetAttribute(attn_tc_dq_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_bwd_smem<HD>()));
    MRG_CUDA_CHECK(cudaFuncSetAttribute(attn_tc_dkv_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_bwd_smem<HD>()));
    attr_set = true;
  }
  const dim3 gq((a.Tq + TC_T - 1) / TC_T, a.B * a.nh), gk((a.Tk + TC_T - 1) / TC_T, a.B * a.nh);
  if (!backward) {
    count_launch();
    attn_tc_fwd_kernel<HD><<<gq, TC_THREADS, tc_fwd_smem<HD>(), stream>>>(a);
  } else {
    count_launch(2);
    attn_tc_dq_kernel<HD><<<gq, TC_THREADS, tc_bwd_smem<HD>(), stream>>>(a);
    MRG_CUDA_CHECK(cudaGetLastError());
    attn_tc_dkv_kernel<HD><<<gk, TC_THREADS, tc_bwd_smem<HD>(), stream>>>(a);
  }
  MRG_CUDA_CHECK(cudaGetLastError());
  return 0;
}

// entry used by the C-ABI functions in mrg_attention.cu (arguments already validated there)
int attn_tc_launch(const AttnArgs& a, int hd, int backward, cudaStream_t stream) {
  return hd == 32 ? attn_tc_launch_hd<32>(a, backward, stream) : attn_tc_launch_hd<64>(a, backward, stream);
}

}  // namespace mrg
