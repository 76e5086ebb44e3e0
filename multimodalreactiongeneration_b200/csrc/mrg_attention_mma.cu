// Fused multi-head attention on the warp-level tensor cores (mma.sync m16n8k8 tf32, 3xTF32): forward, dQ, dK/dV.
//
// Same contract, argument block and three-kernel structure as the CUDA-core kernels of mrg_attention.cu (flash-style,
// scores never reach HBM, heads addressed in place, the mask is a function, deterministic, no atomics); what changes is
// where the five 64 x 64 x d products run.  The CUDA-core kernels reach 22-32 TFLOP/s and are bound by shared-memory
// wavefronts (profiles/r1d_attention.txt).  Measured on B200 (tools/mma_sync_rate.cu, profiles/r2_mma_rate.txt): the
// warp-level tf32 MMA path issues 510 FMA/clk/SM — 4x the FFMA rate — so the 3-term split A.B ~ A_lo.B_hi + A_hi.B_lo +
// A_hi.B_hi (fp32-grade: the parity budget of the path is 1e-5) still leaves a margin, and the reduced-precision modes
// run one pass.  A tcgen05 version was costed and not built: with the scores in tensor memory the P.V / dS.K / P^T.dO /
// dS^T.Q products have N = head_dim = 32 or 64, and a tcgen05.mma costs ~116 cycles for any N <= 128 (tools/mma_rate.cu),
// so those four products would run the big tensor pipe at 1/8 - 1/4 of its rate.
//
// Register-resident score tile (FlashAttention-2 style): a warp owns 16 rows of the resident 64-row tile; the 16 x 64
// score block lives in the accumulator fragments of 8 n-tiles, and is fed back as the A operand of the second product
// WITHOUT leaving registers: the accumulator layout of m16n8k8 (row g = lane / 4: columns 2q, 2q+1 with q = lane % 4)
// is read as an A fragment whose k index is a permutation of the column index (k = q <-> column 2q, k = q + 4 <-> column
// 2q + 1); the B fragments of that product load their two rows in the same permuted order, so the sum is unchanged.
// Operand tiles sit in shared memory in their natural [row][d] layout with a row stride of d + 4 floats, which makes
// every fragment load (scalar 4-byte loads: 32-bit operands cannot use ldmatrix.trans) bank-conflict free; tiles of the
// streamed side arrive with cp.async, double-buffered.
#include <cstddef>
#include <cstdlib>

#include "mrg_mma_common.cuh"

namespace mrg {

constexpr int AM_T = 64;          // rows of the resident tile / of a streamed tile
// A warp owns MT m-tiles of 16 rows (MT = 1: 4 warps per CTA, MT = 2: 2 warps): with MT = 2 every B fragment a warp loads
// and splits feeds two MMAs per pass, which halves the load / split instructions per MMA — but at 230-255 registers per
// thread only 8 warps stay resident per SM, and that costs more than the instructions save.  Measured (B200,
// profiles/r2_attention.txt): MT = 1: 150 / 476 us forward / backward at the cfg 2 shape, 480 / 1757 us at the cfg 4 shape;
// MT = 2: 172 / 489 and 888 / 1855.  MT = 1 is built; the template parameter stays for the next experiment.
constexpr int am_threads(int MT) { return 32 * (4 / MT); }
#ifndef AM_MT_FWD
#define AM_MT_FWD 1
#endif
#ifndef AM_MT_BWD32
#define AM_MT_BWD32 1
#endif

// c[mt][nt] += X[16 MT rows of this warp] . Y^T : c[mt][nt][.] = sum_d X[mt*16 + row][d] Y[nt*8 + col][d]
// (X, Y: [64][HD + 4] tiles)
template <int HD, int PASSES, int MT>
__device__ __forceinline__ void am_prod_nt(float (&c)[MT][8][4], const float* X, const float* Y, int g, int q) {
  constexpr int LD = HD + 4;
#pragma unroll
  for (int ks = 0; ks < HD / 8; ++ks) {
    uint32_t ahi[MT][4], alo[MT][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      const float* xr = X + (mt * 16 + g) * LD + ks * 8 + q;
      const float xa[4] = {xr[0], xr[8 * LD], xr[4], xr[8 * LD + 4]};
      am_split_a<PASSES>(xa, ahi[mt], alo[mt]);
    }
    // the three passes of one accumulator depend on each other: issue pass by pass ACROSS the n-tiles so that
    // consecutive MMAs are independent
    float y0[8], y1[8];
    uint32_t h0[8], h1[8];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float* yr = Y + (nt * 8 + g) * LD + ks * 8 + q;
      y0[nt] = yr[0];
      y1[nt] = yr[4];
      h0[nt] = am_hi<PASSES>(y0[nt]);
      h1[nt] = am_hi<PASSES>(y1[nt]);
    }
    if (PASSES == 3) {
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) am_mma(c[mt][nt], alo[mt], h0[nt], h1[nt]);
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const uint32_t l0 = am_lo(y0[nt], h0[nt]), l1 = am_lo(y1[nt], h1[nt]);
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) am_mma(c[mt][nt], ahi[mt], l0, l1);
      }
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) am_mma(c[mt][nt], ahi[mt], h0[nt], h1[nt]);
  }
}

// o[mt][dt] += P . Y : o[mt][dt][.] = sum_j P[mt][row][j] Y[j][dt*8 + col], P = the 16 MT x 64 block held as accumulator
// fragments
template <int HD, int PASSES, int MT>
__device__ __forceinline__ void am_prod_acc(float (&o)[MT][HD / 8][4], const float (&p)[MT][8][4], const float* Y, int g,
                                            int q) {
  constexpr int LD = HD + 4;
#pragma unroll
  for (int kk = 0; kk < 8; ++kk) {
    // accumulator (row g: cols 2q, 2q+1 | row g+8: cols 2q, 2q+1) read as A with k = q <-> col 2q, k = q+4 <-> col 2q+1
    uint32_t ahi[MT][4], alo[MT][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      const float pa[4] = {p[mt][kk][0], p[mt][kk][2], p[mt][kk][1], p[mt][kk][3]};
      am_split_a<PASSES>(pa, ahi[mt], alo[mt]);
    }
    const float* yp = Y + (kk * 8 + 2 * q) * LD + g;
    float y0[HD / 8], y1[HD / 8];
    uint32_t h0[HD / 8], h1[HD / 8];
#pragma unroll
    for (int dt = 0; dt < HD / 8; ++dt) {
      y0[dt] = yp[dt * 8];
      y1[dt] = yp[LD + dt * 8];
      h0[dt] = am_hi<PASSES>(y0[dt]);
      h1[dt] = am_hi<PASSES>(y1[dt]);
    }
    if (PASSES == 3) {
#pragma unroll
      for (int dt = 0; dt < HD / 8; ++dt)
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) am_mma(o[mt][dt], alo[mt], h0[dt], h1[dt]);
#pragma unroll
      for (int dt = 0; dt < HD / 8; ++dt) {
        const uint32_t l0 = am_lo(y0[dt], h0[dt]), l1 = am_lo(y1[dt], h1[dt]);
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) am_mma(o[mt][dt], ahi[mt], l0, l1);
      }
    }
#pragma unroll
    for (int dt = 0; dt < HD / 8; ++dt)
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) am_mma(o[mt][dt], ahi[mt], h0[dt], h1[dt]);
  }
}

// dst[r][0..HD) (row stride HD + 4) <- src[(r0 + r) * ld + ..]; rows past nrows are zero
template <int HD>
__device__ __forceinline__ void am_load_tile(float* dst, const float* __restrict__ src, int ld, int r0, int nrows) {
  constexpr int C4 = HD / 4, LD = HD + 4;
  for (int f = threadIdx.x; f < AM_T * C4; f += blockDim.x) {
    const int r = f / C4, c = f % C4;
    float* d = dst + r * LD + c * 4;
    if (r0 + r < nrows) am_cp16(d, src + (size_t)(r0 + r) * ld + c * 4);
    else *reinterpret_cast<float4*>(d) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

__device__ __forceinline__ int am_key_tiles(const AttnArgs& a, int i1) {   // key tiles a query tile ending at i1 can see
  int n = (a.Tk + AM_T - 1) / AM_T;
  if (a.mask_mode == 1) n = min(n, (int)((((long long)i1 + 1) * a.rate - 1) / AM_T) + 1);
  else if (a.mask_mode == 2) n = min(n, (i1 / a.rate) / AM_T + 1);
  return n;
}
__device__ __forceinline__ int am_first_query_tile(const AttnArgs& a, int j0) {   // first query tile that sees key j0
  if (a.mask_mode == 1) return (j0 / a.rate) / AM_T;
  if (a.mask_mode == 2) return (int)(((long long)j0 * a.rate) / AM_T);
  return 0;
}

// The mask rule per ROW of a score block, so that no division is left in the per-element code (the first version spent
// ~3000 of its ~3500 instructions per tile on `j / rate`):
//   rows = queries i: key j is visible iff j <= am_row_limit<true>(i)   (mode 1: j / rate <= i  <=>  j <= (i+1) rate - 1;
//                                                                        mode 2: j <= i / rate)
//   rows = keys j:    query i is visible iff i >= am_row_limit<false>(j) (mode 1: i >= j / rate; mode 2: j <= i / rate  <=>
//                                                                        i >= j rate)
// rows past the end get a limit that hides every column.
template <bool ROWS_ARE_QUERIES>
__device__ __forceinline__ int am_row_limit(const AttnArgs& a, int r) {
  if (ROWS_ARE_QUERIES) {
    if (r >= a.Tq) return -1;
    if (a.mask_mode == 1) return (int)min((long long)(r + 1) * a.rate - 1, (long long)0x7fffffff);
    if (a.mask_mode == 2) return r / a.rate;
    return 0x7fffffff;
  }
  if (r >= a.Tk) return 0x7fffffff;
  if (a.mask_mode == 1) return r / a.rate;
  if (a.mask_mode == 2) return (int)min((long long)r * a.rate, (long long)0x7fffffff);
  return 0;
}

// scores of a 16 x 64 block in the log2 domain: rows r0 / r1 (two per thread) with their limits, columns c0 + nt*8 + 2q
// (+1); -inf where masked / out of range.  ROWS_ARE_QUERIES = false: the block is transposed (rows = keys).
template <bool ROWS_ARE_QUERIES>
__device__ __forceinline__ void am_finish_scores(float (&s)[8][4], const AttnArgs& a, int b, int r0, int r1, int lim0,
                                                 int lim1, int c0, int q) {
  const int nrow = ROWS_ARE_QUERIES ? a.Tq : a.Tk, ncol = ROWS_ARE_QUERIES ? a.Tk : a.Tq;
  const unsigned char* prow = ROWS_ARE_QUERIES ? a.pad_q : a.pad_k;
  const unsigned char* pcol = ROWS_ARE_QUERIES ? a.pad_k : a.pad_q;
  unsigned pr0 = 0, pr1 = 0;
  if (prow) {
    pr0 = r0 < nrow ? prow[(size_t)b * nrow + r0] : 0u;
    pr1 = r1 < nrow ? prow[(size_t)b * nrow + r1] : 0u;
  }
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int c = c0 + nt * 8 + 2 * q + e;
      unsigned pc = 0;
      if (pcol) pc = c < ncol ? pcol[(size_t)b * ncol + c] : 0u;
      const bool in = c < ncol;
      const bool ok0 = in && (ROWS_ARE_QUERIES ? c <= lim0 : c >= lim0) && !(pr0 & pc);
      const bool ok1 = in && (ROWS_ARE_QUERIES ? c <= lim1 : c >= lim1) && !(pr1 & pc);
      s[nt][e] = ok0 ? s[nt][e] * a.scale_log2 : -INFINITY;
      s[nt][2 + e] = ok1 ? s[nt][2 + e] * a.scale_log2 : -INFINITY;
    }
}

// ---------------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------------
template <int HD, int PASSES, int MT>
__global__ void __launch_bounds__(am_threads(MT), HD == 32 ? 4 : 2) attn_mma_fwd_kernel(AttnArgs a) {
  constexpr int LD = HD + 4, TILE = AM_T * LD, ND = HD / 8;
  extern __shared__ __align__(16) float am_sm[];
  float* Qs = am_sm;              // [64][LD]
  float* KV = Qs + TILE;          // stage s: K at KV + s*2*TILE, V at + TILE
  const int b = blockIdx.y / a.nh, h = blockIdx.y % a.nh;
  const int i0 = blockIdx.x * AM_T;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
  const float* qb = a.q + (size_t)b * a.Tq * a.ldq + h * HD;
  const float* kb = a.k + (size_t)b * a.Tk * a.ldk + h * HD;
  const float* vb = a.v + (size_t)b * a.Tk * a.ldv + h * HD;
  const int njt = am_key_tiles(a, min(i0 + AM_T, a.Tq) - 1);

  am_load_tile<HD>(Qs, qb, a.ldq, i0, a.Tq);
  am_load_tile<HD>(KV, kb, a.ldk, 0, a.Tk);
  am_load_tile<HD>(KV + TILE, vb, a.ldv, 0, a.Tk);
  am_commit();

  const int wrow = warp * 16 * MT;   // first row of this warp inside the tile
  int row[MT][2], lim[MT][2];
  float m[MT][2], l[MT][2], o[MT][ND][4];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      row[mt][rr] = i0 + wrow + mt * 16 + g + 8 * rr;
      lim[mt][rr] = am_row_limit<true>(a, row[mt][rr]);
      m[mt][rr] = -INFINITY;
      l[mt][rr] = 0.f;
    }
#pragma unroll
    for (int dt = 0; dt < ND; ++dt) o[mt][dt][0] = o[mt][dt][1] = o[mt][dt][2] = o[mt][dt][3] = 0.f;
  }

  for (int jt = 0; jt < njt; ++jt) {
    float* Ks = KV + (jt & 1) * 2 * TILE;
    float* Vs = Ks + TILE;
    if (jt + 1 < njt) {   // the other stage was last read in iteration jt-1, which ended with a block barrier
      float* Kn = KV + ((jt + 1) & 1) * 2 * TILE;
      am_load_tile<HD>(Kn, kb, a.ldk, (jt + 1) * AM_T, a.Tk);
      am_load_tile<HD>(Kn + TILE, vb, a.ldv, (jt + 1) * AM_T, a.Tk);
      am_commit();
      am_wait1();
    } else {
      am_wait0();
    }
    __syncthreads();
    float s[MT][8][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) s[mt][nt][0] = s[mt][nt][1] = s[mt][nt][2] = s[mt][nt][3] = 0.f;
    am_prod_nt<HD, PASSES, MT>(s, Qs + wrow * LD, Ks, g, q);
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      am_finish_scores<true>(s[mt], a, b, row[mt][0], row[mt][1], lim[mt][0], lim[mt][1], jt * AM_T, q);
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        float mx = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) mx = fmaxf(mx, fmaxf(s[mt][nt][rr * 2], s[mt][nt][rr * 2 + 1]));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
        const float mn = fmaxf(m[mt][rr], mx);
        const float alpha = mn == -INFINITY ? 1.f : ex2_ftz(m[mt][rr] - mn);
        float rs = 0.f;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const float p = mn == -INFINITY ? 0.f : ex2_ftz(s[mt][nt][rr * 2 + e] - mn);
            s[mt][nt][rr * 2 + e] = p;
            rs += p;
          }
        l[mt][rr] = l[mt][rr] * alpha + rs;   // this thread's share of the row sum (the quad is reduced at the end)
        m[mt][rr] = mn;
#pragma unroll
        for (int dt = 0; dt < ND; ++dt) {
          o[mt][dt][rr * 2] *= alpha;
          o[mt][dt][rr * 2 + 1] *= alpha;
        }
      }
    }
    am_prod_acc<HD, PASSES, MT>(o, s, Vs, g, q);
    __syncthreads();
  }
  float* ob = a.o + (size_t)b * a.Tq * a.ldo + h * HD;
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      float lt = l[mt][rr];
      lt += __shfl_xor_sync(0xffffffffu, lt, 1);
      lt += __shfl_xor_sync(0xffffffffu, lt, 2);
      const int qi = row[mt][rr];
      if (qi >= a.Tq) continue;
      const float inv = lt > 0.f ? 1.f / lt : 0.f;   // a query with no visible key gives 0 (torch: NaN)
      float* orow = ob + (size_t)qi * a.ldo + 2 * q;
#pragma unroll
      for (int dt = 0; dt < ND; ++dt)
        *reinterpret_cast<float2*>(orow + dt * 8) = make_float2(o[mt][dt][rr * 2] * inv, o[mt][dt][rr * 2 + 1] * inv);
      if (q == 0 && a.lse) a.lse[(size_t)blockIdx.y * a.Tq + qi] = lt > 0.f ? m[mt][rr] + log2f(lt) : 0.f;
    }
}

// ---------------------------------------------------------------------------------------------------------
// backward 1: dQ (and D = dO . O), query tile resident
// ---------------------------------------------------------------------------------------------------------
template <int HD, int PASSES, int MT>
__global__ void __launch_bounds__(am_threads(MT), HD == 32 ? 3 : 2) attn_mma_dq_kernel(AttnArgs a) {
  constexpr int LD = HD + 4, TILE = AM_T * LD, ND = HD / 8;
  extern __shared__ __align__(16) float am_sm[];
  float* Qs = am_sm;              // [64][LD]
  float* dOs = Qs + TILE;
  float* KV = dOs + TILE;         // stage s: K at KV + s*2*TILE, V at + TILE
  const int b = blockIdx.y / a.nh, h = blockIdx.y % a.nh;
  const int i0 = blockIdx.x * AM_T;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
  const float* qb = a.q + (size_t)b * a.Tq * a.ldq + h * HD;
  const float* kb = a.k + (size_t)b * a.Tk * a.ldk + h * HD;
  const float* vb = a.v + (size_t)b * a.Tk * a.ldv + h * HD;
  const float* ob = a.o + (size_t)b * a.Tq * a.ldo + h * HD;
  const float* dob = a.dout + (size_t)b * a.Tq * a.lddo + h * HD;
  const int njt = am_key_tiles(a, min(i0 + AM_T, a.Tq) - 1);

  am_load_tile<HD>(Qs, qb, a.ldq, i0, a.Tq);
  am_load_tile<HD>(dOs, dob, a.lddo, i0, a.Tq);
  am_load_tile<HD>(KV, kb, a.ldk, 0, a.Tk);
  am_load_tile<HD>(KV + TILE, vb, a.ldv, 0, a.Tk);
  am_commit();

  const int wrow = warp * 16 * MT;
  int row[MT][2], lim[MT][2];
  float lse[MT][2], dvec[MT][2], acc[MT][ND][4];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {   // D_i = dO_i . O_i: this thread's columns, then the quad
      const int qi = i0 + wrow + mt * 16 + g + 8 * rr;
      row[mt][rr] = qi;
      lim[mt][rr] = am_row_limit<true>(a, qi);
      float d = 0.f;
      lse[mt][rr] = 0.f;
      if (qi < a.Tq) {
#pragma unroll
        for (int dt = 0; dt < ND; ++dt) {
          const float2 x = __ldg(reinterpret_cast<const float2*>(dob + (size_t)qi * a.lddo + dt * 8 + 2 * q));
          const float2 y = __ldg(reinterpret_cast<const float2*>(ob + (size_t)qi * a.ldo + dt * 8 + 2 * q));
          d = fmaf(x.x, y.x, fmaf(x.y, y.y, d));
        }
        lse[mt][rr] = a.lse[(size_t)blockIdx.y * a.Tq + qi];
      }
      d += __shfl_xor_sync(0xffffffffu, d, 1);
      d += __shfl_xor_sync(0xffffffffu, d, 2);
      dvec[mt][rr] = d;
      if (q == 0 && qi < a.Tq) a.dvec[(size_t)blockIdx.y * a.Tq + qi] = d;
    }
#pragma unroll
    for (int dt = 0; dt < ND; ++dt) acc[mt][dt][0] = acc[mt][dt][1] = acc[mt][dt][2] = acc[mt][dt][3] = 0.f;
  }

  for (int jt = 0; jt < njt; ++jt) {
    float* Ks = KV + (jt & 1) * 2 * TILE;
    float* Vs = Ks + TILE;
    if (jt + 1 < njt) {
      float* Kn = KV + ((jt + 1) & 1) * 2 * TILE;
      am_load_tile<HD>(Kn, kb, a.ldk, (jt + 1) * AM_T, a.Tk);
      am_load_tile<HD>(Kn + TILE, vb, a.ldv, (jt + 1) * AM_T, a.Tk);
      am_commit();
      am_wait1();
    } else {
      am_wait0();
    }
    __syncthreads();
    float s[MT][8][4], dp[MT][8][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        s[mt][nt][0] = s[mt][nt][1] = s[mt][nt][2] = s[mt][nt][3] = 0.f;
        dp[mt][nt][0] = dp[mt][nt][1] = dp[mt][nt][2] = dp[mt][nt][3] = 0.f;
      }
    am_prod_nt<HD, PASSES, MT>(s, Qs + wrow * LD, Ks, g, q);
    am_prod_nt<HD, PASSES, MT>(dp, dOs + wrow * LD, Vs, g, q);
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      am_finish_scores<true>(s[mt], a, b, row[mt][0], row[mt][1], lim[mt][0], lim[mt][1], jt * AM_T, q);
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int rr = c >> 1;
          const float p = ex2_ftz(s[mt][nt][c] - lse[mt][rr]);   // masked: ex2(-inf) = 0
          s[mt][nt][c] = p * (dp[mt][nt][c] - dvec[mt][rr]) * a.scale;
        }
    }
    am_prod_acc<HD, PASSES, MT>(acc, s, Ks, g, q);
    __syncthreads();
  }
  float* dqb = a.dq + (size_t)b * a.Tq * a.lddq + h * HD;
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const int qi = row[mt][rr];
      if (qi >= a.Tq) continue;
      float* rowp = dqb + (size_t)qi * a.lddq + 2 * q;
#pragma unroll
      for (int dt = 0; dt < ND; ++dt)
        *reinterpret_cast<float2*>(rowp + dt * 8) = make_float2(acc[mt][dt][rr * 2], acc[mt][dt][rr * 2 + 1]);
    }
}

// ---------------------------------------------------------------------------------------------------------
// backward 2: dK, dV, key tile resident; the score block is computed TRANSPOSED (rows = keys) so that P^T and dS^T are
// the A operands of dV += P^T dO and dK += dS^T Q straight from the accumulator fragments
// ---------------------------------------------------------------------------------------------------------
template <int HD, int PASSES, int MT>
__global__ void __launch_bounds__(am_threads(MT), HD == 32 ? 3 : 2) attn_mma_dkv_kernel(AttnArgs a) {
  constexpr int LD = HD + 4, TILE = AM_T * LD, ND = HD / 8;
  extern __shared__ __align__(16) float am_sm[];
  float* Ks = am_sm;              // [64][LD] resident
  float* Vs = Ks + TILE;
  float* QD = Vs + TILE;          // stage s: Q at QD + s*2*TILE, dO at + TILE
  float* LS = QD + 4 * TILE;      // stage s: lse[64] at LS + s*128, dvec[64] at + 64
  const int b = blockIdx.y / a.nh, h = blockIdx.y % a.nh;
  const int j0 = blockIdx.x * AM_T;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
  const float* qb = a.q + (size_t)b * a.Tq * a.ldq + h * HD;
  const float* kb = a.k + (size_t)b * a.Tk * a.ldk + h * HD;
  const float* vb = a.v + (size_t)b * a.Tk * a.ldv + h * HD;
  const float* dob = a.dout + (size_t)b * a.Tq * a.lddo + h * HD;
  const float* lseb = a.lse + (size_t)blockIdx.y * a.Tq;
  const float* dvb_ = a.dvec + (size_t)blockIdx.y * a.Tq;
  const int nit = (a.Tq + AM_T - 1) / AM_T, it0 = am_first_query_tile(a, j0);

  auto load_q = [&](int stage, int it) {
    float* Qn = QD + stage * 2 * TILE;
    am_load_tile<HD>(Qn, qb, a.ldq, it * AM_T, a.Tq);
    am_load_tile<HD>(Qn + TILE, dob, a.lddo, it * AM_T, a.Tq);
    for (int x = threadIdx.x; x < AM_T; x += blockDim.x) {
      const int qi = it * AM_T + x;
      LS[stage * 128 + x] = qi < a.Tq ? lseb[qi] : 0.f;
      LS[stage * 128 + 64 + x] = qi < a.Tq ? dvb_[qi] : 0.f;
    }
  };
  am_load_tile<HD>(Ks, kb, a.ldk, j0, a.Tk);
  am_load_tile<HD>(Vs, vb, a.ldv, j0, a.Tk);
  if (it0 < nit) load_q(0, it0);
  am_commit();

  const int wrow = warp * 16 * MT;
  int row[MT][2], lim[MT][2];   // keys of this thread's rows
  float dk[MT][ND][4], dv[MT][ND][4];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      row[mt][rr] = j0 + wrow + mt * 16 + g + 8 * rr;
      lim[mt][rr] = am_row_limit<false>(a, row[mt][rr]);
    }
#pragma unroll
    for (int dt = 0; dt < ND; ++dt) {
      dk[mt][dt][0] = dk[mt][dt][1] = dk[mt][dt][2] = dk[mt][dt][3] = 0.f;
      dv[mt][dt][0] = dv[mt][dt][1] = dv[mt][dt][2] = dv[mt][dt][3] = 0.f;
    }
  }
  for (int it = it0; it < nit; ++it) {
    const int st = (it - it0) & 1;
    float* Qs = QD + st * 2 * TILE;
    float* dOs = Qs + TILE;
    const float* ls = LS + st * 128;
    if (it + 1 < nit) {
      load_q(st ^ 1, it + 1);
      am_commit();
      am_wait1();
    } else {
      am_wait0();
    }
    __syncthreads();
    float s[MT][8][4], dp[MT][8][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        s[mt][nt][0] = s[mt][nt][1] = s[mt][nt][2] = s[mt][nt][3] = 0.f;
        dp[mt][nt][0] = dp[mt][nt][1] = dp[mt][nt][2] = dp[mt][nt][3] = 0.f;
      }
    am_prod_nt<HD, PASSES, MT>(s, Ks + wrow * LD, Qs, g, q);      // S^T[key][query]
    am_prod_nt<HD, PASSES, MT>(dp, Vs + wrow * LD, dOs, g, q);    // dP^T[key][query]
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      am_finish_scores<false>(s[mt], a, b, row[mt][0], row[mt][1], lim[mt][0], lim[mt][1], it * AM_T, q);
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int col = nt * 8 + 2 * q + (c & 1);
          const float p = ex2_ftz(s[mt][nt][c] - ls[col]);           // queries past Tq and masked entries: s = -inf -> 0
          s[mt][nt][c] = p;
          dp[mt][nt][c] = p * (dp[mt][nt][c] - ls[64 + col]) * a.scale;
        }
    }
    am_prod_acc<HD, PASSES, MT>(dv, s, dOs, g, q);    // dV[key] += sum_i P[i][key] dO[i]
    am_prod_acc<HD, PASSES, MT>(dk, dp, Qs, g, q);    // dK[key] += sum_i dS[i][key] Q[i]
    __syncthreads();
  }
  am_wait0();   // (no query tile sees this key tile: the K / V loads are still in flight)
  float* dkb = a.dk + (size_t)b * a.Tk * a.lddk + h * HD;
  float* dvb = a.dv + (size_t)b * a.Tk * a.lddv + h * HD;
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const int kj = row[mt][rr];
      if (kj >= a.Tk) continue;
      float* rk = dkb + (size_t)kj * a.lddk + 2 * q;
      float* rv = dvb + (size_t)kj * a.lddv + 2 * q;
#pragma unroll
      for (int dt = 0; dt < ND; ++dt) {
        *reinterpret_cast<float2*>(rk + dt * 8) = make_float2(dk[mt][dt][rr * 2], dk[mt][dt][rr * 2 + 1]);
        *reinterpret_cast<float2*>(rv + dt * 8) = make_float2(dv[mt][dt][rr * 2], dv[mt][dt][rr * 2 + 1]);
      }
    }
}

template <int HD>
constexpr size_t am_fwd_smem() { return sizeof(float) * 5 * AM_T * (HD + 4); }
template <int HD>
constexpr size_t am_dq_smem() { return sizeof(float) * 6 * AM_T * (HD + 4); }
template <int HD>
constexpr size_t am_dkv_smem() { return sizeof(float) * (6 * AM_T * (HD + 4) + 256); }

// m-tiles per warp of the three kernels (register budget: the backward kernels hold two 16 MT x 64 blocks)
template <int HD> constexpr int am_mt_fwd() { return AM_MT_FWD; }
template <int HD> constexpr int am_mt_bwd() { return HD == 32 ? AM_MT_BWD32 : 1; }

template <int HD, int PASSES>
static int attn_mma_launch_t(const AttnArgs& a, int backward, cudaStream_t stream) {
  constexpr int MF = am_mt_fwd<HD>(), MB = am_mt_bwd<HD>();
  static bool attr_set = false;
  if (!attr_set) {
    MRG_CUDA_CHECK(cudaFuncSetAttribute(attn_mma_fwd_kernel<HD, PASSES, MF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)am_fwd_smem<HD>()));
    MRG_CUDA_CHECK(cudaFuncSetAttribute(attn_mma_dq_kernel<HD, PASSES, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)am_dq_smem<HD>()));
    MRG_CUDA_CHECK(cudaFuncSetAttribute(attn_mma_dkv_kernel<HD, PASSES, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)am_dkv_smem<HD>()));
    attr_set = true;
  }
  const dim3 gq((a.Tq + AM_T - 1) / AM_T, a.B * a.nh), gk((a.Tk + AM_T - 1) / AM_T, a.B * a.nh);
  if (!backward) {
    count_launch();
    attn_mma_fwd_kernel<HD, PASSES, MF><<<gq, am_threads(MF), am_fwd_smem<HD>(), stream>>>(a);
  } else {
    count_launch(2);
    attn_mma_dq_kernel<HD, PASSES, MB><<<gq, am_threads(MB), am_dq_smem<HD>(), stream>>>(a);
    MRG_CUDA_CHECK(cudaGetLastError());
    attn_mma_dkv_kernel<HD, PASSES, MB><<<gk, am_threads(MB), am_dkv_smem<HD>(), stream>>>(a);
  }
  MRG_CUDA_CHECK(cudaGetLastError());
  return 0;
}

// passes: 3 = 3xTF32 (fp32-grade), 1 = one tf32 pass (the reduced-precision modes)
int attn_mma_launch(const AttnArgs& a, int hd, int backward, int passes, cudaStream_t stream) {
  if (hd == 32) return passes == 1 ? attn_mma_launch_t<32, 1>(a, backward, stream) : attn_mma_launch_t<32, 3>(a, backward, stream);
  return passes == 1 ? attn_mma_launch_t<64, 1>(a, backward, stream) : attn_mma_launch_t<64, 3>(a, backward, stream);
}

}  // namespace mrg
