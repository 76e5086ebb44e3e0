// Fused multi-head attention in fp32 (forward + backward), SURVEY.md §8(f) item 3: the cross-modal attention that
// sits between the LSTM stacks (reference: nn.MultiheadAttention called at mr_gen/model/utils/multi_modal_att.py:12-31
// without masks, and at mr_gen/model/utils/for_sequential.py:25-50 with the causal-rectangular + padding mask that
// mr_gen/model/utils/multi_modal_metaformer.py:32-79 materialises as a [B*heads, L, S] bool tensor).
//
// Why a kernel of our own: the path is fp32 (parity 1e-5 against the reference), and for fp32 torch dispatches to the
// sm_80 "memory-efficient" kernels — 254 us forward / 780 us backward per layer at B=64, 8 heads, T=300, d=32, i.e.
// 3.1 ms of the 13.4 ms SimpleLSTM step, serial between the encoders and the decoder.
//
//  * flash-style: scores never touch HBM.  forward = one pass over the key tiles with an online softmax (log2
//    domain, ex2.approx); backward = two kernels, dQ (query tile resident, loops over key tiles; also produces
//    D_i = dO_i . O_i) and dK/dV (key tile resident, loops over query tiles): no atomics, deterministic.
//  * exact fp32 FMA arithmetic (CUDA cores): 64x64 score tiles, 4x4 register micro-tiles, operands staged in shared
//    memory k-major so that every inner-loop load is one 16-byte (broadcast) access per 16 (or 8) FMAs.
//  * q / k / v / o are read and written IN PLACE in the projections' [B, T, heads*d] layout (row stride given), so no
//    head transpose copies are made on either side, and k / v may be the two halves of one fused projection output.
//  * the mask is a FUNCTION, not a tensor: mode 1 = key j visible to query i iff j / rate <= i (keys run `rate` times
//    faster), mode 2 = iff j <= i / rate (queries run faster), plus "both sides padded" from two per-frame byte
//    vectors; key tiles that are masked for a whole query tile are skipped.
#include <cstddef>
#include <cstdlib>

#include "mrg_common.cuh"

namespace mrg {

constexpr int AT_T = 64;         // queries / keys per tile
constexpr int AT_LD = 68;        // row stride (floats) of the k-major [.][64] shared-memory tiles
constexpr int AT_THREADS = 256;  // 16 x 16 threads, 4 x 4 scores each

// dst[d][r] (k-major, for "sum over d") = src[(r0 + r) * ld + d]; rows past nrows are zero
template <int HD>
__device__ __forceinline__ void load_tile_t(float* dst, const float* __restrict__ src, int ld, int r0, int nrows) {
  constexpr int C4 = HD / 4;
  for (int f = threadIdx.x; f < AT_T * C4; f += AT_THREADS) {
    const int r = f / C4, c = f % C4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r0 + r < nrows) v = __ldg(reinterpret_cast<const float4*>(src + (size_t)(r0 + r) * ld + c * 4));
    float* p = dst + (c * 4) * AT_LD + r;
    p[0] = v.x; p[AT_LD] = v.y; p[2 * AT_LD] = v.z; p[3 * AT_LD] = v.w;
  }
}

// dst[r][d] (row-major, for "sum over r") = src[(r0 + r) * ld + d]
template <int HD>
__device__ __forceinline__ void load_tile_n(float* dst, const float* __restrict__ src, int ld, int r0, int nrows) {
  constexpr int C4 = HD / 4;
  for (int f = threadIdx.x; f < AT_T * C4; f += AT_THREADS) {
    const int r = f / C4, c = f % C4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r0 + r < nrows) v = __ldg(reinterpret_cast<const float4*>(src + (size_t)(r0 + r) * ld + c * 4));
    *reinterpret_cast<float4*>(dst + r * HD + c * 4) = v;
  }
}

// A 64 x HD tile held in registers between its global load and its shared-memory store, so that the load of tile
// t+1 is in flight while tile t is being multiplied (the transposed store cannot be done by cp.async).
template <int HD>
struct TileRegs {
  float4 v[HD / 16];
};
template <int HD>
__device__ __forceinline__ void tile_fetch(TileRegs<HD>& t, const float* __restrict__ src, int ld, int r0, int nrows) {
  constexpr int C4 = HD / 4;
#pragma unroll
  for (int n = 0; n < HD / 16; ++n) {
    const int f = threadIdx.x + n * AT_THREADS, r = f / C4, c = f % C4;
    t.v[n] = r0 + r < nrows ? __ldg(reinterpret_cast<const float4*>(src + (size_t)(r0 + r) * ld + c * 4))
                            : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}
template <int HD>
__device__ __forceinline__ void tile_put_t(float* dst, const TileRegs<HD>& t) {  // dst[d][r], stride AT_LD
  constexpr int C4 = HD / 4;
#pragma unroll
  for (int n = 0; n < HD / 16; ++n) {
    const int f = threadIdx.x + n * AT_THREADS, r = f / C4, c = f % C4;
    float* p = dst + (c * 4) * AT_LD + r;
    p[0] = t.v[n].x; p[AT_LD] = t.v[n].y; p[2 * AT_LD] = t.v[n].z; p[3 * AT_LD] = t.v[n].w;
  }
}
template <int HD>
__device__ __forceinline__ void tile_put_n(float* dst, const TileRegs<HD>& t) {  // dst[r][d], stride HD
  constexpr int C4 = HD / 4;
#pragma unroll
  for (int n = 0; n < HD / 16; ++n) {
    const int f = threadIdx.x + n * AT_THREADS, r = f / C4, c = f % C4;
    *reinterpret_cast<float4*>(dst + r * HD + c * 4) = t.v[n];
  }
}

// acc[i][j] += sum_{k < KD} X[k * XS + i] * Y[k * YS + j]   (X, Y already offset to this thread's 4 rows / NC columns)
template <int KD, int NC, int XS, int YS>
__device__ __forceinline__ void mm_acc(float (&acc)[4][NC], const float* X, const float* Y) {
#pragma unroll 8
  for (int k = 0; k < KD; ++k) {
    const float4 x4 = *reinterpret_cast<const float4*>(X + k * XS);
    const float x[4] = {x4.x, x4.y, x4.z, x4.w};
    float y[NC];
    if constexpr (NC == 4) {
      const float4 y4 = *reinterpret_cast<const float4*>(Y + k * YS);
      y[0] = y4.x; y[1] = y4.y; y[2] = y4.z; y[3] = y4.w;
    } else {
      const float2 y2 = *reinterpret_cast<const float2*>(Y + k * YS);
      y[0] = y2.x; y[1] = y2.y;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < NC; ++j) acc[i][j] = fmaf(x[i], y[j], acc[i][j]);
  }
}

struct MaskCtx {
  int mode, rate;
  bool padded;
};
__device__ __forceinline__ bool at_masked(const MaskCtx& m, int i, int j, unsigned pq, unsigned pk) {
  bool r = false;
  if (m.mode == 1) r = (j / m.rate) > i;
  else if (m.mode == 2) r = j > (i / m.rate);
  return r || (pq & pk);
}
// key tiles a query tile [i0, i1] can see
__device__ __forceinline__ int at_key_tiles(const AttnArgs& a, int i1) {
  int n = (a.Tk + AT_T - 1) / AT_T;
  if (a.mask_mode == 1) n = min(n, (int)((((long long)i1 + 1) * a.rate - 1) / AT_T) + 1);
  else if (a.mask_mode == 2) n = min(n, (i1 / a.rate) / AT_T + 1);
  return n;
}
// first query tile that can see key tile starting at j0
__device__ __forceinline__ int at_first_query_tile(const AttnArgs& a, int j0) {
  if (a.mask_mode == 1) return (j0 / a.rate) / AT_T;
  if (a.mask_mode == 2) return (int)(((long long)j0 * a.rate) / AT_T);
  return 0;
}

// scores of this thread's 4 x 4 micro-tile in the log2 domain, -inf where masked / out of range
template <int HD>
__device__ __forceinline__ void at_scores(float (&s)[4][4], const float* Qt, const float* Kt, const AttnArgs& a,
                                          const MaskCtx& mc, int b, int i0, int j0, int ty, int tx) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
  mm_acc<HD, 4, AT_LD, AT_LD>(s, Qt + ty * 4, Kt + tx * 4);
  unsigned pq[4] = {0, 0, 0, 0}, pk[4] = {0, 0, 0, 0};
  if (mc.padded) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int i = i0 + ty * 4 + e, j = j0 + tx * 4 + e;
      pq[e] = i < a.Tq ? a.pad_q[(size_t)b * a.Tq + i] : 0u;
      pk[e] = j < a.Tk ? a.pad_k[(size_t)b * a.Tk + j] : 0u;
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int qi = i0 + ty * 4 + i, kj = j0 + tx * 4 + j;
      const bool ok = qi < a.Tq && kj < a.Tk && !at_masked(mc, qi, kj, pq[i], pk[j]);
      s[i][j] = ok ? s[i][j] * a.scale_log2 : -INFINITY;
    }
}

// ---------------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(AT_THREADS, HD == 32 ? 3 : 2) attn_fwd_kernel(AttnArgs a) {
  constexpr int NC = HD / 16;
  extern __shared__ __align__(16) float at_sm[];
  float* Qt = at_sm;                  // [HD][AT_LD]
  float* Kt = Qt + HD * AT_LD;        // [HD][AT_LD]
  float* Vs = Kt + HD * AT_LD;        // [64][HD]
  float* Pt = Vs + AT_T * HD;         // [64 keys][AT_LD queries]
  const int b = blockIdx.y / a.nh, h = blockIdx.y % a.nh;
  const int i0 = blockIdx.x * AT_T;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const float* qb = a.q + (size_t)b * a.Tq * a.ldq + h * HD;
  const float* kb = a.k + (size_t)b * a.Tk * a.ldk + h * HD;
  const float* vb = a.v + (size_t)b * a.Tk * a.ldv + h * HD;
  const MaskCtx mc = {a.mask_mode, a.rate, a.pad_q != nullptr};

  load_tile_t<HD>(Qt, qb, a.ldq, i0, a.Tq);
  float m[4], l[4], acc[4][NC];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    m[i] = -INFINITY;
    l[i] = 0.f;
#pragma unroll
    for (int c = 0; c < NC; ++c) acc[i][c] = 0.f;
  }
  const int njt = at_key_tiles(a, min(i0 + AT_T, a.Tq) - 1);
  TileRegs<HD> kr, vr;
  tile_fetch<HD>(kr, kb, a.ldk, 0, a.Tk);
  tile_fetch<HD>(vr, vb, a.ldv, 0, a.Tk);
  for (int jt = 0; jt < njt; ++jt) {
    const int j0 = jt * AT_T;
    __syncthreads();  // the previous tile's Kt / Vs / Pt have been consumed (also orders the Qt fill)
    tile_put_t<HD>(Kt, kr);
    tile_put_n<HD>(Vs, vr);
    if (jt + 1 < njt) {  // next tile's global loads fly during this tile's arithmetic
      tile_fetch<HD>(kr, kb, a.ldk, j0 + AT_T, a.Tk);
      tile_fetch<HD>(vr, vb, a.ldv, j0 + AT_T, a.Tk);
    }
    __syncthreads();
    float s[4][4];
    at_scores<HD>(s, Qt, Kt, a, mc, b, i0, j0, ty, tx);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float mx = fmaxf(fmaxf(s[i][0], s[i][1]), fmaxf(s[i][2], s[i][3]));
#pragma unroll
      for (int o = 1; o < 16; o <<= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      const float mn = fmaxf(m[i], mx);
      const float alpha = mn == -INFINITY ? 1.f : ex2_ftz(m[i] - mn);
      float rs = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        s[i][j] = mn == -INFINITY ? 0.f : ex2_ftz(s[i][j] - mn);
        rs += s[i][j];
      }
#pragma unroll
      for (int o = 1; o < 16; o <<= 1) rs += __shfl_xor_sync(0xffffffffu, rs, o);
      l[i] = l[i] * alpha + rs;
      m[i] = mn;
#pragma unroll
      for (int c = 0; c < NC; ++c) acc[i][c] *= alpha;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
      *reinterpret_cast<float4*>(Pt + (tx * 4 + j) * AT_LD + ty * 4) = make_float4(s[0][j], s[1][j], s[2][j], s[3][j]);
    __syncthreads();
    mm_acc<AT_T, NC, AT_LD, HD>(acc, Pt + ty * 4, Vs + tx * NC);
  }
  float* ob = a.o + (size_t)b * a.Tq * a.ldo + h * HD;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int qi = i0 + ty * 4 + i;
    if (qi >= a.Tq) continue;
    const float inv = l[i] > 0.f ? 1.f / l[i] : 0.f;  // a query with no visible key gives 0 (torch: NaN)
    float* orow = ob + (size_t)qi * a.ldo + tx * NC;
    if constexpr (NC == 4) *reinterpret_cast<float4*>(orow) = make_float4(acc[i][0] * inv, acc[i][1] * inv, acc[i][2] * inv, acc[i][3] * inv);
    else *reinterpret_cast<float2*>(orow) = make_float2(acc[i][0] * inv, acc[i][1] * inv);
    if (tx == 0 && a.lse) a.lse[(size_t)blockIdx.y * a.Tq + qi] = l[i] > 0.f ? m[i] + log2f(l[i]) : 0.f;
  }
}

// ---------------------------------------------------------------------------------------------------------
// backward 1: dQ (and D = dO . O), query tile resident
// ---------------------------------------------------------------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(AT_THREADS, HD == 32 ? 3 : 2) attn_bwd_dq_kernel(AttnArgs a) {
  constexpr int NC = HD / 16;
  extern __shared__ __align__(16) float at_sm[];
  float* Qt = at_sm;                   // [HD][AT_LD]
  float* dOt = Qt + HD * AT_LD;        // [HD][AT_LD]
  float* Kt = dOt + HD * AT_LD;        // [HD][AT_LD]
  float* Vt = Kt + HD * AT_LD;         // [HD][AT_LD]
  float* Ks = Vt + HD * AT_LD;         // [64][HD]
  float* dSt = Ks + AT_T * HD;         // [64 keys][AT_LD queries]
  const int b = blockIdx.y / a.nh, h = blockIdx.y % a.nh;
  const int i0 = blockIdx.x * AT_T;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const float* qb = a.q + (size_t)b * a.Tq * a.ldq + h * HD;
  const float* kb = a.k + (size_t)b * a.Tk * a.ldk + h * HD;
  const float* vb = a.v + (size_t)b * a.Tk * a.ldv + h * HD;
  const float* ob = a.o + (size_t)b * a.Tq * a.ldo + h * HD;
  const float* dob = a.dout + (size_t)b * a.Tq * a.lddo + h * HD;
  const MaskCtx mc = {a.mask_mode, a.rate, a.pad_q != nullptr};

  load_tile_t<HD>(Qt, qb, a.ldq, i0, a.Tq);
  load_tile_t<HD>(dOt, dob, a.lddo, i0, a.Tq);
  float dvec[4], lse[4], acc[4][NC];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int qi = i0 + ty * 4 + i;
    float d = 0.f;
    lse[i] = 0.f;
    if (qi < a.Tq) {
#pragma unroll
      for (int c = 0; c < NC; ++c)
        d = fmaf(__ldg(dob + (size_t)qi * a.lddo + tx * NC + c), __ldg(ob + (size_t)qi * a.ldo + tx * NC + c), d);
      lse[i] = a.lse[(size_t)blockIdx.y * a.Tq + qi];
    }
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
    dvec[i] = d;
    if (tx == 0 && qi < a.Tq) a.dvec[(size_t)blockIdx.y * a.Tq + qi] = d;
#pragma unroll
    for (int c = 0; c < NC; ++c) acc[i][c] = 0.f;
  }
  const int njt = at_key_tiles(a, min(i0 + AT_T, a.Tq) - 1);
  TileRegs<HD> kr, vr;
  tile_fetch<HD>(kr, kb, a.ldk, 0, a.Tk);
  tile_fetch<HD>(vr, vb, a.ldv, 0, a.Tk);
  for (int jt = 0; jt < njt; ++jt) {
    const int j0 = jt * AT_T;
    __syncthreads();
    tile_put_t<HD>(Kt, kr);
    tile_put_n<HD>(Ks, kr);
    tile_put_t<HD>(Vt, vr);
    if (jt + 1 < njt) {
      tile_fetch<HD>(kr, kb, a.ldk, j0 + AT_T, a.Tk);
      tile_fetch<HD>(vr, vb, a.ldv, j0 + AT_T, a.Tk);
    }
    __syncthreads();
    float s[4][4], dp[4][4];
    at_scores<HD>(s, Qt, Kt, a, mc, b, i0, j0, ty, tx);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) dp[i][j] = 0.f;
    mm_acc<HD, 4, AT_LD, AT_LD>(dp, dOt + ty * 4, Vt + tx * 4);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float p = ex2_ftz(s[i][j] - lse[i]);  // masked: ex2(-inf) = 0
        s[i][j] = p * (dp[i][j] - dvec[i]) * a.scale;
      }
#pragma unroll
    for (int j = 0; j < 4; ++j)
      *reinterpret_cast<float4*>(dSt + (tx * 4 + j) * AT_LD + ty * 4) = make_float4(s[0][j], s[1][j], s[2][j], s[3][j]);
    __syncthreads();
    mm_acc<AT_T, NC, AT_LD, HD>(acc, dSt + ty * 4, Ks + tx * NC);
  }
  float* dqb = a.dq + (size_t)b * a.Tq * a.lddq + h * HD;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int qi = i0 + ty * 4 + i;
    if (qi >= a.Tq) continue;
    float* row = dqb + (size_t)qi * a.lddq + tx * NC;
    if constexpr (NC == 4) *reinterpret_cast<float4*>(row) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    else *reinterpret_cast<float2*>(row) = make_float2(acc[i][0], acc[i][1]);
  }
}

// ---------------------------------------------------------------------------------------------------------
// backward 2: dK, dV, key tile resident
// ---------------------------------------------------------------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(AT_THREADS, HD == 32 ? 3 : 2) attn_bwd_dkv_kernel(AttnArgs a) {
  constexpr int NC = HD / 16;
  extern __shared__ __align__(16) float at_sm[];
  float* Kt = at_sm;                   // [HD][AT_LD]
  float* Vt = Kt + HD * AT_LD;
  float* Qt = Vt + HD * AT_LD;
  float* dOt = Qt + HD * AT_LD;
  float* Qs = dOt + HD * AT_LD;        // [64][HD]
  float* dOs = Qs + AT_T * HD;         // [64][HD]
  float* Ps = dOs + AT_T * HD;         // [64 queries][AT_LD keys]
  float* dSs = Ps + AT_T * AT_LD;      // [64 queries][AT_LD keys]
  const int b = blockIdx.y / a.nh, h = blockIdx.y % a.nh;
  const int j0 = blockIdx.x * AT_T;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const float* qb = a.q + (size_t)b * a.Tq * a.ldq + h * HD;
  const float* kb = a.k + (size_t)b * a.Tk * a.ldk + h * HD;
  const float* vb = a.v + (size_t)b * a.Tk * a.ldv + h * HD;
  const float* dob = a.dout + (size_t)b * a.Tq * a.lddo + h * HD;
  const MaskCtx mc = {a.mask_mode, a.rate, a.pad_q != nullptr};

  load_tile_t<HD>(Kt, kb, a.ldk, j0, a.Tk);
  load_tile_t<HD>(Vt, vb, a.ldv, j0, a.Tk);
  float dk[4][NC], dv[4][NC];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int c = 0; c < NC; ++c) dk[i][c] = dv[i][c] = 0.f;
  const int nit = (a.Tq + AT_T - 1) / AT_T, it0 = at_first_query_tile(a, j0);
  TileRegs<HD> qr, dr;
  tile_fetch<HD>(qr, qb, a.ldq, it0 * AT_T, a.Tq);
  tile_fetch<HD>(dr, dob, a.lddo, it0 * AT_T, a.Tq);
  for (int it = it0; it < nit; ++it) {
    const int i0 = it * AT_T;
    __syncthreads();
    tile_put_t<HD>(Qt, qr);
    tile_put_n<HD>(Qs, qr);
    tile_put_t<HD>(dOt, dr);
    tile_put_n<HD>(dOs, dr);
    if (it + 1 < nit) {
      tile_fetch<HD>(qr, qb, a.ldq, i0 + AT_T, a.Tq);
      tile_fetch<HD>(dr, dob, a.lddo, i0 + AT_T, a.Tq);
    }
    __syncthreads();
    float s[4][4], dp[4][4];
    at_scores<HD>(s, Qt, Kt, a, mc, b, i0, j0, ty, tx);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) dp[i][j] = 0.f;
    mm_acc<HD, 4, AT_LD, AT_LD>(dp, dOt + ty * 4, Vt + tx * 4);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int qi = i0 + ty * 4 + i;
      float lse = 0.f, dvec = 0.f;
      if (qi < a.Tq) {
        lse = a.lse[(size_t)blockIdx.y * a.Tq + qi];
        dvec = a.dvec[(size_t)blockIdx.y * a.Tq + qi];
      }
      float p[4], ds[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        p[j] = ex2_ftz(s[i][j] - lse);  // rows past Tq and masked entries: s = -inf -> 0
        ds[j] = p[j] * (dp[i][j] - dvec) * a.scale;
      }
      *reinterpret_cast<float4*>(Ps + (ty * 4 + i) * AT_LD + tx * 4) = make_float4(p[0], p[1], p[2], p[3]);
      *reinterpret_cast<float4*>(dSs + (ty * 4 + i) * AT_LD + tx * 4) = make_float4(ds[0], ds[1], ds[2], ds[3]);
    }
    __syncthreads();
    mm_acc<AT_T, NC, AT_LD, HD>(dv, Ps + ty * 4, dOs + tx * NC);   // dV[j] += sum_i P[i][j] dO[i]
    mm_acc<AT_T, NC, AT_LD, HD>(dk, dSs + ty * 4, Qs + tx * NC);   // dK[j] += sum_i dS[i][j] Q[i]
  }
  float* dkb = a.dk + (size_t)b * a.Tk * a.lddk + h * HD;
  float* dvb = a.dv + (size_t)b * a.Tk * a.lddv + h * HD;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int kj = j0 + ty * 4 + i;
    if (kj >= a.Tk) continue;
    float* rk = dkb + (size_t)kj * a.lddk + tx * NC;
    float* rv = dvb + (size_t)kj * a.lddv + tx * NC;
    if constexpr (NC == 4) {
      *reinterpret_cast<float4*>(rk) = make_float4(dk[i][0], dk[i][1], dk[i][2], dk[i][3]);
      *reinterpret_cast<float4*>(rv) = make_float4(dv[i][0], dv[i][1], dv[i][2], dv[i][3]);
    } else {
      *reinterpret_cast<float2*>(rk) = make_float2(dk[i][0], dk[i][1]);
      *reinterpret_cast<float2*>(rv) = make_float2(dv[i][0], dv[i][1]);
    }
  }
}

template <int HD>
constexpr size_t at_fwd_smem() { return sizeof(float) * (2 * HD * AT_LD + AT_T * HD + AT_T * AT_LD); }
template <int HD>
constexpr size_t at_dq_smem() { return sizeof(float) * (4 * HD * AT_LD + AT_T * HD + AT_T * AT_LD); }
template <int HD>
constexpr size_t at_dkv_smem() { return sizeof(float) * (4 * HD * AT_LD + 2 * AT_T * HD + 2 * AT_T * AT_LD); }

template <int HD>
static int attn_launch(const AttnArgs& a, int backward, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    MRG_CUDA_CHECK(cudaFuncSetAttribute(attn_fwd_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)at_fwd_smem<HD>()));
    MRG_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd_dq_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)at_dq_smem<HD>()));
    MRG_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd_dkv_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)at_dkv_smem<HD>()));
    attr_set = true;
  }
  const dim3 gq((a.Tq + AT_T - 1) / AT_T, a.B * a.nh), gk((a.Tk + AT_T - 1) / AT_T, a.B * a.nh);
  if (!backward) {
    count_launch();
    attn_fwd_kernel<HD><<<gq, AT_THREADS, at_fwd_smem<HD>(), stream>>>(a);
  } else {
    count_launch(2);
    attn_bwd_dq_kernel<HD><<<gq, AT_THREADS, at_dq_smem<HD>(), stream>>>(a);
    MRG_CUDA_CHECK(cudaGetLastError());
    attn_bwd_dkv_kernel<HD><<<gk, AT_THREADS, at_dkv_smem<HD>(), stream>>>(a);
  }
  MRG_CUDA_CHECK(cudaGetLastError());
  return 0;
}

static int attn_check(const char* who, const AttnArgs& a, int hd) {
  MRG_REQUIRE(a.B > 0 && a.nh > 0 && a.Tq > 0 && a.Tk > 0, "%s: empty problem", who);
  MRG_REQUIRE(hd == 32 || hd == 64, "%s: head_dim %d not built (32 and 64 are)", who, hd);
  MRG_REQUIRE(a.B * (long long)a.nh <= 65535, "%s: batch x heads = %lld exceeds the grid limit", who, a.B * (long long)a.nh);
  MRG_REQUIRE(a.mask_mode >= 0 && a.mask_mode <= 2 && (a.mask_mode == 0 || a.rate >= 1), "%s: bad mask mode / rate", who);
  MRG_REQUIRE((a.pad_q == nullptr) == (a.pad_k == nullptr), "%s: pad_q and pad_k go together", who);
  return 0;
}
int attn_mma_launch(const AttnArgs& a, int hd, int backward, int passes, cudaStream_t stream);   // mrg_attention_mma.cu

// 0: warp-level tensor cores, 3xTF32 (default, fp32-grade) | 1: one tf32 pass (reduced-precision modes) |
// 2: the CUDA-core fp32 kernels of this file (cross-check; MRG_ATTENTION_SIMT=1 forces it)
static int g_attn_mode = 0;
static int attn_dispatch(const AttnArgs& a, int hd, int backward, cudaStream_t stream) {
  static int force_simt = -1;
  if (force_simt < 0) {
    const char* e = getenv("MRG_ATTENTION_SIMT");
    force_simt = (e && e[0] == '1') ? 1 : 0;
  }
  if (!force_simt && g_attn_mode != 2) return attn_mma_launch(a, hd, backward, g_attn_mode == 1 ? 1 : 3, stream);
  return hd == 32 ? attn_launch<32>(a, backward, stream) : attn_launch<64>(a, backward, stream);
}
#define AT_ALIGNED(p, ld) ((p) != nullptr && (((uintptr_t)(p)) & 15) == 0 && (ld) % 4 == 0 && (ld) >= nh * hd)

}  // namespace mrg

extern "C" int mrg_attention_set_mode(int mode) {
  MRG_REQUIRE(mode >= 0 && mode <= 2, "mrg_attention_set_mode: 0 (3xTF32 tensor cores), 1 (one tf32 pass) or 2 (CUDA cores)");
  mrg::g_attn_mode = mode;
  return 0;
}

extern "C" int mrg_attention_forward(const float* q, int ldq, const float* k, int ldk, const float* v, int ldv, float* o,
                                     int ldo, float* lse, int B, int nh, int Tq, int Tk, int hd, float scale,
                                     int mask_mode, int rate, const uint8_t* pad_q, const uint8_t* pad_k,
                                     void* stream) {
  mrg::AttnArgs a = {};
  a.q = q; a.k = k; a.v = v; a.o = o; a.lse = lse;
  a.B = B; a.nh = nh; a.Tq = Tq; a.Tk = Tk;
  a.ldq = ldq; a.ldk = ldk; a.ldv = ldv; a.ldo = ldo;
  a.scale = scale; a.scale_log2 = scale * 1.4426950408889634f;
  a.mask_mode = mask_mode; a.rate = rate; a.pad_q = pad_q; a.pad_k = pad_k;
  if (int e = mrg::attn_check("mrg_attention_forward", a, hd)) return e;
  MRG_REQUIRE(AT_ALIGNED(q, ldq) && AT_ALIGNED(k, ldk) && AT_ALIGNED(v, ldv) && AT_ALIGNED(o, ldo),
              "mrg_attention_forward: q/k/v/o must be 16-byte aligned with row strides that are multiples of 4");
  return mrg::attn_dispatch(a, hd, 0, (cudaStream_t)stream);
}

extern "C" int mrg_attention_backward(const float* q, int ldq, const float* k, int ldk, const float* v, int ldv,
                                      const float* o, int ldo, const float* lse, const float* dout, int lddo,
                                      float* dq, int lddq, float* dk, int lddk, float* dv, int lddv, float* dvec,
                                      int B, int nh, int Tq, int Tk, int hd, float scale, int mask_mode, int rate,
                                      const uint8_t* pad_q, const uint8_t* pad_k, void* stream) {
  mrg::AttnArgs a = {};
  a.q = q; a.k = k; a.v = v; a.o = const_cast<float*>(o); a.lse = const_cast<float*>(lse);
  a.dout = dout; a.dvec = dvec; a.dq = dq; a.dk = dk; a.dv = dv;
  a.B = B; a.nh = nh; a.Tq = Tq; a.Tk = Tk;
  a.ldq = ldq; a.ldk = ldk; a.ldv = ldv; a.ldo = ldo; a.lddo = lddo; a.lddq = lddq; a.lddk = lddk; a.lddv = lddv;
  a.scale = scale; a.scale_log2 = scale * 1.4426950408889634f;
  a.mask_mode = mask_mode; a.rate = rate; a.pad_q = pad_q; a.pad_k = pad_k;
  if (int e = mrg::attn_check("mrg_attention_backward", a, hd)) return e;
  MRG_REQUIRE(lse && dvec, "mrg_attention_backward: lse / dvec missing");
  MRG_REQUIRE(AT_ALIGNED(q, ldq) && AT_ALIGNED(k, ldk) && AT_ALIGNED(v, ldv) && AT_ALIGNED(o, ldo) &&
                  AT_ALIGNED(dout, lddo) && AT_ALIGNED(dq, lddq) && AT_ALIGNED(dk, lddk) && AT_ALIGNED(dv, lddv),
              "mrg_attention_backward: tensors must be 16-byte aligned with row strides that are multiples of 4");
  return mrg::attn_dispatch(a, hd, 1, (cudaStream_t)stream);
}
