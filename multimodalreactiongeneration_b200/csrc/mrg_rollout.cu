// Persistent on-device rollout of lstm_with_sampling (sm_100a): forward and backward-through-time.
//
// Replaces the Python time loop of the reference,
//   mr_gen/model/lstm_with_sampling/lstm_with_sample.py:379-408 (head_motion_generation) and :410-433
//   (generate_one_step -> forward with T = 1),
// for the part of a step that depends on the previous step.  What a step of the reference computes (quirks Q2, Q5,
// Q6 of SURVEY.md Appendix C kept):
//   prev_t  = t == 0 ? ms[0] : (mask[t-1] ? y_{t-1} : ms[t-1])                      feedback select (:397,:404)
//   x_0     = W_f [s_t | partner_t | prev_t] + b_f                                   feature_projection (:228)
//   x_{l+1} = LayerNorm_l(cell_l(x_l) + x_l),   cell_l = nn.LSTM step from ZERO state (Q2):
//             c = sig(i) * tanh(g), h = sig(o) * tanh(c), (i, f, g, o) = W_ih^l x_l + b_ih^l + b_hh^l
//             (W_hh and the forget gate are inert: h_0 = c_0 = 0)                    lstm_block.py:101-107, :165-169
//   y_t     = W_2 relu(W_1 x_L + b_1) + b_2                                          feed_forward (:230)
// The sampler LSTM (the only carried state) does not depend on the feedback, so it runs once over lead + sequence in
// the recurrent kernel (mrg_rec_fwd2.cu) and enters here through `base` = W_f[:, :Hs+P] [s | partner] + b_f, a
// time-parallel GEMM.  This kernel owns everything that is serial in t.
//
// Design (same ownership as the recurrent kernels): a thread-block cluster of CL = H/32 CTAs shares a group of batch
// rows; CTA c owns hidden units [32c, 32c+32) of EVERY predictor layer.  The three live gate rows (i, g, o) of the
// owned units — 96 x H weights per layer — stay on chip for the whole sequence: the g / o rows in REGISTERS (thread
// tile = 1 unit x H/16 k-values), the i rows in shared memory (one conflict-free LDS.128 per 4 k-values); the forget
// rows are never loaded.  Per layer: FFMA2 partial sums -> shared-memory reduction over the k-slices -> gates + cell
// update + residual (one warp per row, lane = unit) -> the 32 pre-LayerNorm values of every row are written into the
// shared memory of ALL CTAs of the cluster with 16-byte st.async stores that signal the destination's mbarrier
// (complete_tx) -> every CTA waits for its own window to fill and normalises the full row redundantly (it needs all H
// values as the next layer's input).  The bottleneck FFN is split by output unit the same way; the pose, the feedback
// select and x_0 are computed redundantly per CTA (a few hundred FMAs), so a step has NL + 1 exchanges and no other
// communication — no cluster barrier and no fence inside the time loop (the first version used barrier.cluster:
// its release fence + arrive / wait round trip was 28 % of the kernel, profiles/r2_rollout_*).  HBM sees only the
// per-step inputs (base, ground-truth pose, mask) and, in training, the reserve for the backward.
//
// Window reuse without extra synchronisation: the exchanges of a step use DIFFERENT windows in a fixed cyclic order
// (v0, v1, f | stats, partials, ... , prev).  A peer can run at most one exchange ahead of this CTA — to send exchange
// e+1 it must have completed the wait of exchange e, which needs this CTA's e data, which is sent only after this CTA
// has consumed window e-1 — so a window is never written before its previous content was read, and an mbarrier never
// receives bytes of its next phase before the current one completed.  Every window's mbarrier completes once per
// step; it is re-armed (arrive.expect_tx) at the top of the step; bytes that land earlier only drive the
// transaction count negative until then.
//
// Backward (rollout_bwd_kernel): the same clusters walk t = T-1 .. 0 with the chain dy -> FFN^T -> LN^T -> cell^T ->
// W_ih^T -> ... -> W_prev^T -> (mask) -> dy_{t-1} (Q6: gradients flow through fed-back poses).  The transposed
// mat-vec keeps the same weight slice (own gate rows x all k), produces partial dx over the CTA's 96 gate rows and
// reduce-scatters them to the CTA that owns each k (fixed summation order: deterministic); LayerNorm's two row sums
// and the P-wide W_prev^T product are small all-reduces over the same shared-memory windows.  Everything that is
// time-parallel (all weight gradients) is left to the tensor-core GEMMs: the kernel stores d(pre-activations).
#include <cstdlib>

#include "mrg_common.cuh"

namespace mrg {

constexpr int RO_THREADS = 512;
constexpr int RO_RCAP = 8;   // batch rows a cluster advances together (one pass over t)
constexpr int RO_LMAX = 2;   // predictor blocks held on chip

struct RolloutArgs {
  const float* base;           // [T][B][H]
  const float* gt_prev;        // [T][B][P]  ground-truth previous pose of every step (one-frame lag, Q5)
  const unsigned char* mask;   // [T][B]     mask[t] != 0: step t+1 is fed y_t (nullptr = never)
  const float* w_prev;         // [H][P], row stride w_prev_ld
  long long w_prev_ld;
  const float* w_ih[RO_LMAX];  // [4H][H]
  const float* b_ih[RO_LMAX];  // [4H] or nullptr
  const float* b_hh[RO_LMAX];
  const float* ln_g[RO_LMAX];  // [H]
  const float* ln_b[RO_LMAX];
  const float *w1, *b1, *w2, *b2;  // [FB][H], [FB], [P][FB], [P]
  float eps;
  int T, B, P, FB, relu, train, slices;
  float* pred;    // [T][B][P]
  float* xs;      // [NL+1][T][B][H]
  float* gates;   // [NL][T][B][3][H]   post-activation i, g, o
  float* xhat;    // [NL][T][B][H]
  float* rstd;    // [NL][T][B]
  float* fact;    // [T][B][FB]         FFN hidden, post-activation
  float* prev;    // [T][B][P]
  // backward only
  const float* dpred;  // [T][B][P]
  float* dy;           // [T][B][P]
  float* df;           // [T][B][FB]
  float* dpre;         // [NL][T][B][4H]
  float* dbase;        // [T][B][H]
  float* dprev;        // [T][B][P]
  float* dln_g;        // [NL][B][H]
  float* dln_b;        // [NL][B][H]
};

__device__ __forceinline__ void ro_cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void ro_cp_async4(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void ro_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void ro_cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ float ro_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__host__ __device__ inline int ro_align4(int n) { return (n + 3) & ~3; }
// n / d for n < 1024, d <= 64 with one multiply (inv = ro_inv(d)): the runtime divisors of the kernels (pose width,
// bottleneck width) would otherwise put a ~130-clock integer division on the serial path of every step
__host__ __device__ inline unsigned ro_inv(int d) { return ((1u << 20) + (unsigned)d - 1u) / (unsigned)d; }
__device__ __forceinline__ int ro_div(int n, unsigned inv) { return (int)(((unsigned)n * inv) >> 20); }

// Developer-only phase timing (-DMRG_RO_TRACE): thread 0 of CTA 0 accumulates the clocks it spends between the phase
// marks of a step into a global array that mrg_debug_rollout_trace() copies out.
#ifdef MRG_RO_TRACE
__device__ unsigned long long g_ro_trace[2][32];
#define RO_TRACE_DECL unsigned long long ro_t0 = clock64();
#define RO_MARK(dir, k)                                                    \
  if (blockIdx.x == 0 && threadIdx.x == 0) {                               \
    const unsigned long long ro_t1 = clock64();                            \
    g_ro_trace[dir][k] += ro_t1 - ro_t0;                                   \
    ro_t0 = ro_t1;                                                         \
  }
#else
#define RO_TRACE_DECL
#define RO_MARK(dir, k)
#endif

// ---------------------------------------------------------------------------------------------------------
// shared-memory layouts (offsets in floats, every block 16-byte aligned)
// ---------------------------------------------------------------------------------------------------------
struct RoFwdLayout {
  int wi, part, xfull, vbuf, base, fbuf, w1, w2, wprev, lng, lnb, bias, b1, b2, prevsm, bars, total;
};
__host__ __device__ inline RoFwdLayout ro_fwd_layout(int H, int NL, int P, int FB) {
  const int CL = H / 32, KSL = H >= 64 ? 16 : 8, PP = P | 1, FBc = FB / CL;
  RoFwdLayout l;
  int o = 0;
  l.wi = o;    o += NL * H * 32;                    // gate-i rows of the owned units: [NL][H/4][32] float4
  l.part = o;  o += (KSL / 2) * RO_RCAP * 3 * 32;   // partial gate sums [k-slice][row][gate][unit]
  l.xfull = o; o += RO_RCAP * H;                    // input vector of the current layer, all rows
  l.vbuf = o;  o += 2 * RO_RCAP * H;                // pre-LayerNorm rows written by all CTAs (two windows)
  l.base = o;  o += RO_RCAP * H;                    // cp.async landing zone of base[t+1]
  l.fbuf = o;  o += RO_RCAP * ro_align4(FB);        // FFN hidden written by all CTAs
  l.w1 = o;    o += ro_align4(FBc * H);
  l.w2 = o;    o += ro_align4(P * FB);
  l.wprev = o; o += ro_align4(H * PP);
  l.lng = o;   o += NL * H;
  l.lnb = o;   o += NL * H;
  l.bias = o;  o += NL * 3 * 32;
  l.b1 = o;    o += ro_align4(FBc);
  l.b2 = o;    o += ro_align4(P);
  l.prevsm = o; o += ro_align4(RO_RCAP * P);        // the pose fed to the current step, transposed [p][row]
  l.bars = o;  o += 8;                              // mbarriers: v window 0, v window 1, FFN window (8 bytes each)
  l.total = o;
  return l;
}

// partial gate sums of NR rows over this thread's k-slice (forward mat-vec): i rows from shared memory, g / o rows
// from registers; the two k sub-slices of a warp meet by shuffle, sub-slice s stores the rows r with (r & 1) == s
template <int H, int NCH, int NR>
__device__ __forceinline__ void ro_fwd_rows(const float4* __restrict__ wi_l, const float4 (&wg)[NCH],
                                            const float4 (&wo)[NCH], const float* __restrict__ xrow, float* part_row,
                                            int unit, int subk) {
  float2 acc[3][NR];
#pragma unroll
  for (int g = 0; g < 3; ++g)
#pragma unroll
    for (int r = 0; r < NR; ++r) acc[g][r] = make_float2(0.f, 0.f);
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const float4 wi = wi_l[c * 32 + unit];
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      const float4 x4 = *reinterpret_cast<const float4*>(xrow + r * H + 4 * c);
      const float2 xlo = make_float2(x4.x, x4.y), xhi = make_float2(x4.z, x4.w);
      ffma2(acc[0][r], make_float2(wi.x, wi.y), xlo);
      ffma2(acc[0][r], make_float2(wi.z, wi.w), xhi);
      ffma2(acc[1][r], make_float2(wg[c].x, wg[c].y), xlo);
      ffma2(acc[1][r], make_float2(wg[c].z, wg[c].w), xhi);
      ffma2(acc[2][r], make_float2(wo[c].x, wo[c].y), xlo);
      ffma2(acc[2][r], make_float2(wo[c].z, wo[c].w), xhi);
    }
  }
#pragma unroll
  for (int g = 0; g < 3; ++g)
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      float v = acc[g][r].x + acc[g][r].y;
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      if ((r & 1) == subk) part_row[(r * 3 + g) * 32 + unit] = v;
    }
}

// =========================================================================================================
// forward
// =========================================================================================================
template <int H, int NL>
__global__ void __launch_bounds__(RO_THREADS, 1) rollout_fwd_kernel(RolloutArgs a) {
  constexpr int CL = H / 32;
  constexpr int KSL = H >= 64 ? 16 : 8;  // k-slices of the mat-vec (one thread covers KT consecutive k)
  constexpr int KT = H / KSL;
  constexpr int NCH = KT / 4;
  constexpr int RC = RO_RCAP;
  extern __shared__ __align__(16) float sm[];
  const int P = a.P, FB = a.FB, FBc = FB / CL, PP = P | 1, T = a.T, B = a.B;
  const RoFwdLayout lay = ro_fwd_layout(H, NL, P, FB);
  float4* wi_sm = reinterpret_cast<float4*>(sm + lay.wi);
  float* part = sm + lay.part;
  float* xfull = sm + lay.xfull;
  float* vbuf = sm + lay.vbuf;
  float* basebuf = sm + lay.base;
  float* fbuf = sm + lay.fbuf;
  float* w1s = sm + lay.w1;
  float* w2s = sm + lay.w2;
  float* wprev = sm + lay.wprev;
  float* lng = sm + lay.lng;
  float* lnb = sm + lay.lnb;
  float* bias = sm + lay.bias;
  float* b1s = sm + lay.b1;
  float* b2s = sm + lay.b2;
  float* prevsm = sm + lay.prevsm;
  const uint32_t bar0 = smem_u32(sm + lay.bars);   // + 8 * window
  const int FBa = ro_align4(FB);
  constexpr int FIT = 4;   // rounds of phase F: 64 (row, pose) items per round, RO_RCAP * 32 items at most
  const unsigned invP = ro_inv(P), invFBc = ro_inv(FBc);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rank = (int)cluster_ctarank();
  const int cid = blockIdx.x / CL;
  const int j0 = rank * 32;
  const int base_rows = B / a.slices, rem_rows = B % a.slices;
  const int crow0 = cid * base_rows + min(cid, rem_rows);
  const int nrows = base_rows + (cid < rem_rows ? 1 : 0);

  // mat-vec coordinates: lane = (unit within a half, k sub-slice), warp = (k slice pair, unit half)
  const int u16 = lane & 15, subk = lane >> 4;
  const int ks = warp % (KSL / 2), uhalf = warp / (KSL / 2);
  const bool mv = warp < KSL;
  const int unit = (uhalf & 1) * 16 + u16;
  const int k0 = (ks * 2 + subk) * KT;

  // ---- weights on chip -------------------------------------------------------------------------------------------
  float4 wr[NL][2][NCH];  // g and o rows of `unit`, k0 .. k0+KT-1
#pragma unroll
  for (int l = 0; l < NL; ++l)
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      if (mv) {
        wr[l][0][c] = __ldg(reinterpret_cast<const float4*>(a.w_ih[l] + (size_t)(2 * H + j0 + unit) * H + k0 + 4 * c));
        wr[l][1][c] = __ldg(reinterpret_cast<const float4*>(a.w_ih[l] + (size_t)(3 * H + j0 + unit) * H + k0 + 4 * c));
      } else {
        wr[l][0][c] = make_float4(0.f, 0.f, 0.f, 0.f);
        wr[l][1][c] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  for (int idx = tid; idx < NL * (H / 4) * 32; idx += RO_THREADS) {
    const int un = idx & 31, kc = (idx >> 5) % (H / 4), l = idx / ((H / 4) * 32);
    wi_sm[idx] = __ldg(reinterpret_cast<const float4*>(a.w_ih[l] + (size_t)(j0 + un) * H + kc * 4));
  }
  for (int idx = tid; idx < NL * 3 * 32; idx += RO_THREADS) {
    const int un = idx & 31, g = (idx >> 5) % 3, l = idx / 96;
    const int row = (g == 0 ? 0 : g == 1 ? 2 * H : 3 * H) + j0 + un;
    float b = 0.f;
    if (a.b_ih[l]) b += a.b_ih[l][row];
    if (a.b_hh[l]) b += a.b_hh[l][row];
    bias[idx] = b;
  }
  for (int idx = tid; idx < NL * H; idx += RO_THREADS) {
    const int l = idx / H, k = idx % H;
    lng[idx] = a.ln_g[l][k];
    lnb[idx] = a.ln_b[l][k];
  }
  for (int idx = tid; idx < FBc * H; idx += RO_THREADS) w1s[idx] = a.w1[(size_t)(rank * FBc + idx / H) * H + idx % H];
  for (int idx = tid; idx < P * FB; idx += RO_THREADS) w2s[idx] = a.w2[idx];
  for (int idx = tid; idx < H * P; idx += RO_THREADS) wprev[(idx / P) * PP + idx % P] = a.w_prev[(size_t)(idx / P) * a.w_prev_ld + idx % P];
  for (int idx = tid; idx < FBc; idx += RO_THREADS) b1s[idx] = a.b1 ? a.b1[rank * FBc + idx] : 0.f;
  for (int idx = tid; idx < P; idx += RO_THREADS) b2s[idx] = a.b2 ? a.b2[idx] : 0.f;
  if (tid == 0) {
    for (int i = 0; i < 3; ++i) mbar_init(bar0 + 8 * i, 1);
    mbar_init_fence();
  }
  __syncthreads();
  cluster_sync_all();  // every CTA of the cluster is resident (and its mbarriers exist) before anybody sends to it
  const uint32_t sm_base = smem_u32(sm);
  uint32_t it = 0;  // steps done so far (all passes): every window's mbarrier completes once per step
  RO_TRACE_DECL

  const int npass = (nrows + RC - 1) / RC;
  const int rpp = npass > 0 ? (nrows + npass - 1) / npass : 0;
  const size_t TBH = (size_t)T * B * H;
  for (int pass = 0; pass < npass; ++pass) {
    const int prow0 = crow0 + pass * rpp;
    const int R = min(rpp, crow0 + nrows - prow0);
    // ---- state of step 0: ground-truth pose, no feedback; base[0] in flight ----
    for (int i = tid; i < RC * P; i += RO_THREADS) prevsm[i] = 0.f;   // rows beyond R stay finite
    __syncthreads();
    if (tid < R * P) {   // transposed window: prevsm[p][row]
      const int r = ro_div(tid, invP), pp = tid - r * P;
      prevsm[pp * RC + r] = a.gt_prev[(size_t)prow0 * P + tid];
    }
    for (int e = tid; e < R * H / 4; e += RO_THREADS)
      ro_cp_async16(smem_u32(basebuf + 4 * e), a.base + (size_t)prow0 * H + 4 * e);
    ro_cp_async_commit();
    const int nF = R * P;   // phase F items (row, pose component): 8 lanes each, 4 items per warp, <= FIT rounds

    for (int t = 0; t < T; ++t) {
      const size_t tb = (size_t)t * B + prow0;  // first row of this pass at step t
      const uint32_t par = it & 1u;
      ++it;
      if (tid == 0) {  // this step's three exchanges: bytes every window will receive (all CTAs, itself included)
#pragma unroll
        for (int l = 0; l < NL; ++l) mbar_arrive_expect_tx(bar0 + 8 * (l & 1), (uint32_t)(R * H * 4));
        mbar_arrive_expect_tx(bar0 + 16, (uint32_t)(R * FB * 4));
      }
      // ================= phase A: x_0 = base + W_prev prev =================
      ro_cp_async_wait_all();
      __syncthreads();
      RO_MARK(0, 0)
      // every phase that all 16 warps run costs (instructions x 4) issue cycles: one thread per column k computes all
      // rows (one weight load + two 16-byte loads of the transposed pose window feed 8 FMAs)
      if (tid < H) {
        float acc[RC];
#pragma unroll
        for (int r = 0; r < RC; ++r) acc[r] = basebuf[r * H + tid];
        const float* wp = wprev + tid * PP;
        for (int p = 0; p < P; ++p) {
          const float w = wp[p];
          const float4 pa = *reinterpret_cast<const float4*>(prevsm + p * RC);
          const float4 pb = *reinterpret_cast<const float4*>(prevsm + p * RC + 4);
          acc[0] = fmaf(w, pa.x, acc[0]); acc[1] = fmaf(w, pa.y, acc[1]);
          acc[2] = fmaf(w, pa.z, acc[2]); acc[3] = fmaf(w, pa.w, acc[3]);
          acc[4] = fmaf(w, pb.x, acc[4]); acc[5] = fmaf(w, pb.y, acc[5]);
          acc[6] = fmaf(w, pb.z, acc[6]); acc[7] = fmaf(w, pb.w, acc[7]);
        }
#pragma unroll
        for (int r = 0; r < RC; ++r)
          if (r < R) {
            xfull[r * H + tid] = acc[r];
            if (a.train && (tid >> 5) == rank) a.xs[(tb + r) * H + tid] = acc[r];
          }
      }
      if (a.train && rank == 0 && tid < nF) {
        const int r = ro_div(tid, invP), pp = tid - r * P;
        a.prev[tb * P + tid] = prevsm[pp * RC + r];
      }
      RO_MARK(0, 20)
      __syncthreads();
      RO_MARK(0, 21)
      RO_MARK(0, 22)
      if (t + 1 < T)
        for (int e = tid; e < R * H / 4; e += RO_THREADS)
          ro_cp_async16(smem_u32(basebuf + 4 * e), a.base + ((size_t)(t + 1) * B + prow0) * H + 4 * e);
      ro_cp_async_commit();
      RO_MARK(0, 1)

#pragma unroll
      for (int l = 0; l < NL; ++l) {
        // ================= phase B: partial gate sums over this thread's k-slice =================
        if (mv) {
          const float4* wi_l = wi_sm + (l * (H / 4) + (k0 >> 2)) * 32;
          for (int rg = 0; rg < R;) {   // row groups of 4, then a 2- or 1-row tail (no padded row slots)
            const float* xrow = xfull + rg * H + k0;
            float* prow = part + (ks * RC + rg) * 96;
            const int left = R - rg;
            if (left >= 3) { ro_fwd_rows<H, NCH, 4>(wi_l, wr[l][0], wr[l][1], xrow, prow, unit, subk); rg += 4; }
            else if (left == 2) { ro_fwd_rows<H, NCH, 2>(wi_l, wr[l][0], wr[l][1], xrow, prow, unit, subk); rg += 2; }
            else { ro_fwd_rows<H, NCH, 1>(wi_l, wr[l][0], wr[l][1], xrow, prow, unit, subk); rg += 1; }
          }
        }
        RO_MARK(0, 2 + 5 * l)
        __syncthreads();
        RO_MARK(0, 3 + 5 * l)
        // ================= phase C: gates, cell, residual; pre-LayerNorm values to every CTA =================
        float* vwin = vbuf + (l & 1) * RC * H;
        if (warp < R) {
          const int r = warp;
          float pi = bias[(l * 3 + 0) * 32 + lane], pg = bias[(l * 3 + 1) * 32 + lane], po = bias[(l * 3 + 2) * 32 + lane];
#pragma unroll
          for (int s = 0; s < KSL / 2; ++s) {
            pi += part[((s * RC + r) * 3 + 0) * 32 + lane];
            pg += part[((s * RC + r) * 3 + 1) * 32 + lane];
            po += part[((s * RC + r) * 3 + 2) * 32 + lane];
          }
          const float gi = fast_sigmoid(pi), gg = fast_tanh(pg), go = fast_sigmoid(po);
          const float h = go * fast_tanh(gi * gg);
          const float v = h + xfull[r * H + j0 + lane];
          if (a.train) {
            float* gp = a.gates + (((size_t)l * T * B + tb + r) * 3) * H + j0 + lane;
            gp[0] = gi; gp[H] = gg; gp[2 * H] = go;
          }
          float4 hv;
          hv.x = __shfl_sync(0xffffffffu, v, (lane & ~3));
          hv.y = __shfl_sync(0xffffffffu, v, (lane & ~3) + 1);
          hv.z = __shfl_sync(0xffffffffu, v, (lane & ~3) + 2);
          hv.w = __shfl_sync(0xffffffffu, v, (lane & ~3) + 3);
          const uint32_t off = smem_u32(vwin + r * H + j0 + (lane & ~3)) - sm_base;
          const uint32_t boff = bar0 + 8 * (l & 1) - sm_base;
#pragma unroll
          for (int i = 0; i < 2; ++i) {   // this lane's destinations: CTAs (lane & 3) and (lane & 3) + 4
            const int dst = (lane & 3) + 4 * i;
            if (dst < CL) {
              const uint32_t peer = map_to_cta(sm_base, (uint32_t)dst);
              st_async_v4(peer + off, hv, peer + boff);
            }
          }
        }
        RO_MARK(0, 4 + 5 * l)
        // ================= phase D: LayerNorm of the full row (every CTA, redundantly) =================
        if (warp < R) {
          const int r = warp;
          mbar_wait(bar0 + 8 * (l & 1), par);   // all CL slices of all R rows have landed in this CTA's window
          RO_MARK(0, 5 + 5 * l)
          // one butterfly for both moments: sums of (v - c) and (v - c)^2 around a sample c of the row (the row's
          // first value), so that E[d^2] - E[d]^2 does not cancel (|c - mean| is a few sigma at most)
          float v[H / 32];
#pragma unroll
          for (int i = 0; i < H / 32; ++i) v[i] = vwin[r * H + lane + 32 * i];
          const float c0 = __shfl_sync(0xffffffffu, v[0], 0);
          float s1 = 0.f, s2 = 0.f;
#pragma unroll
          for (int i = 0; i < H / 32; ++i) { const float d = v[i] - c0; s1 += d; s2 = fmaf(d, d, s2); }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
          }
          const float md = s1 * (1.0f / H);
          const float mean = c0 + md;
          const float var = fmaxf(fmaf(-md, md, s2 * (1.0f / H)), 0.f) + a.eps;
          float rs = rsqrtf(var);
          rs = rs * fmaf(-0.5f * var * rs, rs, 1.5f);   // one Newton step: full fp32 accuracy
#pragma unroll
          for (int i = 0; i < H / 32; ++i) {
            const int k = lane + 32 * i;
            const float xh = (v[i] - mean) * rs;
            const float out = fmaf(xh, lng[l * H + k], lnb[l * H + k]);
            xfull[r * H + k] = out;
            if (a.train && i == rank) {
              a.xs[(size_t)(l + 1) * TBH + (tb + r) * H + k] = out;
              a.xhat[(size_t)l * TBH + (tb + r) * H + k] = xh;
            }
          }
          if (a.train && rank == 0 && lane == 0) a.rstd[(size_t)l * T * B + tb + r] = rs;
        }
        __syncthreads();
        RO_MARK(0, 6 + 5 * l)
      }

      // loads for the feedback select at the end of this step, by the threads that use them.  Issued HERE, after the
      // last mat-vec: values that are live across the register-hungry mat-vec get spilled right after the load, and
      // the spill store waits for the load (a global round trip on the serial path — measured 860 clocks per step)
      float gt_n[FIT];
      unsigned m_n[FIT];
#pragma unroll
      for (int j = 0; j < FIT; ++j) {
        gt_n[j] = 0.f;
        m_n[j] = 0u;
        if ((tid >> 5) * 4 + 64 * j >= nF) continue;   // warp-uniform: most warps have no item
        const int item = (tid >> 5) * 4 + 64 * j + ((tid >> 3) & 3);
        if ((tid & 7) == 0 && item < nF && t + 1 < T) {
          const int r = ro_div(item, invP);
          gt_n[j] = a.gt_prev[((size_t)(t + 1) * B + prow0) * P + item];
          if (a.mask) m_n[j] = a.mask[tb + r];
        }
      }
      // ================= phase E: bottleneck FFN, this CTA's FBc hidden units =================
      // 8 lanes per (row, unit) item, 4 items per warp = one float4 of the row, sent straight from registers
      for (int i0 = (tid >> 5) * 4; i0 < R * FBc; i0 += (RO_THREADS / 32) * 4) {  // warp-uniform (R*FBc % 4 == 0)
        const int l8 = tid & 7;
        const int r = ro_div(i0, invFBc), o0 = i0 - r * FBc;
        const int o = o0 + ((tid >> 3) & 3);
        float sa = 0.f, sb = 0.f;
#pragma unroll
        for (int k = 0; k < H; k += 64) {
          if (k + l8 * 4 < H) {
            const float4 w4 = *reinterpret_cast<const float4*>(w1s + o * H + k + l8 * 4);
            const float4 x4 = *reinterpret_cast<const float4*>(xfull + r * H + k + l8 * 4);
            sa = fmaf(w4.x, x4.x, sa); sa = fmaf(w4.y, x4.y, sa); sa = fmaf(w4.z, x4.z, sa); sa = fmaf(w4.w, x4.w, sa);
          }
          if (k + 32 + l8 * 4 < H) {
            const float4 w4 = *reinterpret_cast<const float4*>(w1s + o * H + k + 32 + l8 * 4);
            const float4 x4 = *reinterpret_cast<const float4*>(xfull + r * H + k + 32 + l8 * 4);
            sb = fmaf(w4.x, x4.x, sb); sb = fmaf(w4.y, x4.y, sb); sb = fmaf(w4.z, x4.z, sb); sb = fmaf(w4.w, x4.w, sb);
          }
        }
        float sv = sa + sb;
        sv += __shfl_xor_sync(0xffffffffu, sv, 4);
        sv += __shfl_xor_sync(0xffffffffu, sv, 2);
        sv += __shfl_xor_sync(0xffffffffu, sv, 1);
        sv += b1s[o];
        if (a.relu) sv = fmaxf(sv, 0.f);
        float4 f4;
        f4.x = __shfl_sync(0xffffffffu, sv, 0);
        f4.y = __shfl_sync(0xffffffffu, sv, 8);
        f4.z = __shfl_sync(0xffffffffu, sv, 16);
        f4.w = __shfl_sync(0xffffffffu, sv, 24);
        if (lane < CL) {
          const uint32_t peer = map_to_cta(sm_base, (uint32_t)lane);
          st_async_v4(peer + (smem_u32(fbuf + r * FBa + rank * FBc + o0) - sm_base), f4, peer + (bar0 + 16 - sm_base));
        }
        if (a.train && lane == 0) *reinterpret_cast<float4*>(a.fact + (tb + r) * FB + rank * FBc + o0) = f4;
      }
      RO_MARK(0, 13)
      mbar_wait(bar0 + 16, par);   // the FFN hidden of all rows is complete in this CTA's window
      RO_MARK(0, 14)
      // ================= phase F: pose (every CTA), feedback select for the next step =================
#pragma unroll
      for (int j = 0; j < FIT; ++j) {
        const int i0 = (tid >> 5) * 4 + 64 * j;
        if (i0 < nF) {   // warp-uniform
          const int item = i0 + ((tid >> 3) & 3), l8 = tid & 7;
          const bool ok = item < nF;
          const int r = ok ? ro_div(item, invP) : 0, p = ok ? item - r * P : 0;
          float sa = 0.f, sb = 0.f;
          for (int o = l8; o < FB; o += 16) {
            sa = fmaf(w2s[p * FB + o], fbuf[r * FBa + o], sa);
            if (o + 8 < FB) sb = fmaf(w2s[p * FB + o + 8], fbuf[r * FBa + o + 8], sb);
          }
          float sv = sa + sb;
          sv += __shfl_xor_sync(0xffffffffu, sv, 4);
          sv += __shfl_xor_sync(0xffffffffu, sv, 2);
          sv += __shfl_xor_sync(0xffffffffu, sv, 1);
          if (ok && l8 == 0) {
            sv += b2s[p];
            if (rank == 0) a.pred[tb * P + item] = sv;
            prevsm[p * RC + r] = m_n[j] ? sv : gt_n[j];   // fed to step t+1 (read after the __syncthreads of phase A)
          }
        }
      }
      RO_MARK(0, 15)
      // (phase A of the next step starts with a __syncthreads)
    }
    __syncthreads();
  }
  cluster_sync_all();  // nobody leaves while a peer may still write into its shared memory
}

// =========================================================================================================
// backward through time
// =========================================================================================================
struct RoBwdLayout {
  int wi, part, dpre, dv, dx, rbuf, sbuf, pbuf, pl, w1t, w2, wprev, lng, pf, dysm, dfsm, dpn, bars, total;
};
__host__ __device__ inline RoBwdLayout ro_bwd_layout(int H, int NL, int P, int FB) {
  const int CL = H / 32, PP = P | 1, P4 = ro_align4(P);
  RoBwdLayout l;
  int o = 0;
  l.wi = o;    o += NL * H * 32;             // gate-i rows, transposed tiles: [NL][16 j-pairs][H/2 k-pairs] float4
  l.part = o;  o += 4 * RO_RCAP * H;         // partial dx [j-group][row][k]
  l.dpre = o;  o += RO_RCAP * 3 * 32;        // d(pre-activations) of the owned units [row][gate][unit]
  l.dv = o;    o += RO_RCAP * 32;            // residual branch of dx
  l.dx = o;    o += RO_RCAP * 32;            // gradient at the owned units of the current layer's output
  l.rbuf = o;  o += CL * RO_RCAP * 32;       // reduce-scatter window [source CTA][row][unit]
  l.sbuf = o;  o += ro_align4(CL * RO_RCAP * 2);   // LayerNorm row sums [source CTA][row][2]
  l.pbuf = o;  o += CL * RO_RCAP * P4;       // W_prev^T partial products [source CTA][row][p]
  l.pl = o;    o += RO_RCAP * P4;            // this CTA's partial products
  l.w1t = o;   o += FB * 32;                 // W_1[:, owned units] as [FB][32]
  l.w2 = o;    o += ro_align4(P * FB);
  l.wprev = o; o += ro_align4(32 * PP);      // W_prev rows of the owned units
  l.lng = o;   o += NL * 32;
  l.pf = o;                                  // cp.async landing zones (two steps): xhat, i, g, o, rstd, d(pose), FFN hidden
  o += 2 * (NL * 4 * RO_RCAP * 32 + ro_align4(NL * RO_RCAP) + ro_align4(RO_RCAP * P) + RO_RCAP * ro_align4(FB));
  l.dysm = o;  o += ro_align4(RO_RCAP * P);
  l.dfsm = o;  o += RO_RCAP * ro_align4(FB);
  l.dpn = o;   o += ro_align4(RO_RCAP * P);  // d(prev) of the step processed before (t+1)
  l.bars = o;  o += 2 * (2 * RO_LMAX + 1);   // mbarriers: row sums [layer], partials [layer], prev (8 bytes each)
  l.total = o;
  return l;
}

// partial dx of NR rows over this thread's 8 owned units (backward mat-vec): columns 2kp, 2kp+1
template <int H, int NR>
__device__ __forceinline__ void ro_bwd_rows(const float4* __restrict__ wi_l, const float2 (&wg)[4][2],
                                            const float2 (&wo)[4][2], const float* __restrict__ drow, float* part_row) {
  float2 acc[NR][2];
#pragma unroll
  for (int r = 0; r < NR; ++r) acc[r][0] = acc[r][1] = make_float2(0.f, 0.f);
#pragma unroll
  for (int jp = 0; jp < 4; ++jp) {
    const float4 wi = wi_l[jp * (H / 2)];
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      const float* dr = drow + r * 96 + 2 * jp;
      const float2 di = *reinterpret_cast<const float2*>(dr);
      const float2 dg = *reinterpret_cast<const float2*>(dr + 32);
      const float2 dd = *reinterpret_cast<const float2*>(dr + 64);
      ffma2(acc[r][0], make_float2(wi.x, wi.y), di);
      ffma2(acc[r][1], make_float2(wi.z, wi.w), di);
      ffma2(acc[r][0], wg[jp][0], dg);
      ffma2(acc[r][1], wg[jp][1], dg);
      ffma2(acc[r][0], wo[jp][0], dd);
      ffma2(acc[r][1], wo[jp][1], dd);
    }
  }
#pragma unroll
  for (int r = 0; r < NR; ++r)
    *reinterpret_cast<float2*>(part_row + r * H) =
        make_float2(acc[r][0].x + acc[r][0].y, acc[r][1].x + acc[r][1].y);
}

template <int H, int NL>
__global__ void __launch_bounds__(RO_THREADS, 1) rollout_bwd_kernel(RolloutArgs a) {
  constexpr int CL = H / 32;
  constexpr int RC = RO_RCAP;
  constexpr int KP = H / 2;           // k-pairs: a mat-vec thread owns columns 2kp, 2kp+1 and 8 of the 32 owned units
  constexpr int MVT = 4 * KP;         // mat-vec threads (4 unit groups)
  extern __shared__ __align__(16) float sm[];
  const int P = a.P, FB = a.FB, PP = P | 1, P4 = ro_align4(P), FBa = ro_align4(FB), T = a.T, B = a.B;
  const unsigned invP = ro_inv(P), invFB = ro_inv(FB), invP4q = ro_inv(P4 / 4);
  const RoBwdLayout lay = ro_bwd_layout(H, NL, P, FB);
  float4* wi_sm = reinterpret_cast<float4*>(sm + lay.wi);
  float* part = sm + lay.part;
  float* dpre_sm = sm + lay.dpre;
  float* dv_sm = sm + lay.dv;
  float* dx_sm = sm + lay.dx;
  float* rbuf = sm + lay.rbuf;
  float* sbuf = sm + lay.sbuf;
  float* pbuf = sm + lay.pbuf;
  float* pl = sm + lay.pl;
  float* w1t = sm + lay.w1t;
  float* w2s = sm + lay.w2;
  float* wprev = sm + lay.wprev;
  float* lng = sm + lay.lng;
  float* pf = sm + lay.pf;
  float* dysm = sm + lay.dysm;
  float* dfsm = sm + lay.dfsm;
  float* dpn = sm + lay.dpn;
  const uint32_t sm_base = smem_u32(sm);
  const uint32_t bar0 = smem_u32(sm + lay.bars);
  auto sbar = [&](int l) { return bar0 + 8u * (uint32_t)l; };             // row sums of layer l
  auto rbar = [&](int l) { return bar0 + 8u * (uint32_t)(RO_LMAX + l); };   // partial dx of layer l
  const uint32_t pbar = bar0 + 8u * (2 * RO_LMAX);                          // W_prev^T products

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rank = (int)cluster_ctarank();
  const int cid = blockIdx.x / CL;
  const int j0 = rank * 32;
  const int base_rows = B / a.slices, rem_rows = B % a.slices;
  const int crow0 = cid * base_rows + min(cid, rem_rows);
  const int nrows = base_rows + (cid < rem_rows ? 1 : 0);
  const size_t TBH = (size_t)T * B * H;

  // ---- weights on chip -------------------------------------------------------------------------------------------
  const bool mv = tid < MVT;
  const int kp = tid % KP, jg = tid / KP;  // columns 2kp, 2kp+1; owned units 8jg .. 8jg+7
  // wr[l][gate g/o][jp][c]: (W[j][2kp+c], W[j+1][2kp+c]) with j = 8jg + 2jp
  float2 wr[NL][2][4][2];
#pragma unroll
  for (int l = 0; l < NL; ++l)
#pragma unroll
    for (int g = 0; g < 2; ++g)
#pragma unroll
      for (int jp = 0; jp < 4; ++jp) {
        float2 r0 = make_float2(0.f, 0.f), r1 = make_float2(0.f, 0.f);
        if (mv) {
          const size_t row = (size_t)((g == 0 ? 2 * H : 3 * H) + j0 + 8 * jg + 2 * jp);
          r0 = __ldg(reinterpret_cast<const float2*>(a.w_ih[l] + row * H + 2 * kp));
          r1 = __ldg(reinterpret_cast<const float2*>(a.w_ih[l] + (row + 1) * H + 2 * kp));
        }
        wr[l][g][jp][0] = make_float2(r0.x, r1.x);
        wr[l][g][jp][1] = make_float2(r0.y, r1.y);
      }
  for (int idx = tid; idx < NL * 16 * KP; idx += RO_THREADS) {
    const int k2 = idx % KP, jp = (idx / KP) % 16, l = idx / (16 * KP);
    const float2 r0 = __ldg(reinterpret_cast<const float2*>(a.w_ih[l] + (size_t)(j0 + 2 * jp) * H + 2 * k2));
    const float2 r1 = __ldg(reinterpret_cast<const float2*>(a.w_ih[l] + (size_t)(j0 + 2 * jp + 1) * H + 2 * k2));
    wi_sm[idx] = make_float4(r0.x, r1.x, r0.y, r1.y);
  }
  for (int idx = tid; idx < FB * 32; idx += RO_THREADS) w1t[idx] = a.w1[(size_t)(idx >> 5) * H + j0 + (idx & 31)];
  for (int idx = tid; idx < P * FB; idx += RO_THREADS) w2s[idx] = a.w2[idx];
  for (int idx = tid; idx < 32 * P; idx += RO_THREADS) wprev[(idx / P) * PP + idx % P] = a.w_prev[(size_t)(j0 + idx / P) * a.w_prev_ld + idx % P];
  for (int idx = tid; idx < NL * 32; idx += RO_THREADS) lng[idx] = a.ln_g[idx >> 5][j0 + (idx & 31)];
  if (tid == 0) {
    for (int i = 0; i < 2 * RO_LMAX + 1; ++i) mbar_init(bar0 + 8 * i, 1);
    mbar_init_fence();
  }
  __syncthreads();
  cluster_sync_all();
  uint32_t it = 0;  // steps done so far (all passes): every mbarrier completes once per step
  RO_TRACE_DECL

  const int npass = (nrows + RC - 1) / RC;
  const int rpp = npass > 0 ? (nrows + npass - 1) / npass : 0;
  for (int pass = 0; pass < npass; ++pass) {
    const int prow0 = crow0 + pass * rpp;
    const int R = min(rpp, crow0 + nrows - prow0);
    float dgam[NL], dbet[NL];  // thread (row = warp, unit = lane): sums over t
#pragma unroll
    for (int l = 0; l < NL; ++l) dgam[l] = dbet[l] = 0.f;
    if (tid < RC * P) dpn[tid] = 0.f;
    for (int i = tid; i < RC * P4; i += RO_THREADS) pl[i] = 0.f;   // the padding columns travel with the rest
    // reserve (xhat, i, g, o of every layer) for thread (row, unit), fetched with cp.async TWO steps ahead into
    // alternating landing zones: one group per step, so the group of the step being processed is the older one
    const int ZG = NL * 4 * RC * 32, ZR = ro_align4(NL * RC), ZP = ro_align4(RC * P);
    const int ZSZ = ZG + ZR + ZP + RC * FBa;
    auto fetch_reserve = [&](int ts) {
      if (ts >= 0) {
        float* z = pf + (ts & 1) * ZSZ;
        const size_t tbp = (size_t)ts * B + prow0;
        if (warp < R) {
          const size_t tbn = tbp + warp;
#pragma unroll
          for (int l = 0; l < NL; ++l) {
            ro_cp_async4(smem_u32(z + ((l * 4 + 0) * RC + warp) * 32 + lane), a.xhat + (size_t)l * TBH + tbn * H + j0 + lane);
            const float* gp = a.gates + (((size_t)l * T * B + tbn) * 3) * H + j0 + lane;
            ro_cp_async4(smem_u32(z + ((l * 4 + 1) * RC + warp) * 32 + lane), gp);
            ro_cp_async4(smem_u32(z + ((l * 4 + 2) * RC + warp) * 32 + lane), gp + H);
            ro_cp_async4(smem_u32(z + ((l * 4 + 3) * RC + warp) * 32 + lane), gp + 2 * H);
            if (lane == 0) ro_cp_async4(smem_u32(z + ZG + l * RC + warp), a.rstd + (size_t)l * T * B + tbn);
          }
        }
        if ((tid & ~31) < R * P && tid < R * P) ro_cp_async4(smem_u32(z + ZG + ZR + tid), a.dpred + tbp * P + tid);
        if ((tid & ~31) < R * FB && tid < R * FB) ro_cp_async4(smem_u32(z + ZG + ZR + ZP + tid), a.fact + tbp * FB + tid);
      }
      ro_cp_async_commit();
    };
    fetch_reserve(T - 1);
    fetch_reserve(T - 2);
    unsigned m_cur = 0u;   // raw byte of mask[t] for the step being processed (0 for t = T-1: nothing was fed on)
    __syncthreads();

    for (int t = T - 1; t >= 0; --t) {
      const size_t tb = (size_t)t * B + prow0;
      const uint32_t par = it & 1u;
      ++it;
      if (tid == 0) {  // bytes every window of this step will receive (from all CL CTAs, itself included)
#pragma unroll
        for (int l = 0; l < NL; ++l) {
          mbar_arrive_expect_tx(sbar(l), (uint32_t)(CL * R * 8));
          mbar_arrive_expect_tx(rbar(l), (uint32_t)(CL * R * 128));
        }
        mbar_arrive_expect_tx(pbar, (uint32_t)(CL * R * P4 * 4));
      }
      // the staged inputs of this step (issued two steps ago) have landed for everybody
      asm volatile("cp.async.wait_group 1;" ::: "memory");
      __syncthreads();
      const float* z = pf + (t & 1) * ZSZ;
      // ================= phase 0: total gradient at the pose =================
      // mask[t] says whether y_t was fed to step t+1 (whose d(prev) sits in dpn)
      if (tid < R * P) {
        const float dyv = z[ZG + ZR + tid] + (m_cur != 0u ? dpn[tid] : 0.f);
        dysm[tid] = dyv;
        if (rank == 0) a.dy[tb * P + tid] = dyv;
      }
      __syncthreads();
      RO_MARK(1, 0)
      // ================= phase 1: FFN hidden gradient (every CTA, all FB units) =================
      if (tid < R * FB) {
        const int r = ro_div(tid, invFB), o = tid - r * FB;
        float s = 0.f;
        for (int p = 0; p < P; ++p) s = fmaf(w2s[p * FB + o], dysm[r * P + p], s);
        if (a.relu && !(z[ZG + ZR + ZP + tid] > 0.f)) s = 0.f;
        dfsm[r * FBa + o] = s;
        if (rank == 0) a.df[tb * FB + tid] = s;
      }
      __syncthreads();
      RO_MARK(1, 1)
      // ================= phase 2: gradient at x_L, owned units =================
      for (int d0 = (tid >> 5) * 8; d0 < R * 32; d0 += (RO_THREADS / 32) * 8) {  // warp-uniform trip count
        const int d = d0 + ((tid >> 2) & 7), sub = tid & 3;
        const int r = d >> 5, k = d & 31;
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;   // independent chains: the loads pipeline
        for (int o = sub; o < FB; o += 16) {
          s0 = fmaf(w1t[o * 32 + k], dfsm[r * FBa + o], s0);
          if (o + 4 < FB) s1 = fmaf(w1t[(o + 4) * 32 + k], dfsm[r * FBa + o + 4], s1);
          if (o + 8 < FB) s2 = fmaf(w1t[(o + 8) * 32 + k], dfsm[r * FBa + o + 8], s2);
          if (o + 12 < FB) s3 = fmaf(w1t[(o + 12) * 32 + k], dfsm[r * FBa + o + 12], s3);
        }
        float s = (s0 + s1) + (s2 + s3);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        if (sub == 0) dx_sm[d] = s;
      }
      __syncthreads();
      RO_MARK(1, 2)

#pragma unroll
      for (int l = NL - 1; l >= 0; --l) {
        // ================= phase 3: LayerNorm^T row sums, exchanged between the CTAs =================
        float gdx = 0.f, xh = 0.f, gi = 0.f, gg = 0.f, go = 0.f;
        if (warp < R) {
          const int r = warp;
          const float dxo = dx_sm[r * 32 + lane];
          xh = z[((l * 4 + 0) * RC + r) * 32 + lane];
          gi = z[((l * 4 + 1) * RC + r) * 32 + lane];
          gg = z[((l * 4 + 2) * RC + r) * 32 + lane];
          go = z[((l * 4 + 3) * RC + r) * 32 + lane];
          gdx = dxo * lng[l * 32 + lane];
          dgam[l] = fmaf(dxo, xh, dgam[l]);
          dbet[l] += dxo;
          const float s1 = ro_warp_sum(gdx), s2 = ro_warp_sum(gdx * xh);
          if (lane < CL) {
            const uint32_t peer = map_to_cta(sm_base, (uint32_t)lane);
            const uint32_t addr = peer + (smem_u32(sbuf + (rank * RC + r) * 2) - sm_base);
            st_async_f32(addr, s1, peer + (sbar(l) - sm_base));
            st_async_f32(addr + 4, s2, peer + (sbar(l) - sm_base));
          }
        }
        RO_MARK(1, 3 + 6 * l)
        // ================= phase 4: LayerNorm^T, cell^T (owned units) =================
        if (warp < R) {
          const int r = warp;
          mbar_wait(sbar(l), par);   // the row sums of all CTAs have landed
          RO_MARK(1, 4 + 6 * l)
          float s1 = 0.f, s2 = 0.f;
#pragma unroll
          for (int c = 0; c < CL; ++c) { s1 += sbuf[(c * RC + r) * 2]; s2 += sbuf[(c * RC + r) * 2 + 1]; }
          const float dv = z[ZG + l * RC + r] * (gdx - s1 * (1.0f / H) - xh * s2 * (1.0f / H));
          const float c = gi * gg, tc = fast_tanh(c);
          const float d_o = dv * tc * go * (1.f - go);
          const float dc = dv * go * (1.f - tc * tc);
          const float d_i = dc * gg * gi * (1.f - gi);
          const float d_g = dc * gi * (1.f - gg * gg);
          dv_sm[r * 32 + lane] = dv;
          dpre_sm[(r * 3 + 0) * 32 + lane] = d_i;
          dpre_sm[(r * 3 + 1) * 32 + lane] = d_g;
          dpre_sm[(r * 3 + 2) * 32 + lane] = d_o;
          float* dp = a.dpre + ((size_t)l * T * B + tb + r) * 4 * H + j0 + lane;   // torch gate order i, f, g, o
          dp[0] = d_i; dp[H] = 0.f; dp[2 * H] = d_g; dp[3 * H] = d_o;
        }
        __syncthreads();
        RO_MARK(1, 5 + 6 * l)
        // ================= phase 5: W_ih^T over the owned gate rows -> partial dx for every k =================
        if (mv) {
          const float4* wi_l = wi_sm + (l * 16 + jg * 4) * KP + kp;
          for (int rg = 0; rg < R;) {   // row groups of 4, then a 2- or 1-row tail
            const float* drow = dpre_sm + rg * 96 + 8 * jg;
            float* prow = part + (jg * RC + rg) * H + 2 * kp;
            const int left = R - rg;
            if (left >= 3) { ro_bwd_rows<H, 4>(wi_l, wr[l][0], wr[l][1], drow, prow); rg += 4; }
            else if (left == 2) { ro_bwd_rows<H, 2>(wi_l, wr[l][0], wr[l][1], drow, prow); rg += 2; }
            else { ro_bwd_rows<H, 1>(wi_l, wr[l][0], wr[l][1], drow, prow); rg += 1; }
          }
        }
        __syncthreads();
        RO_MARK(1, 6 + 6 * l)
        // sum the 4 unit groups and hand every k to the CTA that owns it (window slot = this CTA's rank)
        for (int i = tid; i < R * (H / 4); i += RO_THREADS) {
          const int r = i / (H / 4), k4 = (i % (H / 4)) * 4;
          float4 s = *reinterpret_cast<const float4*>(part + (0 * RC + r) * H + k4);
#pragma unroll
          for (int g = 1; g < 4; ++g) {
            const float4 q = *reinterpret_cast<const float4*>(part + (g * RC + r) * H + k4);
            s.x += q.x; s.y += q.y; s.z += q.z; s.w += q.w;
          }
          const uint32_t peer = map_to_cta(sm_base, (uint32_t)(k4 >> 5));
          st_async_v4(peer + (smem_u32(rbuf + (rank * RC + r) * 32 + (k4 & 31)) - sm_base), s, peer + (rbar(l) - sm_base));
        }
        RO_MARK(1, 7 + 6 * l)
        // ================= phase 6: gradient at the layer's input, owned units =================
        if (warp < R) {
          const int r = warp;
          mbar_wait(rbar(l), par);   // the partial sums of all CTAs for this CTA's units have landed
          RO_MARK(1, 8 + 6 * l)
          float s = dv_sm[r * 32 + lane];
#pragma unroll
          for (int c = 0; c < CL; ++c) s += rbuf[(c * RC + r) * 32 + lane];
          dx_sm[r * 32 + lane] = s;
          if (l == 0) a.dbase[(tb + r) * H + j0 + lane] = s;
        }
        // dx_sm is read next by the same thread (phase 3 of layer l-1) or after the barrier below (phase 7)
      }
      __syncthreads();
      fetch_reserve(t - 2);   // into the landing zone just consumed (read by the same threads)
      RO_MARK(1, 15)
      // ================= phase 7: W_prev^T over the owned units, all-reduced =================
      for (int d0 = (tid >> 5) * 4; d0 < R * P; d0 += (RO_THREADS / 32) * 4) {  // warp-uniform trip count
        const int d = d0 + ((tid >> 3) & 3), sub = tid & 7;
        const bool ok = d < R * P;
        const int r = ok ? ro_div(d, invP) : 0, p = ok ? d - r * P : 0;
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) s = fmaf(wprev[(sub + 8 * i) * PP + p], dx_sm[r * 32 + sub + 8 * i], s);
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        if (ok && sub == 0) pl[r * P4 + p] = s;
      }
      __syncthreads();
      for (int i = tid; i < R * (P4 / 4) * CL; i += RO_THREADS) {
        const int dst = i % CL, rq = i / CL, r = ro_div(rq, invP4q), q = rq - r * (P4 / 4);
        const float4 v4 = *reinterpret_cast<const float4*>(pl + r * P4 + 4 * q);
        const uint32_t peer = map_to_cta(sm_base, (uint32_t)dst);
        st_async_v4(peer + (smem_u32(pbuf + (rank * RC + r) * P4 + 4 * q) - sm_base), v4, peer + (pbar - sm_base));
      }
      RO_MARK(1, 16)
      mbar_wait(pbar, par);
      RO_MARK(1, 17)
      if (tid < R * P) {
        const int r = ro_div(tid, invP), p = tid - r * P;
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < CL; ++c) s += pbuf[(c * RC + r) * P4 + p];
        dpn[tid] = s;
        if (rank == 0) a.dprev[tb * P + tid] = s;
        // mask[t-1] for the next iteration: loaded after the last mat-vec so that it is never live across one
        m_cur = (a.mask && t > 0) ? a.mask[(size_t)(t - 1) * B + prow0 + r] : 0u;
      }
      __syncthreads();
      RO_MARK(1, 18)
    }
    if (warp < R) {
#pragma unroll
      for (int l = 0; l < NL; ++l) {
        a.dln_g[((size_t)l * B + prow0 + warp) * H + j0 + lane] = dgam[l];
        a.dln_b[((size_t)l * B + prow0 + warp) * H + j0 + lane] = dbet[l];
      }
    }
    __syncthreads();
  }
  cluster_sync_all();
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
template <typename K>
static int ro_max_clusters(K kernel, int CL, size_t smem) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(CL * 64));
  cfg.blockDim = dim3(RO_THREADS);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kernel, &cfg) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

template <int H, int NL>
static int ro_launch(RolloutArgs a, bool backward, cudaStream_t stream) {
  constexpr int CL = H / 32;
  const size_t smem = sizeof(float) * (size_t)(backward ? ro_bwd_layout(H, NL, a.P, a.FB).total
                                                        : ro_fwd_layout(H, NL, a.P, a.FB).total);
  if (smem > 227 * 1024) {
    set_error("mrg_rollout: %zu bytes of shared memory needed (P = %d, bottleneck = %d too large)", smem, a.P, a.FB);
    return MRG_E_UNSUPPORTED;
  }
  auto kernel = backward ? rollout_bwd_kernel<H, NL> : rollout_fwd_kernel<H, NL>;
  static size_t attr_smem[2] = {0, 0};
  static int maxc[2] = {0, 0};
  if (smem > attr_smem[backward]) {
    MRG_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem[backward] = smem;
    maxc[backward] = ro_max_clusters(kernel, CL, smem);
  }
  int clusters = maxc[backward] > 0 ? maxc[backward] : 120 / CL;
  static int forced = -1;  // MRG_ROLLOUT_CLUSTERS: tuning / tests of the multi-pass path
  if (forced < 0) {
    const char* e = getenv("MRG_ROLLOUT_CLUSTERS");
    forced = e ? atoi(e) : 0;
  }
  if (forced > 0) clusters = forced;
  if (clusters > a.B) clusters = a.B;
  a.slices = clusters;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(clusters * CL));
  cfg.blockDim = dim3(RO_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  static char name[2][64];
  if (!name[backward][0])
    snprintf(name[backward], sizeof(name[backward]), "mrg::rollout_%s_kernel<%d, %d>", backward ? "bwd" : "fwd", H, NL);
  ProfScope prof(backward ? PROF_ROLLOUT_BWD : PROF_ROLLOUT_FWD, stream, name[backward]);
  count_launch();
  MRG_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kernel, a));
  return 0;
}

template <int H>
static int ro_dispatch_layers(const RolloutArgs& a, int L, bool backward, cudaStream_t stream) {
  return L == 1 ? ro_launch<H, 1>(a, backward, stream) : ro_launch<H, 2>(a, backward, stream);
}

static int ro_dispatch(const RolloutArgs& a, int H, int L, bool backward, cudaStream_t stream) {
  switch (H) {
    case 32: return ro_dispatch_layers<32>(a, L, backward, stream);
    case 64: return ro_dispatch_layers<64>(a, L, backward, stream);
    case 128: return ro_dispatch_layers<128>(a, L, backward, stream);
    case 256: return ro_dispatch_layers<256>(a, L, backward, stream);
  }
  set_error("mrg_rollout: hidden size %d not built (32, 64, 128, 256 are)", H);
  return MRG_E_UNSUPPORTED;
}

static bool ro_shape_ok(int H, int L, int P, int FB) {
  if (!(H == 32 || H == 64 || H == 128 || H == 256) || L < 1 || L > RO_LMAX) return false;
  const int CL = H / 32;
  if (P < 1 || P > 32 || FB < 4 * CL || FB > 64 || FB % (4 * CL) != 0) return false;
  return true;
}

static int ro_fill(RolloutArgs& a, const mrg_rollout_weights* w, const mrg_rollout_reserve* rs, int T, int B) {
  a.w_prev = w->w_prev; a.w_prev_ld = w->w_prev_ld;
  for (int l = 0; l < w->L; ++l) {
    a.w_ih[l] = w->w_ih[l]; a.b_ih[l] = w->b_ih[l]; a.b_hh[l] = w->b_hh[l];
    a.ln_g[l] = w->ln_g[l]; a.ln_b[l] = w->ln_b[l];
  }
  a.w1 = w->w1; a.b1 = w->b1; a.w2 = w->w2; a.b2 = w->b2;
  a.eps = w->ln_eps; a.T = T; a.B = B; a.P = w->P; a.FB = w->FB; a.relu = w->relu;
  if (rs) {
    a.xs = rs->xs; a.gates = rs->gates; a.xhat = rs->xhat; a.rstd = rs->rstd; a.fact = rs->fact; a.prev = rs->prev;
  }
  return 0;
}

static int ro_check(const char* who, const mrg_rollout_weights* w, int T, int B) {
  MRG_REQUIRE(w != nullptr && T >= 0 && B > 0, "%s: bad arguments", who);
  MRG_REQUIRE(ro_shape_ok(w->H, w->L, w->P, w->FB),
              "%s: shape not built (H = %d in {32,64,128,256}, L = %d <= 2, P = %d <= 32, bottleneck = %d: multiple of "
              "H/8, <= 64)", who, w->H, w->L, w->P, w->FB);
  MRG_REQUIRE(w->w_prev && w->w1 && w->w2 && w->w_prev_ld >= w->P, "%s: null weights", who);
  for (int l = 0; l < w->L; ++l)
    MRG_REQUIRE(w->w_ih[l] && w->ln_g[l] && w->ln_b[l] && (((uintptr_t)w->w_ih[l]) & 15) == 0,
                "%s: layer %d: null or unaligned weights", who, l);
  MRG_REQUIRE((long long)T * B * w->H < (1LL << 31), "%s: T*B*H exceeds the 32-bit index range", who);
  return 0;
}

}  // namespace mrg

using namespace mrg;

#ifdef MRG_RO_TRACE
extern "C" int mrg_debug_rollout_trace(unsigned long long* out, int reset) {  // out[2][32] host array
  MRG_CUDA_CHECK(cudaMemcpyFromSymbol(out, mrg::g_ro_trace, sizeof(unsigned long long) * 64));
  if (reset) {
    unsigned long long z[64] = {0};
    MRG_CUDA_CHECK(cudaMemcpyToSymbol(mrg::g_ro_trace, z, sizeof(z)));
  }
  return 0;
}
#endif

extern "C" int mrg_rollout_supported(int H, int L, int P, int FB) { return ro_shape_ok(H, L, P, FB) ? 1 : 0; }

extern "C" int mrg_rollout_forward(const float* base, const float* gt_prev, const uint8_t* mask,
                                   const mrg_rollout_weights* w, float* pred, const mrg_rollout_reserve* reserve,
                                   int T, int B, void* stream) {
  if (int e = ro_check("mrg_rollout_forward", w, T, B)) return e;
  MRG_REQUIRE(base && gt_prev && pred && (((uintptr_t)base) & 15) == 0, "mrg_rollout_forward: null / unaligned tensors");
  if (T == 0) return 0;
  RolloutArgs a = {};
  ro_fill(a, w, reserve, T, B);
  a.base = base; a.gt_prev = gt_prev; a.mask = mask; a.pred = pred;
  a.train = reserve ? 1 : 0;
  if (reserve)
    MRG_REQUIRE(reserve->xs && reserve->gates && reserve->xhat && reserve->rstd && reserve->fact && reserve->prev,
                "mrg_rollout_forward: incomplete reserve");
  return ro_dispatch(a, w->H, w->L, false, (cudaStream_t)stream);
}

extern "C" int mrg_rollout_backward(const float* dpred, const uint8_t* mask, const mrg_rollout_weights* w,
                                    const mrg_rollout_reserve* reserve, const mrg_rollout_grads* g, int T, int B,
                                    void* stream) {
  if (int e = ro_check("mrg_rollout_backward", w, T, B)) return e;
  MRG_REQUIRE(dpred && reserve && g, "mrg_rollout_backward: null tensors");
  MRG_REQUIRE(reserve->gates && reserve->xhat && reserve->rstd && reserve->fact, "mrg_rollout_backward: incomplete reserve");
  MRG_REQUIRE(g->dy && g->df && g->dpre && g->dbase && g->dprev && g->dln_g && g->dln_b,
              "mrg_rollout_backward: incomplete gradient set");
  if (T == 0) return 0;
  RolloutArgs a = {};
  ro_fill(a, w, reserve, T, B);
  a.mask = mask; a.dpred = dpred;
  a.dy = g->dy; a.df = g->df; a.dpre = g->dpre; a.dbase = g->dbase; a.dprev = g->dprev;
  a.dln_g = g->dln_g; a.dln_b = g->dln_b;
  return ro_dispatch(a, w->H, w->L, true, (cudaStream_t)stream);
}
