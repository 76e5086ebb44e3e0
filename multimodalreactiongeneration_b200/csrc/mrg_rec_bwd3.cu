// BPTT kernel for the REDUCED-PRECISION modes (tf32 / bf16) — the mirror of mrg_rec_fwd3.cu:
// dh_{t-1} = dpre_t W_hh on the warp-level tensor cores.
//
// rec_bwd2_kernel (mrg_rec_bwd2.cu) multiplies with FFMA2 (exact fp32 products for the 1e-4 gradient budget of the fp32
// mode); at B = 256 per GPU a cluster walks 17-18 rows per step and the FFMA2 pipe bounds the step at 4.6 us.  The
// reduced-precision modes run every other product of the backward as ONE tf32 tensor-core pass already (stated bound
// 5e-2 on gradients), so the transposed recurrent product may too.
//
// Same ownership and exchange as rec_bwd2 (cluster of 8 CTAs, CTA c owns hidden units [32c, 32c+32) and their 128 gate
// rows of W_hh; every CTA produces a partial dh over ALL 256 columns from its 128 gate rows, the owner of a column sums
// the 8 partials in fixed order; head warps / body warps meet through mbarriers only).  What changes:
//  * the OUTPUT columns are the M dimension of the MMA and the BATCH rows its N dimension (n-tile = 8 rows = one chunk), as
//    in rec_fwd3: 17-18 rows per cluster pay for 24 row slots instead of two 16-row m-tiles, and the chunks pipeline;
//  * W_hh slice as tf32 A FRAGMENTS in registers: the contraction runs over the CTA's 128 gate rows (16 k-steps), body
//    warp w owns output columns [32w, 32w+32) = two m-tiles: 16 x 2 x 4 = 128 registers per thread, the footprint of
//    the FFMA2 kernel.  Every output column of warp w belongs to CTA w: one destination per warp;
//  * B fragments (dpre) are read with ONE 16-byte load per pair of k-steps — thread q of a quad takes the four gates of
//    unit 4 kp + q as its (k = q, q + 4) elements of two consecutive k-steps, A loaded with the same permutation; row
//    stride 128 + 16 floats: conflict free.  dpre goes to the tensor core as raw fp32 bits;
//  * the m-tile rows are permuted so that a thread's four accumulator rows (two m-tiles x rows g, g + 8) are four
//    CONSECUTIVE output columns: one 16-byte st.async per batch row straight from the accumulators, no shuffle reduce;
//  * head warp hw serves row hw of every chunk.
// Used whenever the caller asks for a reduced-precision mode at H = 256 and a cluster gets <= 48 rows (six chunks of shared
// memory); measurements in mrg_rec_fwd3.cu.
#include <cstddef>
#include <cstdlib>

#include "mrg_mma_common.cuh"

namespace mrg {

constexpr int B3_THREADS = 512;   // warps 0-7: body (MMA) role, warps 8-15: head role
constexpr int B3_RB = 8;          // row capacity of a chunk = one n-tile
constexpr int B3_DP = 128 + 16;   // row stride of dpre in floats
constexpr int B3_MAX_CHUNKS = 6;

template <int H>
struct Bwd3Chunk {
  float part[2][H / 32][B3_RB][32];  // partial dh from every source CTA, double-buffered
  float dpre[B3_RB][B3_DP];          // d(pre-activation) of this CTA's 128 gate columns (unit-major, gate-minor); unused rows stay 0
  float4 g[B3_RB][32];               // prefetched gates of the step
  float4 db[B3_RB][32];              // bias-gradient accumulator (sum over t of dpre)
  float cp[B3_RB][32];               // prefetched c_{t-1} (GRU: h_{t-1})
  float dy[B3_RB][32];               // prefetched dy_t
  float dc[B3_RB][32];               // carried dc
  float c_cur[B3_RB][32];            // c_t of the step being processed
  unsigned long long hbar[2];        // bytes of partial dh landed in part[b]
  unsigned long long dbar;           // rows whose dpre is published
  unsigned long long rbar;           // body warps that have finished reading dpre
};

__device__ __forceinline__ void b3_cp16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void b3_cp8(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void b3_cp4(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void b3_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void b3_arrive_local(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void b3_wait_dyn(int n) {  // n uniform: at most n groups stay in flight
  switch (n) {
    case 0: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
    case 1: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
    case 2: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
    case 3: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
    case 4: asm volatile("cp.async.wait_group 4;" ::: "memory"); break;
    default: asm volatile("cp.async.wait_group 5;" ::: "memory"); break;
  }
}

// GRU = true: the four-slot convention of rec_bwd2_kernel (reserve in (r, z, n, q), out (dr_pre, dz_pre, dn_pre, dn_pre r)).
template <int H, bool GRU>
__global__ void __launch_bounds__(B3_THREADS, 1) rec_bwd3_kernel(RecBwdArgs a, int slices, int nch) {
  using Chunk = Bwd3Chunk<H>;
  constexpr int CL = H / 32;
  static_assert(H == 256, "one output n-tile group of 32 columns per body warp: H = 256");
  extern __shared__ __align__(16) uint8_t smem_dyn[];
  Chunk* chunks = reinterpret_cast<Chunk*>(smem_dyn);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t rank = cluster_ctarank();
  const int cid = blockIdx.x / CL;
  const int d = cid / slices;
  const int T = a.T, B = a.B, D = a.D;
  const uint32_t BH = (uint32_t)B * H;
  const int sl = cid % slices, base_rows = B / slices, rem_rows = B % slices;
  const int row0 = sl * base_rows + min(sl, rem_rows);
  const int nrows = base_rows + (sl < rem_rows ? 1 : 0);  // <= B3_RB * nch
  const int cbase = nrows / nch, crem = nrows % nch;
  const int j0 = rank * 32;

  const bool bf = a.bf16_gates != 0;
  char* gates_b = reinterpret_cast<char*>(a.gates) + (size_t)d * T * B * 4 * H * (bf ? 2 : 4);
  const float* c_ext = (GRU ? a.y_ext : a.c_ext) + (size_t)d * (T + 1) * B * H;

  // ---- shared-memory state (all 16 warps) ---------------------------------------------------------------------
  for (int ch = 0; ch < nch; ++ch) {
    Chunk& C = chunks[ch];
    const int nr = cbase + (ch < crem ? 1 : 0);
    for (int idx = tid; idx < 2 * CL * B3_RB * 32; idx += B3_THREADS) (&C.part[0][0][0][0])[idx] = 0.f;
    for (int idx = tid; idx < B3_RB * B3_DP; idx += B3_THREADS) (&C.dpre[0][0])[idx] = 0.f;
    if (tid == 0) {
      mbar_init(smem_u32(&C.hbar[0]), 1);
      mbar_init(smem_u32(&C.hbar[1]), 1);
      mbar_init(smem_u32(&C.dbar), nr > 0 ? nr : 1);
      mbar_init(smem_u32(&C.rbar), 8);
    }
  }
  if (tid == 0) {
    mbar_init_fence();
    if (T >= 1)
      for (int ch = 0; ch < nch; ++ch) {
        const int nr = cbase + (ch < crem ? 1 : 0);
        if (nr > 0)  // round of iteration 0
          mbar_arrive_expect_tx(smem_u32(&chunks[ch].hbar[1]), (uint32_t)(CL * nr * 32 * sizeof(float)));
      }
  }
  __syncthreads();

  if (warp >= 8) {
    // =========================== head warps: row hw of every chunk, lane = hidden unit ==============================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
    const int hw = warp - 8;
    const int r = hw;
    const int j = j0 + lane;
    const int ngroups = nch;  // cp.async groups committed per step by this thread
    auto prefetch = [&](Chunk& C, int crow0, int step) {
      const int t = d == 0 ? step : T - 1 - step;
      const int prev_slot = d == 0 ? t : t + 1;
      const uint32_t rj = (uint32_t)(crow0 + r) * H + j;
      if (bf) b3_cp8(smem_u32(&C.g[r][lane]), gates_b + (size_t)((uint32_t)t * BH + rj) * 8);
      else b3_cp16(smem_u32(&C.g[r][lane]), gates_b + (size_t)((uint32_t)t * BH + rj) * 16);
      b3_cp4(smem_u32(&C.cp[r][lane]), c_ext + (uint32_t)prev_slot * BH + rj);
      if (a.dy) b3_cp4(smem_u32(&C.dy[r][lane]), a.dy + ((uint32_t)t * B + crow0 + r) * (uint32_t)(D * H) + d * H + j);
    };
    for (int ch = 0; ch < nch; ++ch) {
      Chunk& C = chunks[ch];
      const int nr = cbase + (ch < crem ? 1 : 0);
      const int crow0 = row0 + ch * cbase + min(ch, crem);
      if (r < nr) {
        const size_t row = crow0 + r;
        C.db[r][lane] = make_float4(0.f, 0.f, 0.f, 0.f);
        C.dy[r][lane] = 0.f;
        C.dc[r][lane] = (!GRU && a.dc_n) ? a.dc_n[((size_t)d * B + row) * H + j] : 0.f;
        if (T > 0) {
          const int t_last = d == 0 ? T - 1 : 0;
          const int out_slot = d == 0 ? t_last + 1 : t_last;
          C.c_cur[r][lane] = GRU ? 0.f : c_ext[((size_t)out_slot * B + row) * H + j];
          prefetch(C, crow0, T - 1);
        }
        // the first iteration reads dh_n through source slot 0 of part[0]
        if (a.dh_n) C.part[0][0][r][lane] = a.dh_n[((size_t)d * B + row) * H + j];
      }
      b3_commit();
    }
    cluster_sync_all();
    uint32_t hphases = 0;  // bit (ch*2 + buf) = parity of hbar to wait for next
    for (int iter = 0; iter < T; ++iter) {
      const int step = T - 1 - iter;
      const int t = d == 0 ? step : T - 1 - step;
      const int cur = iter & 1;
      for (int ch = 0; ch < nch; ++ch) {
        Chunk& C = chunks[ch];
        const int nr = cbase + (ch < crem ? 1 : 0);
        const int crow0 = row0 + ch * cbase + min(ch, crem);
        const uint32_t hbar_cur = smem_u32(&C.hbar[cur]);
        if (r < nr) {
          // loads of this step: committed `ngroups` groups ago (one step) by this thread
          b3_wait_dyn(ngroups - 1);
          const float4 g = bf ? unpack_bf16x4(*reinterpret_cast<const uint2*>(&C.g[r][lane])) : C.g[r][lane];
          const float cprev = C.cp[r][lane];
          float dh = C.dy[r][lane];
          const float tc = GRU ? 0.f : fast_tanh(C.c_cur[r][lane]);
          const float dcin = C.dc[r][lane];
          if (iter + 1 < T) prefetch(C, crow0, step - 1);
          if (iter > 0) {  // the partial dh of all source CTAs have landed in part[cur]
            mbar_wait(hbar_cur, (hphases >> (ch * 2 + cur)) & 1u);
            hphases ^= 1u << (ch * 2 + cur);
          }
          // re-arm for the round of iteration iter+1 (which writes part[cur] again)
          if (hw == 0 && lane == 0 && iter + 1 < T)
            mbar_arrive_expect_tx(hbar_cur, (uint32_t)(CL * nr * 32 * sizeof(float)));
          // dpre may only be overwritten once all 8 local body warps have read the previous step
          if (iter > 0) mbar_wait(smem_u32(&C.rbar), (uint32_t)((iter - 1) & 1));
#pragma unroll
          for (int s = 0; s < CL; ++s) dh += C.part[cur][s][r][lane];
          float4 dp;
          float carry;
          if (GRU) {
            dh += dcin;                                            // direct path dh_{t+1} z_{t+1}
            const float dnp = dh * (1.f - g.y) * (1.f - g.z * g.z);
            const float dzp = dh * (cprev - g.z) * g.y * (1.f - g.y);   // cprev = h_{t-1}
            dp = make_float4(dnp * g.w * g.x * (1.f - g.x), dzp, dnp, dnp * g.x);
            carry = dh * g.y;
          } else {
            const float d_o = dh * tc;
            const float dct = dcin + dh * g.w * (1.f - tc * tc);
            const float d_i = dct * g.z, d_g = dct * g.x, d_f = dct * cprev;
            dp = make_float4(d_i * g.x * (1.f - g.x), d_f * g.y * (1.f - g.y), d_g * (1.f - g.z * g.z),
                             d_o * g.w * (1.f - g.w));
            carry = dct * g.y;
          }
          *reinterpret_cast<float4*>(&C.dpre[r][lane * 4]) = dp;
          __syncwarp();
          if (lane == 0) b3_arrive_local(smem_u32(&C.dbar));
          C.dc[r][lane] = carry;
          if (!GRU) C.c_cur[r][lane] = cprev;
          float4 db = C.db[r][lane];
          db.x += dp.x; db.y += dp.y; db.z += dp.z; db.w += dp.w;
          C.db[r][lane] = db;
          {
            const size_t gidx = (size_t)((uint32_t)t * BH + (uint32_t)(crow0 + r) * H + j);
            if (bf) reinterpret_cast<uint2*>(gates_b)[gidx] = pack_bf16x4(dp.x, dp.y, dp.z, dp.w);
            else reinterpret_cast<float4*>(gates_b)[gidx] = dp;
          }
        }
        b3_commit();
      }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    // ---- dh0 / dc0 / bias-gradient partials ----------------------------------------------------------------
    const int fin = T & 1;
    for (int ch = 0; ch < nch; ++ch) {
      Chunk& C = chunks[ch];
      const int nr = cbase + (ch < crem ? 1 : 0);
      const int crow0 = row0 + ch * cbase + min(ch, crem);
      if (r >= nr) continue;
      if (T > 0) mbar_wait(smem_u32(&C.hbar[fin]), (hphases >> (ch * 2 + fin)) & 1u);
      const size_t row = crow0 + r;
      float dh = 0.f;
#pragma unroll
      for (int s = 0; s < CL; ++s) dh += C.part[fin][s][r][lane];
      float* dh0 = d == 0 ? a.dh0[0] : a.dh0[1];
      float* dc0 = d == 0 ? a.dc0[0] : a.dc0[1];
      if (dh0) dh0[row * H + j] = GRU ? dh + C.dc[r][lane] : dh;
      if (dc0 && !GRU) dc0[row * H + j] = C.dc[r][lane];
      *reinterpret_cast<float4*>(a.db_part + (((size_t)d * B + row) * H + j) * 4) = C.db[r][lane];
    }
    return;
  }

  // =========================== body warps ==========================================================================
  asm volatile("setmaxnreg.inc.sync.aligned.u32 184;");
  const int g8 = lane >> 2, q = lane & 3;
  const float* __restrict__ W = d == 0 ? a.w_hh[0] : a.w_hh[1];
  // A fragments (W_hh^T) of this warp's two m-tiles.  Output column of m-tile t, fragment row g8 (+ 8): 32 warp + 4 g8 + 2 t
  // (+ 1) — the thread's four accumulator rows are the consecutive columns 32 warp + 4 g8 .. + 3.  Contraction: per pair of
  // k-steps kp the thread holds the four gate rows of local unit 4 kp + q — (gate 0, gate 1) = the (k = q, q + 4) elements of
  // the even k-step, (gate 2, gate 3) of the odd one — matching the B load below.  (a0, a1, a2, a3) = (row g8 k = q,
  // row g8 + 8 k = q, row g8 k = q + 4, row g8 + 8 k = q + 4).  tf32 rounding once, here.
  uint4 wa[8][2][2];   // [kp][even / odd k-step][m-tile]
#pragma unroll
  for (int kp = 0; kp < 8; ++kp) {
    const float* wrow = W + (size_t)(j0 + 4 * kp + q) * H + 32 * warp + 4 * g8;
    float4 gw[4];
#pragma unroll
    for (int gate = 0; gate < 4; ++gate) gw[gate] = __ldg(reinterpret_cast<const float4*>(wrow + (size_t)gate * H * H));
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      wa[kp][e][0] = make_uint4(tf32_rna(gw[2 * e].x), tf32_rna(gw[2 * e].y), tf32_rna(gw[2 * e + 1].x), tf32_rna(gw[2 * e + 1].y));
      wa[kp][e][1] = make_uint4(tf32_rna(gw[2 * e].z), tf32_rna(gw[2 * e].w), tf32_rna(gw[2 * e + 1].z), tf32_rna(gw[2 * e + 1].w));
    }
  }
  // every output column of this warp belongs to CTA `warp`
  const uint32_t remote_base = map_to_cta(smem_u32(chunks), (uint32_t)warp);
  cluster_sync_all();

  for (int iter = 0; iter < T; ++iter) {
    const int nxt = (iter & 1) ^ 1;
    for (int ch = 0; ch < nch; ++ch) {
      Chunk& C = chunks[ch];
      const int nr = cbase + (ch < crem ? 1 : 0);
      if (nr == 0) continue;
      mbar_wait(smem_u32(&C.dbar), (uint32_t)(iter & 1));  // dpre of (ch, iter) is published
      float acc[2][2][4];   // [m-tile][even / odd k-step]: four independent chains
#pragma unroll
      for (int t = 0; t < 2; ++t)
#pragma unroll
        for (int e = 0; e < 2; ++e) acc[t][e][0] = acc[t][e][1] = acc[t][e][2] = acc[t][e][3] = 0.f;
      const float* dr = &C.dpre[g8][4 * q];   // B fragment: batch row g8 (rows past the chunk's count stay zero)
      uint4 v = *reinterpret_cast<const uint4*>(dr);
#pragma unroll
      for (int kp = 0; kp < 8; ++kp) {
        const uint4 vn = *reinterpret_cast<const uint4*>(dr + (kp < 7 ? kp + 1 : 0) * 16);   // one pair of k-steps ahead
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          const uint32_t a0[4] = {wa[kp][0][t].x, wa[kp][0][t].y, wa[kp][0][t].z, wa[kp][0][t].w};
          const uint32_t a1[4] = {wa[kp][1][t].x, wa[kp][1][t].y, wa[kp][1][t].z, wa[kp][1][t].w};
          am_mma(acc[t][0], a0, v.x, v.y);
          am_mma(acc[t][1], a1, v.z, v.w);
        }
        v = vn;
      }
      __syncwarp();
      if (lane == 0) b3_arrive_local(smem_u32(&C.rbar));  // this warp is done reading dpre of (ch, iter)
      // ---- one 16-byte store per batch row to the owner of this warp's columns ------------------------------------
      // accumulator element e of m-tile t: (row g8 | g8 + 8) x (batch row 2q | 2q + 1) -> column 4 g8 + 2 t + (e >> 1)
      const uint32_t off_bar = (uint32_t)(ch * sizeof(Chunk) + offsetof(Chunk, hbar) + nxt * 8);
      const uint32_t off = (uint32_t)(ch * sizeof(Chunk) + offsetof(Chunk, part) +
                                      (((nxt * CL + (int)rank) * B3_RB + 2 * q) * 32 + 4 * g8) * sizeof(float));
      if (2 * q < nr)
        st_async_v4(remote_base + off,
                    make_float4(acc[0][0][0] + acc[0][1][0], acc[0][0][2] + acc[0][1][2], acc[1][0][0] + acc[1][1][0],
                                acc[1][0][2] + acc[1][1][2]),
                    remote_base + off_bar);
      if (2 * q + 1 < nr)
        st_async_v4(remote_base + off + 32 * sizeof(float),
                    make_float4(acc[0][0][1] + acc[0][1][1], acc[0][0][3] + acc[0][1][3], acc[1][0][1] + acc[1][1][1],
                                acc[1][0][3] + acc[1][1][3]),
                    remote_base + off_bar);
    }
  }
}

template <int H, bool GRU>
static int launch_bwd3(const RecBwdArgs& a, int slices, int nch, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    MRG_CUDA_CHECK(cudaFuncSetAttribute(rec_bwd3_kernel<H, GRU>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)(B3_MAX_CHUNKS * sizeof(Bwd3Chunk<H>))));
    attr_set = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(a.D * slices * (H / 32)));
  cfg.blockDim = dim3(B3_THREADS);
  cfg.dynamicSmemBytes = (size_t)nch * sizeof(Bwd3Chunk<H>);
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = H / 32;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  static char name[64];
  if (!name[0]) snprintf(name, sizeof(name), GRU ? "mrg::rec_bwd3_kernel<%d, gru>" : "mrg::rec_bwd3_kernel<%d>", H);
  ProfScope prof(PROF_REC_BWD, stream, name);
  count_launch();
  MRG_CUDA_CHECK(cudaLaunchKernelEx(&cfg, rec_bwd3_kernel<H, GRU>, a, slices, nch));
  return 0;
}

// Does the tensor-core BPTT apply?  Reduced-precision call, H = 256, one wave of clusters with <= 48 rows each.  The row
// partition is this kernel's own (the reserve is indexed by row: it need not match the forward's).
bool rec_backward_mma_applies(const RecBwdArgs& a, int* slices_out, int* nch_out) {
  static int off = -1;
  if (off < 0) {
    const char* e = getenv("MRG_NO_REC_MMA");   // developer switch: keep the FFMA2 recurrence in the reduced modes too
    off = (e && e[0] == '1') ? 1 : 0;
  }
  if (off || a.H != 256) return false;
  int maxc = max_active_clusters2(a.H);
  if (maxc <= 0) maxc = 15;
  if (a.cluster_budget > 0 && a.cluster_budget < maxc) maxc = a.cluster_budget;
  int per_dir = maxc / a.D;
  if (per_dir < 1) per_dir = 1;
  const int slices = a.B < per_dir ? a.B : per_dir;
  const int rows = (a.B + slices - 1) / slices;
  if (rows > B3_RB * B3_MAX_CHUNKS) return false;
  const int nch = (rows + B3_RB - 1) / B3_RB;
  // chunks of <= 8 rows (one n-tile each): the MMA work is proportional to the number of chunks, so as few as the rows need
  *slices_out = slices;
  *nch_out = nch;
  return true;
}

int rec_backward_cluster3(const RecBwdArgs& a, int slices, int nch, cudaStream_t stream) {
  MRG_REQUIRE((long long)(a.T + 1) * a.B * a.H * 4 * a.D < (1LL << 31),
              "rec_backward_cluster3: T*B*4H*D exceeds the 32-bit index range");
  return a.gru ? launch_bwd3<256, true>(a, slices, nch, stream) : launch_bwd3<256, false>(a, slices, nch, stream);
}

}  // namespace mrg
