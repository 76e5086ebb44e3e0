// Recurrent forward kernel for the REDUCED-PRECISION modes (tf32 / bf16): h_{t-1} W_hh^T on the
// warp-level tensor cores.
//
// rec_fwd2_kernel (mrg_rec_fwd2.cu) multiplies with FFMA2 because the fp32 mode has a 1e-5 parity budget per step: exact
// fp32 products, W_hh in registers as fp32.  At B = 256 per GPU (lstmformer, BASELINE configs[3]) a cluster advances 17-18
// rows per step and the FFMA2 pipe is the bound: 4.5 us per timestep, half of that configuration's step.  The
// reduced-precision modes already run every projection as ONE tf32 tensor-core pass (stated bound 2e-2 / 5e-2), so the
// recurrent product may do the same: mma.sync m16n8k8 tf32 issues 510 FMA/clk/SM (tools/mma_sync_rate.cu), 4x FFMA2.
//
// Same ownership, exchange and chunk pipeline as rec_fwd2 (cluster of 8 CTAs, CTA c owns hidden units [32c, 32c+32),
// st.async + mbarrier hand-off of h, cp.async prefetch of the x-projection, tail warps for the cell); what changes:
//  * the GATE rows are the M dimension of the MMA and the BATCH rows its N dimension (n-tile = 8 rows): a cluster with 17-18
//    rows (B = 256 on 15 clusters) pays for 24 row slots, not for two 16-row m-tiles (the first version of this kernel had
//    the batch rows on M: 1024 MMAs per CTA and step there, 768 now), and rows come in chunks of <= 8 that pipeline like the
//    FFMA2 kernel's chunks;
//  * W_hh slice as tf32 A FRAGMENTS in registers: warp w of the 8 MMA warps owns the m-tile of the 16 gate rows of units
//    4w..4w+3 (m index = local unit * 4 + gate) over all of K = H: 32 k-steps x 4 registers = 128 per thread — the register
//    footprint of the FFMA2 kernel;
//  * B fragments (h) are read from shared memory with ONE 16-byte load per pair of k-steps: the contraction index is
//    permuted (thread q of a quad takes memory columns 4q .. 4q+3 of a 16-column group as its (k = q, q + 4) elements of
//    two consecutive k-steps; the A fragments are loaded with the same permutation, so the sum is unchanged), row stride
//    H + 16 floats makes those loads conflict free; rows past the chunk's count stay zero; h goes to the tensor core as raw
//    fp32 bits (it ignores the low 13 mantissa bits);
//  * every MMA warp produces COMPLETE gate sums for its 16 gate rows (no k-split, no partial-sum reduction in the tail:
//    the tail reads one float4 per (row, unit)); the accumulator (gate row g / g + 8 x batch rows 2q, 2q + 1) is stored
//    with a row stride of 132 floats: conflict free;
//  * tail warp tw serves row tw of every chunk.
// Used whenever the caller asks for a reduced-precision mode at H = 256 and a cluster gets <= 64 rows: measured faster than
// the FFMA2 kernel at every batch size (tools/rec_bench.py REDUCED=1, us per timestep forward / BPTT: B = 16 0.83 / 0.74 vs
// 0.90 / 0.92, B = 64 0.94 / 0.83 vs 1.49 / 1.55, B = 128 1.32 / 1.16 vs 2.43 / 2.57, B = 256 1.98 / 1.75 vs 4.51 / 4.64).
#include <cstddef>
#include <cstdlib>

#include "mrg_mma_common.cuh"

namespace mrg {

constexpr int F3_THREADS = 512;   // warps 0-7: MMA role, warps 8-15: tail role
constexpr int F3_RB = 8;          // row capacity of a chunk = one n-tile
constexpr int F3_PLD = 132;       // floats per row of the gate-sum buffer (128 + 4: conflict-free accumulator stores)

template <int H>
struct Fwd3Chunk {
  float h[2][F3_RB][H + 16];    // h_{t-1} of the chunk's rows, double-buffered (written by all CTAs); unused rows stay 0
  float part[F3_RB][F3_PLD];    // complete gate sums [row][unit * 4 + gate]
  float4 xg[F3_RB][32];         // x-projection of the step being computed
  float c[F3_RB][32];           // cell state
  unsigned long long hbar[2];   // bytes of h landed in h[b]
  unsigned long long pbar;      // MMA warps whose sums are stored
  unsigned long long pad;
};

__device__ __forceinline__ void f3_cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void f3_wait_dyn(int n) {  // n uniform: at most n groups stay in flight
  switch (n) {
    case 0: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
    case 1: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
    case 2: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
    case 3: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
    case 4: asm volatile("cp.async.wait_group 4;" ::: "memory"); break;
    case 5: asm volatile("cp.async.wait_group 5;" ::: "memory"); break;
    case 6: asm volatile("cp.async.wait_group 6;" ::: "memory"); break;
    default: asm volatile("cp.async.wait_group 7;" ::: "memory"); break;
  }
}

template <int H, bool GRU>
__global__ void __launch_bounds__(F3_THREADS, 1) rec_fwd3_kernel(RecArgs a, int slices, int nch) {
  using Chunk = Fwd3Chunk<H>;
  constexpr int CL = H / 32, HP = H + 16, KP = H / 16;   // KP = pairs of k-steps
  extern __shared__ __align__(16) uint8_t smem_dyn[];
  Chunk* chunks = reinterpret_cast<Chunk*>(smem_dyn);

  REC_TRACE_DECL
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t rank = cluster_ctarank();
  const int cid = blockIdx.x / CL;
  const int d = cid / slices;
  const int T = a.T, B = a.B;
  const uint32_t BH = (uint32_t)B * H;
  // uneven row split over the clusters, then over the chunks of a cluster (as rec_fwd2)
  const int sl = cid % slices, base_rows = B / slices, rem_rows = B % slices;
  const int row0 = sl * base_rows + min(sl, rem_rows);
  const int nrows = base_rows + (sl < rem_rows ? 1 : 0);  // <= F3_RB * nch
  const int cbase = nrows / nch, crem = nrows % nch;
  const int j0 = rank * 32;

  const bool bf = a.bf16_gates != 0;
  char* gates_b = reinterpret_cast<char*>(a.gates) + (size_t)d * T * B * 4 * H * (bf ? 2 : 4);
  float* y_ext = a.y_ext + (size_t)d * (T + 1) * B * H;
  float* c_ext = a.c_ext + (size_t)d * (T + 1) * B * H;
  auto fetch_xg = [&](float4* dst, uint32_t idx) {   // x-projection of (t, row, unit) -> shared memory
    if (bf) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(gates_b + (size_t)idx * 8) : "memory");
    else f3_cp_async16(smem_u32(dst), gates_b + (size_t)idx * 16);
  };

  // ---- initial state (all 16 warps) ---------------------------------------------------------
  const int init_slot = d == 0 ? 0 : T;
  for (int ch = 0; ch < nch; ++ch) {
    Chunk& C = chunks[ch];
    const int nr = cbase + (ch < crem ? 1 : 0);
    const int crow0 = row0 + ch * cbase + min(ch, crem);
    for (int idx = tid; idx < F3_RB * HP; idx += F3_THREADS) {
      const int rl = idx / HP, k = idx % HP;
      C.h[0][rl][k] = (rl < nr && k < H) ? y_ext[((size_t)init_slot * B + crow0 + rl) * H + k] : 0.f;
      C.h[1][rl][k] = 0.f;
    }
    if (tid == 0) {
      mbar_init(smem_u32(&C.hbar[0]), 1);
      mbar_init(smem_u32(&C.hbar[1]), 1);
      mbar_init(smem_u32(&C.pbar), 8);
    }
  }
  if (tid == 0) {
    mbar_init_fence();
    if (T >= 2)
      for (int ch = 0; ch < nch; ++ch) {
        const int nr = cbase + (ch < crem ? 1 : 0);
        if (nr > 0)  // round of step 0
          mbar_arrive_expect_tx(smem_u32(&chunks[ch].hbar[1]), (uint32_t)(nr * H * sizeof(float)));
      }
  }
  __syncthreads();
  cluster_sync_all();  // every CTA of the cluster is running and has initialised its buffers and barriers

  if (warp >= 8) {
    // =========================== tail warps: row tw of every chunk, lane = hidden unit ================================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
    const int tw = warp - 8;
    const int j = j0 + lane;
    uint32_t remote_base[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int dst = (lane & 3) + 4 * i;
      remote_base[i] = map_to_cta(smem_u32(chunks), (uint32_t)(dst < CL ? dst : 0));
    }
    const int t0 = d == 0 ? 0 : T - 1;
    const int r = tw;
    // one cp.async group per chunk, in the order they are consumed
    for (int ch = 0; ch < nch; ++ch) {
      Chunk& C = chunks[ch];
      const int nr = cbase + (ch < crem ? 1 : 0);
      const int crow0 = row0 + ch * cbase + min(ch, crem);
      if (r < nr) {
        C.c[r][lane] = GRU ? 0.f : c_ext[((size_t)init_slot * B + crow0 + r) * H + j];
        if (T > 0) fetch_xg(&C.xg[r][lane], (uint32_t)t0 * BH + (uint32_t)(crow0 + r) * H + j);
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    const int ngroups = nch;   // groups committed per step by this thread
    const int tstep = d == 0 ? 1 : -1;
    for (int step = 0; step < T; ++step) {
      const int t = d == 0 ? step : T - 1 - step;
      const uint32_t cur = (uint32_t)(step & 1), nxt = cur ^ 1u;
      const bool send = step + 1 < T;  // nobody consumes the last step's h through shared memory
      const uint32_t obase = (uint32_t)(d == 0 ? t + 1 : t) * BH;  // < 2^31 (host-checked)
      for (int ch = 0; ch < nch; ++ch) {
        Chunk& C = chunks[ch];
        const int nr = cbase + (ch < crem ? 1 : 0);
        const int crow0 = row0 + ch * cbase + min(ch, crem);
        if (r < nr) {
          // x-projection of this step: committed `ngroups` groups ago (one step) by this thread
          f3_wait_dyn(ngroups - 1);
          const float4 xg = bf ? unpack_bf16x4(*reinterpret_cast<const uint2*>(&C.xg[r][lane])) : C.xg[r][lane];
          const float cold = C.c[r][lane];
          const uint32_t rj = (uint32_t)(crow0 + r) * H + j;
          if (send) fetch_xg(&C.xg[r][lane], (uint32_t)(t + tstep) * BH + rj);
          REC_TRACE(10, ch, step);
          mbar_wait(smem_u32(&C.pbar), cur);
          REC_TRACE(11, ch, step);  // all 8 MMA warps have stored their sums
          const float4 p = *reinterpret_cast<const float4*>(&C.part[r][lane * 4]);
          float gi, gf, gg, go, cn, h;
          if (GRU) {
            gi = fast_sigmoid(p.x + xg.x);                 // r
            gf = fast_sigmoid(p.y + xg.y);                 // z
            go = p.w + xg.w;                               // W_hn h + b_hn
            gg = fast_tanh(fmaf(gi, go, xg.z));            // n
            const float hprev = C.h[cur][r][j0 + lane];
            cn = 0.f;
            h = fmaf(gf, hprev - gg, gg);
          } else {
            gi = fast_sigmoid(p.x + xg.x);
            gf = fast_sigmoid(p.y + xg.y);
            gg = fast_tanh(p.z + xg.z);
            go = fast_sigmoid(p.w + xg.w);
            cn = fmaf(gf, cold, gi * gg);
            h = go * fast_tanh(cn);
          }
          if (send) {
            float4 hv;  // h of units 4q .. 4q+3, q = lane / 4
            hv.x = __shfl_sync(0xffffffffu, h, (lane & ~3));
            hv.y = __shfl_sync(0xffffffffu, h, (lane & ~3) + 1);
            hv.z = __shfl_sync(0xffffffffu, h, (lane & ~3) + 2);
            hv.w = __shfl_sync(0xffffffffu, h, (lane & ~3) + 3);
            const uint32_t off_h = (uint32_t)(ch * sizeof(Chunk) + offsetof(Chunk, h) +
                                              ((nxt * F3_RB + r) * HP + j0 + (lane & ~3)) * sizeof(float));
            const uint32_t off_bar = (uint32_t)(ch * sizeof(Chunk) + offsetof(Chunk, hbar) + nxt * 8);
#pragma unroll
            for (int i = 0; i < 2; ++i)
              if ((lane & 3) + 4 * i < CL) st_async_v4(remote_base[i] + off_h, hv, remote_base[i] + off_bar);
          }
          REC_TRACE(12, ch, step);
          y_ext[obase + rj] = h;
          if (!GRU) {
            C.c[r][lane] = cn;
            c_ext[obase + rj] = cn;
          }
          if (a.train) {
            const size_t gidx = (size_t)((uint32_t)t * BH + rj);
            if (bf) reinterpret_cast<uint2*>(gates_b)[gidx] = pack_bf16x4(gi, gf, gg, go);
            else reinterpret_cast<float4*>(gates_b)[gidx] = make_float4(gi, gf, gg, go);
          }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
      }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    return;
  }

  // =========================== MMA warps ==============================================================================
  asm volatile("setmaxnreg.inc.sync.aligned.u32 184;");
  const int g8 = lane >> 2, q = lane & 3;
  const float* __restrict__ W = d == 0 ? a.w_hh[0] : a.w_hh[1];
  // A fragments of this warp's m-tile: m = local unit * 4 + gate, fragment rows g8 / g8 + 8 -> units 4 warp + (g8 >> 2) and
  // 4 warp + 2 + (g8 >> 2), gate g8 & 3.  Per pair of k-steps kp the thread holds memory columns 16 kp + 4q .. + 3 of both
  // rows: (a0, a1, a2, a3) of the even k-step = (row g8 col 0, row g8+8 col 0, row g8 col 1, row g8+8 col 1), of the odd
  // k-step the same with columns 2, 3.  tf32 rounding once, here.
  uint4 wa[KP][2];
  {
    const int gate = g8 & 3;
    const float* wr0 = W + (size_t)(gate * H + j0 + warp * 4 + (g8 >> 2)) * H;
    const float* wr1 = wr0 + (size_t)2 * H;
#pragma unroll
    for (int kp = 0; kp < KP; ++kp) {
      const float4 lo = __ldg(reinterpret_cast<const float4*>(wr0 + kp * 16 + 4 * q));
      const float4 hi = __ldg(reinterpret_cast<const float4*>(wr1 + kp * 16 + 4 * q));
      wa[kp][0] = make_uint4(tf32_rna(lo.x), tf32_rna(hi.x), tf32_rna(lo.y), tf32_rna(hi.y));
      wa[kp][1] = make_uint4(tf32_rna(lo.z), tf32_rna(hi.z), tf32_rna(lo.w), tf32_rna(hi.w));
    }
  }
  uint32_t hphases = 0;  // bit (ch*2 + buf): parity of hbar to wait for next

  for (int step = 0; step < T; ++step) {
    const int cur = step & 1;
    for (int ch = 0; ch < nch; ++ch) {
      const int nr = cbase + (ch < crem ? 1 : 0);
      if (nr == 0) continue;
      Chunk& C = chunks[ch];
      const uint32_t hbar_cur = smem_u32(&C.hbar[cur]);
      REC_TRACE(1, ch, step);
      if (step > 0) {  // h_{t-1} of this chunk from all CTAs has landed in C.h[cur]
        mbar_wait(hbar_cur, (hphases >> (ch * 2 + cur)) & 1u);
        hphases ^= 1u << (ch * 2 + cur);
      }
      REC_TRACE(2, ch, step);
      // re-arm this buffer's barrier for the round of step+1 (which writes C.h[cur] again)
      if (tid == 0 && step + 2 < T) mbar_arrive_expect_tx(hbar_cur, (uint32_t)(nr * H * sizeof(float)));
      float acc[4][4];   // four independent accumulator chains (k-step mod 4), summed at the end
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[c][0] = acc[c][1] = acc[c][2] = acc[c][3] = 0.f;
      const float* hr = &C.h[cur][g8][4 * q];   // B fragment: batch row g8 (rows past the chunk's count are zero)
#pragma unroll
      for (int kp = 0; kp < KP; ++kp) {
        const uint4 v = *reinterpret_cast<const uint4*>(hr + kp * 16);
        const uint32_t a0[4] = {wa[kp][0].x, wa[kp][0].y, wa[kp][0].z, wa[kp][0].w};
        const uint32_t a1[4] = {wa[kp][1].x, wa[kp][1].y, wa[kp][1].z, wa[kp][1].w};
        am_mma(acc[(2 * kp) & 3], a0, v.x, v.y);
        am_mma(acc[(2 * kp + 1) & 3], a1, v.z, v.w);
      }
      // accumulator: gate rows m = g8 / g8 + 8 of this warp's m-tile x batch rows 2q, 2q + 1 -> part[row][16 warp + m]
      {
        const float s0 = (acc[0][0] + acc[1][0]) + (acc[2][0] + acc[3][0]);
        const float s1 = (acc[0][1] + acc[1][1]) + (acc[2][1] + acc[3][1]);
        const float s2 = (acc[0][2] + acc[1][2]) + (acc[2][2] + acc[3][2]);
        const float s3 = (acc[0][3] + acc[1][3]) + (acc[2][3] + acc[3][3]);
        float* pr = &C.part[2 * q][16 * warp + g8];
        pr[0] = s0;
        pr[F3_PLD] = s1;
        pr[8] = s2;
        pr[F3_PLD + 8] = s3;
      }
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&C.pbar)) : "memory");
      REC_TRACE(3, ch, step);
    }
  }
  // Exit safety as in rec_fwd2: the last round of remote stores into this CTA (step T-2) was waited for at step T-1.
}

constexpr int F3_MAX_CHUNKS = 8;

template <int H, bool GRU>
static int launch_fwd3(const RecArgs& a, int slices, int nch, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    MRG_CUDA_CHECK(cudaFuncSetAttribute(rec_fwd3_kernel<H, GRU>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)(F3_MAX_CHUNKS * sizeof(Fwd3Chunk<H>))));
    attr_set = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(a.D * slices * (H / 32)));
  cfg.blockDim = dim3(F3_THREADS);
  cfg.dynamicSmemBytes = (size_t)nch * sizeof(Fwd3Chunk<H>);
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = H / 32;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  static char name[64];
  if (!name[0]) snprintf(name, sizeof(name), GRU ? "mrg::rec_fwd3_kernel<%d, gru>" : "mrg::rec_fwd3_kernel<%d>", H);
  ProfScope prof(PROF_REC_FWD, stream, name);
  count_launch();
  MRG_CUDA_CHECK(cudaLaunchKernelEx(&cfg, rec_fwd3_kernel<H, GRU>, a, slices, nch));
  return 0;
}

// Does the tensor-core forward apply?  Reduced-precision call, H = 256, one wave of clusters with <= 64 rows each (chunks of <= 8 rows).
bool rec_forward_mma_applies(const RecArgs& a, int* slices_out, int* nch_out) {
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
  if (rows > F3_RB * F3_MAX_CHUNKS) return false;
  const int nch = (rows + F3_RB - 1) / F3_RB;
  // chunks of <= 8 rows (one n-tile each): the MMA work is proportional to the number of chunks, so as few as the rows need
  *slices_out = slices;
  *nch_out = nch;
  return true;
}

int rec_forward_cluster3(const RecArgs& a, int slices, int nch, cudaStream_t stream) {
  MRG_REQUIRE((long long)(a.T + 1) * a.B * a.H * 4 < (1LL << 31),
              "rec_forward_cluster3: T*B*4H exceeds the 32-bit index range of one direction");
  return a.gru ? launch_fwd3<256, true>(a, slices, nch, stream) : launch_fwd3<256, false>(a, slices, nch, stream);
}

}  // namespace mrg
