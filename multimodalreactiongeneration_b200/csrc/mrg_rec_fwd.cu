// Persistent recurrent forward kernel (sm_100a): one thread-block cluster of 8 CTAs per slice of
// batch rows.  Batch rows are independent through the recurrence, so the synchronisation scope is the
// cluster that shares a slice, not the grid: W_hh is split by hidden unit across the 8 CTAs and stays
// in REGISTERS for the whole sequence (128 fp32 per thread at H=256), h_{t-1} of the slice lives in
// every CTA's shared memory, and each step is
//   h_{t-1} (smem) x W_hh slice (regs) -> partial sums (FFMA2) -> 16-lane shuffle reduce-scatter
//   -> + x-projection (prefetched one step ahead) -> sigmoid/tanh gates, cell update
//   -> h_t broadcast to the 8 CTAs through distributed shared memory with st.async: every remote
//      store signals the destination CTA's mbarrier (complete_tx), so the only per-step wait is a
//      warp-local try_wait on the CTA's own mbarrier — no cluster barrier, no block barrier, and the
//      global stores of y / c / gates carry no fence.
#include "mrg_common.cuh"

namespace mrg {

constexpr int CL = 8;  // CTAs per cluster

template <int H>
struct FwdCfg {
  static constexpr int U = H / CL;        // hidden units owned by one CTA
  static constexpr int UPT = U / 16;      // units per thread (16 unit groups x 16 k-slices = 256 thr)
  static constexpr int RB = 8 / UPT;      // batch rows per register chunk
  static constexpr int MK = H / 64;       // float4 k-chunks per thread (k = m*64 + ks*4 + i)
  static constexpr int NA = UPT * 4;      // gate columns per thread
  static_assert(U % 16 == 0 && (UPT == 1 || UPT == 2), "unsupported hidden size");
};

template <int H, int NCH>
__global__ void __launch_bounds__(256, 1) rec_fwd_cluster_kernel(RecArgs a, int slices) {
  using Cfg = FwdCfg<H>;
  constexpr int U = Cfg::U, UPT = Cfg::UPT, RB = Cfg::RB, MK = Cfg::MK, NA = Cfg::NA;
  constexpr int R = RB * NCH;
  __shared__ __align__(16) float h_buf[2][R][H];
  __shared__ __align__(8) unsigned long long bars[2];  // bars[b]: bytes landed in h_buf[b]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ks = lane & 15;
  const int ng = warp * 2 + (lane >> 4);
  const uint32_t rank = cluster_ctarank();
  const int cid = blockIdx.x / CL;
  const int d = cid / slices;
  const int T = a.T, B = a.B;
  const uint32_t BH = (uint32_t)B * H;
  // uneven row split: the first (B % slices) clusters take one row more
  const int sl = cid % slices, base_rows = B / slices, rem_rows = B % slices;
  const int row0 = sl * base_rows + min(sl, rem_rows);
  const int nrows = base_rows + (sl < rem_rows ? 1 : 0);  // <= R
  const int j0 = rank * U;

  const float* __restrict__ W = d == 0 ? a.w_hh[0] : a.w_hh[1];
  float* gates = a.gates + (size_t)d * T * B * 4 * H;
  float* y_ext = a.y_ext + (size_t)d * (T + 1) * B * H;
  float* c_ext = a.c_ext + (size_t)d * (T + 1) * B * H;

  // ---- W_hh slice -> registers ------------------------------------------------------------
  float4 w[NA][MK];
#pragma unroll
  for (int uu = 0; uu < UPT; ++uu)
#pragma unroll
    for (int g = 0; g < 4; ++g)
#pragma unroll
      for (int m = 0; m < MK; ++m)
        w[uu * 4 + g][m] = __ldg(reinterpret_cast<const float4*>(
            W + (size_t)(g * H + j0 + ng * UPT + uu) * H + m * 64 + ks * 4));

  // ---- initial state ----------------------------------------------------------------------
  const int init_slot = d == 0 ? 0 : T;
  for (int idx = tid; idx < R * H; idx += 256) {
    const int rl = idx / H, k = idx % H;
    const float v = (rl < nrows) ? y_ext[((size_t)init_slot * B + row0 + rl) * H + k] : 0.f;
    h_buf[0][rl][k] = v;
    h_buf[1][rl][k] = 0.f;
  }
  // owner lanes: after the reduce-scatter lane ks holds the 4 gates of combo q = (ks >> 1) & 7
  const int q = (ks >> 1) & 7;
  const int ob = UPT == 2 ? (q >> 1) : q;   // row within chunk
  const int ouu = UPT == 2 ? (q & 1) : 0;   // unit within thread
  const bool is_owner = (ks & 1) == 0;
  const int ju = ng * UPT + ouu;            // local unit
  const int j = j0 + ju;
  float c_reg[NCH];
  float4 xg[NCH];
  bool valid[NCH];
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) {
    const int row = row0 + ch * RB + ob;
    valid[ch] = is_owner && ch * RB + ob < nrows;
    c_reg[ch] = valid[ch] ? c_ext[((size_t)init_slot * B + row) * H + j] : 0.f;
    xg[ch] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  uint32_t remote[CL], remote_bar[CL];
  {
    const uint32_t base = smem_u32(&h_buf[0][0][0]);
    const uint32_t bbase = smem_u32(&bars[0]);
#pragma unroll
    for (int r = 0; r < CL; ++r) {
      remote[r] = map_to_cta(base, (uint32_t)r);
      remote_bar[r] = map_to_cta(bbase, (uint32_t)r);
    }
  }
  // bytes every CTA receives per step: one fp32 per (valid row of the slice, hidden unit)
  const uint32_t step_bytes = (uint32_t)(nrows * H * sizeof(float));
  if (tid == 0) {
    mbar_init(smem_u32(&bars[0]), 1);
    mbar_init(smem_u32(&bars[1]), 1);
    mbar_init_fence();
    if (T >= 2) mbar_arrive_expect_tx(smem_u32(&bars[1]), step_bytes);  // round of step 0
  }
  uint32_t phase0 = 0, phase1 = 0;
  if (T > 0) {
    const int t0 = d == 0 ? 0 : T - 1;
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch)
      if (valid[ch])
        xg[ch] = __ldcg(reinterpret_cast<const float4*>(
            gates + (((size_t)t0 * B + row0 + ch * RB + ob) * H + j) * 4));
  }
  __syncthreads();
  cluster_sync_all();  // every CTA of the cluster is running and has initialised its h_buf

  for (int step = 0; step < T; ++step) {
    const int t = d == 0 ? step : T - 1 - step;
    const int out_slot = d == 0 ? t + 1 : t;
    const int cur = step & 1, nxt = cur ^ 1;
    // prefetch the next step's x-projection (independent of the recurrence)
    float4 xg_n[NCH];
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      xg_n[ch] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (step + 1 < T && valid[ch]) {
        const int tn = d == 0 ? t + 1 : t - 1;
        xg_n[ch] = __ldcg(reinterpret_cast<const float4*>(gates) + (uint32_t)tn * BH +
                          (uint32_t)(row0 + ch * RB + ob) * H + j);
      }
    }
    if (step > 0) {  // h_buf[cur] holds h_{t-1} of all 8 CTAs once its mbarrier phase completes
      if (cur == 0) { mbar_wait(smem_u32(&bars[0]), phase0); phase0 ^= 1; }
      else { mbar_wait(smem_u32(&bars[1]), phase1); phase1 ^= 1; }
    }
    // re-arm this buffer's barrier for the round of step+1 (which writes h_buf[cur] again)
    if (tid == 0 && step + 2 < T) mbar_arrive_expect_tx(smem_u32(&bars[cur]), step_bytes);
    const bool send = step + 1 < T;  // nobody consumes the last step's h through smem

#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      float2 acc[NA][RB];
      if ((ch + 1) * RB <= nrows) {
        // full chunk (the common case): no per-row predicates, accumulators start from the first product
#pragma unroll
        for (int b = 0; b < RB; ++b) {
          const float* hrow = &h_buf[cur][ch * RB + b][ks * 4];
#pragma unroll
          for (int m = 0; m < MK; ++m) {
            const float4 h4 = *reinterpret_cast<const float4*>(hrow + m * 64);
            const float2 hlo = make_float2(h4.x, h4.y), hhi = make_float2(h4.z, h4.w);
#pragma unroll
            for (int n = 0; n < NA; ++n) {
              if (m == 0) acc[n][b] = fmul2(make_float2(w[n][m].x, w[n][m].y), hlo);
              else ffma2(acc[n][b], make_float2(w[n][m].x, w[n][m].y), hlo);
              ffma2(acc[n][b], make_float2(w[n][m].z, w[n][m].w), hhi);
            }
          }
        }
      } else {
#pragma unroll
        for (int n = 0; n < NA; ++n)
#pragma unroll
          for (int b = 0; b < RB; ++b) acc[n][b] = make_float2(0.f, 0.f);
#pragma unroll
        for (int b = 0; b < RB; ++b) {
          if (ch * RB + b >= nrows) continue;  // uniform: rows this cluster does not own
          const float* hrow = &h_buf[cur][ch * RB + b][ks * 4];
#pragma unroll
          for (int m = 0; m < MK; ++m) {
            const float4 h4 = *reinterpret_cast<const float4*>(hrow + m * 64);
            const float2 hlo = make_float2(h4.x, h4.y), hhi = make_float2(h4.z, h4.w);
#pragma unroll
            for (int n = 0; n < NA; ++n) {
              ffma2(acc[n][b], make_float2(w[n][m].x, w[n][m].y), hlo);
              ffma2(acc[n][b], make_float2(w[n][m].z, w[n][m].w), hhi);
            }
          }
        }
      }
      // v[q*4+g], q = b*UPT+uu
      float v32[32];
#pragma unroll
      for (int b = 0; b < RB; ++b)
#pragma unroll
        for (int uu = 0; uu < UPT; ++uu)
#pragma unroll
          for (int g = 0; g < 4; ++g)
            v32[(b * UPT + uu) * 4 + g] = acc[uu * 4 + g][b].x + acc[uu * 4 + g][b].y;
      // reduce-scatter over the 16 k-slice lanes
      float v16[16], v8[8], v4[4];
      {
        const bool up = (ks & 8) != 0;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float send = up ? v32[i] : v32[16 + i];
          const float keep = up ? v32[16 + i] : v32[i];
          v16[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
      }
      {
        const bool up = (ks & 4) != 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float send = up ? v16[i] : v16[8 + i];
          const float keep = up ? v16[8 + i] : v16[i];
          v8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
        }
      }
      {
        const bool up = (ks & 2) != 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float send = up ? v8[i] : v8[4 + i];
          const float keep = up ? v8[4 + i] : v8[i];
          v4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) v4[i] += __shfl_xor_sync(0xffffffffu, v4[i], 1);

      if (valid[ch]) {
        const int rl = ch * RB + ob;
        const int row = row0 + rl;
        const float gi = gate_sigmoid(v4[0] + xg[ch].x);
        const float gf = gate_sigmoid(v4[1] + xg[ch].y);
        const float gg = gate_tanh(v4[2] + xg[ch].z);
        const float go = gate_sigmoid(v4[3] + xg[ch].w);
        const float c = gf * c_reg[ch] + gi * gg;
        const float h = go * gate_tanh(c);
        c_reg[ch] = c;
        if (send) {
          const uint32_t off = (uint32_t)(((nxt * R + rl) * H + j) * sizeof(float));
#pragma unroll
          for (int r = 0; r < CL; ++r) st_async_f32(remote[r] + off, h, remote_bar[r] + nxt * 8);
        }
        const uint32_t oidx = (uint32_t)out_slot * BH + (uint32_t)row * H + j;  // < 2^31 (host-checked)
        y_ext[oidx] = h;
        c_ext[oidx] = c;
        if (a.train)
          reinterpret_cast<float4*>(gates)[(uint32_t)t * BH + (uint32_t)row * H + j] =
              make_float4(gi, gf, gg, go);
      }
    }
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) xg[ch] = xg_n[ch];
  }
  // Exit safety: the last round of remote stores into this CTA (step T-2) was waited for at step
  // T-1, and nobody sends at step T-1, so no store can target the shared memory of an exited CTA.
}

template <int H, int NCH>
static int launch_fwd(const RecArgs& a, int slices, cudaStream_t stream) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(a.D * slices * CL));
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  ProfScope prof(PROF_REC_FWD, stream);
  count_launch();
  MRG_CUDA_CHECK(cudaLaunchKernelEx(&cfg, rec_fwd_cluster_kernel<H, NCH>, a, slices));
  return 0;
}

template <int H>
static int max_clusters_fwd() {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(CL * 64);
  cfg.blockDim = dim3(256);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, rec_fwd_cluster_kernel<H, 1>, &cfg) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int max_active_clusters(int H) {
  static int cache256 = -1, cache128 = -1;
  if (H == 256) { if (cache256 < 0) cache256 = max_clusters_fwd<256>(); return cache256; }
  if (H == 128) { if (cache128 < 0) cache128 = max_clusters_fwd<128>(); return cache128; }
  return 0;
}

bool rec_cluster_supported(int H) { return H == 128 || H == 256; }

// Partition of the batch: `slices` clusters per direction, rows split as evenly as possible, and
// NCH register chunks of RB rows per cluster.  One wave of co-resident clusters whenever B allows.
void pick_partition(int H, int B, int D, int* slices_out, int* nch_out) {
  const int RB = H == 256 ? FwdCfg<256>::RB : FwdCfg<128>::RB;
  int maxc = max_active_clusters(H);
  if (maxc <= 0) maxc = 15;
  int per_dir = maxc / D;
  if (per_dir < 1) per_dir = 1;
  int slices = (B + RB - 1) / RB;            // one chunk per cluster if they all fit
  if (slices > per_dir) slices = per_dir;     // otherwise use every cluster slot and add chunks
  int rows = (B + slices - 1) / slices;
  int nch = (rows + RB - 1) / RB;
  if (nch > 4) {                              // more rows than 4 chunks hold: several waves
    nch = 4;
    slices = (B + 4 * RB - 1) / (4 * RB);
  }
  if (nch == 3) nch = 4;
  *slices_out = slices;
  *nch_out = nch;
}

int rec_forward_cluster(const RecArgs& a, cudaStream_t stream) {
  MRG_REQUIRE((long long)(a.T + 1) * a.B * a.H * 4 < (1LL << 31),
              "rec_forward_cluster: T*B*4H exceeds the 32-bit index range of one direction");
  int slices, nch;
  pick_partition(a.H, a.B, a.D, &slices, &nch);
  if (a.H == 256) {
    if (nch == 1) return launch_fwd<256, 1>(a, slices, stream);
    if (nch == 2) return launch_fwd<256, 2>(a, slices, stream);
    return launch_fwd<256, 4>(a, slices, stream);
  }
  if (a.H == 128) {
    if (nch == 1) return launch_fwd<128, 1>(a, slices, stream);
    if (nch == 2) return launch_fwd<128, 2>(a, slices, stream);
    return launch_fwd<128, 4>(a, slices, stream);
  }
  set_error("rec_forward_cluster: unsupported hidden size %d", a.H);
  return MRG_E_UNSUPPORTED;
}

}  // namespace mrg
