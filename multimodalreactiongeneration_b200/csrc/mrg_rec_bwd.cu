// Persistent BPTT kernel (sm_100a), same partition as the forward: a cluster of 8 CTAs per slice of
// batch rows, W_hh slice [4U gate rows of this CTA's U hidden units] x [H] resident in registers.
// Per step (descending in processing order):
//   dh_t = dy_t + sum over the 8 source CTAs of their partial dh (slots in local smem, fixed order
//          -> deterministic) ; gate derivatives -> dpre (written over the gates reserve + local smem)
//   partial dh_{t-1}[b][k] = sum over this CTA's 4U gate columns of dpre * W_hh   (FFMA2, W in regs)
//   8-lane shuffle reduce-scatter, one 16-byte st.async per lane to the CTA that owns unit k; the
//   store signals that CTA's mbarrier (complete_tx), so the per-step wait is a try_wait on the local
//   mbarrier instead of a cluster barrier.
// The bias gradient (sum of dpre over t) is accumulated in registers and written per batch row.
#include "mrg_common.cuh"

namespace mrg {

constexpr int CLB = 8;

template <int H>
struct BwdCfg {
  static constexpr int U = H / CLB;      // hidden units per CTA
  static constexpr int NL = 4 * U;       // local gate columns (reduction dim of the matvec)
  static constexpr int TNR = NL / 8;     // gate columns per thread (8 n-slices)
  static constexpr int MM = TNR / 4;     // float4 chunks of dpre per thread per row
  static constexpr int TKO = H / 32;     // output k per thread (32 k-groups)
  static constexpr int RB = 32 / TKO;    // rows per register chunk
  static_assert(TKO == 4 || TKO == 8, "unsupported hidden size");
};

template <int H, int NCH>
__global__ void __launch_bounds__(256, 1) rec_bwd_cluster_kernel(RecBwdArgs a, int slices) {
  using Cfg = BwdCfg<H>;
  constexpr int U = Cfg::U, NL = Cfg::NL, MM = Cfg::MM, TKO = Cfg::TKO, RB = Cfg::RB;
  constexpr int R = RB * NCH;
  constexpr int NITEMS = R * U;                       // (row, unit) pairs owned by this CTA
  constexpr int NIT = (NITEMS + 255) / 256;           // per thread
  __shared__ __align__(16) float part_buf[2][CLB][R][U];
  __shared__ __align__(16) float dpre_s[R][NL];
  __shared__ __align__(8) unsigned long long bars[2];  // bars[b]: bytes landed in part_buf[b]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ns = lane & 7;
  const int kg = warp * 4 + (lane >> 3);
  const uint32_t rank = cluster_ctarank();
  const int cid = blockIdx.x / CLB;
  const int d = cid / slices;
  const int T = a.T, B = a.B, D = a.D;
  const uint32_t BH = (uint32_t)B * H;
  const int sl = cid % slices, base_rows = B / slices, rem_rows = B % slices;
  const int row0 = sl * base_rows + min(sl, rem_rows);
  const int nrows = base_rows + (sl < rem_rows ? 1 : 0);  // <= R, same split as the forward
  const int j0 = rank * U;

  const float* __restrict__ W = d == 0 ? a.w_hh[0] : a.w_hh[1];
  float* gates = a.gates + (size_t)d * T * B * 4 * H;
  const float* c_ext = a.c_ext + (size_t)d * (T + 1) * B * H;

  // ---- W_hh slice -> registers: w2[mm][i][kp] = (W[row][k], W[row][k+1]) ---------------------
  float2 w2[MM][4][TKO / 2];
#pragma unroll
  for (int mm = 0; mm < MM; ++mm)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float* src = W + (size_t)(i * H + j0 + mm * 8 + ns) * H + kg * TKO;
#pragma unroll
      for (int k4 = 0; k4 < TKO / 4; ++k4) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(src + k4 * 4));
        w2[mm][i][k4 * 2 + 0] = make_float2(v.x, v.y);
        w2[mm][i][k4 * 2 + 1] = make_float2(v.z, v.w);
      }
    }

  // ---- per-item state -----------------------------------------------------------------------
  float dc_reg[NIT], c_cur[NIT];
  float4 dbacc[NIT];
  bool valid[NIT];
  int it_rl[NIT], it_u[NIT];
#pragma unroll
  for (int i = 0; i < NIT; ++i) {
    const int it = tid + i * 256;
    it_rl[i] = it / U;
    it_u[i] = it % U;
    valid[i] = it < NITEMS && it_rl[i] < nrows;
    dbacc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    dc_reg[i] = 0.f;
    c_cur[i] = 0.f;
    if (valid[i]) {
      const size_t row = row0 + it_rl[i];
      const int j = j0 + it_u[i];
      if (a.dc_n) dc_reg[i] = a.dc_n[((size_t)d * B + row) * H + j];
      if (T > 0) {
        const int t_last = d == 0 ? T - 1 : 0;
        const int out_slot = d == 0 ? t_last + 1 : t_last;
        c_cur[i] = c_ext[((size_t)out_slot * B + row) * H + j];
      }
    }
  }
  for (int idx = tid; idx < 2 * CLB * R * U; idx += 256) (&part_buf[0][0][0][0])[idx] = 0.f;
  for (int idx = tid; idx < R * NL; idx += 256) (&dpre_s[0][0])[idx] = 0.f;
  __syncthreads();
  // first iteration reads dh_n through slot src=0 of part_buf[0]
#pragma unroll
  for (int i = 0; i < NIT; ++i)
    if (valid[i] && a.dh_n)
      part_buf[0][0][it_rl[i]][it_u[i]] = a.dh_n[((size_t)d * B + row0 + it_rl[i]) * H + j0 + it_u[i]];

  // destination of this lane's reduced float4: owner CTA of k and the offset inside its part_buf
  const int khalf = TKO == 8 ? (ns & 1) : 0;
  const int ob = TKO == 8 ? (ns >> 1) : ns;             // row within chunk held after the reduce
  const int kfirst = kg * TKO + khalf * 4;
  const uint32_t owner = (uint32_t)(kfirst / U);
  const int kin = kfirst % U;
  const uint32_t remote_base = map_to_cta(smem_u32(&part_buf[0][0][0][0]), owner);
  const uint32_t remote_bar = map_to_cta(smem_u32(&bars[0]), owner);
  // every lane of every CTA sends one float4 per chunk per step: R*U fp32 from each of the 8 sources
  constexpr uint32_t STEP_BYTES = (uint32_t)(CLB * R * U * sizeof(float));
  if (tid == 0) {
    mbar_init(smem_u32(&bars[0]), 1);
    mbar_init(smem_u32(&bars[1]), 1);
    mbar_init_fence();
    if (T >= 1) mbar_arrive_expect_tx(smem_u32(&bars[1]), STEP_BYTES);  // round of iteration 0
  }
  uint32_t phase0 = 0, phase1 = 0;

  // prefetch registers for the first processed step
  float4 g4[NIT];
  float cp[NIT], dyv[NIT];
  auto prefetch = [&](int step) {
    const int t = d == 0 ? step : T - 1 - step;
    const int prev_slot = d == 0 ? t : t + 1;
#pragma unroll
    for (int i = 0; i < NIT; ++i) {
      g4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      cp[i] = 0.f;
      dyv[i] = 0.f;
      if (valid[i]) {
        const uint32_t rj = (uint32_t)(row0 + it_rl[i]) * H + j0 + it_u[i];
        g4[i] = __ldcg(reinterpret_cast<const float4*>(gates) + (uint32_t)t * BH + rj);
        cp[i] = c_ext[(uint32_t)prev_slot * BH + rj];
        if (a.dy) dyv[i] = a.dy[((uint32_t)t * B + row0 + it_rl[i]) * (uint32_t)(D * H) + d * H + j0 + it_u[i]];
      }
    }
  };
  if (T > 0) prefetch(T - 1);
  __syncthreads();
  cluster_sync_all();

  for (int iter = 0; iter < T; ++iter) {
    const int step = T - 1 - iter;
    const int t = d == 0 ? step : T - 1 - step;
    const int cur = iter & 1, nxt = cur ^ 1;
    if (iter > 0) {
      if (cur == 0) { mbar_wait(smem_u32(&bars[0]), phase0); phase0 ^= 1; }
      else { mbar_wait(smem_u32(&bars[1]), phase1); phase1 ^= 1; }
    }
    if (tid == 0 && iter + 1 < T) mbar_arrive_expect_tx(smem_u32(&bars[cur]), STEP_BYTES);

    // ---- elementwise: dh, gate derivatives -> dpre ------------------------------------------
#pragma unroll
    for (int i = 0; i < NIT; ++i) {
      if (valid[i]) {
        const int rl = it_rl[i], u = it_u[i];
        float dh = dyv[i];
#pragma unroll
        for (int s = 0; s < CLB; ++s) dh += part_buf[cur][s][rl][u];
        const float4 g = g4[i];
        const float tc = gate_tanh(c_cur[i]);
        const float d_o = dh * tc;
        const float dct = dc_reg[i] + dh * g.w * (1.f - tc * tc);
        const float d_i = dct * g.z, d_g = dct * g.x, d_f = dct * cp[i];
        dc_reg[i] = dct * g.y;
        c_cur[i] = cp[i];
        const float4 dp = make_float4(d_i * g.x * (1.f - g.x), d_f * g.y * (1.f - g.y),
                                      d_g * (1.f - g.z * g.z), d_o * g.w * (1.f - g.w));
        dbacc[i].x += dp.x; dbacc[i].y += dp.y; dbacc[i].z += dp.z; dbacc[i].w += dp.w;
        *reinterpret_cast<float4*>(&dpre_s[rl][u * 4]) = dp;
        reinterpret_cast<float4*>(gates)[(uint32_t)t * BH + (uint32_t)(row0 + rl) * H + j0 + u] = dp;
      }
    }
    if (iter + 1 < T) prefetch(step - 1);
    __syncthreads();

    // ---- partial dh_{prev}[b][k] over this CTA's gate columns --------------------------------
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      float2 acc[TKO / 2][RB];
      if ((ch + 1) * RB <= nrows) {
        // full chunk: no per-row predicates, accumulators start from the first product
#pragma unroll
        for (int b = 0; b < RB; ++b) {
#pragma unroll
          for (int mm = 0; mm < MM; ++mm) {
            const float4 dv = *reinterpret_cast<const float4*>(&dpre_s[ch * RB + b][mm * 32 + ns * 4]);
            const float dvv[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float2 dup = make_float2(dvv[i], dvv[i]);
#pragma unroll
              for (int kp = 0; kp < TKO / 2; ++kp) {
                if (mm == 0 && i == 0) acc[kp][b] = fmul2(dup, w2[mm][i][kp]);
                else ffma2(acc[kp][b], dup, w2[mm][i][kp]);
              }
            }
          }
        }
      } else {
#pragma unroll
        for (int kp = 0; kp < TKO / 2; ++kp)
#pragma unroll
          for (int b = 0; b < RB; ++b) acc[kp][b] = make_float2(0.f, 0.f);
#pragma unroll
        for (int b = 0; b < RB; ++b) {
          if (ch * RB + b >= nrows) continue;  // uniform: rows this cluster does not own
#pragma unroll
          for (int mm = 0; mm < MM; ++mm) {
            const float4 dv = *reinterpret_cast<const float4*>(&dpre_s[ch * RB + b][mm * 32 + ns * 4]);
            const float dvv[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float2 dup = make_float2(dvv[i], dvv[i]);
#pragma unroll
              for (int kp = 0; kp < TKO / 2; ++kp) ffma2(acc[kp][b], dup, w2[mm][i][kp]);
            }
          }
        }
      }
      // v[q*4 + (kk&3)], q = b*(TKO/4) + (kk>>2)
      float v32[32];
#pragma unroll
      for (int b = 0; b < RB; ++b)
#pragma unroll
        for (int kp = 0; kp < TKO / 2; ++kp) {
          const int kk = kp * 2;
          const int qq = b * (TKO / 4) + (kk >> 2);
          v32[qq * 4 + (kk & 3)] = acc[kp][b].x;
          v32[qq * 4 + (kk & 3) + 1] = acc[kp][b].y;
        }
      float v16[16], v8[8], v4[4];
      {
        const bool up = (ns & 4) != 0;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float send = up ? v32[i] : v32[16 + i];
          const float keep = up ? v32[16 + i] : v32[i];
          v16[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
        }
      }
      {
        const bool up = (ns & 2) != 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float send = up ? v16[i] : v16[8 + i];
          const float keep = up ? v16[8 + i] : v16[i];
          v8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
        }
      }
      {
        const bool up = (ns & 1) != 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float send = up ? v8[i] : v8[4 + i];
          const float keep = up ? v8[4 + i] : v8[i];
          v4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
        }
      }
      const int rl = ch * RB + ob;
      const uint32_t off = (uint32_t)((((nxt * CLB + (int)rank) * R + rl) * U + kin) * sizeof(float));
      st_async_v4(remote_base + off, make_float4(v4[0], v4[1], v4[2], v4[3]), remote_bar + nxt * 8);
    }
  }
  if (T > 0) {  // the last round (iteration T-1) carries dh0
    if ((T & 1) == 0) mbar_wait(smem_u32(&bars[0]), phase0);
    else mbar_wait(smem_u32(&bars[1]), phase1);
  }

  // ---- dh0 / dc0 / bias-gradient partials -----------------------------------------------------
  const int fin = T & 1;
#pragma unroll
  for (int i = 0; i < NIT; ++i) {
    if (valid[i]) {
      const int rl = it_rl[i], u = it_u[i];
      const size_t row = row0 + rl;
      const int j = j0 + u;
      float dh = 0.f;
#pragma unroll
      for (int s = 0; s < CLB; ++s) dh += part_buf[fin][s][rl][u];
      float* dh0 = d == 0 ? a.dh0[0] : a.dh0[1];
      float* dc0 = d == 0 ? a.dc0[0] : a.dc0[1];
      if (dh0) dh0[row * H + j] = dh;
      if (dc0) dc0[row * H + j] = dc_reg[i];
      *reinterpret_cast<float4*>(a.db_part + (((size_t)d * B + row) * H + j) * 4) = dbacc[i];
    }
  }
}

template <int H, int NCH>
static int launch_bwd(const RecBwdArgs& a, int slices, cudaStream_t stream) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(a.D * slices * CLB));
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CLB;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  ProfScope prof(PROF_REC_BWD, stream);
  count_launch();
  MRG_CUDA_CHECK(cudaLaunchKernelEx(&cfg, rec_bwd_cluster_kernel<H, NCH>, a, slices));
  return 0;
}

void pick_partition(int H, int B, int D, int* slices_out, int* nch_out);

int rec_backward_cluster(const RecBwdArgs& a, cudaStream_t stream) {
  MRG_REQUIRE((long long)(a.T + 1) * a.B * a.H * 4 * a.D < (1LL << 31),
              "rec_backward_cluster: T*B*4H*D exceeds the 32-bit index range");
  int slices, nch;
  pick_partition(a.H, a.B, a.D, &slices, &nch);
  if (a.H == 256) {
    if (nch == 1) return launch_bwd<256, 1>(a, slices, stream);
    if (nch == 2) return launch_bwd<256, 2>(a, slices, stream);
    return launch_bwd<256, 4>(a, slices, stream);
  }
  if (a.H == 128) {
    if (nch == 1) return launch_bwd<128, 1>(a, slices, stream);
    if (nch == 2) return launch_bwd<128, 2>(a, slices, stream);
    return launch_bwd<128, 4>(a, slices, stream);
  }
  set_error("rec_backward_cluster: unsupported hidden size %d", a.H);
  return MRG_E_UNSUPPORTED;
}

}  // namespace mrg
