// Persistent BPTT kernel, second generation (sm_100a) — the mirror of mrg_rec_fwd2.cu.
//
// A cluster of CL = H/32 CTAs owns a slice of batch rows; CTA `rank` owns hidden units [32*rank, 32*rank+32)
// and keeps the 128 gate rows of W_hh that belong to them ([128][H]) in registers.  The rows of a cluster
// are cut into `nch` chunks (the forward's partition), each an independent backward recurrence with its own
// double-buffered exchange buffer and mbarriers, processed round-robin.  Per chunk and step:
//   head  (8 dedicated warps, 8-15: ONE warp per (chunk slot, row), lane = hidden unit)
//         dh_t = dy_t + sum over the CL source CTAs of their partial dh (fixed order -> deterministic);
//         gate derivatives -> dpre, written over the gates reserve and into local shared memory; the
//         bias-gradient / dc / c state and the cp.async prefetch slots of the next step live in shared
//         memory; one lane arrives on the chunk's CTA-local mbarrier;
//   body  (all 256 threads) partial dh_{t-1}[row][k] = sum over this CTA's 128 gate columns of
//         dpre[row][n] * W_hh[n][k]: thread tile 4 k x H/8 n, FFMA2 over (n, n+1) pairs, shuffle reduce over
//         the 1024/H lanes that share a k-group, one 16-byte st.async per (row, k-group) to the CTA that owns
//         those k, signalling its mbarrier.
// Head and body warps only meet through mbarriers: the head warps run ahead through their latency chain
// (smem sums, MUFU, global store) while the 8 body warps run the FFMA2 block of another chunk, and the DSMEM
// flight of chunk i is covered by the bodies of chunks i+1, i+2 (nch >= 3).
#include <cstddef>

#include "mrg_common.cuh"

namespace mrg {

template <int H>
struct Bwd2Cfg {
  static constexpr int CL = H / 32;    // CTAs per cluster
  static constexpr int RS = 1024 / H;  // lanes that split the 128 local gate columns of one k-group
  static constexpr int MM = H / 32;    // float4 chunks of dpre per thread and row: n = m*(RS*4) + rs*4 + i
  static_assert(H == 128 || H == 256, "unsupported hidden size");
};

__device__ __forceinline__ void cpb_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cpb_async4(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cpb_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cpb_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void mbarb_arrive_local(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

template <int H, int RBC>
struct Bwd2Chunk {
  float part[2][H / 32][RBC][32];  // partial dh from every source CTA, double-buffered
  float dpre[RBC][128];            // d(pre-activation) of this CTA's 128 gate columns (unit-major, gate-minor)
  float4 g[RBC][32];               // prefetched gates of the step
  float4 db[RBC][32];              // bias-gradient accumulator (sum over t of dpre)
  float cp[RBC][32];               // prefetched c_{t-1}
  float dy[RBC][32];               // prefetched dy_t
  float dc[RBC][32];               // carried dc
  float c_cur[RBC][32];            // c_t of the step being processed
  unsigned long long hbar[2];      // bytes of partial dh landed in part[b]
  unsigned long long dbar;         // head warps that have published dpre
  unsigned long long rbar;         // body warps that have finished reading dpre
};

// v[b*4+kk] = sum over this lane's gate columns of dpre[b][n] * W[n][k0+kk]
template <int MM, int RS, int NR, int RBC>
__device__ __forceinline__ void bwd2_matvec(const float4 (&w)[MM][4], const float* drow0, float (&v)[RBC * 4]) {
  float2 acc[4][NR];
#pragma unroll
  for (int m = 0; m < MM; ++m) {
#pragma unroll
    for (int b = 0; b < NR; ++b) {
      const float4 d4 = *reinterpret_cast<const float4*>(drow0 + b * 128 + m * (RS * 4));
      const float2 dlo = make_float2(d4.x, d4.y), dhi = make_float2(d4.z, d4.w);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        if (m == 0) acc[kk][b] = fmul2(make_float2(w[m][kk].x, w[m][kk].y), dlo);
        else ffma2(acc[kk][b], make_float2(w[m][kk].x, w[m][kk].y), dlo);
        ffma2(acc[kk][b], make_float2(w[m][kk].z, w[m][kk].w), dhi);
      }
    }
  }
#pragma unroll
  for (int b = 0; b < RBC; ++b)
#pragma unroll
    for (int kk = 0; kk < 4; ++kk)
      v[b * 4 + kk] = b < NR ? acc[kk][b < NR ? b : 0].x + acc[kk][b < NR ? b : 0].y : 0.f;
}

__device__ __forceinline__ void cpb_wait_dyn(int n) {  // n uniform: at most n groups stay in flight
  switch (n) {
    case 0: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
    case 1: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
    case 2: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
    case 3: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
    case 4: asm volatile("cp.async.wait_group 4;" ::: "memory"); break;
    case 5: asm volatile("cp.async.wait_group 5;" ::: "memory"); break;
    case 6: asm volatile("cp.async.wait_group 6;" ::: "memory"); break;
    case 7: asm volatile("cp.async.wait_group 7;" ::: "memory"); break;
    case 8: asm volatile("cp.async.wait_group 8;" ::: "memory"); break;
    case 9: asm volatile("cp.async.wait_group 9;" ::: "memory"); break;
    case 10: asm volatile("cp.async.wait_group 10;" ::: "memory"); break;
    default: asm volatile("cp.async.wait_group 11;" ::: "memory"); break;
  }
}

constexpr int BWD2_THREADS = 512;  // warps 0-7: body (FFMA2) role, warps 8-15: head role

// GRU = true (see rec_fwd2_kernel): reserve in = (r, z, n, q) with q = W_hn h + b_hn; with dh the gradient at h_t,
//     d(n) = dh (1 - z), d(z) = dh (h_{t-1} - n); dn_pre = d(n) (1 - n^2), dr_pre = dn_pre q r (1 - r), dz_pre = d(z) z (1 - z)
// reserve out = (dr_pre, dz_pre, dn_pre, dn_pre r): slots (0, 1, 2) are the gradient at the x-projection, slots (0, 1, 3) at
// the recurrent projection — with W_ih rows (r, z, n, 0) and W_hh rows (r, z, 0, n) ONE four-slot vector serves dX, both
// weight gradients, both bias gradients and the transposed mat-vec of the body warps.  The direct path dh z to h_{t-1}
// is carried in the slot the LSTM uses for d(c).
template <int H, int RBC, bool GRU>
__global__ void __launch_bounds__(BWD2_THREADS, 1) rec_bwd2_kernel(RecBwdArgs a, int slices, int nch) {
  using Cfg = Bwd2Cfg<H>;
  using Chunk = Bwd2Chunk<H, RBC>;
  constexpr int CL = Cfg::CL, RS = Cfg::RS, MM = Cfg::MM;
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
  const int nrows = base_rows + (sl < rem_rows ? 1 : 0);  // same split as the forward
  const int cbase = nrows / nch, crem = nrows % nch;
  const int j0 = rank * 32;

  // reserve of this direction: (i, f, g, o) in, d(pre-activations) out; 16 bytes per unit in fp32, 8 in the bf16 mode
  const bool bf = a.bf16_gates != 0;
  char* gates_b = reinterpret_cast<char*>(a.gates) + (size_t)d * T * B * 4 * H * (bf ? 2 : 4);
  // state saved by the forward that the head reads one step back: c_{t-1} (LSTM) or h_{t-1} (GRU)
  const float* c_ext = (GRU ? a.y_ext : a.c_ext) + (size_t)d * (T + 1) * B * H;

  // ---- shared-memory state (all 12 warps) -------------------------------------------------------------------
  for (int ch = 0; ch < nch; ++ch) {
    Chunk& C = chunks[ch];
    const int nr = cbase + (ch < crem ? 1 : 0);
    for (int idx = tid; idx < 2 * CL * RBC * 32; idx += BWD2_THREADS) (&C.part[0][0][0][0])[idx] = 0.f;
    for (int idx = tid; idx < RBC * 128; idx += BWD2_THREADS) (&C.dpre[0][0])[idx] = 0.f;
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
    // =========================== head warps: row r of every chunk, lane = hidden unit ===========================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
    constexpr int NSLOT = 8 / RBC;  // head warp hw serves row hw % RBC of the chunks ch = hw / RBC (mod NSLOT)
    const int hw = warp - 8;
    const int r = hw % RBC, slot = hw / RBC;
    const int j = j0 + lane;
    const int ntc = nch > slot ? (nch - slot + NSLOT - 1) / NSLOT : 0;  // chunks this warp serves
    // prefetch of one step for (chunk, row r); the caller commits the group
    auto prefetch = [&](Chunk& C, int crow0, int step) {
      const int t = d == 0 ? step : T - 1 - step;
      const int prev_slot = d == 0 ? t : t + 1;
      const uint32_t rj = (uint32_t)(crow0 + r) * H + j;
      if (bf) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(&C.g[r][lane])),
                           "l"(gates_b + (size_t)((uint32_t)t * BH + rj) * 8) : "memory");
      else cpb_async16(smem_u32(&C.g[r][lane]), gates_b + (size_t)((uint32_t)t * BH + rj) * 16);
      cpb_async4(smem_u32(&C.cp[r][lane]), c_ext + (uint32_t)prev_slot * BH + rj);
      if (a.dy) cpb_async4(smem_u32(&C.dy[r][lane]), a.dy + ((uint32_t)t * B + crow0 + r) * (uint32_t)(D * H) + d * H + j);
    };
    {
      for (int ch = slot; ch < nch; ch += NSLOT) {
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
        cpb_commit();
      }
    }
    cluster_sync_all();
    uint32_t hphases = 0;  // bit (ch*2 + buf) = parity of hbar to wait for next
    for (int iter = 0; iter < T; ++iter) {
      const int step = T - 1 - iter;
      const int t = d == 0 ? step : T - 1 - step;
      const int cur = iter & 1;
      for (int ch = slot; ch < nch; ch += NSLOT) {
        Chunk& C = chunks[ch];
        const int nr = cbase + (ch < crem ? 1 : 0);
        if (r < nr) {
          const int crow0 = row0 + ch * cbase + min(ch, crem);
          // loads of this step: committed ntc groups ago (one step) by this thread
          if (ntc == 1) cpb_wait_all(); else cpb_wait_dyn(ntc - 1);
          const float4 g = bf ? unpack_bf16x4(*reinterpret_cast<const uint2*>(&C.g[r][lane])) : C.g[r][lane];
          const float cprev = C.cp[r][lane];
          float dh = C.dy[r][lane];
          const float tc = GRU ? 0.f : fast_tanh(C.c_cur[r][lane]);
          const float dcin = C.dc[r][lane];
          if (iter + 1 < T) prefetch(C, crow0, step - 1);
          const uint32_t hbar_cur = smem_u32(&C.hbar[cur]);
          if (iter > 0) {  // the partial dh of all source CTAs have landed in part[cur]
            mbar_wait(hbar_cur, (hphases >> (ch * 2 + cur)) & 1u);
            hphases ^= 1u << (ch * 2 + cur);
          }
          // re-arm for the round of iteration iter+1 (which writes part[cur] again)
          if (r == 0 && lane == 0 && iter + 1 < T)
            mbar_arrive_expect_tx(hbar_cur, (uint32_t)(CL * nr * 32 * sizeof(float)));
          // The exchange only orders this head behind ONE body warp per source CTA (the one whose k-group this
          // CTA owns); dpre may only be overwritten once all 8 local body warps have read the previous step.
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
          if (lane == 0) mbarb_arrive_local(smem_u32(&C.dbar));
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
        cpb_commit();
      }
    }
    cpb_wait_dyn(0);
    // ---- dh0 / dc0 / bias-gradient partials ----------------------------------------------------------------
    const int fin = T & 1;
    for (int ch = slot; ch < nch; ch += NSLOT) {
      Chunk& C = chunks[ch];
      const int nr = cbase + (ch < crem ? 1 : 0);
      if (r >= nr) continue;
      const int crow0 = row0 + ch * cbase + min(ch, crem);
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
  const int rs = tid % RS, kg = tid / RS;
  const float* __restrict__ W = d == 0 ? a.w_hh[0] : a.w_hh[1];
  // W_hh slice -> registers: w[m][kk] = 4 gates (i,f,g,o) of local unit m*RS+rs at column kg*4+kk
  float4 w[MM][4];
#pragma unroll
  for (int m = 0; m < MM; ++m) {
    float4 r4[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
      r4[i] = __ldg(reinterpret_cast<const float4*>(W + (size_t)(i * H + j0 + m * RS + rs) * H + kg * 4));
    w[m][0] = make_float4(r4[0].x, r4[1].x, r4[2].x, r4[3].x);
    w[m][1] = make_float4(r4[0].y, r4[1].y, r4[2].y, r4[3].y);
    w[m][2] = make_float4(r4[0].z, r4[1].z, r4[2].z, r4[3].z);
    w[m][3] = make_float4(r4[0].w, r4[1].w, r4[2].w, r4[3].w);
  }
  // destination of this lane's reduced float4 = the CTA that owns k = kg*4 .. kg*4+3
  constexpr int ROWBITS = RBC == 2 ? 1 : 2;                          // reduce-scatter stages (row bits)
  const int rrow = rs >> (RS == 4 ? (2 - ROWBITS) : (3 - ROWBITS));  // row held after the reduce
  const bool sender = (rs & ((RS >> ROWBITS) - 1)) == 0;
  const uint32_t owner = (uint32_t)(kg >> 3);
  const int kin = (kg & 7) * 4;
  const uint32_t remote_base = map_to_cta(smem_u32(chunks), owner);
  cluster_sync_all();

  for (int iter = 0; iter < T; ++iter) {
    const int nxt = (iter & 1) ^ 1;
    for (int ch = 0; ch < nch; ++ch) {
      Chunk& C = chunks[ch];
      const int nr = cbase + (ch < crem ? 1 : 0);
      if (nr == 0) continue;
      mbar_wait(smem_u32(&C.dbar), (uint32_t)(iter & 1));  // dpre of (ch, iter) is published
      // ---- partial dh_{prev}[row][k] over this CTA's gate columns ----------------------------------------
      const float* drow0 = &C.dpre[0][rs * 4];
      float v[RBC * 4];
      switch (nr) {
        case 1: bwd2_matvec<MM, RS, 1, RBC>(w, drow0, v); break;
        case 2: bwd2_matvec<MM, RS, 2, RBC>(w, drow0, v); break;
        case 3: bwd2_matvec<MM, RS, (RBC > 2 ? 3 : 1), RBC>(w, drow0, v); break;
        default: bwd2_matvec<MM, RS, (RBC > 2 ? 4 : 1), RBC>(w, drow0, v); break;
      }
      __syncwarp();
      if (lane == 0) mbarb_arrive_local(smem_u32(&C.rbar));  // this warp is done reading dpre of (ch, iter)
      // reduce over the RS lanes of the k-group: scatter over the row bits, all-reduce over the rest
      float v4[4];
      if (RBC == 4) {
        float v8[8];
        {
          const bool up = (rs & (RS / 2)) != 0;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float sendv = up ? v[i] : v[(8 + i) % (RBC * 4)];
            const float keep = up ? v[(8 + i) % (RBC * 4)] : v[i];
            v8[i] = keep + __shfl_xor_sync(0xffffffffu, sendv, RS / 2);
          }
        }
        {
          const bool up = (rs & (RS / 4)) != 0;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float sendv = up ? v8[i] : v8[4 + i];
            const float keep = up ? v8[4 + i] : v8[i];
            v4[i] = keep + __shfl_xor_sync(0xffffffffu, sendv, RS / 4);
          }
        }
#pragma unroll
        for (int s = RS / 8; s >= 1; s >>= 1)
#pragma unroll
          for (int i = 0; i < 4; ++i) v4[i] += __shfl_xor_sync(0xffffffffu, v4[i], s);
      } else {
        {
          const bool up = (rs & (RS / 2)) != 0;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float sendv = up ? v[i] : v[4 + i];
            const float keep = up ? v[4 + i] : v[i];
            v4[i] = keep + __shfl_xor_sync(0xffffffffu, sendv, RS / 2);
          }
        }
#pragma unroll
        for (int s = RS / 4; s >= 1; s >>= 1)
#pragma unroll
          for (int i = 0; i < 4; ++i) v4[i] += __shfl_xor_sync(0xffffffffu, v4[i], s);
      }
      // ---- one 16-byte store per (row, k-group) to the owner of these 4 k -----------------------------------
      if (sender && rrow < nr) {
        const uint32_t off = (uint32_t)(ch * sizeof(Chunk) + offsetof(Chunk, part) +
                                        (((nxt * CL + (int)rank) * RBC + rrow) * 32 + kin) * sizeof(float));
        const uint32_t off_bar = (uint32_t)(ch * sizeof(Chunk) + offsetof(Chunk, hbar) + nxt * 8);
        st_async_v4(remote_base + off, make_float4(v4[0], v4[1], v4[2], v4[3]), remote_base + off_bar);
      }
    }
  }
}

template <int H, int RBC, bool GRU>
static int launch_bwd2(const RecBwdArgs& a, int slices, int nch, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    MRG_CUDA_CHECK(cudaFuncSetAttribute(rec_bwd2_kernel<H, RBC, GRU>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)(rec2_max_chunks(H, RBC) * sizeof(Bwd2Chunk<H, RBC>))));
    attr_set = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(a.D * slices * Bwd2Cfg<H>::CL));
  cfg.blockDim = dim3(BWD2_THREADS);
  cfg.dynamicSmemBytes = (size_t)nch * sizeof(Bwd2Chunk<H, RBC>);
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = Bwd2Cfg<H>::CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  static char name[64];
  if (!name[0]) snprintf(name, sizeof(name), GRU ? "mrg::rec_bwd2_kernel<%d, %d, gru>" : "mrg::rec_bwd2_kernel<%d, %d>", H, RBC);
  ProfScope prof(PROF_REC_BWD, stream, name);
  count_launch();
  MRG_CUDA_CHECK(cudaLaunchKernelEx(&cfg, rec_bwd2_kernel<H, RBC, GRU>, a, slices, nch));
  return 0;
}

int rec_backward_cluster2(const RecBwdArgs& a, cudaStream_t stream) {
  MRG_REQUIRE((long long)(a.T + 1) * a.B * a.H * 4 * a.D < (1LL << 31),
              "rec_backward_cluster2: T*B*4H*D exceeds the 32-bit index range");
  int slices, nch, rbc;
  pick_partition2(a.H, a.B, a.D, a.cluster_budget, &slices, &nch, &rbc);
  if (a.gru) {
    if (a.H == 256) return rbc == 2 ? launch_bwd2<256, 2, true>(a, slices, nch, stream) : launch_bwd2<256, 4, true>(a, slices, nch, stream);
    if (a.H == 128) return rbc == 2 ? launch_bwd2<128, 2, true>(a, slices, nch, stream) : launch_bwd2<128, 4, true>(a, slices, nch, stream);
  }
  if (a.H == 256) return rbc == 2 ? launch_bwd2<256, 2, false>(a, slices, nch, stream) : launch_bwd2<256, 4, false>(a, slices, nch, stream);
  if (a.H == 128) return rbc == 2 ? launch_bwd2<128, 2, false>(a, slices, nch, stream) : launch_bwd2<128, 4, false>(a, slices, nch, stream);
  set_error("rec_backward_cluster2: unsupported hidden size %d", a.H);
  return MRG_E_UNSUPPORTED;
}

}  // namespace mrg
