// Persistent recurrent forward kernel, second generation (sm_100a).
//
// Same ownership as the first kernel — a thread-block cluster of CL = H/32 CTAs shares a slice of batch
// rows, W_hh is split by hidden unit (32 units per CTA) and stays in REGISTERS for the whole sequence,
// h_t is broadcast through distributed shared memory with st.async + mbarrier complete_tx — but the step
// is re-cut after the ncu source view of the first kernel (profiles/r1_rec_gemm_full.md): only 28 % of its
// issued instructions were FFMA2 and the SM issued in < 45 % of the cycles.  The rest was the shuffle
// reduce-scatter over the k-slice lanes, gate math replicated on every lane that ended up with a copy of
// the sums, and a dependent chain shuffle -> MUFU -> shuffle -> DSMEM round trip at the end of every step
// with only two warps per scheduler to hide it (17 warp-instructions per produced (row, unit) value).
//
//  * thread tile = 4 gate columns (ONE hidden unit) x H/8 k-values, FFMA2 over (k, k+1) pairs;
//  * WARP SPECIALISATION: 8 FFMA2 warps (256 threads hold the W_hh slice; setmaxnreg.inc to 184 registers — the
//    pool is the CTA's own launch allocation, 512 x 128) and 8 TAIL warps (setmaxnreg.dec to 72), one per
//    (chunk slot, row), lane = hidden unit: a tail is ~120 dependent instructions, so several must be in
//    flight at once;
//  * the k reduction goes through SHARED MEMORY: every FFMA2 thread stores its <= 4 rows x 4 gate partial
//    sums (conflict-free 16-byte stores), one lane per warp arrives on a CTA-local mbarrier, and the tail
//    warp of the row sums the 8 partials, applies the gates and the cell update and sends h: ~6x fewer
//    instructions than the shuffle tree, nothing replicated, y / c / gates stores coalesced;
//  * the rows of a cluster are cut into `nch` CHUNKS, each an independent recurrence with its own
//    double-buffered h and its own mbarriers, processed round-robin.  The FFMA2 warps never wait for a
//    tail: they go straight on to the next chunk whose h has landed, so the tail's latency chain and the
//    DSMEM flight of chunk i are covered by the arithmetic of chunks i+1, i+2 (fully once nch >= 3);
//  * the tail gathers h of 4 consecutive units with indexed shuffles and every lane sends 16-byte
//    st.async stores (one per destination CTA) that signal the destination's mbarrier;
//  * nch is a runtime value: per-chunk state (c, prefetched x-projection) lives in shared memory, the
//    x-projection of step t+1 is fetched with cp.async one full step ahead by the thread that consumes it.
#include <cstddef>
#include <cstdlib>

#include "mrg_common.cuh"

namespace mrg {

template <int H>
struct Fwd2Cfg {
  static constexpr int CL = H / 32;  // CTAs per cluster (32 hidden units each)
  static constexpr int MK = H / 32;  // float4 k-chunks per thread: k = m*32 + ks*4 + i
  static_assert(H == 128 || H == 256, "unsupported hidden size");
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_local(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// per-chunk shared-memory block (RBC = row capacity of a chunk)
template <int H, int RBC>
struct Fwd2Chunk {
  float h[2][RBC][H];          // h_{t-1} of the chunk's rows, double-buffered (written by all CTAs)
  float4 part[RBC][32][8];     // partial gate sums [row][unit][k-slice]
  float4 xg[RBC][32];          // x-projection of the step being computed
  float c[RBC][32];            // cell state
  unsigned long long hbar[2];  // bytes of h landed in h[b]
  unsigned long long pbar;     // warps whose partial sums are stored
  unsigned long long pad;
};

// partial sums of NR rows over this lane's k-slice: stored as one float4 (4 gates) per row
template <int H, int NR>
__device__ __forceinline__ void fwd2_matvec(const float4 (&w)[4][H / 32], const float* hrow0, float4* part, int pidx) {
  constexpr int MK = H / 32;
  float2 acc[4][NR];
#pragma unroll
  for (int m = 0; m < MK; ++m) {
#pragma unroll
    for (int b = 0; b < NR; ++b) {
      const float4 h4 = *reinterpret_cast<const float4*>(hrow0 + b * H + m * 32);
      const float2 hlo = make_float2(h4.x, h4.y), hhi = make_float2(h4.z, h4.w);
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        if (m == 0) acc[g][b] = fmul2(make_float2(w[g][m].x, w[g][m].y), hlo);
        else ffma2(acc[g][b], make_float2(w[g][m].x, w[g][m].y), hlo);
        ffma2(acc[g][b], make_float2(w[g][m].z, w[g][m].w), hhi);
      }
    }
  }
#pragma unroll
  for (int b = 0; b < NR; ++b)
    part[b * 256 + pidx] = make_float4(acc[0][b].x + acc[0][b].y, acc[1][b].x + acc[1][b].y,
                                       acc[2][b].x + acc[2][b].y, acc[3][b].x + acc[3][b].y);
}

__device__ __forceinline__ void cp_async_wait_dyn(int n) {  // n uniform: at most n groups stay in flight
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

constexpr int FWD2_THREADS = 512;  // warps 0-7: FFMA2 role, warps 8-15: tail role

// GRU = true: the same recurrence machinery runs a GRU (lstmformer's GRU mixer, nn.GRU at mixer_block.py:194).  The caller
// hands the weights in four-gate form — W_ih rows (r, z, n, 0), W_hh rows (r, z, 0, n), bias (b_ir + b_hr, b_iz + b_hz,
// b_in, b_hn) — so that per hidden unit the x-projection slot holds (x_r, x_z, x_n, b_hn) and the recurrent sums are
// (h_r, h_z, 0, h_n); only the cell in the tail differs:
//     r = sig(x_r + h_r), z = sig(x_z + h_z), n = tanh(x_n + r (h_n + b_hn)), h' = n + z (h - n)       (torch.nn.GRU)
// reserve = (r, z, n, h_n + b_hn); there is no cell state.
template <int H, int RBC, bool GRU>
__global__ void __launch_bounds__(FWD2_THREADS, 1) rec_fwd2_kernel(RecArgs a, int slices, int nch) {
  using Cfg = Fwd2Cfg<H>;
  using Chunk = Fwd2Chunk<H, RBC>;
  constexpr int CL = Cfg::CL, MK = Cfg::MK;
  constexpr int NSLOT = 8 / RBC;  // tail warp tw serves row tw % RBC of the chunks ch = tw / RBC (mod NSLOT)
  extern __shared__ __align__(16) uint8_t smem_dyn[];
  Chunk* chunks = reinterpret_cast<Chunk*>(smem_dyn);

  REC_TRACE_DECL
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t rank = cluster_ctarank();
  const int cid = blockIdx.x / CL;
  const int d = cid / slices;
  const int T = a.T, B = a.B;
  const uint32_t BH = (uint32_t)B * H;
  // uneven row split over the clusters, then over the chunks of a cluster
  const int sl = cid % slices, base_rows = B / slices, rem_rows = B % slices;
  const int row0 = sl * base_rows + min(sl, rem_rows);
  const int nrows = base_rows + (sl < rem_rows ? 1 : 0);  // <= RBC * nch
  const int cbase = nrows / nch, crem = nrows % nch;
  const int j0 = rank * 32;

  // reserve of this direction: one (i, f, g, o) group per (t, row, unit): 16 bytes in fp32, 8 bytes in the bf16 mode
  const bool bf = a.bf16_gates != 0;
  char* gates_b = reinterpret_cast<char*>(a.gates) + (size_t)d * T * B * 4 * H * (bf ? 2 : 4);
  float* y_ext = a.y_ext + (size_t)d * (T + 1) * B * H;
  float* c_ext = a.c_ext + (size_t)d * (T + 1) * B * H;
  auto fetch_xg = [&](float4* dst, uint32_t idx) {   // x-projection of (t, row, unit) -> shared memory
    if (bf) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(gates_b + (size_t)idx * 8) : "memory");
    else cp_async16(smem_u32(dst), gates_b + (size_t)idx * 16);
  };

  // ---- initial state (all 16 warps) ---------------------------------------------------------
  const int init_slot = d == 0 ? 0 : T;
  for (int ch = 0; ch < nch; ++ch) {
    Chunk& C = chunks[ch];
    const int nr = cbase + (ch < crem ? 1 : 0);
    const int crow0 = row0 + ch * cbase + min(ch, crem);
    for (int idx = tid; idx < RBC * H; idx += FWD2_THREADS) {
      const int rl = idx / H, k = idx % H;
      C.h[0][rl][k] = rl < nr ? y_ext[((size_t)init_slot * B + crow0 + rl) * H + k] : 0.f;
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
    // =========================== tail warps: one row of a chunk, lane = hidden unit ===========================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
    const int tw = warp - 8;
    const int r = tw % RBC, slot = tw / RBC;
    const int j = j0 + lane;
    const int ntc = nch > slot ? (nch - slot + NSLOT - 1) / NSLOT : 0;  // chunks this warp serves
    // destinations: lane sends the float4 of units 4*(lane/4).. to CTAs (lane & 3) and (lane & 3) + 4
    uint32_t remote_base[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int dst = (lane & 3) + 4 * i;
      remote_base[i] = map_to_cta(smem_u32(chunks), (uint32_t)(dst < CL ? dst : 0));
    }
    // cell state and the x-projection of the first step: one cp.async group per served chunk, in order
    const int t0 = d == 0 ? 0 : T - 1;
    for (int ch = slot; ch < nch; ch += NSLOT) {
      Chunk& C = chunks[ch];
      const int nr = cbase + (ch < crem ? 1 : 0);
      const int crow0 = row0 + ch * cbase + min(ch, crem);
      if (r < nr) {
        C.c[r][lane] = c_ext[((size_t)init_slot * B + crow0 + r) * H + j];
        if (T > 0) fetch_xg(&C.xg[r][lane], (uint32_t)t0 * BH + (uint32_t)(crow0 + r) * H + j);
      }
      cp_async_commit();
    }
    const int tstep = d == 0 ? 1 : -1;
    for (int step = 0; step < T; ++step) {
      const int t = d == 0 ? step : T - 1 - step;
      const uint32_t cur = (uint32_t)(step & 1), nxt = cur ^ 1u;
      const bool send = step + 1 < T;  // nobody consumes the last step's h through shared memory
      const uint32_t obase = (uint32_t)(d == 0 ? t + 1 : t) * BH;  // < 2^31 (host-checked)
      for (int ch = slot; ch < nch; ch += NSLOT) {
        Chunk& C = chunks[ch];
        const int nr = cbase + (ch < crem ? 1 : 0);
        if (r < nr) {
          // x-projection of this step: committed ntc groups ago (one step) by this thread
          if (ntc == 1) cp_async_wait_all(); else cp_async_wait_dyn(ntc - 1);
          const float4 xg = bf ? unpack_bf16x4(*reinterpret_cast<const uint2*>(&C.xg[r][lane])) : C.xg[r][lane];
          const float cold = C.c[r][lane];
          const uint32_t rj = (uint32_t)(row0 + ch * cbase + min(ch, crem) + r) * H + j;
          if (send)  // x-projection of the next step into the slot just read
            fetch_xg(&C.xg[r][lane], (uint32_t)(t + tstep) * BH + rj);
          REC_TRACE(10, ch, step);
          mbar_wait(smem_u32(&C.pbar), cur);  // all 8 FFMA2 warps have stored their partial sums
          REC_TRACE(11, ch, step);
          float4 p[8];
#pragma unroll
          for (int i = 0; i < 8; ++i)  // rotated start: the lanes of a quarter-warp hit 8 different 16-byte columns
            p[i] = C.part[r][lane][(i + lane) & 7];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            p[i].x += p[i + 4].x; p[i].y += p[i + 4].y; p[i].z += p[i + 4].z; p[i].w += p[i + 4].w;
          }
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            p[i].x += p[i + 2].x; p[i].y += p[i + 2].y; p[i].z += p[i + 2].z; p[i].w += p[i + 2].w;
          }
          float gi, gf, gg, go, cn, h;
          if (GRU) {
            gi = fast_sigmoid((p[0].x + p[1].x) + xg.x);                 // r
            gf = fast_sigmoid((p[0].y + p[1].y) + xg.y);                 // z
            go = (p[0].w + p[1].w) + xg.w;                               // W_hn h + b_hn
            gg = fast_tanh(fmaf(gi, go, xg.z));                          // n
            const float hprev = C.h[cur][r][j0 + lane];
            cn = 0.f;
            h = fmaf(gf, hprev - gg, gg);
          } else {
            gi = fast_sigmoid((p[0].x + p[1].x) + xg.x);
            gf = fast_sigmoid((p[0].y + p[1].y) + xg.y);
            gg = fast_tanh((p[0].z + p[1].z) + xg.z);
            go = fast_sigmoid((p[0].w + p[1].w) + xg.w);
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
                                              ((nxt * RBC + r) * H + j0 + (lane & ~3)) * sizeof(float));
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
        cp_async_commit();
      }
    }
    cp_async_wait_all();
    return;
  }

  // =========================== FFMA2 warps ===========================================================================
  asm volatile("setmaxnreg.inc.sync.aligned.u32 184;");
  const int ks = lane & 7, ugl = lane >> 3;
  const int u = warp * 4 + ugl;  // local hidden unit
  const float* __restrict__ W = d == 0 ? a.w_hh[0] : a.w_hh[1];
  float4 w[4][MK];  // W_hh slice, resident for the whole sequence
#pragma unroll
  for (int g = 0; g < 4; ++g)
#pragma unroll
    for (int m = 0; m < MK; ++m)
      w[g][m] = __ldg(reinterpret_cast<const float4*>(W + (size_t)(g * H + j0 + u) * H + m * 32 + ks * 4));
  const int pidx = u * 8 + ks;
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
      const float* hrow0 = &C.h[cur][0][ks * 4];
      float4* part = &C.part[0][0][0];
      switch (nr) {
        case 1: fwd2_matvec<H, 1>(w, hrow0, part, pidx); break;
        case 2: fwd2_matvec<H, 2>(w, hrow0, part, pidx); break;
        case 3: fwd2_matvec<H, (RBC > 2 ? 3 : 1)>(w, hrow0, part, pidx); break;
        default: fwd2_matvec<H, (RBC > 2 ? 4 : 1)>(w, hrow0, part, pidx); break;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive_local(smem_u32(&C.pbar));
      REC_TRACE(3, ch, step);
    }
  }
  // Exit safety: the last round of remote stores into this CTA (step T-2) was waited for at step T-1, and
  // nobody sends at step T-1, so no store can target the shared memory of an exited CTA.
}

// chunks one CTA can hold: the forward and the backward kernel both keep < 26 KB per 4-row chunk
int rec2_max_chunks(int H, int rbc) {
  (void)H;
  return rbc == 2 ? 12 : 8;
}

template <int H, int RBC, bool GRU>
static int launch_fwd2(const RecArgs& a, int slices, int nch, cudaStream_t stream) {
  static bool attr_set = false;
  const size_t smem = (size_t)nch * sizeof(Fwd2Chunk<H, RBC>);
  if (!attr_set) {
    MRG_CUDA_CHECK(cudaFuncSetAttribute(rec_fwd2_kernel<H, RBC, GRU>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)(rec2_max_chunks(H, RBC) * sizeof(Fwd2Chunk<H, RBC>))));
    attr_set = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(a.D * slices * Fwd2Cfg<H>::CL));
  cfg.blockDim = dim3(FWD2_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = Fwd2Cfg<H>::CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  static char name[64];
  if (!name[0]) snprintf(name, sizeof(name), GRU ? "mrg::rec_fwd2_kernel<%d, %d, gru>" : "mrg::rec_fwd2_kernel<%d, %d>", H, RBC);
  ProfScope prof(PROF_REC_FWD, stream, name);
  count_launch();
  MRG_CUDA_CHECK(cudaLaunchKernelEx(&cfg, rec_fwd2_kernel<H, RBC, GRU>, a, slices, nch));
  return 0;
}

template <int H>
static int max_clusters_fwd2() {
  if (cudaFuncSetAttribute(rec_fwd2_kernel<H, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)(12 * sizeof(Fwd2Chunk<H, 2>))) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(Fwd2Cfg<H>::CL * 64);
  cfg.blockDim = dim3(FWD2_THREADS);
  cfg.dynamicSmemBytes = 3 * sizeof(Fwd2Chunk<H, 2>);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = Fwd2Cfg<H>::CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, rec_fwd2_kernel<H, 2, false>, &cfg) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

// co-resident clusters of the second-generation kernels (one CTA per SM: 256 threads at > 128 registers)
int max_active_clusters2(int H) {
  static int cache[2] = {-1, -1};
  const int i = H == 256 ? 0 : H == 128 ? 1 : -1;
  if (i < 0) return 0;
  if (cache[i] < 0) cache[i] = H == 256 ? max_clusters_fwd2<256>() : max_clusters_fwd2<128>();
  return cache[i];
}

bool rec2_supported(int H) { return H == 128 || H == 256; }

// Partition of the batch for the second-generation kernels: `slices` clusters per direction (one wave of
// co-resident clusters whenever B allows, rows split as evenly as possible) and `nch` chunks of <= `rbc`
// rows per cluster.  Measured on B200 (tools/rec_bench.py): every chunk-iteration costs the FFMA2 warps
// ~500 cycles of exposed latency (barrier probe, first shared-memory load, epilogue) on top of 272 cycles of
// FFMA2 issue per row, so few large chunks win as soon as two of them cover each other's tail + DSMEM
// flight: 4-row chunks, at least 2 of them; clusters with <= 3 rows run one row per chunk.
// MRG_REC_NCH / MRG_REC_RBC override nch / rbc (tuning experiments).
void pick_partition2(int H, int B, int D, int budget, int* slices_out, int* nch_out, int* rbc_out) {
  int maxc = max_active_clusters2(H);
  if (maxc <= 0) maxc = H == 256 ? 15 : 30;
  if (budget > 0 && budget < maxc) maxc = budget;
  int per_dir = maxc / D;
  if (per_dir < 1) per_dir = 1;
  int slices = B < per_dir ? B : per_dir;
  int rows = (B + slices - 1) / slices;
  if (rows > 32) {  // more rows than one cluster holds: several waves
    rows = 32;
    slices = (B + rows - 1) / rows;
    rows = (B + slices - 1) / slices;
  }
  static int forced_nch = -1, forced_rbc = -1;
  if (forced_nch < 0) {
    const char* e = getenv("MRG_REC_NCH");
    forced_nch = e ? atoi(e) : 0;
    e = getenv("MRG_REC_RBC");
    forced_rbc = e ? atoi(e) : 0;
  }
  int rbc = rows <= 3 ? 2 : 4;
  if ((forced_rbc == 2 && rows <= 24) || forced_rbc == 4) rbc = forced_rbc;
  int nch = rows <= 3 ? rows : (rows + rbc - 1) / rbc;
  if (nch < 2 && rows >= 2) nch = 2;
  if (forced_nch > 0 && forced_nch * rbc >= rows && forced_nch <= rec2_max_chunks(H, rbc)) nch = forced_nch;
  *slices_out = slices;
  *nch_out = nch;
  *rbc_out = rbc;
}

int rec_forward_cluster2(const RecArgs& a, cudaStream_t stream) {
  MRG_REQUIRE((long long)(a.T + 1) * a.B * a.H * 4 < (1LL << 31),
              "rec_forward_cluster2: T*B*4H exceeds the 32-bit index range of one direction");
  int slices, nch, rbc;
  pick_partition2(a.H, a.B, a.D, a.cluster_budget, &slices, &nch, &rbc);
  if (a.gru) {
    if (a.H == 256) return rbc == 2 ? launch_fwd2<256, 2, true>(a, slices, nch, stream) : launch_fwd2<256, 4, true>(a, slices, nch, stream);
    if (a.H == 128) return rbc == 2 ? launch_fwd2<128, 2, true>(a, slices, nch, stream) : launch_fwd2<128, 4, true>(a, slices, nch, stream);
  }
  if (a.H == 256) return rbc == 2 ? launch_fwd2<256, 2, false>(a, slices, nch, stream) : launch_fwd2<256, 4, false>(a, slices, nch, stream);
  if (a.H == 128) return rbc == 2 ? launch_fwd2<128, 2, false>(a, slices, nch, stream) : launch_fwd2<128, 4, false>(a, slices, nch, stream);
  set_error("rec_forward_cluster2: unsupported hidden size %d", a.H);
  return MRG_E_UNSUPPORTED;
}

}  // namespace mrg
