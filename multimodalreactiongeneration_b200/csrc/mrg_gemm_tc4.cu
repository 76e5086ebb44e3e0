// Projection GEMM for a WEIGHT B operand, third design: persistent 128 x 256 tiles, B pre-split.
//
// What the measurements of round 2 say about the 128 x 128 kernel (mrg_gemm_tc2.cu; tools/mma_rate.cu,
// tools/gemm_trace.py, profiles/r2_gemm_*.txt):
//  * a tcgen05.mma with M = 128 costs the SAME ~116 cycles for N = 64 and N = 128 and 128 cycles for N = 256: an
//    N = 128 tile runs the tensor pipe at 55 % of the rate of an N = 256 tile.  The 3xTF32 main loop of the old kernel
//    is MMA-bound at that rate (12 MMAs per k-block: ~1900 cycles per CTA with two CTAs per SM);
//  * a CTA spends ~4000 cycles before its first MMA (barrier / TMEM setup, cold TMA, first conversion) and ~5000 after
//    its last one (pipe drain, TMEM -> registers -> shared memory -> global) against ~17000 of main loop at K = 256.
// This kernel therefore
//  * issues 128 x 256 x 8 MMAs (one CTA per SM, 256 accumulator columns of tensor memory);
//  * is PERSISTENT: a CTA walks the tile list, the operand pipelines never drain between tiles, and the epilogue warps
//    (TMEM -> registers -> private shared-memory transpose -> global) of tile n run under the main loop of tile n+1;
//  * takes B = W (a weight matrix) ALREADY SPLIT into tf32 hi / lo planes (written once per step by pack_kernel /
//    mrg_split_tf32), so that B goes TMA -> shared memory -> tensor core with no conversion pass: shared-memory bytes per
//    128 x 256 x 32 k-block: 16 KB (A in) + 16 KB (A converter reads) + 64 KB (B in) + 96 KB (MMA operand reads)
//    = 192 KB = 1500 cycles at 128 B/clk against 1536 cycles of MMA;
//  * keeps the A path of the second generation: raw fp32 (or bfloat16) tile by TMA, converter warps split it in registers
//    and write hi | lo into tensor memory, the MMAs take A from TMEM (.ts form) — 3 shared-memory stages (free again as
//    soon as the converter has read them) feeding 3 tensor-memory stages, decoupled from the 2 B stages (64 KB each; a
//    third does not fit beside the epilogue's transpose buffers — the hi / lo barrier split below hides that);
//  * drains the accumulator in one go: every epilogue thread pulls its 128 columns into registers (setmaxnreg moves
//    registers from the producer / MMA warps to the epilogue warps), hands the accumulator back to the MMA issuer ~300
//    cycles after the last MMA and only then adds the bias and stores through a per-warp shared-memory transpose (full
//    128-byte row segments per store instruction) — the first version held the accumulator through the global stores,
//    ~5900 cycles per tile with every SM writing at once; storing 16 bytes per lane straight from the registers was
//    slower still (profiles/r2_gemm_tc4_trace.txt, r2_gemm_split_check.txt).
// Same arithmetic as the 128 x 128 kernel: D += A_lo B_hi + A_hi B_lo + A_hi B_hi per k-step, fp32 accumulation in TMEM.
//
// Tiles are handed out DYNAMICALLY (warp 2: atomic counter -> 4-deep shared-memory ring -> every role): the step runs
// several streams side by side (two encoder stacks, weight-gradient GEMMs), so a CTA of this kernel may become resident
// long after its siblings — a static tile list per CTA made the kernel as slow as its latest CTA.
//
// Warp roles (512 threads): 0 A producer (TMA) | 1 MMA issuer | 2 TMEM allocator + tile scheduler | 3 B producer (TMA) |
// 4-7 A converter (TMEM lane quarter = warp % 4) | 8-15 epilogue (lane quarter = warp % 4, column half = (warp - 8) / 4).
// TMEM: accumulator at columns [0, 256), A stage s at 256 + 64 s (hi | lo).
// Barriers: a_full[s] TMA -> converters, a_free[s] converters -> A producer, a_cvt[t] converters -> MMA, t_empty[t] MMA ->
// converters, bh_full / bl_full[s] TMA -> MMA, bh_empty / bl_empty[s] MMA -> B producer (the hi and the lo plane of a B
// stage are separate buffers with their own barriers: the MMAs of a k-block use B_hi first — A_lo B_hi, A_hi B_hi for the
// four k-steps — and B_lo last, so the hi buffer is refilled while the last third of the k-block still runs),
// acc_full MMA -> epilogue, acc_empty epilogue -> MMA.
#include <atomic>
#include <cstdlib>

#include "mrg_tc_common.cuh"

namespace mrg {

constexpr int NT4 = 256;                                  // tile columns
constexpr int SA4 = 3, ST4 = 3, SB4 = 2;                  // A shared-memory / A tensor-memory / B pipeline stages
constexpr int B4_HALF = 2 * TILE_BYTES;                   // one 256 x 32 tf32 plane (hi or lo): 32 KB
constexpr int B4_STAGE = 2 * B4_HALF;
constexpr int EPI4_WARPS = 8;
constexpr int EPI4_LD = 36;                               // floats per row of an epilogue staging buffer ([32][36] per warp)
constexpr int EPI4_BYTES = EPI4_WARPS * 32 * EPI4_LD * 4;
constexpr int SMEM4_BYTES = SA4 * TILE_BYTES + SB4 * B4_STAGE + EPI4_BYTES + 1024 + 256;
constexpr int TC4_THREADS = 512;

#ifdef MRG_REC_TRACE
#define TC4_TRACE_DECL unsigned trace_n = 0; const bool trace_cta = p.trace && blockIdx.x == gridDim.x / 2;
#define TC4_TRACE(evt, i)                                                                                   \
  if (trace_cta && (threadIdx.x & 31) == 0 && trace_n < 1024u) {                                            \
    unsigned long long* tp = p.trace + ((size_t)(threadIdx.x >> 5) * 1024 + trace_n) * 2;                   \
    tp[0] = clock64();                                                                                      \
    tp[1] = ((unsigned long long)(threadIdx.x >> 5) << 48) | ((unsigned long long)(evt) << 32) | (unsigned long long)(i); \
    ++trace_n;                                                                                              \
  }
#else
#define TC4_TRACE_DECL
#define TC4_TRACE(evt, i)
#endif

struct Tc4Work {
  int tiles_n, items;
  unsigned int* counter;   // [0] tiles handed out beyond the first one per CTA, [1] CTAs that have finished; both 0 at launch
};

constexpr int SCHED4 = 4;            // depth of the tile ring
constexpr int SCHED4_CONSUMERS = 15; // A producer, B producer, MMA issuer, 4 converter warps, 8 epilogue warps

__global__ void __launch_bounds__(TC4_THREADS, 1)
gemm_tc4_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_bh,
                const __grid_constant__ CUtensorMap tma_bl, TcParams p, Tc4Work w) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = base;
  const uint32_t b_base = base + SA4 * TILE_BYTES;
  const uint32_t epi_base = b_base + SB4 * B4_STAGE;
  const uint32_t bar_base = epi_base + EPI4_BYTES;
  auto a_full = [&](int s) { return bar_base + 8u * s; };
  auto a_free = [&](int s) { return bar_base + 8u * (SA4 + s); };
  auto a_cvt = [&](int s) { return bar_base + 8u * (2 * SA4 + s); };
  auto t_empty = [&](int s) { return bar_base + 8u * (2 * SA4 + ST4 + s); };
  auto bh_full = [&](int s) { return bar_base + 8u * (2 * SA4 + 2 * ST4 + s); };
  auto bh_empty = [&](int s) { return bar_base + 8u * (2 * SA4 + 2 * ST4 + SB4 + s); };
  auto bl_full = [&](int s) { return bar_base + 8u * (2 * SA4 + 2 * ST4 + 2 * SB4 + s); };
  auto bl_empty = [&](int s) { return bar_base + 8u * (2 * SA4 + 2 * ST4 + 3 * SB4 + s); };
  const uint32_t accf_bar = bar_base + 8u * (2 * SA4 + 2 * ST4 + 4 * SB4);
  const uint32_t acce_bar = accf_bar + 8u;
  auto sched_full = [&](int s) { return acce_bar + 8u + 8u * s; };
  auto sched_empty = [&](int s) { return acce_bar + 8u + 8u * (SCHED4 + s); };
  const uint32_t sched_tile = acce_bar + 8u + 8u * (2 * SCHED4);   // int[SCHED4]
  const uint32_t tmem_slot = sched_tile + 4u * SCHED4;
  uint8_t* smem_gen = smem_raw + (base - smem_u32(smem_raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // One-pass mode with a K-major fp32 A: the raw tile IS a valid tf32 operand (the tensor core ignores the low 13 mantissa
  // bits), so the MMAs take A straight from the TMA stage (.ss form) — no converter pass, no tensor-memory staging; the
  // shared-memory stage is released by tcgen05.commit.  (bfloat16 or MN-major A still goes through the converters.)
  const bool ss = p.single_pass && !p.a_bf16 && !p.a_mn;
  TC4_TRACE_DECL
  TC4_TRACE(1, 0);

  if (threadIdx.x == 0) {
    for (int s = 0; s < SA4; ++s) {
      mbar_init(a_full(s), 1);
      mbar_init(a_free(s), ss ? 1 : 4);   // the four A-converter warps have read the raw tile (.ss: one tcgen05.commit)
    }
    for (int s = 0; s < ST4; ++s) {
      mbar_init(a_cvt(s), 4);    // ... have written hi | lo into tensor memory
      mbar_init(t_empty(s), 1);
    }
    for (int s = 0; s < SB4; ++s) {
      mbar_init(bh_full(s), 1);
      mbar_init(bh_empty(s), 1);
      mbar_init(bl_full(s), 1);
      mbar_init(bl_empty(s), 1);
    }
    mbar_init(accf_bar, 1);
    mbar_init(acce_bar, EPI4_WARPS);
    for (int s = 0; s < SCHED4; ++s) {
      mbar_init(sched_full(s), 1);
      mbar_init(sched_empty(s), SCHED4_CONSUMERS);
    }
    mbar_init_fence();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(tmem_slot) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - base));
  TC4_TRACE(2, 0);

  const int nkb = p.kb_total;
  // work item -> tile; n fastest so that concurrent CTAs share the A row panel in L2
  auto item_coords = [&](int item, int& m0, int& n0) {
    m0 = (item / w.tiles_n) * TBM;
    n0 = (item % w.tiles_n) * NT4;
  };
  // next tile of this CTA from the scheduler's ring (-1: no more); `sn` = tiles this role has taken so far.  Called by one
  // thread (producers, MMA issuer) or by all lanes of a warp (converters, epilogue): one arrival per role / warp.
  auto next_item = [&](uint32_t& sn, bool whole_warp) -> int {
    const int slot = sn % SCHED4;
    mbar_wait(sched_full(slot), (sn / SCHED4) & 1u);
    const int t = *reinterpret_cast<volatile int*>(smem_gen + (sched_tile - base) + 4 * slot);
    if (whole_warp) __syncwarp();
    if (!whole_warp || lane == 0) mbar_arrive(sched_empty(slot));
    ++sn;
    return t;
  };

  // registers: 512 x 128 at launch; producers / MMA issuer / allocator (warps 0-3) and the converters give theirs to the
  // epilogue warps (setmaxnreg at the head of every role's code)
  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
  if (warp == 0) {
    // ===================== A producer =====================
    if (lane == 0) {
      uint32_t it = 0, sn = 0;
      const uint32_t a_bytes = p.a_bf16 ? TILE_BYTES / 2 : TILE_BYTES;
      for (int item = next_item(sn, false); item >= 0; item = next_item(sn, false)) {
        int m0, n0;
        item_coords(item, m0, n0);
        for (int i = 0; i < nkb; ++i, ++it) {
          const int s = it % SA4, ph = (it / SA4) & 1;
          mbar_wait(a_free(s), ph ^ 1);
          TC4_TRACE(10, it);
          const uint32_t a_dst = a_base + s * TILE_BYTES;
          mbar_arrive_expect_tx(a_full(s), a_bytes);
          const int k0 = i * TBK;
          if (!p.a_mn) {
            tma_load_2d(a_dst, &tma_a, a_full(s), k0, m0);   // fp32: 128-byte rows (swizzled); bf16: 64-byte rows
          } else {
            const uint32_t box_bytes = p.a_bf16 ? 2048u : 4096u;
#pragma unroll
            for (int j = 0; j < 4; ++j) tma_load_2d(a_dst + j * box_bytes, &tma_a, a_full(s), m0 + j * 32, k0);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 3) {
    // ===================== B producer: hi and lo planes, no conversion =====================
    if (lane == 0) {
      uint32_t it = 0, sn = 0;
      for (int item = next_item(sn, false); item >= 0; item = next_item(sn, false)) {
        int m0, n0;
        item_coords(item, m0, n0);
        for (int i = 0; i < nkb; ++i, ++it) {
          const int s = it % SB4, ph = (it / SB4) & 1;
          const uint32_t b_dst = b_base + s * B4_STAGE;
          const int k0 = i * TBK;
          mbar_wait(bh_empty(s), ph ^ 1);
          TC4_TRACE(15, it);
          mbar_arrive_expect_tx(bh_full(s), B4_HALF);
          if (!p.b_mn) {
            tma_load_2d(b_dst, &tma_bh, bh_full(s), k0, n0);
          } else {
#pragma unroll
            for (int j = 0; j < NT4 / 32; ++j) tma_load_2d(b_dst + j * 4096, &tma_bh, bh_full(s), n0 + j * 32, k0);
          }
          if (!p.single_pass) {
            mbar_wait(bl_empty(s), ph ^ 1);
            mbar_arrive_expect_tx(bl_full(s), B4_HALF);
            if (!p.b_mn) {
              tma_load_2d(b_dst + B4_HALF, &tma_bl, bl_full(s), k0, n0);
            } else {
#pragma unroll
              for (int j = 0; j < NT4 / 32; ++j)
                tma_load_2d(b_dst + B4_HALF + j * 4096, &tma_bl, bl_full(s), n0 + j * 32, k0);
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      // D=f32, A=B=tf32, A from TMEM (K-major by construction), B major from the operand, N = 256
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)p.b_mn << 16) |
                             ((uint32_t)(NT4 >> 3) << 17) | ((uint32_t)(TBM >> 4) << 24);
      const uint32_t b_lbo = p.b_mn ? 4096u : 16u, b_kstep = p.b_mn ? 1024u : 32u;
      const uint32_t b_sbo = p.b_mn ? 512u : 1024u, b_lt = p.b_mn ? 1u : 2u;
      uint32_t it = 0, n = 0, sn = 0;
      for (int item = next_item(sn, false); item >= 0; item = next_item(sn, false), ++n) {
        mbar_wait(acce_bar, (n & 1u) ^ 1u);  // the epilogue has drained the accumulator of the previous tile
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        TC4_TRACE(19, n);
        for (int i = 0; i < nkb; ++i, ++it) {
          const int sa = it % ST4, pa = (it / ST4) & 1;
          const int sb = it % SB4, pb = (it / SB4) & 1;
          const int sr = it % SA4, pr = (it / SA4) & 1;   // raw A stage (.ss mode)
          if (ss) mbar_wait(a_full(sr), pr);
          else mbar_wait(a_cvt(sa), pa);
          mbar_wait(bh_full(sb), pb);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          TC4_TRACE(20, it);
          const uint32_t b_hi = b_base + sb * B4_STAGE, b_lo = b_hi + B4_HALF;
          const uint32_t ta_hi = tmem_base + 256u + sa * 64, ta_lo = ta_hi + 32;
          const bool lo_a = !p.single_pass && !p.a_bf16;   // A has a lo part (bfloat16 A is exact in tf32)
#pragma unroll
          for (int k = 0; k < TBK / 8; ++k) {
            const uint64_t dbh = make_smem_desc(b_hi + k * b_kstep, b_lbo, b_sbo, b_lt);
            if (ss) {
              umma_tf32(tmem_base, make_smem_desc(a_base + sr * TILE_BYTES + k * 32u, 16u, 1024u, 2u), dbh, idesc,
                        (i > 0 || k > 0) ? 1u : 0u);
            } else if (lo_a) {
              umma_tf32_ts(tmem_base, ta_lo + k * 8, dbh, idesc, (i > 0 || k > 0) ? 1u : 0u);
              umma_tf32_ts(tmem_base, ta_hi + k * 8, dbh, idesc, 1u);
            } else {
              umma_tf32_ts(tmem_base, ta_hi + k * 8, dbh, idesc, (i > 0 || k > 0) ? 1u : 0u);
            }
          }
          umma_commit(bh_empty(sb));
          if (!p.single_pass) {
            mbar_wait(bl_full(sb), pb);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int k = 0; k < TBK / 8; ++k) {
              const uint64_t dbl = make_smem_desc(b_lo + k * b_kstep, b_lbo, b_sbo, b_lt);
              umma_tf32_ts(tmem_base, ta_hi + k * 8, dbl, idesc, 1u);
            }
            umma_commit(bl_empty(sb));
          }
          umma_commit(ss ? a_free(sr) : t_empty(sa));
          TC4_TRACE(21, it);
        }
        umma_commit(accf_bar);
      }
    }
    __syncwarp();
  } else {
    // ===================== tile scheduler (warp 2) =====================
    if (lane == 0) {
      int t = blockIdx.x;   // the first tile is static (grid <= tiles), the rest come from the global counter
      for (uint32_t sn = 0;; ++sn) {
        const int slot = sn % SCHED4;
        mbar_wait(sched_empty(slot), ((sn / SCHED4) & 1u) ^ 1u);
        *reinterpret_cast<volatile int*>(smem_gen + (sched_tile - base) + 4 * slot) = t < w.items ? t : -1;
        mbar_arrive(sched_full(slot));
        if (t >= w.items) break;
        t = (int)gridDim.x + (int)atomicAdd(&w.counter[0], 1u);
      }
    }
    __syncwarp();
  }
  } else if (warp < 8) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 104;");
    // ===================== A converter: smem (raw) -> registers -> TMEM (hi | lo) =====================
    const int q = warp & 3;
    const int row = q * 32 + lane;  // row of the tile == TMEM lane
    uint32_t it = 0, sn = 0;
    for (int item = next_item(sn, true); item >= 0; item = next_item(sn, true)) {
      if (ss) continue;   // nothing to convert: the role only keeps the tile ring moving
      for (int i = 0; i < nkb; ++i, ++it) {
        const int s = it % SA4, ph = (it / SA4) & 1;
        const int ts = it % ST4, tph = (it / ST4) & 1;
        mbar_wait(a_full(s), ph);
        TC4_TRACE(30, it);
        const uint8_t* at = smem_gen + s * TILE_BYTES;
        uint32_t hi[32], lo[32];
        if (p.a_bf16) {
          // bfloat16 tile (no swizzle): widening to fp32 is exact and fits tf32, so there is no lo part
          if (!p.a_mn) {   // K-major: row r = 32 k x 2 bytes
            const uint4* rp = reinterpret_cast<const uint4*>(at + row * 64);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const uint4 v = rp[c];
              const uint32_t wv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                hi[c * 8 + 2 * e] = wv[e] << 16;
                hi[c * 8 + 2 * e + 1] = wv[e] & 0xFFFF0000u;
              }
            }
          } else {         // MN-major: four boxes of [32 k][32 m] bfloat16
            const unsigned short* cp = reinterpret_cast<const unsigned short*>(at + (row >> 5) * 2048) + (row & 31);
#pragma unroll
            for (int k = 0; k < 32; ++k) hi[k] = (uint32_t)cp[k * 32] << 16;
          }
#pragma unroll
          for (int k = 0; k < 32; ++k) lo[k] = 0u;
        } else if (!p.a_mn) {
          // K-major tile: row r = 128 bytes, 16-byte chunk c stored at chunk (c ^ (r & 7))  [SWIZZLE_128B]
          const uint8_t* rp = at + row * 128;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float4 v = *reinterpret_cast<const float4*>(rp + ((c ^ (row & 7)) << 4));
            const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              hi[c * 4 + e] = tf32_rna(vv[e]);
              lo[c * 4 + e] = __float_as_uint(vv[e] - __uint_as_float(hi[c * 4 + e]));
            }
          }
        } else {
          // MN-major tile: four boxes of [32 k][32 m] fp32, no swizzle; lanes read consecutive m
          const float* cp = reinterpret_cast<const float*>(at + (row >> 5) * 4096) + (row & 31);
#pragma unroll
          for (int k = 0; k < 32; ++k) {
            const float v = cp[k * 32];
            hi[k] = tf32_rna(v);
            lo[k] = __float_as_uint(v - __uint_as_float(hi[k]));
          }
        }
        __syncwarp();                                   // every lane holds its row in registers:
        if (lane == 0) mbar_arrive(a_free(s));          // the shared-memory stage can be refilled
        mbar_wait(t_empty(ts), tph ^ 1);                // the MMAs that read this tensor-memory stage have completed
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + 256u + ts * 64;
        tmem_st32(taddr, hi);
        if (!p.single_pass && !p.a_bf16) tmem_st32(taddr + 32, lo);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(a_cvt(ts));
        TC4_TRACE(31, it);
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 168;");
    // ===================== epilogue: TMEM -> registers (whole accumulator, then release it) -> global ============
    const int q = warp & 3;
    const int half = (warp - 8) >> 2;   // columns [128 half, 128 half + 128)
    float* ebuf = reinterpret_cast<float*>(smem_gen + (epi_base - base)) + (warp - 8) * 32 * EPI4_LD;
    uint32_t n = 0, sn = 0;
    for (int item = next_item(sn, true); item >= 0; item = next_item(sn, true), ++n) {
      int m0, n0;
      item_coords(item, m0, n0);
      TC4_TRACE(50, n);
      mbar_wait(accf_bar, n & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      TC4_TRACE(51, n);
      uint32_t r[4][32];
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * 128 + cc * 32);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[cc][0]), "=r"(r[cc][1]), "=r"(r[cc][2]), "=r"(r[cc][3]), "=r"(r[cc][4]), "=r"(r[cc][5]),
              "=r"(r[cc][6]), "=r"(r[cc][7]), "=r"(r[cc][8]), "=r"(r[cc][9]), "=r"(r[cc][10]), "=r"(r[cc][11]),
              "=r"(r[cc][12]), "=r"(r[cc][13]), "=r"(r[cc][14]), "=r"(r[cc][15]), "=r"(r[cc][16]), "=r"(r[cc][17]),
              "=r"(r[cc][18]), "=r"(r[cc][19]), "=r"(r[cc][20]), "=r"(r[cc][21]), "=r"(r[cc][22]), "=r"(r[cc][23]),
              "=r"(r[cc][24]), "=r"(r[cc][25]), "=r"(r[cc][26]), "=r"(r[cc][27]), "=r"(r[cc][28]), "=r"(r[cc][29]),
              "=r"(r[cc][30]), "=r"(r[cc][31])
            : "r"(taddr)
            : "memory");
      }
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(acce_bar);   // the MMA issuer may start the next tile
      TC4_TRACE(52, n);
      // out through a private [32][36] transpose buffer, 32 columns at a time: every store instruction writes four
      // full 128-byte row segments
      const int nb = n0 + half * 128;
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        float* mine = ebuf + lane * EPI4_LD;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<uint4*>(mine + 4 * j) =
              make_uint4(r[cc][4 * j], r[cc][4 * j + 1], r[cc][4 * j + 2], r[cc][4 * j + 3]);
        __syncwarp();
        const int nn = nb + cc * 32 + (lane & 7) * 4;
        float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.bias && nn < p.N) bias4 = __ldg(reinterpret_cast<const float4*>(p.bias + nn));
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int rr = j * 4 + (lane >> 3);
          const int m = m0 + q * 32 + rr;
          if (m < p.M && nn < p.N) {
            float4 v = *reinterpret_cast<const float4*>(ebuf + rr * EPI4_LD + (lane & 7) * 4);
            v.x += bias4.x; v.y += bias4.y; v.z += bias4.z; v.w += bias4.w;
            const int rowo = p.deint_H > 0 ? ((m & 3) * p.deint_H + (m >> 2)) : m;
            if (p.c_bf16) {
              *reinterpret_cast<uint2*>(reinterpret_cast<unsigned short*>(p.c) + (long long)rowo * p.ldc + nn) =
                  pack_bf16x4(v.x, v.y, v.z, v.w);
            } else {
              float4* o = reinterpret_cast<float4*>(p.c + (long long)rowo * p.ldc + nn);
              if (p.accumulate) {
                const float4 old = *o;
                v.x += old.x; v.y += old.y; v.z += old.z; v.w += old.w;
              }
              *o = v;
            }
          }
        }
        __syncwarp();
      }
      TC4_TRACE(53, n);
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
  // the last CTA to finish leaves the counters at zero for the next launch that uses this slot (every CTA has taken its
  // "no more tiles" answer by now, so nobody touches counter[0] any more)
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(&w.counter[1], 1u) == gridDim.x - 1) {
      w.counter[0] = 0u;
      w.counter[1] = 0u;
      __threadfence();
    }
  }
}

// tile counters: zero at module load, left at zero by every launch; consecutive launches take consecutive slots, so two
// launches share a slot only when they are TC4_SLOTS launches apart (they cannot overlap in time)
constexpr int TC4_SLOTS = 1024;
__device__ unsigned int g_tc4_counters[2 * TC4_SLOTS];

// host side ------------------------------------------------------------------------------------
int make_tc_map(CUtensorMap* map, const float* ptr, long long s_r, long long s_k, int rows, int K, int* mn_major,
                int a_through_tmem, int bf16, int box_rows);
bool gemm_tc_supported(const GemmArgs& g);

// B given as pre-split tf32 planes, more than one 128-column tile wide, K a whole number of k-blocks is NOT required
// (TMA zero-fills), no split-K (weight GEMMs have short K)
bool gemm_tc4_supported(const GemmArgs& g) {
  if (!g.b_hi || !g.b_lo || g.N <= 128) return false;
  if (g.M < 4096) return false;   // a few tiles only: the one-tile-per-CTA kernel starts and drains faster (measured)
  if ((((uintptr_t)g.b_hi) | ((uintptr_t)g.b_lo)) & 15) return false;
  return gemm_tc_supported(g);
}

int gemm_tc4(const GemmArgs& g, cudaStream_t stream) {
  static int sms = 0;
  if (sms == 0) {
    MRG_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM4_BYTES));
    int dev0 = 0;
    MRG_CUDA_CHECK(cudaGetDevice(&dev0));
    MRG_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev0));
  }
  CUtensorMap ma, mbh, mbl;
  TcParams p = {};
  int b_mn2 = 0;
  if (int e = make_tc_map(&ma, g.a, g.a_sm, g.a_sk, g.M, g.K, &p.a_mn, 1, g.a_bf16, TBM)) return e;
  if (int e = make_tc_map(&mbh, g.b_hi, g.b_sn, g.b_sk, g.N, g.K, &p.b_mn, 0, 0, NT4)) return e;
  if (int e = make_tc_map(&mbl, g.b_lo, g.b_sn, g.b_sk, g.N, g.K, &b_mn2, 0, 0, NT4)) return e;
  p.a_bf16 = g.a_bf16; p.c_bf16 = g.c_bf16;
  p.M = g.M; p.N = g.N; p.K = g.K;
  p.kb_total = (g.K + TBK - 1) / TBK;
  p.kb_per_split = p.kb_total;
  p.c = g.c; p.ldc = g.ldc; p.bias = g.bias; p.accumulate = g.accumulate; p.deint_H = g.row_deinterleave_H;
  p.partial = nullptr;
  p.single_pass = g.single_pass;
  p.trace = debug_trace_buffer();
  Tc4Work w;
  w.tiles_n = (g.N + NT4 - 1) / NT4;
  w.items = w.tiles_n * ((g.M + TBM - 1) / TBM);
  static unsigned int* counters_of[64] = {};   // the symbol has one instance per device
  static std::atomic<unsigned int> seq{0};
  int dev = 0;
  MRG_CUDA_CHECK(cudaGetDevice(&dev));
  MRG_REQUIRE(dev >= 0 && dev < 64, "gemm_tc4: device index %d out of range", dev);
  if (!counters_of[dev]) MRG_CUDA_CHECK(cudaGetSymbolAddress((void**)&counters_of[dev], g_tc4_counters));
  w.counter = counters_of[dev] + 2 * (seq.fetch_add(1) % TC4_SLOTS);
  ProfScope prof(PROF_GEMM, stream);
  count_launch();
  gemm_tc4_kernel<<<w.items < sms ? w.items : sms, TC4_THREADS, SMEM4_BYTES, stream>>>(ma, mbh, mbl, p, w);
  MRG_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace mrg
