// Projection GEMM, second generation: the A operand goes through TENSOR MEMORY.
//
// ncu on the first kernel (profiles/r1_rec_gemm_full.md) shows the tensor pipe 33-49 % active: with
// both operands in shared memory every one of the three 128x128x8 tf32 MMAs of the 3xTF32 split re-reads
// 8 KB of operands in ~69 cycles — the full 128 B/clk of the shared-memory port — while the converter
// warps and the TMA engine need the same port.  Here the A tile is read from shared memory ONCE per
// k-block by the converter warps (thread = row), split into hi / lo in registers and written with
// tcgen05.st into TMEM (lane = row, column = k); the MMAs take A from TMEM (.ts form) and only B from
// shared memory.  Shared-memory bytes per 128x128x32 k-block: 224 KB -> 144 KB, and because the converter
// reads the raw tile element-wise it also handles MN-major A (weight-gradient GEMMs) without the
// 32-byte-atom swizzle.  TMEM budget: 128 accumulator columns + STAGES x (32 hi + 32 lo) columns.
//
// Warp roles: 0 TMA producer | 1 MMA issuer | 2 TMEM allocator | 3 idle | 4-7 A converter (TMEM lane
// quarter = warp % 4) | 8-11 B converter (hi / lo tiles in shared memory) | all 12: epilogue.
#include <cstdlib>

#include "mrg_tc_common.cuh"

namespace mrg {

#ifdef MRG_REC_TRACE
#define GEMM_TRACE_DECL unsigned trace_n = 0; const bool trace_cta = p.trace && blockIdx.x == 0 && blockIdx.z == 0 && blockIdx.y == gridDim.y / 2;
#define GEMM_TRACE(evt, i)                                                                                  \
  if (trace_cta && (threadIdx.x & 31) == 0 && trace_n < 1024u) {                                            \
    unsigned long long* tp = p.trace + ((size_t)(threadIdx.x >> 5) * 1024 + trace_n) * 2;                   \
    tp[0] = clock64();                                                                                      \
    tp[1] = ((unsigned long long)(threadIdx.x >> 5) << 48) | ((unsigned long long)(evt) << 32) | (unsigned long long)(i); \
    ++trace_n;                                                                                              \
  }
#else
#define GEMM_TRACE_DECL
#define GEMM_TRACE(evt, i)
#endif

constexpr int STAGE2_BYTES = 3 * TILE_BYTES;                // A raw, B_hi, B_lo
// S2 = pipeline stages: 4 with one CTA per SM (long K), or 2 with TWO co-resident CTAs per SM (short K: one CTA's
// prologue / epilogue overlaps the other's main loop; 2 x (128 accumulator + 2 x 64 operand) TMEM columns)
template <int S2>
constexpr int smem2_bytes() { return S2 * STAGE2_BYTES + 1024 + 256; }
constexpr int TC2_THREADS = 384;
constexpr int ACC_COLS = 128;

template <int S2>
__global__ void __launch_bounds__(TC2_THREADS, S2 == 2 ? 2 : 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b, TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = base + S2 * STAGE2_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto cvt_bar = [&](int s) { return bar_base + 8u * (S2 + s); };
  auto empty_bar = [&](int s) { return bar_base + 8u * (2 * S2 + s); };
  const uint32_t acc_bar = bar_base + 8u * (3 * S2);
  const uint32_t tmem_slot = bar_base + 8u * (3 * S2 + 1);
  uint8_t* smem_gen = smem_raw + (base - smem_u32(smem_raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * TBM, n0 = blockIdx.x * TBN;
  const int kb_begin = blockIdx.z * p.kb_per_split;
  const int kb_end = min(p.kb_total, kb_begin + p.kb_per_split);
  const int nkb = max(0, kb_end - kb_begin);
  GEMM_TRACE_DECL
  GEMM_TRACE(1, 0);

  if (threadIdx.x == 0) {
    for (int s = 0; s < S2; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(cvt_bar(s), 8);  // 4 A-converter warps + 4 B-converter warps
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(acc_bar, 1);
    mbar_init_fence();
  }
  if (warp == 2) {  // all 512 columns: 128 accumulator + 4 stages x 64 operand columns
    if (S2 == 2) asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(tmem_slot) : "memory");
    else asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(tmem_slot) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - base));
  GEMM_TRACE(2, 0);

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      for (int i = 0; i < nkb; ++i) {
        const int s = i % S2, ph = (i / S2) & 1;
        mbar_wait(empty_bar(s), ph ^ 1);
        GEMM_TRACE(10, i);
        const uint32_t a_dst = base + s * STAGE2_BYTES;
        const uint32_t b_dst = a_dst + TILE_BYTES;
        mbar_arrive_expect_tx(full_bar(s), (p.a_bf16 ? TILE_BYTES / 2 : TILE_BYTES) + TILE_BYTES);
        const int k0 = (kb_begin + i) * TBK;
        if (!p.a_mn) {
          tma_load_2d(a_dst, &tma_a, full_bar(s), k0, m0);   // fp32: 128-byte rows (swizzled); bf16: 64-byte rows
        } else {
          const uint32_t box_bytes = p.a_bf16 ? 2048u : 4096u;
#pragma unroll
          for (int j = 0; j < 4; ++j) tma_load_2d(a_dst + j * box_bytes, &tma_a, full_bar(s), m0 + j * 32, k0);
        }
        if (!p.b_mn) {
          tma_load_2d(b_dst, &tma_b, full_bar(s), k0, n0);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) tma_load_2d(b_dst + j * 4096, &tma_b, full_bar(s), n0 + j * 32, k0);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      // D=f32, A=B=tf32, A from TMEM (K-major by construction), B major from the operand
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)p.b_mn << 16) |
                             ((uint32_t)(TBN >> 3) << 17) | ((uint32_t)(TBM >> 4) << 24);
      const uint32_t b_lbo = p.b_mn ? 4096u : 16u, b_kstep = p.b_mn ? 1024u : 32u;
      const uint32_t b_sbo = p.b_mn ? 512u : 1024u, b_lt = p.b_mn ? 1u : 2u;
      for (int i = 0; i < nkb; ++i) {
        const int s = i % S2, ph = (i / S2) & 1;
        mbar_wait(cvt_bar(s), ph);
        GEMM_TRACE(20, i);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t b_hi = base + s * STAGE2_BYTES + TILE_BYTES, b_lo = b_hi + TILE_BYTES;
        const uint32_t ta_hi = tmem_base + ACC_COLS + s * 64, ta_lo = ta_hi + 32;
#pragma unroll
        for (int k = 0; k < TBK / 8; ++k) {
          const uint64_t dbh = make_smem_desc(b_hi + k * b_kstep, b_lbo, b_sbo, b_lt);
          const uint64_t dbl = make_smem_desc(b_lo + k * b_kstep, b_lbo, b_sbo, b_lt);
          if (p.single_pass) {
            umma_tf32_ts(tmem_base, ta_hi + k * 8, dbh, idesc, (i > 0 || k > 0) ? 1u : 0u);
          } else if (p.a_bf16) {   // A is exact in its hi part: A B = A B_lo + A B_hi
            umma_tf32_ts(tmem_base, ta_hi + k * 8, dbl, idesc, (i > 0 || k > 0) ? 1u : 0u);
            umma_tf32_ts(tmem_base, ta_hi + k * 8, dbh, idesc, 1u);
          } else {
            umma_tf32_ts(tmem_base, ta_lo + k * 8, dbh, idesc, (i > 0 || k > 0) ? 1u : 0u);
            umma_tf32_ts(tmem_base, ta_hi + k * 8, dbl, idesc, 1u);
            umma_tf32_ts(tmem_base, ta_hi + k * 8, dbh, idesc, 1u);
          }
        }
        umma_commit(empty_bar(s));
        GEMM_TRACE(21, i);
      }
      umma_commit(acc_bar);
    }
    __syncwarp();
  } else if (warp >= 4 && warp < 8) {
    // ===================== A converter: smem (raw fp32) -> registers -> TMEM (hi | lo) =====================
    const int q = warp & 3;
    const int row = q * 32 + lane;  // row of the tile == TMEM lane
    for (int i = 0; i < nkb; ++i) {
      const int s = i % S2, ph = (i / S2) & 1;
      mbar_wait(full_bar(s), ph);
      GEMM_TRACE(30, i);
      const uint8_t* at = smem_gen + s * STAGE2_BYTES;
      uint32_t hi[32], lo[32];
      if (p.a_bf16) {
        // bfloat16 tile (no swizzle): widening to fp32 is exact and fits tf32, so there is no lo part
        if (!p.a_mn) {   // K-major: row r = 32 k x 2 bytes
          const uint4* rp = reinterpret_cast<const uint4*>(at + row * 64);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const uint4 v = rp[c];
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              hi[c * 8 + 2 * e] = w[e] << 16;
              hi[c * 8 + 2 * e + 1] = w[e] & 0xFFFF0000u;
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
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + ACC_COLS + s * 64;
      tmem_st32(taddr, hi);
      if (!p.single_pass && !p.a_bf16) tmem_st32(taddr + 32, lo);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(cvt_bar(s));
      GEMM_TRACE(31, i);
    }
  } else if (warp >= 8) {
    // ===================== B converter: raw tile -> hi (in place) / lo tiles in shared memory ===========
    const int ct = threadIdx.x - 256;  // 0..127
    for (int i = 0; i < nkb; ++i) {
      const int s = i % S2, ph = (i / S2) & 1;
      mbar_wait(full_bar(s), ph);
      GEMM_TRACE(40, i);
      float4* hi = reinterpret_cast<float4*>(smem_gen + s * STAGE2_BYTES + TILE_BYTES);
      float4* lo = reinterpret_cast<float4*>(smem_gen + s * STAGE2_BYTES + 2 * TILE_BYTES);
#pragma unroll
      for (int j = 0; j < TILE_BYTES / 16 / 128; ++j) {
        const int e = ct + j * 128;
        const float4 v = hi[e];
        uint4 h, l;
        h.x = tf32_rna(v.x); h.y = tf32_rna(v.y); h.z = tf32_rna(v.z); h.w = tf32_rna(v.w);
        l.x = __float_as_uint(v.x - __uint_as_float(h.x));
        l.y = __float_as_uint(v.y - __uint_as_float(h.y));
        l.z = __float_as_uint(v.z - __uint_as_float(h.z));
        l.w = __float_as_uint(v.w - __uint_as_float(h.w));
        reinterpret_cast<uint4*>(hi)[e] = h;
        if (!p.single_pass) reinterpret_cast<uint4*>(lo)[e] = l;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(cvt_bar(s));
      GEMM_TRACE(41, i);
    }
  }

  // ===================== epilogue (all 12 warps) =====================
  {
    constexpr int CS = TBN + 4;
    float* csm = reinterpret_cast<float*>(smem_gen);
    const int q = warp & 3;
    const int half = warp >> 2;
    __syncwarp();
    GEMM_TRACE(50, 0);
    if (nkb > 0) {
      mbar_wait(acc_bar, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    GEMM_TRACE(51, 0);
#pragma unroll 1
    for (int cc = 0; cc < (half < 2 ? 2 : 0); ++cc) {
      uint32_t r[32];
      const int col0 = half * 64 + cc * 32;
      if (nkb > 0) {
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)col0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
              "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
              "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),
              "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]),
              "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr)
            : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = 0u;
      }
      float* dst = csm + (size_t)(q * 32 + lane) * CS + col0;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<uint4*>(dst + 4 * j) = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    GEMM_TRACE(52, 0);
    const int n = n0 + lane * 4;
    if (n < p.N) {
      float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p.bias && !p.partial) bias4 = *reinterpret_cast<const float4*>(p.bias + n);
#pragma unroll 4
      for (int rr = warp; rr < TBM; rr += TC2_THREADS / 32) {
        const int m = m0 + rr;
        if (m >= p.M) break;
        float4 v = *reinterpret_cast<const float4*>(csm + (size_t)rr * CS + lane * 4);
        if (p.partial) {
          *reinterpret_cast<float4*>(p.partial + ((size_t)blockIdx.z * p.M + m) * p.N + n) = v;
        } else {
          v.x += bias4.x; v.y += bias4.y; v.z += bias4.z; v.w += bias4.w;
          const int rowo = p.deint_H > 0 ? ((m & 3) * p.deint_H + (m >> 2)) : m;
          if (p.c_bf16) {
            *reinterpret_cast<uint2*>(reinterpret_cast<unsigned short*>(p.c) + (long long)rowo * p.ldc + n) =
                pack_bf16x4(v.x, v.y, v.z, v.w);
            continue;
          }
          float4* o = reinterpret_cast<float4*>(p.c + (long long)rowo * p.ldc + n);
          if (p.accumulate) {
            const float4 old = *o;
            v.x += old.x; v.y += old.y; v.z += old.z; v.w += old.w;
          }
          *o = v;
        }
      }
    }
  }
  GEMM_TRACE(53, 0);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (S2 == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem_base) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

// host side ------------------------------------------------------------------------------------
int make_tc_map(CUtensorMap* map, const float* ptr, long long s_r, long long s_k, int rows, int K, int* mn_major,
                int a_through_tmem, int bf16, int box_rows);
int tc_splits(int M, int N, int K);
__global__ void tc_splitk_reduce_kernel(TcParams p, int splits);

int gemm_tc2(const GemmArgs& g, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    MRG_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc2_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2_bytes<4>()));
    MRG_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc2_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2_bytes<2>()));
    attr_set = true;
  }
  CUtensorMap ma, mb;
  TcParams p = {};
  if (int e = make_tc_map(&ma, g.a, g.a_sm, g.a_sk, g.M, g.K, &p.a_mn, 1, g.a_bf16, TBM)) return e;
  if (int e = make_tc_map(&mb, g.b, g.b_sn, g.b_sk, g.N, g.K, &p.b_mn, 0, 0, TBN)) return e;
  p.a_bf16 = g.a_bf16; p.c_bf16 = g.c_bf16;
  p.M = g.M; p.N = g.N; p.K = g.K;
  p.kb_total = (g.K + TBK - 1) / TBK;
  const int splits = tc_splits(g.M, g.N, g.K);
  p.kb_per_split = (p.kb_total + splits - 1) / splits;
  const int zdim = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;
  p.c = g.c; p.ldc = g.ldc; p.bias = g.bias; p.accumulate = g.accumulate; p.deint_H = g.row_deinterleave_H;
  p.partial = nullptr;
  p.single_pass = g.single_pass;
  p.trace = debug_trace_buffer();
  if (zdim > 1) {
    const size_t need = (size_t)zdim * g.M * g.N * sizeof(float);
    if (workspace == nullptr || workspace_bytes < need) {
      set_error("gemm_tc2: workspace too small (%zu needed)", need);
      return MRG_E_WORKSPACE;
    }
    p.partial = (float*)workspace;
  }
  dim3 grid((g.N + TBN - 1) / TBN, (g.M + TBM - 1) / TBM, zdim);
  ProfScope prof(PROF_GEMM, stream);
  count_launch(zdim > 1 ? 2 : 1);
  static int force_stages = -1;
  if (force_stages < 0) {
    const char* e = getenv("MRG_GEMM_STAGES");
    force_stages = e ? atoi(e) : 0;
  }
  // short K (<= 16 k-blocks per CTA) or more CTAs than SMs: two co-resident CTAs per SM with 2 stages each
  const long long ctas = (long long)grid.x * grid.y * grid.z;
  const bool two = force_stages ? force_stages == 2 : (p.kb_per_split <= 16 || ctas > 148);
  if (two) gemm_tc2_kernel<2><<<grid, TC2_THREADS, smem2_bytes<2>(), stream>>>(ma, mb, p);
  else gemm_tc2_kernel<4><<<grid, TC2_THREADS, smem2_bytes<4>(), stream>>>(ma, mb, p);
  MRG_CUDA_CHECK(cudaGetLastError());
  if (zdim > 1) {
    const long long total = (long long)g.M * g.N;
    tc_splitk_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(p, zdim);
    MRG_CUDA_CHECK(cudaGetLastError());
  }
  return 0;
}

}  // namespace mrg
