// Projection GEMM, third generation: PERSISTENT tiles with a double-buffered accumulator.
//
// The second-generation kernel (mrg_gemm_tc2.cu) runs one 128x128 tile per CTA: TMA fill, converter latency, the MMA
// main loop and the epilogue (TMEM -> registers -> shared memory -> global) are serial inside a CTA, and with K = 256
// (8 k-blocks) the main loop is only half of a CTA's life — two co-resident CTAs per SM hide part of that.  Here one
// CTA per SM walks a list of work items (tile x split-K slice), the operand pipeline (4 stages) never drains between
// items, and the accumulator lives in TWO tensor-memory buffers: while the four epilogue warps drain item n, the MMA
// issuer is already accumulating item n+1.  tcgen05 / TMA / 3xTF32 arithmetic, operand layouts and the A-through-TMEM
// converter are those of the second generation (bit-identical results).
//
// Warp roles (512 threads): 0 TMA producer | 1 MMA issuer | 2 TMEM allocator | 3 idle | 4-7 A converter |
// 8-11 B converter | 12-15 epilogue (TMEM lane quarter = warp % 4; 32x32 chunks transposed through a private
// shared-memory buffer so that every global store instruction writes four full 128-byte row segments).
// TMEM: accumulators at columns [0,128) and [128,256), operand stage s at 256 + 64 s (hi | lo).
// Barriers: full[s] TMA -> converters, cvt[s] converters -> MMA, empty[s] MMA -> TMA, acc_full[b] MMA -> epilogue,
// acc_empty[b] epilogue -> MMA.
#include <cstdlib>

#include "mrg_tc_common.cuh"

namespace mrg {

constexpr int S3 = 4;                                   // operand pipeline stages
constexpr int STAGE3_BYTES = 3 * TILE_BYTES;            // A raw, B_hi, B_lo
constexpr int EPI_LD = 36;                              // floats per row of an epilogue staging buffer
constexpr int EPI_BYTES = 4 * 32 * EPI_LD * 4;          // one [32][36] buffer per epilogue warp
constexpr int SMEM3_BYTES = S3 * STAGE3_BYTES + EPI_BYTES + 1024 + 256;
constexpr int TC3_THREADS = 512;

struct Tc3Work {
  int tiles_n, n_tiles, items;  // items = n_tiles * splits
};

__global__ void __launch_bounds__(TC3_THREADS, 1)
gemm_tc3_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b, TcParams p,
                Tc3Work w) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t epi_base = base + S3 * STAGE3_BYTES;
  const uint32_t bar_base = epi_base + EPI_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto cvt_bar = [&](int s) { return bar_base + 8u * (S3 + s); };
  auto empty_bar = [&](int s) { return bar_base + 8u * (2 * S3 + s); };
  auto accf_bar = [&](int b) { return bar_base + 8u * (3 * S3 + b); };
  auto acce_bar = [&](int b) { return bar_base + 8u * (3 * S3 + 2 + b); };
  const uint32_t tmem_slot = bar_base + 8u * (3 * S3 + 4);
  uint8_t* smem_gen = smem_raw + (base - smem_u32(smem_raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S3; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(cvt_bar(s), 8);  // 4 A-converter warps + 4 B-converter warps
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(accf_bar(b), 1);
      mbar_init(acce_bar(b), 4);  // the four epilogue warps
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

  // work item -> (split-K slice, tile row, tile column); n fastest so that concurrent CTAs share the A row panel in L2
  auto item_coords = [&](int item, int& m0, int& n0, int& z, int& kb_begin, int& nkb) {
    z = item / w.n_tiles;
    const int tile = item % w.n_tiles;
    m0 = (tile / w.tiles_n) * TBM;
    n0 = (tile % w.tiles_n) * TBN;
    kb_begin = z * p.kb_per_split;
    nkb = min(p.kb_total, kb_begin + p.kb_per_split) - kb_begin;
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t it = 0;
      for (int item = blockIdx.x; item < w.items; item += gridDim.x) {
        int m0, n0, z, kb_begin, nkb;
        item_coords(item, m0, n0, z, kb_begin, nkb);
        for (int i = 0; i < nkb; ++i, ++it) {
          const int s = it % S3, ph = (it / S3) & 1;
          mbar_wait(empty_bar(s), ph ^ 1);
          const uint32_t a_dst = base + s * STAGE3_BYTES;
          const uint32_t b_dst = a_dst + TILE_BYTES;
          mbar_arrive_expect_tx(full_bar(s), 2 * TILE_BYTES);
          const int k0 = (kb_begin + i) * TBK;
          if (!p.a_mn) {
            tma_load_2d(a_dst, &tma_a, full_bar(s), k0, m0);
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) tma_load_2d(a_dst + j * 4096, &tma_a, full_bar(s), m0 + j * 32, k0);
          }
          if (!p.b_mn) {
            tma_load_2d(b_dst, &tma_b, full_bar(s), k0, n0);
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) tma_load_2d(b_dst + j * 4096, &tma_b, full_bar(s), n0 + j * 32, k0);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)p.b_mn << 16) |
                             ((uint32_t)(TBN >> 3) << 17) | ((uint32_t)(TBM >> 4) << 24);
      const uint32_t b_lbo = p.b_mn ? 4096u : 16u, b_kstep = p.b_mn ? 1024u : 32u;
      const uint32_t b_sbo = p.b_mn ? 512u : 1024u, b_lt = p.b_mn ? 1u : 2u;
      uint32_t it = 0, n = 0;
      for (int item = blockIdx.x; item < w.items; item += gridDim.x, ++n) {
        int m0, n0, z, kb_begin, nkb;
        item_coords(item, m0, n0, z, kb_begin, nkb);
        const uint32_t buf = n & 1u;
        mbar_wait(acce_bar(buf), ((n >> 1) & 1u) ^ 1u);  // the epilogue has drained this accumulator
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tacc = tmem_base + buf * 128u;
        for (int i = 0; i < nkb; ++i, ++it) {
          const int s = it % S3, ph = (it / S3) & 1;
          mbar_wait(cvt_bar(s), ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t b_hi = base + s * STAGE3_BYTES + TILE_BYTES, b_lo = b_hi + TILE_BYTES;
          const uint32_t ta_hi = tmem_base + 256u + s * 64, ta_lo = ta_hi + 32;
#pragma unroll
          for (int k = 0; k < TBK / 8; ++k) {
            const uint64_t dbh = make_smem_desc(b_hi + k * b_kstep, b_lbo, b_sbo, b_lt);
            const uint64_t dbl = make_smem_desc(b_lo + k * b_kstep, b_lbo, b_sbo, b_lt);
            if (p.single_pass) {
              umma_tf32_ts(tacc, ta_hi + k * 8, dbh, idesc, (i > 0 || k > 0) ? 1u : 0u);
            } else {
              umma_tf32_ts(tacc, ta_lo + k * 8, dbh, idesc, (i > 0 || k > 0) ? 1u : 0u);
              umma_tf32_ts(tacc, ta_hi + k * 8, dbl, idesc, 1u);
              umma_tf32_ts(tacc, ta_hi + k * 8, dbh, idesc, 1u);
            }
          }
          umma_commit(empty_bar(s));
        }
        umma_commit(accf_bar(buf));
      }
    }
    __syncwarp();
  } else if (warp >= 4 && warp < 8) {
    // ===================== A converter: smem (raw fp32) -> registers -> TMEM (hi | lo) =====================
    const int q = warp & 3;
    const int row = q * 32 + lane;  // row of the tile == TMEM lane
    uint32_t it = 0;
    for (int item = blockIdx.x; item < w.items; item += gridDim.x) {
      int m0, n0, z, kb_begin, nkb;
      item_coords(item, m0, n0, z, kb_begin, nkb);
      for (int i = 0; i < nkb; ++i, ++it) {
        const int s = it % S3, ph = (it / S3) & 1;
        mbar_wait(full_bar(s), ph);
        const uint8_t* at = smem_gen + s * STAGE3_BYTES;
        uint32_t hi[32], lo[32];
        if (!p.a_mn) {
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
          const float* cp = reinterpret_cast<const float*>(at + (row >> 5) * 4096) + (row & 31);
#pragma unroll
          for (int k = 0; k < 32; ++k) {
            const float v = cp[k * 32];
            hi[k] = tf32_rna(v);
            lo[k] = __float_as_uint(v - __uint_as_float(hi[k]));
          }
        }
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + 256u + s * 64;
        tmem_st32(taddr, hi);
        if (!p.single_pass) tmem_st32(taddr + 32, lo);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(cvt_bar(s));
      }
    }
  } else if (warp >= 8 && warp < 12) {
    // ===================== B converter: raw tile -> hi (in place) / lo tiles in shared memory ===========
    const int ct = threadIdx.x - 256;  // 0..127
    uint32_t it = 0;
    for (int item = blockIdx.x; item < w.items; item += gridDim.x) {
      int m0, n0, z, kb_begin, nkb;
      item_coords(item, m0, n0, z, kb_begin, nkb);
      for (int i = 0; i < nkb; ++i, ++it) {
        const int s = it % S3, ph = (it / S3) & 1;
        mbar_wait(full_bar(s), ph);
        float4* hi = reinterpret_cast<float4*>(smem_gen + s * STAGE3_BYTES + TILE_BYTES);
        float4* lo = reinterpret_cast<float4*>(smem_gen + s * STAGE3_BYTES + 2 * TILE_BYTES);
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
      }
    }
  } else if (warp >= 12) {
    // ===================== epilogue: TMEM -> registers -> private smem transpose -> global =====================
    const int q = warp & 3;
    float* ebuf = reinterpret_cast<float*>(smem_gen + (epi_base - base)) + q * 32 * EPI_LD;
    uint32_t n = 0;
    for (int item = blockIdx.x; item < w.items; item += gridDim.x, ++n) {
      int m0, n0, z, kb_begin, nkb;
      item_coords(item, m0, n0, z, kb_begin, nkb);
      const uint32_t buf = n & 1u;
      mbar_wait(accf_bar(buf), (n >> 1) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
      for (int cc = 0; cc < TBN / 32; ++cc) {
        uint32_t r[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * 128u + (uint32_t)(cc * 32);
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
        if (cc == TBN / 32 - 1) {  // the whole accumulator is in registers / on its way out: hand the buffer back
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(acce_bar(buf));
        }
        float* mine = ebuf + lane * EPI_LD;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<uint4*>(mine + 4 * j) = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
        __syncwarp();
        const int nn = n0 + cc * 32 + (lane & 7) * 4;
        float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.bias && !p.partial && nn < p.N) bias4 = *reinterpret_cast<const float4*>(p.bias + nn);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int rr = j * 4 + (lane >> 3);
          const int m = m0 + q * 32 + rr;
          if (m < p.M && nn < p.N) {
            float4 v = *reinterpret_cast<const float4*>(ebuf + rr * EPI_LD + (lane & 7) * 4);
            if (p.partial) {
              *reinterpret_cast<float4*>(p.partial + ((size_t)z * p.M + m) * p.N + nn) = v;
            } else {
              v.x += bias4.x; v.y += bias4.y; v.z += bias4.z; v.w += bias4.w;
              const int rowo = p.deint_H > 0 ? ((m & 3) * p.deint_H + (m >> 2)) : m;
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
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

// host side ------------------------------------------------------------------------------------
int make_tc_map(CUtensorMap* map, const float* ptr, long long s_r, long long s_k, int rows, int K, int* mn_major,
                int a_through_tmem);
int tc_splits(int M, int N, int K);
__global__ void tc_splitk_reduce_kernel(TcParams p, int splits);

int gemm_tc3(const GemmArgs& g, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  static int sms = 0;
  if (sms == 0) {
    MRG_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM3_BYTES));
    int dev = 0;
    MRG_CUDA_CHECK(cudaGetDevice(&dev));
    MRG_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  }
  CUtensorMap ma, mb;
  TcParams p = {};
  if (int e = make_tc_map(&ma, g.a, g.a_sm, g.a_sk, g.M, g.K, &p.a_mn, 1)) return e;
  if (int e = make_tc_map(&mb, g.b, g.b_sn, g.b_sk, g.N, g.K, &p.b_mn, 0)) return e;
  p.M = g.M; p.N = g.N; p.K = g.K;
  p.kb_total = (g.K + TBK - 1) / TBK;
  const int splits = tc_splits(g.M, g.N, g.K);
  p.kb_per_split = (p.kb_total + splits - 1) / splits;
  const int zdim = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;
  p.c = g.c; p.ldc = g.ldc; p.bias = g.bias; p.accumulate = g.accumulate; p.deint_H = g.row_deinterleave_H;
  p.partial = nullptr;
  p.single_pass = g.single_pass;
  if (zdim > 1) {
    const size_t need = (size_t)zdim * g.M * g.N * sizeof(float);
    if (workspace == nullptr || workspace_bytes < need) {
      set_error("gemm_tc3: workspace too small (%zu needed)", need);
      return MRG_E_WORKSPACE;
    }
    p.partial = (float*)workspace;
  }
  Tc3Work w;
  w.tiles_n = (g.N + TBN - 1) / TBN;
  w.n_tiles = w.tiles_n * ((g.M + TBM - 1) / TBM);
  w.items = w.n_tiles * zdim;
  ProfScope prof(PROF_GEMM, stream);
  count_launch(zdim > 1 ? 2 : 1);
  gemm_tc3_kernel<<<w.items < sms ? w.items : sms, TC3_THREADS, SMEM3_BYTES, stream>>>(ma, mb, p, w);
  MRG_CUDA_CHECK(cudaGetLastError());
  if (zdim > 1) {
    const long long total = (long long)g.M * g.N;
    tc_splitk_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(p, zdim);
    MRG_CUDA_CHECK(cudaGetLastError());
  }
  return 0;
}

}  // namespace mrg
