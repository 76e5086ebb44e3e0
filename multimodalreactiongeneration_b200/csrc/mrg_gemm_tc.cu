// Time-parallel projection GEMMs on the 5th-generation tensor cores (sm_100a): tcgen05.mma kind::tf32
// with fp32-grade accuracy ("3xTF32": A*B ~= Ahi*Bhi + Alo*Bhi + Ahi*Blo, hi = rna_tf32(x),
// lo = rna_tf32(x - hi)), operands staged by TMA (cp.async.bulk.tensor, 128-byte swizzle), fp32
// accumulator in tensor memory, epilogue through tcgen05.ld with bias / accumulate / gate
// de-interleave / split-K partials.
//
// One 128x128 output tile per CTA (x split-K slices), BLOCK_K = 32 fp32 = one 128-byte swizzle row.
// Warp roles:  0 TMA producer | 1 MMA issuer (one elected thread) | 2 TMEM allocator | 3 idle |
//              4-11 converter (raw fp32 tile -> hi / lo TF32 tiles in shared memory); all 8 warps run the
//              epilogue (TMEM -> padded smem tile -> coalesced 512-byte row stores).
// Pipelines (mbarriers): full[s] TMA->converter, cvt[s] converter->MMA, empty[s] MMA->TMA
// (tcgen05.commit), acc_full MMA->epilogue.
//
// Either operand may be K-major (contraction index contiguous) or MN-major (row/column index
// contiguous, used by the weight-gradient GEMMs dW = dG^T X): the raw tile is position-preserving under
// the hi/lo split, so only the TMA box + swizzle mode (128B vs 128B_ATOM_32B), the shared-memory
// descriptor (layout type, LBO/SBO, K advance) and the major bits of the instruction descriptor change.
#include "mrg_tc_common.cuh"

namespace mrg {

constexpr int CVT_WARPS = 8;
constexpr int TC_THREADS = 128 + CVT_WARPS * 32;

__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b, TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = base + STAGES * STAGE_BYTES;
  // barrier layout (8 bytes each): full[STAGES], cvt[STAGES], empty[STAGES], acc_full, then tmem ptr
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto cvt_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto empty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
  const uint32_t acc_bar = bar_base + 8u * (3 * STAGES);
  const uint32_t tmem_slot = bar_base + 8u * (3 * STAGES + 1);
  uint8_t* smem_gen = smem_raw + (base - smem_u32(smem_raw));  // generic pointer to the aligned base

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * TBM, n0 = blockIdx.x * TBN;
  const int kb_begin = blockIdx.z * p.kb_per_split;
  const int kb_end = min(p.kb_total, kb_begin + p.kb_per_split);
  const int nkb = max(0, kb_end - kb_begin);

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(cvt_bar(s), CVT_WARPS);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(acc_bar, 1);
    mbar_init_fence();
  }
  if (warp == 2) {  // TMEM: 128 columns x 128 lanes of fp32 accumulator
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(tmem_slot) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - base));

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      for (int i = 0; i < nkb; ++i) {
        const int s = i % STAGES, ph = (i / STAGES) & 1;
        mbar_wait(empty_bar(s), ph ^ 1);
        const uint32_t a_dst = base + s * STAGE_BYTES;                   // A raw -> becomes A_hi
        const uint32_t b_dst = base + s * STAGE_BYTES + 2 * TILE_BYTES;  // B raw -> becomes B_hi
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
    __syncwarp();  // lanes 1-31 park here (no spinning) until the producer lane is done
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      // instruction descriptor: D=f32, A=B=tf32, majors, N>>3 at [17,23), M>>4 at [24,29)
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)p.a_mn << 15) |
                             ((uint32_t)p.b_mn << 16) | ((uint32_t)(TBN >> 3) << 17) |
                             ((uint32_t)(TBM >> 4) << 24);
      const uint32_t a_lbo = p.a_mn ? 4096u : 16u, b_lbo = p.b_mn ? 4096u : 16u;
      const uint32_t a_kstep = p.a_mn ? 1024u : 32u, b_kstep = p.b_mn ? 1024u : 32u;
      const uint32_t a_sbo = p.a_mn ? 512u : 1024u, b_sbo = p.b_mn ? 512u : 1024u;
      const uint32_t a_lt = p.a_mn ? 1u : 2u, b_lt = p.b_mn ? 1u : 2u;
      for (int i = 0; i < nkb; ++i) {
        const int s = i % STAGES, ph = (i / STAGES) & 1;
        mbar_wait(cvt_bar(s), ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a_hi = base + s * STAGE_BYTES, a_lo = a_hi + TILE_BYTES;
        const uint32_t b_hi = a_hi + 2 * TILE_BYTES, b_lo = a_hi + 3 * TILE_BYTES;
#pragma unroll
        for (int k = 0; k < TBK / 8; ++k) {
          const uint64_t dah = make_smem_desc(a_hi + k * a_kstep, a_lbo, a_sbo, a_lt);
          const uint64_t dal = make_smem_desc(a_lo + k * a_kstep, a_lbo, a_sbo, a_lt);
          const uint64_t dbh = make_smem_desc(b_hi + k * b_kstep, b_lbo, b_sbo, b_lt);
          const uint64_t dbl = make_smem_desc(b_lo + k * b_kstep, b_lbo, b_sbo, b_lt);
          if (p.single_pass) {
            umma_tf32(tmem_base, dah, dbh, idesc, (i > 0 || k > 0) ? 1u : 0u);
          } else {
            umma_tf32(tmem_base, dal, dbh, idesc, (i > 0 || k > 0) ? 1u : 0u);
            umma_tf32(tmem_base, dah, dbl, idesc, 1u);
            umma_tf32(tmem_base, dah, dbh, idesc, 1u);
          }
        }
        umma_commit(empty_bar(s));  // stage reusable once these MMAs have read it
      }
      umma_commit(acc_bar);         // accumulator complete
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ===================== converter, then epilogue =====================
    const int ct = threadIdx.x - 128;  // 0 .. CVT_WARPS*32-1
    for (int i = 0; i < nkb; ++i) {
      const int s = i % STAGES, ph = (i / STAGES) & 1;
      mbar_wait(full_bar(s), ph);
      uint8_t* st = smem_gen + s * STAGE_BYTES;
#pragma unroll
      for (int op = 0; op < 2; ++op) {
        float4* hi = reinterpret_cast<float4*>(st + op * 2 * TILE_BYTES);
        float4* lo = reinterpret_cast<float4*>(st + op * 2 * TILE_BYTES + TILE_BYTES);
#pragma unroll
        for (int j = 0; j < TILE_BYTES / 16 / (CVT_WARPS * 32); ++j) {
          const int e = ct + j * (CVT_WARPS * 32);
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
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes -> tensor-core reads
      __syncwarp();
      if (lane == 0) mbar_arrive(cvt_bar(s));
    }
  }

  // ===================== epilogue (all 8 warps) =====================
  // TMEM -> registers -> padded shared tile (conflict-free, reuses the now idle stage buffers)
  // -> fully coalesced 512-byte row stores with bias / accumulate / de-interleave.
  {
    constexpr int CS = TBN + 4;  // padded row stride (floats): 16-byte lanes of 8 rows cover all 32 banks
    float* csm = reinterpret_cast<float*>(smem_gen);
    const int q = warp & 3;      // TMEM lane quarter this warp may access
    const int half = warp >> 2;  // which 64 accumulator columns (warps 8+ only take part in the row stores)
    if (nkb > 0) {
      mbar_wait(acc_bar, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
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
    const int n = n0 + lane * 4;
    if (n < p.N) {  // N % 4 == 0 is a precondition of this path
      float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p.bias && !p.partial) bias4 = *reinterpret_cast<const float4*>(p.bias + n);
#pragma unroll 4
      for (int rr = warp; rr < TBM; rr += TC_THREADS / 32) {
        const int m = m0 + rr;
        if (m >= p.M) break;
        float4 v = *reinterpret_cast<const float4*>(csm + (size_t)rr * CS + lane * 4);
        if (p.partial) {
          *reinterpret_cast<float4*>(p.partial + ((size_t)blockIdx.z * p.M + m) * p.N + n) = v;
        } else {
          v.x += bias4.x; v.y += bias4.y; v.z += bias4.z; v.w += bias4.w;
          const int row = p.deint_H > 0 ? ((m & 3) * p.deint_H + (m >> 2)) : m;
          float4* o = reinterpret_cast<float4*>(p.c + (long long)row * p.ldc + n);
          if (p.accumulate) {
            const float4 old = *o;
            v.x += old.x; v.y += old.y; v.z += old.z; v.w += old.w;
          }
          *o = v;
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem_base) : "memory");
  }
}

__global__ void tc_splitk_reduce_kernel(TcParams p, int splits) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)p.M * p.N) return;
  const int m = (int)(idx / p.N), n = (int)(idx % p.N);
  float v = 0.f;
  for (int s = 0; s < splits; ++s) v += p.partial[(size_t)s * p.M * p.N + idx];
  if (p.bias) v += p.bias[n];
  const int row = p.deint_H > 0 ? ((m & 3) * p.deint_H + (m >> 2)) : m;
  float* o = p.c + (long long)row * p.ldc + n;
  *o = p.accumulate ? *o + v : v;
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
    else
      cudaGetLastError();
  }
  return fn;
}

// operand X(r, k): element (r,k) at ptr[r*s_r + k*s_k]; K-major if s_k == 1, MN-major if s_r == 1
// `plain_mn`: an MN-major operand that is read element-wise by converter threads (the TMEM-operand kernel) is
// loaded without swizzle (box = 32 rows x 32 k, 4 boxes per tile).
int make_tc_map(CUtensorMap* map, const float* ptr, long long s_r, long long s_k, int rows, int K, int* mn_major,
                int plain_mn) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { set_error("cuTensorMapEncodeTiled unavailable"); return MRG_E_UNSUPPORTED; }
  cuuint64_t dims[2], strides[1];
  cuuint32_t box[2], estr[2] = {1, 1};
  if (s_k == 1) {           // K-major: inner = K
    *mn_major = 0;
    dims[0] = (cuuint64_t)K; dims[1] = (cuuint64_t)rows;
    strides[0] = (cuuint64_t)s_r * 4;
    box[0] = TBK; box[1] = TBM;
  } else {                  // MN-major: inner = rows
    *mn_major = 1;
    dims[0] = (cuuint64_t)rows; dims[1] = (cuuint64_t)K;
    strides[0] = (cuuint64_t)s_k * 4;
    box[0] = 32; box[1] = TBK;
  }
  const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)ptr, dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE,
                         *mn_major ? (plain_mn ? CU_TENSOR_MAP_SWIZZLE_NONE : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)
                                   : CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return MRG_E_INVALID; }
  return 0;
}

static bool operand_ok(const float* ptr, long long s_r, long long s_k) {
  if (((uintptr_t)ptr & 15) != 0) return false;
  if (s_k == 1) return s_r >= 4 && s_r % 4 == 0;
  if (s_r == 1) return s_k >= 4 && s_k % 4 == 0;
  return false;
}

bool gemm_tc_supported(const GemmArgs& g) {
  if (g.M <= 0 || g.N <= 0 || g.K <= 0) return false;
  if (g.N % 4 != 0 || g.ldc % 4 != 0 || ((uintptr_t)g.c & 15) != 0) return false;
  if (g.bias && ((uintptr_t)g.bias & 15) != 0) return false;
  if ((long long)g.M * g.N < 64 * 64) return false;  // tiny problems: SIMT path
  return operand_ok(g.a, g.a_sm, g.a_sk) && operand_ok(g.b, g.b_sn, g.b_sk) && get_encode_fn() != nullptr;
}

int tc_splits(int M, int N, int K) {
  const int tiles = ((M + TBM - 1) / TBM) * ((N + TBN - 1) / TBN);
  const int kb = (K + TBK - 1) / TBK;
  if (tiles >= 120 || kb < 16) return 1;
  int s = 148 / tiles;  // one wave: never more CTAs than SMs
  if (s > kb / 8) s = kb / 8;
  return s < 1 ? 1 : s;
}

size_t gemm_tc_workspace_bytes(int M, int N, int K) {
  const int s = tc_splits(M, N, K);
  return s > 1 ? (size_t)s * M * N * sizeof(float) : 0;
}

int gemm_tc(const GemmArgs& g, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    MRG_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_set = true;
  }
  CUtensorMap ma, mb;
  TcParams p = {};
  if (int e = make_tc_map(&ma, g.a, g.a_sm, g.a_sk, g.M, g.K, &p.a_mn, 0)) return e;
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
      set_error("gemm_tc: workspace too small (%zu needed)", need);
      return MRG_E_WORKSPACE;
    }
    p.partial = (float*)workspace;
  }
  dim3 grid((g.N + TBN - 1) / TBN, (g.M + TBM - 1) / TBM, zdim);
  ProfScope prof(PROF_GEMM, stream);
  count_launch(zdim > 1 ? 2 : 1);
  gemm_tc_kernel<<<grid, TC_THREADS, SMEM_BYTES, stream>>>(ma, mb, p);
  MRG_CUDA_CHECK(cudaGetLastError());
  if (zdim > 1) {
    const long long total = (long long)g.M * g.N;
    tc_splitk_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(p, zdim);
    MRG_CUDA_CHECK(cudaGetLastError());
  }
  return 0;
}

}  // namespace mrg
