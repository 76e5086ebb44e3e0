// Host side shared by the tcgen05 GEMM kernels: TMA tensor maps for K-major / MN-major fp32 operands, the split-K
// policy with its deterministic second-pass reduction, and the shape test that routes a GEMM to the tensor-core
// kernel (mrg_gemm_tc2.cu) or to the SIMT fp32 cross-check kernel (mrg_gemm_simt.cu).
#include <cstdlib>

#include "mrg_tc_common.cuh"

namespace mrg {

__global__ void tc_splitk_reduce_kernel(TcParams p, int splits) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)p.M * p.N) return;
  const int m = (int)(idx / p.N), n = (int)(idx % p.N);
  float v = 0.f;
  for (int s = 0; s < splits; ++s) v += p.partial[(size_t)s * p.M * p.N + idx];
  if (p.bias) v += p.bias[n];
  const int row = p.deint_H > 0 ? ((m & 3) * p.deint_H + (m >> 2)) : m;
  if (p.c_bf16) {
    reinterpret_cast<unsigned short*>(p.c)[(long long)row * p.ldc + n] = (unsigned short)(pack_bf16x2(v, 0.f) & 0xFFFFu);
    return;
  }
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
// `box_rows`: rows of a K-major tile (128, or 256 for the B operand of the 128 x 256 kernel).
// `bf16`: the operand is stored as bfloat16 (only for an A operand that goes through the converter warps: plain tiles,
// 64-byte rows, no swizzle).
int make_tc_map(CUtensorMap* map, const float* ptr, long long s_r, long long s_k, int rows, int K, int* mn_major,
                int plain_mn, int bf16, int box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { set_error("cuTensorMapEncodeTiled unavailable"); return MRG_E_UNSUPPORTED; }
  cuuint64_t dims[2], strides[1];
  cuuint32_t box[2], estr[2] = {1, 1};
  const cuuint64_t esz = bf16 ? 2 : 4;
  if (s_k == 1) {           // K-major: inner = K
    *mn_major = 0;
    dims[0] = (cuuint64_t)K; dims[1] = (cuuint64_t)rows;
    strides[0] = (cuuint64_t)s_r * esz;
    box[0] = TBK; box[1] = (cuuint32_t)box_rows;
  } else {                  // MN-major: inner = rows
    *mn_major = 1;
    dims[0] = (cuuint64_t)rows; dims[1] = (cuuint64_t)K;
    strides[0] = (cuuint64_t)s_k * esz;
    box[0] = 32; box[1] = TBK;
  }
  const CUtensorMapSwizzle swz =
      bf16 ? CU_TENSOR_MAP_SWIZZLE_NONE
           : (*mn_major ? (plain_mn ? CU_TENSOR_MAP_SWIZZLE_NONE : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)
                        : CU_TENSOR_MAP_SWIZZLE_128B);
  const CUresult r = enc(map, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)ptr,
                         dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return MRG_E_INVALID; }
  return 0;
}

static bool operand_ok(const float* ptr, long long s_r, long long s_k, int bf16 = 0) {
  if (((uintptr_t)ptr & 15) != 0) return false;
  const long long q = bf16 ? 8 : 4;   // global strides must be multiples of 16 bytes
  if (s_k == 1) return s_r >= q && s_r % q == 0;
  if (s_r == 1) return s_k >= q && s_k % q == 0;
  return false;
}

bool gemm_tc_supported(const GemmArgs& g) {
  if (g.M <= 0 || g.N <= 0 || g.K <= 0) return false;
  if (g.N % 4 != 0 || g.ldc % 4 != 0 || ((uintptr_t)g.c & (g.c_bf16 ? 7 : 15)) != 0) return false;
  if (g.c_bf16 && (g.accumulate || g.row_deinterleave_H > 0)) return false;
  if (g.bias && ((uintptr_t)g.bias & 15) != 0) return false;
  if ((long long)g.M * g.N < 64 * 64) return false;  // tiny problems: SIMT path
  return operand_ok(g.a, g.a_sm, g.a_sk, g.a_bf16) && operand_ok(g.b, g.b_sn, g.b_sk) && get_encode_fn() != nullptr;
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

}  // namespace mrg
