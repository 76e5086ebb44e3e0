// fp32 SIMT GEMM with arbitrary operand strides, split-K and a gate-de-interleaving epilogue.
// It is the always-available CUDA path for shapes the tcgen05 3xTF32 GEMM does not take
// (unaligned K / tiny problems) and the cross-check for it (MRG_F_SIMT_GEMM).
#include <cstdint>

#include "mrg_common.cuh"

namespace mrg {

constexpr int BM = 128, BN = 128, BK = 16, PAD = 4;

template <bool A_KCONTIG, bool B_NCONTIG>
__global__ void __launch_bounds__(256) gemm_simt_kernel(GemmArgs g, float* __restrict__ partial, int klen) {
  __shared__ __align__(16) float As[BK][BM + PAD];
  __shared__ __align__(16) float Bs[BK][BN + PAD];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kbeg = blockIdx.z * klen;
  const int kend = min(g.K, kbeg + klen);

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float ra[8], rb[8];
  auto load_tile = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int e = tid + i * 256;
      int m, k;
      if (A_KCONTIG) { k = e % BK; m = e / BK; } else { m = e % BM; k = e / BM; }
      const int gm = m0 + m, gk = k0 + k;
      const long long ai = (long long)gm * g.a_sm + (long long)gk * g.a_sk;
      ra[i] = (gm < g.M && gk < kend)
                  ? (g.a_bf16 ? bf16_lo_to_f32(reinterpret_cast<const unsigned short*>(g.a)[ai]) : g.a[ai])
                  : 0.f;
      int n, kb;
      if (B_NCONTIG) { n = e % BN; kb = e / BN; } else { kb = e % BK; n = e / BK; }
      const int gn = n0 + n, gkb = k0 + kb;
      rb[i] = (gn < g.N && gkb < kend) ? g.b[(long long)gkb * g.b_sk + (long long)gn * g.b_sn] : 0.f;
    }
  };
  auto store_tile = [&]() {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int e = tid + i * 256;
      int m, k;
      if (A_KCONTIG) { k = e % BK; m = e / BK; } else { m = e % BM; k = e / BM; }
      As[k][m] = ra[i];
      int n, kb;
      if (B_NCONTIG) { n = e % BN; kb = e / BN; } else { kb = e % BK; n = e / BK; }
      Bs[kb][n] = rb[i];
    }
  };

  if (kbeg < kend) load_tile(kbeg);
  for (int k0 = kbeg; k0 < kend; k0 += BK) {
    store_tile();
    __syncthreads();
    if (k0 + BK < kend) load_tile(k0 + BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[k][64 + tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

  const bool direct = (partial == nullptr);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (n >= g.N) continue;
      if (direct) {
        const int r = g.row_deinterleave_H > 0 ? ((m & 3) * g.row_deinterleave_H + (m >> 2)) : m;
        float v = acc[i][j];
        if (g.bias) v += g.bias[n];
        if (g.c_bf16) {
          reinterpret_cast<unsigned short*>(g.c)[(long long)r * g.ldc + n] = (unsigned short)(pack_bf16x2(v, 0.f) & 0xFFFFu);
        } else {
          float* o = g.c + (long long)r * g.ldc + n;
          *o = g.accumulate ? *o + v : v;
        }
      } else {
        partial[((size_t)blockIdx.z * g.M + m) * g.N + n] = acc[i][j];
      }
    }
  }
}

__global__ void splitk_reduce_kernel(GemmArgs g, const float* __restrict__ partial, int splits) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)g.M * g.N) return;
  const int m = (int)(idx / g.N), n = (int)(idx % g.N);
  float v = 0.f;
  for (int s = 0; s < splits; ++s) v += partial[(size_t)s * g.M * g.N + idx];
  if (g.bias) v += g.bias[n];
  const int r = g.row_deinterleave_H > 0 ? ((m & 3) * g.row_deinterleave_H + (m >> 2)) : m;
  if (g.c_bf16) {
    reinterpret_cast<unsigned short*>(g.c)[(long long)r * g.ldc + n] = (unsigned short)(pack_bf16x2(v, 0.f) & 0xFFFFu);
    return;
  }
  float* o = g.c + (long long)r * g.ldc + n;
  *o = g.accumulate ? *o + v : v;
}


// ---------------------------------------------------------------------------------------------------------------
// Skinny shapes of the 6-wide output head (Linear(256, 6): decoder output of every model, once per streamed frame).
// The 128 x 128 tile above spends 95 % of its work on padding there (16.7 us for M = 1024, 86 / 114 us for M = 76800).
// ---------------------------------------------------------------------------------------------------------------
// C[M][N <= 8] = A[M][K] . W[N][K]^T (+ bias): W staged in shared memory once per CTA, one warp per row, 16-byte loads,
// warp-shuffle reduction in a fixed order (deterministic).
constexpr int SKN_MAXN = 8, SKN_MAXK = 1024;
__global__ void __launch_bounds__(256) gemm_skinny_n_kernel(GemmArgs g) {
  __shared__ __align__(16) float ws[SKN_MAXN * SKN_MAXK];
  const int K = g.K, N = g.N, K4 = K >> 2;
  for (int e = threadIdx.x; e < N * K4; e += blockDim.x) {
    const int n = e / K4, k4 = e % K4;
    reinterpret_cast<float4*>(ws)[n * K4 + k4] = __ldg(reinterpret_cast<const float4*>(g.b + (long long)n * g.b_sn) + k4);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int m = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; m < g.M; m += warps) {
    const float4* ap = reinterpret_cast<const float4*>(g.a + (long long)m * g.a_sm);
    float acc[SKN_MAXN];
#pragma unroll
    for (int n = 0; n < SKN_MAXN; ++n) acc[n] = 0.f;
    for (int k4 = lane; k4 < K4; k4 += 32) {
      const float4 a = __ldg(ap + k4);
#pragma unroll
      for (int n = 0; n < SKN_MAXN; ++n)
        if (n < N) {
          const float4 w = reinterpret_cast<const float4*>(ws)[n * K4 + k4];
          acc[n] = fmaf(a.x, w.x, fmaf(a.y, w.y, fmaf(a.z, w.z, fmaf(a.w, w.w, acc[n]))));
        }
    }
    float out = 0.f;
#pragma unroll
    for (int n = 0; n < SKN_MAXN; ++n) {
      float v = acc[n];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == n) out = v;
    }
    if (lane < N) {
      if (g.bias) out += g.bias[lane];
      float* o = g.c + (long long)m * g.ldc + lane;
      *o = g.accumulate ? *o + out : out;
    }
  }
}

// partial[slab][M <= 8][N] = sum over the slab's k of A(m, k) B(k, n) with B rows contiguous (weight gradient of the
// head: A = dY^T, B = X): thread = output column, the <= 8 values of A per k are warp-uniform loads; the slabs are summed
// by splitk_reduce_kernel in a fixed order.
constexpr int SKM_MAXM = 8;
__global__ void __launch_bounds__(256) gemm_skinny_m_kernel(GemmArgs g, float* __restrict__ partial, int klen) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int kbeg = blockIdx.y * klen, kend = min(g.K, kbeg + klen);
  if (n >= g.N) return;
  float acc[SKM_MAXM];
#pragma unroll
  for (int m = 0; m < SKM_MAXM; ++m) acc[m] = 0.f;
  const float* bp = g.b + n;
#pragma unroll 4
  for (int k = kbeg; k < kend; ++k) {
    const float x = __ldg(bp + (long long)k * g.b_sk);
    const float* ak = g.a + (long long)k * g.a_sk;
#pragma unroll
    for (int m = 0; m < SKM_MAXM; ++m)
      if (m < g.M) acc[m] = fmaf(__ldg(ak + (long long)m * g.a_sm), x, acc[m]);
  }
#pragma unroll
  for (int m = 0; m < SKM_MAXM; ++m)
    if (m < g.M) partial[((size_t)blockIdx.y * g.M + m) * g.N + n] = acc[m];
}

static bool skinny_n_ok(const GemmArgs& g) {
  return g.N <= SKN_MAXN && g.M >= 64 && g.K >= 4 && g.K <= SKN_MAXK && g.K % 4 == 0 && g.a_sk == 1 && g.b_sk == 1 &&
         g.a_sm % 4 == 0 && g.b_sn % 4 == 0 && ((uintptr_t)g.a & 15) == 0 && ((uintptr_t)g.b & 15) == 0 && !g.a_bf16 &&
         !g.c_bf16 && g.row_deinterleave_H == 0;
}
static int skinny_m_slabs(const GemmArgs& g) {   // 0: the shape is not taken
  if (!(g.M <= SKM_MAXM && g.K >= 4096 && g.b_sn == 1 && !g.a_bf16 && !g.c_bf16 && g.row_deinterleave_H == 0)) return 0;
  const int col_blocks = (g.N + 255) / 256;
  int slabs = (148 * 4) / col_blocks;
  if (slabs > g.K / 64) slabs = g.K / 64;
  return slabs < 1 ? 1 : slabs;
}

static int pick_splits(int M, int N, int K) {
  const int tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
  if (tiles >= 96 || K < 1024) return 1;
  int s = 148 / tiles;
  const int smax = K / 256;
  if (s > smax) s = smax;
  if (s < 1) s = 1;
  return s;
}

size_t gemm_simt_workspace_bytes(int M, int N, int K) {
  const int s = pick_splits(M, N, K);
  size_t bytes = s > 1 ? (size_t)s * M * N * sizeof(float) : 0;
  if (M <= SKM_MAXM && K >= 4096) {   // the skinny-M kernel's slabs (at most 148 x 4 of them)
    const size_t sk = (size_t)148 * 4 * M * N * sizeof(float);
    if (sk > bytes) bytes = sk;
  }
  return bytes;
}

int gemm_simt(const GemmArgs& g, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  if (g.M <= 0 || g.N <= 0) return 0;
  MRG_REQUIRE(g.K >= 0, "gemm: negative K");
  MRG_REQUIRE(!(g.c_bf16 && g.accumulate), "gemm: a bfloat16 output cannot accumulate");
  if (skinny_n_ok(g)) {
    int grid = (g.M + 7) / 8;
    if (grid > 148 * 4) grid = 148 * 4;
    ProfScope prof(PROF_GEMM, stream);
    count_launch();
    gemm_skinny_n_kernel<<<grid, 256, 0, stream>>>(g);
    MRG_CUDA_CHECK(cudaGetLastError());
    return 0;
  }
  if (const int slabs = skinny_m_slabs(g)) {
    if (workspace != nullptr && workspace_bytes >= (size_t)slabs * g.M * g.N * sizeof(float)) {
      const int klen = (g.K + slabs - 1) / slabs;
      ProfScope prof(PROF_GEMM, stream);
      count_launch(2);
      gemm_skinny_m_kernel<<<dim3((g.N + 255) / 256, slabs), 256, 0, stream>>>(g, (float*)workspace, klen);
      MRG_CUDA_CHECK(cudaGetLastError());
      const long long total = (long long)g.M * g.N;
      splitk_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(g, (const float*)workspace, slabs);
      MRG_CUDA_CHECK(cudaGetLastError());
      return 0;
    }
  }
  const int splits = pick_splits(g.M, g.N, g.K);
  float* partial = nullptr;
  if (splits > 1) {
    if (workspace_bytes < (size_t)splits * g.M * g.N * sizeof(float) || workspace == nullptr) {
      set_error("gemm_simt: workspace too small (%zu needed)", (size_t)splits * g.M * g.N * sizeof(float));
      return MRG_E_WORKSPACE;
    }
    partial = (float*)workspace;
  }
  int klen = (g.K + splits - 1) / splits;
  klen = (klen + BK - 1) / BK * BK;
  if (klen == 0) klen = BK;
  dim3 grid((g.N + BN - 1) / BN, (g.M + BM - 1) / BM, splits);
  ProfScope prof(PROF_GEMM, stream);
  count_launch(splits > 1 ? 2 : 1);
  const bool ak = (g.a_sk == 1), bn = (g.b_sn == 1);
  if (ak && bn) gemm_simt_kernel<true, true><<<grid, 256, 0, stream>>>(g, partial, klen);
  else if (ak && !bn) gemm_simt_kernel<true, false><<<grid, 256, 0, stream>>>(g, partial, klen);
  else if (!ak && bn) gemm_simt_kernel<false, true><<<grid, 256, 0, stream>>>(g, partial, klen);
  else gemm_simt_kernel<false, false><<<grid, 256, 0, stream>>>(g, partial, klen);
  MRG_CUDA_CHECK(cudaGetLastError());
  if (splits > 1) {
    const long long total = (long long)g.M * g.N;
    splitk_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(g, partial, splits);
    MRG_CUDA_CHECK(cudaGetLastError());
  }
  return 0;
}

}  // namespace mrg
