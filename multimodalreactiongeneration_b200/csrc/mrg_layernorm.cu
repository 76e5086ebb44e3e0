// Fused residual + LayerNorm (ResidualConnection of the reference: out = LN(module(x) + x),
// mr_gen/model/utils/residual_connection.py:29-32) — the HBM-bound epilogue next to every LSTM block.
// One warp per row; rows are addressed as (i0, i1) with independent strides per tensor so that the
// time-major LSTM output and the batch-first block input are consumed in place (no transposed copy).
// Backward recomputes s = y + x (both are alive for the LSTM backward anyway), so the only extra state is
// mean / rstd (8 bytes per row).  d(gamma), d(beta): per-CTA column partials + a second pass, no atomics.
#include <cstdlib>

#include "mrg_common.cuh"

namespace mrg {

struct LnArgs {
  const float* y; long long y_s0, y_s1;
  const float* x; long long x_s0, x_s1;      // x may be nullptr (plain LayerNorm)
  float* out; long long o_s0, o_s1;
  const float* dout; long long d_s0, d_s1;   // backward only
  float* dsum; long long g_s0, g_s1;         // backward only
  const float* gamma; const float* beta;
  float* mean; float* rstd;                  // [n0*n1]
  float* partial;                            // backward: [gridDim.x][2][H]
  int n0, n1, H;
  float eps;
};

template <int NV>  // H = 128 * NV
__global__ void __launch_bounds__(256) ln_fwd_kernel(LnArgs a) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const long long rows = (long long)a.n0 * a.n1;
  float4 gm[NV], bt[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    gm[v] = reinterpret_cast<const float4*>(a.gamma)[v * 32 + lane];
    bt[v] = reinterpret_cast<const float4*>(a.beta)[v * 32 + lane];
  }
  for (long long r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < rows; r += warps) {
    const long long i0 = r / a.n1, i1 = r % a.n1;
    const float4* yp = reinterpret_cast<const float4*>(a.y + i0 * a.y_s0 + i1 * a.y_s1);
    float4 s[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) s[v] = __ldcs(yp + v * 32 + lane);
    if (a.x) {
      const float4* xp = reinterpret_cast<const float4*>(a.x + i0 * a.x_s0 + i1 * a.x_s1);
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const float4 t = __ldg(xp + v * 32 + lane);
        s[v].x += t.x; s[v].y += t.y; s[v].z += t.z; s[v].w += t.w;
      }
    }
    float sum = 0.f;
#pragma unroll
    for (int v = 0; v < NV; ++v) sum += (s[v].x + s[v].y) + (s[v].z + s[v].w);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum / (float)a.H;
    float var = 0.f;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const float dx = s[v].x - mean, dy = s[v].y - mean, dz = s[v].z - mean, dw = s[v].w - mean;
      var += (dx * dx + dy * dy) + (dz * dz + dw * dw);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
    const float rstd = rsqrtf(var / (float)a.H + a.eps);
    float4* op = reinterpret_cast<float4*>(a.out + i0 * a.o_s0 + i1 * a.o_s1);
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      float4 o;
      o.x = (s[v].x - mean) * rstd * gm[v].x + bt[v].x;
      o.y = (s[v].y - mean) * rstd * gm[v].y + bt[v].y;
      o.z = (s[v].z - mean) * rstd * gm[v].z + bt[v].z;
      o.w = (s[v].w - mean) * rstd * gm[v].w + bt[v].w;
      op[v * 32 + lane] = o;
    }
    if (lane == 0 && a.mean) { a.mean[r] = mean; a.rstd[r] = rstd; }
  }
}

template <int NV>
__global__ void __launch_bounds__(256) ln_bwd_kernel(LnArgs a) {
  __shared__ float4 red[2][8][NV * 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const long long rows = (long long)a.n0 * a.n1;
  float4 gm[NV], dg[NV], db[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    gm[v] = reinterpret_cast<const float4*>(a.gamma)[v * 32 + lane];
    dg[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    db[v] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (long long r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < rows; r += warps) {
    const long long i0 = r / a.n1, i1 = r % a.n1;
    const float4* yp = reinterpret_cast<const float4*>(a.y + i0 * a.y_s0 + i1 * a.y_s1);
    const float4* dp = reinterpret_cast<const float4*>(a.dout + i0 * a.d_s0 + i1 * a.d_s1);
    const float mean = a.mean[r], rstd = a.rstd[r];
    float4 xh[NV], g[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) xh[v] = __ldg(yp + v * 32 + lane);
    if (a.x) {
      const float4* xp = reinterpret_cast<const float4*>(a.x + i0 * a.x_s0 + i1 * a.x_s1);
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const float4 t = __ldg(xp + v * 32 + lane);
        xh[v].x += t.x; xh[v].y += t.y; xh[v].z += t.z; xh[v].w += t.w;
      }
    }
    float c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const float4 d = __ldcs(dp + v * 32 + lane);
      xh[v].x = (xh[v].x - mean) * rstd; xh[v].y = (xh[v].y - mean) * rstd;
      xh[v].z = (xh[v].z - mean) * rstd; xh[v].w = (xh[v].w - mean) * rstd;
      dg[v].x += d.x * xh[v].x; dg[v].y += d.y * xh[v].y; dg[v].z += d.z * xh[v].z; dg[v].w += d.w * xh[v].w;
      db[v].x += d.x; db[v].y += d.y; db[v].z += d.z; db[v].w += d.w;
      g[v] = make_float4(d.x * gm[v].x, d.y * gm[v].y, d.z * gm[v].z, d.w * gm[v].w);
      c1 += (g[v].x + g[v].y) + (g[v].z + g[v].w);
      c2 += (g[v].x * xh[v].x + g[v].y * xh[v].y) + (g[v].z * xh[v].z + g[v].w * xh[v].w);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      c1 += __shfl_xor_sync(0xffffffffu, c1, o);
      c2 += __shfl_xor_sync(0xffffffffu, c2, o);
    }
    c1 /= (float)a.H; c2 /= (float)a.H;
    float4* gp = reinterpret_cast<float4*>(a.dsum + i0 * a.g_s0 + i1 * a.g_s1);
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      float4 o;
      o.x = rstd * (g[v].x - c1 - xh[v].x * c2);
      o.y = rstd * (g[v].y - c1 - xh[v].y * c2);
      o.z = rstd * (g[v].z - c1 - xh[v].z * c2);
      o.w = rstd * (g[v].w - c1 - xh[v].w * c2);
      gp[v * 32 + lane] = o;
    }
  }
  // column partials of this CTA: fixed-order sum over its 8 warps
#pragma unroll
  for (int v = 0; v < NV; ++v) { red[0][warp][v * 32 + lane] = dg[v]; red[1][warp][v * 32 + lane] = db[v]; }
  __syncthreads();
  for (int e = threadIdx.x; e < 2 * NV * 32; e += blockDim.x) {
    const int which = e / (NV * 32), col4 = e % (NV * 32);
    float4 s = red[which][0][col4];
    for (int w = 1; w < 8; ++w) {
      const float4 t = red[which][w][col4];
      s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
    }
    reinterpret_cast<float4*>(a.partial)[((size_t)blockIdx.x * 2 + which) * (NV * 32) + col4] = s;
  }
}

// 32 columns per CTA, 8 slices of the partial blocks per column, fixed-order tree: deterministic
__global__ void __launch_bounds__(256) ln_param_reduce_kernel(const float* __restrict__ partial,
                                                              float* __restrict__ dgamma,
                                                              float* __restrict__ dbeta, int nblocks, int H) {
  __shared__ float red[8][32];
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);  // column in [0, 2H)
  const int part = threadIdx.x >> 5;
  const int which = c / H, col = c % H;
  float s = 0.f;
  if (c < 2 * H)
    for (int b = part; b < nblocks; b += 8) s += partial[((size_t)b * 2 + which) * H + col];
  red[part][threadIdx.x & 31] = s;
  __syncthreads();
  if (part == 0 && c < 2 * H) {
    float t = red[0][threadIdx.x];
    for (int w = 1; w < 8; ++w) t += red[w][threadIdx.x];
    (which == 0 ? dgamma : dbeta)[col] = t;
  }
}

static int ln_grid(long long rows, int per_sm) {
  static int forced = -1;
  if (forced < 0) {
    const char* e = getenv("MRG_LN_CTAS");   // tuning experiments: resident CTAs per SM (<= 4: the partial-sum workspace)
    forced = e ? atoi(e) : 0;
  }
  if (forced >= 1 && forced <= 4) per_sm = forced;
  long long blocks = (rows + 7) / 8;
  if (blocks > 148 * per_sm) blocks = 148 * per_sm;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace mrg

using namespace mrg;

extern "C" size_t mrg_layernorm_workspace_bytes(int H) { return (size_t)148 * 4 * 2 * H * sizeof(float); }

extern "C" int mrg_residual_layernorm_forward(const float* y, long long y_s0, long long y_s1, const float* x,
                                              long long x_s0, long long x_s1, const float* gamma,
                                              const float* beta, float* out, long long o_s0, long long o_s1,
                                              float* mean, float* rstd, int n0, int n1, int H, float eps,
                                              void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MRG_REQUIRE(y && gamma && beta && out && n0 >= 0 && n1 >= 0, "mrg_residual_layernorm_forward: bad arguments");
  MRG_REQUIRE(H == 128 || H == 256 || H == 512, "mrg_residual_layernorm_forward: H must be 128, 256 or 512");
  MRG_REQUIRE((y_s0 | y_s1 | x_s0 | x_s1 | o_s0 | o_s1) % 4 == 0, "mrg_residual_layernorm_forward: unaligned strides");
  const long long rows = (long long)n0 * n1;
  if (rows == 0) return 0;
  LnArgs a = {};
  a.y = y; a.y_s0 = y_s0; a.y_s1 = y_s1; a.x = x; a.x_s0 = x_s0; a.x_s1 = x_s1;
  a.out = out; a.o_s0 = o_s0; a.o_s1 = o_s1; a.gamma = gamma; a.beta = beta; a.mean = mean; a.rstd = rstd;
  a.n0 = n0; a.n1 = n1; a.H = H; a.eps = eps;
  const int grid = ln_grid(rows, 4);   // 48 registers: 4 CTAs per SM in flight (B200, 76800 rows: 48.6 -> 37.7 us, tools/ln_bench.py)
  count_launch();
  if (H == 128) ln_fwd_kernel<1><<<grid, 256, 0, stream>>>(a);
  else if (H == 256) ln_fwd_kernel<2><<<grid, 256, 0, stream>>>(a);
  else ln_fwd_kernel<4><<<grid, 256, 0, stream>>>(a);
  MRG_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int mrg_residual_layernorm_backward(const float* dout, long long d_s0, long long d_s1, const float* y,
                                               long long y_s0, long long y_s1, const float* x, long long x_s0,
                                               long long x_s1, const float* gamma, const float* mean,
                                               const float* rstd, float* dsum, long long g_s0, long long g_s1,
                                               float* dgamma, float* dbeta, void* workspace,
                                               size_t workspace_bytes, int n0, int n1, int H, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MRG_REQUIRE(dout && y && gamma && mean && rstd && dsum && dgamma && dbeta && workspace,
              "mrg_residual_layernorm_backward: null pointer");
  MRG_REQUIRE(H == 128 || H == 256 || H == 512, "mrg_residual_layernorm_backward: H must be 128, 256 or 512");
  MRG_REQUIRE((d_s0 | d_s1 | y_s0 | y_s1 | x_s0 | x_s1 | g_s0 | g_s1) % 4 == 0,
              "mrg_residual_layernorm_backward: unaligned strides");
  if (workspace_bytes < mrg_layernorm_workspace_bytes(H)) {
    set_error("mrg_residual_layernorm_backward: workspace too small");
    return MRG_E_WORKSPACE;
  }
  const long long rows = (long long)n0 * n1;
  LnArgs a = {};
  a.dout = dout; a.d_s0 = d_s0; a.d_s1 = d_s1; a.y = y; a.y_s0 = y_s0; a.y_s1 = y_s1;
  a.x = x; a.x_s0 = x_s0; a.x_s1 = x_s1; a.dsum = dsum; a.g_s0 = g_s0; a.g_s1 = g_s1;
  a.gamma = gamma; a.mean = const_cast<float*>(mean); a.rstd = const_cast<float*>(rstd); a.partial = (float*)workspace;
  a.n0 = n0; a.n1 = n1; a.H = H;
  const int grid = ln_grid(rows, 3);   // 75 registers: 3 resident CTAs per SM (71.1 -> 59.9 us; 4 is slower: a second wave)
  count_launch(2);
  if (H == 128) ln_bwd_kernel<1><<<grid, 256, 0, stream>>>(a);
  else if (H == 256) ln_bwd_kernel<2><<<grid, 256, 0, stream>>>(a);
  else ln_bwd_kernel<4><<<grid, 256, 0, stream>>>(a);
  MRG_CUDA_CHECK(cudaGetLastError());
  ln_param_reduce_kernel<<<(2 * H + 31) / 32, 256, 0, stream>>>(a.partial, dgamma, dbeta, grid, H);
  MRG_CUDA_CHECK(cudaGetLastError());
  return 0;
}
