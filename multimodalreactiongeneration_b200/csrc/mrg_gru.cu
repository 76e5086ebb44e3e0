// GRU recurrence: the token mixer lstmformer selects with ``emb_mixers: gru``
// (reference: nn.GRU constructed at mr_gen/model/utils/mixer_block.py:194, config mr_gen/model/lstmformer/config_gru.yaml).
//
//   r = sigmoid(gx_r + W_hr h + b_hr)        z = sigmoid(gx_z + W_hz h + b_hz)
//   n = tanh(gx_n + r * (W_hn h + b_hn))     h' = (1 - z) * n + z * h            (torch.nn.GRU, gate order r, z, n)
//
// gx = x W_ih^T + b_ih is time-parallel and comes from the tcgen05 projection GEMM (host side: gru.py); this file holds
// the sequential part.  First generation = the generic scheme of mrg_rec_generic.cu: one CTA per 1-4 batch rows (as few as it takes to cover the SMs), W_hh
// streamed from L2 every step, any hidden size.  (The cluster-resident scheme of the LSTM kernels — W_hh in registers,
// DSMEM exchange — carries over with 3 gate columns per unit instead of 4; not built yet.)
//
// Layouts: gx [T][B][3H]; y_ext [T+1][B][H] with slot 0 = h0 and slot t+1 = h_t (so H_prev of all steps is one
// contiguous matrix for dW_hh); reserve [T][B][4][H] = r, z, n, hn (hn = W_hn h + b_hn is needed by the backward).
// Backward writes dgx [T][B][3H] = d(pre-activation) seen from the input side (r, z, n) and dgh [T][B][3H] = the same
// seen from the hidden side (r, z, hn); the weight / input gradients are plain GEMMs over those two matrices.
#include "mrg_common.cuh"

namespace mrg {

constexpr int GRU_R = 4;          // row capacity of a CTA (shared-memory layout)
constexpr int GRU_THREADS = 512;  // 16 warps: the W_hh stream from L2 is latency-bound, it needs loads in flight
constexpr int GRU_NR = 4;         // gate rows per warp iteration (forward)
constexpr int GRU_NP = 8;         // parts the gate rows are cut into (backward, vector path)

// rows per CTA: as few as it takes to put a CTA on every SM (B=64 -> 1 row per CTA, 64 CTAs), at most GRU_R
static int gru_rows_per_cta(int B) {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int r = (B + sms - 1) / sms;
  return r < 1 ? 1 : (r > GRU_R ? GRU_R : r);
}

// smem: h_s[R][H], pre_s[R][3H]
__global__ void __launch_bounds__(GRU_THREADS) gru_fwd_kernel(const float* __restrict__ gx, const float* __restrict__ w_hh,
                                                              const float* __restrict__ b_hh, float* __restrict__ y_ext,
                                                              float* __restrict__ reserve, int T, int B, int H, int train,
                                                              int rpc) {
  extern __shared__ __align__(16) float smem[];
  float* h_s = smem;
  float* pre_s = h_s + GRU_R * H;
  const int row0 = blockIdx.x * rpc;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = GRU_THREADS / 32;
  for (int idx = tid; idx < GRU_R * H; idx += GRU_THREADS) {
    const int b = idx / H, j = idx % H;
    h_s[idx] = (b < rpc && row0 + b < B) ? y_ext[(size_t)(row0 + b) * H + j] : 0.f;
  }
  __syncthreads();
  const bool vec = (H % 4) == 0;
  for (int t = 0; t < T; ++t) {
    // pre[b][n] = sum_k W_hh[n][k] h[b][k]: a warp takes GRU_NR gate rows at a time, lanes over k (16-byte loads when
    // H % 4 == 0), so that GRU_NR x H/128 independent 16-byte loads per lane are in flight: the stream of W_hh from L2
    // is latency-bound and lives on memory-level parallelism
    for (int n = warp * GRU_NR; n < 3 * H; n += NW * GRU_NR) {
      float acc[GRU_NR][GRU_R];
      const float* wr[GRU_NR];
#pragma unroll
      for (int r = 0; r < GRU_NR; ++r) {
        wr[r] = w_hh + (size_t)min(n + r, 3 * H - 1) * H;
#pragma unroll
        for (int b = 0; b < GRU_R; ++b) acc[r][b] = 0.f;
      }
      if (vec) {
        for (int k4 = lane; k4 < H / 4; k4 += 32) {
          float4 a[GRU_NR];
#pragma unroll
          for (int r = 0; r < GRU_NR; ++r) a[r] = __ldg(reinterpret_cast<const float4*>(wr[r]) + k4);
#pragma unroll
          for (int b = 0; b < GRU_R; ++b) {
            const float4 h4 = *reinterpret_cast<const float4*>(h_s + b * H + k4 * 4);
#pragma unroll
            for (int r = 0; r < GRU_NR; ++r)
              acc[r][b] = fmaf(a[r].x, h4.x, fmaf(a[r].y, h4.y, fmaf(a[r].z, h4.z, fmaf(a[r].w, h4.w, acc[r][b]))));
          }
        }
      } else {
        for (int k = lane; k < H; k += 32) {
#pragma unroll
          for (int r = 0; r < GRU_NR; ++r) {
            const float av = __ldg(wr[r] + k);
#pragma unroll
            for (int b = 0; b < GRU_R; ++b) acc[r][b] = fmaf(av, h_s[b * H + k], acc[r][b]);
          }
        }
      }
#pragma unroll
      for (int r = 0; r < GRU_NR; ++r)
#pragma unroll
        for (int b = 0; b < GRU_R; ++b) {
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) acc[r][b] += __shfl_xor_sync(0xffffffffu, acc[r][b], o);
        }
#pragma unroll
      for (int r = 0; r < GRU_NR; ++r)
        if (lane == r && n + r < 3 * H) {
          const float bias = b_hh ? b_hh[n + r] : 0.f;
#pragma unroll
          for (int b = 0; b < GRU_R; ++b) pre_s[b * 3 * H + n + r] = acc[r][b] + bias;
        }
    }
    __syncthreads();
    for (int idx = tid; idx < rpc * H; idx += GRU_THREADS) {
      const int b = idx / H, j = idx % H;
      if (row0 + b >= B) continue;
      const size_t fb = (size_t)t * B + row0 + b;
      const float* g = gx + fb * 3 * H;
      const float* p = pre_s + b * 3 * H;
      const float r = sigmoid_acc(g[j] + p[j]);
      const float z = sigmoid_acc(g[H + j] + p[H + j]);
      const float hn = p[2 * H + j];
      const float n = tanhf(g[2 * H + j] + r * hn);
      const float h = (1.f - z) * n + z * h_s[idx];
      h_s[idx] = h;
      y_ext[((size_t)(t + 1) * B + row0 + b) * H + j] = h;
      if (train) {
        float* rs = reserve + fb * 4 * H;
        rs[j] = r; rs[H + j] = z; rs[2 * H + j] = n; rs[3 * H + j] = hn;
      }
    }
    __syncthreads();
  }
}

// Forward with W_hh TRANSPOSED (w_t [H][3H], made once per call by the host): the reduction index k becomes the row of the
// matrix, so every thread streams 16-byte pieces of consecutive gate columns n with GRU_NP-way split over k — the same
// access pattern as the backward kernel, no cross-lane reduction, 8 loads in flight per thread.
// smem: h_s[R][H], part_s[GRU_NP][R][3H]
__global__ void __launch_bounds__(GRU_THREADS) gru_fwd_t_kernel(const float* __restrict__ gx, const float* __restrict__ w_t,
                                                                const float* __restrict__ b_hh, float* __restrict__ y_ext,
                                                                float* __restrict__ reserve, int T, int B, int H, int train,
                                                                int rpc) {
  extern __shared__ __align__(16) float smem[];
  float* h_s = smem;
  float* part_s = h_s + GRU_R * H;
  const int row0 = blockIdx.x * rpc;
  const int tid = threadIdx.x;
  const int G = 3 * H, G4 = G / 4;
  for (int idx = tid; idx < GRU_R * H; idx += GRU_THREADS) {
    const int b = idx / H, j = idx % H;
    h_s[idx] = (b < rpc && row0 + b < B) ? y_ext[(size_t)(row0 + b) * H + j] : 0.f;
  }
  __syncthreads();
  const int per = (H + GRU_NP - 1) / GRU_NP;
  for (int t = 0; t < T; ++t) {
    for (int w = tid; w < GRU_NP * G4; w += GRU_THREADS) {
      const int part = w / G4, n4 = w % G4;
      const int k_lo = part * per, k_hi = min(H, k_lo + per);
      float4 acc[GRU_R];
#pragma unroll
      for (int b = 0; b < GRU_R; ++b) acc[b] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
      for (int k = k_lo; k < k_hi; ++k) {
        const float4 wv = __ldg(reinterpret_cast<const float4*>(w_t + (size_t)k * G) + n4);
#pragma unroll
        for (int b = 0; b < GRU_R; ++b) {
          const float hv = h_s[b * H + k];
          acc[b].x = fmaf(wv.x, hv, acc[b].x); acc[b].y = fmaf(wv.y, hv, acc[b].y);
          acc[b].z = fmaf(wv.z, hv, acc[b].z); acc[b].w = fmaf(wv.w, hv, acc[b].w);
        }
      }
#pragma unroll
      for (int b = 0; b < GRU_R; ++b) reinterpret_cast<float4*>(part_s + (size_t)(part * GRU_R + b) * G)[n4] = acc[b];
    }
    __syncthreads();
    for (int idx = tid; idx < rpc * H; idx += GRU_THREADS) {
      const int b = idx / H, j = idx % H;
      if (row0 + b >= B) continue;
      float pr = b_hh ? b_hh[j] : 0.f, pz = b_hh ? b_hh[H + j] : 0.f, hn = b_hh ? b_hh[2 * H + j] : 0.f;
#pragma unroll
      for (int q = 0; q < GRU_NP; ++q) {
        const float* p = part_s + (size_t)(q * GRU_R + b) * G;
        pr += p[j]; pz += p[H + j]; hn += p[2 * H + j];
      }
      const size_t fb = (size_t)t * B + row0 + b;
      const float* g = gx + fb * G;
      const float r = sigmoid_acc(g[j] + pr);
      const float z = sigmoid_acc(g[H + j] + pz);
      const float n = tanhf(g[2 * H + j] + r * hn);
      const float h = (1.f - z) * n + z * h_s[idx];
      h_s[idx] = h;   // only this thread touches h_s[idx] in this phase; the next matvec starts after the barrier
      y_ext[((size_t)(t + 1) * B + row0 + b) * H + j] = h;
      if (train) {
        float* rs = reserve + fb * 4 * H;
        rs[j] = r; rs[H + j] = z; rs[2 * H + j] = n; rs[3 * H + j] = hn;
      }
    }
    __syncthreads();
  }
}

// smem: dh_s[R][H], dgh_s[R][3H], part_s[GRU_NP][R][H] (partial sums of the gate-row parts)
__global__ void __launch_bounds__(GRU_THREADS) gru_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ dh_n,
                                                              const float* __restrict__ reserve,
                                                              const float* __restrict__ y_ext,
                                                              const float* __restrict__ w_hh, float* __restrict__ dgx,
                                                              float* __restrict__ dgh, float* __restrict__ dh0, int T, int B,
                                                              int H, int rpc) {
  extern __shared__ __align__(16) float smem[];
  float* dh_s = smem;
  float* dgh_s = dh_s + GRU_R * H;
  float* part_s = dgh_s + GRU_R * 3 * H;
  const int row0 = blockIdx.x * rpc;
  const int tid = threadIdx.x;
  for (int idx = tid; idx < GRU_R * H; idx += GRU_THREADS) {
    const int b = idx / H, j = idx % H;
    dh_s[idx] = (dh_n && b < rpc && row0 + b < B) ? dh_n[(size_t)(row0 + b) * H + j] : 0.f;
  }
  for (int idx = tid; idx < GRU_R * 3 * H; idx += GRU_THREADS) dgh_s[idx] = 0.f;
  __syncthreads();
  for (int t = T - 1; t >= 0; --t) {
    for (int idx = tid; idx < rpc * H; idx += GRU_THREADS) {
      const int b = idx / H, j = idx % H;
      if (row0 + b >= B) continue;
      const size_t fb = (size_t)t * B + row0 + b;
      const float* rs = reserve + fb * 4 * H;
      const float r = rs[j], z = rs[H + j], n = rs[2 * H + j], hn = rs[3 * H + j];
      const float hp = y_ext[fb * H + j];  // slot t = h_{t-1}
      float dh = dh_s[idx];
      if (dy) dh += dy[fb * H + j];
      const float dpn = dh * (1.f - z) * (1.f - n * n);   // d pre-activation of n
      const float dpz = dh * (hp - n) * z * (1.f - z);
      const float dhn = dpn * r;
      const float dpr = dpn * hn * r * (1.f - r);
      float* gxo = dgx + fb * 3 * H;
      float* gho = dgh + fb * 3 * H;
      gxo[j] = dpr; gxo[H + j] = dpz; gxo[2 * H + j] = dpn;
      gho[j] = dpr; gho[H + j] = dpz; gho[2 * H + j] = dhn;
      float* ds = dgh_s + b * 3 * H;
      ds[j] = dpr; ds[H + j] = dpz; ds[2 * H + j] = dhn;
      dh_s[idx] = dh * z;  // the direct path h_{t-1} -> h_t; the path through W_hh is added below
    }
    __syncthreads();
    // dh[b][k] += sum_n dgh[b][n] W_hh[n][k]
    if ((H % 4) == 0 && H / 4 * GRU_NP <= GRU_THREADS) {
      // thread = (one of GRU_NP parts of the gate rows, 4 consecutive columns k): 16-byte loads coalesced over k, 8 of
      // them in flight per thread; the GRU_NP partial sums meet in shared memory
      const int part = tid / (H / 4), k4 = tid % (H / 4);
      if (part < GRU_NP) {
        const int per = (3 * H + GRU_NP - 1) / GRU_NP;
        const int n_lo = part * per, n_hi = min(3 * H, n_lo + per);
        float4 acc[GRU_R];
#pragma unroll
        for (int b = 0; b < GRU_R; ++b) acc[b] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
        for (int n = n_lo; n < n_hi; ++n) {
          const float4 wv = __ldg(reinterpret_cast<const float4*>(w_hh + (size_t)n * H) + k4);
#pragma unroll
          for (int b = 0; b < GRU_R; ++b) {
            const float g = dgh_s[b * 3 * H + n];
            acc[b].x = fmaf(wv.x, g, acc[b].x); acc[b].y = fmaf(wv.y, g, acc[b].y);
            acc[b].z = fmaf(wv.z, g, acc[b].z); acc[b].w = fmaf(wv.w, g, acc[b].w);
          }
        }
#pragma unroll
        for (int b = 0; b < GRU_R; ++b) reinterpret_cast<float4*>(part_s + (part * GRU_R + b) * H)[k4] = acc[b];
      }
      __syncthreads();
      for (int idx = tid; idx < GRU_R * H; idx += GRU_THREADS) {
        float v = dh_s[idx];
#pragma unroll
        for (int q = 0; q < GRU_NP; ++q) v += part_s[q * GRU_R * H + idx];
        dh_s[idx] = v;
      }
      __syncthreads();
    } else {
      // scalar path (any H): thread = (half of the gate rows, column k)
      const int half = tid / (GRU_THREADS / 2), kk0 = tid % (GRU_THREADS / 2);
      const int n_lo = half * ((3 * H + 1) / 2), n_hi = half == 0 ? (3 * H + 1) / 2 : 3 * H;
      for (int k = kk0; k < H; k += GRU_THREADS / 2) {
        float acc[GRU_R] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 8
        for (int n = n_lo; n < n_hi; ++n) {
          const float wv = __ldg(w_hh + (size_t)n * H + k);
#pragma unroll
          for (int b = 0; b < GRU_R; ++b) acc[b] = fmaf(wv, dgh_s[b * 3 * H + n], acc[b]);
        }
        if (half == 0) {
#pragma unroll
          for (int b = 0; b < GRU_R; ++b) dh_s[b * H + k] += acc[b];
        } else {
#pragma unroll
          for (int b = 0; b < GRU_R; ++b) part_s[b * H + k] = acc[b];
        }
      }
      __syncthreads();
      for (int idx = tid; idx < GRU_R * H; idx += GRU_THREADS) dh_s[idx] += part_s[idx];
      __syncthreads();
    }
  }
  if (dh0)
    for (int idx = tid; idx < rpc * H; idx += GRU_THREADS) {
      const int b = idx / H, j = idx % H;
      if (row0 + b < B) dh0[(size_t)(row0 + b) * H + j] = dh_s[idx];
    }
}

}  // namespace mrg

extern "C" int mrg_gru_forward(const float* gx, const float* w_hh, const float* w_hh_t, const float* b_hh, float* y_ext,
                               float* reserve, int T, int B, int H, int train, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MRG_REQUIRE(gx && w_hh && y_ext && T > 0 && B > 0 && H > 0, "mrg_gru_forward: bad arguments");
  MRG_REQUIRE(!train || reserve, "mrg_gru_forward: training needs the reserve buffer");
  const size_t smem = (size_t)(mrg::GRU_R * H + mrg::GRU_R * 3 * H) * sizeof(float);
  const int rpc = mrg::gru_rows_per_cta(B);
  MRG_REQUIRE(smem <= 200 * 1024, "mrg_gru_forward: hidden size %d too large", H);
  if (smem > 48 * 1024)
    MRG_CUDA_CHECK(cudaFuncSetAttribute(mrg::gru_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const size_t smem_t = (size_t)(mrg::GRU_R * H + mrg::GRU_NP * mrg::GRU_R * 3 * H) * sizeof(float);
  if (w_hh_t && H % 4 == 0 && smem_t <= 200 * 1024) {   // transposed weights supplied: the column-streaming kernel
    MRG_REQUIRE(((uintptr_t)w_hh_t & 15) == 0, "mrg_gru_forward: w_hh_t must be 16-byte aligned");
    if (smem_t > 48 * 1024)
      MRG_CUDA_CHECK(cudaFuncSetAttribute(mrg::gru_fwd_t_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_t));
    mrg::ProfScope prof(mrg::PROF_REC_FWD, stream);
    mrg::count_launch();
    mrg::gru_fwd_t_kernel<<<(B + rpc - 1) / rpc, mrg::GRU_THREADS, smem_t, stream>>>(gx, w_hh_t, b_hh, y_ext, reserve, T, B,
                                                                                   H, train, rpc);
    MRG_CUDA_CHECK(cudaGetLastError());
    return 0;
  }
  mrg::ProfScope prof(mrg::PROF_REC_FWD, stream);
  mrg::count_launch();
  mrg::gru_fwd_kernel<<<(B + rpc - 1) / rpc, mrg::GRU_THREADS, smem, stream>>>(gx, w_hh, b_hh, y_ext, reserve, T, B, H,
                                                                              train, rpc);
  MRG_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int mrg_gru_backward(const float* dy, const float* dh_n, const float* reserve, const float* y_ext,
                                const float* w_hh, float* dgx, float* dgh, float* dh0, int T, int B, int H,
                                void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MRG_REQUIRE(reserve && y_ext && w_hh && dgx && dgh && T > 0 && B > 0 && H > 0, "mrg_gru_backward: bad arguments");
  const size_t smem = (size_t)((1 + mrg::GRU_NP) * mrg::GRU_R * H + mrg::GRU_R * 3 * H) * sizeof(float);
  const int rpc = mrg::gru_rows_per_cta(B);
  MRG_REQUIRE(smem <= 200 * 1024, "mrg_gru_backward: hidden size %d too large", H);
  if (smem > 48 * 1024)
    MRG_CUDA_CHECK(cudaFuncSetAttribute(mrg::gru_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mrg::ProfScope prof(mrg::PROF_REC_BWD, stream);
  mrg::count_launch();
  mrg::gru_bwd_kernel<<<(B + rpc - 1) / rpc, mrg::GRU_THREADS, smem, stream>>>(dy, dh_n, reserve, y_ext, w_hh, dgx, dgh,
                                                                              dh0, T, B, H, rpc);
  MRG_CUDA_CHECK(cudaGetLastError());
  return 0;
}
