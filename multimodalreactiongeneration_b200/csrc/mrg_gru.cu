// GRU recurrence: the token mixer lstmformer selects with ``emb_mixers: gru``
// (reference: nn.GRU constructed at mr_gen/model/utils/mixer_block.py:194, config mr_gen/model/lstmformer/config_gru.yaml).
//
//   r = sigmoid(gx_r + W_hr h + b_hr)        z = sigmoid(gx_z + W_hz h + b_hz)
//   n = tanh(gx_n + r * (W_hn h + b_hn))     h' = (1 - z) * n + z * h            (torch.nn.GRU, gate order r, z, n)
//
// gx = x W_ih^T + b_ih is time-parallel and comes from the tcgen05 projection GEMM (host side: gru.py); this file holds
// the sequential part.  First generation = the generic scheme of mrg_rec_generic.cu: one CTA per 4 batch rows, W_hh
// streamed from L2 every step, any hidden size.  (The cluster-resident scheme of the LSTM kernels — W_hh in registers,
// DSMEM exchange — carries over with 3 gate columns per unit instead of 4; not built yet.)
//
// Layouts: gx [T][B][3H]; y_ext [T+1][B][H] with slot 0 = h0 and slot t+1 = h_t (so H_prev of all steps is one
// contiguous matrix for dW_hh); reserve [T][B][4][H] = r, z, n, hn (hn = W_hn h + b_hn is needed by the backward).
// Backward writes dgx [T][B][3H] = d(pre-activation) seen from the input side (r, z, n) and dgh [T][B][3H] = the same
// seen from the hidden side (r, z, hn); the weight / input gradients are plain GEMMs over those two matrices.
#include "mrg_common.cuh"

namespace mrg {

constexpr int GRU_R = 4;  // batch rows per CTA

// smem: h_s[R][H], pre_s[R][3H]
__global__ void __launch_bounds__(256) gru_fwd_kernel(const float* __restrict__ gx, const float* __restrict__ w_hh,
                                                      const float* __restrict__ b_hh, float* __restrict__ y_ext,
                                                      float* __restrict__ reserve, int T, int B, int H, int train) {
  extern __shared__ float smem[];
  float* h_s = smem;
  float* pre_s = h_s + GRU_R * H;
  const int row0 = blockIdx.x * GRU_R;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int idx = tid; idx < GRU_R * H; idx += 256) {
    const int b = idx / H, j = idx % H;
    h_s[idx] = row0 + b < B ? y_ext[(size_t)(row0 + b) * H + j] : 0.f;
  }
  __syncthreads();
  for (int t = 0; t < T; ++t) {
    // pre[b][n] = sum_k W_hh[n][k] h[b][k]: one warp per gate row n, lanes over k (coalesced reads of W_hh)
    for (int n = warp; n < 3 * H; n += 8) {
      const float* wr = w_hh + (size_t)n * H;
      float acc[GRU_R] = {0.f, 0.f, 0.f, 0.f};
      for (int k = lane; k < H; k += 32) {
        const float wv = __ldg(wr + k);
#pragma unroll
        for (int b = 0; b < GRU_R; ++b) acc[b] = fmaf(wv, h_s[b * H + k], acc[b]);
      }
#pragma unroll
      for (int b = 0; b < GRU_R; ++b) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[b] += __shfl_xor_sync(0xffffffffu, acc[b], o);
      }
      if (lane == 0) {
        const float bias = b_hh ? b_hh[n] : 0.f;
#pragma unroll
        for (int b = 0; b < GRU_R; ++b) pre_s[b * 3 * H + n] = acc[b] + bias;
      }
    }
    __syncthreads();
    for (int idx = tid; idx < GRU_R * H; idx += 256) {
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

// smem: dh_s[R][H], dgh_s[R][3H]
__global__ void __launch_bounds__(256) gru_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ dh_n,
                                                      const float* __restrict__ reserve,
                                                      const float* __restrict__ y_ext, const float* __restrict__ w_hh,
                                                      float* __restrict__ dgx, float* __restrict__ dgh,
                                                      float* __restrict__ dh0, int T, int B, int H) {
  extern __shared__ float smem[];
  float* dh_s = smem;
  float* dgh_s = dh_s + GRU_R * H;
  const int row0 = blockIdx.x * GRU_R;
  const int tid = threadIdx.x;
  for (int idx = tid; idx < GRU_R * H; idx += 256) {
    const int b = idx / H, j = idx % H;
    dh_s[idx] = (dh_n && row0 + b < B) ? dh_n[(size_t)(row0 + b) * H + j] : 0.f;
  }
  for (int idx = tid; idx < GRU_R * 3 * H; idx += 256) dgh_s[idx] = 0.f;
  __syncthreads();
  for (int t = T - 1; t >= 0; --t) {
    for (int idx = tid; idx < GRU_R * H; idx += 256) {
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
    for (int k = tid; k < H; k += 256) {
      float acc[GRU_R] = {0.f, 0.f, 0.f, 0.f};
      for (int n = 0; n < 3 * H; ++n) {
        const float wv = __ldg(w_hh + (size_t)n * H + k);
#pragma unroll
        for (int b = 0; b < GRU_R; ++b) acc[b] = fmaf(wv, dgh_s[b * 3 * H + n], acc[b]);
      }
#pragma unroll
      for (int b = 0; b < GRU_R; ++b) dh_s[b * H + k] += acc[b];
    }
    __syncthreads();
  }
  if (dh0)
    for (int idx = tid; idx < GRU_R * H; idx += 256) {
      const int b = idx / H, j = idx % H;
      if (row0 + b < B) dh0[(size_t)(row0 + b) * H + j] = dh_s[idx];
    }
}

}  // namespace mrg

extern "C" int mrg_gru_forward(const float* gx, const float* w_hh, const float* b_hh, float* y_ext, float* reserve, int T,
                               int B, int H, int train, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MRG_REQUIRE(gx && w_hh && y_ext && T > 0 && B > 0 && H > 0, "mrg_gru_forward: bad arguments");
  MRG_REQUIRE(!train || reserve, "mrg_gru_forward: training needs the reserve buffer");
  const size_t smem = (size_t)(mrg::GRU_R * H + mrg::GRU_R * 3 * H) * sizeof(float);
  MRG_REQUIRE(smem <= 200 * 1024, "mrg_gru_forward: hidden size %d too large", H);
  if (smem > 48 * 1024)
    MRG_CUDA_CHECK(cudaFuncSetAttribute(mrg::gru_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mrg::ProfScope prof(mrg::PROF_REC_FWD, stream);
  mrg::count_launch();
  mrg::gru_fwd_kernel<<<(B + mrg::GRU_R - 1) / mrg::GRU_R, 256, smem, stream>>>(gx, w_hh, b_hh, y_ext, reserve, T, B, H,
                                                                               train);
  MRG_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int mrg_gru_backward(const float* dy, const float* dh_n, const float* reserve, const float* y_ext,
                                const float* w_hh, float* dgx, float* dgh, float* dh0, int T, int B, int H,
                                void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MRG_REQUIRE(reserve && y_ext && w_hh && dgx && dgh && T > 0 && B > 0 && H > 0, "mrg_gru_backward: bad arguments");
  const size_t smem = (size_t)(mrg::GRU_R * H + mrg::GRU_R * 3 * H) * sizeof(float);
  MRG_REQUIRE(smem <= 200 * 1024, "mrg_gru_backward: hidden size %d too large", H);
  if (smem > 48 * 1024)
    MRG_CUDA_CHECK(cudaFuncSetAttribute(mrg::gru_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mrg::ProfScope prof(mrg::PROF_REC_BWD, stream);
  mrg::count_launch();
  mrg::gru_bwd_kernel<<<(B + mrg::GRU_R - 1) / mrg::GRU_R, 256, smem, stream>>>(dy, dh_n, reserve, y_ext, w_hh, dgx, dgh, dh0,
                                                                               T, B, H);
  MRG_CUDA_CHECK(cudaGetLastError());
  return 0;
}
