// Generic recurrent kernels: any hidden size, one CTA per 4 batch rows, W_hh streamed from L2.
// They serve hidden sizes the cluster kernels are not instantiated for (anything but 128 / 256)
// and cross-check them (MRG_F_GENERIC_REC).  Same reserve layout as the cluster kernels.
#include "mrg_common.cuh"

namespace mrg {

constexpr int GR = 4;  // batch rows per CTA

__device__ __forceinline__ void slots(int dir, int T, int step, int& t, int& prev_slot, int& out_slot) {
  if (dir == 0) { t = step; prev_slot = t; out_slot = t + 1; }
  else { t = T - 1 - step; prev_slot = t + 1; out_slot = t; }
}

// smem: h_s[GR][H], c_s[GR][H], pre_s[GR][4H]
__global__ void __launch_bounds__(256) rec_fwd_generic_kernel(RecArgs a) {
  extern __shared__ float smem[];
  const int H = a.H, B = a.B, T = a.T;
  float* h_s = smem;
  float* c_s = h_s + GR * H;
  float* pre_s = c_s + GR * H;
  const int slices = (B + GR - 1) / GR;
  const int d = blockIdx.x / slices;
  const int row0 = (blockIdx.x % slices) * GR;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* __restrict__ W = d == 0 ? a.w_hh[0] : a.w_hh[1];
  float* gates = a.gates + (size_t)d * T * B * 4 * H;
  float* y_ext = a.y_ext + (size_t)d * (T + 1) * B * H;
  float* c_ext = a.c_ext + (size_t)d * (T + 1) * B * H;

  {
    const int init_slot = d == 0 ? 0 : T;
    for (int idx = tid; idx < GR * H; idx += 256) {
      const int b = idx / H, j = idx % H;
      const bool ok = row0 + b < B;
      h_s[idx] = ok ? y_ext[((size_t)init_slot * B + row0 + b) * H + j] : 0.f;
      c_s[idx] = ok ? c_ext[((size_t)init_slot * B + row0 + b) * H + j] : 0.f;
    }
  }
  __syncthreads();

  for (int step = 0; step < T; ++step) {
    int t, prev_slot, out_slot;
    slots(d, T, step, t, prev_slot, out_slot);
    // 1. pre[b][n] = sum_k W[row(n)][k] * h[b][k], one warp per interleaved column n = j*4+g
    for (int n = warp; n < 4 * H; n += 8) {
      const float* wr = W + (size_t)((n & 3) * H + (n >> 2)) * H;
      float acc[GR] = {0.f, 0.f, 0.f, 0.f};
      for (int k = lane; k < H; k += 32) {
        const float wv = wr[k];
#pragma unroll
        for (int b = 0; b < GR; ++b) acc[b] = fmaf(wv, h_s[b * H + k], acc[b]);
      }
#pragma unroll
      for (int b = 0; b < GR; ++b) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[b] += __shfl_xor_sync(0xffffffffu, acc[b], o);
      }
      if (lane == 0) {
#pragma unroll
        for (int b = 0; b < GR; ++b) pre_s[b * 4 * H + n] = acc[b];
      }
    }
    __syncthreads();
    // 2. gates + cell update
    for (int idx = tid; idx < GR * H; idx += 256) {
      const int b = idx / H, j = idx % H;
      if (row0 + b >= B) continue;
      float4* gp = reinterpret_cast<float4*>(gates + (((size_t)t * B + row0 + b) * H + j) * 4);
      const float4 x = *gp;
      const float* p = pre_s + b * 4 * H + j * 4;
      const float gi = sigmoid_acc(x.x + p[0]);
      const float gf = sigmoid_acc(x.y + p[1]);
      const float gg = tanhf(x.z + p[2]);
      const float go = sigmoid_acc(x.w + p[3]);
      const float c = gf * c_s[idx] + gi * gg;
      const float h = go * tanhf(c);
      c_s[idx] = c;
      h_s[idx] = h;
      if (a.train) *gp = make_float4(gi, gf, gg, go);
      y_ext[((size_t)out_slot * B + row0 + b) * H + j] = h;
      c_ext[((size_t)out_slot * B + row0 + b) * H + j] = c;
    }
    __syncthreads();
  }
}

// smem: dh_s[GR][H], dc_s[GR][H], dpre_s[GR][4H], db_s[GR][4H]
__global__ void __launch_bounds__(256) rec_bwd_generic_kernel(RecBwdArgs a) {
  extern __shared__ float smem[];
  const int H = a.H, B = a.B, T = a.T, D = a.D;
  float* dh_s = smem;
  float* dc_s = dh_s + GR * H;
  float* dpre_s = dc_s + GR * H;
  float* db_s = dpre_s + GR * 4 * H;
  const int slices = (B + GR - 1) / GR;
  const int d = blockIdx.x / slices;
  const int row0 = (blockIdx.x % slices) * GR;
  const int tid = threadIdx.x;
  const float* __restrict__ W = d == 0 ? a.w_hh[0] : a.w_hh[1];
  float* gates = a.gates + (size_t)d * T * B * 4 * H;
  const float* c_ext = a.c_ext + (size_t)d * (T + 1) * B * H;

  for (int idx = tid; idx < GR * H; idx += 256) {
    const int b = idx / H, j = idx % H;
    const bool ok = row0 + b < B;
    dh_s[idx] = (ok && a.dh_n) ? a.dh_n[((size_t)d * B + row0 + b) * H + j] : 0.f;
    dc_s[idx] = (ok && a.dc_n) ? a.dc_n[((size_t)d * B + row0 + b) * H + j] : 0.f;
  }
  for (int idx = tid; idx < GR * 4 * H; idx += 256) { db_s[idx] = 0.f; dpre_s[idx] = 0.f; }
  __syncthreads();

  for (int step = T - 1; step >= 0; --step) {
    int t, prev_slot, out_slot;
    slots(d, T, step, t, prev_slot, out_slot);
    for (int idx = tid; idx < GR * H; idx += 256) {
      const int b = idx / H, j = idx % H;
      if (row0 + b >= B) continue;
      const size_t rb = row0 + b;
      float4* gp = reinterpret_cast<float4*>(gates + (((size_t)t * B + rb) * H + j) * 4);
      const float4 g4 = *gp;
      const float c = c_ext[((size_t)out_slot * B + rb) * H + j];
      const float cp = c_ext[((size_t)prev_slot * B + rb) * H + j];
      float dh = dh_s[idx];
      if (a.dy) dh += a.dy[((size_t)t * B + rb) * D * H + (size_t)d * H + j];
      const float tc = tanhf(c);
      const float d_o = dh * tc;
      const float dct = dc_s[idx] + dh * g4.w * (1.f - tc * tc);
      const float d_i = dct * g4.z, d_g = dct * g4.x, d_f = dct * cp;
      dc_s[idx] = dct * g4.y;
      const float4 dp = make_float4(d_i * g4.x * (1.f - g4.x), d_f * g4.y * (1.f - g4.y),
                                    d_g * (1.f - g4.z * g4.z), d_o * g4.w * (1.f - g4.w));
      *gp = dp;
      float* ds = dpre_s + b * 4 * H + j * 4;
      float* bs = db_s + b * 4 * H + j * 4;
      ds[0] = dp.x; ds[1] = dp.y; ds[2] = dp.z; ds[3] = dp.w;
      bs[0] += dp.x; bs[1] += dp.y; bs[2] += dp.z; bs[3] += dp.w;
    }
    __syncthreads();
    // dh_rec[b][k] = sum_n dpre[b][n] * W[row(n)][k]
    for (int k = tid; k < H; k += 256) {
      float acc[GR] = {0.f, 0.f, 0.f, 0.f};
      for (int n = 0; n < 4 * H; ++n) {
        const float wv = W[(size_t)((n & 3) * H + (n >> 2)) * H + k];
#pragma unroll
        for (int b = 0; b < GR; ++b) acc[b] = fmaf(wv, dpre_s[b * 4 * H + n], acc[b]);
      }
#pragma unroll
      for (int b = 0; b < GR; ++b) dh_s[b * H + k] = acc[b];
    }
    __syncthreads();
  }
  for (int idx = tid; idx < GR * H; idx += 256) {
    const int b = idx / H, j = idx % H;
    if (row0 + b >= B) continue;
    if (a.dh0[d]) a.dh0[d][(size_t)(row0 + b) * H + j] = dh_s[idx];
    if (a.dc0[d]) a.dc0[d][(size_t)(row0 + b) * H + j] = dc_s[idx];
  }
  for (int idx = tid; idx < GR * 4 * H; idx += 256) {
    const int b = idx / (4 * H), n = idx % (4 * H);
    if (row0 + b >= B) continue;
    a.db_part[((size_t)d * B + row0 + b) * 4 * H + n] = db_s[idx];
  }
}

int rec_forward_generic(const RecArgs& a, cudaStream_t stream) {
  const size_t smem = (size_t)(2 * GR * a.H + GR * 4 * a.H) * sizeof(float);
  MRG_REQUIRE(smem <= 200 * 1024, "generic recurrent kernel: hidden size %d too large", a.H);
  if (smem > 48 * 1024)
    MRG_CUDA_CHECK(cudaFuncSetAttribute(rec_fwd_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)smem));
  const int slices = (a.B + GR - 1) / GR;
  ProfScope prof(PROF_REC_FWD, stream);
  count_launch();
  rec_fwd_generic_kernel<<<a.D * slices, 256, smem, stream>>>(a);
  MRG_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int rec_backward_generic(const RecBwdArgs& a, cudaStream_t stream) {
  const size_t smem = (size_t)(2 * GR * a.H + 2 * GR * 4 * a.H) * sizeof(float);
  MRG_REQUIRE(smem <= 200 * 1024, "generic recurrent kernel: hidden size %d too large", a.H);
  if (smem > 48 * 1024)
    MRG_CUDA_CHECK(cudaFuncSetAttribute(rec_bwd_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)smem));
  const int slices = (a.B + GR - 1) / GR;
  ProfScope prof(PROF_REC_BWD, stream);
  count_launch();
  rec_bwd_generic_kernel<<<a.D * slices, 256, smem, stream>>>(a);
  MRG_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace mrg
