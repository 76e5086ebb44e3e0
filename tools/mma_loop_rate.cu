// Developer microbenchmark: the MMA phase of rec_fwd3 / rec_bwd3 in isolation (one chunk = 8 k-step pairs x 4 n-tiles x 2
// MMAs per warp, B fragments in 128 registers, A fragments from shared memory), to separate the tensor-pipe rate from
// operand delivery.  Variants: 0 = A as the kernels load it (two row loads + register interleave), 1 = A from a
// row-interleaved layout (one 16-byte load IS the fragment), 2 = A constant in registers (no loads), 3 = variant 1 with
// 16 warps.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void mma_tf32(float (&c)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <int MODE>
__global__ void __launch_bounds__(512, 1) k(int iters, float* out, long long* clk, const float* w) {
  __shared__ __align__(16) float h[16][272];
  const int lane = threadIdx.x & 31, g8 = lane >> 2, q = lane & 3;
  for (int i = threadIdx.x; i < 16 * 272; i += blockDim.x) (&h[0][0])[i] = 0.001f * i;
  uint4 wb[8][4];
#pragma unroll
  for (int kp = 0; kp < 8; ++kp)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) wb[kp][nt] = *reinterpret_cast<const uint4*>(w + ((threadIdx.x * 8 + kp) * 4 + nt) * 4);
  __syncthreads();
  float tot = 0.f;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    float acc[4][4], acc2[4][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[nt][e] = acc2[nt][e] = 0.f;
    const float* hr = &h[g8][4 * q + (it & 1) * 128];
#pragma unroll
    for (int kp = 0; kp < 8; ++kp) {
      unsigned a0[4], a1[4];
      if (MODE == 0) {
        const uint4 lo = *reinterpret_cast<const uint4*>(hr + kp * 16);
        const uint4 hi = *reinterpret_cast<const uint4*>(hr + 8 * 272 + kp * 16);
        a0[0] = lo.x; a0[1] = hi.x; a0[2] = lo.y; a0[3] = hi.y;
        a1[0] = lo.z; a1[1] = hi.z; a1[2] = lo.w; a1[3] = hi.w;
      } else if (MODE == 1 || MODE == 3) {
        const uint4 x = *reinterpret_cast<const uint4*>(hr + kp * 16);
        const uint4 y = *reinterpret_cast<const uint4*>(hr + 8 * 272 + kp * 16);
        a0[0] = x.x; a0[1] = x.y; a0[2] = x.z; a0[3] = x.w;
        a1[0] = y.x; a1[1] = y.y; a1[2] = y.z; a1[3] = y.w;
      } else {
        a0[0] = a1[1] = lane; a0[1] = a1[0] = 2; a0[2] = a1[3] = 3; a0[3] = a1[2] = it;
      }
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        mma_tf32(acc[nt], a0, wb[kp][nt].x, wb[kp][nt].y);
        mma_tf32(acc2[nt], a1, wb[kp][nt].z, wb[kp][nt].w);
      }
    }
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) tot += acc[nt][0] + acc2[nt][3] + acc[nt][1] + acc2[nt][2];
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = tot;
  if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}
int main() {
  float *out, *w; long long* clk;
  cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&clk, 8); cudaMalloc(&w, 512 * 128 * 4); cudaMemset(w, 0x3c, 512 * 128 * 4);
  const int iters = 2000;
  for (int mode = 0; mode < 4; ++mode) {
    const int warps = mode == 3 ? 16 : 8;
    for (int rep = 0; rep < 2; ++rep) {
      if (mode == 0) k<0><<<148, warps * 32>>>(iters, out, clk, w);
      if (mode == 1) k<1><<<148, warps * 32>>>(iters, out, clk, w);
      if (mode == 2) k<2><<<148, warps * 32>>>(iters, out, clk, w);
      if (mode == 3) k<3><<<148, warps * 32>>>(iters, out, clk, w);
    }
    long long c = 0; cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost);
    printf("mode %d, %2d warps: %.0f clk per chunk (64 MMAs per warp), %.2f clk per MMA per SM  [%s]\n", mode, warps, (double)c / iters,
           (double)c / iters / (64.0 * warps), cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
