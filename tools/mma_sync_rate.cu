// Developer microbenchmark: throughput of the warp-level (legacy) mma.sync tensor-core path on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void mma_tf32(float (&c)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
template <int MODE>
__global__ void k(int iters, float* out, long long* clk) {
  float c[8][4] = {};
  unsigned a[4] = {threadIdx.x, 2, 3, 4}, b[2] = {5, threadIdx.x};
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) { if (MODE == 0) mma_tf32(c[j], a, b); else mma_bf16(c[j], a, b); }
  }
  __syncthreads();
  long long t1 = clock64();
  float s = 0; for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}
int main() {
  float* out; long long* clk; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&clk, 8);
  for (int warps : {4, 8, 16}) for (int mode = 0; mode < 2; ++mode) {
    const int iters = 4096;
    if (mode == 0) { k<0><<<148, warps * 32>>>(iters, out, clk); k<0><<<148, warps * 32>>>(iters, out, clk); }
    else { k<1><<<148, warps * 32>>>(iters, out, clk); k<1><<<148, warps * 32>>>(iters, out, clk); }
    long long c = 0; cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost);
    const double mmas = (double)iters * 8 * warps;
    const double fma = mmas * 16 * 8 * (mode == 0 ? 8 : 16);
    printf("%s %2d warps/SM: %.2f clk per MMA per SM, %.0f FMA/clk/SM -> %.0f TFLOP/s at 1.965 GHz x 148 SMs\n",
           mode == 0 ? "mma.sync m16n8k8 tf32 " : "mma.sync m16n8k16 bf16", warps, c / mmas, fma / c, fma / c * 2 * 1.965e9 * 148 / 1e12);
  }
  return 0;
}
