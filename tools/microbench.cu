// Hardware facts the recurrent-kernel design depends on (developer tool, B200):
//   1. FFMA vs FFMA2 (fma.rn.f32x2) issue throughput per SM
//   2. per-step exchange cost in an 8-CTA cluster, 128 threads x 4 B to each of the 8 CTAs:
//      (a) st.shared::cluster + barrier.cluster.arrive.release / wait.acquire
//      (b) st.async ... mbarrier::complete_tx::bytes + local mbarrier try_wait (two barriers, by parity)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void ffma_kernel(float* out, int iters, long long* cyc) {
  float a[8], x = out[threadIdx.x], y = out[threadIdx.x + 1];
  for (int i = 0; i < 8; ++i) a[i] = i;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], x, y);
  }
  long long t1 = clock64();
  float s = 0; for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void ffma2_kernel(float* out, int iters, long long* cyc) {
  unsigned long long a[8], x, y;
  float2 xf = make_float2(out[threadIdx.x], out[threadIdx.x + 1]);
  x = *reinterpret_cast<unsigned long long*>(&xf); y = x;
  for (int i = 0; i < 8; ++i) { float2 v = make_float2(i, i + 1); a[i] = *reinterpret_cast<unsigned long long*>(&v); }
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a[i]) : "l"(x), "l"(y));
  }
  long long t1 = clock64();
  float s = 0; for (int i = 0; i < 8; ++i) { float2 v = *reinterpret_cast<float2*>(&a[i]); s += v.x + v.y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t a, uint32_t r) { uint32_t o; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(o) : "r"(a), "r"(r)); return o; }

// (a) plain remote stores + cluster barrier
__global__ void __cluster_dims__(8, 1, 1) xchg_barrier_kernel(float* out, int iters, long long* cyc, int payload) {
  __shared__ float buf[2][1024];
  uint32_t rank; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  uint32_t remote[8];
  for (int r = 0; r < 8; ++r) remote[r] = mapa(smem_u32(&buf[0][0]), r);
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) (&buf[0][0])[i] = 0.f;
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  float v = threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const int nxt = (it + 1) & 1, cur = it & 1;
    v += buf[cur][(threadIdx.x * 7) & 1023];
    if (payload && threadIdx.x < 128) {
      const uint32_t off = (nxt * 1024 + rank * 128 + threadIdx.x) * 4;
#pragma unroll
      for (int r = 0; r < 8; ++r) asm volatile("st.shared::cluster.f32 [%0], %1;" :: "r"(remote[r] + off), "f"(v) : "memory");
    }
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = v;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

// (b) st.async with complete_tx on the destination's mbarrier, consumers wait on the local mbarrier
__global__ void __cluster_dims__(8, 1, 1) xchg_mbar_kernel(float* out, int iters, long long* cyc) {
  __shared__ float buf[2][1024];
  __shared__ __align__(8) unsigned long long bar[2];
  uint32_t rank; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  uint32_t remote[8], rbar[8];
  for (int r = 0; r < 8; ++r) { remote[r] = mapa(smem_u32(&buf[0][0]), r); rbar[r] = mapa(smem_u32(&bar[0]), r); }
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) (&buf[0][0])[i] = 0.f;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar[1])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    // arm both phases' first use
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&bar[1])), "r"(4096) : "memory");
  }
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  float v = threadIdx.x;
  uint32_t phase[2] = {0, 0};
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const int nxt = (it + 1) & 1, cur = it & 1;
    v += buf[cur][(threadIdx.x * 7) & 1023];
    if (threadIdx.x < 128) {
      const uint32_t off = (nxt * 1024 + rank * 128 + threadIdx.x) * 4;
#pragma unroll
      for (int r = 0; r < 8; ++r)
        asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];"
                     :: "r"(remote[r] + off), "r"(__float_as_uint(v)), "r"(rbar[r] + nxt * 8) : "memory");
    }
    // wait for the 8 x 128 x 4 B of buffer nxt to land here
    const uint32_t b = smem_u32(&bar[nxt]);
    uint32_t done = 0;
    while (!done)
      asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p;}"
                   : "=r"(done) : "r"(b), "r"(phase[nxt]) : "memory");
    phase[nxt] ^= 1;
    __syncthreads();  // everyone has observed the phase before it is re-armed / buffer cur is rewritten
    if (threadIdx.x == 0)  // re-arm the barrier of buffer cur for its next use (iteration it+1 writes cur)
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&bar[cur])), "r"(4096) : "memory");
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = v;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

int main() {
  float* out; long long* cyc; long long h;
  CK(cudaMalloc(&out, 1 << 22)); CK(cudaMemset(out, 0, 1 << 22)); CK(cudaMalloc(&cyc, 8));
  const int iters = 2000;
  for (int warps = 4; warps <= 16; warps *= 2) {
    ffma_kernel<<<148, warps * 32>>>(out, iters, cyc); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
    printf("FFMA  %2d warps/SM: %.1f FMA/clk/SM\n", warps, (double)iters * 64 * warps * 32 / h);
    ffma2_kernel<<<148, warps * 32>>>(out, iters, cyc); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
    printf("FFMA2 %2d warps/SM: %.1f FMA/clk/SM\n", warps, (double)iters * 64 * 2 * warps * 32 / h);
  }
  for (int payload = 0; payload <= 1; ++payload) {
    xchg_barrier_kernel<<<128, 256>>>(out, iters, cyc, payload); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
    printf("cluster barrier exchange (payload=%d): %.0f clk/step\n", payload, (double)h / iters);
  }
  xchg_mbar_kernel<<<128, 256>>>(out, iters, cyc); CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
  printf("st.async + mbarrier exchange: %.0f clk/step\n", (double)h / iters);
  int clk; CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0)); printf("clock rate attr %d kHz\n", clk);
  return 0;
}
