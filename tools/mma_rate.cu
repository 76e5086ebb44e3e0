// Developer microbenchmark: back-to-back tcgen05.mma issue rate on one SM (and on all SMs at once) for the operand
// forms the projection GEMM can use.  Operands are whatever shared / tensor memory holds (zeros): only timing matters.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I multimodalreactiongeneration_b200/csrc -o tools/mma_rate tools/mma_rate.cu
#include <cstdio>
#include "mrg_tc_common.cuh"
namespace mrg { void set_error(const char*, ...) {} }
using namespace mrg;

__device__ __forceinline__ void umma_f16_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_f16_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

// mode 0: tf32 .ss N   | 1: tf32 .ts N | 2: bf16 .ss N | 3: bf16 .ts N
__global__ void __launch_bounds__(128, 1) rate_kernel(int mode, int N, int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar = base + 96 * 1024, slot = bar + 8;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init_fence(); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(slot) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem_raw + (slot - smem_u32(smem_raw)));
  if (threadIdx.x == 0) {
    const bool f16 = mode >= 2;
    // D=f32 (1<<4), A/B format: tf32 = 2, bf16 = 1 at bits 7 / 10; K-major both; N>>3 at 17, M>>4 at 24
    const uint32_t fmt = f16 ? 1u : 2u;
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t da = make_smem_desc(base, 16, 1024, 2), db = make_smem_desc(base + 32 * 1024, 16, 1024, 2);
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const uint32_t koff = (i & 3) * 32;   // walk the 4 k-steps of a 128-byte swizzled row, like the GEMM
      if (mode == 0) umma_tf32(tmem, da + (koff >> 4), db + (koff >> 4), idesc, 1);
      else if (mode == 1) umma_tf32_ts(tmem, tmem + 256 + (i & 3) * 8, db + (koff >> 4), idesc, 1);
      else if (mode == 2) umma_f16_ss(tmem, da + (koff >> 4), db + (koff >> 4), idesc, 1);
      else umma_f16_ts(tmem, tmem + 256 + (i & 3) * 8, db + (koff >> 4), idesc, 1);
    }
    umma_commit(bar);
    mbar_wait(bar, 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

int main() {
  long long* out;
  cudaMalloc(&out, 8);
  const int smem = 96 * 1024 + 1024 + 64;
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const char* names[4] = {"tf32 .ss", "tf32 .ts", "bf16 .ss", "bf16 .ts"};
  for (int grid : {1, 148})
    for (int mode = 0; mode < 4; ++mode)
      for (int N : {64, 128, 256}) {
        const int iters = 2048;
        rate_kernel<<<grid, 128, smem>>>(mode, N, iters, out);
        rate_kernel<<<grid, 128, smem>>>(mode, N, iters, out);
        long long clk = 0;
        cudaError_t e = cudaMemcpy(&clk, out, 8, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        const double per = (double)clk / iters;
        const int K = mode >= 2 ? 16 : 8;
        printf("grid %3d  %s  M=128 N=%3d K=%2d: %7.1f clk per MMA  -> %6.0f FMA/clk/SM\n", grid, names[mode], N, K, per,
               128.0 * N * K / per);
      }
  return 0;
}
