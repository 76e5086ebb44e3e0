// Warp-level tensor-core helpers of the attention kernels (mrg_attention_mma.cu): mma.sync m16n8k8 tf32 with fp32
// accumulation and the 3xTF32 operand split.
#pragma once
#include "mrg_common.cuh"

namespace mrg {

__device__ __forceinline__ void am_mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t am_lo(float x, uint32_t hi) { return __float_as_uint(x - __uint_as_float(hi)); }
// tf32 operand of one value: the 3-pass split needs the ROUNDED hi part (lo = x - hi must be small); the one-pass modes
// hand the fp32 bits over as they are — the tensor core ignores the low 13 mantissa bits (truncation, 2^-10 instead of
// 2^-11 relative: inside the bound of the reduced-precision modes) and two integer instructions per operand disappear
template <int PASSES>
__device__ __forceinline__ uint32_t am_hi(float x) { return PASSES == 3 ? tf32_rna(x) : __float_as_uint(x); }

// c += A . B with A given as raw fp32 fragment values and B as two raw fp32 values (split here)
template <int PASSES>
__device__ __forceinline__ void am_mma_split(float (&c)[4], const uint32_t (&ahi)[4], const uint32_t (&alo)[4], float b0,
                                             float b1) {
  const uint32_t bh0 = am_hi<PASSES>(b0), bh1 = am_hi<PASSES>(b1);
  if (PASSES == 3) {
    am_mma(c, alo, bh0, bh1);
    am_mma(c, ahi, am_lo(b0, bh0), am_lo(b1, bh1));
  }
  am_mma(c, ahi, bh0, bh1);
}
template <int PASSES>
__device__ __forceinline__ void am_split_a(const float (&x)[4], uint32_t (&hi)[4], uint32_t (&lo)[4]) {
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    hi[e] = am_hi<PASSES>(x[e]);
    lo[e] = PASSES == 3 ? am_lo(x[e], hi[e]) : 0u;
  }
}

__device__ __forceinline__ void am_cp16(float* dst, const float* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void am_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void am_wait0() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void am_wait1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

}  // namespace mrg
