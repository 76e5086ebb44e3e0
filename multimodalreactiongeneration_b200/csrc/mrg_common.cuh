// Shared helpers for the sm_100a kernels of the LSTM hot path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/mrg_lstm.h"

namespace mrg {

void set_error(const char* fmt, ...);
const char* last_error();

#define MRG_CUDA_CHECK(expr)                                                           \
  do {                                                                                 \
    cudaError_t _e = (expr);                                                           \
    if (_e != cudaSuccess) {                                                           \
      mrg::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                     __LINE__);                                                        \
      return (int)_e;                                                                  \
    }                                                                                  \
  } while (0)

#define MRG_REQUIRE(cond, ...)     \
  do {                             \
    if (!(cond)) {                 \
      mrg::set_error(__VA_ARGS__); \
      return MRG_E_INVALID;        \
    }                              \
  } while (0)

// ---------------------------------------------------------------------------------------------
// PTX helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_id_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_arrive_release() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_wait_acquire() {
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  cluster_arrive_release();
  cluster_wait_acquire();
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
// address of the same shared-memory location in CTA `rank` of this cluster
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_f32(uint32_t addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, float4 v) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}

// ---- mbarrier + st.async (remote store that signals the destination CTA's mbarrier) -----------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void st_async_f32(uint32_t remote_addr, float v, uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(
                   remote_addr),
               "r"(__float_as_uint(v)), "r"(remote_bar)
               : "memory");
}
__device__ __forceinline__ void st_async_v4(uint32_t remote_addr, float4 v, uint32_t remote_bar) {
  asm volatile(
      "st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(
          remote_addr),
      "r"(__float_as_uint(v.x)), "r"(__float_as_uint(v.y)), "r"(__float_as_uint(v.z)),
      "r"(__float_as_uint(v.w)), "r"(remote_bar)
      : "memory");
}

// Predicated forms (one instruction under a predicate instead of a branch around it): the software-pipelined
// recurrent kernels need their whole iteration in ONE basic block so that ptxas can interleave the
// latency-bound tail of one chunk with the FFMA2 block of the next.
__device__ __forceinline__ void st_async_v4_if(bool p, uint32_t remote_addr, float4 v, uint32_t remote_bar) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %6, 0;\n"
      "@p st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];\n}" ::"r"(
          remote_addr),
      "r"(__float_as_uint(v.x)), "r"(__float_as_uint(v.y)), "r"(__float_as_uint(v.z)),
      "r"(__float_as_uint(v.w)), "r"(remote_bar), "r"((uint32_t)p)
      : "memory");
}
__device__ __forceinline__ void st_global_v4_if(bool p, float* addr, float4 v) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %5, 0;\n@p st.global.v4.f32 [%0], {%1, %2, %3, %4};\n}" ::"l"(addr),
               "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "r"((uint32_t)p)
               : "memory");
}
__device__ __forceinline__ void st_global_f32_if(bool p, float* addr, float v) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %2, 0;\n@p st.global.f32 [%0], %1;\n}" ::"l"(addr), "f"(v),
               "r"((uint32_t)p)
               : "memory");
}
__device__ __forceinline__ void cp_async16_if(bool p, uint32_t dst, const void* src) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %2, 0;\n@p cp.async.cg.shared.global [%0], [%1], 16;\n}" ::"r"(dst),
               "l"(src), "r"((uint32_t)p)
               : "memory");
}
__device__ __forceinline__ void cp_async4_if(bool p, uint32_t dst, const void* src) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %2, 0;\n@p cp.async.ca.shared.global [%0], [%1], 4;\n}" ::"r"(dst),
               "l"(src), "r"((uint32_t)p)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx_if(bool p, uint32_t bar, uint32_t bytes) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %2, 0;\n@p mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n}" ::"r"(bar),
               "r"(bytes), "r"((uint32_t)p)
               : "memory");
}

// Gate non-linearities of the second-generation kernels: ex2.approx.ftz / rcp.approx.ftz directly (2 MUFU + 2-3
// FP32 ops, no denormal range fix-up): |error| <= ~2e-7 absolute, the same order as one fp32 rounding of the
// pre-activation; saturates correctly (ex2 -> inf gives rcp -> 0).
__device__ __forceinline__ float ex2_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_ftz(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fast_sigmoid(float x) { return rcp_ftz(1.0f + ex2_ftz(x * -1.4426950408889634f)); }
__device__ __forceinline__ float fast_tanh(float x) {
  return fmaf(-2.0f, rcp_ftz(1.0f + ex2_ftz(x * 2.8853900817779268f)), 1.0f);
}

// Developer-only event trace of the recurrent kernels (-DMRG_REC_TRACE): lane 0 of every warp of CTA 0 appends
// (clock, warp, event, chunk, step) records to a global buffer set with mrg_debug_set_trace().
#ifdef MRG_REC_TRACE
// 1024 (clock, packed event) slots per warp, plain stores (no atomics: the trace must not stall the warp)
__device__ __forceinline__ void rec_trace(unsigned long long* buf, unsigned& n, int evt, int ch, int step) {
  if (buf && blockIdx.x == 0 && (threadIdx.x & 31) == 0 && n < 1024u) {
    unsigned long long* p = buf + ((size_t)(threadIdx.x >> 5) * 1024 + n) * 2;
    p[0] = clock64();
    p[1] = ((unsigned long long)(threadIdx.x >> 5) << 48) | ((unsigned long long)evt << 32) |
           ((unsigned long long)ch << 16) | (unsigned long long)step;
    ++n;
  }
}
#define REC_TRACE_DECL unsigned trace_n = 0;
#define REC_TRACE(evt, ch, step) rec_trace(a.trace, trace_n, evt, ch, step)
#else
#define REC_TRACE_DECL
#define REC_TRACE(evt, ch, step)
#endif

// packed fp32x2 FMA (Blackwell FFMA2): d = a * b + d on both halves
__device__ __forceinline__ void ffma2(float2& d, const float2& a, const float2& b) {
  unsigned long long dd = *reinterpret_cast<unsigned long long*>(&d);
  const unsigned long long aa = *reinterpret_cast<const unsigned long long*>(&a);
  const unsigned long long bb = *reinterpret_cast<const unsigned long long*>(&b);
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(dd) : "l"(aa), "l"(bb));
  d = *reinterpret_cast<float2*>(&dd);
}

__device__ __forceinline__ float2 fmul2(const float2& a, const float2& b) {
  unsigned long long dd;
  const unsigned long long aa = *reinterpret_cast<const unsigned long long*>(&a);
  const unsigned long long bb = *reinterpret_cast<const unsigned long long*>(&b);
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(dd) : "l"(aa), "l"(bb));
  return *reinterpret_cast<float2*>(&dd);
}

__device__ __forceinline__ float sigmoid_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

// Gate non-linearities on the serial critical path of the cluster kernels: ex2.approx based, absolute
// error <= ~2e-7 (same order as one fp32 rounding of the pre-activation), ~8 instructions instead of ~30.
// -DMRG_ACCURATE_GATES switches back to expf / tanhf / IEEE division.
#ifdef MRG_ACCURATE_GATES
__device__ __forceinline__ float gate_sigmoid(float x) { return sigmoid_acc(x); }
__device__ __forceinline__ float gate_tanh(float x) { return tanhf(x); }
#else
__device__ __forceinline__ float gate_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float gate_tanh(float x) {
  return 1.0f - __fdividef(2.0f, 1.0f + __expf(2.0f * x));
}
#endif

// ---------------------------------------------------------------------------------------------
// internal entry points (host)
// ---------------------------------------------------------------------------------------------
struct GemmArgs {
  const float* a;  // A(m,k) = a[m*a_sm + k*a_sk]
  long long a_sm, a_sk;
  const float* b;  // B(k,n) = b[k*b_sk + n*b_sn]
  long long b_sk, b_sn;
  const float* bias;  // [N] or nullptr
  float* c;           // C(m,n) = c[row(m)*ldc + n]
  long long ldc;
  int M, N, K;
  int accumulate;      // C += result
  int row_deinterleave_H;  // >0: output row m=(j*4+g) is stored at row g*H+j (gate de-interleave)
  int single_pass;         // tensor-core path only: one tf32 pass (MRG_F_TF32) instead of 3xTF32
  int a_bf16;              // bf16 mode: the A operand is stored as bfloat16 (a points to uint16 data, strides in elements)
  int c_bf16;              // bf16 mode: C is stored as bfloat16 (round to nearest even); no accumulate, no de-interleave
  const float* b_hi;       // optional: B already split into tf32 hi / lo planes (same strides as b; weights, split once
  const float* b_lo;       // per step by pack_kernel / mrg_split_tf32) -> the persistent 128 x 256 kernel (mrg_gemm_tc4.cu)
};

// bfloat16 storage helpers of the bf16 mode (the reserve of the recurrent kernels and the GEMM operands next to it)
__device__ __forceinline__ float bf16_lo_to_f32(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi_to_f32(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {   // low half = lo, high half = hi (RNE)
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint2 pack_bf16x4(float a, float b, float c, float d) {
  return make_uint2(pack_bf16x2(a, b), pack_bf16x2(c, d));
}
__device__ __forceinline__ float4 unpack_bf16x4(uint2 v) {
  return make_float4(bf16_lo_to_f32(v.x), bf16_hi_to_f32(v.x), bf16_lo_to_f32(v.y), bf16_hi_to_f32(v.y));
}

// hi = x rounded to TF32 (nearest, ties away: add half an ulp of the 10-bit mantissa, clear the low 13
// bits) with two full-rate integer ops instead of cvt.rna.tf32.f32 (a quarter-rate conversion-pipe op).
// lo = x - hi is exact in fp32 and is handed to the tensor core as is: it ignores the low 13 mantissa
// bits of a tf32 operand, an error of 2^-10 relative to lo, i.e. 2^-21 relative to x.
__device__ __forceinline__ uint32_t tf32_rna(float x) { return (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u; }

int gemm_simt(const GemmArgs& g, void* workspace, size_t workspace_bytes, cudaStream_t stream);
size_t gemm_simt_workspace_bytes(int M, int N, int K);

// w_pack: three planes of [D][4H][I] floats — gate-interleaved W_ih as is, its tf32 hi part, its lo part (x - hi)
int pack_weights(const mrg_lstm_dir_weights* w, float* w_pack, float* bias_pack, float* whh_pack, int I, int H,
                 int D, cudaStream_t stream, float* wcat_pack = nullptr);
int concat_xh(const float* x, const float* h, float* dst, int B, int I, int H, cudaStream_t stream);

struct RecArgs {
  float* gates;        // [D][T][B][H][4]
  const float* w_hh[2];
  float* y_ext;        // [D][T+1][B][H]
  float* c_ext;        // [D][T+1][B][H]
  int T, B, H, D;
  int train;
  unsigned long long* trace;  // developer event trace (-DMRG_REC_TRACE builds), else nullptr
  int cluster_budget;         // > 0: use at most this many clusters (MRG_F_CLUSTER_BUDGET)
  int bf16_gates;             // bf16 mode: `gates` holds 4 x bfloat16 per hidden unit (8 bytes) instead of 4 x fp32
  int gru;                    // MRG_F_GRU: the four gate rows are (r, z, 0, n) of a GRU, the tail applies the GRU cell
};
unsigned long long* debug_trace_buffer();
int rec_forward_generic(const RecArgs& a, cudaStream_t stream);

struct RecBwdArgs {
  float* gates;         // in: gates, out: dpre   [D][T][B][H][4]
  const float* w_hh[2];
  const float* y_ext;
  const float* c_ext;
  const float* dy;      // [T][B][D*H] or nullptr
  const float* dh_n;    // [D][B][H] or nullptr
  const float* dc_n;    // [D][B][H] or nullptr
  float* dh0[2];        // [B][H] or nullptr
  float* dc0[2];
  float* db_part;       // [D][B][H][4] per-row bias-gradient partials (sum over t)
  int T, B, H, D;
  int cluster_budget;   // same meaning as in RecArgs
  int bf16_gates;       // same meaning as in RecArgs: gates in, d(pre-activations) out, both bfloat16
  int gru;              // same meaning as in RecArgs
};
int rec_backward_generic(const RecBwdArgs& a, cudaStream_t stream);
// cluster kernels (chunk-pipelined, H in {128, 256})
int rec_forward_cluster2(const RecArgs& a, cudaStream_t stream);
int rec_backward_cluster2(const RecBwdArgs& a, cudaStream_t stream);
bool rec2_supported(int H);
// tensor-core forward recurrence of the reduced-precision modes (mrg_rec_fwd3.cu)
bool rec_forward_mma_applies(const RecArgs& a, int* slices, int* nch);
int rec_forward_cluster3(const RecArgs& a, int slices, int nch, cudaStream_t stream);
// tensor-core BPTT of the reduced-precision modes (mrg_rec_bwd3.cu)
bool rec_backward_mma_applies(const RecBwdArgs& a, int* slices, int* nch);
int rec_backward_cluster3(const RecBwdArgs& a, int slices, int nch, cudaStream_t stream);
int max_active_clusters2(int H);
int rec2_max_chunks(int H, int rbc);
void pick_partition2(int H, int B, int D, int budget, int* slices_out, int* nch_out, int* rbc_out);

int cell_zero_state_forward(float* gates, float* y_ext, float* c_ext, int B, int H, int D, int train,
                            int has_state, cudaStream_t stream, const float* c0_d0 = nullptr, const float* c0_d1 = nullptr);
int cell_zero_state_backward(float* gates, const float* c_ext, const float* dy, const float* dh_n,
                             const float* dc_n, float* db_part, int B, int H, int D, cudaStream_t stream);

int colsum_deinterleave(const float* part, float* db, int B, int H, int accumulate,
                        cudaStream_t stream);
// fused attention (mrg_attention.cu)
struct AttnArgs {
  const float *q, *k, *v;
  float* o;
  float* lse;   // [B, heads, Tq]: m + log2(sum), log2 domain
  const float* dout;
  float* dvec;  // [B, heads, Tq]: dO . O
  float *dq, *dk, *dv;
  int B, nh, Tq, Tk;
  int ldq, ldk, ldv, ldo, lddo, lddq, lddk, lddv;  // row strides in floats; batch stride = T * ld
  float scale, scale_log2;
  int mask_mode, rate;
  const unsigned char *pad_q, *pad_k;  // [B, Tq], [B, Tk] or both null
};

// launch accounting / live kernel timing (bench.py's gpu_launches and roofline numbers)
void count_launch(int n = 1);
struct ProfScope {  // brackets a launch with CUDA events on its stream when profiling is enabled
  ProfScope(int kind, cudaStream_t stream, const char* kernel_name = nullptr);
  ~ProfScope();
  int slot;
  cudaStream_t stream;
};
enum { PROF_REC_FWD = 0, PROF_REC_BWD = 1, PROF_GEMM = 2, PROF_ROLLOUT_FWD = 3, PROF_ROLLOUT_BWD = 4, PROF_KINDS = 5 };

}  // namespace mrg
