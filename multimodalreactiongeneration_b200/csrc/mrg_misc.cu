// Small streamed kernels: weight packing (gate interleave), bias-gradient column sums, Philox mask.
#include <stdarg.h>

#include "mrg_common.cuh"

namespace mrg {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }

// ---------------------------------------------------------------------------------------------
// launch counter + event-timed launches
// ---------------------------------------------------------------------------------------------
static unsigned long long g_launches = 0;
static int g_prof_on = 0;
constexpr int PROF_MAX = 8192;
static cudaEvent_t g_ev[PROF_MAX][2];
static int g_ev_kind[PROF_MAX];
static int g_ev_created = 0, g_ev_used = 0;
static char g_kind_name[PROF_KINDS][96];   // name of the last named kernel launched per kind

void count_launch(int n) { __atomic_fetch_add(&g_launches, (unsigned long long)n, __ATOMIC_RELAXED); }

ProfScope::ProfScope(int kind, cudaStream_t st, const char* name) : slot(-1), stream(st) {
  if (!g_prof_on || g_ev_used >= PROF_MAX) return;
  if (name) snprintf(g_kind_name[kind], sizeof(g_kind_name[kind]), "%s", name);
  slot = g_ev_used++;
  if (slot >= g_ev_created) {
    cudaEventCreate(&g_ev[slot][0]);
    cudaEventCreate(&g_ev[slot][1]);
    g_ev_created = slot + 1;
  }
  g_ev_kind[slot] = kind;
  cudaEventRecord(g_ev[slot][0], stream);
}
ProfScope::~ProfScope() {
  if (slot >= 0) cudaEventRecord(g_ev[slot][1], stream);
}

// w_pack[d][j*4+g][i] = w_ih[d][g*H+j][i];  bias_pack[d][j*4+g] = b_ih[g*H+j] + b_hh[g*H+j]
struct PackArgs {
  const float* w_ih[2];
  const float* b_ih[2];
  const float* b_hh[2];
  const float* w_hh[2];
};

__global__ void pack_kernel(PackArgs p, float* __restrict__ w_pack, float* __restrict__ bias_pack,
                            float* __restrict__ whh_pack, float* __restrict__ wcat_pack, int I, int H) {
  const int d = blockIdx.y;
  const int row = blockIdx.x;  // packed row j*4+g
  const int j = row >> 2, g = row & 3;
  const float* src = p.w_ih[d] + (size_t)(g * H + j) * I;
  float* dst = w_pack + ((size_t)d * 4 * H + row) * I;
  const size_t plane = (size_t)gridDim.y * 4 * H * I;   // raw | tf32 hi | lo planes (B operand of mrg_gemm_tc4.cu)
  for (int i = threadIdx.x; i < I; i += blockDim.x) {
    const float v = src[i];
    const float hi = __uint_as_float(tf32_rna(v));
    dst[i] = v;
    dst[plane + i] = hi;
    dst[2 * plane + i] = v - hi;
  }
  if (whh_pack != nullptr) {  // single-step path: h0 * W_hh^T is a second projection GEMM
    const float* hs = p.w_hh[d] + (size_t)(g * H + j) * H;
    float* hd = whh_pack + ((size_t)d * 4 * H + row) * H;
    for (int i = threadIdx.x; i < H; i += blockDim.x) hd[i] = hs[i];
  }
  if (wcat_pack != nullptr) {  // single-step inference: [W_ih | W_hh] rows for ONE projection GEMM over [x | h0]
    const float* hs = p.w_hh[d] + (size_t)(g * H + j) * H;
    float* cd = wcat_pack + ((size_t)d * 4 * H + row) * (I + H);
    for (int i = threadIdx.x; i < I; i += blockDim.x) cd[i] = src[i];
    for (int i = threadIdx.x; i < H; i += blockDim.x) cd[I + i] = hs[i];
  }
  if (threadIdx.x == 0 && bias_pack != nullptr) {
    float b = 0.f;
    if (p.b_ih[d]) b += p.b_ih[d][g * H + j];
    if (p.b_hh[d]) b += p.b_hh[d][g * H + j];
    bias_pack[(size_t)d * 4 * H + row] = b;
  }
}

// dst[b][0 .. I) = x[b][0 .. I), dst[b][I .. I + H) = h[b][0 .. H): the A operand of the merged single-step projection
__global__ void concat_xh_kernel(const float* __restrict__ x, const float* __restrict__ h, float* __restrict__ dst, int B,
                                 int I4, int H4) {
  const int W4 = I4 + H4;
  const long long total = (long long)B * W4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % W4);
    const long long b = i / W4;
    reinterpret_cast<float4*>(dst)[i] = c < I4 ? __ldg(reinterpret_cast<const float4*>(x) + b * I4 + c)
                                               : __ldg(reinterpret_cast<const float4*>(h) + b * H4 + (c - I4));
  }
}

int concat_xh(const float* x, const float* h, float* dst, int B, int I, int H, cudaStream_t stream) {
  const long long total = (long long)B * ((I + H) / 4);
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  count_launch();
  concat_xh_kernel<<<(unsigned)blocks, 256, 0, stream>>>(x, h, dst, B, I / 4, H / 4);
  MRG_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int pack_weights(const mrg_lstm_dir_weights* w, float* w_pack, float* bias_pack, float* whh_pack, int I, int H,
                 int D, cudaStream_t stream, float* wcat_pack) {
  PackArgs p;
  for (int d = 0; d < 2; ++d) {
    p.w_hh[d] = d < D ? w[d].w_hh : nullptr;
    p.w_ih[d] = d < D ? w[d].w_ih : nullptr;
    p.b_ih[d] = d < D ? w[d].b_ih : nullptr;
    p.b_hh[d] = d < D ? w[d].b_hh : nullptr;
  }
  dim3 grid(4 * H, D);
  pack_kernel<<<grid, 128, 0, stream>>>(p, w_pack, bias_pack, whh_pack, wcat_pack, I, H);
  MRG_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return 0;
}

// dst[i0][i1][0..H) (contiguous) = src[i0 * s0 + i1 * s1 + ..]: the batch-first <-> time-major relayout in front of / behind
// an LSTM layer as a row copy (every row is H contiguous floats on both sides; 16-byte accesses, grid-stride)
__global__ void copy_rows_kernel(const float* __restrict__ src, long long s0, long long s1, float* __restrict__ dst, int n0,
                                 int n1, int H4) {
  const long long total = (long long)n0 * n1 * H4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % H4);
    const long long r = i / H4;
    const int i1 = (int)(r % n1);
    const long long i0 = r / n1;
    reinterpret_cast<float4*>(dst)[i] = __ldg(reinterpret_cast<const float4*>(src + i0 * s0 + i1 * s1) + c);
  }
}

__global__ void split_tf32_kernel(const float* __restrict__ w, float* __restrict__ hi, float* __restrict__ lo, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float v = w[i];
    const float h = __uint_as_float(tf32_rna(v));
    hi[i] = h;
    lo[i] = v - h;
  }
}

}  // namespace mrg
extern "C" int mrg_copy_rows(const float* src, long long s0, long long s1, float* dst, int n0, int n1, int H, void* stream) {
  MRG_REQUIRE(src && dst && n0 >= 0 && n1 >= 0 && H > 0 && H % 4 == 0 && s0 % 4 == 0 && s1 % 4 == 0 &&
                  (((uintptr_t)src | (uintptr_t)dst) & 15) == 0,
              "mrg_copy_rows: rows of H %% 4 == 0 floats, 16-byte aligned, strides multiples of 4");
  const long long total = (long long)n0 * n1 * (H / 4);
  if (total == 0) return 0;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  mrg::copy_rows_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(src, s0, s1, dst, n0, n1, H / 4);
  MRG_CUDA_CHECK(cudaGetLastError());
  mrg::count_launch();
  return 0;
}

extern "C" int mrg_split_tf32(const float* w, float* hi, float* lo, size_t n, void* stream) {
  MRG_REQUIRE(w && hi && lo, "mrg_split_tf32: null pointer");
  if (n == 0) return 0;
  size_t blocks = (n + 255) / 256;
  if (blocks > 1184) blocks = 1184;
  mrg::split_tf32_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(w, hi, lo, n);
  MRG_CUDA_CHECK(cudaGetLastError());
  mrg::count_launch();
  return 0;
}
namespace mrg {

// db[g*H+j] (+)= sum_b part[b][j][g]
__global__ void colsum_kernel(const float* __restrict__ part, float* __restrict__ db, int B, int H,
                              int accumulate) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;  // interleaved column j*4+g
  if (n >= 4 * H) return;
  float s = 0.f;
  for (int b = 0; b < B; ++b) s += part[(size_t)b * 4 * H + n];
  const int j = n >> 2, g = n & 3;
  float* o = db + g * H + j;
  *o = accumulate ? *o + s : s;
}

int colsum_deinterleave(const float* part, float* db, int B, int H, int accumulate,
                        cudaStream_t stream) {
  colsum_kernel<<<(4 * H + 127) / 128, 128, 0, stream>>>(part, db, B, H, accumulate);
  MRG_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// T == 1 from zero state: the LSTM layer degenerates to a pointwise cell on the projection
// (c = i*g, h = o*tanh(c); W_hh and the forget gate are inert).  This is every predictor step of the
// reference's rollout.  One thread per (direction, row, unit).
// ---------------------------------------------------------------------------------------------
__global__ void cell_zero_fwd_kernel(float* __restrict__ gates, float* __restrict__ y_ext,
                                     float* __restrict__ c_ext, int B, int H, int D, int train, int has_state,
                                     const float* __restrict__ c0_d0, const float* __restrict__ c0_d1) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long per_dir = (long long)B * H;
  if (idx >= per_dir * D) return;
  const int d = (int)(idx / per_dir);
  const long long r = idx % per_dir;
  float4* gp = reinterpret_cast<float4*>(gates) + idx;                // [D][1][B][H] float4
  const float4 x = *gp;
  const float gi = sigmoid_acc(x.x), gf = sigmoid_acc(x.y), gg = tanhf(x.z), go = sigmoid_acc(x.w);
  const long long out = (long long)d * 2 * per_dir + (d == 0 ? per_dir : 0) + r;  // dir0: slot 1, dir1: slot 0
  const long long init = (long long)d * 2 * per_dir + (d == 0 ? 0 : per_dir) + r; // the other slot holds c0
  float c = gi * gg;
  if (has_state) {  // carried state: the h0 * W_hh term is already in `gates`; c0 from the caller's tensor (inference: no
    const float* c0 = d == 0 ? c0_d0 : c0_d1;   // staging copy) or from the init slot of c_ext
    c = fmaf(gf, c0 ? c0[r] : c_ext[init], c);
  }
  const float h = go * tanhf(c);
  y_ext[out] = h;
  c_ext[out] = c;
  if (train) *gp = make_float4(gi, gf, gg, go);
}

__global__ void cell_zero_bwd_kernel(float* __restrict__ gates, const float* __restrict__ c_ext,
                                     const float* __restrict__ dy, const float* __restrict__ dh_n,
                                     const float* __restrict__ dc_n, float* __restrict__ db_part, int B,
                                     int H, int D) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long per_dir = (long long)B * H;
  if (idx >= per_dir * D) return;
  const int d = (int)(idx / per_dir);
  const long long r = idx % per_dir;
  const int b = (int)(r / H), j = (int)(r % H);
  float4* gp = reinterpret_cast<float4*>(gates) + idx;
  const float4 g = *gp;
  const float c = c_ext[(long long)d * 2 * per_dir + (d == 0 ? per_dir : 0) + r];
  float dh = dh_n ? dh_n[idx] : 0.f;
  if (dy) dh += dy[(long long)b * D * H + (long long)d * H + j];
  const float tc = tanhf(c);
  const float dct = (dc_n ? dc_n[idx] : 0.f) + dh * g.w * (1.f - tc * tc);
  const float4 dp = make_float4(dct * g.z * g.x * (1.f - g.x), 0.f, dct * g.x * (1.f - g.z * g.z),
                                dh * tc * g.w * (1.f - g.w));
  *gp = dp;
  reinterpret_cast<float4*>(db_part)[idx] = dp;                        // [D][B][H][4]
}

int cell_zero_state_forward(float* gates, float* y_ext, float* c_ext, int B, int H, int D, int train,
                            int has_state, cudaStream_t stream, const float* c0_d0, const float* c0_d1) {
  const long long n = (long long)B * H * D;
  ProfScope prof(PROF_REC_FWD, stream);
  count_launch();
  cell_zero_fwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(gates, y_ext, c_ext, B, H, D, train,
                                                                        has_state, c0_d0, c0_d1);
  MRG_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int cell_zero_state_backward(float* gates, const float* c_ext, const float* dy, const float* dh_n,
                             const float* dc_n, float* db_part, int B, int H, int D, cudaStream_t stream) {
  const long long n = (long long)B * H * D;
  ProfScope prof(PROF_REC_BWD, stream);
  count_launch();
  cell_zero_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(gates, c_ext, dy, dh_n, dc_n, db_part,
                                                                        B, H, D);
  MRG_CUDA_CHECK(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11) — must match oracle/philox.py bit for bit.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t philox4x32_10_first(uint32_t c0, uint32_t c1, uint32_t c2,
                                                         uint32_t c3, uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  return c0;
}

__global__ void philox_mask_kernel(uint64_t seed, uint64_t offset, float prob, int T, int B, int shared,
                                   uint8_t* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= T * B) return;
  const int t = idx / B, b = idx % B;
  const uint64_t pos = offset + (uint64_t)t;
  const uint32_t r = philox4x32_10_first((uint32_t)pos, (uint32_t)(pos >> 32), shared ? 0u : (uint32_t)b,
                                         0u, (uint32_t)seed, (uint32_t)(seed >> 32));
  const float u = (float)(r >> 8) * 5.9604644775390625e-08f;  // 2^-24
  out[idx] = u < prob ? 1 : 0;
}

}  // namespace mrg

extern "C" int mrg_philox_mask(uint64_t seed, uint64_t offset, float prob, int T, int B, int shared,
                               uint8_t* out, void* stream) {
  MRG_REQUIRE(T >= 0 && B >= 0 && out != nullptr, "mrg_philox_mask: bad arguments");
  if (T * B == 0) return 0;
  mrg::philox_mask_kernel<<<(T * B + 255) / 256, 256, 0, (cudaStream_t)stream>>>(seed, offset, prob, T, B,
                                                                                shared, out);
  MRG_CUDA_CHECK(cudaGetLastError());
  mrg::count_launch();
  return 0;
}

extern "C" unsigned long long mrg_launch_count(void) { return mrg::g_launches; }

extern "C" int mrg_profile_enable(int on) {
  mrg::g_prof_on = on ? 1 : 0;
  if (on) mrg::g_ev_used = 0;
  return 0;
}

extern "C" const char* mrg_profile_kernel_name(int kind) {
  return (kind >= 0 && kind < mrg::PROF_KINDS) ? mrg::g_kind_name[kind] : "";
}

// Sums the event-timed durations recorded since mrg_profile_enable(1): ms[k], n[k] for k in
// {0: recurrent forward, 1: recurrent backward, 2: GEMM, 3: rollout forward, 4: rollout backward}.
// Synchronises on the recorded events.
extern "C" int mrg_profile_read(float* ms, int* n) {
  for (int k = 0; k < mrg::PROF_KINDS; ++k) { ms[k] = 0.f; n[k] = 0; }
  for (int i = 0; i < mrg::g_ev_used; ++i) {
    MRG_CUDA_CHECK(cudaEventSynchronize(mrg::g_ev[i][1]));
    float t = 0.f;
    MRG_CUDA_CHECK(cudaEventElapsedTime(&t, mrg::g_ev[i][0], mrg::g_ev[i][1]));
    ms[mrg::g_ev_kind[i]] += t;
    n[mrg::g_ev_kind[i]] += 1;
  }
  mrg::g_ev_used = 0;
  return 0;
}

extern "C" int mrg_version(void) { return MRG_VERSION; }
extern "C" const char* mrg_last_error_string(void) { return mrg::last_error(); }

// ---------------------------------------------------------------------------------------------------------
// Column sums out[n] = sum_m x[m][n] — the bias gradient of the Linear layers on the [B*T, N] stream
// (db = dy^T 1).  HBM-bound, deterministic: pass 1 sums slabs of COLSUM_ROWS rows with coalesced 16-byte
// loads into partial[slab][N], pass 2 adds the slabs in a fixed order.
// ---------------------------------------------------------------------------------------------------------
namespace mrg {
constexpr int COLSUM_ROWS = 128;
__global__ void __launch_bounds__(256) colsum_slab_kernel(const float* __restrict__ x, float* __restrict__ partial,
                                                          int M, int N) {
  // blockDim = 256: 64 column quads x 4 row phases; grid = (ceil(N/256), slabs)
  const int cq = threadIdx.x & 63, rp = threadIdx.x >> 6;
  const int n = blockIdx.x * 256 + cq * 4;
  const int m0 = blockIdx.y * COLSUM_ROWS;
  const int m1 = min(M, m0 + COLSUM_ROWS);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (n + 3 < N) {
#pragma unroll 8
    for (int m = m0 + rp; m < m1; m += 4) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(x + (size_t)m * N + n));
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  } else {
    for (int m = m0 + rp; m < m1; m += 4) {
      const float* r = x + (size_t)m * N;
      if (n < N) acc.x += r[n];
      if (n + 1 < N) acc.y += r[n + 1];
      if (n + 2 < N) acc.z += r[n + 2];
    }
  }
  __shared__ float4 red[4][64];
  red[rp][cq] = acc;
  __syncthreads();
  if (rp == 0) {
    float4 s = red[0][cq];
#pragma unroll
    for (int i = 1; i < 4; ++i) { s.x += red[i][cq].x; s.y += red[i][cq].y; s.z += red[i][cq].z; s.w += red[i][cq].w; }
    float* o = partial + (size_t)blockIdx.y * N;
    if (n < N) o[n] = s.x;
    if (n + 1 < N) o[n + 1] = s.y;
    if (n + 2 < N) o[n + 2] = s.z;
    if (n + 3 < N) o[n + 3] = s.w;
  }
}
// pass 1 for rows that are not float4-addressable (N % 4 != 0 or an unaligned base: the 6- / 18-wide pose heads):
// 32 columns x 8 row phases per block, scalar loads
__global__ void __launch_bounds__(256) colsum_slab_scalar_kernel(const float* __restrict__ x,
                                                                 float* __restrict__ partial, int M, int N) {
  __shared__ float red[8][33];
  const int c = threadIdx.x & 31, rp = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + c;
  const int m0 = blockIdx.y * COLSUM_ROWS;
  const int m1 = min(M, m0 + COLSUM_ROWS);
  float acc = 0.f;
  if (n < N)
    for (int m = m0 + rp; m < m1; m += 8) acc += __ldg(x + (size_t)m * N + n);
  red[rp][c] = acc;
  __syncthreads();
  if (rp == 0 && n < N) {
    float s = red[0][c];
#pragma unroll
    for (int i = 1; i < 8; ++i) s += red[i][c];
    partial[(size_t)blockIdx.y * N + n] = s;
  }
}
// pass 2: 32 columns x 8 slab lanes per block; lane l adds slabs l, l+8, ... then the 8 lanes are added in order
__global__ void __launch_bounds__(256) colsum_final_kernel(const float* __restrict__ partial, float* __restrict__ out,
                                                           int slabs, int N, int accumulate) {
  __shared__ float red[8][32];
  const int c = threadIdx.x & 31, l = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + c;
  float s = 0.f;
  if (n < N)
    for (int i = l; i < slabs; i += 8) s += partial[(size_t)i * N + n];
  red[l][c] = s;
  __syncthreads();
  if (l == 0 && n < N) {
    float t = accumulate ? out[n] : 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][c];
    out[n] = t;
  }
}
}  // namespace mrg

extern "C" size_t mrg_colsum_workspace_bytes(int M, int N) {
  if (M <= 0 || N <= 0) return 0;
  return (size_t)((M + mrg::COLSUM_ROWS - 1) / mrg::COLSUM_ROWS) * N * sizeof(float);
}

extern "C" int mrg_colsum(const float* x, float* out, int M, int N, int accumulate, void* workspace,
                          size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MRG_REQUIRE(out && M >= 0 && N > 0 && (x || M == 0), "mrg_colsum: bad arguments");
  if (M == 0) {
    if (!accumulate) MRG_CUDA_CHECK(cudaMemsetAsync(out, 0, (size_t)N * sizeof(float), stream));
    return 0;
  }
  if (workspace == nullptr || workspace_bytes < mrg_colsum_workspace_bytes(M, N)) {
    mrg::set_error("mrg_colsum: workspace too small");
    return MRG_E_WORKSPACE;
  }
  const int slabs = (M + mrg::COLSUM_ROWS - 1) / mrg::COLSUM_ROWS;
  mrg::count_launch(2);
  if (N % 4 == 0 && ((uintptr_t)x & 15) == 0)
    mrg::colsum_slab_kernel<<<dim3((N + 255) / 256, slabs), 256, 0, stream>>>(x, (float*)workspace, M, N);
  else
    mrg::colsum_slab_scalar_kernel<<<dim3((N + 31) / 32, slabs), 256, 0, stream>>>(x, (float*)workspace, M, N);
  MRG_CUDA_CHECK(cudaGetLastError());
  mrg::colsum_final_kernel<<<(N + 31) / 32, 256, 0, stream>>>((const float*)workspace, out, slabs, N, accumulate);
  MRG_CUDA_CHECK(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------------------
// Flat AdamW: the optimizer step of the trainer (reference: torch.optim.AdamW built in configure_optimizers,
// mr_gen/model/simple_lstm/simple_lstm.py:193-221) as ONE streamed kernel over the flat parameter / gradient
// buckets instead of ~230 foreach launches.  HBM-bound: 4 reads + 3 (4 with zero_grad) writes of 4 bytes per
// parameter.  state = {step, 1/(1-b1^step), 1/sqrt(1-b2^step)} lives on the device so that a captured CUDA graph
// advances the step count on every replay.
// ---------------------------------------------------------------------------------------------------------
namespace mrg {
__global__ void adamw_state_kernel(float* state, float b1, float b2) {
  const float step = state[0] + 1.0f;
  state[0] = step;
  state[1] = 1.0f / (1.0f - powf(b1, step));
  state[2] = 1.0f / sqrtf(1.0f - powf(b2, step));
}

__global__ void __launch_bounds__(256) adamw_flat_kernel(float4* __restrict__ p, float4* __restrict__ g,
                                                         float4* __restrict__ m, float4* __restrict__ v,
                                                         size_t n4, const float* __restrict__ lr_dev, float lr_host,
                                                         float b1, float b2, float eps, float wd,
                                                         const float* __restrict__ state, float gscale,
                                                         int zero_grad) {
  const float lr = lr_dev ? *lr_dev : lr_host;
  const float step_size = lr * state[1];
  const float inv_sqrt_bc2 = state[2];
  const float decay = 1.0f - lr * wd;
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 pp = p[i], gg = g[i], mm = m[i], vv = v[i];
    float* P = &pp.x; float* G = &gg.x; float* M = &mm.x; float* V = &vv.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gr = G[k] * gscale;
      M[k] = M[k] + (gr - M[k]) * (1.0f - b1);
      V[k] = V[k] * b2 + gr * gr * (1.0f - b2);
      const float denom = sqrtf(V[k]) * inv_sqrt_bc2 + eps;
      P[k] = P[k] * decay - step_size * (M[k] / denom);
    }
    p[i] = pp; m[i] = mm; v[i] = vv;
    if (zero_grad) g[i] = zero;
  }
}
}  // namespace mrg

extern "C" int mrg_adamw_flat(float* p, float* g, float* m, float* v, size_t n, const float* lr_dev, float lr,
                              float beta1, float beta2, float eps, float weight_decay, float* state,
                              float grad_scale, int zero_grad, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MRG_REQUIRE(p && g && m && v && state, "mrg_adamw_flat: null pointer");
  MRG_REQUIRE(n % 4 == 0, "mrg_adamw_flat: the flat buckets must be padded to a multiple of 4 floats");
  if (n == 0) return 0;
  mrg::adamw_state_kernel<<<1, 1, 0, stream>>>(state, beta1, beta2);
  const size_t n4 = n / 4;
  int blocks = (int)((n4 + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  mrg::count_launch(2);
  mrg::adamw_flat_kernel<<<blocks, 256, 0, stream>>>((float4*)p, (float4*)g, (float4*)m, (float4*)v, n4, lr_dev, lr,
                                                     beta1, beta2, eps, weight_decay, state, grad_scale, zero_grad);
  MRG_CUDA_CHECK(cudaGetLastError());
  return 0;
}

namespace mrg {
static unsigned long long* g_trace_buf = nullptr;
unsigned long long* debug_trace_buffer() { return g_trace_buf; }
}  // namespace mrg
// developer hook: device buffer of 1 + 2*65535 uint64 that -DMRG_REC_TRACE builds of the recurrent kernels fill
extern "C" int mrg_debug_set_trace(unsigned long long* buf) {
  mrg::g_trace_buf = buf;
  return 0;
}
