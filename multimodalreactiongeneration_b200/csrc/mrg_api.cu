// C-ABI entry points (include/mrg_lstm.h): orchestration of pack -> projection GEMM -> recurrent
// kernel for the forward, and recurrent BPTT kernel -> dX / dW GEMMs -> bias column sums for the
// backward.  Everything is queued on the caller's stream; nothing synchronises.
#include <cstdint>
#include <cstdlib>

#include "mrg_common.cuh"

namespace mrg {
int gemm_tc2(const GemmArgs& g, void* workspace, size_t workspace_bytes, cudaStream_t stream);
int gemm_tc4(const GemmArgs& g, cudaStream_t stream);
bool gemm_tc4_supported(const GemmArgs& g);
bool gemm_tc_supported(const GemmArgs& g);
size_t gemm_tc_workspace_bytes(int M, int N, int K);

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static int run_gemm(const GemmArgs& g_in, void* ws, size_t ws_bytes, int flags, cudaStream_t stream) {
  GemmArgs g = g_in;
  g.single_pass = (flags & (MRG_F_TF32 | MRG_F_BF16)) ? 1 : 0;
  static int no_tc4 = -1;
  if (no_tc4 < 0) {
    const char* e = getenv("MRG_NO_TC4");   // developer switch: keep the 128 x 128 kernel for every GEMM
    no_tc4 = (e && e[0] == '1') ? 1 : 0;
  }
  if (!(flags & MRG_F_SIMT_GEMM) && !no_tc4 && g.K > 0 && gemm_tc4_supported(g)) return gemm_tc4(g, stream);
  if (!(flags & MRG_F_SIMT_GEMM) && gemm_tc_supported(g)) return gemm_tc2(g, ws, ws_bytes, stream);
  return gemm_simt(g, ws, ws_bytes, stream);
}

static size_t gemm_ws(int M, int N, int K) {
  size_t s = gemm_simt_workspace_bytes(M, N, K);
  const size_t t = gemm_tc_workspace_bytes(M, N, K);
  if (t > s) s = t;
  return s;
}

static int check_device() {
  static int ok = -1;
  if (ok < 0) {
    int dev = 0;
    cudaDeviceProp p;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&p, dev) != cudaSuccess) {
      cudaGetLastError();
      ok = 0;
    } else {
      ok = (p.major == 10) ? 1 : 0;
    }
  }
  return ok;
}

}  // namespace mrg

using namespace mrg;

extern "C" int mrg_device_info(int* sm_count, int* max_clusters_h256, int* max_clusters_h128,
                               int* cc_major, int* cc_minor) {
  int dev = 0;
  MRG_CUDA_CHECK(cudaGetDevice(&dev));
  cudaDeviceProp p;
  MRG_CUDA_CHECK(cudaGetDeviceProperties(&p, dev));
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  if (max_clusters_h256) *max_clusters_h256 = max_active_clusters2(256);
  if (max_clusters_h128) *max_clusters_h128 = max_active_clusters2(128);
  return 0;
}

extern "C" size_t mrg_lstm_workspace_bytes(int T, int B, int I, int H, int D) {
  if (T < 0 || B <= 0 || I <= 0 || H <= 0 || D < 1 || D > 2) return 0;
  const size_t head = align_up((size_t)D * 4 * H * sizeof(float), 256) +
                      align_up((size_t)D * B * 4 * H * sizeof(float), 256) +
                      (T == 1 ? align_up((size_t)D * 4 * H * H * sizeof(float), 256) +
                                    align_up((size_t)D * 4 * H * (I + H) * sizeof(float), 256) +   // [W_ih | W_hh] pack
                                    align_up((size_t)D * B * (I + H) * sizeof(float), 256)         // [x | h0] staging
                              : 0);
  size_t g = gemm_ws(T * B, 4 * H, I);
  if (T == 1) {
    const size_t m = gemm_ws(B, 4 * H, I + H);
    if (m > g) g = m;
  }
  size_t v = gemm_ws(T * B, I, 4 * H);
  if (v > g) g = v;
  v = gemm_ws(4 * H, I, T * B);
  if (v > g) g = v;
  v = gemm_ws(4 * H, H, T * B);
  if (v > g) g = v;
  return head + align_up(g, 256) + 256;
}

static int check_shape(const char* who, int T, int B, int I, int H, int D) {
  MRG_REQUIRE(T >= 0 && B > 0 && I > 0 && H > 0 && (D == 1 || D == 2),
              "%s: bad shape T=%d B=%d I=%d H=%d D=%d", who, T, B, I, H, D);
  MRG_REQUIRE(check_device() == 1, "%s: this library only runs on compute capability 10.x (B200)", who);
  return 0;
}

extern "C" int mrg_lstm_layer_forward(const float* x, const mrg_lstm_dir_weights* w, float* w_pack,
                                      float* gates, float* y_ext, float* c_ext, void* workspace,
                                      size_t workspace_bytes, int T, int B, int I, int H, int D, int flags,
                                      void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (int e = check_shape("mrg_lstm_layer_forward", T, B, I, H, D)) return e;
  MRG_REQUIRE(w && w_pack && gates && y_ext && c_ext && (x || T == 0), "mrg_lstm_layer_forward: null pointer");
  for (int d = 0; d < D; ++d)
    MRG_REQUIRE(w[d].w_ih && w[d].w_hh, "mrg_lstm_layer_forward: null weights for direction %d", d);
  if (workspace_bytes < mrg_lstm_workspace_bytes(T, B, I, H, D) || workspace == nullptr) {
    set_error("mrg_lstm_layer_forward: workspace too small");
    return MRG_E_WORKSPACE;
  }
  const bool bf16 = (flags & MRG_F_BF16) != 0;
  MRG_REQUIRE(!bf16 || (T > 1 && rec2_supported(H) && !(flags & MRG_F_GENERIC_REC)),
              "mrg_lstm_layer_forward: MRG_F_BF16 needs the cluster kernels (H in {128, 256}, T > 1)");
  const bool gru = (flags & MRG_F_GRU) != 0;
  MRG_REQUIRE(!gru || (T > 1 && rec2_supported(H) && !(flags & MRG_F_GENERIC_REC)),
              "mrg_lstm_layer_forward: MRG_F_GRU needs the cluster kernels (H in {128, 256}, T > 1)");
  char* ws = (char*)workspace;
  float* bias_pack = (float*)ws;
  ws += align_up((size_t)D * 4 * H * sizeof(float), 256);
  ws += align_up((size_t)D * B * 4 * H * sizeof(float), 256);
  // T == 1: no recurrence to run — the step is projection GEMM(s) + a pointwise cell
  bool single_zero = (T == 1) && (flags & MRG_F_ZERO_STATE);
  for (int d = 0; d < D; ++d) single_zero = single_zero && !w[d].h0 && !w[d].c0;
  const bool single_state = (T == 1) && !single_zero;
  float* whh_pack = nullptr;
  float* wcat_pack = nullptr;   // single-step inference with carried state: ONE projection over [x | h0] (below)
  float* xh = nullptr;
  if (T == 1) {
    if (single_state) whh_pack = (float*)ws;
    ws += align_up((size_t)D * 4 * H * H * sizeof(float), 256);
    if (single_state) wcat_pack = (float*)ws;
    ws += align_up((size_t)D * 4 * H * (I + H) * sizeof(float), 256);
    xh = (float*)ws;
    ws += align_up((size_t)D * B * (I + H) * sizeof(float), 256);
  }
  const size_t ws_left = workspace_bytes - (size_t)(ws - (char*)workspace);

  // MRG_F_PACK_VALID: w_pack and the head of the workspace (bias pack, W_hh pack of the single-step path) still hold the
  // packs an earlier call with the same weights, shape and buffers wrote (inference with frozen weights): skip the launch
  if (!(flags & MRG_F_PACK_VALID))
    if (int e = pack_weights(w, w_pack, bias_pack, whh_pack, I, H, D, stream, wcat_pack)) return e;
  const size_t slot = (size_t)B * H;
  // single-step inference: nobody reads the init slots — a zero-state step needs none, a carried-state step takes h0 / c0
  // straight from the caller's tensors (no staging copy); training keeps them (the backward reads h_{-1}, c_{-1} there)
  bool direct_state = single_state && !(flags & MRG_F_TRAIN);
  for (int d = 0; d < D; ++d) direct_state = direct_state && w[d].h0 && w[d].c0;
  const bool skip_init = (single_zero && !(flags & MRG_F_TRAIN)) || direct_state;
  for (int d = 0; d < D && !skip_init; ++d) {
    float* ys = y_ext + (size_t)d * (T + 1) * slot + (d == 0 ? 0 : (size_t)T * slot);
    float* cs = c_ext + (size_t)d * (T + 1) * slot + (d == 0 ? 0 : (size_t)T * slot);
    if (w[d].h0) MRG_CUDA_CHECK(cudaMemcpyAsync(ys, w[d].h0, slot * sizeof(float), cudaMemcpyDeviceToDevice, stream));
    else MRG_CUDA_CHECK(cudaMemsetAsync(ys, 0, slot * sizeof(float), stream));
    if (w[d].c0) MRG_CUDA_CHECK(cudaMemcpyAsync(cs, w[d].c0, slot * sizeof(float), cudaMemcpyDeviceToDevice, stream));
    else MRG_CUDA_CHECK(cudaMemsetAsync(cs, 0, slot * sizeof(float), stream));
  }
  if (T == 0) return 0;
  // carried-state single step in inference: gates = [x | h0] . [W_ih | W_hh]^T + b as ONE GEMM (the two-GEMM form below
  // pays launch + fill + drain twice, and the second one reads the output back to accumulate)
  if (direct_state && I % 4 == 0 && H % 4 == 0 && ((uintptr_t)x & 15) == 0) {
    bool aligned = true;
    for (int d = 0; d < D; ++d) aligned = aligned && ((uintptr_t)w[d].h0 & 15) == 0;
    if (aligned) {
      for (int d = 0; d < D; ++d) {
        float* a_cat = xh + (size_t)d * B * (I + H);
        if (int e = concat_xh(x, w[d].h0, a_cat, B, I, H, stream)) return e;
        GemmArgs g = {};
        g.a = a_cat; g.a_sm = I + H; g.a_sk = 1;
        g.b = wcat_pack + (size_t)d * 4 * H * (I + H); g.b_sk = 1; g.b_sn = I + H;
        g.bias = bias_pack + (size_t)d * 4 * H;
        g.c = gates + (size_t)d * B * 4 * H; g.ldc = 4 * H;
        g.M = B; g.N = 4 * H; g.K = I + H;
        if (int e = run_gemm(g, ws, ws_left, flags, stream)) return e;
      }
      return cell_zero_state_forward(gates, y_ext, c_ext, B, H, D, 0, 1, stream, w[0].c0, D > 1 ? w[1].c0 : nullptr);
    }
  }
  for (int d = 0; d < D; ++d) {
    GemmArgs g = {};
    g.a = x; g.a_sm = I; g.a_sk = 1;
    g.b = w_pack + (size_t)d * 4 * H * I; g.b_sk = 1; g.b_sn = I;
    g.b_hi = g.b + (size_t)D * 4 * H * I; g.b_lo = g.b_hi + (size_t)D * 4 * H * I;
    g.bias = bias_pack + (size_t)d * 4 * H;
    // bf16 mode: the reserve is bfloat16, direction d starts half as many bytes in
    g.c = bf16 ? reinterpret_cast<float*>(reinterpret_cast<unsigned short*>(gates) + (size_t)d * T * B * 4 * H)
               : gates + (size_t)d * T * B * 4 * H;
    g.ldc = 4 * H; g.c_bf16 = bf16 ? 1 : 0;
    g.M = T * B; g.N = 4 * H; g.K = I;
    if (int e = run_gemm(g, ws, ws_left, flags, stream)) return e;
  }
  if (single_zero)
    return cell_zero_state_forward(gates, y_ext, c_ext, B, H, D, (flags & MRG_F_TRAIN) ? 1 : 0, 0, stream);
  if (single_state) {
    // gates += h0 * W_hh^T (second projection), then the pointwise cell with the carried c0
    for (int d = 0; d < D; ++d) {
      GemmArgs g = {};
      g.a = direct_state ? w[d].h0 : y_ext + (size_t)d * 2 * slot + (d == 0 ? 0 : slot); g.a_sm = H; g.a_sk = 1;
      g.b = whh_pack + (size_t)d * 4 * H * H; g.b_sk = 1; g.b_sn = H;
      g.c = gates + (size_t)d * B * 4 * H; g.ldc = 4 * H;
      g.M = B; g.N = 4 * H; g.K = H; g.accumulate = 1;
      if (int e = run_gemm(g, ws, ws_left, flags, stream)) return e;
    }
    return cell_zero_state_forward(gates, y_ext, c_ext, B, H, D, (flags & MRG_F_TRAIN) ? 1 : 0, 1, stream,
                                   direct_state ? w[0].c0 : nullptr,
                                   direct_state && D > 1 ? w[1].c0 : nullptr);
  }
  RecArgs r = {};
  r.gates = gates;
  r.w_hh[0] = w[0].w_hh;
  r.w_hh[1] = D > 1 ? w[1].w_hh : nullptr;
  r.y_ext = y_ext; r.c_ext = c_ext;
  r.T = T; r.B = B; r.H = H; r.D = D;
  r.train = (flags & MRG_F_TRAIN) ? 1 : 0;
  r.trace = debug_trace_buffer();
  r.cluster_budget = (flags >> 16) & 0xFF;
  r.bf16_gates = bf16 ? 1 : 0;
  r.gru = gru ? 1 : 0;
  if (!(flags & MRG_F_GENERIC_REC) && rec2_supported(H)) {
    int sl3 = 0, nc3 = 0;   // reduced-precision modes with many rows per cluster: h W_hh^T on the tensor cores
    if ((flags & (MRG_F_TF32 | MRG_F_BF16)) && rec_forward_mma_applies(r, &sl3, &nc3))
      return rec_forward_cluster3(r, sl3, nc3, stream);
    return rec_forward_cluster2(r, stream);
  }
  return rec_forward_generic(r, stream);
}

extern "C" int mrg_lstm_layer_backward(const float* x, const mrg_lstm_dir_weights* w, const float* w_pack,
                                       const float* dy, const float* dh_n, const float* dc_n, float* gates,
                                       const float* y_ext, const float* c_ext, float* dx,
                                       const mrg_lstm_dir_grads* g, void* workspace, size_t workspace_bytes,
                                       int T, int B, int I, int H, int D, int flags, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (int e = check_shape("mrg_lstm_layer_backward", T, B, I, H, D)) return e;
  MRG_REQUIRE(w && w_pack && gates && y_ext && c_ext && g && (x || T == 0),
              "mrg_lstm_layer_backward: null pointer");
  if (workspace_bytes < mrg_lstm_workspace_bytes(T, B, I, H, D) || workspace == nullptr) {
    set_error("mrg_lstm_layer_backward: workspace too small");
    return MRG_E_WORKSPACE;
  }
  const bool bf16 = (flags & MRG_F_BF16) != 0;
  MRG_REQUIRE(!bf16 || (T > 1 && rec2_supported(H) && !(flags & MRG_F_GENERIC_REC)),
              "mrg_lstm_layer_backward: MRG_F_BF16 needs the cluster kernels (H in {128, 256}, T > 1)");
  char* ws = (char*)workspace;
  ws += align_up((size_t)D * 4 * H * sizeof(float), 256);
  float* db_part = (float*)ws;
  ws += align_up((size_t)D * B * 4 * H * sizeof(float), 256);
  if (T == 1) ws += align_up((size_t)D * 4 * H * H * sizeof(float), 256);
  const size_t ws_left = workspace_bytes - (size_t)(ws - (char*)workspace);
  const int acc = (flags & (MRG_F_ACCUMULATE | MRG_F_ACC_WEIGHTS)) ? 1 : 0;  // weight matrices
  const int acc_b = (flags & MRG_F_ACCUMULATE) ? 1 : 0;                      // bias

  RecBwdArgs r = {};
  r.gates = gates;
  r.w_hh[0] = w[0].w_hh;
  r.w_hh[1] = D > 1 ? w[1].w_hh : nullptr;
  r.y_ext = y_ext; r.c_ext = c_ext;
  r.dy = dy; r.dh_n = dh_n; r.dc_n = dc_n;
  for (int d = 0; d < 2; ++d) {
    r.dh0[d] = d < D ? g[d].dh0 : nullptr;
    r.dc0[d] = d < D ? g[d].dc0 : nullptr;
  }
  r.db_part = db_part;
  r.T = T; r.B = B; r.H = H; r.D = D;
  r.cluster_budget = (flags >> 16) & 0xFF;
  r.bf16_gates = bf16 ? 1 : 0;
  r.gru = (flags & MRG_F_GRU) ? 1 : 0;
  MRG_REQUIRE(!r.gru || (T > 1 && rec2_supported(H) && !(flags & MRG_F_GENERIC_REC)),
              "mrg_lstm_layer_backward: MRG_F_GRU needs the cluster kernels (H in {128, 256}, T > 1)");
  int e = 0;
  // two-phase backward: MRG_F_BWD_NO_WGRAD = BPTT + bias sums + dX (what the previous layer waits for),
  // MRG_F_BWD_WGRAD_ONLY = the two weight-gradient GEMMs from the d(pre-activations) an earlier NO_WGRAD call left in
  // `gates` (the caller may queue it on another stream so that it overlaps the previous layer's BPTT)
  const bool do_rec = !(flags & MRG_F_BWD_WGRAD_ONLY), do_wgrad = !(flags & MRG_F_BWD_NO_WGRAD);
  bool pointwise = (T == 1) && (flags & MRG_F_ZERO_STATE);
  for (int d = 0; d < D; ++d) pointwise = pointwise && !g[d].dh0 && !g[d].dc0;
  if (!do_rec) e = 0;
  else if (pointwise) e = cell_zero_state_backward(gates, c_ext, dy, dh_n, dc_n, db_part, B, H, D, stream);
  else if (!(flags & MRG_F_GENERIC_REC) && rec2_supported(H)) {
    int sl3 = 0, nc3 = 0;   // reduced-precision modes with many rows per cluster: dpre W_hh on the tensor cores
    if ((flags & (MRG_F_TF32 | MRG_F_BF16)) && rec_backward_mma_applies(r, &sl3, &nc3))
      e = rec_backward_cluster3(r, sl3, nc3, stream);
    else e = rec_backward_cluster2(r, stream);
  }
  else e = rec_backward_generic(r, stream);
  if (e) return e;

  const size_t slot = (size_t)B * H;
  for (int d = 0; d < D; ++d) {
    const float* dpre = bf16 ? reinterpret_cast<const float*>(reinterpret_cast<const unsigned short*>(gates) +
                                                               (size_t)d * T * B * 4 * H)
                             : gates + (size_t)d * T * B * 4 * H;
    if (g[d].db && do_rec)
      if ((e = colsum_deinterleave(db_part + (size_t)d * B * 4 * H, g[d].db, B, H, acc_b, stream))) return e;
    if (g[d].dw_ih && do_wgrad) {
      GemmArgs m = {};
      m.a = dpre; m.a_sm = 1; m.a_sk = 4 * H;
      m.b = x; m.b_sk = I; m.b_sn = 1;
      m.c = g[d].dw_ih; m.ldc = I;
      m.M = 4 * H; m.N = I; m.K = T * B;
      m.accumulate = acc; m.row_deinterleave_H = H; m.a_bf16 = bf16 ? 1 : 0;
      if ((e = run_gemm(m, ws, ws_left, flags, stream))) return e;
    }
    if (!do_wgrad) {
    } else if (g[d].dw_hh && pointwise) {  // h_prev = 0: no contribution
      if (!acc) MRG_CUDA_CHECK(cudaMemsetAsync(g[d].dw_hh, 0, (size_t)4 * H * H * sizeof(float), stream));
    } else if (g[d].dw_hh) {
      const float* hprev = y_ext + (size_t)d * (T + 1) * slot + (d == 0 ? 0 : slot);
      GemmArgs m = {};
      m.a = dpre; m.a_sm = 1; m.a_sk = 4 * H;
      m.b = hprev; m.b_sk = H; m.b_sn = 1;
      m.c = g[d].dw_hh; m.ldc = H;
      m.M = 4 * H; m.N = H; m.K = T * B;
      m.accumulate = acc; m.row_deinterleave_H = H; m.a_bf16 = bf16 ? 1 : 0;
      if ((e = run_gemm(m, ws, ws_left, flags, stream))) return e;
    }
    if (dx && do_rec) {
      GemmArgs m = {};
      m.a = dpre; m.a_sm = 4 * H; m.a_sk = 1;
      m.b = w_pack + (size_t)d * 4 * H * I; m.b_sk = I; m.b_sn = 1;
      m.b_hi = m.b + (size_t)D * 4 * H * I; m.b_lo = m.b_hi + (size_t)D * 4 * H * I;
      m.c = dx; m.ldc = I;
      m.M = T * B; m.N = I; m.K = 4 * H;
      m.accumulate = d > 0 ? 1 : 0; m.a_bf16 = bf16 ? 1 : 0;
      if ((e = run_gemm(m, ws, ws_left, flags, stream))) return e;
    }
  }
  return 0;
}

extern "C" int mrg_gemm_nt(const float* a, const float* b, const float* bias, float* c, int M, int N, int K,
                           void* workspace, size_t workspace_bytes, int flags, void* stream) {
  MRG_REQUIRE(a && b && c && M >= 0 && N >= 0 && K >= 0, "mrg_gemm_nt: bad arguments");
  MRG_REQUIRE(check_device() == 1, "mrg_gemm_nt: this library only runs on compute capability 10.x (B200)");
  GemmArgs g = {};
  g.a = a; g.a_sm = K; g.a_sk = 1;
  g.b = b; g.b_sk = 1; g.b_sn = K;
  g.bias = bias; g.c = c; g.ldc = N;
  g.M = M; g.N = N; g.K = K;
  return run_gemm(g, workspace, workspace_bytes, flags, (cudaStream_t)stream);
}

extern "C" int mrg_gemm_strided(const float* a, long long a_sm, long long a_sk, const float* b, long long b_sk,
                                long long b_sn, const float* bias, float* c, long long ldc, int M, int N, int K,
                                int accumulate, int deint_H, void* workspace, size_t workspace_bytes, int flags,
                                void* stream) {
  MRG_REQUIRE(a && b && c && M >= 0 && N >= 0 && K >= 0, "mrg_gemm_strided: bad arguments");
  MRG_REQUIRE(check_device() == 1, "mrg_gemm_strided: this library only runs on compute capability 10.x (B200)");
  GemmArgs g = {};
  g.a = a; g.a_sm = a_sm; g.a_sk = a_sk;
  g.b = b; g.b_sk = b_sk; g.b_sn = b_sn;
  g.bias = bias; g.c = c; g.ldc = ldc;
  g.M = M; g.N = N; g.K = K;
  g.accumulate = accumulate; g.row_deinterleave_H = deint_H;
  return run_gemm(g, workspace, workspace_bytes, flags, (cudaStream_t)stream);
}

extern "C" int mrg_gemm_strided_split(const float* a, long long a_sm, long long a_sk, const float* b_hi,
                                      const float* b_lo, long long b_sk, long long b_sn, const float* bias, float* c,
                                      long long ldc, int M, int N, int K, int accumulate, void* workspace,
                                      size_t workspace_bytes, int flags, void* stream) {
  MRG_REQUIRE(a && b_hi && b_lo && c && M >= 0 && N >= 0 && K >= 0, "mrg_gemm_strided_split: bad arguments");
  MRG_REQUIRE(check_device() == 1, "mrg_gemm_strided_split: this library only runs on compute capability 10.x (B200)");
  GemmArgs g = {};
  g.a = a; g.a_sm = a_sm; g.a_sk = a_sk;
  g.b_hi = b_hi; g.b_lo = b_lo; g.b_sk = b_sk; g.b_sn = b_sn;
  g.b = b_hi;   // operand checks of the tensor-core path look at b
  g.bias = bias; g.c = c; g.ldc = ldc;
  g.M = M; g.N = N; g.K = K;
  g.accumulate = accumulate;
  g.single_pass = (flags & (MRG_F_TF32 | MRG_F_BF16)) ? 1 : 0;
  MRG_REQUIRE(!(flags & MRG_F_SIMT_GEMM) && g.K > 0 && gemm_tc4_supported(g),
              "mrg_gemm_strided_split: shape / alignment not covered by the pre-split kernel (see mrg_gemm_split_supported)");
  (void)workspace; (void)workspace_bytes;
  return gemm_tc4(g, (cudaStream_t)stream);
}

extern "C" int mrg_gemm_split_supported(int M, int N, int K, long long a_sm, long long a_sk, long long b_sk,
                                        long long b_sn, long long ldc) {
  // alignment of the pointers is the caller's (16 bytes); this checks the shape and stride rules
  if (M < 4096 || K <= 0 || N <= 128 || N % 4 != 0 || ldc % 4 != 0) return 0;
  auto ok = [](long long s_r, long long s_k) {
    if (s_k == 1) return s_r >= 4 && s_r % 4 == 0;
    if (s_r == 1) return s_k >= 4 && s_k % 4 == 0;
    return false;
  };
  return ok(a_sm, a_sk) && ok(b_sn, b_sk) ? 1 : 0;
}

extern "C" size_t mrg_lstm_pack_floats(int I, int H, int D) { return (size_t)3 * D * 4 * H * I; }

extern "C" size_t mrg_gemm_workspace_bytes(int M, int N, int K) { return gemm_ws(M, N, K) + 256; }
