/*
 * mrg_lstm.h — C-ABI of the B200 (sm_100a) LSTM hot path of MultimodalReactionGeneration.
 *
 * The reference has no FFI layer: every LSTM call is `torch.nn.LSTM(...)(x, hx)` held by three
 * thin wrappers.  These entry points are what a binding for that seam calls instead:
 *
 *   mrg_lstm_layer_forward / _backward   replace  nn.LSTM.forward and its autograd backward at
 *       mr_gen/model/utils/lstm_block.py:21,41    (LSTMModule.lstm_module)
 *       mr_gen/model/utils/lstm_sampler.py:16,29  (LSTMSampler.sampler)
 *       mr_gen/model/utils/mixer_block.py:237,251 (LSTMMixer.mixer)
 *   mrg_rollout_forward / _backward      replace the Python time loop of
 *       mr_gen/model/lstm_with_sampling/lstm_with_sample.py:379-408 (head_motion_generation)
 *       + :410-433 (generate_one_step) with one persistent cluster kernel per direction (csrc/mrg_rollout.cu)
 *   mrg_philox_mask                      replaces `torch.rand(length) < epoch/max_epochs`
 *       mr_gen/model/lstm_with_sampling/lstm_with_sample.py:389
 *
 * Conventions: plain pointers and sizes only; every pointer is a DEVICE pointer to fp32 data owned by
 * the caller (a torch tensor) unless stated otherwise, and must stay alive until the work queued on
 * `stream` (a cudaStream_t passed as void*) has finished.  All calls are asynchronous on `stream`.
 * Return value: 0 = OK, <0 = invalid argument (MRG_E_*), >0 = cudaError_t.  No exceptions cross the
 * boundary; mrg_last_error_string() describes the last failure on the calling thread.
 *
 * Tensor layouts (all contiguous, fp32):
 *   x       [T][B][I]            time-major input of the layer
 *   w_ih    [4H][I], w_hh [4H][H], b_* [4H]   torch.nn.LSTM layout, gate order i,f,g,o (rnn.py:842-847)
 *   gates   [D][T][B][H][4]      reserve: post-activation (i,f,g,o) per hidden unit, gate-interleaved;
 *                                the backward overwrites it with d(pre-activation); fp32, or bfloat16 with MRG_F_BF16
 *   y_ext   [D][T+1][B][H]       hidden states with one extra slot for h0:
 *                                  direction 0 (forward in time): slot 0 = h0, slot t+1 = h_t
 *                                  direction 1 (reverse):          slot T = h0, slot t   = h_t
 *   c_ext   [D][T+1][B][H]       cell states, same slot convention
 */
#ifndef MRG_LSTM_H_
#define MRG_LSTM_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MRG_VERSION 100

#define MRG_E_INVALID   (-1)  /* bad shape / null pointer */
#define MRG_E_UNSUPPORTED (-2)
#define MRG_E_WORKSPACE (-3)  /* workspace too small */

/* flags */
#define MRG_F_TRAIN        1   /* forward: keep gates / c for the backward                     */
#define MRG_F_GENERIC_REC  2   /* force the generic (non-cluster) recurrent kernels             */
#define MRG_F_SIMT_GEMM    4   /* force the SIMT fp32 GEMM instead of the tcgen05 3xTF32 GEMM   */
#define MRG_F_ACCUMULATE   8   /* backward: add into dw_ih / dw_hh / db instead of overwriting  */
#define MRG_F_TF32        32   /* reduced-precision mode: the projection GEMMs run ONE tf32 tensor-core
                                  pass (10-bit mantissa, >= bf16 precision) instead of the 3-pass
                                  fp32-grade split.  H = 256 layers (<= 48 batch rows per cluster)
                                  also run h W_hh^T / dpre W_hh as one tf32 pass on the warp-level tensor
                                  cores (csrc/mrg_rec_fwd3.cu, mrg_rec_bwd3.cu); states, cell and
                                  accumulation stay fp32.  Bound: 2e-2 per step / 5e-2 on gradients   */
#define MRG_F_BF16        64   /* bf16 mode (with MRG_F_TF32): the reserve `gates` holds 4 x bfloat16 per hidden unit
                                  (8 bytes: x-projection in, gates out, d(pre-activations) back) — half the bytes of
                                  the largest buffer of the path; the GEMMs around it take / write bfloat16 and run one
                                  tf32 tensor-core pass (bf16 values are exact in tf32), accumulation and the
                                  recurrence stay fp32.  Cluster kernels only: H in {128, 256}, T > 1               */
#define MRG_F_GRU        128   /* the layer is a GRU run on the LSTM machinery (nn.GRU at mixer_block.py:194): the caller passes
                                  the weights in four-gate form — w_ih rows (r, z, n, 0), w_hh rows (r, z, 0, n), b_ih =
                                  (b_ir + b_hr, b_iz + b_hz, b_in, b_hn), b_hh = NULL — the recurrent kernels apply the
                                  GRU cell, the reserve holds (r, z, n, W_hn h + b_hn) and comes back as (d r_pre, d z_pre,
                                  d n_pre, d n_pre r); c_ext / c0 / dc_n are unused.  H in {128, 256}, T > 1            */
#define MRG_F_CLUSTER_BUDGET(n) (((n) & 0xFF) << 16)  /* recurrent kernels use at most n clusters (0 = all): lets two
                                  independent LSTM stacks (audio / motion encoders) run side by side on two streams */
#define MRG_F_ACC_WEIGHTS 256  /* backward: add into dw_ih / dw_hh only (db is overwritten): weight gradients that
                                  accumulate straight into the trainer's flat gradient bucket            */
#define MRG_F_PACK_VALID  512  /* forward: `w_pack` and the first two regions of `workspace` (bias pack, and the W_hh pack
                                  of the T == 1 carried-state path) still hold what an earlier call with the SAME weights,
                                  shape, `w_pack` and `workspace` wrote — the pack launch is skipped (frame-by-frame
                                  inference with frozen weights; the caller owns both buffers and the invalidation)    */
#define MRG_F_BWD_NO_WGRAD   2048  /* backward: BPTT + bias sums + dX only (leaves d(pre-activations) in `gates`)    */
#define MRG_F_BWD_WGRAD_ONLY 4096  /* backward: only dW_ih / dW_hh from a previous NO_WGRAD call (any stream)        */
#define MRG_F_ZERO_STATE  16   /* caller guarantees h0 = c0 = 0 (hx=None): with T == 1 the layer is a
                                  pointwise cell on the projection (no recurrence, W_hh inert) — the
                                  stateless predictor steps of the lstm_with_sampling rollout       */

typedef struct mrg_lstm_dir_weights {
  const float* w_ih;  /* [4H][I] */
  const float* w_hh;  /* [4H][H] */
  const float* b_ih;  /* [4H] or NULL */
  const float* b_hh;  /* [4H] or NULL */
  const float* h0;    /* [B][H] or NULL (zeros) */
  const float* c0;    /* [B][H] or NULL (zeros) */
} mrg_lstm_dir_weights;

typedef struct mrg_lstm_dir_grads {
  float* dw_ih;  /* [4H][I] */
  float* dw_hh;  /* [4H][H] */
  float* db;     /* [4H]  gradient of b_ih and of b_hh (they enter as a sum) */
  float* dh0;    /* [B][H] or NULL */
  float* dc0;    /* [B][H] or NULL */
} mrg_lstm_dir_grads;

int mrg_version(void);
const char* mrg_last_error_string(void);

/* Device facts the host side sizes its launches with. Returns 0 and fills the outputs. */
int mrg_device_info(int* sm_count, int* max_clusters_h256, int* max_clusters_h128, int* cc_major,
                    int* cc_minor);

/* Bytes of scratch mrg_lstm_layer_forward / _backward need for this shape (max of both). */
size_t mrg_lstm_workspace_bytes(int T, int B, int I, int H, int D);

/* One nn.LSTM layer, D = 1 (uni) or 2 (bidirectional) directions in one launch.
 * w_pack: caller-owned buffer of mrg_lstm_pack_floats(I, H, D) = 3*D*4H*I floats; receives W_ih with gate-interleaved
 * rows in three planes — as is, its tf32 hi part, its lo part — (the B operand of the projection and dX GEMMs, split
 * once per forward instead of once per tile) and is consumed again by the backward. */
size_t mrg_lstm_pack_floats(int I, int H, int D);
int mrg_lstm_layer_forward(const float* x, const mrg_lstm_dir_weights* w, float* w_pack, float* gates,
                           float* y_ext, float* c_ext, void* workspace, size_t workspace_bytes, int T,
                           int B, int I, int H, int D, int flags, void* stream);

/* BPTT of the same layer.  dy is [T][B][D*H] (NULL = zeros); dh_n/dc_n are [D][B][H] or NULL.
 * dx [T][B][I] may be NULL when the input needs no gradient. */
int mrg_lstm_layer_backward(const float* x, const mrg_lstm_dir_weights* w, const float* w_pack,
                            const float* dy, const float* dh_n, const float* dc_n, float* gates,
                            const float* y_ext, const float* c_ext, float* dx,
                            const mrg_lstm_dir_grads* g, void* workspace, size_t workspace_bytes, int T,
                            int B, int I, int H, int D, int flags, void* stream);

/* C[M][N] = A[M][K] * B[N][K]^T (+ bias[N]) — the time-parallel projection GEMM, exported for tests
 * and for the Linear layers adjacent to the LSTMs. */
int mrg_gemm_nt(const float* a, const float* b, const float* bias, float* c, int M, int N, int K,
                void* workspace, size_t workspace_bytes, int flags, void* stream);

/* General strided form used by the backward (and by the tests): A(m,k) = a[m*a_sm + k*a_sk],
 * B(k,n) = b[k*b_sk + n*b_sn], C[row(m)][n] (+)= sum_k A*B (+ bias[n]); row(m) = (m%4)*deint_H + m/4 when
 * deint_H > 0 (gate de-interleave of weight gradients), else m. */
int mrg_gemm_strided(const float* a, long long a_sm, long long a_sk, const float* b, long long b_sk,
                     long long b_sn, const float* bias, float* c, long long ldc, int M, int N, int K,
                     int accumulate, int deint_H, void* workspace, size_t workspace_bytes, int flags,
                     void* stream);
size_t mrg_gemm_workspace_bytes(int M, int N, int K);

/* dst[n0][n1][H] (contiguous) = src[i0 * s0 + i1 * s1 + h]: the batch-first <-> time-major relayout of a [B, T, H] activation
 * in front of / behind a layer (B200LSTM takes time-major input; the reference's tensors are batch_first).  H % 4 == 0,
 * strides multiples of 4 floats, 16-byte aligned pointers. */
int mrg_copy_rows(const float* src, long long s0, long long s1, float* dst, int n0, int n1, int H, void* stream);

/* Weight operand split once: hi[i] = w[i] rounded to tf32 (nearest, ties away), lo[i] = w[i] - hi[i] (exact).
 * mrg_gemm_strided_split is mrg_gemm_strided (no de-interleave) with B given as those two planes: it runs the persistent
 * 128 x 256 tcgen05 kernel (csrc/mrg_gemm_tc4.cu), which needs no B conversion pass.  Covered shapes:
 * mrg_gemm_split_supported() != 0 (N > 128, strides as for mrg_gemm_strided, 16-byte aligned pointers); it fails loudly
 * otherwise — the caller falls back to mrg_gemm_strided with the unsplit weight. */
int mrg_split_tf32(const float* w, float* hi, float* lo, size_t n, void* stream);
int mrg_gemm_split_supported(int M, int N, int K, long long a_sm, long long a_sk, long long b_sk, long long b_sn,
                             long long ldc);
int mrg_gemm_strided_split(const float* a, long long a_sm, long long a_sk, const float* b_hi, const float* b_lo,
                           long long b_sk, long long b_sn, const float* bias, float* c, long long ldc, int M, int N,
                           int K, int accumulate, void* workspace, size_t workspace_bytes, int flags, void* stream);

/* Fused residual + LayerNorm — ResidualConnection.forward, mr_gen/model/utils/residual_connection.py:29-32:
 * out = LN(y + x) * gamma + beta over the last dimension (H in {128, 256, 512}).  Rows are indexed (i0, i1),
 * i0 < n0, i1 < n1, with element strides (s0, s1) per tensor, so a time-major LSTM output and a batch-first
 * block input are read in place.  x may be NULL (plain LayerNorm).  mean / rstd: [n0*n1], kept for the
 * backward, which recomputes y + x and returns dsum = d(y) = d(x) plus d(gamma), d(beta). */
size_t mrg_layernorm_workspace_bytes(int H);
int mrg_residual_layernorm_forward(const float* y, long long y_s0, long long y_s1, const float* x,
                                   long long x_s0, long long x_s1, const float* gamma, const float* beta,
                                   float* out, long long o_s0, long long o_s1, float* mean, float* rstd,
                                   int n0, int n1, int H, float eps, void* stream);
int mrg_residual_layernorm_backward(const float* dout, long long d_s0, long long d_s1, const float* y,
                                    long long y_s0, long long y_s1, const float* x, long long x_s0,
                                    long long x_s1, const float* gamma, const float* mean, const float* rstd,
                                    float* dsum, long long g_s0, long long g_s1, float* dgamma, float* dbeta,
                                    void* workspace, size_t workspace_bytes, int n0, int n1, int H,
                                    void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Persistent on-device rollout of lstm_with_sampling — replaces the Python time loop
 * mr_gen/model/lstm_with_sampling/lstm_with_sample.py:379-408 (head_motion_generation) + :410-433
 * (generate_one_step) for everything that depends on the previous step (scheduled sampling, step-wise teacher
 * forcing, free-running generation, streaming with T = 1).  Per step t and batch row b:
 *     prev   = t == 0 ? gt_prev[0] : (mask[t-1] ? pred[t-1] : gt_prev[t])        (quirks Q5, Q6)
 *     x_0    = base[t] + W_prev prev
 *     x_l+1  = LayerNorm_l(cell_l(x_l) + x_l)   cell_l: one nn.LSTM step from zero state (quirk Q2): W_hh and the
 *                                               forget gate are inert, c = sig(i) tanh(g), h = sig(o) tanh(c)
 *     pred[t] = W_2 act(W_1 x_L + b_1) + b_2    act = ReLU when relu != 0
 * base = feature_projection over [sampler output | partner pose] + bias (time-parallel GEMM, the caller's);
 * W_prev = the last P columns of feature_projection.weight (row stride w_prev_ld).  All tensors time-major.
 * H in {32, 64, 128, 256}, L in {1, 2}, P <= 32, FB <= 64 and a multiple of H/8: mrg_rollout_supported().
 * mask: uint8 [T][B] or NULL (never feed back); produce it with mrg_philox_mask for the counter-based RNG. */
typedef struct mrg_rollout_weights {
  int H, L, P, FB, relu;
  float ln_eps;
  const float* w_prev;        /* [H][P], row stride w_prev_ld */
  long long w_prev_ld;
  const float* w_ih[2];       /* [4H][H]  torch gate order i, f, g, o */
  const float* b_ih[2];       /* [4H] or NULL */
  const float* b_hh[2];       /* [4H] or NULL */
  const float* ln_g[2];       /* [H] */
  const float* ln_b[2];       /* [H] */
  const float* w1;            /* [FB][H] */
  const float* b1;            /* [FB] or NULL */
  const float* w2;            /* [P][FB] */
  const float* b2;            /* [P] or NULL */
} mrg_rollout_weights;

/* Written by the forward when non-NULL (training), read by the backward. */
typedef struct mrg_rollout_reserve {
  float* xs;     /* [L+1][T][B][H]  x_0 .. x_L */
  float* gates;  /* [L][T][B][3][H] post-activation i, g, o */
  float* xhat;   /* [L][T][B][H]    normalised pre-affine LayerNorm values */
  float* rstd;   /* [L][T][B] */
  float* fact;   /* [T][B][FB]      FFN hidden after the activation */
  float* prev;   /* [T][B][P]       the pose that was fed to every step */
} mrg_rollout_reserve;

/* Outputs of the backward-through-time kernel.  Weight gradients are time-parallel GEMMs / column sums over these:
 * dW_2 = dy^T fact, db_2 = colsum dy; dW_1 = df^T xs[L], db_1 = colsum df; dW_ih^l = dpre[l]^T xs[l],
 * db_ih^l = db_hh^l = colsum dpre[l]; d(ln weight) = colsum dln_g[l], d(ln bias) = colsum dln_b[l];
 * dW_prev = dbase^T prev; d(base) = dbase; d(gt_prev[t]) = dprev[t] where step t was NOT fed back. */
typedef struct mrg_rollout_grads {
  float* dy;     /* [T][B][P]   total gradient at pred (loss + fed-back) */
  float* df;     /* [T][B][FB]  gradient at the FFN hidden pre-activation */
  float* dpre;   /* [L][T][B][4H] gradient at the gate pre-activations, torch gate order, forget columns zero */
  float* dbase;  /* [T][B][H] */
  float* dprev;  /* [T][B][P] */
  float* dln_g;  /* [L][B][H]   per-row sums over t */
  float* dln_b;  /* [L][B][H] */
} mrg_rollout_grads;

int mrg_rollout_supported(int H, int L, int P, int FB);
int mrg_rollout_forward(const float* base, const float* gt_prev, const uint8_t* mask, const mrg_rollout_weights* w,
                        float* pred, const mrg_rollout_reserve* reserve, int T, int B, void* stream);
int mrg_rollout_backward(const float* dpred, const uint8_t* mask, const mrg_rollout_weights* w,
                         const mrg_rollout_reserve* reserve, const mrg_rollout_grads* g, int T, int B, void* stream);

/* Scheduled-sampling mask: out[t*B+b] = philox4x32_10(ctr=(lo(offset+t), hi(offset+t), shared?0:b, 0),
 * key=(lo(seed), hi(seed)))[0] >> 8 as a 24-bit uniform < prob.  out is a DEVICE uint8 buffer. */
int mrg_philox_mask(uint64_t seed, uint64_t offset, float prob, int T, int B, int shared, uint8_t* out,
                    void* stream);

/* out[N] (+)= column sums of x[M][N] (any N; float4 loads when N % 4 == 0 and x is 16-byte aligned): the bias gradient of the Linear layers next
 * to the LSTMs (LSTMModule.mixer, the bottleneck FFN, embed / projection layers).  Deterministic two-pass sum. */
size_t mrg_colsum_workspace_bytes(int M, int N);
int mrg_colsum(const float* x, float* out, int M, int N, int accumulate, void* workspace, size_t workspace_bytes,
               void* stream);

/* Optimizer step of the training loop (mr_gen/model/simple_lstm/simple_lstm.py:193-221 builds torch.optim.AdamW):
 * decoupled-weight-decay Adam over FLAT fp32 buckets p / g / m / v of n floats (n % 4 == 0), torch.optim.AdamW
 * arithmetic.  state[3] (device) = {step count, 1/(1-b1^step), 1/sqrt(1-b2^step)}, advanced by the call itself so a
 * captured CUDA graph keeps counting.  lr is read from lr_dev when non-NULL.  g is scaled by grad_scale first
 * (1/world_size after the all-reduce) and cleared afterwards when zero_grad != 0. */
int mrg_adamw_flat(float* p, float* g, float* m, float* v, size_t n, const float* lr_dev, float lr, float beta1,
                   float beta2, float eps, float weight_decay, float* state, float grad_scale, int zero_grad,
                   void* stream);

/* Fused fp32 multi-head attention between the LSTM stacks.  Replaces the
 * nn.MultiheadAttention core called at mr_gen/model/utils/multi_modal_att.py:12-31 (no mask) and at
 * mr_gen/model/utils/for_sequential.py:25-50 with the mask of mr_gen/model/utils/multi_modal_metaformer.py:32-79.
 * q [B,Tq,*], k / v [B,Tk,*], o [B,Tq,*]: head h occupies columns h*hd .. h*hd+hd-1 of a row, ld* = row stride in
 * floats (batch stride = T*ld; k and v may be the halves of one fused projection output).  hd in {32, 64}; pointers
 * 16-byte aligned, strides multiples of 4.  o = softmax(scale * q k^T + mask) v per (batch, head).
 * mask_mode 0: none | 1: key j visible to query i iff j / rate <= i | 2: iff j <= i / rate; additionally (i, j) is
 * masked when pad_q[b*Tq+i] and pad_k[b*Tk+j] are both non-zero (both NULL = no padding).  A query with no visible
 * key yields 0 (torch: NaN).  lse [B,nh,Tq] (log2 domain) is saved for the backward; dvec [B,nh,Tq] is scratch. */
/* Which kernels the two calls below run (process-wide): 0 = warp-level tensor cores with the 3xTF32 split (default,
 * fp32-grade: csrc/mrg_attention_mma.cu), 1 = one tf32 pass (the tf32 / bf16 precision modes), 2 = the CUDA-core fp32
 * kernels (csrc/mrg_attention.cu, cross-check). */
int mrg_attention_set_mode(int mode);
int mrg_attention_forward(const float* q, int ldq, const float* k, int ldk, const float* v, int ldv, float* o,
                          int ldo, float* lse, int B, int nh, int Tq, int Tk, int hd, float scale, int mask_mode,
                          int rate, const uint8_t* pad_q, const uint8_t* pad_k, void* stream);
int mrg_attention_backward(const float* q, int ldq, const float* k, int ldk, const float* v, int ldv, const float* o,
                           int ldo, const float* lse, const float* dout, int lddo, float* dq, int lddq, float* dk,
                           int lddk, float* dv, int lddv, float* dvec, int B, int nh, int Tq, int Tk, int hd,
                           float scale, int mask_mode, int rate, const uint8_t* pad_q, const uint8_t* pad_k,
                           void* stream);

/* GRU recurrence of lstmformer's GRU mixer (nn.GRU at mr_gen/model/utils/mixer_block.py:194).
 * gx [T][B][3H] = x W_ih^T + b_ih (gate order r, z, n; from the projection GEMM), w_hh [3H][H], w_hh_t [H][3H] = its
 * transpose or NULL (when given and H % 4 == 0 the forward streams columns: no cross-lane reduction), b_hh [3H] or NULL,
 * y_ext [T+1][B][H]: slot 0 = h0 on entry, slot t+1 = h_t on return; reserve [T][B][4][H] (r, z, n, W_hn h + b_hn),
 * written when train != 0.  Backward: dy [T][B][H] or NULL, dh_n [B][H] or NULL -> dgx / dgh [T][B][3H] (gradients of
 * the pre-activations seen from the input side (r, z, n) and from the hidden side (r, z, hn)), dh0 [B][H] or NULL;
 * dX = dgx W_ih, dW_ih = dgx^T X, dW_hh = dgh^T y_ext[0:T], db_ih / db_hh = column sums are the caller's GEMMs. */
int mrg_gru_forward(const float* gx, const float* w_hh, const float* w_hh_t, const float* b_hh, float* y_ext,
                    float* reserve, int T, int B, int H, int train, void* stream);
int mrg_gru_backward(const float* dy, const float* dh_n, const float* reserve, const float* y_ext, const float* w_hh,
                     float* dgx, float* dgh, float* dh0, int T, int B, int H, void* stream);

/* On-GPU audio feature front-end: what AudioPreprocessor.__call__ computes after reading the
 * waveform — mr_gen/utils/preprocess/audio.py:30-37 (log-mel + log-power, 27-d at the reference's 26 mels) and :55-67
 * (delta / delta-delta by first differences -> 81-d).  The windowed real DFT is a GEMM of this library (mrg_gemm_strided
 * over the waveform viewed as overlapping frames, row stride = hop) and is the caller's; this call turns its output into
 * features.  spec: [rows][ld_spec] with re of bin k at column k and im at column bins + k (bins = nfft/2 + 1); frame j of
 * sequence b is row b * frame_stride + j and starts at sample b * samples_per_seq + j * hop of `wave`.  mel_fb
 * [bins][nmels] (torchaudio.functional.melscale_fbanks layout).  feat [B][frames_per_seq][nmels + 1] is scratch (the
 * static features); out [B][frames_per_seq - delta_order][(nmels + 1) * (delta_order + 1)]. */
int mrg_audio_features(const float* spec, int ld_spec, const float* wave, const float* mel_fb, float* feat, float* out,
                       int B, int frames_per_seq, long long frame_stride, long long samples_per_seq, int hop, int nfft,
                       int nmels, int delta_order, void* stream);

/* Developer hook: device buffer of 16*1024*2 uint64 that -DMRG_REC_TRACE builds of the recurrent kernels fill
 * with (clock, event) records; NULL disables.  No effect in regular builds. */
int mrg_debug_set_trace(unsigned long long* buf);

/* Measurement hooks used by bench.py: number of kernels this library has launched since it was loaded,
 * and CUDA-event timing of the recurrent / GEMM / rollout launches on their own stream.  mrg_profile_read fills
 * ms[MRG_PROF_KINDS], n[MRG_PROF_KINDS] for {recurrent forward, recurrent backward, GEMM, rollout forward, rollout
 * backward} and resets the record; mrg_profile_kernel_name(kind) names the last kernel of that kind that was timed. */
#define MRG_PROF_KINDS 5
unsigned long long mrg_launch_count(void);
int mrg_profile_enable(int on);
int mrg_profile_read(float* ms, int* n);
const char* mrg_profile_kernel_name(int kind);

#ifdef __cplusplus
}
#endif
#endif /* MRG_LSTM_H_ */
