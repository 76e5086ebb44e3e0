// On-GPU audio feature front-end — replaces AudioPreprocessor.__call__,
// mr_gen/utils/preprocess/audio.py:24-67 of the reference, after the waveform has been read:
//     MelSpectrogram(n_fft, hop, n_mels, center=False, power 2) -> log(max(., 1e-6))      audio.py:14-22,30-31
//     log-power of the raw (un-windowed) frame, max(., 1e-10)                              audio.py:41-53 (a Python loop)
//     [mel | log-power] -> delta / delta-delta by first differences, leading frames cut   audio.py:55-67
// The windowed real DFT is ONE tensor-core GEMM of this library over a strided view of the waveform (frame g = samples
// g*hop .. g*hop + nfft - 1: row stride hop, the frames are never materialised) against a [nfft x 2*bins] cos | -sin basis
// with the Hann window folded in; the kernels here turn the spectrum into features.
#include "mrg_common.cuh"

namespace mrg {

// static features of FPB frames per block: power spectrum -> mel filterbank -> log, plus the raw-frame log-power
constexpr int FE_FPB = 8;
constexpr int FE_THREADS = 256;

__global__ void __launch_bounds__(FE_THREADS)
fe_static_kernel(const float* __restrict__ spec, int ld_spec, const float* __restrict__ wave, const float* __restrict__ fb,
                 float* __restrict__ feat, long long frames_total, int frames_per_seq, long long frame_stride,
                 long long samples_per_seq, int hop, int nfft, int bins, int nmels) {
  extern __shared__ float fe_sm[];
  float* pw = fe_sm;                       // [FE_FPB][bins] power spectrum
  float* red = pw + FE_FPB * bins;         // [FE_FPB] raw-frame energy
  const long long f0 = (long long)blockIdx.x * FE_FPB;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // frame index f (over the whole batch) -> spectrum row g = b * frame_stride + j, first sample b * samples_per_seq + j*hop
  for (int idx = threadIdx.x; idx < FE_FPB * bins; idx += FE_THREADS) {
    const int r = idx / bins, k = idx % bins;
    const long long f = f0 + r;
    float v = 0.f;
    if (f < frames_total) {
      const long long b = f / frames_per_seq, j = f % frames_per_seq;
      const float* row = spec + (size_t)(b * frame_stride + j) * ld_spec;
      const float re = row[k], im = row[bins + k];
      v = re * re + im * im;
    }
    pw[idx] = v;
  }
  if (warp < FE_FPB) {   // one warp per frame: sum of squares of the raw samples
    const long long f = f0 + warp;
    float s = 0.f;
    if (f < frames_total) {
      const long long b = f / frames_per_seq, j = f % frames_per_seq;
      const float* x = wave + (size_t)(b * samples_per_seq + j * hop);
      for (int i = lane; i < nfft; i += 32) s = fmaf(x[i], x[i], s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) red[warp] = s;
  }
  __syncthreads();
  const int nf = nmels + 1;
  for (int idx = threadIdx.x; idx < FE_FPB * nf; idx += FE_THREADS) {
    const int r = idx / nf, m = idx % nf;
    const long long f = f0 + r;
    if (f >= frames_total) continue;
    float out;
    if (m < nmels) {
      float acc = 0.f;
      const float* p = pw + r * bins;
      for (int k = 0; k < bins; ++k) acc = fmaf(p[k], fb[(size_t)k * nmels + m], acc);
      out = logf(fmaxf(acc, 1e-6f));     // log(clamp(clamp(x, 1e-10), 1e-6))
    } else {
      out = logf(fmaxf(red[r], 1e-10f));
    }
    feat[(size_t)f * nf + m] = out;
  }
}

// out[b][j] = [x[j+o] | x[j+o] - x[j+o-1] | (x[j+2] - x[j+1]) - (x[j+1] - x[j])]   (o = delta order, audio.py:55-67)
__global__ void fe_delta_kernel(const float* __restrict__ feat, float* __restrict__ out, int B, int frames_per_seq,
                                int nf, int order) {
  const int out_frames = frames_per_seq - order;
  const long long total = (long long)B * out_frames * nf;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int m = (int)(i % nf);
    const long long bj = i / nf;
    const int j = (int)(bj % out_frames);
    const long long b = bj / out_frames;
    const float* x = feat + ((size_t)b * frames_per_seq + j) * nf + m;
    float* o = out + (size_t)bj * nf * (order + 1);
    const float x0 = x[0];
    if (order == 0) { o[m] = x0; continue; }
    const float x1 = x[nf];
    if (order == 1) { o[m] = x1; o[nf + m] = x1 - x0; continue; }
    const float x2 = x[2 * nf];
    const float d1a = x1 - x0, d1b = x2 - x1;
    o[m] = x2; o[nf + m] = d1b; o[2 * nf + m] = d1b - d1a;
  }
}

}  // namespace mrg

using namespace mrg;

extern "C" int mrg_audio_features(const float* spec, int ld_spec, const float* wave, const float* mel_fb, float* feat,
                                  float* out, int B, int frames_per_seq, long long frame_stride,
                                  long long samples_per_seq, int hop, int nfft, int nmels, int delta_order,
                                  void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MRG_REQUIRE(spec && wave && mel_fb && feat && out, "mrg_audio_features: null pointer");
  MRG_REQUIRE(B > 0 && frames_per_seq > delta_order && nfft > 0 && hop > 0 && nmels > 0 && delta_order >= 0 &&
                  delta_order <= 2, "mrg_audio_features: bad shape (delta_order must be 0, 1 or 2)");
  const int bins = nfft / 2 + 1;
  MRG_REQUIRE(ld_spec >= 2 * bins, "mrg_audio_features: spectrum rows hold re | im of %d bins", bins);
  const long long frames_total = (long long)B * frames_per_seq;
  const size_t smem = (size_t)(FE_FPB * bins + FE_FPB) * sizeof(float);
  MRG_REQUIRE(smem <= 48 * 1024, "mrg_audio_features: nfft too large for the feature kernel");
  const unsigned blocks = (unsigned)((frames_total + FE_FPB - 1) / FE_FPB);
  fe_static_kernel<<<blocks, FE_THREADS, smem, stream>>>(spec, ld_spec, wave, mel_fb, feat, frames_total, frames_per_seq,
                                                         frame_stride, samples_per_seq, hop, nfft, bins, nmels);
  MRG_CUDA_CHECK(cudaGetLastError());
  const long long total = (long long)B * (frames_per_seq - delta_order) * (nmels + 1);
  long long db = (total + 255) / 256;
  if (db > 148 * 8) db = 148 * 8;
  fe_delta_kernel<<<(unsigned)db, 256, 0, stream>>>(feat, out, B, frames_per_seq, nmels + 1, delta_order);
  MRG_CUDA_CHECK(cudaGetLastError());
  count_launch(2);
  return 0;
}
