// Log-mel front end (K9 of SURVEY.md §2.2): the reference computes the STFT as a dense windowed-DFT conv1d
// on the host (code/audioldm/audio/stft.py:52-81), then |.|, the mel matmul and log(clamp(.,1e-5))
// (stft.py:159-180, audio_processing.py:85-91) and transposes to [frames, mels] (tools.py:78-79).
// Here one CTA owns one frame: the 1024 windowed samples (reflect padding resolved at load time) and a
// 1024-entry twiddle table live in shared memory, each thread evaluates DFT bins directly (the index
// (k*n) mod n_fft walks the table), magnitudes stay in shared memory for the mel projection.  The kernel is
// HBM-trivial (0.9 MB in+out at 10 s) and compute-light (1.1 GMAC); it exists to keep the whole path on device.
#include "common.cuh"

namespace aedit {
namespace {

__global__ void __launch_bounds__(256) stft_mel_kernel(const float* __restrict__ wav, int n_samples, int n_fft, int hop,
                                                       const float* __restrict__ window,
                                                       const float* __restrict__ mel_basis, int n_mels, int n_frames,
                                                       float* __restrict__ mag_out, float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float sm[];
  float* frame = sm;              // [n_fft]
  float* ctab = frame + n_fft;    // [n_fft]
  float* stab = ctab + n_fft;     // [n_fft]
  float* mag = stab + n_fft;      // [n_fft/2+1]
  const int f = blockIdx.x;
  const int nb = n_fft / 2 + 1;
  const int pad = n_fft / 2;
  for (int n = threadIdx.x; n < n_fft; n += blockDim.x) {
    int idx = f * hop + n - pad;  // position in the un-padded signal
    if (idx < 0) idx = -idx;                                   // reflect (no edge repeat), stft.py:61-65
    if (idx >= n_samples) idx = 2 * (n_samples - 1) - idx;
    idx = max(0, min(n_samples - 1, idx));
    frame[n] = wav[idx] * window[n];
    const float ang = 2.0f * (float)n / (float)n_fft;          // in units of pi
    ctab[n] = cospif(ang);
    stab[n] = sinpif(ang);
  }
  __syncthreads();
  const int mask = n_fft - 1;  // n_fft is a power of two
  for (int k = threadIdx.x; k < nb; k += blockDim.x) {
    float re = 0.f, im = 0.f;
    int idx = 0;
    for (int n = 0; n < n_fft; ++n) {
      const float x = frame[n];
      re = fmaf(x, ctab[idx], re);
      im = fmaf(x, stab[idx], im);
      idx = (idx + k) & mask;
    }
    const float m = sqrtf(re * re + im * im);
    mag[k] = m;
    if (mag_out) mag_out[(long long)f * nb + k] = m;
  }
  __syncthreads();
  for (int j = threadIdx.x; j < n_mels; j += blockDim.x) {
    const float* mb = mel_basis + (long long)j * nb;
    float acc = 0.f;
    for (int k = 0; k < nb; ++k) acc = fmaf(__ldg(mb + k), mag[k], acc);
    out[(long long)f * n_mels + j] = logf(fmaxf(acc, 1e-5f));
  }
}

}  // namespace
}  // namespace aedit

using namespace aedit;

extern "C" int ae_stft_mel(const float* wav, int n_samples, int n_fft, int hop, const float* window,
                           const float* mel_basis, int n_mels, int n_frames, float* mag_workspace, float* out_logmel,
                           ae_stream stream) {
  AE_CHECK_ARG(wav && window && mel_basis && out_logmel, "ae_stft_mel: null pointer");
  AE_CHECK_ARG(n_fft >= 64 && (n_fft & (n_fft - 1)) == 0 && n_fft <= 4096, "ae_stft_mel: n_fft must be a power of two");
  AE_CHECK_ARG(n_samples > n_fft / 2 && hop > 0 && n_mels > 0 && n_frames > 0, "ae_stft_mel: bad sizes");
  const size_t smem = (size_t)(3 * n_fft + n_fft / 2 + 1) * sizeof(float);
  static bool attr_set = false;
  if (!attr_set && smem > 48 * 1024) {
    cudaFuncSetAttribute(stft_mel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    attr_set = true;
  }
  launch_kernel(stft_mel_kernel, dim3(n_frames), dim3(256), (size_t)(smem), as_stream(stream), wav, n_samples, n_fft, hop, window, mel_basis, n_mels,
                                                            n_frames, mag_workspace, out_logmel);
  return launched("ae_stft_mel");
}
