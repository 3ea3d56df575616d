// Log-mel front end (K9 of SURVEY.md §2.2): the reference computes the STFT as a dense windowed-DFT conv1d on the host
// (code/audioldm/audio/stft.py:52-81: 1026 basis rows x 1024 taps per frame, 1.1 GMAC per 10 s clip), then |.|, the mel
// matmul and log(clamp(., 1e-5)) (stft.py:159-180, audio_processing.py:85-91), transposed to [frames, mels]
// (tools.py:78-79).  Here the whole chain is ONE kernel and the DFT is a shared-memory FFT:
//   * a CTA owns kFramesPerCta = 4 consecutive frames; two real frames are packed into one complex radix-2 FFT of
//     n_fft points (frame A -> real part, frame B -> imaginary part; X_A[k] = (Z[k] + conj Z[N-k]) / 2,
//     X_B[k] = (Z[k] - conj Z[N-k]) / 2i), so 4 frames cost two n_fft-point FFTs: 50 kFLOP instead of 4.2 MFLOP;
//   * waveform reads are coalesced (consecutive threads read consecutive samples; reflect padding of stft.py:61-65 is
//     resolved in the index), the window is applied at load, the bit-reversed scatter goes to shared memory;
//   * the n_fft/2 twiddles are computed once per CTA (sincospif) into shared memory;
//   * the mel projection runs warp-per-mel-row with lanes striding the bins (coalesced 128-byte reads of the basis,
//     shared by the CTA's 4 frames), a shuffle tree per (row, frame), log-clamp fused; the magnitudes are written out
//     (coalesced) only if the caller asks for them.
// HBM traffic = the waveform (640 KiB) + outputs; the kernel is latency / L2 bound (the 131 KiB mel basis is re-read
// from L2 by every CTA), measured in profiles/r02_hbm_kernels_*.
#include "common.cuh"

namespace aedit {
namespace {

constexpr int kFramesPerCta = 4;
constexpr int kStftThreads = 256;

__global__ void __launch_bounds__(kStftThreads) stft_mel_kernel(const float* __restrict__ wav, int n_samples, int n_fft,
                                                                int log2n, int hop, const float* __restrict__ window,
                                                                const float* __restrict__ mel_basis, int n_mels,
                                                                int n_frames, float* __restrict__ mag_out,
                                                                float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float sm[];
  const int nb = n_fft / 2 + 1;
  float* re = sm;                      // [n_fft]
  float* im = re + n_fft;              // [n_fft]
  float* twc = im + n_fft;             // [n_fft/2]  cos(-2 pi j / n_fft)
  float* tws = twc + n_fft / 2;        // [n_fft/2]  sin(-2 pi j / n_fft)
  float* mag = tws + n_fft / 2;        // [kFramesPerCta][nb]
  const int f0 = blockIdx.x * kFramesPerCta;
  const int pad = n_fft / 2;
  const int tid = threadIdx.x;
  for (int j = tid; j < n_fft / 2; j += kStftThreads) {
    float s, c;
    sincospif(-2.0f * (float)j / (float)n_fft, &s, &c);
    twc[j] = c;
    tws[j] = s;
  }
  for (int pair = 0; pair < kFramesPerCta / 2; ++pair) {
    const int fa = f0 + 2 * pair, fb = fa + 1;
    __syncthreads();                   // previous pair's unpack has finished reading re / im
    for (int n = tid; n < n_fft; n += kStftThreads) {
      const float w = window[n];
      float xa = 0.f, xb = 0.f;
      if (fa < n_frames) {
        int idx = fa * hop + n - pad;                              // position in the un-padded signal
        if (idx < 0) idx = -idx;                                   // reflect (no edge repeat), stft.py:61-65
        if (idx >= n_samples) idx = 2 * (n_samples - 1) - idx;
        xa = wav[max(0, min(n_samples - 1, idx))] * w;
      }
      if (fb < n_frames) {
        int idx = fb * hop + n - pad;
        if (idx < 0) idx = -idx;
        if (idx >= n_samples) idx = 2 * (n_samples - 1) - idx;
        xb = wav[max(0, min(n_samples - 1, idx))] * w;
      }
      const int r = (int)(__brev((unsigned)n) >> (32 - log2n));
      re[r] = xa;
      im[r] = xb;
    }
    __syncthreads();
    for (int s = 1; s <= log2n; ++s) {                             // radix-2 decimation in time, in place
      const int half = 1 << (s - 1);
      const int tstride = n_fft >> s;
      for (int b = tid; b < n_fft / 2; b += kStftThreads) {
        const int j = b & (half - 1);
        const int i0 = ((b >> (s - 1)) << s) + j;
        const int i1 = i0 + half;
        const float c = twc[j * tstride], sn = tws[j * tstride];
        const float vr = re[i1] * c - im[i1] * sn;
        const float vi = re[i1] * sn + im[i1] * c;
        const float ur = re[i0], ui = im[i0];
        re[i0] = ur + vr;
        im[i0] = ui + vi;
        re[i1] = ur - vr;
        im[i1] = ui - vi;
      }
      __syncthreads();
    }
    for (int k = tid; k < nb; k += kStftThreads) {                 // split the two real spectra, magnitudes
      const int kn = (n_fft - k) & (n_fft - 1);
      const float zr = re[k], zi = im[k], yr = re[kn], yi = im[kn];
      const float ar = 0.5f * (zr + yr), ai = 0.5f * (zi - yi);
      const float br = 0.5f * (zi + yi), bi = -0.5f * (zr - yr);
      const float ma = sqrtf(ar * ar + ai * ai);                   // stft.py:76
      const float mb = sqrtf(br * br + bi * bi);
      mag[(2 * pair) * nb + k] = ma;
      mag[(2 * pair + 1) * nb + k] = mb;
      if (mag_out) {
        if (fa < n_frames) mag_out[(long long)fa * nb + k] = ma;
        if (fb < n_frames) mag_out[(long long)fb * nb + k] = mb;
      }
    }
  }
  __syncthreads();
  // mel projection (stft.py:176-177) + log(clamp(., 1e-5)) (audio_processing.py:85-91): warp per mel row
  const int lane = tid & 31, warp = tid >> 5;
  for (int j = warp; j < n_mels; j += kStftThreads / 32) {
    const float* mb = mel_basis + (long long)j * nb;
    float acc[kFramesPerCta];
#pragma unroll
    for (int q = 0; q < kFramesPerCta; ++q) acc[q] = 0.f;
    for (int k = lane; k < nb; k += 32) {
      const float w = __ldg(mb + k);
#pragma unroll
      for (int q = 0; q < kFramesPerCta; ++q) acc[q] = fmaf(w, mag[q * nb + k], acc[q]);
    }
#pragma unroll
    for (int q = 0; q < kFramesPerCta; ++q) {
      float v = acc[q];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0 && f0 + q < n_frames) out[(long long)(f0 + q) * n_mels + j] = logf(fmaxf(v, 1e-5f));
    }
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
  int log2n = 0;
  while ((1 << log2n) < n_fft) ++log2n;
  const size_t smem = (size_t)(3 * n_fft + kFramesPerCta * (n_fft / 2 + 1)) * sizeof(float);
  static size_t attr_set = 0;
  if (smem > 48 * 1024 && smem > attr_set) {
    cudaFuncSetAttribute(stft_mel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr_set = smem;
  }
  const int ctas = (n_frames + kFramesPerCta - 1) / kFramesPerCta;
  launch_kernel(stft_mel_kernel, dim3(ctas), dim3(kStftThreads), smem, as_stream(stream), wav, n_samples, n_fft, log2n, hop,
                window, mel_basis, n_mels, n_frames, mag_workspace, out_logmel);
  return launched("ae_stft_mel");
}
