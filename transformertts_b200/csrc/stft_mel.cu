// STFT (n_fft 1024, hop 256, periodic Hann, reflect padding) -> magnitude -> banded mel filterbank -> log, fused.
// Replaces data/audio.py:81-92 (librosa.stft + librosa.feature.melspectrogram(S=|D|)) and the normalisers
// (data/audio.py:209-231).  HBM-bound by design: the audio is read once (neighbouring frames hit L1/L2) and only the
// (frames, 80) log-mel is written; the 513-bin spectrum never leaves the SM.
//
// One warp transforms TWO consecutive real frames as one 1024-point complex FFT (frame A -> real, frame B -> imag):
// 1024 = 32 x 32 Cooley-Tukey, both 32-point passes fully in registers (one column per lane), one shared-memory
// transpose in between, twiddles from a table in shared memory.  The two spectra are separated with
// X_A[k] = (Z[k] + conj Z[N-k])/2, X_B[k] = (Z[k] - conj Z[N-k])/(2i).
#include <math.h>

#include "../../include/ttsb.h"
#include "ttsb_common.cuh"
#include "ttsb_host.h"

namespace ttsb {

constexpr int NFFT = 1024;
constexpr int HOP = 256;
constexpr int NBINS = 513;
constexpr int STFT_WARPS = 8;
constexpr int MAX_MELS = 128;

__device__ float g_tw_re[NFFT];   // cos(2 pi i / 1024)
__device__ float g_tw_im[NFFT];   // -sin(2 pi i / 1024)
__device__ float g_window[NFFT];  // periodic Hann
__device__ int g_band[2 * MAX_MELS];

// cos/sin(2 pi j / 32), j = 0..15, for the in-register 32-point transforms
__device__ constexpr float C32[16] = {1.0f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254524f,
                                      0.70710678118654757f, 0.55557023301960229f, 0.38268343236508984f, 0.19509032201612833f,
                                      0.0f, -0.19509032201612819f, -0.38268343236508973f, -0.55557023301960196f,
                                      -0.70710678118654746f, -0.83146961230254535f, -0.92387953251128674f, -0.98078528040323043f};
__device__ constexpr float S32[16] = {0.0f, 0.19509032201612825f, 0.38268343236508978f, 0.55557023301960218f,
                                      0.70710678118654746f, 0.83146961230254524f, 0.92387953251128674f, 0.98078528040323043f,
                                      1.0f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254546f,
                                      0.70710678118654757f, 0.55557023301960218f, 0.38268343236508989f, 0.19509032201612861f};

__host__ __device__ constexpr int bitrev5(int i) {
  return ((i & 1) << 4) | ((i & 2) << 2) | (i & 4) | ((i & 8) >> 2) | ((i & 16) >> 4);
}

// In-place decimation-in-frequency radix-2 FFT of 32 complex values held in registers (forward, e^{-i...}).
// Result: X[bitrev5(i)] is left in slot i.
__device__ __forceinline__ void fft32(float (&re)[32], float (&im)[32]) {
#pragma unroll
  for (int len = 32; len >= 2; len >>= 1) {
    const int half = len >> 1;
    const int step = 32 / len;
#pragma unroll
    for (int start = 0; start < 32; start += len) {
#pragma unroll
      for (int j = 0; j < half; ++j) {
        const int a = start + j, b = a + half;
        const float tr = re[a] - re[b], ti = im[a] - im[b];
        re[a] += re[b];
        im[a] += im[b];
        const float c = C32[j * step], s = S32[j * step];  // W = c - i s
        re[b] = tr * c + ti * s;
        im[b] = ti * c - tr * s;
      }
    }
  }
}

__device__ __forceinline__ float sample_reflect(const float* __restrict__ clip, int n, int i) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return __ldg(clip + i);
}

__global__ void mel_band_kernel(const float* __restrict__ basis, int n_mels) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= n_mels) return;
  int lo = NBINS, hi = 0;
  for (int k = 0; k < NBINS; ++k)
    if (basis[(size_t)m * NBINS + k] != 0.f) {
      lo = min(lo, k);
      hi = max(hi, k + 1);
    }
  if (lo > hi) lo = hi = 0;
  g_band[2 * m] = lo;
  g_band[2 * m + 1] = hi;
}

struct StftSmem {
  float tw_re[NFFT];
  float tw_im[NFFT];
  float buf_re[STFT_WARPS][32 * 33];
  float buf_im[STFT_WARPS][32 * 33];
  float mag[STFT_WARPS][2][NBINS + 7];
};

__global__ void __launch_bounds__(STFT_WARPS * 32, 2)
stft_mel_kernel(const float* __restrict__ wav, int n_samples, int n_frames, int pairs_per_clip, const float* __restrict__ basis,
                int n_mels, int normalizer, float* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  StftSmem& sm = *reinterpret_cast<StftSmem*>(smem_raw);
  for (int i = threadIdx.x; i < NFFT; i += blockDim.x) {
    sm.tw_re[i] = g_tw_re[i];
    sm.tw_im[i] = g_tw_im[i];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int clip = blockIdx.y;
  const int pair = blockIdx.x * STFT_WARPS + warp;
  if (pair >= pairs_per_clip) return;
  const float* x = wav + (size_t)clip * n_samples;
  const int fA = 2 * pair, fB = 2 * pair + 1;
  const bool hasB = fB < n_frames;
  float* bre = sm.buf_re[warp];
  float* bim = sm.buf_im[warp];

  // ---- pass 1: lane n2 transforms x[32*n1 + n2], n1 = 0..31
  float re[32], im[32];
  const int sA = fA * HOP - NFFT / 2, sB = fB * HOP - NFFT / 2;
  const bool interior = sA >= 0 && (hasB ? sB : sA) + NFFT <= n_samples;
#pragma unroll
  for (int n1 = 0; n1 < 32; ++n1) {
    const int i = 32 * n1 + lane;
    const float w = g_window[i];
    float a, b;
    if (interior) {
      a = __ldg(x + sA + i);
      b = hasB ? __ldg(x + sB + i) : 0.f;
    } else {
      a = sample_reflect(x, n_samples, sA + i);
      b = hasB ? sample_reflect(x, n_samples, sB + i) : 0.f;
    }
    re[n1] = a * w;
    im[n1] = b * w;
  }
  fft32(re, im);
  // ---- twiddle by W_1024^(n2*k1) and transpose through shared memory: slot i holds k1 = bitrev5(i)
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const int k1 = bitrev5(i);
    const int tw = (lane * k1) & (NFFT - 1);
    const float c = sm.tw_re[tw], s = sm.tw_im[tw];
    bre[k1 * 33 + lane] = re[i] * c - im[i] * s;
    bim[k1 * 33 + lane] = re[i] * s + im[i] * c;
  }
  __syncwarp();
  // ---- pass 2: lane k1 transforms over n2
#pragma unroll
  for (int n2 = 0; n2 < 32; ++n2) {
    re[n2] = bre[lane * 33 + n2];
    im[n2] = bim[lane * 33 + n2];
  }
  __syncwarp();
  fft32(re, im);
  // Z[k1 + 32*k2] with k2 = bitrev5(i) -> shared (aliases the transpose buffer; 1024 <= 32*33)
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const int k = lane + 32 * bitrev5(i);
    bre[k] = re[i];
    bim[k] = im[i];
  }
  __syncwarp();
  // ---- separate the two real spectra, magnitudes
  float* magA = sm.mag[warp][0];
  float* magB = sm.mag[warp][1];
  for (int k = lane; k < NBINS; k += 32) {
    const int k2 = (NFFT - k) & (NFFT - 1);
    const float z1r = bre[k], z1i = bim[k], z2r = bre[k2], z2i = bim[k2];
    const float ar = 0.5f * (z1r + z2r), ai = 0.5f * (z1i - z2i);
    const float br = 0.5f * (z1i + z2i), bi = -0.5f * (z1r - z2r);
    magA[k] = sqrtf(ar * ar + ai * ai);
    magB[k] = sqrtf(br * br + bi * bi);
  }
  __syncwarp();
  // ---- banded mel filterbank + normaliser
  for (int m = lane; m < n_mels; m += 32) {
    const int lo = g_band[2 * m], hi = g_band[2 * m + 1];
    float accA = 0.f, accB = 0.f;
    for (int k = lo; k < hi; ++k) {
      const float wgt = __ldg(basis + (size_t)m * NBINS + k);
      accA = fmaf(wgt, magA[k], accA);
      accB = fmaf(wgt, magB[k], accB);
    }
    float ya, yb;
    if (normalizer == 0) {
      ya = logf(fmaxf(accA, 1e-5f));
      yb = logf(fmaxf(accB, 1e-5f));
    } else {
      ya = 20.f * log10f(fmaxf(accA, 1e-5f));
      yb = 20.f * log10f(fmaxf(accB, 1e-5f));
      ya = fminf(fmaxf((ya + 100.f) / 100.f, 0.f), 1.f) * 8.f - 4.f;
      yb = fminf(fmaxf((yb + 100.f) / 100.f, 0.f), 1.f) * 8.f - 4.f;
    }
    out[((size_t)clip * n_frames + fA) * n_mels + m] = ya;
    if (hasB) out[((size_t)clip * n_frames + fB) * n_mels + m] = yb;
  }
}

static int init_tables() {
  static PerDevice<bool> done_pd;  // __constant__ tables and function attributes are per device
  bool& done = done_pd.get();
  if (done) return 0;
  static float tr[NFFT], ti[NFFT], win[NFFT];
  for (int i = 0; i < NFFT; ++i) {
    const double a = 2.0 * M_PI * (double)i / (double)NFFT;
    tr[i] = (float)cos(a);
    ti[i] = (float)(-sin(a));
    win[i] = (float)(0.5 - 0.5 * cos(a));
  }
  TTSB_CUDA_OK(cudaMemcpyToSymbol(g_tw_re, tr, sizeof(tr)));
  TTSB_CUDA_OK(cudaMemcpyToSymbol(g_tw_im, ti, sizeof(ti)));
  TTSB_CUDA_OK(cudaMemcpyToSymbol(g_window, win, sizeof(win)));
  TTSB_CUDA_OK(cudaFuncSetAttribute(stft_mel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(StftSmem)));
  done = true;
  return 0;
}

}  // namespace ttsb

using namespace ttsb;

extern "C" int ttsb_stft_mel_log(const float* wav, int n_clips, int n_samples, const float* mel_basis, int n_mels,
                                 int normalizer, float* out, void* stream_v) {
  if (!wav || !mel_basis || !out) { set_last_error("ttsb_stft_mel_log: NULL tensor"); return TTSB_ERR_INVALID_ARGUMENT; }
  if (n_clips <= 0 || n_samples <= NFFT / 2 || n_mels <= 0 || n_mels > MAX_MELS || (normalizer != 0 && normalizer != 1)) {
    set_last_error("ttsb_stft_mel_log: need n_samples > 512, 0 < n_mels <= 128, normalizer in {0,1}");
    return TTSB_ERR_INVALID_ARGUMENT;
  }
  int rc = init_tables();
  if (rc) return rc;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  const int n_frames = 1 + n_samples / HOP;
  const int pairs = (n_frames + 1) / 2;
  mel_band_kernel<<<1, MAX_MELS, 0, stream>>>(mel_basis, n_mels);
  count_launch();
  dim3 grid((pairs + STFT_WARPS - 1) / STFT_WARPS, n_clips);
  stft_mel_kernel<<<grid, STFT_WARPS * 32, sizeof(StftSmem), stream>>>(wav, n_samples, n_frames, pairs, mel_basis, n_mels,
                                                                        normalizer, out);
  count_launch();
  return check_cuda(cudaGetLastError(), "stft_mel_kernel launch");
}
