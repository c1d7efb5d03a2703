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
#include <stdlib.h>

#include "../../include/ttsb.h"
#include "ttsb_common.cuh"
#include "fft32.cuh"
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

// ----------------------------------------------------------------------------------------------------
// Version 2 of the fused kernel (the one ttsb_stft_mel_log launches; the kernel above is kept as TTSB_STFT_V1=1 for A/B runs).
// ncu of version 1 (profiles/r02_stft.md): 38 % of the issue slots used, long-scoreboard stalls dominate (4.9 per issued
// instruction: global loads of samples, window, filter weights and band bounds with nothing to hide them behind), 17 M
// shared-memory bank conflicts in the mel loop, ~3700 instructions per frame pair.  Changes:
//   * persistent warps; the samples of the NEXT frame pair are requested into the (dead) FFT registers before the
//     magnitude / mel phase of the current pair, so their latency is hidden behind ~500 instructions of other work;
//   * window, twiddles (as (cos, -sin) pairs), packed filter weights and band bounds live in shared memory;
//   * complex values cross shared memory as 8-byte (re, im) pairs (half the load / store instructions);
//   * only the bins below the highest non-zero filter tap are separated (372 of 513 at f_max = 8 kHz), magnitudes with
//     x * rsqrt(x) instead of the IEEE sqrt sequence;
//   * the filterbank is stored banded and packed (727 weights), a lane owns bands m, m+32, m+64.
// ----------------------------------------------------------------------------------------------------
constexpr int V2_WARPS = 8;
constexpr int MAX_TAPS = 2048;          // packed non-zero filter weights (a triangular basis has < 2 * 513)

__device__ float g_mel_w[MAX_TAPS];     // row j: weights of bins [lo_j, lo_j + len_j) at offset off_j
__device__ int g_mel_lo[MAX_MELS], g_mel_len[MAX_MELS], g_mel_off[MAX_MELS];
__device__ int g_mel_kmax;              // 1 + highest bin with a non-zero weight
__device__ int g_mel_ok;                // 0: the basis is not banded enough for the packed table

__global__ void mel_pack_kernel(const float* __restrict__ basis, int n_mels) {
  __shared__ int lo_s[MAX_MELS], len_s[MAX_MELS];
  const int m = threadIdx.x;
  if (m < n_mels) {
    int lo = NBINS, hi = 0;
    for (int k = 0; k < NBINS; ++k)
      if (basis[(size_t)m * NBINS + k] != 0.f) {
        lo = min(lo, k);
        hi = max(hi, k + 1);
      }
    if (lo > hi) lo = hi = 0;
    lo_s[m] = lo;
    len_s[m] = hi - lo;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int off = 0, kmax = 0, ok = 1;
    for (int j = 0; j < n_mels; ++j) {
      g_mel_lo[j] = lo_s[j];
      g_mel_len[j] = len_s[j];
      g_mel_off[j] = off;
      kmax = max(kmax, lo_s[j] + len_s[j]);
      off += len_s[j];
      if (off > MAX_TAPS) { ok = 0; break; }
    }
    g_mel_kmax = kmax;
    g_mel_ok = ok;
  }
  __syncthreads();
  if (m < n_mels && g_mel_ok)
    for (int i = 0; i < len_s[m]; ++i) g_mel_w[g_mel_off[m] + i] = basis[(size_t)m * NBINS + lo_s[m] + i];
}

struct StftSmem2 {
  float2 tw[NFFT];                       // (cos, -sin)(2 pi i / 1024)
  float window[NFFT];
  float mel_w[MAX_TAPS];
  int mel_lo[MAX_MELS], mel_len[MAX_MELS], mel_off[MAX_MELS];
  float2 buf[V2_WARPS][32 * 33];         // per-warp transpose / spectrum buffer; later (|X_A|, |X_B|) per bin, in place
};

__device__ __forceinline__ void load_pair(const float* __restrict__ wav, int n_samples, int n_frames, int pairs_per_clip, int item,
                                          int lane, float (&re)[32], float (&im)[32]) {
  const int clip = item / pairs_per_clip, pair = item - clip * pairs_per_clip;
  const float* x = wav + (size_t)clip * n_samples;
  const int fA = 2 * pair, fB = fA + 1;
  const bool hasB = fB < n_frames;
  const int sA = fA * HOP - NFFT / 2, sB = fB * HOP - NFFT / 2;
  if (sA >= 0 && (hasB ? sB : sA) + NFFT <= n_samples) {
#pragma unroll
    for (int n1 = 0; n1 < 32; ++n1) {
      re[n1] = __ldg(x + sA + 32 * n1 + lane);
      im[n1] = hasB ? __ldg(x + sB + 32 * n1 + lane) : 0.f;
    }
  } else {
#pragma unroll
    for (int n1 = 0; n1 < 32; ++n1) {
      re[n1] = sample_reflect(x, n_samples, sA + 32 * n1 + lane);
      im[n1] = hasB ? sample_reflect(x, n_samples, sB + 32 * n1 + lane) : 0.f;
    }
  }
}

__global__ void __launch_bounds__(V2_WARPS * 32, 2)
stft_mel_v2_kernel(const float* __restrict__ wav, int n_samples, int n_frames, int pairs_per_clip, int n_items, int n_mels,
                   int normalizer, float* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  StftSmem2& sm = *reinterpret_cast<StftSmem2*>(smem_raw);
  for (int i = threadIdx.x; i < NFFT; i += blockDim.x) {
    sm.tw[i] = make_float2(g_tw_re[i], g_tw_im[i]);
    sm.window[i] = g_window[i];
  }
  for (int i = threadIdx.x; i < MAX_TAPS; i += blockDim.x) sm.mel_w[i] = g_mel_w[i];
  for (int i = threadIdx.x; i < MAX_MELS; i += blockDim.x) {
    sm.mel_lo[i] = g_mel_lo[i];
    sm.mel_len[i] = i < n_mels ? g_mel_len[i] : 0;
    sm.mel_off[i] = g_mel_off[i];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kmax = g_mel_kmax;
  const int gw = blockIdx.x * V2_WARPS + warp, nw = gridDim.x * V2_WARPS;
  float2* buf = sm.buf[warp];
  float re[32], im[32];
  int item = gw;
  if (item < n_items) load_pair(wav, n_samples, n_frames, pairs_per_clip, item, lane, re, im);
  for (; item < n_items; item += nw) {
    const int clip = item / pairs_per_clip, pair = item - clip * pairs_per_clip;
    const int fA = 2 * pair, fB = fA + 1;
    const bool hasB = fB < n_frames;
    // ---- window, pass 1 (lane n2 transforms over n1)
#pragma unroll
    for (int n1 = 0; n1 < 32; ++n1) {
      const float w = sm.window[32 * n1 + lane];
      re[n1] *= w;
      im[n1] *= w;
    }
    fft32(re, im);
    // ---- twiddle by W_1024^(n2*k1) and transpose through shared memory: slot i holds k1 = bitrev5(i)
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const int k1 = bitrev5(i);
      const float2 t = sm.tw[(lane * k1) & (NFFT - 1)];
      buf[k1 * 33 + lane] = make_float2(re[i] * t.x - im[i] * t.y, re[i] * t.y + im[i] * t.x);
    }
    __syncwarp();
    // ---- pass 2 (lane k1 transforms over n2)
#pragma unroll
    for (int n2 = 0; n2 < 32; ++n2) {
      const float2 v = buf[lane * 33 + n2];
      re[n2] = v.x;
      im[n2] = v.y;
    }
    __syncwarp();
    fft32(re, im);
#pragma unroll
    for (int i = 0; i < 32; ++i) buf[lane + 32 * bitrev5(i)] = make_float2(re[i], im[i]);   // Z[k1 + 32 k2], natural order
    __syncwarp();
    // ---- the FFT registers are dead: request the next pair's samples now (consumed at the top of the next iteration)
    if (item + nw < n_items) load_pair(wav, n_samples, n_frames, pairs_per_clip, item + nw, lane, re, im);
    // ---- separate the two real spectra (bins below the highest filter tap); the magnitudes (|X_A[k]|, |X_B[k]|) replace
    //      Z[k] in place: bin k is only ever read together with bin N-k > 512 by the lane that owns k
    for (int k = lane; k < kmax; k += 32) {
      const float2 z1 = buf[k], z2 = buf[(NFFT - k) & (NFFT - 1)];
      const float ar = z1.x + z2.x, ai = z1.y - z2.y;      // 2 X_A[k]
      const float br = z1.y + z2.y, bi = z1.x - z2.x;      // 2 X_B[k] (up to a rotation by -i: same magnitude)
      const float pa = ar * ar + ai * ai, pb = br * br + bi * bi;
      buf[k] = make_float2(pa > 0.f ? 0.5f * pa * rsqrtf(pa) : 0.f, pb > 0.f ? 0.5f * pb * rsqrtf(pb) : 0.f);
    }
    __syncwarp();
    // ---- banded mel filterbank + normaliser: lane owns bands lane, lane + 32, lane + 64, ...
    for (int m = lane; m < n_mels; m += 32) {
      const int lo = sm.mel_lo[m], len = sm.mel_len[m];
      const float* w = sm.mel_w + sm.mel_off[m];
      float accA = 0.f, accB = 0.f;
      for (int i = 0; i < len; ++i) {
        const float wgt = w[i];
        const float2 mg = buf[lo + i];
        accA = fmaf(wgt, mg.x, accA);
        accB = fmaf(wgt, mg.y, accB);
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
    __syncwarp();   // buf / mag are rewritten by the next iteration
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
  TTSB_CUDA_OK(cudaFuncSetAttribute(stft_mel_v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(StftSmem2)));
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
  static const bool use_v1 = getenv("TTSB_STFT_V1") != nullptr;
  if (!use_v1) {
    // the packed filterbank table is rebuilt when another basis pointer shows up (a basis is a per-Audio constant)
    static PerDevice<const float*> packed_for;
    int ok = 1;
    if (packed_for.get() != mel_basis) {
      mel_pack_kernel<<<1, MAX_MELS, 0, stream>>>(mel_basis, n_mels);
      count_launch();
      TTSB_CUDA_OK(cudaMemcpyFromSymbolAsync(&ok, g_mel_ok, sizeof(int), 0, cudaMemcpyDeviceToHost, stream));
      TTSB_CUDA_OK(cudaStreamSynchronize(stream));
      packed_for.get() = ok ? mel_basis : nullptr;
    }
    if (ok) {
      const int n_items = pairs * n_clips;
      const int grid2 = min((n_items + V2_WARPS - 1) / V2_WARPS, 2 * num_sms());
      stft_mel_v2_kernel<<<grid2, V2_WARPS * 32, sizeof(StftSmem2), stream>>>(wav, n_samples, n_frames, pairs, n_items, n_mels, normalizer, out);
      count_launch();
      return check_cuda(cudaGetLastError(), "stft_mel_v2_kernel launch");
    }
  }
  mel_band_kernel<<<1, MAX_MELS, 0, stream>>>(mel_basis, n_mels);
  count_launch();
  dim3 grid((pairs + STFT_WARPS - 1) / STFT_WARPS, n_clips);
  stft_mel_kernel<<<grid, STFT_WARPS * 32, sizeof(StftSmem), stream>>>(wav, n_samples, n_frames, pairs, mel_basis, n_mels,
                                                                        normalizer, out);
  count_launch();
  return check_cuda(cudaGetLastError(), "stft_mel_kernel launch");
}
