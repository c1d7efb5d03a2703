// 32-point complex FFT on register-resident values (shared by stft_mel.cu and griffin_lim.cu): decimation in frequency,
// radix 2, forward (e^{-i...}); X[bitrev5(i)] is left in slot i.  The twiddle W_32^t = cos(2 pi t / 32) - i sin(2 pi t / 32)
// of every butterfly is a compile-time constant after unrolling; under IEEE rules the compiler may NOT drop a multiplication
// by 1.0f or 0.0f, so the trivial cases are written out: t = 0 (no multiply), t = 8 (multiply by -i: swap), t = 4 / 12
// (45 degrees: two adds + two multiplies); only 20 of the 80 butterflies keep a general complex multiply.
#pragma once

namespace ttsb {

__device__ constexpr float FFT32_C[16] = {1.0f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254524f,
                                          0.70710678118654757f, 0.55557023301960229f, 0.38268343236508984f, 0.19509032201612833f,
                                          0.0f, -0.19509032201612819f, -0.38268343236508973f, -0.55557023301960196f,
                                          -0.70710678118654746f, -0.83146961230254535f, -0.92387953251128674f, -0.98078528040323043f};
__device__ constexpr float FFT32_S[16] = {0.0f, 0.19509032201612825f, 0.38268343236508978f, 0.55557023301960218f,
                                          0.70710678118654746f, 0.83146961230254524f, 0.92387953251128674f, 0.98078528040323043f,
                                          1.0f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254546f,
                                          0.70710678118654757f, 0.55557023301960218f, 0.38268343236508989f, 0.19509032201612861f};

__host__ __device__ constexpr int bitrev5(int i) {
  return ((i & 1) << 4) | ((i & 2) << 2) | (i & 4) | ((i & 8) >> 2) | ((i & 16) >> 4);
}

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
        const int t = j * step;                        // compile-time after unrolling
        const float tr = re[a] - re[b], ti = im[a] - im[b];
        re[a] += re[b];
        im[a] += im[b];
        if (t == 0) {                                  // W = 1
          re[b] = tr;
          im[b] = ti;
        } else if (t == 8) {                           // W = -i
          re[b] = ti;
          im[b] = -tr;
        } else if (t == 4) {                           // W = (1 - i) / sqrt(2)
          re[b] = (tr + ti) * 0.70710678118654757f;
          im[b] = (ti - tr) * 0.70710678118654757f;
        } else if (t == 12) {                          // W = -(1 + i) / sqrt(2)
          re[b] = (ti - tr) * 0.70710678118654757f;
          im[b] = -(tr + ti) * 0.70710678118654757f;
        } else {
          const float c = FFT32_C[t], s = FFT32_S[t];  // W = c - i s
          re[b] = tr * c + ti * s;
          im[b] = ti * c - tr * s;
        }
      }
    }
  }
}

}  // namespace ttsb
