// Fused variable-length self-attention on tcgen05 (flash-style, never materialises the (B,H,T,T) tensors):
//   per CTA one (batch row, head, 128-query tile); loop over 64-key tiles:
//     S = Q K^T  (tcgen05.mma, fp32 in TMEM)  ->  online softmax in registers (one thread per query row, exp2 domain,
//     keys >= kv_len masked)  ->  P (bf16 hi/lo or fp16) written to 128B-swizzled shared memory  ->  O += P V
//     (tcgen05.mma; V is read MN-major straight from the QKV buffer, no transposed copy)  ->  O rescaled in TMEM only
//     when a row maximum moved.
// Replaces model/layers.py:123-129,138-147 (split/merge heads) and :176-195 (ScaledDotProductAttention) of the
// reference, which materialise logits / softmax / dropout tensors of shape (B,H,T,T) in fp32.
//
// Masking semantics: the reference adds mask*(-1e9) to the logits of padded KEYS (layers.py:186-187); after the
// fp32 add every such logit equals -1e9 and its softmax weight underflows to exactly 0 whenever the row has at least
// one valid key, so skipping those keys is exact.  Rows of padded QUERIES are zeroed by the caller's row mask
// (layers.py:229,262), so they are not computed here (written as zeros).
#include <cstdlib>
#include <cuda_fp16.h>

#include "../../include/ttsb.h"
#include "ttsb_common.cuh"
#include "ttsb_host.h"

namespace ttsb {

constexpr int ATT_BQ = 128;
constexpr int ATT_BKV = 64;
// narrow layout: warps 0-3 softmax/epilogue (TMEM lane quarters 0-3), warp 4 TMA + MMA issue               (160 threads)
// wide layout:   warps 0-7 softmax/epilogue, warp w owns lane quarter w&3 and column half w>>2 of every S/P/O tile,
//                warp 8 TMA + MMA issue                                                                        (288 threads)
constexpr int ATT_THREADS_NARROW = 160;
constexpr int ATT_THREADS_WIDE = 288;

// single MUFU.EX2 (no denormal fix-up code: p < 2^-126 flushes to 0, far below bf16/fp16 resolution of P)
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct MhaKParams {
  int B, T, H;
  int Tk;            // key/value rows (== T for self-attention)
  int causal;        // keys > query index are masked
  int full_queries;  // 0: query rows >= kv_len[b] are written as zeros
  int q_col0, k_col0, v_col0;
  const int* kv_len;
  __nv_bfloat16* out_hi;
  __nv_bfloat16* out_lo;
  int ld_out;
  float scale_log2;  // log2(e)/sqrt(dh)
#ifdef TTSB_ATT_TRACE
  long long* trace;  // debug build only: clock64 stamps of CTA (0,0,0): [2 roles][64 tiles][8 events]
#endif
};

#ifdef TTSB_ATT_TRACE
#define ATT_TRACE(role, j, ev)                                                                         \
  do {                                                                                                 \
    if (p.trace && (threadIdx.x & 31) == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (j) < 64)                 \
      p.trace[((role) * 64 + (j)) * 8 + (ev)] = clock64();                                             \
  } while (0)
#else
#define ATT_TRACE(role, j, ev) do {} while (0)
#endif

template <int DH, bool kSplit>
struct MhaCfg {
  static constexpr int kPlanes = kSplit ? 2 : 1;
  static constexpr int Q_BYTES = ATT_BQ * DH * 2;    // DH/64 panels of [128 x 64]
  static constexpr int K_BYTES = ATT_BKV * DH * 2;   // DH/64 panels of [64 x 64]
  static constexpr int V_BYTES = DH * ATT_BKV * 2;   // DH/64 boxes of [64 keys x 64 dh] (MN-major B operand)
  static constexpr int P_BYTES = ATT_BQ * ATT_BKV * 2;
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_K = OFF_Q + kPlanes * Q_BYTES;
  // three key/value slots: slot j%3 holds K_j until S_j has completed, then V_j until PV_j has completed, then K_{j+3}.
  // Every load is therefore issued about one key tile ahead of its consumer (no TMA latency on the critical path).
  static constexpr int K_STAGE = kPlanes * K_BYTES;
  static constexpr int OFF_P = OFF_K + 3 * K_STAGE;
  static constexpr int OFF_BAR = OFF_P + kPlanes * P_BYTES;
  static constexpr int OFF_RED = OFF_BAR + 128;      // wide layout: row max [2 tiles][2 halves][128] + row sum [2][128]
  static constexpr int kSmemBytes = OFF_RED + 3072 + 1024;
  static constexpr int kTmemCols = (2 * ATT_BKV + DH) <= 128 ? 128 : ((2 * ATT_BKV + DH) <= 256 ? 256 : 512);
  static constexpr int S_COL = 0;             // two S accumulators of ATT_BKV columns
  static constexpr int O_COL = 2 * ATT_BKV;
};

template <int DH, bool kSplit, bool kF16, bool kWide>
__global__ void __launch_bounds__(kWide ? ATT_THREADS_WIDE : ATT_THREADS_NARROW, (kWide && !kSplit && DH <= 128) ? 2 : 1)
mha_tc_kernel(const __grid_constant__ CUtensorMap tmQh, const __grid_constant__ CUtensorMap tmQl,
              const __grid_constant__ CUtensorMap tmKh, const __grid_constant__ CUtensorMap tmKl, const MhaKParams p) {
  using Cfg = MhaCfg<DH, kSplit>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* bar_q = bars + 0;
  uint64_t* bar_kv = bars + 1;  // [3] slot full: completes once for K_j (parity 0) and once for V_j (parity 1) per use
  uint64_t* bar_s = bars + 4;   // [2]
  uint64_t* bar_p = bars + 6;
  uint64_t* bar_pv = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  constexpr int NH = kWide ? 2 : 1;        // column halves per S/P/O tile
  constexpr int kMmaWarp = 4 * NH;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int quarter = warp & 3;            // TMEM lane quarter this warp may access
  const int half = kWide ? (warp >> 2) & 1 : 0;
  const int q0 = blockIdx.x * ATT_BQ;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  int len = __ldg(p.kv_len + b);
  len = len < 0 ? 0 : (len > p.Tk ? p.Tk : len);

  if (len == 0 || (!p.full_queries && q0 >= len)) {
    // whole query tile is padding (or no key at all): defined (zero) output, no tensor work
    if (warp < kMmaWarp) {
      const int t = q0 + quarter * 32 + lane;
      if (t < p.T) {
        const size_t o = ((size_t)b * p.T + t) * p.ld_out + h * DH + half * (DH / NH);
        for (int c = 0; c < DH / NH; c += 8) {
          st_global_v4(p.out_hi + o + c, 0, 0, 0, 0);
          if (p.out_lo) st_global_v4(p.out_lo + o + c, 0, 0, 0, 0);
        }
      }
    }
    return;
  }
  int n_kv = (len + ATT_BKV - 1) / ATT_BKV;
  if (p.causal) n_kv = min(n_kv, (q0 + ATT_BQ + ATT_BKV - 1) / ATT_BKV);  // key tiles beyond the last query row are all masked

  if (threadIdx.x == 0) {
    mbar_init(bar_q, 1);
    mbar_init(bar_kv, 1);
    mbar_init(bar_kv + 1, 1);
    mbar_init(bar_kv + 2, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_s + 1, 1);
    mbar_init(bar_p, 4 * NH);
    mbar_init(bar_pv, 1);
    fence_mbar_init();
  }
  if (warp == kMmaWarp) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == kMmaWarp) {
    {
      // ===================== TMA + MMA issue warp: warp-uniform control flow, one elected lane issues =====================
      const bool leader = elect_one();
      uint8_t* sQ = smem + Cfg::OFF_Q;
      uint8_t* sKV = smem + Cfg::OFF_K;
      uint8_t* sP = smem + Cfg::OFF_P;
      const uint32_t idesc_s = kF16 ? make_idesc_f16(ATT_BQ, ATT_BKV) : make_idesc_bf16(ATT_BQ, ATT_BKV);
      const uint32_t idesc_o = (kF16 ? make_idesc_f16(ATT_BQ, DH) : make_idesc_bf16(ATT_BQ, DH)) | (1u << 16);  // B (= V) MN-major
      const uint32_t t_s = tmem_base + Cfg::S_COL;
      const uint32_t t_o = tmem_base + Cfg::O_COL;

      if (leader) {
        mbar_arrive_expect_tx(bar_q, Cfg::kPlanes * Cfg::Q_BYTES);
        for (int pn = 0; pn < DH / 64; ++pn) {
          tma_load_3d(&tmQh, bar_q, sQ + pn * (ATT_BQ * 128), p.q_col0 + h * DH + pn * 64, q0, b);
          if (kSplit) tma_load_3d(&tmQl, bar_q, sQ + Cfg::Q_BYTES + pn * (ATT_BQ * 128), p.q_col0 + h * DH + pn * 64, q0, b);
        }
      }
      // K_j and V_j are the same [64 keys x 64 columns] boxes of the QKV buffer, at the K resp. V columns
      auto load_kv = [&](int j, int col0) {
        uint8_t* dst = sKV + (j % 3) * Cfg::K_STAGE;
        uint64_t* bar = bar_kv + (j % 3);
        if (leader) {
          mbar_arrive_expect_tx(bar, Cfg::kPlanes * Cfg::K_BYTES);
          for (int pn = 0; pn < DH / 64; ++pn) {
            tma_load_3d(&tmKh, bar, dst + pn * (ATT_BKV * 128), col0 + h * DH + pn * 64, j * ATT_BKV, b);
            if (kSplit) tma_load_3d(&tmKl, bar, dst + Cfg::K_BYTES + pn * (ATT_BKV * 128), col0 + h * DH + pn * 64, j * ATT_BKV, b);
          }
        }
      };
      // S_j = Q K_j^T into S accumulator j&1
      auto issue_s = [&](int j) {
        const uint8_t* kst = sKV + (j % 3) * Cfg::K_STAGE;
        const uint32_t ts = t_s + (j & 1) * ATT_BKV;
        mbar_wait(bar_kv + (j % 3), 0);
        tc_fence_after();
        if (leader) {
#pragma unroll
        for (int kk = 0; kk < DH / 16; ++kk) {
          const uint64_t a = make_smem_desc_sw128(smem_u32(sQ + (kk / 4) * (ATT_BQ * 128))) + 2 * (kk % 4);
          const uint64_t bb = make_smem_desc_sw128(smem_u32(kst + (kk / 4) * (ATT_BKV * 128))) + 2 * (kk % 4);
          umma_bf16(ts, a, bb, idesc_s, kk != 0);
        }
        if (kSplit) {
#pragma unroll
          for (int kk = 0; kk < DH / 16; ++kk) {
            const uint64_t a = make_smem_desc_sw128(smem_u32(sQ + Cfg::Q_BYTES + (kk / 4) * (ATT_BQ * 128))) + 2 * (kk % 4);
            const uint64_t bb = make_smem_desc_sw128(smem_u32(kst + (kk / 4) * (ATT_BKV * 128))) + 2 * (kk % 4);
            umma_bf16(ts, a, bb, idesc_s, 1);
          }
#pragma unroll
          for (int kk = 0; kk < DH / 16; ++kk) {
            const uint64_t a = make_smem_desc_sw128(smem_u32(sQ + (kk / 4) * (ATT_BQ * 128))) + 2 * (kk % 4);
            const uint64_t bb = make_smem_desc_sw128(smem_u32(kst + Cfg::K_BYTES + (kk / 4) * (ATT_BKV * 128))) + 2 * (kk % 4);
            umma_bf16(ts, a, bb, idesc_s, 1);
          }
        }
        umma_commit(bar_s + (j & 1));
        }
      };
      load_kv(0, p.k_col0);
      if (n_kv > 1) load_kv(1, p.k_col0);
      if (n_kv > 2) load_kv(2, p.k_col0);
      mbar_wait(bar_q, 0);
      issue_s(0);
      mbar_wait(bar_s, 0);        // S_0 complete: its K slot takes V_0
      load_kv(0, p.v_col0);
      for (int j = 0; j < n_kv; ++j) {
        const uint32_t ph = j & 1;
        // ---- software pipeline: S_{j+1} goes to the tensor pipe now, so the softmax warps find it ready when they finish
        //      tile j (its accumulator was last read for tile j-1, whose bar_p this thread has already seen)
        ATT_TRACE(1, j, 0);
        if (j + 1 < n_kv) issue_s(j + 1);
        ATT_TRACE(1, j, 1);
        // ---- slot (j-1)%3 is free once PV_{j-1} has completed: refill it with K_{j+2}
        if (j >= 1 && j + 2 < n_kv) {
          mbar_wait(bar_pv, (j - 1) & 1);
          load_kv(j + 2, p.k_col0);
        }
        // ---- O += P V
        const uint8_t* sV = sKV + (j % 3) * Cfg::K_STAGE;
        ATT_TRACE(1, j, 2);
        mbar_wait(bar_kv + (j % 3), 1);
        ATT_TRACE(1, j, 3);
        mbar_wait(bar_p, ph);
        ATT_TRACE(1, j, 4);
        tc_fence_after();
        if (leader) {
#pragma unroll
        for (int kk = 0; kk < ATT_BKV / 16; ++kk) {
          const uint64_t a = make_smem_desc_sw128(smem_u32(sP)) + 2 * kk;
          const uint64_t bb = make_smem_desc_mn_sw128(smem_u32(sV), ATT_BKV * 128) + 128 * kk;  // 16 keys = 2048 B
          umma_bf16(t_o, a, bb, idesc_o, (j | kk) != 0);
        }
        if (kSplit) {
#pragma unroll
          for (int kk = 0; kk < ATT_BKV / 16; ++kk) {
            const uint64_t a = make_smem_desc_sw128(smem_u32(sP + Cfg::P_BYTES)) + 2 * kk;
            const uint64_t bb = make_smem_desc_mn_sw128(smem_u32(sV), ATT_BKV * 128) + 128 * kk;
            umma_bf16(t_o, a, bb, idesc_o, 1);
          }
#pragma unroll
          for (int kk = 0; kk < ATT_BKV / 16; ++kk) {
            const uint64_t a = make_smem_desc_sw128(smem_u32(sP)) + 2 * kk;
            const uint64_t bb = make_smem_desc_mn_sw128(smem_u32(sV + Cfg::K_BYTES), ATT_BKV * 128) + 128 * kk;
            umma_bf16(t_o, a, bb, idesc_o, 1);
          }
        }
        umma_commit(bar_pv);
        }
        ATT_TRACE(1, j, 5);
        // ---- S_{j+1} (issued above, ahead of PV_j) completes first: its K slot takes V_{j+1}, a whole softmax ahead of PV_{j+1}
        if (j + 1 < n_kv) {
          mbar_wait(bar_s + ((j + 1) & 1), ((j + 1) >> 1) & 1);
          load_kv(j + 1, p.v_col0);
        }
        ATT_TRACE(1, j, 6);
      }
    }
  } else {
    // ===================== softmax / epilogue warps =====================
    // one thread per (query row, column half): with the wide layout two partner warps (w, w+4) share a row, each
    // reading its half of S, writing its half of P and owning its half of O; row max / row sum are combined through
    // shared memory behind a 64-thread named barrier.
    constexpr int SW = ATT_BKV / NH;   // S / P columns of this thread
    constexpr int OW = DH / NH;        // O columns of this thread
    const int row = quarter * 32 + lane;
    const int tq = q0 + row;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    const uint32_t t_s = tmem_base + lane_base + Cfg::S_COL + half * SW;
    const uint32_t t_o = tmem_base + lane_base + Cfg::O_COL + half * OW;
    uint8_t* sP = smem + Cfg::OFF_P;
    float* red_max = reinterpret_cast<float*>(smem + Cfg::OFF_RED);          // [2][2][128]
    float* red_sum = red_max + 2 * 2 * ATT_BQ;                               // [2][128]
    float m = -INFINITY, l = 0.f;
    uint32_t r[16];
    for (int j = 0; j < n_kv; ++j) {
      if (threadIdx.x == 0) ATT_TRACE(0, j, 0);
      mbar_wait(bar_s + (j & 1), (j >> 1) & 1);
      if (threadIdx.x == 0) ATT_TRACE(0, j, 1);
      tc_fence_after();
      const uint32_t t_sj = t_s + (j & 1) * ATT_BKV;
      // raw scores (un-scaled): the running max is kept in raw units, the softmax scale is folded into one FFMA per
      // element: p = exp2(s * scale_log2 - m * scale_log2)
      float s[SW];
      {
        // all TMEM loads are issued back to back and waited for once
        uint32_t qq[SW / 16][16];
#pragma unroll
        for (int c = 0; c < SW / 16; ++c) tmem_ld16(t_sj + c * 16, qq[c]);
        tmem_wait_ld();
        if (threadIdx.x == 0) ATT_TRACE(0, j, 2);
#pragma unroll
        for (int c = 0; c < SW / 16; ++c)
#pragma unroll
          for (int i = 0; i < 16; ++i) s[c * 16 + i] = __uint_as_float(qq[c][i]);
      }
      if ((j + 1) * ATT_BKV > len) {  // only the last key tile can hold padded keys
#pragma unroll
        for (int i = 0; i < SW; ++i)
          if (j * ATT_BKV + half * SW + i >= len) s[i] = -INFINITY;
      }
      if (p.causal && (j + 1) * ATT_BKV - 1 > q0) {  // look-ahead mask: tiles that reach past the first query row of this CTA
#pragma unroll
        for (int i = 0; i < SW; ++i)
          if (j * ATT_BKV + half * SW + i > tq) s[i] = -INFINITY;
      }
      float mx = s[0];
#pragma unroll
      for (int i = 1; i < SW; ++i) mx = fmaxf(mx, s[i]);
      if (kWide) {
        float* rm = red_max + (j & 1) * 2 * ATT_BQ;
        rm[half * ATT_BQ + row] = mx;
        asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");
        mx = fmaxf(mx, rm[(half ^ 1) * ATT_BQ + row]);
      }
      // lazy running max: the reference point only moves when the row max grew by more than 2^8 (in exponent units), so
      // P stays <= 256 (exact range for fp16/bf16, fp32 accumulation) and the O rescale below becomes rare.  l and O
      // use the same reference, so the final O / l is unchanged.
      const float m_cand = fmaxf(m, mx);
      const bool grow = (m_cand - m) * p.scale_log2 > 8.f;        // first tile: m = -inf -> true
      const float m_new = grow ? m_cand : m;
      const float alpha = grow ? fast_exp2((m - m_new) * p.scale_log2) : 1.f;  // first tile: exp2(-inf) = 0
      const float neg_m = -m_new * p.scale_log2;
      float psum = 0.f;
#pragma unroll
      for (int i = 0; i < SW; ++i) {
        s[i] = fast_exp2(fmaf(s[i], p.scale_log2, neg_m));
        psum += s[i];
      }
      l = l * alpha + psum;   // wide layout: partial sum over this thread's column half
      m = m_new;
      if (threadIdx.x == 0) ATT_TRACE(0, j, 3);
      if (j > 0) {
        mbar_wait(bar_pv, (j - 1) & 1);  // O_{j-1} complete, P buffer free
        if (threadIdx.x == 0) ATT_TRACE(0, j, 4);
        tc_fence_after();
        if (__any_sync(0xffffffffu, alpha != 1.f)) {
#pragma unroll
          for (int c = 0; c < OW / 16; ++c) {
            tmem_ld16(t_o + c * 16, r);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
            tmem_st16(t_o + c * 16, r);
          }
          tmem_wait_st();
        }
      }
      if (threadIdx.x == 0) ATT_TRACE(0, j, 5);
      // P -> shared memory, K-major rows of 128 B with the 128B swizzle (16-byte chunk index ^= row & 7)
#pragma unroll
      for (int c8 = 0; c8 < SW / 8; ++c8) {
        const int ch = half * (SW / 8) + c8;
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (kF16) {
            const __half2 hh = __floats2half2_rn(s[c8 * 8 + 2 * i], s[c8 * 8 + 2 * i + 1]);
            hi[i] = *reinterpret_cast<const uint32_t*>(&hh);
            lo[i] = 0;
          } else {
            __nv_bfloat16 h0, l0, h1, l1;
            split_bf16(s[c8 * 8 + 2 * i], h0, l0);
            split_bf16(s[c8 * 8 + 2 * i + 1], h1, l1);
            hi[i] = pack_bf16(h0, h1);
            lo[i] = pack_bf16(l0, l1);
          }
        }
        const uint32_t off = row * 128 + ((ch ^ (row & 7)) << 4);
        st_shared_v4(sP + off, hi[0], hi[1], hi[2], hi[3]);
        if (kSplit) st_shared_v4(sP + Cfg::P_BYTES + off, lo[0], lo[1], lo[2], lo[3]);
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_p);
      if (threadIdx.x == 0) ATT_TRACE(0, j, 6);
    }
    // ---- epilogue: O / l -> bf16 hi/lo at columns [h*DH, (h+1)*DH)
    if (kWide) {
      red_sum[half * ATT_BQ + row] = l;
      asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");
      l += red_sum[(half ^ 1) * ATT_BQ + row];
    }
    mbar_wait(bar_pv, (n_kv - 1) & 1);
    tc_fence_after();
    const float inv = l > 0.f ? 1.f / l : 0.f;
    const size_t o = ((size_t)b * p.T + (tq < p.T ? tq : 0)) * p.ld_out + h * DH + half * OW;
#pragma unroll
    for (int c = 0; c < OW / 16; ++c) {
      tmem_ld16(t_o + c * 16, r);
      tmem_wait_ld();
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        __nv_bfloat16 h0, l0, h1, l1;
        split_bf16(__uint_as_float(r[2 * i]) * inv, h0, l0);
        split_bf16(__uint_as_float(r[2 * i + 1]) * inv, h1, l1);
        hi[i] = pack_bf16(h0, h1);
        lo[i] = pack_bf16(l0, l1);
      }
      if (tq < p.T) {
        st_global_v8(p.out_hi + o + c * 16, hi);
        if (p.out_lo) st_global_v8(p.out_lo + o + c * 16, lo);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ----------------------------------------------------------------------------------------------------
// SIMT bring-up kernel (one block per (b,h,query)); also produces the reference-exact attention weights
// of one batch row on request (padded keys get logit -1e9 exactly as layers.py:186-187).
// ----------------------------------------------------------------------------------------------------
struct MhaSimtPtrs {
  const __nv_bfloat16* qk_hi;
  const __nv_bfloat16* qk_lo;
  int ld_qk;
  const __nv_bfloat16* kv_hi;  // == qk_hi for self-attention
  const __nv_bfloat16* kv_lo;
  int ld_kv;
  int dh;
  int f16;           // operands hold IEEE fp16 bit patterns
  float scale;       // 1/sqrt(dh)
  float* weights;    // (H,T,Tk), (B,H,T,Tk) with weights_all, or null
  int weights_b;
  int weights_all;
  int weights_only;  // 1: only fill weights (for batch row weights_b, or every row with weights_all)
};

__device__ __forceinline__ float ld_split(const __nv_bfloat16* hi, const __nv_bfloat16* lo, size_t i, int f16 = 0) {
  if (f16) return __half2float(reinterpret_cast<const __half*>(hi)[i]);
  float v = __bfloat162float(hi[i]);
  if (lo) v += __bfloat162float(lo[i]);
  return v;
}

__global__ void mha_simt_kernel(const MhaKParams p, const MhaSimtPtrs q) {
  const int tq = blockIdx.x, h = blockIdx.y;
  const int b = (q.weights_only && !q.weights_all) ? q.weights_b : blockIdx.z;
  extern __shared__ float sm[];  // Tk logits + dh query + scratch
  float* logit = sm;
  float* qv = sm + p.Tk;
  __shared__ float red[32];
  const int len = min(max(p.kv_len[b], 0), p.Tk);
  const size_t qrow = ((size_t)b * p.T + tq) * q.ld_qk;
  for (int c = threadIdx.x; c < q.dh; c += blockDim.x) qv[c] = ld_split(q.qk_hi, q.qk_lo, qrow + p.q_col0 + h * q.dh + c, q.f16);
  __syncthreads();
  for (int tk = threadIdx.x; tk < p.Tk; tk += blockDim.x) {
    const size_t krow = ((size_t)b * p.Tk + tk) * q.ld_kv + p.k_col0 + h * q.dh;
    float acc = 0.f;
    for (int c = 0; c < q.dh; ++c) acc = fmaf(qv[c], ld_split(q.kv_hi, q.kv_lo, krow + c, q.f16), acc);
    acc = acc * q.scale;
    // reference: logits += mask * -1e9 with mask = max(padding, look-ahead) in {0,1} (models.py:136-138, layers.py:186-187)
    if (tk >= len || (p.causal && tk > tq)) acc += -1e9f;
    logit[tk] = acc;
  }
  __syncthreads();
  float mx = -INFINITY;
  for (int tk = threadIdx.x; tk < p.Tk; tk += blockDim.x) mx = fmaxf(mx, logit[tk]);
  for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = red[0];
  for (int w = 1; w < (int)(blockDim.x >> 5); ++w) mx = fmaxf(mx, red[w]);
  __syncthreads();
  float sum = 0.f;
  for (int tk = threadIdx.x; tk < p.Tk; tk += blockDim.x) {
    const float e = expf(logit[tk] - mx);
    logit[tk] = e;
    sum += e;
  }
  for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
  __syncthreads();
  sum = 0.f;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) sum += red[w];
  const float inv = 1.f / sum;
  if (q.weights && (q.weights_all || b == q.weights_b)) {
    float* wrow = q.weights + (((size_t)(q.weights_all ? b : 0) * p.H + h) * p.T + tq) * p.Tk;
    for (int tk = threadIdx.x; tk < p.Tk; tk += blockDim.x) wrow[tk] = logit[tk] * inv;
  }
  if (q.weights_only) return;
  for (int c = threadIdx.x; c < q.dh; c += blockDim.x) {
    const size_t vcol = (size_t)p.v_col0 + h * q.dh + c;
    float acc = 0.f;
    for (int tk = 0; tk < p.Tk; ++tk)
      acc = fmaf(logit[tk], ld_split(q.kv_hi, q.kv_lo, ((size_t)b * p.Tk + tk) * q.ld_kv + vcol, q.f16), acc);
    acc *= inv;
    __nv_bfloat16 hi, lo;
    split_bf16(acc, hi, lo);
    const size_t o = ((size_t)b * p.T + tq) * p.ld_out + h * q.dh + c;
    p.out_hi[o] = hi;
    if (p.out_lo) p.out_lo[o] = lo;
  }
}

// ----------------------------------------------------------------------------------------------------
// Attention weights as a model output (reference-exact fp32: logits + mask * -1e9, softmax; layers.py:176-195).
// Block = 32 queries of one (batch row, head); key tiles of 32 rows are staged through shared memory as fp32
// (hi + lo recombined), lane = key, each warp owns 4 queries.  Three light passes over the (L2-resident) output row:
// logits + running max, exp + sum, scale.  Row stride DH + 4 floats keeps the 128-bit shared loads conflict-free.
// ----------------------------------------------------------------------------------------------------
constexpr int AW_Q = 32;   // queries per block
constexpr int AW_K = 32;   // keys per staged tile

__global__ void __launch_bounds__(256) mha_weights_kernel(const MhaKParams p, const MhaSimtPtrs q) {
  extern __shared__ float aw_sm[];
  const int dh = q.dh;
  const int kst = dh + 4;
  float* Qs = aw_sm;                 // [AW_Q][dh]
  float* Ks = aw_sm + AW_Q * dh;     // [AW_K][dh + 4]
  const int q0 = blockIdx.x * AW_Q, h = blockIdx.y;
  const int b = q.weights_all ? blockIdx.z : q.weights_b;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int len = min(max(p.kv_len[b], 0), p.Tk);
  for (int i = threadIdx.x; i < AW_Q * dh; i += blockDim.x) {
    const int r = i / dh, c = i - r * dh;
    const int tq = q0 + r;
    Qs[i] = tq < p.T ? ld_split(q.qk_hi, q.qk_lo, ((size_t)b * p.T + tq) * q.ld_qk + p.q_col0 + h * dh + c, q.f16) : 0.f;
  }
  float* wbase = q.weights + ((size_t)(q.weights_all ? b : 0) * p.H + h) * (size_t)p.T * p.Tk;
  float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
  const int n_tiles = (p.Tk + AW_K - 1) / AW_K;
  for (int kt = 0; kt < n_tiles; ++kt) {
    __syncthreads();  // Qs ready (first tile) / previous tile consumed
    for (int i = threadIdx.x; i < AW_K * dh; i += blockDim.x) {
      const int r = i / dh, c = i - r * dh;
      const int tk = kt * AW_K + r;
      Ks[r * kst + c] = tk < p.Tk ? ld_split(q.kv_hi, q.kv_lo, ((size_t)b * p.Tk + tk) * q.ld_kv + p.k_col0 + h * dh + c, q.f16) : 0.f;
    }
    __syncthreads();
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const float4* krow = reinterpret_cast<const float4*>(Ks + lane * kst);
#pragma unroll 2
    for (int c4 = 0; c4 < dh / 4; ++c4) {
      const float4 kv = krow[c4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 qv = reinterpret_cast<const float4*>(Qs + (warp * 4 + i) * dh)[c4];
        acc[i] = fmaf(qv.x, kv.x, acc[i]);
        acc[i] = fmaf(qv.y, kv.y, acc[i]);
        acc[i] = fmaf(qv.z, kv.z, acc[i]);
        acc[i] = fmaf(qv.w, kv.w, acc[i]);
      }
    }
    const int tk = kt * AW_K + lane;
    if (tk < p.Tk) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int tq = q0 + warp * 4 + i;
        if (tq < p.T) {
          float l = acc[i] * q.scale;
          if (tk >= len || (p.causal && tk > tq)) l += -1e9f;  // mask = max(padding, look-ahead) in {0,1}, times -1e9
          wbase[(size_t)tq * p.Tk + tk] = l;
          mx[i] = fmaxf(mx[i], l);
        }
      }
    }
  }
  // every lane re-reads only what it wrote itself: no fence needed
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int tq = q0 + warp * 4 + i;
    if (tq >= p.T) continue;   // warp-uniform
    float m = mx[i];
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float* wrow = wbase + (size_t)tq * p.Tk;
    float sum = 0.f;
    for (int tk = lane; tk < p.Tk; tk += 32) {
      const float e = expf(wrow[tk] - m);
      wrow[tk] = e;
      sum += e;
    }
    for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float inv = 1.f / sum;
    for (int tk = lane; tk < p.Tk; tk += 32) wrow[tk] *= inv;
  }
}

template <int DH, bool kSplit, bool kF16, bool kWide>
static int launch_tc(const ttsb_mha_args* a, const MhaKParams& p, cudaStream_t stream) {
  using Cfg = MhaCfg<DH, kSplit>;
  CUtensorMap tmQ[2], tmK[2];
  const bool cross = a->kv_hi != nullptr;
  const int Tk = cross ? a->Tk : a->T;
  const int ld_kv = cross ? a->ld_kv : a->ld_qk;
  for (int hl = 0; hl < 2; ++hl) {
    const void* qk = hl == 0 ? a->qk_hi : (kSplit ? a->qk_lo : a->qk_hi);
    const void* kv = cross ? (hl == 0 ? a->kv_hi : (kSplit ? a->kv_lo : a->kv_hi)) : qk;
    int rc = make_tmap_bf16_3d(&tmQ[hl], qk, (uint64_t)a->ld_qk, a->T, a->B, a->ld_qk, (uint64_t)a->ld_qk * a->T, 64, ATT_BQ);
    if (rc) return rc;
    rc = make_tmap_bf16_3d(&tmK[hl], kv, (uint64_t)ld_kv, Tk, a->B, ld_kv, (uint64_t)ld_kv * Tk, 64, ATT_BKV);
    if (rc) return rc;
  }
  static PerDevice<bool> attr_set;
  if (!attr_set.get()) {
    TTSB_CUDA_OK(cudaFuncSetAttribute(mha_tc_kernel<DH, kSplit, kF16, kWide>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set.get() = true;
  }
  dim3 grid((a->T + ATT_BQ - 1) / ATT_BQ, a->H, a->B);
  mha_tc_kernel<DH, kSplit, kF16, kWide><<<grid, kWide ? ATT_THREADS_WIDE : ATT_THREADS_NARROW, Cfg::kSmemBytes, stream>>>(tmQ[0], tmQ[1], tmK[0], tmK[1], p);
  count_launch();
  return check_cuda(cudaGetLastError(), "mha_tc_kernel launch");
}

}  // namespace ttsb

using namespace ttsb;

extern "C" int ttsb_mha_fwd(const ttsb_mha_args* a, void* stream_v) {
  if (!a) { set_last_error("ttsb_mha_fwd: args is NULL"); return TTSB_ERR_INVALID_ARGUMENT; }
  if (a->B <= 0 || a->T <= 0 || a->H <= 0) { set_last_error("ttsb_mha_fwd: B,T,H must be positive"); return TTSB_ERR_INVALID_ARGUMENT; }
  if (!a->qk_hi || !a->kv_len || !a->out_hi) { set_last_error("ttsb_mha_fwd: NULL tensor"); return TTSB_ERR_INVALID_ARGUMENT; }
  const bool split = a->precision == TTSB_PREC_BF16X3;
  const bool f16 = a->precision == TTSB_PREC_FP16;
  if (split && (!a->qk_lo || !a->out_lo)) { set_last_error("ttsb_mha_fwd: bf16x3 needs the lo planes"); return TTSB_ERR_INVALID_ARGUMENT; }
  const bool cross = a->kv_hi != nullptr;
  if (cross && (a->Tk <= 0 || a->ld_kv % 8 || (split && !a->kv_lo))) {
    set_last_error("ttsb_mha_fwd: cross-attention needs Tk > 0, ld_kv %% 8 == 0 and (bf16x3) the kv lo plane");
    return TTSB_ERR_INVALID_ARGUMENT;
  }
  if (cross && !a->full_queries) { set_last_error("ttsb_mha_fwd: cross-attention computes every query row (set full_queries = 1)"); return TTSB_ERR_INVALID_ARGUMENT; }
  if (a->causal && cross) { set_last_error("ttsb_mha_fwd: the look-ahead mask applies to self-attention only"); return TTSB_ERR_INVALID_ARGUMENT; }
  if (a->ld_qk % 8 || a->ld_out % 16 || a->q_col0 % 8 || a->k_col0 % 8 || a->v_col0 % 8) {
    set_last_error("ttsb_mha_fwd: leading dimensions / column offsets must be multiples of 8");
    return TTSB_ERR_INVALID_ARGUMENT;
  }
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  MhaKParams p{};
  p.B = a->B; p.T = a->T; p.H = a->H;
  p.Tk = cross ? a->Tk : a->T;
  p.causal = a->causal ? 1 : 0;
  p.full_queries = a->full_queries ? 1 : 0;
  p.q_col0 = a->q_col0; p.k_col0 = a->k_col0; p.v_col0 = a->v_col0; p.kv_len = a->kv_len;
  p.out_hi = static_cast<__nv_bfloat16*>(a->out_hi);
  p.out_lo = static_cast<__nv_bfloat16*>(a->out_lo);  // optional second plane of the OUTPUT (consumer may be bf16x3)
  p.ld_out = a->ld_out;
  p.scale_log2 = 1.4426950408889634f / sqrtf((float)a->dh);
#ifdef TTSB_ATT_TRACE
  {
    const char* e = getenv("TTSB_ATT_TRACE_PTR");  // device pointer (hex) of a 2*64*8 int64 buffer
    p.trace = e ? reinterpret_cast<long long*>(strtoull(e, nullptr, 16)) : nullptr;
  }
#endif

  MhaSimtPtrs q{};
  q.qk_hi = static_cast<const __nv_bfloat16*>(a->qk_hi);
  q.qk_lo = split ? static_cast<const __nv_bfloat16*>(a->qk_lo) : nullptr;
  q.ld_qk = a->ld_qk;
  q.kv_hi = cross ? static_cast<const __nv_bfloat16*>(a->kv_hi) : q.qk_hi;
  q.kv_lo = cross ? (split ? static_cast<const __nv_bfloat16*>(a->kv_lo) : nullptr) : q.qk_lo;
  q.ld_kv = cross ? a->ld_kv : a->ld_qk;
  q.dh = a->dh;
  q.scale = 1.f / sqrtf((float)a->dh);
  q.f16 = f16 ? 1 : 0;
  q.weights = a->weights_out;
  q.weights_b = a->weights_batch_index;
  q.weights_all = a->weights_all ? 1 : 0;
  const size_t simt_smem = (size_t)(p.Tk + a->dh) * sizeof(float);
  const bool need_simt = a->impl == TTSB_IMPL_SIMT;  // the weights-only pass has its own tiled kernel (any Tk)
  if (need_simt && simt_smem > 200 * 1024) { set_last_error("ttsb_mha_fwd: Tk too large for the weights / SIMT kernel"); return TTSB_ERR_UNSUPPORTED; }
  if (need_simt && simt_smem > 48 * 1024) {
    static PerDevice<size_t> simt_attr_pd;
    size_t& simt_attr = simt_attr_pd.get();
    if (simt_smem > simt_attr) {
      TTSB_CUDA_OK(cudaFuncSetAttribute(mha_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)simt_smem));
      simt_attr = simt_smem;
    }
  }

  if (a->impl == TTSB_IMPL_SIMT) {
    q.weights_only = 0;
    mha_simt_kernel<<<dim3(a->T, a->H, a->B), 128, simt_smem, stream>>>(p, q);
    count_launch();
    return check_cuda(cudaGetLastError(), "mha_simt_kernel launch");
  }
  // wide (8 softmax warps) is the default layout; TTSB_ATT_NARROW=1 selects the 4-warp layout for A/B measurements
  static const bool narrow = [] { const char* e = getenv("TTSB_ATT_NARROW"); return e && e[0] == '1'; }();
  int rc;
#define TTSB_MHA_DISPATCH(DH_, W_)                                                                                  \
  (split ? launch_tc<DH_, true, false, W_>(a, p, stream)                                                            \
         : (f16 ? launch_tc<DH_, false, true, W_>(a, p, stream) : launch_tc<DH_, false, false, W_>(a, p, stream)))
#define TTSB_MHA_DISPATCH1(DH_, W_) \
  (f16 ? launch_tc<DH_, false, true, W_>(a, p, stream) : launch_tc<DH_, false, false, W_>(a, p, stream))
  if (a->dh == 128) rc = narrow ? TTSB_MHA_DISPATCH(128, false) : TTSB_MHA_DISPATCH(128, true);
  else if (a->dh == 64) rc = narrow ? TTSB_MHA_DISPATCH(64, false) : TTSB_MHA_DISPATCH(64, true);
  else if (a->dh == 192 && !split) rc = narrow ? TTSB_MHA_DISPATCH1(192, false) : TTSB_MHA_DISPATCH1(192, true);
  else if (a->dh == 256 && !split) rc = narrow ? TTSB_MHA_DISPATCH1(256, false) : TTSB_MHA_DISPATCH1(256, true);
  else { set_last_error("ttsb_mha_fwd: head_dim %d not supported by the tcgen05 kernel (64, 128; 192, 256 in single-pass modes)", a->dh); return TTSB_ERR_UNSUPPORTED; }
  if (rc) return rc;
  if (a->weights_out) {
    if (!a->weights_all && (a->weights_batch_index < 0 || a->weights_batch_index >= a->B)) { set_last_error("ttsb_mha_fwd: weights_batch_index out of range"); return TTSB_ERR_INVALID_ARGUMENT; }
    q.weights_only = 1;
    if (a->dh % 4) { set_last_error("ttsb_mha_fwd: the weights kernel needs head_dim %% 4 == 0"); return TTSB_ERR_UNSUPPORTED; }
    const size_t aw_smem = (size_t)(AW_Q * a->dh + AW_K * (a->dh + 4)) * sizeof(float);
    static PerDevice<size_t> aw_attr_pd;
    size_t& aw_attr = aw_attr_pd.get();
    if (aw_smem > 48 * 1024 && aw_smem > aw_attr) {
      TTSB_CUDA_OK(cudaFuncSetAttribute(mha_weights_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)aw_smem));
      aw_attr = aw_smem;
    }
    mha_weights_kernel<<<dim3((a->T + AW_Q - 1) / AW_Q, a->H, a->weights_all ? a->B : 1), 256, aw_smem, stream>>>(p, q);
    count_launch();
    return check_cuda(cudaGetLastError(), "mha weights kernel launch");
  }
  return 0;
}
