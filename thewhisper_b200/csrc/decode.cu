// Decoder-step kernels for sm_100a (q_len = 1): every one is a stream over weights or KV in HBM, so they are
// plain coalesced 16-byte-load kernels sized to cover all 148 SMs; no tensor cores (the step is
// HBM-bound, 1.60 GB of weights + 245.76 MB cross-KV per audio per token).
//
// Replaces, per generated token, TF/models/whisper/modeling_whisper.py:449-506 (decoder layer), :738-763
// (embeddings), :1081 (tied LM head) and the logits processors + argmax of TF/generation/logits_process.py:1812-2043
// and TF/generation/utils.py:2762-2797, which the reference runs as ~30 launches per layer plus Python loops.
#include <math.h>

#include "decode.cuh"
#include "kernels.h"

namespace bw {

namespace {

constexpr int GEMV_THREADS = 256;
constexpr int GEMV_WARPS = GEMV_THREADS / 32;

// ------------------------------------------------------------------------------------------------
// GEMV with optional fused LayerNorm prologue and bias / scale / GELU / residual / KV-scatter epilogue
// ------------------------------------------------------------------------------------------------
template <int MB>
__global__ void __launch_bounds__(GEMV_THREADS) gemv_kernel(const GemvArgs a, const int rows_per_warp) {
  extern __shared__ float xs[];  // [MB][K]
  __shared__ float red[GEMV_WARPS][MB];
  __shared__ float stat[2][MB];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int K = a.K;

  for (int i = threadIdx.x; i < MB * K; i += GEMV_THREADS) {
    const int m = i / K, k = i - m * K;
    xs[i] = (m < a.M) ? a.x[(long long)m * a.ldx + k] : 0.f;
  }
  __syncthreads();
  if (a.ln_g) {
    // two-pass LayerNorm (biased variance, eps 1e-5) like torch.nn.LayerNorm
    float part[MB];
#pragma unroll
    for (int m = 0; m < MB; ++m) part[m] = 0.f;
    for (int k = threadIdx.x; k < K; k += GEMV_THREADS) {
#pragma unroll
      for (int m = 0; m < MB; ++m) part[m] += xs[m * K + k];
    }
#pragma unroll
    for (int m = 0; m < MB; ++m) {
      const float s = warp_sum(part[m]);
      if (lane == 0) red[warp][m] = s;
    }
    __syncthreads();
    if (threadIdx.x < MB) {
      float s = 0.f;
      for (int w = 0; w < GEMV_WARPS; ++w) s += red[w][threadIdx.x];
      stat[0][threadIdx.x] = s / (float)K;
    }
    __syncthreads();
#pragma unroll
    for (int m = 0; m < MB; ++m) part[m] = 0.f;
    for (int k = threadIdx.x; k < K; k += GEMV_THREADS) {
#pragma unroll
      for (int m = 0; m < MB; ++m) {
        const float d = xs[m * K + k] - stat[0][m];
        part[m] = fmaf(d, d, part[m]);
      }
    }
#pragma unroll
    for (int m = 0; m < MB; ++m) {
      const float s = warp_sum(part[m]);
      if (lane == 0) red[warp][m] = s;
    }
    __syncthreads();
    if (threadIdx.x < MB) {
      float s = 0.f;
      for (int w = 0; w < GEMV_WARPS; ++w) s += red[w][threadIdx.x];
      stat[1][threadIdx.x] = rsqrtf(s / (float)K + 1e-5f);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < K; k += GEMV_THREADS) {
      const float g = a.ln_g[k], bb = a.ln_b[k];
#pragma unroll
      for (int m = 0; m < MB; ++m) xs[m * K + k] = (xs[m * K + k] - stat[0][m]) * stat[1][m] * g + bb;
    }
    __syncthreads();
  }

  const int gw = blockIdx.x * GEMV_WARPS + warp;
  const int n_begin = gw * rows_per_warp;
  const int n_end = min(a.N, n_begin + rows_per_warp);
  const int pos = a.pos ? *a.pos : 0;
  for (int n = n_begin; n < n_end; n += 2) {
    const bool two = (n + 1) < n_end;
    const bf16* w0 = a.W + (long long)n * K;
    const bf16* w1 = a.W + (long long)(two ? n + 1 : n) * K;
    float acc0[MB], acc1[MB];
#pragma unroll
    for (int m = 0; m < MB; ++m) acc0[m] = acc1[m] = 0.f;
#pragma unroll 4
    for (int k = lane * 8; k < K; k += 256) {
      const uint4 u0 = ld_nc_u4(w0 + k);
      const uint4 u1 = ld_nc_u4(w1 + k);
      float f0[8], f1[8];
      float2 t;
      t = unpack_bf16(u0.x); f0[0] = t.x; f0[1] = t.y;
      t = unpack_bf16(u0.y); f0[2] = t.x; f0[3] = t.y;
      t = unpack_bf16(u0.z); f0[4] = t.x; f0[5] = t.y;
      t = unpack_bf16(u0.w); f0[6] = t.x; f0[7] = t.y;
      t = unpack_bf16(u1.x); f1[0] = t.x; f1[1] = t.y;
      t = unpack_bf16(u1.y); f1[2] = t.x; f1[3] = t.y;
      t = unpack_bf16(u1.z); f1[4] = t.x; f1[5] = t.y;
      t = unpack_bf16(u1.w); f1[6] = t.x; f1[7] = t.y;
#pragma unroll
      for (int m = 0; m < MB; ++m) {
        const float4 xa = *reinterpret_cast<const float4*>(&xs[m * K + k]);
        const float4 xb = *reinterpret_cast<const float4*>(&xs[m * K + k + 4]);
        acc0[m] = fmaf(f0[0], xa.x, acc0[m]); acc0[m] = fmaf(f0[1], xa.y, acc0[m]);
        acc0[m] = fmaf(f0[2], xa.z, acc0[m]); acc0[m] = fmaf(f0[3], xa.w, acc0[m]);
        acc0[m] = fmaf(f0[4], xb.x, acc0[m]); acc0[m] = fmaf(f0[5], xb.y, acc0[m]);
        acc0[m] = fmaf(f0[6], xb.z, acc0[m]); acc0[m] = fmaf(f0[7], xb.w, acc0[m]);
        acc1[m] = fmaf(f1[0], xa.x, acc1[m]); acc1[m] = fmaf(f1[1], xa.y, acc1[m]);
        acc1[m] = fmaf(f1[2], xa.z, acc1[m]); acc1[m] = fmaf(f1[3], xa.w, acc1[m]);
        acc1[m] = fmaf(f1[4], xb.x, acc1[m]); acc1[m] = fmaf(f1[5], xb.y, acc1[m]);
        acc1[m] = fmaf(f1[6], xb.z, acc1[m]); acc1[m] = fmaf(f1[7], xb.w, acc1[m]);
      }
    }
#pragma unroll
    for (int m = 0; m < MB; ++m) {
      acc0[m] = warp_sum(acc0[m]);
      acc1[m] = warp_sum(acc1[m]);
    }
    // lanes [0, MB) finish row n, lanes [16, 16+MB) finish row n+1
    const int m = lane & 15;
    const int nn = n + (lane >> 4);
    if (m < MB && m < a.M && (lane < 16 || two)) {
      float v = 0.f;
#pragma unroll
      for (int mm = 0; mm < MB; ++mm)
        if (mm == m) v = (lane < 16) ? acc0[mm] : acc1[mm];
      if (a.bias) v += a.bias[nn];
      if (nn < a.alpha_cols) v *= a.alpha;
      if (a.act == 1) v = gelu_erf(v);
      if (a.residual) v += a.residual[(long long)m * a.ldo + nn];
      a.out[(long long)m * a.ldo + nn] = v;
      if (a.kc && nn >= a.D) {
        const long long row = ((long long)(a.seq0 + m) * a.Tmax + pos) * a.D;
        if (nn < 2 * a.D) a.kc[row + nn - a.D] = __float2bfloat16(v);
        else a.vc[row + nn - 2 * a.D] = __float2bfloat16(v);
      }
    }
  }
}

// x[q, :] = E[token[q, pos], :] + P[pos, :]
__global__ void embed_kernel(const bf16* __restrict__ E, const float* __restrict__ P, const int* __restrict__ tokens,
                             const int* __restrict__ pos_ptr, float* __restrict__ x, int D, int Tmax) {
  const int q = blockIdx.x;
  const int pos = *pos_ptr;
  const int tok = tokens[q * Tmax + pos];
  for (int d = threadIdx.x; d < D; d += blockDim.x)
    x[(long long)q * D + d] = __bfloat162float(E[(long long)tok * D + d]) + P[(long long)pos * D + d];
}

// ------------------------------------------------------------------------------------------------
// causal self-attention over the cached positions 0..pos of one (sequence, head); 8 lanes share a key row
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) self_attn_kernel(const SelfAttnArgs a) {
  extern __shared__ float sc[];  // [Tmax] scores, then [2][64] partials
  __shared__ float redm[4], reds[4];
  const int h = blockIdx.x, q = blockIdx.y;
  const int pos = *a.pos;
  const int n = pos + 1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane & 7;
  const float* qp = a.qkv + (long long)q * 3 * a.D + h * 64 + sub * 8;
  float qv[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) qv[i] = qp[i];
  float lmax = -INFINITY;
  for (int sb = warp * 4; sb < n; sb += 16) {  // warp-uniform trip count: the shuffles below need all 32 lanes
    const int s = sb + (lane >> 3);
    const bool valid = s < n;
    uint4 u = make_uint4(0u, 0u, 0u, 0u);
    if (valid) {
      const int slot = a.anc ? a.anc[q * a.Tmax + s] : q;
      u = ld_nc_u4(a.kc + ((long long)slot * a.Tmax + s) * a.D + h * 64 + sub * 8);
    }
    float2 t;
    float d = 0.f;
    t = unpack_bf16(u.x); d = fmaf(qv[0], t.x, d); d = fmaf(qv[1], t.y, d);
    t = unpack_bf16(u.y); d = fmaf(qv[2], t.x, d); d = fmaf(qv[3], t.y, d);
    t = unpack_bf16(u.z); d = fmaf(qv[4], t.x, d); d = fmaf(qv[5], t.y, d);
    t = unpack_bf16(u.w); d = fmaf(qv[6], t.x, d); d = fmaf(qv[7], t.y, d);
    d += __shfl_xor_sync(0xffffffffu, d, 1);
    d += __shfl_xor_sync(0xffffffffu, d, 2);
    d += __shfl_xor_sync(0xffffffffu, d, 4);
    if (valid) {
      if (sub == 0) sc[s] = d;
      lmax = fmaxf(lmax, d);
    }
  }
  lmax = warp_max(lmax);
  if (lane == 0) redm[warp] = lmax;
  __syncthreads();
  const float mx = fmaxf(fmaxf(redm[0], redm[1]), fmaxf(redm[2], redm[3]));
  float lsum = 0.f;
  for (int s = threadIdx.x; s < n; s += 128) {
    const float e = __expf(sc[s] - mx);
    sc[s] = e;
    lsum += e;
  }
  lsum = warp_sum(lsum);
  if (lane == 0) reds[warp] = lsum;
  __syncthreads();
  const float inv = 1.0f / (reds[0] + reds[1] + reds[2] + reds[3]);
  const int d = threadIdx.x & 63, part = threadIdx.x >> 6;
  float acc = 0.f;
  for (int s = part; s < n; s += 2) {
    const int slot = a.anc ? a.anc[q * a.Tmax + s] : q;
    acc = fmaf(sc[s], __bfloat162float(a.vc[((long long)slot * a.Tmax + s) * a.D + h * 64 + d]), acc);
  }
  float* po = sc + a.Tmax;
  po[part * 64 + d] = acc;
  __syncthreads();
  if (threadIdx.x < 64) a.out[(long long)q * a.D + h * 64 + d] = (po[d] + po[64 + d]) * inv;
}

// ------------------------------------------------------------------------------------------------
// cross-attention over the encoder K/V of one audio, shared by its G beams; XSPLIT key splits per (audio, head),
// the last-arriving split block merges the partials (flash-decoding) so no extra launch is needed
// ------------------------------------------------------------------------------------------------
constexpr int XKEYS = 256;  // max keys per split  (S <= XSPLIT * XKEYS = 2048)

__global__ void __launch_bounds__(128) cross_attn_kernel(const CrossAttnArgs a) {
  __shared__ float sc[MAXG][XKEYS];
  __shared__ float redm[4], reds[4];
  __shared__ float po[2][MAXG][64];
  __shared__ float gm[MAXG], gl[MAXG];
  __shared__ unsigned is_last;
  const int split = blockIdx.x, h = blockIdx.y, au = blockIdx.z;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, sub = lane & 7;
  const int G = a.G;
  const int ks = (a.S + XSPLIT - 1) / XSPLIT;
  const int s0 = split * ks;
  const int n = max(0, min(a.S, s0 + ks) - s0);
  const bf16* kbase = a.kc + (((long long)au * a.H + h) * a.S + s0) * 64;
  const bf16* vbase = a.vc + (((long long)au * a.H + h) * a.S + s0) * 64;

  float qv[MAXG][8];
#pragma unroll
  for (int g = 0; g < MAXG; ++g) {
    if (g < G) {
      const float* qp = a.q + (long long)(au * G + g) * a.D + h * 64 + sub * 8;
#pragma unroll
      for (int i = 0; i < 8; ++i) qv[g][i] = qp[i];
    }
  }
  const int slot = a.head_slots ? a.head_slots[h] : -1;
  float* align_row = nullptr;
  if (slot >= 0 && a.align) {
    const int step = *a.pos - a.step_base;
    if (step >= 0 && step < a.Tcap) align_row = a.align + (((long long)au * a.Ha + slot) * a.Tcap + step) * a.S + s0;
  }
  for (int kb = warp * 4; kb < n; kb += 16) {  // warp-uniform trip count (shuffles)
    const int kk = kb + (lane >> 3);
    const bool valid = kk < n;
    uint4 u = make_uint4(0u, 0u, 0u, 0u);
    if (valid) u = ld_nc_u4(kbase + (long long)kk * 64 + sub * 8);
    float kf[8];
    float2 t;
    t = unpack_bf16(u.x); kf[0] = t.x; kf[1] = t.y;
    t = unpack_bf16(u.y); kf[2] = t.x; kf[3] = t.y;
    t = unpack_bf16(u.z); kf[4] = t.x; kf[5] = t.y;
    t = unpack_bf16(u.w); kf[6] = t.x; kf[7] = t.y;
#pragma unroll
    for (int g = 0; g < MAXG; ++g) {
      if (g < G) {
        float d = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) d = fmaf(qv[g][i], kf[i], d);
        d += __shfl_xor_sync(0xffffffffu, d, 1);
        d += __shfl_xor_sync(0xffffffffu, d, 2);
        d += __shfl_xor_sync(0xffffffffu, d, 4);
        if (sub == 0 && valid) {
          sc[g][kk] = d;
          if (g == 0 && align_row) align_row[kk] = d;
        }
      }
    }
  }
  __syncthreads();
  for (int g = 0; g < G; ++g) {
    float lmax = -INFINITY;
    for (int kk = threadIdx.x; kk < n; kk += 128) lmax = fmaxf(lmax, sc[g][kk]);
    lmax = warp_max(lmax);
    if (lane == 0) redm[warp] = lmax;
    __syncthreads();
    const float mx = fmaxf(fmaxf(redm[0], redm[1]), fmaxf(redm[2], redm[3]));
    float lsum = 0.f;
    for (int kk = threadIdx.x; kk < n; kk += 128) {
      const float e = __expf(sc[g][kk] - mx);
      sc[g][kk] = e;
      lsum += e;
    }
    lsum = warp_sum(lsum);
    if (lane == 0) reds[warp] = lsum;
    __syncthreads();
    if (threadIdx.x == 0) {
      gm[g] = mx;
      gl[g] = reds[0] + reds[1] + reds[2] + reds[3];
    }
    __syncthreads();
  }
  const int d = threadIdx.x & 63, part = threadIdx.x >> 6;
  float acc[MAXG];
#pragma unroll
  for (int g = 0; g < MAXG; ++g) acc[g] = 0.f;
#pragma unroll 4
  for (int kk = part; kk < n; kk += 2) {
    const float v = __bfloat162float(vbase[(long long)kk * 64 + d]);
#pragma unroll
    for (int g = 0; g < MAXG; ++g)
      if (g < G) acc[g] = fmaf(sc[g][kk], v, acc[g]);
  }
#pragma unroll
  for (int g = 0; g < MAXG; ++g)
    if (g < G) po[part][g][d] = acc[g];
  __syncthreads();
  const long long pbase = (((long long)au * a.H + h) * XSPLIT + split) * G;
  if (threadIdx.x < 64) {
    for (int g = 0; g < G; ++g) a.part_o[(pbase + g) * 64 + d] = po[0][g][d] + po[1][g][d];
  }
  if (threadIdx.x < G) {
    a.part_ml[(pbase + threadIdx.x) * 2 + 0] = gm[threadIdx.x];
    a.part_ml[(pbase + threadIdx.x) * 2 + 1] = gl[threadIdx.x];
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned prev = atomicAdd(&a.counters[au * a.H + h], 1u);
    is_last = (prev == XSPLIT - 1) ? 1u : 0u;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  if (threadIdx.x < 64) {
    const long long hb = ((long long)au * a.H + h) * XSPLIT * G;
    for (int g = 0; g < G; ++g) {
      float M = -INFINITY;
      for (int sp = 0; sp < XSPLIT; ++sp) {
        const float l = __ldcg(&a.part_ml[(hb + (long long)sp * G + g) * 2 + 1]);
        if (l > 0.f) M = fmaxf(M, __ldcg(&a.part_ml[(hb + (long long)sp * G + g) * 2 + 0]));
      }
      float L = 0.f, o = 0.f;
      for (int sp = 0; sp < XSPLIT; ++sp) {
        const float l = __ldcg(&a.part_ml[(hb + (long long)sp * G + g) * 2 + 1]);
        if (l > 0.f) {
          const float w = __expf(__ldcg(&a.part_ml[(hb + (long long)sp * G + g) * 2 + 0]) - M);
          L = fmaf(l, w, L);
          o = fmaf(__ldcg(&a.part_o[(hb + (long long)sp * G + g) * 64 + d]), w, o);
        }
      }
      a.out[(long long)(au * G + g) * a.D + h * 64 + d] = o / L;
    }
  }
  if (threadIdx.x == 0) a.counters[au * a.H + h] = 0u;
}

// ------------------------------------------------------------------------------------------------
// token selection: suppress masks + Whisper timestamp rules + greedy argmax, one block per sequence
// (SuppressTokens -> SuppressTokensAtBegin -> WhisperTimeStamp, the order observed in SURVEY.md §3.3)
// ------------------------------------------------------------------------------------------------
struct MaxIdx {
  float v;
  int i;
};
__device__ __forceinline__ MaxIdx better(MaxIdx a, MaxIdx b) {  // larger value, ties -> smaller index (torch.argmax)
  if (b.v > a.v || (b.v == a.v && b.i < a.i)) return b;
  return a;
}
__device__ __forceinline__ MaxIdx warp_best(MaxIdx m) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    MaxIdx t;
    t.v = __shfl_xor_sync(0xffffffffu, m.v, o);
    t.i = __shfl_xor_sync(0xffffffffu, m.i, o);
    m = better(m, t);
  }
  return m;
}

constexpr int SEL_THREADS = 1024;

__global__ void __launch_bounds__(SEL_THREADS) select_kernel(const SelectArgs a) {
  __shared__ MaxIdx s_text[32], s_ts[32];
  __shared__ float s_sum[32];
  __shared__ float s_f[4];
  __shared__ int s_i[8];
  const int q = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int pos = *a.pos;          // index of the token just consumed
  const int cur_len = pos + 1;     // tokens present; the new token goes to index cur_len
  const bool generating = cur_len >= a.begin_index && cur_len < a.Tmax;
  if (generating) {
    const float* lg = a.logits + (long long)q * a.V;
    const int* seq = a.tokens + q * a.Tmax;
    // ---- mask ranges from the token history (WhisperTimeStampLogitsProcessor, logits_process.py:1995-2033)
    if (threadIdx.x == 0) {
      int block_ts_all = 0, block_text_below_eos = 0, ts_lo_block_end = a.ts_begin, first_step = 0;
      if (a.ts_rules) {
        const int ngen = cur_len - a.begin_index;
        const bool last_ts = ngen >= 1 && seq[cur_len - 1] >= a.ts_begin;
        const bool penult_ts = ngen < 2 || seq[cur_len - 2] >= a.ts_begin;
        if (last_ts) {
          if (penult_ts) block_ts_all = 1;
          else block_text_below_eos = 1;
        }
        int last_tsv = -1;
        for (int i = cur_len - 1; i >= a.begin_index; --i)
          if (seq[i] >= a.ts_begin) { last_tsv = seq[i]; break; }
        if (last_tsv >= 0) ts_lo_block_end = (last_ts && !penult_ts) ? last_tsv : last_tsv + 1;
        first_step = (cur_len == a.begin_index);
      }
      s_i[0] = block_ts_all; s_i[1] = block_text_below_eos; s_i[2] = ts_lo_block_end; s_i[3] = first_step;
    }
    __syncthreads();
    const int block_ts_all = s_i[0], block_text_below_eos = s_i[1], ts_lo_end = s_i[2], first_step = s_i[3];
    const bool at_begin = (cur_len == a.begin_index);
    const int last_allowed = (a.max_initial_ts >= 0) ? a.ts_begin + a.max_initial_ts : a.V;
    auto masked = [&](int v) -> bool {
      if (a.suppress_bits[v >> 5] >> (v & 31) & 1u) return true;
      if (at_begin && a.begin_suppress_bits && (a.begin_suppress_bits[v >> 5] >> (v & 31) & 1u)) return true;
      if (a.ts_rules) {
        if (v == a.no_ts) return true;
        if (v >= a.ts_begin) {
          if (block_ts_all) return true;
          if (v < ts_lo_end) return true;
          if (first_step && v > last_allowed) return true;
        } else {
          if (block_text_below_eos && v < a.eos) return true;
          if (first_step) return true;
        }
      }
      return false;
    };
    // ---- pass 1: masked max / argmax of the text range and of the timestamp range; raw max for the lse
    MaxIdx bt{-INFINITY, 0x7fffffff}, bs{-INFINITY, 0x7fffffff};
    float rawmax = -INFINITY;
    const int tsb = a.ts_rules ? a.ts_begin : a.V;
    for (int v = threadIdx.x; v < a.V; v += SEL_THREADS) {
      const float x = lg[v];
      rawmax = fmaxf(rawmax, x);
      if (!masked(v)) {
        MaxIdx c{x, v};
        if (v < tsb) bt = better(bt, c);
        else bs = better(bs, c);
      }
    }
    bt = warp_best(bt);
    bs = warp_best(bs);
    rawmax = warp_max(rawmax);
    if (lane == 0) { s_text[warp] = bt; s_ts[warp] = bs; s_sum[warp] = rawmax; }
    __syncthreads();
    if (warp == 0) {
      MaxIdx t = s_text[lane], s = s_ts[lane];
      float r = s_sum[lane];
      t = warp_best(t); s = warp_best(s); r = warp_max(r);
      if (lane == 0) { s_text[0] = t; s_ts[0] = s; s_f[0] = r; }
    }
    __syncthreads();
    bt = s_text[0];
    bs = s_ts[0];
    rawmax = s_f[0];
    __syncthreads();
    // ---- pass 2: sum exp over unmasked timestamps (relative to their max) and over all raw logits
    float ts_sum = 0.f, raw_sum = 0.f;
    for (int v = threadIdx.x; v < a.V; v += SEL_THREADS) {
      const float x = lg[v];
      raw_sum += __expf(x - rawmax);
      if (v >= tsb && !masked(v)) ts_sum += __expf(x - bs.v);
    }
    ts_sum = warp_sum(ts_sum);
    raw_sum = warp_sum(raw_sum);
    if (lane == 0) { s_sum[warp] = ts_sum; s_text[warp].v = raw_sum; }
    __syncthreads();
    if (threadIdx.x == 0) {
      float tsum = 0.f, rsum = 0.f;
      for (int w = 0; w < SEL_THREADS / 32; ++w) { tsum += s_sum[w]; rsum += s_text[w].v; }
      if (a.out_lse) a.out_lse[q] = rawmax + logf(rsum);
      int choice;
      if (a.ts_rules) {
        // "sum of timestamp probability > max text probability -> sample a timestamp" (logits_process.py:2036-2041);
        // both sides share the log-softmax normaliser, so compare logsumexp(ts logits) with max(text logits)
        const float ts_lse = (bs.v == -INFINITY) ? -INFINITY : bs.v + logf(tsum);
        const bool force_ts = ts_lse > bt.v;
        if (force_ts) choice = bs.i;
        else choice = better(bt, bs).i;
      } else {
        choice = bt.i;
      }
      if (a.finished[q]) choice = a.pad;
      else if (choice == a.eos) a.finished[q] = 1;
      a.tokens[q * a.Tmax + cur_len] = choice;
    }
  }
  // ---- last block advances the shared position counter
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned prev = atomicAdd(a.done_ctr, 1u);
    if (prev == (unsigned)(a.Q - 1)) {
      *a.done_ctr = 0u;
      *a.pos = pos + 1;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm over rows (encoder): fp32 in, bf16 (GEMM operand) or fp32 out.  One warp per row.
// ------------------------------------------------------------------------------------------------
template <typename OutT>
__global__ void layernorm_rows_kernel(const float* __restrict__ x, const float* __restrict__ g, const float* __restrict__ b,
                                      OutT* __restrict__ y, int rows, int D) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + (long long)row * D;
  float s = 0.f;
  for (int k = lane * 4; k < D; k += 128) {
    const float4 v = *reinterpret_cast<const float4*>(xr + k);
    s += v.x + v.y + v.z + v.w;
  }
  const float mean = warp_sum(s) / (float)D;
  float ss = 0.f;
  for (int k = lane * 4; k < D; k += 128) {
    const float4 v = *reinterpret_cast<const float4*>(xr + k);
    const float a0 = v.x - mean, a1 = v.y - mean, a2 = v.z - mean, a3 = v.w - mean;
    ss += a0 * a0 + a1 * a1 + a2 * a2 + a3 * a3;
  }
  const float rstd = rsqrtf(warp_sum(ss) / (float)D + 1e-5f);
  for (int k = lane * 4; k < D; k += 128) {
    const float4 v = *reinterpret_cast<const float4*>(xr + k);
    const float4 gg = *reinterpret_cast<const float4*>(g + k);
    const float4 bb = *reinterpret_cast<const float4*>(b + k);
    const float o0 = (v.x - mean) * rstd * gg.x + bb.x, o1 = (v.y - mean) * rstd * gg.y + bb.y;
    const float o2 = (v.z - mean) * rstd * gg.z + bb.z, o3 = (v.w - mean) * rstd * gg.w + bb.w;
    if constexpr (sizeof(OutT) == 2) {
      uint2 w;
      w.x = pack_bf16(o0, o1);
      w.y = pack_bf16(o2, o3);
      *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(y) + (long long)row * D + k) = w;
    } else {
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(y) + (long long)row * D + k) = make_float4(o0, o1, o2, o3);
    }
  }
}

}  // namespace

int launch_gemv(cudaStream_t st, const GemvArgs& a) {
  BW_CHECK(a.M >= 1 && a.M <= 8, "gemv: M=%d must be in 1..8", a.M);
  BW_CHECK(a.K % 8 == 0, "gemv: K=%d must be a multiple of 8", a.K);
  const int mb = a.M <= 1 ? 1 : (a.M <= 2 ? 2 : (a.M <= 4 ? 4 : 8));
  // ~2 CTAs per SM on 148 SMs, at least 2 rows per warp so each lane keeps two 16-byte loads per k-step in flight
  int rpw = (a.N + 296 * GEMV_WARPS - 1) / (296 * GEMV_WARPS);
  if (rpw < 2) rpw = 2;
  if (rpw & 1) ++rpw;
  const int grid = (a.N + rpw * GEMV_WARPS - 1) / (rpw * GEMV_WARPS);
  const size_t smem = (size_t)mb * a.K * sizeof(float);
  BW_CHECK(smem <= 200 * 1024, "gemv: K=%d too large for the smem stage", a.K);
#define BW_GEMV_CASE(MB)                                                                                          \
  case MB: {                                                                                                      \
    static bool attr = false;                                                                                     \
    if (!attr) {                                                                                                  \
      BW_CUDA_OK(cudaFuncSetAttribute(gemv_kernel<MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); \
      attr = true;                                                                                                \
    }                                                                                                             \
    gemv_kernel<MB><<<grid, GEMV_THREADS, smem, st>>>(a, rpw);                                                    \
  } break;
  switch (mb) {
    BW_GEMV_CASE(1)
    BW_GEMV_CASE(2)
    BW_GEMV_CASE(4)
    BW_GEMV_CASE(8)
  }
#undef BW_GEMV_CASE
  BW_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_embed(cudaStream_t st, const bf16* E, const float* P, const int* tokens, const int* pos, float* x, int Q, int D, int Tmax) {
  embed_kernel<<<Q, 256, 0, st>>>(E, P, tokens, pos, x, D, Tmax);
  BW_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_self_attn(cudaStream_t st, const SelfAttnArgs& a, int Q) {
  const size_t smem = (size_t)(a.Tmax + 128) * sizeof(float);
  self_attn_kernel<<<dim3(a.H, Q), 128, smem, st>>>(a);
  BW_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_cross_attn(cudaStream_t st, const CrossAttnArgs& a, int A) {
  BW_CHECK(a.G >= 1 && a.G <= MAXG, "cross_attn: G=%d must be in 1..%d", a.G, MAXG);
  BW_CHECK(a.S <= XSPLIT * XKEYS, "cross_attn: S=%d exceeds %d", a.S, XSPLIT * XKEYS);
  cross_attn_kernel<<<dim3(XSPLIT, a.H, A), 128, 0, st>>>(a);
  BW_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_select(cudaStream_t st, const SelectArgs& a) {
  select_kernel<<<a.Q, SEL_THREADS, 0, st>>>(a);
  BW_CUDA_OK(cudaGetLastError());
  return 0;
}

int layernorm_bf16(cudaStream_t st, const float* x, const float* g, const float* b, bf16* y, int rows, int D) {
  BW_CHECK(D % 4 == 0, "layernorm: D=%d must be a multiple of 4", D);
  layernorm_rows_kernel<bf16><<<(rows + 7) / 8, 256, 0, st>>>(x, g, b, y, rows, D);
  BW_CUDA_OK(cudaGetLastError());
  return 0;
}

int layernorm_f32(cudaStream_t st, const float* x, const float* g, const float* b, float* y, int rows, int D) {
  BW_CHECK(D % 4 == 0, "layernorm: D=%d must be a multiple of 4", D);
  layernorm_rows_kernel<float><<<(rows + 7) / 8, 256, 0, st>>>(x, g, b, y, rows, D);
  BW_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace bw
