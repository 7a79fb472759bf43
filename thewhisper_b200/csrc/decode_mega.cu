// One persistent kernel per decoder step (q_len = 1, up to 8 sequences, one beam per audio).
//
// Why: with one kernel per op the step is 259 launches; on B200 each kernel boundary costs ~4 us of drain + launch and
// exposes one DRAM round trip, so a step that should take 0.32 ms (2.07 GB at the measured 6.5 TB/s) took 1.7 ms
// (profiles/r1_v1_launches_summary.md).  Here the whole step -- embedding, 32 x (LN1+QKV, self-attention, out-proj,
// LN2+cross-q, cross-attention, out-proj, LN3+fc1+GELU, fc2), final LN + tied LM head -- runs in ONE kernel of one CTA
// per SM.  Phases are separated by a grid barrier (~1 us, a global atomic counter), and every phase requests its first
// weight rows BEFORE waiting at the barrier that precedes it, so the DRAM latency of phase p+1 hides behind the tail of
// phase p and the barrier itself.  Token selection stays a separate small kernel (select_kernel).
//
// Work split: 16 warps per CTA, global warp id gw; a GEMV phase gives warp gw the rows {gw*R + i*GW*R + r}; attention
// phases hand (sequence, head[, key split]) items to CTAs round-robin.  Activations that cross CTAs (dx, dqkv, dattn, dq, dh,
// partials) are read with ld.global.cg (L2) because L1 is not coherent across SMs; weights use ld.global.nc.
#include <math.h>

#include "decode.cuh"
#include "kernels.h"

namespace bw {

namespace {

constexpr int MT = 384;        // threads per CTA (12 warps: <= 170 registers per thread, no spills)
constexpr int MW = MT / 32;    // warps per CTA
constexpr int KG = MT / 8;     // key groups of 8 lanes in the attention phases
constexpr int MAXKEYS = 448;   // self-attention keys held in smem (Tmax)
constexpr int XKMAX = 256;     // cross-attention keys per work item held in smem

__device__ __forceinline__ void unpack8m(const uint4& u, float (&f)[8]) {
  float2 t;
  t = unpack_bf16(u.x); f[0] = t.x; f[1] = t.y;
  t = unpack_bf16(u.y); f[2] = t.x; f[3] = t.y;
  t = unpack_bf16(u.z); f[4] = t.x; f[5] = t.y;
  t = unpack_bf16(u.w); f[6] = t.x; f[7] = t.y;
}
__device__ __forceinline__ void cp_async16m(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_allm() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// grid barrier: monotonically increasing arrival counter (zeroed by a memset node before the kernel)
struct GridBar {
  unsigned* ctr;
  unsigned nblocks;
  unsigned epoch;
  __device__ __forceinline__ void sync() {
    __syncthreads();
    if (threadIdx.x == 0) {
      ++epoch;
      __threadfence();
      atomicAdd(ctr, 1u);
      const unsigned target = epoch * nblocks;
      const long long t0 = clock64();
      while (ld_acquire_u32(ctr) < target) {
        if (clock64() - t0 > (1ll << 32)) {
          printf("[bw] decode_mega: grid barrier %u timed out (block %d)\n", epoch, blockIdx.x);
          __trap();
        }
      }
      __threadfence();
    }
    __syncthreads();
  }
};

template <int NC, int R>
struct WR {
  uint4 w[R][NC];
};

template <int NC, int R>
__device__ __forceinline__ void load_rows_m(WR<NC, R>& wr, const bf16* __restrict__ W, int K, int n, int N, int lane) {
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int row = min(n + r, N - 1);
    const bf16* wp = W + (long long)row * K;
#pragma unroll
    for (int i = 0; i < NC; ++i) {
      const int k = lane * 8 + i * 256;
      wr.w[r][i] = (k < K) ? ld_nc_u4(wp + k) : make_uint4(0u, 0u, 0u, 0u);
    }
  }
}

struct PhaseOut {
  const float* bias;
  float alpha;
  int alpha_cols;
  int act;
  float* residual;  // may alias out
  float* out;
  int ldo;
  bf16* kc;  // optional KV scatter (fused QKV)
  bf16* vc;
  int D, Tmax, pos;
};

// stage M rows of K floats into smem (ld.global.cg), optionally LayerNorm them.  All MT threads.
template <int MB>
__device__ void stage_x(float* xs, float* red, const float* __restrict__ src, int ld, int K, int M, const float* __restrict__ g,
                        const float* __restrict__ b) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < MB * K; i += MT) {
    const int m = i / K, k = i - m * K;
    xs[i] = (m < M) ? __ldcg(src + (long long)m * ld + k) : 0.f;
  }
  __syncthreads();
  if (!g) return;
  float part[MB];
#pragma unroll
  for (int m = 0; m < MB; ++m) part[m] = 0.f;
  for (int k = threadIdx.x; k < K; k += MT) {
#pragma unroll
    for (int m = 0; m < MB; ++m) part[m] += xs[m * K + k];
  }
#pragma unroll
  for (int m = 0; m < MB; ++m) {
    const float s = warp_sum(part[m]);
    if (lane == 0) red[warp * MB + m] = s;
  }
  __syncthreads();
  float mean[MB];
#pragma unroll
  for (int m = 0; m < MB; ++m) {
    float s = 0.f;
    for (int w = 0; w < MW; ++w) s += red[w * MB + m];
    mean[m] = s / (float)K;
  }
  __syncthreads();
#pragma unroll
  for (int m = 0; m < MB; ++m) part[m] = 0.f;
  for (int k = threadIdx.x; k < K; k += MT) {
#pragma unroll
    for (int m = 0; m < MB; ++m) {
      const float d = xs[m * K + k] - mean[m];
      part[m] = fmaf(d, d, part[m]);
    }
  }
#pragma unroll
  for (int m = 0; m < MB; ++m) {
    const float s = warp_sum(part[m]);
    if (lane == 0) red[warp * MB + m] = s;
  }
  __syncthreads();
  float rstd[MB];
#pragma unroll
  for (int m = 0; m < MB; ++m) {
    float s = 0.f;
    for (int w = 0; w < MW; ++w) s += red[w * MB + m];
    rstd[m] = rsqrtf(s / (float)K + 1e-5f);
  }
  for (int k = threadIdx.x; k < K; k += MT) {
    const float gg = g[k], bb = b[k];
#pragma unroll
    for (int m = 0; m < MB; ++m) xs[m * K + k] = (xs[m * K + k] - mean[m]) * rstd[m] * gg + bb;
  }
  __syncthreads();
}

// rows {gw*R + i*GW*R + r}; `cur` already holds the first pass (prefetched before the preceding barrier)
template <int MB, int NC, int R, bool PIPE>
__device__ void gemv_phase(WR<NC, R>& cur, const bf16* __restrict__ W, int N, int K, const float* xs, int M, const PhaseOut& o, int gw,
                           int GW, int lane) {
  for (int n = gw * R; n < N; n += GW * R) {
    WR<NC, R> nxt;  // (dead when !PIPE)
    const int n2 = n + GW * R;
    const bool has_next = n2 < N;
    if (PIPE && has_next) load_rows_m<NC, R>(nxt, W, K, n2, N, lane);
    float acc[R][MB];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int m = 0; m < MB; ++m) acc[r][m] = 0.f;
#pragma unroll
    for (int i = 0; i < NC; ++i) {
      const int k = lane * 8 + i * 256;
      if (k < K) {
        float wf[R][8];
#pragma unroll
        for (int r = 0; r < R; ++r) unpack8m(cur.w[r][i], wf[r]);
#pragma unroll
        for (int m = 0; m < MB; ++m) {
          const float4 xa = *reinterpret_cast<const float4*>(&xs[m * K + k]);
          const float4 xb = *reinterpret_cast<const float4*>(&xs[m * K + k + 4]);
#pragma unroll
          for (int r = 0; r < R; ++r) {
            float s = acc[r][m];
            s = fmaf(wf[r][0], xa.x, s); s = fmaf(wf[r][1], xa.y, s); s = fmaf(wf[r][2], xa.z, s); s = fmaf(wf[r][3], xa.w, s);
            s = fmaf(wf[r][4], xb.x, s); s = fmaf(wf[r][5], xb.y, s); s = fmaf(wf[r][6], xb.z, s); s = fmaf(wf[r][7], xb.w, s);
            acc[r][m] = s;
          }
        }
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int m = 0; m < MB; ++m) acc[r][m] = warp_sum(acc[r][m]);
    const int m = lane & 7, r_sel = lane >> 3;  // lanes [8r, 8r + MB) finish row n + r  (R <= 4, MB <= 8)
    const int nn = n + r_sel;
    if (r_sel < R && m < MB && m < M && nn < N) {
      float v = 0.f;
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int mm = 0; mm < MB; ++mm)
          if (r == r_sel && mm == m) v = acc[r][mm];
      if (o.bias) v += o.bias[nn];
      if (nn < o.alpha_cols) v *= o.alpha;
      if (o.act == 1) v = gelu_erf(v);
      if (o.residual) v += __ldcg(o.residual + (long long)m * o.ldo + nn);
      o.out[(long long)m * o.ldo + nn] = v;
      if (o.kc && nn >= o.D) {
        const long long row = ((long long)m * o.Tmax + o.pos) * o.D;
        if (nn < 2 * o.D) o.kc[row + nn - o.D] = __float2bfloat16(v);
        else o.vc[row + nn - 2 * o.D] = __float2bfloat16(v);
      }
    }
    if (PIPE) {
      if (has_next) cur = nxt;
    } else if (has_next) {
      load_rows_m<NC, R>(cur, W, K, n2, N, lane);
    }
  }
}

template <int NC, int R>
__device__ __forceinline__ void prefetch_rows(WR<NC, R>& w, const bf16* W, int N, int K, int gw, int lane) {
  if (gw * R < N) load_rows_m<NC, R>(w, W, K, gw * R, N, lane);
}

__device__ __forceinline__ float block_max_m(float v, float* red) {
  v = warp_max(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = red[0];
#pragma unroll
  for (int w = 1; w < MW; ++w) r = fmaxf(r, red[w]);
  __syncthreads();
  return r;
}
__device__ __forceinline__ float block_sum_m(float v, float* red) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
#pragma unroll
  for (int w = 0; w < MW; ++w) r += red[w];
  __syncthreads();
  return r;
}

// smem carve-up (dynamic): red [MW*8 + 64] floats | union { xs [MB*ffn] floats (GEMV phases), attention scratch }
template <int MB>
__global__ void __launch_bounds__(MT, 1) decode_mega_kernel(const __grid_constant__ MegaArgs a) {
  extern __shared__ __align__(16) uint8_t dyn[];
  float* red = reinterpret_cast<float*>(dyn);
  float* xs = red + MW * 8 + 64;
  uint8_t* att = reinterpret_cast<uint8_t*>(xs);  // attention phases never overlap a GEMV phase
  __shared__ unsigned s_last;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gw = blockIdx.x * MW + warp, GW = gridDim.x * MW;
  const int D = a.D, H = a.H, Q = a.Q, ffn = a.ffn;
  const int pos = *a.pos;
  GridBar bar{a.bar, gridDim.x, 0u};

  // ---- phase 0: embedding (CTA 0 writes the residual stream), first QKV rows requested meanwhile
  WR<5, 3> w53;  // QKV and fc1: 3 rows per warp (1776 warps x 3 >= 5120 rows: one pass)
  WR<5, 2> w52;  // out-proj / cross-q / LM head
  prefetch_rows<5, 3>(w53, a.layers[0].wqkv, 3 * D, D, gw, lane);
  if (blockIdx.x == 0) {
    for (int i = threadIdx.x; i < Q * D; i += MT) {
      const int q = i / D, d = i - q * D;
      const int tok = a.tokens[q * a.Tmax + pos];
      a.dx[i] = __bfloat162float(a.embed[(long long)tok * D + d]) + a.dec_pos[(long long)pos * D + d];
    }
  }
  bar.sync();

  for (int l = 0; l < a.L; ++l) {
    const MegaLayer& L = a.layers[l];
    // ---------------- A: LN1 + fused QKV (+ self-KV append) ----------------
    stage_x<MB>(xs, red, a.dx, D, D, Q, L.ln1g, L.ln1b);
    {
      PhaseOut o{L.bqkv, 0.125f, D, 0, nullptr, a.dqkv, 3 * D, L.self_k, L.self_v, D, a.Tmax, pos};
      gemv_phase<MB, 5, 3, false>(w53, L.wqkv, 3 * D, D, xs, Q, o, gw, GW, lane);
    }
    prefetch_rows<5, 2>(w52, L.wo, D, D, gw, lane);
    bar.sync();
    // ---------------- B: causal self-attention, one (sequence, head) per CTA ----------------
    for (int item = blockIdx.x; item < Q * H; item += gridDim.x) {
      const int q = item / H, h = item - q * H;
      const int n = pos + 1;
      const int grp = threadIdx.x >> 3, sub = threadIdx.x & 7;  // KG key groups x 8 lanes
      uint8_t* sK = att;
      uint8_t* sV = att + (size_t)MAXKEYS * 128;
      float* sc = reinterpret_cast<float*>(att + (size_t)MAXKEYS * 256);
      float* redo = sc + MAXKEYS;  // [KG][64]
      for (int s = grp; s < n; s += KG) {
        const long long off = ((long long)q * a.Tmax + s) * D + h * 64 + sub * 8;
        cp_async16m(sK + s * 128 + sub * 16, L.self_k + off);
        cp_async16m(sV + s * 128 + sub * 16, L.self_v + off);
      }
      float qv[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) qv[j] = __ldcg(a.dqkv + (long long)q * 3 * D + h * 64 + sub * 8 + j);
      cp_async_wait_allm();
      float lmax = -INFINITY;
      for (int sb = 0; sb < n; sb += KG) {
        const int s = sb + grp;
        float d = 0.f;
        if (s < n) {
          float kf[8];
          unpack8m(*reinterpret_cast<const uint4*>(sK + s * 128 + sub * 16), kf);
#pragma unroll
          for (int j = 0; j < 8; ++j) d = fmaf(qv[j], kf[j], d);
        }
        d += __shfl_xor_sync(0xffffffffu, d, 1);
        d += __shfl_xor_sync(0xffffffffu, d, 2);
        d += __shfl_xor_sync(0xffffffffu, d, 4);
        if (s < n) {
          if (sub == 0) sc[s] = d;
          lmax = fmaxf(lmax, d);
        }
      }
      const float mx = block_max_m(lmax, red);
      float lsum = 0.f;
      for (int s = threadIdx.x; s < n; s += MT) {
        const float e = __expf(sc[s] - mx);
        sc[s] = e;
        lsum += e;
      }
      const float inv = 1.0f / block_sum_m(lsum, red);
      float acc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = 0.f;
      for (int s = grp; s < n; s += KG) {
        float vf[8];
        unpack8m(*reinterpret_cast<const uint4*>(sV + s * 128 + sub * 16), vf);
        const float p = sc[s];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(p, vf[j], acc[j]);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) redo[grp * 64 + sub * 8 + j] = acc[j];
      __syncthreads();
      if (threadIdx.x < 64) {
        float ov = 0.f;
        for (int g = 0; g < KG; ++g) ov += redo[g * 64 + threadIdx.x];
        a.dattn[(long long)q * D + h * 64 + threadIdx.x] = ov * inv;
      }
      __syncthreads();
    }
    bar.sync();
    // ---------------- C: self out-proj + residual ----------------
    stage_x<MB>(xs, red, a.dattn, D, D, Q, nullptr, nullptr);
    {
      PhaseOut o{L.bo, 1.f, 0, 0, a.dx, a.dx, D, nullptr, nullptr, D, a.Tmax, pos};
      gemv_phase<MB, 5, 2, false>(w52, L.wo, D, D, xs, Q, o, gw, GW, lane);
    }
    prefetch_rows<5, 2>(w52, L.xwq, D, D, gw, lane);
    bar.sync();
    // ---------------- D: LN2 + cross q projection ----------------
    stage_x<MB>(xs, red, a.dx, D, D, Q, L.ln2g, L.ln2b);
    {
      PhaseOut o{L.xbq, 0.125f, D, 0, nullptr, a.dq, D, nullptr, nullptr, D, a.Tmax, pos};
      gemv_phase<MB, 5, 2, false>(w52, L.xwq, D, D, xs, Q, o, gw, GW, lane);
    }
    prefetch_rows<5, 2>(w52, L.xwo, D, D, gw, lane);
    bar.sync();
    // ---------------- E: cross-attention, (audio, head, key split) items; last split of a head merges ----------------
    {
      const int nsplit = a.nsplit;
      const int ks = (a.S + nsplit - 1) / nsplit;
      const int grp = threadIdx.x >> 3, sub = threadIdx.x & 7;
      uint8_t* sK = att;
      uint8_t* sV = att + XKMAX * 128;
      float* sc = reinterpret_cast<float*>(att + 2 * XKMAX * 128);
      float* redo = sc + XKMAX;  // [KG][64]
      for (int item = blockIdx.x; item < Q * H * nsplit; item += gridDim.x) {
        const int split = item % nsplit;
        const int h = (item / nsplit) % H;
        const int q = item / (nsplit * H);
        const int s0 = split * ks;
        const int n = max(0, min(a.S, s0 + ks) - s0);
        const bf16* kbase = L.cross_k + (((long long)q * H + h) * a.S + s0) * 64 + sub * 8;
        const bf16* vbase = L.cross_v + (((long long)q * H + h) * a.S + s0) * 64 + sub * 8;
        for (int kk = grp; kk < n; kk += KG) {
          cp_async16m(sK + kk * 128 + sub * 16, kbase + (long long)kk * 64);
          cp_async16m(sV + kk * 128 + sub * 16, vbase + (long long)kk * 64);
        }
        float qv[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) qv[j] = __ldcg(a.dq + (long long)q * D + h * 64 + sub * 8 + j);
        float* align_row = nullptr;
        if (a.align && L.head_slots) {
          const int slot = L.head_slots[h];
          const int step = pos - a.step_base;
          if (slot >= 0 && step >= 0 && step < a.Tcap) align_row = a.align + (((long long)q * a.Ha + slot) * a.Tcap + step) * a.S + s0;
        }
        cp_async_wait_allm();
        float lmax = -INFINITY;
        for (int kb = 0; kb < n; kb += KG) {
          const int kk = kb + grp;
          float d = 0.f;
          if (kk < n) {
            float kf[8];
            unpack8m(*reinterpret_cast<const uint4*>(sK + kk * 128 + sub * 16), kf);
#pragma unroll
            for (int j = 0; j < 8; ++j) d = fmaf(qv[j], kf[j], d);
          }
          d += __shfl_xor_sync(0xffffffffu, d, 1);
          d += __shfl_xor_sync(0xffffffffu, d, 2);
          d += __shfl_xor_sync(0xffffffffu, d, 4);
          if (kk < n) {
            if (sub == 0) {
              sc[kk] = d;
              if (align_row) align_row[kk] = d;
            }
            lmax = fmaxf(lmax, d);
          }
        }
        const float mx = block_max_m(lmax, red);
        float lsum = 0.f;
        for (int kk = threadIdx.x; kk < n; kk += MT) {
          const float e = __expf(sc[kk] - mx);
          sc[kk] = e;
          lsum += e;
        }
        const float lsumt = block_sum_m(lsum, red);
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
        for (int kk = grp; kk < n; kk += KG) {
          float vf[8];
          unpack8m(*reinterpret_cast<const uint4*>(sV + kk * 128 + sub * 16), vf);
          const float p = sc[kk];
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] = fmaf(p, vf[j], acc[j]);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) redo[grp * 64 + sub * 8 + j] = acc[j];
        __syncthreads();
        const long long pb = ((long long)q * H + h) * nsplit + split;
        if (threadIdx.x < 64) {
          float ov = 0.f;
          for (int g = 0; g < KG; ++g) ov += redo[g * 64 + threadIdx.x];
          a.part_o[pb * 64 + threadIdx.x] = ov;
        }
        if (threadIdx.x == 0) {
          a.part_ml[pb * 2 + 0] = mx;
          a.part_ml[pb * 2 + 1] = lsumt;
        }
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) {
          const unsigned prev = atomicAdd(&a.xcounters[q * H + h], 1u);
          s_last = (prev == (unsigned)(nsplit - 1)) ? 1u : 0u;
        }
        __syncthreads();
        if (s_last) {
          __threadfence();
          if (threadIdx.x < 64) {
            const long long hb = ((long long)q * H + h) * nsplit;
            float M = -INFINITY;
            for (int sp = 0; sp < nsplit; ++sp)
              if (__ldcg(&a.part_ml[(hb + sp) * 2 + 1]) > 0.f) M = fmaxf(M, __ldcg(&a.part_ml[(hb + sp) * 2]));
            float Lsum = 0.f, ov = 0.f;
            for (int sp = 0; sp < nsplit; ++sp) {
              const float lv = __ldcg(&a.part_ml[(hb + sp) * 2 + 1]);
              if (lv > 0.f) {
                const float w = __expf(__ldcg(&a.part_ml[(hb + sp) * 2]) - M);
                Lsum = fmaf(lv, w, Lsum);
                ov = fmaf(__ldcg(&a.part_o[(hb + sp) * 64 + threadIdx.x]), w, ov);
              }
            }
            a.dattn[(long long)q * D + h * 64 + threadIdx.x] = ov / Lsum;
          }
          if (threadIdx.x == 0) a.xcounters[q * H + h] = 0u;
        }
        __syncthreads();
      }
    }
    bar.sync();
    // ---------------- F: cross out-proj + residual ----------------
    stage_x<MB>(xs, red, a.dattn, D, D, Q, nullptr, nullptr);
    {
      PhaseOut o{L.xbo, 1.f, 0, 0, a.dx, a.dx, D, nullptr, nullptr, D, a.Tmax, pos};
      gemv_phase<MB, 5, 2, false>(w52, L.xwo, D, D, xs, Q, o, gw, GW, lane);
    }
    prefetch_rows<5, 3>(w53, L.w1, ffn, D, gw, lane);
    bar.sync();
    // ---------------- G: LN3 + fc1 + GELU ----------------
    stage_x<MB>(xs, red, a.dx, D, D, Q, L.ln3g, L.ln3b);
    {
      PhaseOut o{L.b1, 1.f, 0, 1, nullptr, a.dh, ffn, nullptr, nullptr, D, a.Tmax, pos};
      gemv_phase<MB, 5, 3, false>(w53, L.w1, ffn, D, xs, Q, o, gw, GW, lane);
    }
    {
      // ---------------- H: fc2 + residual (K = ffn: one row per warp, 20 loads in flight) ----------------
      WR<20, 1> w201;
      prefetch_rows<20, 1>(w201, L.w2, D, ffn, gw, lane);
      bar.sync();
      stage_x<MB>(xs, red, a.dh, ffn, ffn, Q, nullptr, nullptr);
      PhaseOut o{L.b2, 1.f, 0, 0, a.dx, a.dx, D, nullptr, nullptr, D, a.Tmax, pos};
      gemv_phase<MB, 20, 1, false>(w201, L.w2, D, ffn, xs, Q, o, gw, GW, lane);
    }
    if (l + 1 < a.L) prefetch_rows<5, 3>(w53, a.layers[l + 1].wqkv, 3 * D, D, gw, lane);
    else prefetch_rows<5, 2>(w52, a.embed, a.V, D, gw, lane);
    bar.sync();
  }
  // ---------------- final LayerNorm + tied LM head ----------------
  stage_x<MB>(xs, red, a.dx, D, D, Q, a.lnf_g, a.lnf_b);
  {
    PhaseOut o{nullptr, 1.f, 0, 0, nullptr, a.logits, a.V, nullptr, nullptr, D, a.Tmax, pos};
    gemv_phase<MB, 5, 2, true>(w52, a.embed, a.V, D, xs, Q, o, gw, GW, lane);
  }
}

}  // namespace

size_t mega_smem_bytes(int mb, int ffn) {
  const size_t attn = (size_t)MAXKEYS * 256 + (size_t)(MAXKEYS + KG * 64) * sizeof(float);
  const size_t xattn = (size_t)2 * XKMAX * 128 + (size_t)(XKMAX + KG * 64) * sizeof(float);
  const size_t xs = (size_t)mb * ffn * sizeof(float);
  const size_t u = xs > attn ? (xs > xattn ? xs : xattn) : (attn > xattn ? attn : xattn);
  return (MW * 8 + 64) * sizeof(float) + u + 64;
}

// Launches the persistent step kernel on `st`.  Returns -3 when the configuration is outside what it supports
// (the caller then uses the per-op path).
int launch_decode_mega(cudaStream_t st, const MegaArgs& a, int num_sms) {
  const int Q = a.Q;
  if (a.L > MEGA_MAXL || Q > 8 || a.D > 1280 || a.ffn > 5120 || a.D % 8 != 0 || a.Tmax > MAXKEYS) return -3;
  const int mb = Q <= 1 ? 1 : (Q <= 2 ? 2 : (Q <= 4 ? 4 : 8));
  const size_t smem = mega_smem_bytes(mb, a.ffn);
  if (smem > 226 * 1024) return -3;  // 227 KB opt-in limit includes the few bytes of static smem
  const int ks = (a.S + a.nsplit - 1) / a.nsplit;
  if (ks > XKMAX) return -3;
  BW_CUDA_OK(cudaMemsetAsync(a.bar, 0, sizeof(unsigned), st));
#define BW_MEGA_CASE(MB)                                                                                              \
  case MB: {                                                                                                          \
    static size_t attr = 0;                                                                                           \
    if (smem > attr) {                                                                                                \
      BW_CUDA_OK(cudaFuncSetAttribute(decode_mega_kernel<MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      attr = smem;                                                                                                    \
    }                                                                                                                 \
    decode_mega_kernel<MB><<<num_sms, MT, smem, st>>>(a);                                                             \
  } break;
  switch (mb) {
    BW_MEGA_CASE(1)
    BW_MEGA_CASE(2)
    BW_MEGA_CASE(4)
    BW_MEGA_CASE(8)
  }
#undef BW_MEGA_CASE
  BW_CUDA_OK(cudaGetLastError());
  return 0;
}


}  // namespace bw
