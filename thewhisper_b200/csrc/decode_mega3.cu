// Third generation of the persistent decoder-step kernel (one sequence, greedy): FOUR grid-wide hand-overs per layer
// instead of eight.
//
// Why (profiles/r1_v8_summary.md): a 148-CTA hand-over through L2 costs 1.3 us (barrier) + 0.3 us (every CTA reads the
// result), nothing that polls data is cheaper, and 257 of them are a third of the step.  The two attention phases and
// the out-projections that follow them are restructured so that their results never have to be gathered:
//   * the 7 CTAs (h, j) of head h all compute head h's self-attention (32 KB of K/V each, redundantly), then CTA (h, j)
//     multiplies the head's 64 outputs with rows [j*RS, (j+1)*RS) of the head's 64-column block of Wo -- a [rows][64] slab
//     that is contiguous in the head-major copy of Wo bound as MegaArgs::wo_hm -- and ADDS its partial row sums into a
//     64-bit fixed-point accumulator (2^-24 resolution: integer adds commute, so the sum is bit-reproducible whatever
//     the order of the 20 heads -- fp32 atomics would not be);
//   * cross-attention the same way with (head, key split) items: the 7 splits of a head exchange their (max, sum, 64
//     outputs) through global memory behind a per-head counter (7 arrivals, 7 pollers), every one merges them, and split j
//     out-projects row slice j of the head's block of the cross Wo;
//   * the consumers (LN2 + cross-q, LN3 + fc1) add residual + bias + accumulator while they stage x; CTA 0 writes the
//     new residual stream back (ping-pong between dx and dx2) and clears the accumulator that is no longer needed;
//   * in front of both attention phases only the producers of a head's q / k / v rows are waited for (per-head counters).
// Per layer: QKV -(per-head)-> self-attn + out-proj -(grid)-> cross-q -(per-head)-> cross-attn + out-proj -(grid)->
// fc1 -(grid)-> fc2 -(grid)->.
// Everything else (TMA slab pipeline, L2 prefetch, LayerNorm staging, dot products, LM head, fused greedy selection) is
// decode_mega.cu's.  Slab phases keep their parity (even phases in region 0, odd ones at the pool's start); the slab of
// the phase that follows an attention phase can only be requested when the attention scratch (which overlays region 0)
// is free again, i.e. just before the grid barrier, which hides the copy.
#include "decode_mega_common.cuh"

namespace bw {

namespace {

using namespace mega;

constexpr int RSMAX = MW * 16;  // out-projection rows of one (head, slice) item: 16 per warp
constexpr float FIX_SCALE = 16777216.f;  // 2^24
constexpr int CNT_QKV = 256, CNT_XQ = 288, CNT_XHEAD = 384;  // word offsets of the per-head counters inside MegaArgs::bar


__device__ __forceinline__ unsigned long long f2fix(float v) { return (unsigned long long)__float2ll_rn(v * FIX_SCALE); }
__device__ __forceinline__ float fix2f(unsigned long long a) { return (float)((double)(long long)a * (1.0 / 16777216.0)); }

// grid barrier (decode_mega.cu's, without the trace): arrival by the CTA's last thread, which never has a load in flight
struct GridBar {
  unsigned* ctr;
  unsigned nblocks;
  unsigned epoch;
  __device__ __forceinline__ void arrive() {
    if (threadIdx.x == MT - 1) red_release_add(ctr, 1u);
  }
  __device__ __forceinline__ void wait() {
    if (threadIdx.x == MT - 1) {
      const unsigned target = (epoch + 1) * nblocks;
      if (ld_acquire_u32(ctr) < target) {
        const long long t0 = clock64();
        while (ld_acquire_u32(ctr) < target) {
          if (clock64() - t0 > (1ll << 32)) {
            printf("[bw] decode_mega3: grid barrier %u timed out (block %d)\n", epoch, blockIdx.x);
            __trap();
          }
        }
      }
    }
    ++epoch;
    __syncthreads();
  }
  __device__ __forceinline__ void sync() {
    __syncthreads();
    arrive();
    wait();
  }
};

// per-head readiness counters (zeroed with the barrier words before every launch; they count layers)
__device__ __forceinline__ void head_signal(unsigned* ctr, int n0, int nend, int D) {
  if (threadIdx.x == MT - 1 && n0 < nend) {
    const int h0 = (n0 % D) >> 6, h1 = ((nend - 1) % D) >> 6;  // a CTA owns fewer than 64 rows: at most two head ranges
    red_release_add(ctr + h0, 1u);
    if (h1 != h0) red_release_add(ctr + h1, 1u);
  }
}
__device__ __forceinline__ unsigned head_expected(int h, int D, int nblk, int rc) {
  unsigned n = 0;
  for (int b = 0; b < nblk; ++b) {
    const int s0 = b * D + h * 64;
    n += (unsigned)((s0 + 63) / rc - s0 / rc + 1);
  }
  return n;
}
__device__ __forceinline__ void counter_wait(const unsigned* ctr, unsigned target) {
  if (threadIdx.x == MT - 1 && ld_acquire_u32(ctr) < target) {
    const long long t0 = clock64();
    while (ld_acquire_u32(ctr) < target) {
      if (clock64() - t0 > (1ll << 32)) {
        printf("[bw] decode_mega3: counter wait timed out (block %d, have %u, want %u)\n", blockIdx.x, ld_acquire_u32(ctr), target);
        __trap();
      }
    }
  }
  __syncthreads();
}

// One GEMV phase: out[n] = epi(sum_k W[n][k] * LN?(x)[k] + bias[n]), x = src (+ cbias + fixed-point accumulator).
struct GemvDesc {
  const bf16* W;
  const float* bias;
  int N, K, R;            // R rows per warp (1, 2 or 3)
  int n0, nend;           // rows of this CTA: [n0, nend), contiguous, ceil(N / CTAs) each (the LM head streams: [0, N))
  bool lm;
  const float* src;       // [K] fp32
  const float *lng, *lnb; // LayerNorm applied while staging (nullptr: none)
  // x = src + cbias + fix2f(acc) while staging (the fused attention + out-projection phases left their sums in acc);
  // CTA 0 stores the resulting residual stream to xout.  zero: an accumulator that nobody reads any more (CTA 0 clears it).
  const unsigned long long* acc;
  const float* cbias;
  float* xout;
  unsigned long long* zero;
  float* out;
  int ldo;
  const float* residual;  // may alias out
  int act;                // 1: GELU
  float alpha;
  int alpha_cols;
  bf16 *kc, *vc;          // optional self-KV append
};

__device__ __forceinline__ void split_rows(GemvDesc& d) {
  const int rc = (d.N + (int)gridDim.x - 1) / (int)gridDim.x;
  d.R = (rc + MW - 1) / MW;
  d.n0 = min(d.N, (int)blockIdx.x * rc);
  d.nend = min(d.N, d.n0 + rc);
}
// gi: 0 LN1+QKV | 1 LN2+cross q | 2 LN3+fc1+GELU | 3 fc2; l == a.L: final LN + LM head
__device__ __forceinline__ GemvDesc make_desc(const MegaArgs& a, const MegaLayer* layers, int l, int gi) {
  GemvDesc d;
  d.lng = d.lnb = nullptr;
  d.residual = nullptr;
  d.act = 0;
  d.alpha = 1.f;
  d.alpha_cols = 0;
  d.kc = d.vc = nullptr;
  d.acc = nullptr;
  d.cbias = nullptr;
  d.xout = nullptr;
  d.zero = nullptr;
  d.N = d.K = d.ldo = a.D;
  d.lm = false;
  if (l >= a.L) {
    d.W = a.embed; d.bias = nullptr; d.N = a.V; d.R = 2; d.n0 = 0; d.nend = a.V; d.lm = true; d.src = a.dx; d.lng = a.lnf_g; d.lnb = a.lnf_b;
    d.out = a.logits; d.ldo = a.V; d.zero = a.acc_b;
    return d;
  }
  const MegaLayer& L = layers[l];
  switch (gi) {
    case 0:
      d.W = L.wqkv; d.bias = L.bqkv; d.N = 3 * a.D; d.src = a.dx; d.lng = L.ln1g; d.lnb = L.ln1b; d.out = a.dqkv; d.ldo = 3 * a.D;
      d.alpha = 0.125f; d.alpha_cols = a.D; d.kc = L.self_k; d.vc = L.self_v; d.zero = (l > 0) ? a.acc_b : nullptr;
      break;
    case 1:  // x1 = x0 + bo + sum_h (Wo_h a_h): dx + acc_a -> dx2
      d.W = L.xwq; d.bias = L.xbq; d.src = a.dx; d.acc = a.acc_a; d.cbias = L.bo; d.xout = a.dx2; d.lng = L.ln2g; d.lnb = L.ln2b; d.out = a.dq;
      d.alpha = 0.125f; d.alpha_cols = a.D;
      break;
    case 2:  // x2 = x1 + xbo + sum_h (Wxo_h a2_h): dx2 + acc_b -> dx; acc_a is clear again after this phase
      d.W = L.w1; d.bias = L.b1; d.N = a.ffn; d.src = a.dx2; d.acc = a.acc_b; d.cbias = L.xbo; d.xout = a.dx; d.zero = a.acc_a;
      d.lng = L.ln3g; d.lnb = L.ln3b; d.out = a.dh; d.ldo = a.ffn; d.act = 1;
      break;
    default:
      d.W = L.w2; d.bias = L.b2; d.K = a.ffn; d.src = a.dh; d.out = a.dx; d.residual = a.dx;
      break;
  }
  split_rows(d);
  return d;
}

// The [rows][64] slab of a fused attention + out-projection item: CTA b = (h, j) multiplies head h's 64 attention outputs
// with rows [j*RS, (j+1)*RS) of the head's column block.  whm: head-major copy [H][D][64] of the out-projection matrix.
struct SliceDesc {
  const bf16* W;  // first row of the slice
  int r0, rows;   // output rows [r0, r0 + rows)
};
__device__ __forceinline__ SliceDesc make_slice(const MegaArgs& a, const bf16* whm) {
  SliceDesc s;
  s.W = nullptr;
  s.r0 = 0;
  s.rows = 0;
  const int nj = a.nsplit;
  if ((int)blockIdx.x < a.H * nj) {
    const int h = blockIdx.x / nj, j = blockIdx.x - h * nj;
    const int rs = (a.D + nj - 1) / nj;
    s.r0 = min(a.D, j * rs);
    s.rows = min(a.D, s.r0 + rs) - s.r0;
    s.W = whm + ((long long)h * a.D + s.r0) * 64;
  }
  return s;
}
__device__ __forceinline__ void issue_slice(const SliceDesc& s, uint8_t* region, uint64_t* cbar) {
  if (threadIdx.x == DMA_T && s.rows > 0) {
    const uint32_t bytes = (uint32_t)s.rows * 128u;
    mbar_arrive_expect_tx(cbar, bytes);
    bulk_g2s(region, s.W, bytes, cbar);
  }
}

// What a warp requests before the barrier that precedes a GEMV phase: bias values and this thread's LayerNorm slice.
struct Pre {
  float bias;   // of the row this lane finishes (lanes [8r, 8r + 1) finish row n + r)
  float4 g, b;  // gamma / beta of elements [4*tid, 4*tid + 4)
};

// The rows of a CTA are contiguous in memory and so are their slabs in smem: one TMA operation per CTA and phase.
__device__ __forceinline__ void l2_prefetch_phase(const GemvDesc& d) {
  if (threadIdx.x == DMA_T && !d.lm && d.n0 < d.nend) l2_prefetch(d.W + (long long)d.n0 * d.K, (uint32_t)(d.nend - d.n0) * d.K * 2);
}
__device__ __forceinline__ void l2_prefetch_slice(const SliceDesc& s) {
  if (threadIdx.x == DMA_T && s.rows > 0) l2_prefetch(s.W, (uint32_t)s.rows * 128u);
}
__device__ __forceinline__ void issue_slabs(const GemvDesc& d, uint8_t* region, uint64_t* cbar) {
  if (threadIdx.x == DMA_T && d.n0 < d.nend) {
    const uint32_t bytes = (uint32_t)(d.nend - d.n0) * d.K * 2;
    mbar_arrive_expect_tx(cbar, bytes);
    bulk_g2s(region, d.W + (long long)d.n0 * d.K, bytes, cbar);
  }
}

__device__ __forceinline__ void prefetch_phase(const GemvDesc& d, Pre& p, uint8_t* pool, uint64_t* wbar, int gw, int warp, int lane) {
  int n;
  if (d.lm) {
    n = gw * d.R;
    if (n < d.N) issue_rows(pool + (size_t)warp * d.R * d.K * 2, wbar, d.W, d.K, d.R, n, d.N, lane);
  } else {
    n = d.n0 + warp * d.R;
  }
  const int r_sel = lane >> 3;
  p.bias = (d.bias && (lane & 7) < 1 && r_sel < d.R && n + r_sel < d.nend) ? d.bias[n + r_sel] : 0.f;
  const int k = threadIdx.x * 4;
  if (d.lng && k < d.K) {
    p.g = *reinterpret_cast<const float4*>(d.lng + k);
    p.b = *reinterpret_cast<const float4*>(d.lnb + k);
  }
}

__device__ __forceinline__ void stage_sync() { asm volatile("bar.sync 1, %0;" ::"n"(MT - 32) : "memory"); }
// the staging warps synchronise among themselves and only signal the DMA warp, which joins after its TMA requests
__device__ __forceinline__ void stage_done_stagers() {
  asm volatile("bar.sync 1, %0;" ::"n"(MT - 32) : "memory");
  asm volatile("bar.arrive 3, %0;" ::"n"(MT) : "memory");
}
__device__ __forceinline__ void stage_done_dma() { asm volatile("bar.sync 3, %0;" ::"n"(MT) : "memory"); }

// stage x (one sequence) into smem: src (+ bias + fixed-point accumulator), LayerNormed when the phase has one.  The last
// warp runs `dma` meanwhile.  CTA 0 stores the combined residual stream and clears the accumulator nobody needs any more.
template <class Dma>
__device__ __forceinline__ void stage_x(float* xs, float* red, const GemvDesc& d, const Pre& p, Dma&& dma) {
  const int K = d.K;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == MW - 1) {
    dma();
    stage_done_dma();
    return;
  }
  constexpr int ST = MT - 32;  // staging threads
  if (!d.lng) {
    constexpr int U = 4;
    for (int base = threadIdx.x * 4; base < K; base += ST * 4 * U) {
      float4 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = base + u * ST * 4;
        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < K) v[u] = __ldcg(reinterpret_cast<const float4*>(d.src + i));
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = base + u * ST * 4;
        if (i < K) *reinterpret_cast<float4*>(xs + i) = v[u];
      }
    }
    stage_done_stagers();
    return;
  }
  // LayerNorm (K <= 4 * ST): one float4 per thread, two-pass statistics through two reductions on register values
  const int k = threadIdx.x * 4;
  const bool have = k < K;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (have) v = __ldcg(reinterpret_cast<const float4*>(d.src + k));
  if (d.acc) {
    if (have) {
      const ulonglong2 a0 = __ldcg(reinterpret_cast<const ulonglong2*>(d.acc + k));
      const ulonglong2 a1 = __ldcg(reinterpret_cast<const ulonglong2*>(d.acc + k + 2));
      const float4 cb = *reinterpret_cast<const float4*>(d.cbias + k);
      v.x = (v.x + cb.x) + fix2f(a0.x);
      v.y = (v.y + cb.y) + fix2f(a0.y);
      v.z = (v.z + cb.z) + fix2f(a1.x);
      v.w = (v.w + cb.w) + fix2f(a1.y);
      if (blockIdx.x == 0) *reinterpret_cast<float4*>(d.xout + k) = v;
    }
  }
  if (d.zero && blockIdx.x == 0 && have) {
    *reinterpret_cast<ulonglong2*>(d.zero + k) = make_ulonglong2(0ull, 0ull);
    *reinterpret_cast<ulonglong2*>(d.zero + k + 2) = make_ulonglong2(0ull, 0ull);
  }
  {
    const float s = warp_sum((v.x + v.y) + (v.z + v.w));
    if (lane == 0) red[warp] = s;
  }
  stage_sync();
  float mean;
  {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < MW - 1; ++w) s += red[w];
    mean = s / (float)K;
    float ss = 0.f;
    if (have) {
      const float a0 = v.x - mean, a1 = v.y - mean, a2 = v.z - mean, a3 = v.w - mean;
      ss = (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
    }
    ss = warp_sum(ss);
    if (lane == 0) red[MW + warp] = ss;
  }
  stage_sync();
  {
    float ss = 0.f;
#pragma unroll
    for (int w = 0; w < MW - 1; ++w) ss += red[MW + w];
    const float rstd = rsqrtf(ss / (float)K + 1e-5f);
    if (have) {
      float4 o;
      o.x = (v.x - mean) * rstd * p.g.x + p.b.x;
      o.y = (v.y - mean) * rstd * p.g.y + p.b.y;
      o.z = (v.z - mean) * rstd * p.g.z + p.b.z;
      o.w = (v.w - mean) * rstd * p.g.w + p.b.w;
      *reinterpret_cast<float4*>(xs + k) = o;
    }
  }
  stage_done_stagers();
}

// lanes 8r finish row n + r  (R <= 3, one sequence)
__device__ __forceinline__ void finish_rows(const GemvDesc& d, const float (&acc)[3][1], float bias, int n, float res, int D, int Tmax, int pos, int lane) {
  const int m = lane & 7, r_sel = lane >> 3;
  const int nn = n + r_sel;
  if (r_sel < d.R && m == 0 && nn < d.nend) {
    float v = r_sel == 0 ? acc[0][0] : (r_sel == 1 ? acc[1][0] : acc[2][0]);
    v += bias;
    if (nn < d.alpha_cols) v *= d.alpha;
    if (d.act == 1) v = gelu_erf(v);
    if (d.residual) v += res;
    d.out[nn] = v;
    if (d.kc && nn >= D) {
      const long long row = (long long)pos * D;
      if (nn < 2 * D) d.kc[row + nn - D] = __float2bfloat16(v);
      else d.vc[row + nn - 2 * D] = __float2bfloat16(v);
    }
  }
}

// Out-projection slice of a fused attention item: rows [0, s.rows) of the [rows][64] bf16 slab against the head's 64
// attention outputs (a_s, smem); warp w owns rows 16w .. 16w + 15, lane l columns 2l, 2l + 1; the totals are ADDED to the
// 64-bit fixed-point accumulator acc[s.r0 + row] (integer adds commute: the result does not depend on the order of the heads).
__device__ __forceinline__ void outproj_slice(const uint8_t* slab, const float* a_s, const SliceDesc& s, unsigned long long* acc, int warp, int lane) {
  const float2 x = *reinterpret_cast<const float2*>(a_s + 2 * lane);
  float v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int r = warp * 16 + i;
    float t = 0.f;
    if (r < s.rows) {
      const float2 w = unpack_bf16(*reinterpret_cast<const uint32_t*>(slab + (size_t)r * 128 + lane * 4));
      t = fmaf(w.y, x.y, w.x * x.x);
    }
    v[i] = t;
  }
  const float tot = treduce<16>(v, lane);
  const int row = warp * 16 + (lane >> 1);
  if ((lane & 1) == 0 && row < s.rows) atomicAdd(acc + s.r0 + row, f2fix(tot));
}

// smem carve-up (dynamic): red [64] | xs [ffn] | pool: weight slabs from 0, attention scratch from ATT_OFF
__global__ void __launch_bounds__(MT, 1) decode_mega3_kernel(const __grid_constant__ MegaArgs a) {
  constexpr int MB = 1;
  extern __shared__ __align__(128) uint8_t dyn[];
  float* red = reinterpret_cast<float*>(dyn);
  float* xs = red + 64;
  uint8_t* pool = reinterpret_cast<uint8_t*>(xs + (size_t)MB * a.ffn);
  uint8_t* att = pool + ATT_OFF;
  __shared__ __align__(8) uint64_t wbar[2 * MW];  // LM head: per warp, two slab stages
  __shared__ __align__(8) uint64_t xbar;          // cross-attention K/V item
  __shared__ __align__(8) uint64_t cbar[2];       // the CTA's weight slab of a phase (one per slab region)
  __shared__ __align__(16) float a_s[64];         // the head's attention outputs of a fused item
  __shared__ __align__(16) MegaLayer sl[MEGA_MAXL];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gw = blockIdx.x * MW + warp, GW = gridDim.x * MW;
  const int D = a.D, H = a.H;
  const int pos0 = *a.pos;
  GridBar bar{a.bar, gridDim.x, 0u};
  const int nsplit = a.nsplit;  // row slices of the self-attention items = key splits of the cross-attention items
  const int ks = (a.S + nsplit - 1) / nsplit;
  const int grp = threadIdx.x >> 3, sub = threadIdx.x & 7;
  const bool item_cta = (int)blockIdx.x < H * nsplit;
  const int ih = item_cta ? (int)blockIdx.x / nsplit : 0;       // head of this CTA's items
  const int ij = item_cta ? (int)blockIdx.x - ih * nsplit : 0;  // row slice / key split
  uint32_t wpar = 0, wpar1 = 0, xpar = 0, cpar0 = 0, cpar1 = 0;  // mbarrier phase parities
  uint8_t* const reg0 = pool + a.p0_off;  // even slab phases (QKV, cross-q, fc1)
  uint8_t* const reg1 = pool;             // odd slab phases (the two out-projection slices, fc2)

  {
    static_assert(sizeof(MegaLayer) % 8 == 0, "MegaLayer is copied in 8-byte words");
    const unsigned long long* src = reinterpret_cast<const unsigned long long*>(a.layers);
    unsigned long long* dst = reinterpret_cast<unsigned long long*>(sl);
    for (int i = threadIdx.x; i < a.L * (int)(sizeof(MegaLayer) / 8); i += MT) dst[i] = src[i];
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2 * MW; ++i) mbar_init(&wbar[i], 1);
    mbar_init(&xbar, 1);
    mbar_init(&cbar[0], 1);
    mbar_init(&cbar[1], 1);
    fence_mbar_init();
  }
  __syncthreads();

  // MegaArgs::n_steps decoder steps per launch: the previous step's last CTA wrote the token and advanced the position
  // before it arrived at the barrier that ends a step, so both are read from L2; the per-head counters count layers across
  // the steps of a launch
  const int nsteps = a.n_steps > 1 ? a.n_steps : 1;
  for (int step = 0; step < nsteps; ++step) {
  const int pos = step > 0 ? __ldcg(a.pos) : pos0;
  const int lbase = step * a.L;
  // ---- embedding (CTA 0 writes the residual stream); first QKV rows + LN1 params requested meanwhile
  GemvDesc cur = make_desc(a, sl, 0, 0);
  Pre pre;
  issue_slabs(cur, reg0, &cbar[0]);
  prefetch_phase(cur, pre, pool, &wbar[warp], gw, warp, lane);
  if (blockIdx.x == 0) {
    for (int i = threadIdx.x; i < D; i += MT) {
      const int tok = step > 0 ? __ldcg(a.tokens + pos) : a.tokens[pos];
      a.dx[i] = __bfloat162float(a.embed[(long long)tok * D + i]) + a.dec_pos[(long long)pos * D + i];
    }
  }
  bar.sync();

  for (int l = 0; l < a.L; ++l) {
    const MegaLayer& L = sl[l];
    for (int gi = 0; gi < 4; ++gi) {
      // slab phases of a layer: 0 QKV | 1 self out-proj slice | 2 cross-q | 3 cross out-proj slice | 4 fc1 | 5 fc2
      const int sp = gi < 2 ? 2 * gi : gi + 2;
      const bool even = (sp & 1) == 0;
      // ---------------- GEMV phase ----------------
      {
        const int n = cur.n0 + warp * cur.R;
        const bool active = n < cur.nend;
        float res = 0.f;
        {
          const int m = lane & 7, r_sel = lane >> 3;
          if (cur.residual && active && r_sel < cur.R && m == 0 && n + r_sel < cur.nend) res = __ldcg(cur.residual + n + r_sel);
        }
        auto ahead = [&]() {
          // smem copies that can be requested now (the target region is free) and DRAM -> L2 prefetches of what comes later:
          //   QKV:     slice 1 -> region 1;  L2: cross-q, cross out-proj slice, this layer's cross K/V item
          //   cross-q: slice 3 -> region 1;  L2: fc1, fc2
          //   fc1:     fc2 -> region 1;      L2: next layer's QKV and self out-proj slice (or the LM head's first rows)
          //   fc2:     next QKV -> region 0
          // (cross-q and fc1 go to region 0, which the attention scratch overlays: they are requested after the attention)
          if (gi == 0) {
            issue_slice(make_slice(a, a.wo_hm[l]), reg1, &cbar[1]);
            l2_prefetch_phase(make_desc(a, sl, l, 1));
            l2_prefetch_slice(make_slice(a, a.xwo_hm[l]));
            if (threadIdx.x == DMA_T + 1 && item_cta) {
              const int s0 = ij * ks;
              const int nk = max(0, min(a.S, s0 + ks) - s0);
              if (nk > 0) {
                l2_prefetch(L.cross_k + ((long long)ih * a.S + s0) * 64, (uint32_t)nk * 128);
                l2_prefetch(L.cross_v + ((long long)ih * a.S + s0) * 64, (uint32_t)nk * 128);
              }
            }
          } else if (gi == 1) {
            issue_slice(make_slice(a, a.xwo_hm[l]), reg1, &cbar[1]);
            l2_prefetch_phase(make_desc(a, sl, l, 2));
            l2_prefetch_phase(make_desc(a, sl, l, 3));
          } else if (gi == 2) {
            issue_slabs(make_desc(a, sl, l, 3), reg1, &cbar[1]);
            if (l + 1 < a.L) {
              l2_prefetch_phase(make_desc(a, sl, l + 1, 0));
              l2_prefetch_slice(make_slice(a, a.wo_hm[l + 1]));
            }
          } else {
            if (l + 1 < a.L) issue_slabs(make_desc(a, sl, l + 1, 0), reg0, &cbar[0]);
          }
        };
        stage_x(xs, red, cur, pre, ahead);
        if (active) {
          mbar_wait(&cbar[even ? 0 : 1], even ? cpar0 : cpar1);
          const uint8_t* slab = (even ? reg0 : reg1) + (size_t)warp * cur.R * cur.K * 2;
          float acc[3][MB];
          if (cur.R == 3) dot_rows<MB, 3>(slab, xs, cur.K, acc, lane);
          else if (cur.R == 2) dot_rows<MB, 2>(slab, xs, cur.K, acc, lane);
          else dot_rows<MB, 1>(slab, xs, cur.K, acc, lane);
          finish_rows(cur, acc, pre.bias, n, res, D, a.Tmax, pos, lane);
        }
        if (cur.n0 < cur.nend) {  // (uniform per CTA: the phase's copy was issued iff the CTA owns rows)
          if (even) cpar0 ^= 1u;
          else cpar1 ^= 1u;
        }
      }
      __syncthreads();  // every warp is done with its slab and with xs

      if (gi >= 2) {
        // ---------------- fc1 -> fc2 -> next layer: grid barrier; the next phase's bias / LayerNorm slice requested meanwhile
        bar.arrive();
        cur = (gi == 2) ? make_desc(a, sl, l, 3) : make_desc(a, sl, l + 1 < a.L ? l + 1 : a.L, 0);
        prefetch_phase(cur, pre, pool, &wbar[warp], gw, warp, lane);
        bar.wait();
        continue;
      }

      // ---------------- QKV / cross-q done: signal the heads this CTA's rows belong to, then the fused attention phase
      head_signal(a.bar + (gi == 0 ? CNT_QKV : CNT_XQ), cur.n0, cur.nend, D);
      cur = make_desc(a, sl, l, gi == 0 ? 1 : 2);
      prefetch_phase(cur, pre, pool, &wbar[warp], gw, warp, lane);
      const SliceDesc sd = make_slice(a, gi == 0 ? a.wo_hm[l] : a.xwo_hm[l]);

      if (gi == 0) {
        // ===== self-attention of head ih (all nsplit CTAs of the head compute it) + row slice ij of the out-projection =====
        if (item_cta) {
          uint8_t* sK = att;
          uint8_t* sV = att + (size_t)MAXKEYS * 128;
          float* redo = reinterpret_cast<float*>(att + (size_t)MAXKEYS * 256);  // [MW][72]
          // rows < pos do not depend on this step: request them before the wait
          for (int s = grp; s < pos; s += KG) {
            const long long off = (long long)s * D + ih * 64 + sub * 8;
            cp_async16m(sK + s * 128 + sub * 16, L.self_k + off);
            cp_async16m(sV + s * 128 + sub * 16, L.self_v + off);
          }
          counter_wait(a.bar + CNT_QKV + ih, (unsigned)(lbase + l + 1) * head_expected(ih, D, 3, (3 * D + (int)gridDim.x - 1) / (int)gridDim.x));
          const int n = pos + 1;
          for (int s = pos + grp; s < n; s += KG) {
            const long long off = (long long)s * D + ih * 64 + sub * 8;
            cp_async16m(sK + s * 128 + sub * 16, L.self_k + off);
            cp_async16m(sV + s * 128 + sub * 16, L.self_v + off);
          }
          float qv[8];
          {
            const float4 q0 = __ldcg(reinterpret_cast<const float4*>(a.dqkv + ih * 64 + sub * 8));
            const float4 q1 = __ldcg(reinterpret_cast<const float4*>(a.dqkv + ih * 64 + sub * 8 + 4));
            qv[0] = q0.x; qv[1] = q0.y; qv[2] = q0.z; qv[3] = q0.w; qv[4] = q1.x; qv[5] = q1.y; qv[6] = q1.z; qv[7] = q1.w;
          }
          cp_async_wait_allm();
          __syncthreads();
          float mx, sum, ov;
          if (n <= 3 * KG) attend_smem<3>(sK, sV, redo, red, qv, n, nullptr, mx, sum, ov);
          else attend_smem<(MAXKEYS + KG - 1) / KG>(sK, sV, redo, red, qv, n, nullptr, mx, sum, ov);
          if (threadIdx.x < 64) a_s[threadIdx.x] = ov / sum;
          __syncthreads();
          if (sd.rows > 0) {
            mbar_wait(&cbar[1], cpar1);
            outproj_slice(reg1, a_s, sd, a.acc_a, warp, lane);
          }
          fence_proxy_async_smem();  // scratch writes (generic proxy) before later TMA writes to the same bytes
        }
        if (sd.rows > 0) cpar1 ^= 1u;
      } else {
        // ===== cross-attention: key split ij of head ih; the head's splits exchange their partials; row slice ij of the
        //       cross out-projection =====
        if (item_cta) {
          uint8_t* sK = att;
          uint8_t* sV = att + XKMAX * 128;
          float* redo = reinterpret_cast<float*>(att + 2 * XKMAX * 128);  // [MW][72]
          const int s0 = ij * ks;
          const int n = max(0, min(a.S, s0 + ks) - s0);
          if (threadIdx.x == 0) {  // the encoder K/V slice is constant during decoding: one bulk copy each, before the wait
            mbar_arrive_expect_tx(&xbar, (uint32_t)n * 256);
            if (n > 0) {
              bulk_g2s(sK, L.cross_k + ((long long)ih * a.S + s0) * 64, (uint32_t)n * 128, &xbar);
              bulk_g2s(sV, L.cross_v + ((long long)ih * a.S + s0) * 64, (uint32_t)n * 128, &xbar);
            }
          }
          counter_wait(a.bar + CNT_XQ + ih, (unsigned)(lbase + l + 1) * head_expected(ih, D, 1, (D + (int)gridDim.x - 1) / (int)gridDim.x));
          float qv[8];
          {
            const float4 q0 = __ldcg(reinterpret_cast<const float4*>(a.dq + ih * 64 + sub * 8));
            const float4 q1 = __ldcg(reinterpret_cast<const float4*>(a.dq + ih * 64 + sub * 8 + 4));
            qv[0] = q0.x; qv[1] = q0.y; qv[2] = q0.z; qv[3] = q0.w; qv[4] = q1.x; qv[5] = q1.y; qv[6] = q1.z; qv[7] = q1.w;
          }
          float* align_row = nullptr;
          if (a.align && L.head_slots) {
            const int slot = L.head_slots[ih];
            const int step = pos - a.step_base;
            if (slot >= 0 && step >= 0 && step < a.Tcap) align_row = a.align + ((long long)slot * a.Tcap + step) * a.S + s0;
          }
          mbar_wait(&xbar, xpar);
          xpar ^= 1u;
          float mx, sum, ov;
          attend_smem<(XKMAX + KG - 1) / KG>(sK, sV, redo, red, qv, n, align_row, mx, sum, ov);
          const long long pb = (long long)ih * nsplit + ij;
          if (threadIdx.x < 64) a.part_o[pb * 64 + threadIdx.x] = ov;
          if (threadIdx.x == 0) {
            a.part_ml[pb * 2 + 0] = mx;
            a.part_ml[pb * 2 + 1] = sum;
          }
          __syncthreads();
          // exchange among the nsplit CTAs of this head: everybody arrives, everybody waits, everybody merges
          if (threadIdx.x == MT - 1) red_release_add(a.bar + CNT_XHEAD + ih, 1u);
          counter_wait(a.bar + CNT_XHEAD + ih, (unsigned)(lbase + l + 1) * (unsigned)nsplit);
          if (threadIdx.x < 64) {
            const long long hb = (long long)ih * nsplit;
            float pm[XSPLIT], pl[XSPLIT], po[XSPLIT];
#pragma unroll
            for (int sp2 = 0; sp2 < XSPLIT; ++sp2) {
              pm[sp2] = -INFINITY; pl[sp2] = 0.f; po[sp2] = 0.f;
              if (sp2 < nsplit) {
                pm[sp2] = __ldcg(&a.part_ml[(hb + sp2) * 2]);
                pl[sp2] = __ldcg(&a.part_ml[(hb + sp2) * 2 + 1]);
                po[sp2] = __ldcg(&a.part_o[(hb + sp2) * 64 + threadIdx.x]);
              }
            }
            float M = -INFINITY;
#pragma unroll
            for (int sp2 = 0; sp2 < XSPLIT; ++sp2)
              if (pl[sp2] > 0.f) M = fmaxf(M, pm[sp2]);
            float Lsum = 0.f, o = 0.f;
#pragma unroll
            for (int sp2 = 0; sp2 < XSPLIT; ++sp2) {
              if (pl[sp2] > 0.f) {
                const float w = __expf(pm[sp2] - M);
                Lsum = fmaf(pl[sp2], w, Lsum);
                o = fmaf(po[sp2], w, o);
              }
            }
            a_s[threadIdx.x] = o / Lsum;
          }
          __syncthreads();
          if (sd.rows > 0) {
            mbar_wait(&cbar[1], cpar1);
            outproj_slice(reg1, a_s, sd, a.acc_b, warp, lane);
          }
          fence_proxy_async_smem();
        }
        if (sd.rows > 0) cpar1 ^= 1u;
      }
      __syncthreads();  // the attention scratch (which overlays region 0) is free: the next GEMV phase's slab can come
      issue_slabs(cur, reg0, &cbar[0]);
      bar.arrive();
      bar.wait();
    }
  }

  // ---------------- final LayerNorm + tied LM head: row pairs, two slab stages per warp ----------------
  stage_x(xs, red, cur, pre, [] {});
  unsigned long long best = 0ull;  // of the logits this lane finished: (order-preserving value bits << 32) | ~token
  {
    const int K = cur.K, N = cur.N;
    const size_t slab_bytes = (size_t)2 * K * 2;
    const size_t set_bytes = slab_bytes * MW;
    int buf = 0;
    const bool at_begin = (pos + 1 == a.begin_index) && a.begin_suppress_bits;
    for (int n = gw * 2; n < N; n += GW * 2) {
      const int n2 = n + GW * 2;
      if (n2 < N) {
        __syncwarp();  // every lane is done reading the stage that is refilled now
        issue_rows(pool + (buf ^ 1) * set_bytes + (size_t)warp * slab_bytes, &wbar[(buf ^ 1) * MW + warp], cur.W, K, 2, n2, N, lane);
      }
      if (buf == 0) {
        mbar_wait(&wbar[warp], wpar);
        wpar ^= 1u;
      } else {
        mbar_wait(&wbar[MW + warp], wpar1);
        wpar1 ^= 1u;
      }
      float acc[3][MB];
      dot_rows<MB, 2>(pool + buf * set_bytes + (size_t)warp * slab_bytes, xs, K, acc, lane);
      finish_rows(cur, acc, 0.f, n, 0.f, D, a.Tmax, pos, lane);
      if (a.fuse_select) {
        const int m = lane & 7, r_sel = lane >> 3, nn = n + r_sel;
        if (r_sel < 2 && m == 0 && nn < N) {
          float v = 0.f;
#pragma unroll
          for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int mm = 0; mm < MB; ++mm)
              if (r == r_sel && mm == m) v = acc[r][mm];
          bool masked = (a.suppress_bits[nn >> 5] >> (nn & 31)) & 1u;
          if (at_begin) masked = masked || ((a.begin_suppress_bits[nn >> 5] >> (nn & 31)) & 1u);
          if (!masked) {
            unsigned u = __float_as_uint(v);
            u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
            const unsigned long long key = ((unsigned long long)u << 32) | (unsigned long long)(0xffffffffu - (unsigned)nn);
            best = key > best ? key : best;
          }
        }
      }
      buf ^= 1;
    }
  }
  if (a.fuse_select) {
    // lanes m and 8 + m hold sequence m's candidates; fold per warp, per CTA, then one atomicMax per CTA and sequence
    {
      const unsigned long long o = __shfl_xor_sync(0xffffffffu, best, 8);
      best = o > best ? o : best;
    }
    unsigned long long* sb = reinterpret_cast<unsigned long long*>(xs);  // x is no longer needed: [MB][MW]
    __syncthreads();
    if (lane < MB) sb[lane * MW + warp] = best;
    __syncthreads();
    if (threadIdx.x < MB) {
      unsigned long long b = 0ull;
      for (int w = 0; w < MW; ++w) b = sb[threadIdx.x * MW + w] > b ? sb[threadIdx.x * MW + w] : b;
      if (b) atomicMax(a.sel_best + threadIdx.x, b);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned prev = atom_acq_rel_add(a.sel_ctr, 1u);
      if (prev == gridDim.x - 1) {  // every CTA's maxima are in: this is SelectArgs' greedy branch (decode.cu select_kernel)
        const int cur_len = pos + 1;
        const bool generating = cur_len >= a.begin_index && cur_len < a.Tmax;
        for (int q = 0; q < 1; ++q) {
          const unsigned long long b = __ldcg(a.sel_best + q);
          a.sel_best[q] = 0ull;
          if (generating) {
            int choice = (int)(0xffffffffu - (unsigned)(b & 0xffffffffull));
            if (a.finished[q]) choice = a.pad;
            else if (choice == a.eos) a.finished[q] = 1;
            a.tokens_rw[q * a.Tmax + cur_len] = choice;
          }
        }
        *a.sel_ctr = 0u;
        *a.pos_rw = pos + 1;
      }
    }
  }
  if (step + 1 < nsteps) bar.sync();  // the token and the position of the next step are in global memory
  }  // step
}

}  // namespace

// Launches the third-generation persistent step kernel on `st`.  Returns -3 when the configuration is outside what it
// supports (more than one sequence, no head-major out-projection copies bound, no fused greedy selection needed is fine).
int launch_decode_mega3(cudaStream_t st, const MegaArgs& a, int num_sms) {
  if (a.Q != 1 || a.trace || !a.dx2 || !a.acc_a || !a.acc_b) return -3;
  if (a.n_steps > 1 && !a.fuse_select) return -3;
  for (int l = 0; l < a.L; ++l)
    if (!a.wo_hm[l] || !a.xwo_hm[l]) return -3;
  if (a.L > MEGA_MAXL || a.D > MAXD || a.ffn > 5120 || a.D % 8 != 0 || a.ffn % 8 != 0 || a.Tmax > MAXKEYS || a.H > 32) return -3;
  if ((size_t)MW * a.D * 2 > (size_t)ATT_OFF) return -3;
  {
    const int nmax = 3 * a.D > a.ffn ? 3 * a.D : a.ffn;
    if (((nmax + num_sms - 1) / num_sms + MW - 1) / MW > 3) return -3;
  }
  if (a.nsplit > XSPLIT || a.H * a.nsplit > num_sms) return -3;
  {
    const int rs = (a.D + a.nsplit - 1) / a.nsplit;
    if (rs > RSMAX || (size_t)rs * 128 > (size_t)ATT_OFF) return -3;  // a slice: 16 rows per warp, below the attention scratch
  }
  MegaArgs b = a;
  const size_t smem = mega_smem_plan(1, a.D, a.ffn, num_sms, true, &b.p0_off);
  if (b.p0_off == 0 || smem + 8 * 1024 > 227 * 1024) return -3;  // needs the two slab regions
  if ((size_t)((a.D + num_sms - 1) / num_sms) * a.D * 2 > (size_t)ATT_OFF) return -3;
  const int ks = (a.S + a.nsplit - 1) / a.nsplit;
  if (ks > XKMAX) return -3;
  BW_CUDA_OK(cudaMemsetAsync(a.bar, 0, 1024 * sizeof(unsigned), st));
  static size_t attr = 0;
  if (smem > attr) {
    BW_CUDA_OK(cudaFuncSetAttribute(decode_mega3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  {  // cooperative launch: the whole grid is co-resident or the launch fails (see launch_decode_mega)
    if (g_mega_coop < 0) {
      const char* ev = getenv("BW_MEGA_COOP");
      g_mega_coop = (ev && ev[0] == '0') ? 0 : 1;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(num_sms); cfg.blockDim = dim3(MT); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeCooperative; at[0].val.cooperative = 1;
    cfg.attrs = at; cfg.numAttrs = g_mega_coop ? 1 : 0;
    BW_CUDA_OK(cudaLaunchKernelEx(&cfg, decode_mega3_kernel, b));
  }
  BW_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace bw
