// One persistent kernel per decoder step (q_len = 1, one beam per audio, up to 2 sequences).
//
// Why: with one kernel per op the step is 259 launches and every op pays its own chain of dependent global round
// trips; a step that should take 0.32 ms (2.07 GB at the measured 6.5 TB/s) took 1.7 ms
// (profiles/r1_v1_launches_summary.md).  Here the whole step -- embedding, 32 x (LN1+QKV, self-attention, out-proj,
// LN2+cross-q, cross-attention, out-proj, LN3+fc1+GELU, fc2), final LN + tied LM head -- runs in ONE kernel of one CTA
// per SM, phases separated by a grid barrier.
//
// The cost model that shaped it (ncu of the first version, profiles/r1_v2_*): 60% of all stall samples were warps parked
// at the barrier -- a phase is only as fast as its chain of *dependent* L2/DRAM round trips (~0.6-1 us each), not its
// bytes.  So everything that does not depend on the previous phase is requested BEFORE the barrier that precedes a phase:
//   * the phase's weight rows (cp.async into a per-warp smem slab: all rows of the phase are in flight at once, at no
//     register cost -- the register-prefetch version spilled them, and every spill store waited for its load, turning
//     the one round trip into several: 1.4 TB/s effective) and bias values,
//   * its LayerNorm gamma/beta and the attention K/V rows that are already final (cp.async into smem),
//   * the barrier itself is one red.release + ld.acquire polling loop (no membar.sc / L1 invalidation; activations that
//     cross CTAs are read with ld.global.cg),
//   * LayerNorm statistics are computed redundantly by every warp from the staged row (no block reductions), and the
//     normalisation is applied on the fly inside the dot product.
// After a barrier only the x row (and the residual values of the rows a warp owns) have to be fetched.
//
// Work split: 12 warps per CTA, global warp id gw; a GEMV phase gives warp gw the rows {gw*R + i*GW*R + r}; attention
// phases hand (sequence, head[, key split]) items to CTAs round-robin.  Token selection stays a separate small kernel.
#include <math.h>

#include "decode.cuh"
#include "kernels.h"

namespace bw {

namespace {

constexpr int MT = 384;        // threads per CTA (12 warps: <= 170 registers per thread)
constexpr int MW = MT / 32;    // warps per CTA
constexpr int KG = MT / 8;     // key groups of 8 lanes in the attention phases
constexpr int MAXKEYS = 448;   // self-attention keys held in smem (Tmax)
constexpr int XKMAX = 256;     // cross-attention keys per work item held in smem
constexpr int MAXD = 1280;
constexpr int ATT_OFF = 32 * 1024;  // attention scratch starts here inside the pool (above the R=1 weight slabs)

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
__device__ __forceinline__ void red_release_add(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// grid barrier: monotonically increasing arrival counter (zeroed by a memset node before the kernel).  bar.sync orders
// the CTA's writes before thread 0's release; the acquire poll + bar.sync orders the other CTAs' writes before our reads.
__device__ __forceinline__ long long global_ns() {
  long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

struct GridBar {
  unsigned* ctr;
  unsigned nblocks;
  unsigned epoch;
  long long* trace;  // optional [nblocks][2*MEGA_TRACE_N]: arrival / release time of every barrier (BW_MEGA_TRACE=1)
  __device__ __forceinline__ void sync() {
    __syncthreads();
    if (threadIdx.x == 0) {
      if (trace && epoch < MEGA_TRACE_N) trace[((long long)blockIdx.x * MEGA_TRACE_N + epoch) * 2] = global_ns();
      ++epoch;
      red_release_add(ctr, 1u);
      const unsigned target = epoch * nblocks;
      if (ld_acquire_u32(ctr) < target) {
        const long long t0 = clock64();
        while (ld_acquire_u32(ctr) < target) {
          if (clock64() - t0 > (1ll << 32)) {
            printf("[bw] decode_mega: grid barrier %u timed out (block %d)\n", epoch, blockIdx.x);
            __trap();
          }
        }
      }
      if (trace && epoch <= MEGA_TRACE_N) trace[((long long)blockIdx.x * MEGA_TRACE_N + epoch - 1) * 2 + 1] = global_ns();
    }
    __syncthreads();
  }
};

// Per-warp weight slab in shared memory: R rows of K bf16, lane l owns the 16-byte pieces at element offsets l*8 + i*256
// of each row (it copies them with cp.async and later reads exactly those back, so no CTA sync is needed for the data).
template <int R>
struct WB {
  float bias[R];
};

template <int NC, int R>
__device__ __forceinline__ void slab_load(uint8_t* slab, WB<R>& wb, const bf16* __restrict__ W, const float* __restrict__ bias, int K, int n,
                                          int N, int lane) {
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int row = min(n + r, N - 1);
    const bf16* wp = W + (long long)row * K;
#pragma unroll
    for (int i = 0; i < NC; ++i) {
      const int k = lane * 8 + i * 256;
      if (k < K) cp_async16m(slab + ((size_t)r * K + k) * 2, wp + k);
    }
    wb.bias[r] = bias ? bias[row] : 0.f;
  }
}

template <int NC, int R>
__device__ __forceinline__ void prefetch_rows(uint8_t* pool, WB<R>& wb, const bf16* W, const float* bias, int N, int K, int gw, int warp,
                                              int lane) {
  if (gw * R < N) slab_load<NC, R>(pool + (size_t)warp * R * K * 2, wb, W, bias, K, gw * R, N, lane);
}

// DRAM -> L2 prefetch of whole rows, one layer ahead of their use (no register / smem cost).  Measured on the barrier
// timeline: a slab requested just before its phase's barrier arrives ~5 us later, i.e. the phase stalls 3-4 us on DRAM;
// with the row already in L2 the cp.async that fills the slab is an L2 hit.
__device__ __forceinline__ void l2_prefetch(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
template <int R>
__device__ __forceinline__ void l2_prefetch_rows(const bf16* W, int N, int K, int gw, int GW, int lane) {
  if (lane < R) {
    const int row = gw * R + lane;
    if (row < N) l2_prefetch(W + (long long)row * K, (uint32_t)K * 2);
  }
}
// one warp-instruction per 4 KB: lane l prefetches the 128-byte line l of the chunk (plain LSU prefetch, no TMA op)
__device__ __forceinline__ void l2_line_prefetch_rows(const bf16* W, int N, int K, int R, int gw, int lane) {
  for (int r = 0; r < R; ++r) {
    const int row = gw * R + r;
    if (row >= N) break;
    const char* base = reinterpret_cast<const char*>(W + (long long)row * K);
    for (int off = lane * 128; off < K * 2; off += 4096) asm volatile("prefetch.global.L2 [%0];" ::"l"(base + off) : "memory");
  }
}

__device__ __forceinline__ void l2_prefetch_layer(const MegaLayer& L, const MegaArgs& a, int gw, int GW, int lane) {
  l2_prefetch_rows<3>(L.wqkv, 3 * a.D, a.D, gw, GW, lane);
  l2_prefetch_rows<1>(L.wo, a.D, a.D, gw, GW, lane);
  l2_prefetch_rows<1>(L.xwq, a.D, a.D, gw, GW, lane);
  l2_prefetch_rows<1>(L.xwo, a.D, a.D, gw, GW, lane);
  l2_prefetch_rows<3>(L.w1, a.ffn, a.D, gw, GW, lane);
  l2_prefetch_rows<1>(L.w2, a.D, a.ffn, gw, GW, lane);
  if (threadIdx.x == 0 && blockIdx.x < a.Q * a.H * a.nsplit) {  // this CTA's first cross-attention item
    const int item = blockIdx.x, nsplit = a.nsplit;
    const int ks = (a.S + nsplit - 1) / nsplit;
    const int split = item % nsplit, h = (item / nsplit) % a.H, q = item / (nsplit * a.H);
    const int s0 = split * ks;
    const int n = max(0, min(a.S, s0 + ks) - s0);
    if (n > 0) {
      l2_prefetch(L.cross_k + (((long long)q * a.H + h) * a.S + s0) * 64, (uint32_t)n * 128);
      l2_prefetch(L.cross_v + (((long long)q * a.H + h) * a.S + s0) * 64, (uint32_t)n * 128);
    }
  }
}

struct PhaseOut {
  float alpha;
  int alpha_cols;
  int act;
  const float* residual;  // may alias out
  float* out;
  int ldo;
  bf16* kc;  // optional KV scatter (fused QKV)
  bf16* vc;
  int D, Tmax, pos;
};

// LayerNorm gamma/beta slice of this thread (elements [4*tid, 4*tid+4), K <= 4*MT), requested before the barrier
struct GB {
  float4 g, b;
};
__device__ __forceinline__ void prefetch_gb(GB& gb, const float* __restrict__ g, const float* __restrict__ b, int D) {
  const int i = threadIdx.x * 4;
  if (i < D) {
    gb.g = *reinterpret_cast<const float4*>(g + i);
    gb.b = *reinterpret_cast<const float4*>(b + i);
  }
}

// stage M rows of K floats into smem (ld.global.cg).  With `ln` the rows are LayerNormed in place: every warp derives
// mean / rstd on its own from the staged row (two passes over smem, no block reduction), then each thread rewrites
// its 4-element slice.  (Normalising on the fly inside every warp's dot product tripled the smem traffic of a phase;
// the GEMV phases are bound by the 128 B/clk shared-memory port, not by HBM -- see the barrier timeline in profiles/.)
template <int MB>
__device__ __forceinline__ void stage_x(float* xs, const float* __restrict__ src, int ld, int K, int M, const GB* gb,
                                        long long* mk = nullptr) {
  const int lane = threadIdx.x & 31;
  for (int i = threadIdx.x * 4; i < MB * K; i += MT * 4) {
    const int m = i / K, k = i - m * K;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (m < M) v = __ldcg(reinterpret_cast<const float4*>(src + (long long)m * ld + k));
    *reinterpret_cast<float4*>(xs + i) = v;
  }
  if (mk && threadIdx.x == 0) mk[0] = global_ns();  // x loads returned
  cp_async_wait_allm();                             // this thread's slab pieces landed
  if (mk && threadIdx.x == 0) mk[1] = global_ns();
  __syncthreads();
  if (mk && threadIdx.x == 0) mk[2] = global_ns();  // whole CTA staged
  if (!gb) return;
  float mean[MB], rstd[MB];
#pragma unroll
  for (int m = 0; m < MB; ++m) {
    float s = 0.f;
    for (int k = lane * 4; k < K; k += 128) {
      const float4 v = *reinterpret_cast<const float4*>(xs + m * K + k);
      s += (v.x + v.y) + (v.z + v.w);
    }
    const float mu = warp_sum(s) / (float)K;
    float ss = 0.f;
    for (int k = lane * 4; k < K; k += 128) {
      const float4 v = *reinterpret_cast<const float4*>(xs + m * K + k);
      const float a0 = v.x - mu, a1 = v.y - mu, a2 = v.z - mu, a3 = v.w - mu;
      ss += (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
    }
    mean[m] = mu;
    rstd[m] = rsqrtf(warp_sum(ss) / (float)K + 1e-5f);
  }
  __syncthreads();  // every warp has its statistics before anybody overwrites the raw row
  const int k = threadIdx.x * 4;
  if (k < K) {
#pragma unroll
    for (int m = 0; m < MB; ++m) {
      float4 v = *reinterpret_cast<float4*>(xs + m * K + k);
      v.x = (v.x - mean[m]) * rstd[m] * gb->g.x + gb->b.x;
      v.y = (v.y - mean[m]) * rstd[m] * gb->g.y + gb->b.y;
      v.z = (v.z - mean[m]) * rstd[m] * gb->g.z + gb->b.z;
      v.w = (v.w - mean[m]) * rstd[m] * gb->g.w + gb->b.w;
      *reinterpret_cast<float4*>(xs + m * K + k) = v;
    }
  }
  __syncthreads();
}

// rows {gw*R + i*GW*R + r}; the first pass already sits in this warp's slab (requested before the preceding barrier).
// PIPE (LM head): two slab sets, the next pass is requested while the current one is consumed.
template <int MB, int NC, int R, bool PIPE>
__device__ __forceinline__ void gemv_phase(uint8_t* pool, WB<R>& wb, const bf16* __restrict__ W, const float* __restrict__ bias, int N,
                                           int K, const float* xs, int M, const PhaseOut& o, float res0, int gw, int GW, int warp,
                                           int lane) {
  const size_t slab_bytes = (size_t)R * K * 2;
  const size_t set_bytes = slab_bytes * MW;
  bool first = true;
  int buf = 0;
  for (int n = gw * R; n < N; n += GW * R) {
    const uint8_t* slab = pool + (PIPE ? buf * set_bytes : 0) + (size_t)warp * slab_bytes;
    const int n2 = n + GW * R;
    const bool has_next = n2 < N;
    WB<R> nb;
    if (PIPE) {
      if (has_next) slab_load<NC, R>(pool + (buf ^ 1) * set_bytes + (size_t)warp * slab_bytes, nb, W, bias, K, n2, N, lane);
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 1;" ::: "memory");  // everything but the pass just requested has landed
      __syncwarp();  // lanes read pieces other lanes of the warp copied
    } else if (!first) {
      cp_async_wait_allm();
      __syncwarp();
    }
    float acc[R][MB];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int m = 0; m < MB; ++m) acc[r][m] = 0.f;
    // per 256-element chunk a lane takes elements [4*lane, +4) and [128 + 4*lane, +4): both the fp32 x reads (LDS.128) and
    // the bf16 weight reads (LDS.64) are contiguous across the warp, i.e. bank-conflict free
#pragma unroll
    for (int i = 0; i < NC; ++i) {
      const int k0 = i * 256 + lane * 4;
      if (k0 < K) {
        const bool hi = (k0 + 128) < K;
        float x0[MB][4], x1[MB][4];
#pragma unroll
        for (int m = 0; m < MB; ++m) {
          *reinterpret_cast<float4*>(x0[m]) = *reinterpret_cast<const float4*>(&xs[m * K + k0]);
          *reinterpret_cast<float4*>(x1[m]) = hi ? *reinterpret_cast<const float4*>(&xs[m * K + k0 + 128]) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const uint2 wa = *reinterpret_cast<const uint2*>(slab + ((size_t)r * K + k0) * 2);
          const uint2 wc = hi ? *reinterpret_cast<const uint2*>(slab + ((size_t)r * K + k0 + 128) * 2) : make_uint2(0u, 0u);
          const float2 a0 = unpack_bf16(wa.x), a1 = unpack_bf16(wa.y), c0 = unpack_bf16(wc.x), c1 = unpack_bf16(wc.y);
#pragma unroll
          for (int m = 0; m < MB; ++m) {
            float s = acc[r][m];
            s = fmaf(a0.x, x0[m][0], s); s = fmaf(a0.y, x0[m][1], s); s = fmaf(a1.x, x0[m][2], s); s = fmaf(a1.y, x0[m][3], s);
            s = fmaf(c0.x, x1[m][0], s); s = fmaf(c0.y, x1[m][1], s); s = fmaf(c1.x, x1[m][2], s); s = fmaf(c1.y, x1[m][3], s);
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
      float v = 0.f, bv = 0.f;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if (r == r_sel) bv = wb.bias[r];
#pragma unroll
        for (int mm = 0; mm < MB; ++mm)
          if (r == r_sel && mm == m) v = acc[r][mm];
      }
      v += bv;
      if (nn < o.alpha_cols) v *= o.alpha;
      if (o.act == 1) v = gelu_erf(v);
      if (o.residual) v += first ? res0 : __ldcg(o.residual + (long long)m * o.ldo + nn);
      o.out[(long long)m * o.ldo + nn] = v;
      if (o.kc && nn >= o.D) {
        const long long row = ((long long)m * o.Tmax + o.pos) * o.D;
        if (nn < 2 * o.D) o.kc[row + nn - o.D] = __float2bfloat16(v);
        else o.vc[row + nn - 2 * o.D] = __float2bfloat16(v);
      }
    }
    first = false;
    if (PIPE) {
      if (has_next) {
#pragma unroll
        for (int r = 0; r < R; ++r) wb.bias[r] = nb.bias[r];
      }
      buf ^= 1;
    } else if (has_next) {
      __syncwarp();
      slab_load<NC, R>(pool + (size_t)warp * slab_bytes, wb, W, bias, K, n2, N, lane);
    }
  }
}

// the residual value of the output element this lane will finish in the first pass (fetched right after the barrier so
// its L2 round trip overlaps the x staging)
template <int MB, int R>
__device__ __forceinline__ float fetch_residual(const float* residual, int ldo, int N, int M, int gw, int lane) {
  const int m = lane & 7, r_sel = lane >> 3;
  const int nn = gw * R + r_sel;
  if (residual && r_sel < R && m < MB && m < M && nn < N) return __ldcg(residual + (long long)m * ldo + nn);
  return 0.f;
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

// smem carve-up (dynamic): red [32] | xs [MB*ffn] | gb [2][2*MAXD] | attention scratch (K rows, V rows, scores, partial out)
template <int MB>
__global__ void __launch_bounds__(MT, 1) decode_mega_kernel(const __grid_constant__ MegaArgs a) {
  extern __shared__ __align__(16) uint8_t dyn[];
  float* red = reinterpret_cast<float*>(dyn);
  float* xs = red + 32;
  uint8_t* pool = reinterpret_cast<uint8_t*>(xs + (size_t)MB * a.ffn);  // weight slabs [0, ...) ; attention scratch from ATT_OFF
  uint8_t* att = pool + ATT_OFF;  // only R=1 slabs (<= 30 KB) are live while an attention phase runs
  __shared__ unsigned s_last;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gw = blockIdx.x * MW + warp, GW = gridDim.x * MW;
  const int D = a.D, H = a.H, Q = a.Q, ffn = a.ffn;
  const int pos = *a.pos;
  GridBar bar{a.bar, gridDim.x, 0u, a.trace};
  GB gb;  // gamma/beta slice of the next LayerNorm phase
  auto marks = [&]() -> long long* {  // 4 marks per barrier epoch, behind the arrive/release table
    return (a.trace && bar.epoch < MEGA_TRACE_N) ? a.trace + (long long)gridDim.x * MEGA_TRACE_N * 2 + ((long long)blockIdx.x * MEGA_TRACE_N + bar.epoch) * 4 : nullptr;
  };  // which gamma/beta buffer the next LN phase uses

  const int nsplit = a.nsplit;
  const int ks = (a.S + nsplit - 1) / nsplit;
  const int grp = threadIdx.x >> 3, sub = threadIdx.x & 7;  // KG key groups x 8 lanes

  // ---- phase 0: embedding (CTA 0 writes the residual stream); first QKV rows + LN1 params requested meanwhile
  WB<3> b3;  // QKV and fc1: 3 rows per warp (1776 warps x 3 >= 5120 rows: one pass)
  WB<1> b1;  // out-proj / cross-q / fc2: one row per warp
  WB<2> b2;  // LM head: pipelined row pairs
  if (((a.flags >> 1) & 3) == 1) l2_prefetch_layer(a.layers[0], a, gw, GW, lane);
  prefetch_rows<5, 3>(pool, b3, a.layers[0].wqkv, a.layers[0].bqkv, 3 * D, D, gw, warp, lane);
  prefetch_gb(gb, a.layers[0].ln1g, a.layers[0].ln1b, D);
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
    const int pfm = (a.flags >> 1) & 3;  // 0 none, 1 bulk prefetch of the whole next layer, 2 per-phase line prefetch
    if (pfm == 1 && l + 1 < a.L) l2_prefetch_layer(a.layers[l + 1], a, gw, GW, lane);  // DRAM -> L2, a whole layer ahead
    else if (pfm == 1 && lane < 8) {  // first LM-head passes of this warp
      const long long row = (long long)gw * 2 + (lane & 1) + (long long)(lane >> 1) * GW * 2;
      if (row < a.V) l2_prefetch(a.embed + row * D, (uint32_t)D * 2);
    }
    const MegaLayer& Ln = a.layers[(l + 1 < a.L) ? l + 1 : l];
    const bool pf2 = (pfm == 2);  // L2 line prefetch two GEMV phases ahead, issued AFTER this phase's own work
    // ---------------- A: LN1 + fused QKV (+ self-KV append) ----------------
    stage_x<MB>(xs, a.dx, D, D, Q, &gb, marks());
    {
      PhaseOut o{0.125f, D, 0, nullptr, a.dqkv, 3 * D, L.self_k, L.self_v, D, a.Tmax, pos};
      gemv_phase<MB, 5, 3, false>(pool, b3, L.wqkv, L.bqkv, 3 * D, D, xs, Q, o, 0.f, gw, GW, warp, lane);
      if (a.trace && threadIdx.x == 0 && bar.epoch < MEGA_TRACE_N) marks()[3] = global_ns();
    }
    __syncthreads();  // every warp is done with its slab: the pool can be re-carved for the next phases
    prefetch_rows<5, 1>(pool, b1, L.wo, L.bo, D, D, gw, warp, lane);
    if (pf2) l2_line_prefetch_rows(L.xwq, D, D, 1, gw, lane);
    // past K/V rows of this CTA's self-attention item do not depend on this step: request them now
    if (blockIdx.x < Q * H) {
      const int q = blockIdx.x / H, h = blockIdx.x - q * H;
      uint8_t* sK = att;
      uint8_t* sV = att + (size_t)MAXKEYS * 128;
      for (int s = grp; s < pos; s += KG) {
        const long long off = ((long long)q * a.Tmax + s) * D + h * 64 + sub * 8;
        cp_async16m(sK + s * 128 + sub * 16, L.self_k + off);
        cp_async16m(sV + s * 128 + sub * 16, L.self_v + off);
      }
    }
    bar.sync();
    // ---------------- B: causal self-attention, one (sequence, head) per CTA ----------------
    for (int item = blockIdx.x; item < Q * H; item += gridDim.x) {
      const int q = item / H, h = item - q * H;
      const int n = pos + 1;
      uint8_t* sK = att;
      uint8_t* sV = att + (size_t)MAXKEYS * 128;
      float* sc = reinterpret_cast<float*>(att + (size_t)MAXKEYS * 256);
      float* redo = sc + MAXKEYS;  // [KG][64]
      const int s_first = (item == blockIdx.x) ? pos : 0;  // rows < pos of the first item were prefetched
      for (int s = s_first + grp; s < n; s += KG) {
        const long long off = ((long long)q * a.Tmax + s) * D + h * 64 + sub * 8;
        cp_async16m(sK + s * 128 + sub * 16, L.self_k + off);
        cp_async16m(sV + s * 128 + sub * 16, L.self_v + off);
      }
      float qv[8];
      {
        const float4 q0 = __ldcg(reinterpret_cast<const float4*>(a.dqkv + (long long)q * 3 * D + h * 64 + sub * 8));
        const float4 q1 = __ldcg(reinterpret_cast<const float4*>(a.dqkv + (long long)q * 3 * D + h * 64 + sub * 8 + 4));
        qv[0] = q0.x; qv[1] = q0.y; qv[2] = q0.z; qv[3] = q0.w; qv[4] = q1.x; qv[5] = q1.y; qv[6] = q1.z; qv[7] = q1.w;
      }
      cp_async_wait_allm();
      __syncthreads();  // the row of position `pos` was copied by group 0, whatever group reads it below
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
    {
      const float res0 = fetch_residual<MB, 1>(a.dx, D, D, Q, gw, lane);
      stage_x<MB>(xs, a.dattn, D, D, Q, nullptr, marks());
      PhaseOut o{1.f, 0, 0, a.dx, a.dx, D, nullptr, nullptr, D, a.Tmax, pos};
      gemv_phase<MB, 5, 1, false>(pool, b1, L.wo, L.bo, D, D, xs, Q, o, res0, gw, GW, warp, lane);
      if (a.trace && threadIdx.x == 0 && bar.epoch < MEGA_TRACE_N) marks()[3] = global_ns();
    }
    __syncthreads();
    prefetch_rows<5, 1>(pool, b1, L.xwq, L.xbq, D, D, gw, warp, lane);
    if (pf2) l2_line_prefetch_rows(L.xwo, D, D, 1, gw, lane);
    prefetch_gb(gb, L.ln2g, L.ln2b, D);
    bar.sync();
    // ---------------- D: LN2 + cross q projection ----------------
    stage_x<MB>(xs, a.dx, D, D, Q, &gb, marks());
    {
      PhaseOut o{0.125f, D, 0, nullptr, a.dq, D, nullptr, nullptr, D, a.Tmax, pos};
      gemv_phase<MB, 5, 1, false>(pool, b1, L.xwq, L.xbq, D, D, xs, Q, o, 0.f, gw, GW, warp, lane);
      if (a.trace && threadIdx.x == 0 && bar.epoch < MEGA_TRACE_N) marks()[3] = global_ns();
    }
    __syncthreads();
    prefetch_rows<5, 1>(pool, b1, L.xwo, L.xbo, D, D, gw, warp, lane);
    if (pf2) l2_line_prefetch_rows(L.w1, ffn, D, 3, gw, lane);
    // the encoder K/V slice of this CTA's first cross-attention item is constant during decoding: request it now
    if (blockIdx.x < Q * H * nsplit) {
      const int item = blockIdx.x;
      const int split = item % nsplit, h = (item / nsplit) % H, q = item / (nsplit * H);
      const int s0 = split * ks;
      const int n = max(0, min(a.S, s0 + ks) - s0);
      const bf16* kbase = L.cross_k + (((long long)q * H + h) * a.S + s0) * 64 + sub * 8;
      const bf16* vbase = L.cross_v + (((long long)q * H + h) * a.S + s0) * 64 + sub * 8;
      uint8_t* sK = att;
      uint8_t* sV = att + XKMAX * 128;
      for (int kk = grp; kk < n; kk += KG) {
        cp_async16m(sK + kk * 128 + sub * 16, kbase + (long long)kk * 64);
        cp_async16m(sV + kk * 128 + sub * 16, vbase + (long long)kk * 64);
      }
    }
    bar.sync();
    // ---------------- E: cross-attention, (audio, head, key split) items; last split of a head merges ----------------
    {
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
        if (item != blockIdx.x) {  // later items of this CTA were not prefetched
          const bf16* kbase = L.cross_k + (((long long)q * H + h) * a.S + s0) * 64 + sub * 8;
          const bf16* vbase = L.cross_v + (((long long)q * H + h) * a.S + s0) * 64 + sub * 8;
          for (int kk = grp; kk < n; kk += KG) {
            cp_async16m(sK + kk * 128 + sub * 16, kbase + (long long)kk * 64);
            cp_async16m(sV + kk * 128 + sub * 16, vbase + (long long)kk * 64);
          }
        }
        float qv[8];
        {
          const float4 q0 = __ldcg(reinterpret_cast<const float4*>(a.dq + (long long)q * D + h * 64 + sub * 8));
          const float4 q1 = __ldcg(reinterpret_cast<const float4*>(a.dq + (long long)q * D + h * 64 + sub * 8 + 4));
          qv[0] = q0.x; qv[1] = q0.y; qv[2] = q0.z; qv[3] = q0.w; qv[4] = q1.x; qv[5] = q1.y; qv[6] = q1.z; qv[7] = q1.w;
        }
        float* align_row = nullptr;
        if (a.align && L.head_slots) {
          const int slot = L.head_slots[h];
          const int step = pos - a.step_base;
          if (slot >= 0 && step >= 0 && step < a.Tcap) align_row = a.align + (((long long)q * a.Ha + slot) * a.Tcap + step) * a.S + s0;
        }
        cp_async_wait_allm();  // (same thread -> piece map as the prefetch: each thread reads back its own copies)
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
        if (!(a.flags & 1)) {  // merge by the last-arriving split of this (sequence, head)
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
        }
        __syncthreads();
      }
    }
    bar.sync();
    // ---------------- F: cross out-proj + residual ----------------
    {
      const float res0 = fetch_residual<MB, 1>(a.dx, D, D, Q, gw, lane);
      // x = merged cross-attention output: every CTA merges the key-split partials itself (flash-decoding combine;
      // 3 * nsplit independent L2 loads per element instead of a fence + atomic + last-block chain in phase E)
      if (!(a.flags & 1)) stage_x<MB>(xs, a.dattn, D, D, Q, nullptr);
      else
      for (int i = threadIdx.x; i < MB * D; i += MT) {
        const int q = i / D, hd = i - q * D;
        float v = 0.f;
        if (q < Q) {
          const int h = hd >> 6, d = hd & 63;
          const long long hb = ((long long)q * H + h) * nsplit;
          float pm[XSPLIT], pl[XSPLIT], po[XSPLIT];
#pragma unroll
          for (int sp = 0; sp < XSPLIT; ++sp) {
            if (sp < nsplit) {
              pm[sp] = __ldcg(&a.part_ml[(hb + sp) * 2]);
              pl[sp] = __ldcg(&a.part_ml[(hb + sp) * 2 + 1]);
              po[sp] = __ldcg(&a.part_o[(hb + sp) * 64 + d]);
            }
          }
          float M = -INFINITY;
#pragma unroll
          for (int sp = 0; sp < XSPLIT; ++sp)
            if (sp < nsplit && pl[sp] > 0.f) M = fmaxf(M, pm[sp]);
          float Lsum = 0.f, ov = 0.f;
#pragma unroll
          for (int sp = 0; sp < XSPLIT; ++sp) {
            if (sp < nsplit && pl[sp] > 0.f) {
              const float w = __expf(pm[sp] - M);
              Lsum = fmaf(pl[sp], w, Lsum);
              ov = fmaf(po[sp], w, ov);
            }
          }
          v = ov / Lsum;
        }
        xs[i] = v;
      }
      cp_async_wait_allm();
      __syncthreads();
      PhaseOut o{1.f, 0, 0, a.dx, a.dx, D, nullptr, nullptr, D, a.Tmax, pos};
      gemv_phase<MB, 5, 1, false>(pool, b1, L.xwo, L.xbo, D, D, xs, Q, o, res0, gw, GW, warp, lane);
    }
    __syncthreads();
    prefetch_rows<5, 3>(pool, b3, L.w1, L.b1, ffn, D, gw, warp, lane);
    if (pf2) l2_line_prefetch_rows(L.w2, D, ffn, 1, gw, lane);
    prefetch_gb(gb, L.ln3g, L.ln3b, D);
    bar.sync();
    // ---------------- G: LN3 + fc1 + GELU ----------------
    stage_x<MB>(xs, a.dx, D, D, Q, &gb, marks());
    {
      PhaseOut o{1.f, 0, 1, nullptr, a.dh, ffn, nullptr, nullptr, D, a.Tmax, pos};
      gemv_phase<MB, 5, 3, false>(pool, b3, L.w1, L.b1, ffn, D, xs, Q, o, 0.f, gw, GW, warp, lane);
      if (a.trace && threadIdx.x == 0 && bar.epoch < MEGA_TRACE_N) marks()[3] = global_ns();
    }
    {
      // ---------------- H: fc2 + residual (K = ffn: one row per warp, 20 loads in flight) ----------------
      __syncthreads();
      prefetch_rows<20, 1>(pool, b1, L.w2, L.b2, D, ffn, gw, warp, lane);
      if (pf2 && l + 1 < a.L) l2_line_prefetch_rows(Ln.wqkv, 3 * D, D, 3, gw, lane);
      bar.sync();
      const float res0 = fetch_residual<MB, 1>(a.dx, D, D, Q, gw, lane);
      stage_x<MB>(xs, a.dh, ffn, ffn, Q, nullptr, marks());
      PhaseOut o{1.f, 0, 0, a.dx, a.dx, D, nullptr, nullptr, D, a.Tmax, pos};
      gemv_phase<MB, 20, 1, false>(pool, b1, L.w2, L.b2, D, ffn, xs, Q, o, res0, gw, GW, warp, lane);
      if (a.trace && threadIdx.x == 0 && bar.epoch < MEGA_TRACE_N) marks()[3] = global_ns();
    }
    __syncthreads();
    if (l + 1 < a.L) {
      prefetch_rows<5, 3>(pool, b3, a.layers[l + 1].wqkv, a.layers[l + 1].bqkv, 3 * D, D, gw, warp, lane);
      if (pf2) l2_line_prefetch_rows(Ln.wo, D, D, 1, gw, lane);
      prefetch_gb(gb, a.layers[l + 1].ln1g, a.layers[l + 1].ln1b, D);
    } else {
      prefetch_rows<5, 2>(pool, b2, a.embed, nullptr, a.V, D, gw, warp, lane);
      prefetch_gb(gb, a.lnf_g, a.lnf_b, D);
    }
    bar.sync();
  }
  // ---------------- final LayerNorm + tied LM head ----------------
  stage_x<MB>(xs, a.dx, D, D, Q, &gb);
  {
    PhaseOut o{1.f, 0, 0, nullptr, a.logits, a.V, nullptr, nullptr, D, a.Tmax, pos};
    gemv_phase<MB, 5, 2, true>(pool, b2, a.embed, nullptr, a.V, D, xs, Q, o, 0.f, gw, GW, warp, lane);
  }
}

size_t mega_smem_bytes(int mb, int ffn) {
  const size_t attn = (size_t)MAXKEYS * 256 + (size_t)(MAXKEYS + KG * 64) * sizeof(float);
  const size_t xattn = (size_t)2 * XKMAX * 128 + (size_t)(XKMAX + KG * 64) * sizeof(float);
  const size_t att = ATT_OFF + (attn > xattn ? attn : xattn);
  const size_t wts = (size_t)MW * 2 * 2 * MAXD * 2 > (size_t)MW * ffn * 2 ? (size_t)MW * 2 * 2 * MAXD * 2 : (size_t)MW * ffn * 2;  // LM head 2 sets of row pairs | fc2 rows
  const size_t w3 = (size_t)MW * 3 * MAXD * 2;
  size_t pool = att > wts ? att : wts;
  if (w3 > pool) pool = w3;
  return 32 * sizeof(float) + (size_t)mb * ffn * sizeof(float) + pool + 64;
}

}  // namespace

// Launches the persistent step kernel on `st`.  Returns -3 when the configuration is outside what it supports
// (the caller then uses the per-op path).
int launch_decode_mega(cudaStream_t st, const MegaArgs& a, int num_sms) {
  const int Q = a.Q;
  if (a.L > MEGA_MAXL || Q > 2 || a.D > MAXD || a.ffn > 5120 || a.D % 8 != 0 || a.ffn % 8 != 0 || a.Tmax > MAXKEYS) return -3;
  const int mb = Q <= 1 ? 1 : 2;
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
  }
#undef BW_MEGA_CASE
  BW_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace bw
