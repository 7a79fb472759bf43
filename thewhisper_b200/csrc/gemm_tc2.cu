// Second-generation encoder GEMM for sm_100a: CTA pairs (tcgen05 cta_group::2), persistent tile loop, two TMEM accumulator stages.
//     C[r, n] = epi( sum_k A[r, k] * W[n, k] ),  A [M, K] bf16 row-major, W [N, K] bf16 row-major (torch Linear layout)
//
// Why (profiles/r2a_*: ncu of gemm_tc_kernel<128>): a 128 x 128 tile per CTA pulls 32 KB from L2 per 64-wide k-block for 256 MMA
// clocks = 128 B/clk/SM, three times what the L2 can feed 148 SMs (~6.3 KB/clk chip-wide, B300_MICROARCH.md "LTS throughput
// cap"), so the first-generation kernel tops out near a third of the tensor peak whatever the pipeline depth; and its single
// accumulator serialises epilogue and main loop inside a CTA.  Here
//   * a CTA PAIR owns a 256 x BN tile: each CTA stages its own 128 rows of A and HALF of the W tile (BN / 2 rows); the leader's
//     tcgen05.mma.cta_group::2 reads both halves, so per SM and k-block 32 KB feed 128 x 256 x 64 MACs (64 B/clk/SM at BN = 256);
//   * CTAs are persistent (one pair per two SMs) and walk tiles n-fastest: the pairs running together work on a few 256-row bands of
//     A against ALL of W (<= 13 MB: L2-resident), so A streams from DRAM once (m-fastest re-read the 245 MB A of a 64-chunk batch once
//     per W tile: 3.7 GB of DRAM reads for one fc1, profiles/r2b_summary.md);
//   * accumulators are double-buffered in TMEM (2 x BN columns): the epilogue of tile i (bias / GELU / residual / bf16 pack, four
//     warps per CTA in round 2b, eight since r2j: two per TMEM lane quadrant, each on half of the tile's columns) runs under the main loop of
//     tile i + 1.
// Barrier protocol (CUTLASS sm100 2-SM pipeline, restated in raw PTX): TMA loads of BOTH CTAs complete on the LEADER's full[s]
// (cp.async.bulk.tensor ... .cta_group::2 with the peer bit of the mbarrier address cleared; the leader alone posts the expected
// byte count of both); tcgen05.commit ... multicast::cluster releases stage s in both CTAs and publishes a finished accumulator to
// both epilogues; the epilogue warps of both CTAs arrive remotely on the leader's accum_empty[as].
#include <limits.h>
#include <stdlib.h>

#include "kernels.h"

namespace BW_NS {

namespace {

constexpr int BM = 128;  // rows per CTA (the pair covers 256)
constexpr int BK = 64;
constexpr int A_STAGE_BYTES = BM * BK * 2;  // 16 KB
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;  // clears the CTA-rank bit of a shared::cluster address: rank 0's copy

// epilogue specialisations (the runtime-parameterised generic one executes every option's instructions predicated off: 25 per element)
enum { EPI_GENERIC = 0, EPI_GELU = 1 /* bias, GELU -> 16 bit */, EPI_RESID = 2 /* bias, + fp32 residual -> fp32 */, EPI_PLAIN = 3 /* bias -> 16 bit */ };

template <int BN, int NEW>
struct Cfg2 {
  static_assert(NEW == 8 || NEW == 16, "epilogue warps");
  static constexpr int B_STAGE_BYTES = (BN / 2) * BK * 2;  // this CTA's half of the W tile
  static constexpr int EPI_BYTES = NEW * 4096;              // one 32 x 32 fp32 patch per epilogue warp
  // ring depth: what fits beside the epilogue patches in 227 KB (64 K-elements per stage, 512 MMA clocks at BN = 256)
  static constexpr int STAGES = (BN == 256) ? (NEW == 16 ? 5 : 6) : (NEW == 16 ? 6 : 8);
  static constexpr int TMEM_COLS = 2 * BN;  // two accumulator stages (512 or 256 columns)
  static constexpr int THREADS = NEW * 32 + 128;  // epilogue warpgroups + {TMA warp, MMA warp, two idle warps}
  static constexpr int SMEM_BYTES = STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) + EPI_BYTES + 1024 /*align slack*/ + 512 /*barriers*/;
};

struct Gemm2Params {
  int M, N, K;
  int tiles_m2, tiles_n;
  int rows_per_item;  // epilogue address map: b = r / rows_per_item, t = r % rows_per_item
  GemmEpi epi;
};

// exact (erf) GELU to 4e-7 absolute: Phi(x) through Abramowitz-Stegun 7.1.26 (|erf error| <= 1.5e-7) -- 2 MUFU + 12 fp32 operations where erff
// costs ~25; the result is rounded to 16 bits right after (half an ulp there is >= 2.4e-4 relative).  tests/test_ops_gpu.py pins it to erf.
__device__ __forceinline__ float gelu_as(float x) {
  const float ax = fabsf(x);
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f * 0.70710678f, ax, 1.0f)));
  float q = 0.5f * 1.061405429f;
  q = fmaf(q, t, 0.5f * -1.453152027f);
  q = fmaf(q, t, 0.5f * 1.421413741f);
  q = fmaf(q, t, 0.5f * -0.284496736f);
  q = fmaf(q, t, 0.5f * 0.254829592f);
  q = (q * t) * ex2_approx((x * -0.72134752f) * x);  // 0.5 erfc(|x| / sqrt 2) = Phi(-|x|)
  return fmaf(-ax, q, fmaxf(x, 0.f));               // x >= 0: x - x q;  x < 0: x q
}

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// TMA load into THIS CTA's smem, completion bytes on the LEADER CTA's mbarrier (same offset, rank bit cleared)
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & PEER_MASK), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// all MMAs issued so far by this thread -> one arrival on the mbarrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
// arrive on rank 0's copy of `bar`.  Relaxed: what the arrival publishes is "my tcgen05.ld of this accumulator stage have completed" --
// tcgen05.wait::ld has already blocked on that -- not this warp's global stores; a release here compiled to MEMBAR + ERRBAR and held every
// epilogue warp until its output stores were acknowledged (8 % of all stall samples in r2l's ncu of the fc1 GEMM).
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & PEER_MASK) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}

template <int BN, int MODE, int NEW>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NEW * 32 + 128, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const Gemm2Params p) {
  using C = Cfg2<BN, NEW>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + C::STAGES * A_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + C::STAGES * C::B_STAGE_BYTES + C::EPI_BYTES);
  uint64_t* full = bars;                             // [STAGES]  (the leader's copies are the live ones)
  uint64_t* empty = bars + C::STAGES;                // [STAGES]
  uint64_t* accum_full = bars + 2 * C::STAGES;       // [2]
  uint64_t* accum_empty = bars + 2 * C::STAGES + 2;  // [2]       (leader's copies)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int nk = p.K / BK;
  const int ntiles = p.tiles_m2 * p.tiles_n;

  if (threadIdx.x == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&accum_full[s], 1);
      mbar_init(&accum_empty[s], 2 * NEW);  // the epilogue warps of both CTAs
    }
    fence_mbar_init();
  }
  if (warp == NEW && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
  }
  if (warp == NEW + 1) tmem_alloc2(tmem_slot, C::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer's barriers are initialised before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // programmatic dependent launch (the encoder runs as one CUDA graph of ~300 kernels): everything below reads what the predecessor wrote
  // (A, the residual) or overwrites what it read; the next kernel's CTAs may be queued once all of ours are past this point
  pdl_wait();
  pdl_launch();

  // NEW / 4 epilogue warpgroups and one of {TMA warp, MMA warp, two idle warps}; the last hands registers to the epilogue
  // (8 warps: two residual chunks in flight + the turned accumulator chunk want ~190; 16 warps: 640 threads leave 96 each, the epilogue takes 104.
  //  setmaxnreg moves registers inside the CTA's OWN allocation: what the epilogue threads gain must not exceed what the last warpgroup gives up
  //  -- 8 warps: 256 x (216 - 168) <= 128 x (168 - 56); 16 warps: 512 x (104 - 96) <= 128 x (96 - 48).  r2n asked for 112 (8192 > 6144), counting the
  //  SM's unallocated registers as available: setmaxnreg.inc then waits for ever, and no mbarrier timeout catches that.)
  if (warp >= NEW) {
  if (NEW == 8) asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
  else asm volatile("setmaxnreg.dec.sync.aligned.u32 48;");
  if (warp == NEW) {
    // ---------------- TMA producer (one lane, both CTAs) ----------------
    if (lane == 0) {
      int it = 0;
      for (int tile = pair; tile < ntiles; tile += npairs) {
        const int m2 = tile / p.tiles_n, nt = tile - m2 * p.tiles_n;
        const int row0 = m2 * 2 * BM + (int)rank * BM;
        const int wrow0 = nt * BN + (int)rank * (BN / 2);
        for (int kb = 0; kb < nk; ++kb, ++it) {
          const int s = it % C::STAGES;
          const uint32_t ph = (it / C::STAGES) & 1;
          mbar_wait(&empty[s], ph ^ 1);
          if (leader) mbar_arrive_expect_tx(&full[s], 2 * (A_STAGE_BYTES + C::B_STAGE_BYTES));
          tma_load_2d_pair(sA + s * A_STAGE_BYTES, &tmA, &full[s], kb * BK, row0);
          tma_load_2d_pair(sB + s * C::B_STAGE_BYTES, &tmW, &full[s], kb * BK, wrow0);
        }
      }
    }
  } else if (warp == NEW + 1) {
    // ---------------- MMA issuer (one lane of the leader CTA) ----------------
    if (leader && lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(2 * BM, BN);
      int it = 0, at = 0;
      for (int tile = pair; tile < ntiles; tile += npairs, ++at) {
        const int as = at & 1;
        const uint32_t aph = (at >> 1) & 1;
        mbar_wait(&accum_empty[as], aph ^ 1);  // both epilogues have drained this accumulator stage
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * BN);
        for (int kb = 0; kb < nk; ++kb, ++it) {
          const int s = it % C::STAGES;
          const uint32_t ph = (it / C::STAGES) & 1;
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint64_t a0 = umma_desc_sw128(smem_u32(sA + s * A_STAGE_BYTES));
          const uint64_t b0 = umma_desc_sw128(smem_u32(sB + s * C::B_STAGE_BYTES));
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) umma_bf16_pair(tmem_d, a0 + 2 * k, b0 + 2 * k, idesc, (uint32_t)((kb | k) != 0));
          umma_commit_pair(&empty[s]);
        }
        umma_commit_pair(&accum_full[as]);
      }
    }
  }
  } else {
    if (NEW == 8) asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
    else asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    // ---------------- epilogue: NEW warps; warp w owns TMEM lanes [32 (w & 3), +32) = rows of this CTA's half tile, and the column slice w >> 2 ----------------
    // (16 warps for the GELU / plain 16-bit epilogues: with 8, one warp's pass over its 4 chunks took about as long as the MMAs of a K = 1280 tile,
    //  and with two accumulator stages the slowest of the 16 warps of a pair sets the pace: fc1 ran at 69 % tensor pipe, QKV at 85 %, r2m ncu)
    // tcgen05.ld hands a thread one ROW of the accumulator; stored like that a warp instruction touches 32 different 128-byte lines
    // (r2j ncu of the out-proj GEMM: lg_throttle 3.3 + long_scoreboard 16 per issue, 37 % tensor pipe -- 16 K L1 cycles of row-per-lane
    // residual loads and stores per tile against 10 K cycles of MMA).  Every 32 x 32 chunk is therefore turned through a 4 KB swizzled
    // smem patch: afterwards 8 lanes cover 128 contiguous bytes of one row and a warp instruction covers 4 full lines.
    const GemmEpi& e = p.epi;
    constexpr bool GEN = MODE == EPI_GENERIC;
    constexpr int WCOLS = BN / (NEW / 4);  // columns per warp
    constexpr int NC = WCOLS / 32;         // 32-column chunks per warp
    const int quad = warp & 3, cpart = warp >> 2;
    const bool has_bias = e.bias != nullptr;
    const bool do_alpha = GEN && e.alpha != 1.0f;
    const bool do_act = GEN ? e.act == 1 : MODE == EPI_GELU;
    const bool has_res = GEN ? e.residual != nullptr : MODE == EPI_RESID;
    const bool f32out = GEN ? e.out_f32 != nullptr : MODE == EPI_RESID;
    // [32 rows][8 chunks of 4 floats], chunk ^= row & 7.  Addressed in the shared window explicitly: through the aligned generic pointer the
    // compiler emitted generic LD / ST (r2l ncu: stall_lg on every patch access).
    const uint32_t sE = smem_u32(sB + C::STAGES * C::B_STAGE_BYTES) + (uint32_t)warp * 4096u;
    const int lr = lane >> 3, lc = lane & 7;  // after the turn: rows 4 i + lr (i = 0..7), columns 4 lc .. 4 lc + 3 of the chunk
    int at = 0;
    for (int tile = pair; tile < ntiles; tile += npairs, ++at) {
      const int m2 = tile / p.tiles_n, nt = tile - m2 * p.tiles_n;
      const int as = at & 1;
      const uint32_t aph = (at >> 1) & 1;
      const int row_base = m2 * 2 * BM + (int)rank * BM + quad * 32 + lr;
      long long roff[8];
      uint32_t rmask = 0;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = row_base + 4 * i;
        if (r < p.M) rmask |= 1u << i;
        const int b = r / p.rows_per_item, t = r - b * p.rows_per_item;
        roff[i] = (long long)b * e.batch_stride + (long long)t * e.row_stride;
      }
      const int n0 = nt * BN + cpart * WCOLS;
      const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * BN + cpart * WCOLS);
      auto col_off = [&](int c) {
        const int n = n0 + c * 32 + lc * 4;
        return (long long)(n >> 6) * e.head_stride + (n & 63);
      };
      // the residual of chunk c + 1 is fetched while chunk c is processed, chunk 0's before the accumulator is even complete
      float4 rb[2][8];
      auto fetch_res = [&](int c, float4 (&dst)[8]) {
        const long long co = col_off(c);
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (rmask >> i & 1) dst[i] = __ldcs(reinterpret_cast<const float4*>(e.residual + roff[i] + co));
      };
      if (has_res) fetch_res(0, rb[0]);
      float4 bias4[NC];  // this lane's 4 bias columns of every chunk, fetched before the accumulator is waited for
#pragma unroll
      for (int c = 0; c < NC; ++c) bias4[c] = has_bias ? __ldg(reinterpret_cast<const float4*>(e.bias + n0 + c * 32 + lc * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
      mbar_wait(&accum_full[as], aph);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(trow + c * 32, v);
        if (c + 1 < NC && has_res) fetch_res(c + 1, rb[(c + 1) & 1]);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 8; ++q) sts128(sE + lane * 128 + ((q ^ (lane & 7)) << 4), v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        __syncwarp();
        float4 f[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rr = 4 * i + lr;
          f[i] = lds128(sE + rr * 128 + ((lc ^ (rr & 7)) << 4));
        }
        __syncwarp();  // the patch is free for the next chunk
        const int n = n0 + c * 32 + lc * 4;
        const long long co = col_off(c);
        if (has_bias) {
          const float4 bb = bias4[c];
#pragma unroll
          for (int i = 0; i < 8; ++i) { f[i].x += bb.x; f[i].y += bb.y; f[i].z += bb.z; f[i].w += bb.w; }
        }
        if (do_alpha && n0 + c * 32 < e.alpha_cols) {
#pragma unroll
          for (int i = 0; i < 8; ++i) { f[i].x *= e.alpha; f[i].y *= e.alpha; f[i].z *= e.alpha; f[i].w *= e.alpha; }
        }
        if (do_act) {
#pragma unroll
          for (int i = 0; i < 8; ++i) { f[i].x = gelu_as(f[i].x); f[i].y = gelu_as(f[i].y); f[i].z = gelu_as(f[i].z); f[i].w = gelu_as(f[i].w); }
        }
        if (has_res) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 q = rb[c & 1][i];
            f[i].x += q.x; f[i].y += q.y; f[i].z += q.z; f[i].w += q.w;
          }
        }
        if (f32out) {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (rmask >> i & 1) *reinterpret_cast<float4*>(e.out_f32 + roff[i] + co) = f[i];
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (rmask >> i & 1) *reinterpret_cast<uint2*>(e.out_bf16 + roff[i] + co) = make_uint2(pack_bf16(f[i].x, f[i].y), pack_bf16(f[i].z, f[i].w));
        }
      }
      tc_fence_before();  // this warp's tcgen05.ld of the stage are complete before the issuer may overwrite it
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(&accum_empty[as]);
    }
  }
  // the leader's MMAs read the peer's smem and write its TMEM, commits land on the peer's barriers: nobody leaves early
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == NEW + 1) {
    __syncwarp();
    tmem_dealloc2(tmem_base, C::TMEM_COLS);
  }
}

template <int BN, int MODE, int NEW>
int launch_tc2(cudaStream_t st, const bf16* A, const bf16* W, const Gemm2Params& p, int num_sms) {
  using C = Cfg2<BN, NEW>;
  CUtensorMap tmA, tmW;
  if (int rc = make_tmap_2d_bf16(&tmA, A, (uint64_t)p.M, (uint64_t)p.K, (uint64_t)p.K * 2, BM, BK)) return rc;
  if (int rc = make_tmap_2d_bf16(&tmW, W, (uint64_t)p.N, (uint64_t)p.K, (uint64_t)p.K * 2, BN / 2, BK)) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    BW_CUDA_OK(cudaFuncSetAttribute(gemm_tc2_kernel<BN, MODE, NEW>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_set = true;
  }
  int pairs = num_sms / 2;
  const int ntiles = p.tiles_m2 * p.tiles_n;
  if (pairs > ntiles) pairs = ntiles;
  BW_CUDA_OK(launch_k(gemm_tc2_kernel<BN, MODE, NEW>, dim3(2 * pairs), dim3(C::THREADS), (size_t)C::SMEM_BYTES, st, tmA, tmW, p));
  return 0;
}

}  // namespace

bool gemm_tc2_supported(int M, int N, int K) { return K % BK == 0 && K >= BK && (N % 256 == 0 || N % 128 == 0) && M >= 1; }

int gemm_tc2(cudaStream_t st, const bf16* A, const bf16* W, int M, int N, int K, int rows_per_item, const GemmEpi& epi, int force_bn) {
  BW_CHECK((epi.out_f32 != nullptr) != (epi.out_bf16 != nullptr), "gemm_tc2: exactly one of out_f32/out_bf16 must be set");
  BW_CHECK(gemm_tc2_supported(M, N, K), "gemm_tc2: unsupported shape M=%d N=%d K=%d", M, N, K);
  BW_CHECK(!epi.pos, "gemm_tc2: positional-table epilogue is not supported (conv stem stays on gemm_tc)");
  BW_CHECK(epi.row_stride % 8 == 0 && epi.batch_stride % 8 == 0 && epi.head_stride % 8 == 0, "gemm_tc2: output strides must be multiples of 8");
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    BW_CUDA_OK(cudaGetDevice(&dev));
    BW_CUDA_OK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  Gemm2Params p;
  p.M = M; p.N = N; p.K = K;
  p.tiles_m2 = (M + 2 * BM - 1) / (2 * BM);
  p.rows_per_item = rows_per_item > 0 ? rows_per_item : INT_MAX;
  p.epi = epi;
  int force_mode = -1;
  if (force_bn >= 1000) {  // tests: 1000 + bn = the generic (runtime-parameterised) epilogue instead of the specialised one
    force_mode = EPI_GENERIC;
    force_bn -= 1000;
  }
  int bn = force_bn;
  if (bn == 0) {
    // 256-wide tiles halve the L2 traffic per MAC; fall back to 128 when 256 does not divide N or leaves the last wave thin
    bn = (N % 256 == 0) ? 256 : 128;
    if (bn == 256) {
      const int pairs = num_sms / 2;
      const long long t256 = (long long)p.tiles_m2 * (N / 256);
      const long long waves = (t256 + pairs - 1) / pairs;
      if (t256 * 10 < waves * pairs * 8) bn = 128;  // < 80 % of the last wave's slots used
    }
  }
  BW_CHECK(N % bn == 0, "gemm_tc2: N=%d is not a multiple of the tile width %d", N, bn);
  p.tiles_n = N / bn;
  int mode = EPI_GENERIC;
  if (epi.alpha == 1.0f) {
    if (epi.act == 1 && !epi.residual && epi.out_bf16) mode = EPI_GELU;
    else if (epi.act == 0 && epi.residual && epi.out_f32) mode = EPI_RESID;
    else if (epi.act == 0 && !epi.residual && epi.out_bf16) mode = EPI_PLAIN;
  }
  if (force_mode >= 0) mode = force_mode;
  // (NEW = 16 epilogue warps for the GELU / plain epilogues was measured in r2o: 182.1 ms per 64 chunks against 177.3 with 8 -- the fifth
  //  warpgroup costs a ring stage and 640 threads leave 104 registers per epilogue thread; only NEW = 8 is instantiated)
  const int nw = 8;
#define BW_TC2_CASE(BNV, MODEV, NEWV) \
  if (bn == BNV && mode == MODEV && nw == NEWV) return launch_tc2<BNV, MODEV, NEWV>(st, A, W, p, num_sms);
  BW_TC2_CASE(256, EPI_GENERIC, 8) BW_TC2_CASE(256, EPI_GELU, 8) BW_TC2_CASE(256, EPI_RESID, 8) BW_TC2_CASE(256, EPI_PLAIN, 8)
  BW_TC2_CASE(128, EPI_GENERIC, 8) BW_TC2_CASE(128, EPI_GELU, 8) BW_TC2_CASE(128, EPI_RESID, 8) BW_TC2_CASE(128, EPI_PLAIN, 8)
#undef BW_TC2_CASE
  BW_CHECK(false, "gemm_tc2: no kernel for bn=%d mode=%d", bn, mode);
  return 1;
}

}  // namespace bw
