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
//     warps per CTA) runs under the main loop of tile i + 1.
// Barrier protocol (CUTLASS sm100 2-SM pipeline, restated in raw PTX): TMA loads of BOTH CTAs complete on the LEADER's full[s]
// (cp.async.bulk.tensor ... .cta_group::2 with the peer bit of the mbarrier address cleared; the leader alone posts the expected
// byte count of both); tcgen05.commit ... multicast::cluster releases stage s in both CTAs and publishes a finished accumulator to
// both epilogues; the epilogue warps of both CTAs arrive remotely on the leader's accum_empty[as].
#include <limits.h>

#include "kernels.h"

namespace BW_NS {

namespace {

constexpr int BM = 128;  // rows per CTA (the pair covers 256)
constexpr int BK = 64;
constexpr int A_STAGE_BYTES = BM * BK * 2;  // 16 KB
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;  // clears the CTA-rank bit of a shared::cluster address: rank 0's copy

template <int BN>
struct Cfg2 {
  static constexpr int B_STAGE_BYTES = (BN / 2) * BK * 2;  // this CTA's half of the W tile
  static constexpr int STAGES = (BN == 256) ? 6 : 8;
  static constexpr int TMEM_COLS = 2 * BN;  // two accumulator stages (512 or 256 columns)
  static constexpr int SMEM_BYTES = STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) + 1024 /*align slack*/ + 512 /*barriers*/;
};

struct Gemm2Params {
  int M, N, K;
  int tiles_m2, tiles_n;
  int rows_per_item;  // epilogue address map: b = r / rows_per_item, t = r % rows_per_item
  GemmEpi epi;
};

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
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {  // arrive on rank 0's copy of `bar`
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & PEER_MASK) : "memory");
}

template <int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(192, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const Gemm2Params p) {
  using C = Cfg2<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + C::STAGES * A_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + C::STAGES * C::B_STAGE_BYTES);
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
      mbar_init(&accum_empty[s], 8);  // 4 epilogue warps x 2 CTAs
    }
    fence_mbar_init();
  }
  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
  }
  if (warp == 5) tmem_alloc2(tmem_slot, C::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer's barriers are initialised before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
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
  } else if (warp == 5) {
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
  } else {
    // ---------------- epilogue: warp w owns TMEM lanes [32w, 32w + 32) = rows of this CTA's half tile ----------------
    const GemmEpi& e = p.epi;
    int at = 0;
    for (int tile = pair; tile < ntiles; tile += npairs, ++at) {
      const int m2 = tile / p.tiles_n, nt = tile - m2 * p.tiles_n;
      const int as = at & 1;
      const uint32_t aph = (at >> 1) & 1;
      mbar_wait(&accum_full[as], aph);
      tc_fence_after();
      const int r = m2 * 2 * BM + (int)rank * BM + warp * 32 + lane;
      const bool row_ok = r < p.M;
      const int b = r / p.rows_per_item, t = r - b * p.rows_per_item;
      const long long row_off = (long long)b * e.batch_stride + (long long)t * e.row_stride;
      const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(as * BN);
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        const int n = nt * BN + c * 32;
        uint32_t v[32];
        tmem_ld_32x32(trow + c * 32, v);
        tmem_ld_wait();
        if (row_ok && n < p.N) {
          const long long off = row_off + (long long)(n >> 6) * e.head_stride + (n & 63);
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
          if (e.bias) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 bb = *reinterpret_cast<const float4*>(e.bias + n + j);
              f[j] += bb.x; f[j + 1] += bb.y; f[j + 2] += bb.z; f[j + 3] += bb.w;
            }
          }
          if (e.alpha != 1.0f && n < e.alpha_cols) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] *= e.alpha;
          }
          if (e.act == 1) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = gelu_erf(f[j]);
          }
          if (e.residual) {
            const float* rp = e.residual + off;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 q = *reinterpret_cast<const float4*>(rp + j);
              f[j] += q.x; f[j + 1] += q.y; f[j + 2] += q.z; f[j + 3] += q.w;
            }
          }
          if (e.out_f32) {
            float* op = e.out_f32 + off;
#pragma unroll
            for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(op + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
          } else {
            bf16* op = e.out_bf16 + off;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              uint4 q;
              q.x = pack_bf16(f[j], f[j + 1]);
              q.y = pack_bf16(f[j + 2], f[j + 3]);
              q.z = pack_bf16(f[j + 4], f[j + 5]);
              q.w = pack_bf16(f[j + 6], f[j + 7]);
              *reinterpret_cast<uint4*>(op + j) = q;
            }
          }
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
  if (warp == 5) {
    __syncwarp();
    tmem_dealloc2(tmem_base, C::TMEM_COLS);
  }
}

template <int BN>
int launch_tc2(cudaStream_t st, const bf16* A, const bf16* W, const Gemm2Params& p, int num_sms) {
  using C = Cfg2<BN>;
  CUtensorMap tmA, tmW;
  if (int rc = make_tmap_2d_bf16(&tmA, A, (uint64_t)p.M, (uint64_t)p.K, (uint64_t)p.K * 2, BM, BK)) return rc;
  if (int rc = make_tmap_2d_bf16(&tmW, W, (uint64_t)p.N, (uint64_t)p.K, (uint64_t)p.K * 2, BN / 2, BK)) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    BW_CUDA_OK(cudaFuncSetAttribute(gemm_tc2_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_set = true;
  }
  int pairs = num_sms / 2;
  const int ntiles = p.tiles_m2 * p.tiles_n;
  if (pairs > ntiles) pairs = ntiles;
  gemm_tc2_kernel<BN><<<2 * pairs, 192, C::SMEM_BYTES, st>>>(tmA, tmW, p);
  BW_CUDA_OK(cudaGetLastError());
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
  if (bn == 256) return launch_tc2<256>(st, A, W, p, num_sms);
  return launch_tc2<128>(st, A, W, p, num_sms);
}

}  // namespace bw
