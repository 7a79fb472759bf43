// tcgen05 / TMEM / TMA GEMM for sm_100a:  C[b,t,n] = epi( sum_k A(b,t,k) * W[n,k] ).
//
// Replaces the cuBLAS GEMMs + cuDNN convs the reference reaches through torch
// (TF/models/whisper/modeling_whisper.py:279-336 q/k/v/out projections, :404-407 fc1/fc2, :619-620 conv stem).
//
// Structure (one 128 x BN output tile per CTA, 192 threads):
//   warps 0-3  epilogue: tcgen05.ld accumulator rows -> bias/alpha/GELU/pos/residual -> vectorised st.global
//   warp  4    TMA producer (one lane): 128B-swizzled K-major boxes of A (3-D map, wrapping k for the conv
//              stem) and W into a STAGES-deep smem ring, completion on mbarriers
//   warp  5    TMEM allocator + MMA issuer (one lane): 4 x tcgen05.mma (K=16) per 64-wide k-block,
//              tcgen05.commit releases the smem stage / signals the epilogue
// BN=128 uses 3 stages (96 KB) so two CTAs share an SM and one CTA's epilogue overlaps the other's main loop.
#include <limits.h>

#include "kernels.h"

namespace BW_NS {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int A_STAGE_BYTES = BM * BK * 2;  // 16 KB

// BN = 32 is the decoder-step shape (q_len = 1 for up to 128 sequences per tile): the GEMM is a stream over the weight matrix, so
// the tile is narrow (N / 32 CTAs cover the SMs without split-K for N >= 3840) and the ring is deep (8 stages x 20 KB in
// flight per SM hide the DRAM latency; one CTA per SM).
template <int BN>
struct Cfg {
  static constexpr int STAGES = (BN == 128) ? 3 : (BN == 32) ? 8 : 4;
  static constexpr int B_STAGE_BYTES = BN * BK * 2;
  static constexpr int TMEM_COLS = (BN < 32) ? 32 : BN;  // power of two >= 32
  static constexpr int SMEM_BYTES = STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr int MIN_CTAS = (BN == 256 || BN == 32) ? 1 : 2;
};

struct GemmParams {
  int B, rows, N, K;
  int kwrap;
  int tiles_m;  // per item
  int ksplit;   // gridDim.z: split z handles k-blocks [z * kper, min(nk, (z + 1) * kper)) and writes its partial sums at
  int kper;     // out + z * split_stride (no bias / residual: the consumer adds the partials -- deterministic, no atomics)
  long long split_stride;
  GemmEpi epi;
};

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

template <int BN>
__global__ void __launch_bounds__(192, Cfg<BN>::MIN_CTAS)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const GemmParams p) {
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + C::STAGES * A_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + C::STAGES * C::B_STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + C::STAGES;
  uint64_t* accum_full = bars + 2 * C::STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::STAGES + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.y / p.tiles_m;
  const int t0 = (blockIdx.y % p.tiles_m) * BM;
  const int n0 = blockIdx.x * BN;
  const int nk_all = (p.K + BK - 1) / BK;
  const int kb0 = (int)blockIdx.z * p.kper;
  const int nk = min(nk_all, kb0 + p.kper) - kb0;  // >= 1 (the launcher picks ksplit so that every split owns a k-block)

  if (threadIdx.x == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(accum_full, 1);
    fence_mbar_init();
  }
  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
  }
  if (warp == 5) tmem_alloc(tmem_slot, C::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    if (lane == 0) {
      // The weight tiles of the first ring pass do not depend on the previous kernel: under programmatic dependent launch they are
      // requested before the wait (their DRAM latency runs beside the predecessor's tail); the activations after it.
      const int npre = nk < C::STAGES ? nk : C::STAGES;
      for (int kb = 0; kb < npre; ++kb) {
        mbar_arrive_expect_tx(&full[kb], A_STAGE_BYTES + C::B_STAGE_BYTES);
        tma_load_2d(sB + kb * C::B_STAGE_BYTES, &tmW, &full[kb], (kb0 + kb) * BK, n0);
      }
      pdl_wait();
      pdl_launch();
      for (int kb = 0; kb < nk; ++kb) {
        const int s = kb % C::STAGES;
        const uint32_t ph = (kb / C::STAGES) & 1;
        const int k = (kb0 + kb) * BK;
        if (kb >= npre) {
          mbar_wait(&empty[s], ph ^ 1);
          mbar_arrive_expect_tx(&full[s], A_STAGE_BYTES + C::B_STAGE_BYTES);
          tma_load_2d(sB + s * C::B_STAGE_BYTES, &tmW, &full[s], k, n0);
        }
        tma_load_3d(sA + s * A_STAGE_BYTES, &tmA, &full[s], k % p.kwrap, t0 + k / p.kwrap, b);
      }
    }
  } else if (warp == 5) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN);
      for (int kb = 0; kb < nk; ++kb) {
        const int s = kb % C::STAGES;
        const uint32_t ph = (kb / C::STAGES) & 1;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint64_t a0 = umma_desc_sw128(smem_u32(sA + s * A_STAGE_BYTES));
        const uint64_t b0 = umma_desc_sw128(smem_u32(sB + s * C::B_STAGE_BYTES));
#pragma unroll
        for (int k = 0; k < BK / 16; ++k)  // +32 B per K=16 step inside the 128 B swizzle atom
          umma_bf16(tmem_base, a0 + 2 * k, b0 + 2 * k, idesc, (uint32_t)((kb | k) != 0));
        umma_commit(&empty[s]);
      }
      umma_commit(accum_full);
    }
  } else {
    // ---------------- epilogue: warp w owns TMEM lanes [32w, 32w+32) = tile rows ----------------
    pdl_wait();  // (residual reads / output stores: after the predecessor grid)
    mbar_wait(accum_full, 0);
    tc_fence_after();
    const int t = t0 + warp * 32 + lane;
    const bool row_ok = t < p.rows;
    const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
    const GemmEpi& e = p.epi;
    const long long row_off = (long long)b * e.batch_stride + (long long)t * e.row_stride + (long long)blockIdx.z * p.split_stride;
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      const int n = n0 + c * 32;
      if (n >= p.N) break;  // warp-uniform
      uint32_t v[32];
      tmem_ld_32x32(trow + c * 32, v);
      tmem_ld_wait();
      if (row_ok) {
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
        if (e.pos) {
          const float* pp = e.pos + (long long)t * p.N + n;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 q = *reinterpret_cast<const float4*>(pp + j);
            f[j] += q.x; f[j + 1] += q.y; f[j + 2] += q.z; f[j + 3] += q.w;
          }
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
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    __syncwarp();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// CUDA-core sibling (comparator / bring-up fallback).  One thread per output, 16x16 tile, no staging.
// ------------------------------------------------------------------------------------------------
__global__ void gemm_simt_kernel(GemmA a, const bf16* __restrict__ W, GemmParams p) {
  const int n = blockIdx.x * 16 + threadIdx.x;
  const int t = (blockIdx.y % p.tiles_m) * 16 + threadIdx.y;
  const int b = blockIdx.y / p.tiles_m;
  if (n >= p.N || t >= p.rows) return;
  const bf16* ab = a.base + (long long)b * a.batch_stride;
  const bf16* w = W + (long long)n * p.K;
  float acc = 0.f;
  const int kend = (p.epi.n_valid > 0 && n >= p.epi.n_valid) ? 0 : p.K;  // rows of W beyond n_valid do not exist: zero
  for (int k = 0; k < kend; ++k) {
    const float av = e2f(ab[(long long)(t + k / p.kwrap) * a.pitch + (k % p.kwrap)]);
    acc = fmaf(av, e2f(w[k]), acc);
  }
  const GemmEpi& e = p.epi;
  const long long off = (long long)b * e.batch_stride + (long long)t * e.row_stride + (long long)(n >> 6) * e.head_stride + (n & 63);
  if (e.bias) acc += e.bias[n];
  if (n < e.alpha_cols) acc *= e.alpha;
  if (e.act == 1) acc = gelu_erf(acc);
  if (e.pos) acc += e.pos[(long long)t * p.N + n];
  if (e.residual) acc += e.residual[off];
  if (e.out_f32) e.out_f32[off] = acc;
  else e.out_bf16[off] = f2e(acc);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !p) {
    set_error("cudaGetDriverEntryPoint(cuTensorMapEncodeTiled) failed: %s", cudaGetErrorString(e));
    return nullptr;
  }
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

int check_epi(const GemmEpi& e, int N) {
  BW_CHECK((e.out_f32 != nullptr) != (e.out_bf16 != nullptr), "gemm: exactly one of out_f32/out_bf16 must be set");
  BW_CHECK(N % 32 == 0, "gemm: N=%d must be a multiple of 32", N);
  BW_CHECK(e.row_stride % 8 == 0 && e.batch_stride % 8 == 0 && e.head_stride % 8 == 0, "gemm: output strides must be multiples of 8");
  return 0;
}

}  // namespace

int make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t row_pitch_bytes,
                      uint32_t box_rows, uint32_t box_cols) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return -1;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {row_pitch_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = fn(out, BW_TMAP_DTYPE, 2, const_cast<void*>(base), dims, strides, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  BW_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(2d rows=%llu cols=%llu pitch=%llu box=%ux%u) -> %d",
           (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)row_pitch_bytes, box_rows, box_cols, (int)r);
  return 0;
}

static int make_tmap_3d_bf16(CUtensorMap* out, const void* base, uint64_t batch, uint64_t rows, uint64_t cols,
                             uint64_t row_pitch_bytes, uint64_t batch_pitch_bytes, uint32_t box_rows, uint32_t box_cols) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return -1;
  cuuint64_t dims[3] = {cols, rows, batch};
  cuuint64_t strides[2] = {row_pitch_bytes, batch_pitch_bytes};
  cuuint32_t box[3] = {box_cols, box_rows, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = fn(out, BW_TMAP_DTYPE, 3, const_cast<void*>(base), dims, strides, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  BW_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(3d batch=%llu rows=%llu cols=%llu pitch=%llu/%llu) -> %d",
           (unsigned long long)batch, (unsigned long long)rows, (unsigned long long)cols,
           (unsigned long long)row_pitch_bytes, (unsigned long long)batch_pitch_bytes, (int)r);
  return 0;
}

template <int BN>
static int launch_tc(cudaStream_t st, const CUtensorMap& tmA, const GemmA& a, const bf16* W, const GemmParams& p) {
  using C = Cfg<BN>;
  CUtensorMap tmW;
  if (int rc = make_tmap_2d_bf16(&tmW, W, p.epi.n_valid > 0 ? p.epi.n_valid : p.N, p.K, (uint64_t)p.K * 2, BN, BK)) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    BW_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_set = true;
  }
  dim3 grid((p.N + BN - 1) / BN, p.tiles_m * p.B, p.ksplit);
  BW_CUDA_OK(launch_k(gemm_tc_kernel<BN>, grid, dim3(192), (size_t)C::SMEM_BYTES, st, tmA, tmW, p));
  return 0;
}

int gemm_tc(cudaStream_t st, const GemmA& a, const bf16* W, int B, int rows, int N, int K, const GemmEpi& epi, int force_bn) {
  return gemm_tc_split(st, a, W, B, rows, N, K, epi, force_bn, 1, 0);
}

int gemm_tc_ksplit(int K, int ksplit) {
  const int nk = K / BK;
  if (ksplit > nk) ksplit = nk;
  if (ksplit < 1) ksplit = 1;
  const int kper = (nk + ksplit - 1) / ksplit;
  return (nk + kper - 1) / kper;
}

int gemm_tc_split(cudaStream_t st, const GemmA& a, const bf16* W, int B, int rows, int N, int K, const GemmEpi& epi, int force_bn,
                  int ksplit, long long split_stride) {
  if (int rc = check_epi(epi, N)) return rc;
  BW_CHECK(ksplit >= 1 && (ksplit == 1 || (epi.out_f32 && !epi.bias && !epi.residual && !epi.pos && epi.act == 0 && epi.alpha == 1.0f)),
           "gemm_tc: split-K writes raw fp32 partial sums (no bias / activation / residual)");
  BW_CHECK(K % 64 == 0, "gemm_tc: K=%d must be a multiple of 64", K);
  BW_CHECK(a.pitch % 8 == 0 && a.batch_stride % 8 == 0, "gemm_tc: A pitch/batch stride must be multiples of 8 elements");
  BW_CHECK(a.kwrap >= K || a.kwrap % 64 == 0, "gemm_tc: kwrap=%d must be a multiple of 64", a.kwrap);
  GemmParams p;
  p.B = B; p.rows = rows; p.N = N; p.K = K;
  p.kwrap = a.kwrap >= K ? INT_MAX : a.kwrap;
  p.tiles_m = (rows + BM - 1) / BM;
  p.epi = epi;
  {
    const int nk = K / BK;
    p.ksplit = gemm_tc_ksplit(K, ksplit);  // every split owns >= 1 k-block
    p.kper = (nk + p.ksplit - 1) / p.ksplit;
    p.split_stride = split_stride;
  }
  const uint64_t inner = (uint64_t)(a.kwrap >= K ? K : a.kwrap);
  CUtensorMap tmA;
  if (int rc = make_tmap_3d_bf16(&tmA, a.base, (uint64_t)B, (uint64_t)a.rows_base, inner, (uint64_t)a.pitch * 2,
                                 (uint64_t)(B > 1 ? a.batch_stride : (long long)a.rows_base * a.pitch) * 2, BM, BK))
    return rc;
  int bn = force_bn;
  if (bn == 0) {
    // enough CTAs to cover the 148 SMs matters more than the wider tile when the grid is small
    const long long tiles128 = (long long)((N + 127) / 128) * p.tiles_m * B;
    bn = (N % 256 == 0 && tiles128 >= 4 * 148) ? 256 : 128;
    if (N < 128) bn = 64;
  }
  switch (bn) {
    case 32: return launch_tc<32>(st, tmA, a, W, p);
    case 64: return launch_tc<64>(st, tmA, a, W, p);
    case 128: return launch_tc<128>(st, tmA, a, W, p);
    case 256: return launch_tc<256>(st, tmA, a, W, p);
  }
  BW_CHECK(false, "gemm_tc: unsupported BN=%d", bn);
}

int gemm_simt(cudaStream_t st, const GemmA& a, const bf16* W, int B, int rows, int N, int K, const GemmEpi& epi) {
  if (int rc = check_epi(epi, N)) return rc;
  GemmParams p;
  p.B = B; p.rows = rows; p.N = N; p.K = K;
  p.kwrap = a.kwrap >= K ? INT_MAX : a.kwrap;
  p.tiles_m = (rows + 15) / 16;
  p.ksplit = 1; p.kper = 0; p.split_stride = 0;
  p.epi = epi;
  dim3 grid((N + 15) / 16, p.tiles_m * B);
  gemm_simt_kernel<<<grid, dim3(16, 16), 0, st>>>(a, W, p);
  BW_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace bw
