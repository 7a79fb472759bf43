// Word-level timestamps on the device: alignment-head cross-attention scores -> softmax -> crop -> z-score over
// tokens -> median filter (width 7, reflect) -> mean over heads -> dynamic time warping -> jump times.
//
// Replaces WhisperGenerationMixin._extract_token_timestamps (TF/models/whisper/generation_whisper.py:241-381),
// _median_filter (:43-61) and the pure-Python/NumPy double loop _dynamic_time_warping (:64-115).  The reference needs
// eager attention for all 32x20 heads with every step's weights kept; here only the alignment heads' raw scores are
// recorded by the cross-attention kernel and the rest happens in four small kernels per audio.
// Semantics kept bit-for-bit where it matters: population std (unbiased=False) with no epsilon (std 0 -> inf/NaN like
// torch), float32 cost cells fed by a float64 sum (:70,367), strict '<' tie-breaking with the else-branch choosing
// c2 (:80-85), trace[0,:]=2 / trace[:,0]=1 before the backtrace (:93-94).
#include <math.h>

#include "kernels.h"

namespace BW_NS {

namespace {

// One descriptor per audio of a batch (blockIdx.z): which alignment block, how many generated tokens / valid frames.
// Per-item scratch lives at work + z * work_stride (floats), results at out + z * out_stride.
struct TsBatch {
  const int* items;     // [n][3] device: slot index of the alignment block, T, NF
  const int* slot_map;  // optional [n][map_pitch]: row t of item i is read from slot slot_map[i * map_pitch + t] (beam search: the
  int map_pitch;        // slot that was the returned sequence's ancestor at step t), null = the item's own slot for every row
  long long work_stride, out_stride, align_stride;
  int Ha, Tcap, S;
};
struct TsWork {
  float* probs;
  float* z;
  double* negm;
  signed char* trace;
};
__device__ __forceinline__ TsWork ts_work(float* work, const TsBatch& b, int item) {
  TsWork w;
  w.probs = work + (long long)item * b.work_stride;
  w.z = w.probs + (size_t)b.Ha * b.Tcap * b.S;
  w.negm = reinterpret_cast<double*>(w.z + (size_t)b.Ha * b.Tcap * b.S);
  w.trace = reinterpret_cast<signed char*>(w.negm + (size_t)b.Tcap * b.S);
  return w;
}

// scores [Ha][Tcap][S] (one audio) -> probs [Ha][T][NF] = softmax over all S keys, first NF kept
__global__ void ts_softmax_kernel(const float* __restrict__ align, float* __restrict__ work, const TsBatch b) {
  __shared__ float red[32];
  const int t = blockIdx.x, ha = blockIdx.y, item = blockIdx.z;
  const int T = b.items[item * 3 + 1], NF = b.items[item * 3 + 2], Tcap = b.Tcap, S = b.S;
  if (t >= T) return;
  const int audio = b.slot_map ? b.slot_map[item * b.map_pitch + t] : b.items[item * 3];
  const float* scores = align + (long long)audio * b.align_stride;
  float* probs = ts_work(work, b, item).probs;
  const float* row = scores + ((long long)ha * Tcap + t) * S;
  float mx = -INFINITY;
  for (int j = threadIdx.x; j < S; j += blockDim.x) mx = fmaxf(mx, row[j]);
  mx = warp_max(mx);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = -INFINITY;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) mx = fmaxf(mx, red[i]);
  __syncthreads();
  float s = 0.f;
  for (int j = threadIdx.x; j < S; j += blockDim.x) s += expf(row[j] - mx);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  float tot = 0.f;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) tot += red[i];
  float* out = probs + ((long long)ha * T + t) * NF;
  for (int j = threadIdx.x; j < NF; j += blockDim.x) out[j] = expf(row[j] - mx) / tot;
}

// z-score over the token axis for each (head, frame)
__global__ void ts_zscore_kernel(float* __restrict__ work, const TsBatch b) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int ha = blockIdx.y, item = blockIdx.z;
  const int T = b.items[item * 3 + 1], NF = b.items[item * 3 + 2];
  const TsWork w = ts_work(work, b, item);
  const float* probs = w.probs;
  float* z = w.z;
  if (j >= NF) return;
  const float* p = probs + (long long)ha * T * NF + j;
  float s = 0.f;
  for (int t = 0; t < T; ++t) s += p[(long long)t * NF];
  const float mean = s / (float)T;
  float ss = 0.f;
  for (int t = 0; t < T; ++t) {
    const float d = p[(long long)t * NF] - mean;
    ss += d * d;
  }
  const float sd = sqrtf(ss / (float)T);
  float* o = z + (long long)ha * T * NF + j;
  for (int t = 0; t < T; ++t) o[(long long)t * NF] = (p[(long long)t * NF] - mean) / sd;
}

__device__ __forceinline__ void cswap(float& a, float& b) {
  // NaN-tolerant compare-exchange (NaNs sort last, as torch.sort does)
  const bool sw = (a > b) || (isnan(a) && !isnan(b));
  if (sw) { const float t = a; a = b; b = t; }
}

// median-7 along frames (reflect padding), then mean over heads; output NEGATED in double for the DTW
__global__ void ts_median_mean_kernel(float* __restrict__ work, const TsBatch b) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int t = blockIdx.y, item = blockIdx.z;
  const int T = b.items[item * 3 + 1], NF = b.items[item * 3 + 2], Ha = b.Ha;
  const TsWork w = ts_work(work, b, item);
  const float* z = w.z;
  double* negm = w.negm;
  if (j >= NF || t >= T) return;
  float acc = 0.f;
  for (int ha = 0; ha < Ha; ++ha) {
    const float* row = z + ((long long)ha * T + t) * NF;
    float v;
    if (NF <= 3) {
      v = row[j];
    } else {
      float w[7];
#pragma unroll
      for (int k = 0; k < 7; ++k) {
        int idx = j + k - 3;
        if (idx < 0) idx = -idx;
        if (idx >= NF) idx = 2 * (NF - 1) - idx;
        w[k] = row[idx];
      }
      // sort 7 (insertion network)
#pragma unroll
      for (int a = 0; a < 7; ++a)
#pragma unroll
        for (int b = 0; b + 1 < 7 - a; ++b) cswap(w[b], w[b + 1]);
      v = w[3];
    }
    acc += v;
  }
  const float m = acc / (float)Ha;  // torch mean over heads in float32
  negm[(long long)t * NF + j] = -(double)m;
}

// anti-diagonal wavefront DTW for one audio, then a serial backtrace.  One block per audio of the batch.
__global__ void __launch_bounds__(512) ts_dtw_kernel(float* __restrict__ work, const TsBatch b, double time_precision, float* __restrict__ out_all) {
  extern __shared__ float diag[];  // 3 x (Tcap + 2)
  const int item = blockIdx.x;
  const int T = b.items[item * 3 + 1], NF = b.items[item * 3 + 2];
  const TsWork w = ts_work(work, b, item);
  const double* negm = w.negm;
  signed char* trace = w.trace;
  float* out = out_all + (long long)item * b.out_stride;
  float* d0 = diag;               // diagonal d-2
  float* d1 = diag + (T + 2);     // diagonal d-1
  float* d2 = diag + 2 * (T + 2); // diagonal d
  const int W = NF + 1;
  // cost[i][j], i in [0,T], j in [0,NF]; diagonal index = i + j, stored by i
  for (int i = threadIdx.x; i <= T; i += blockDim.x) {
    d0[i] = (i == 0) ? 0.f : INFINITY;  // d = 0: only (0,0)
    d1[i] = INFINITY;                   // d = 1: (0,1),(1,0) both inf
  }
  __syncthreads();
  for (int d = 2; d <= T + NF; ++d) {
    const int ilo = max(1, d - NF), ihi = min(T, d - 1);
    for (int i = ilo + threadIdx.x; i <= ihi; i += blockDim.x) {
      const int j = d - i;
      const float c0 = d0[i - 1];  // cost[i-1][j-1]
      const float c1 = d1[i - 1];  // cost[i-1][j]
      const float c2 = d1[i];      // cost[i][j-1]
      float c;
      signed char tr;
      if (c0 < c1 && c0 < c2) { c = c0; tr = 0; }
      else if (c1 < c0 && c1 < c2) { c = c1; tr = 1; }
      else { c = c2; tr = 2; }
      d2[i] = (float)(negm[(long long)(i - 1) * NF + (j - 1)] + (double)c);
      trace[(long long)i * W + j] = tr;
    }
    // boundary cells of this diagonal are infinite: (0, d) and (d, 0)
    if (threadIdx.x == 0) {
      d2[0] = INFINITY;
      if (d <= T) d2[d] = INFINITY;
    }
    __syncthreads();
    float* tmp = d0; d0 = d1; d1 = d2; d2 = tmp;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    int i = T, j = NF;
    while (i > 0 || j > 0) {
      out[i - 1 >= 0 ? i - 1 : 0] = (float)((double)(j - 1) * time_precision);  // float32(index * python float), as the reference stores it  // forward-first visit wins (overwrites)
      signed char tr;
      if (i == 0) tr = 2;
      else if (j == 0) tr = 1;
      else tr = trace[(long long)i * W + j];
      if (tr == 0) { --i; --j; }
      else if (tr == 1) { --i; }
      else { --j; }
    }
    out[T] = out[T - 1];
  }
}

}  // namespace

// per-item work layout (floats): probs [Ha*Tcap*S] | z [Ha*Tcap*S] | negm (double) [Tcap*S] | trace (int8) [(Tcap+1)*(S+1)]
size_t word_timestamps_work_floats(int Ha, int Tcap, int S) {
  const size_t n = (size_t)2 * Ha * Tcap * S + (size_t)2 * Tcap * S + ((size_t)(Tcap + 2) * (S + 2) + 3) / 4 + 64;
  return (n + 3) / 4 * 4;  // items stay 16-byte aligned (the double section starts at an even float offset)
}

// n audios in one pass (4 launches whatever n is): items_dev [n][3] = (audio, T, NF); out_dev [n][Tcap + 8] seconds
int word_timestamps_batch_device(cudaStream_t st, const float* align, int Ha, int Tcap, int S, const int* items_dev, const int* slot_map_dev,
                                 int map_pitch, int n, int maxT, int maxNF, double time_precision, float* work, float* out_dev) {
  TsBatch b;
  b.items = items_dev;
  b.slot_map = slot_map_dev;
  b.map_pitch = map_pitch;
  b.work_stride = (long long)word_timestamps_work_floats(Ha, Tcap, S);
  b.out_stride = Tcap + 8;
  b.align_stride = (long long)Ha * Tcap * S;
  b.Ha = Ha; b.Tcap = Tcap; b.S = S;
  ts_softmax_kernel<<<dim3(maxT, Ha, n), 256, 0, st>>>(align, work, b);
  BW_CUDA_OK(cudaGetLastError());
  ts_zscore_kernel<<<dim3((maxNF + 127) / 128, Ha, n), 128, 0, st>>>(work, b);
  BW_CUDA_OK(cudaGetLastError());
  ts_median_mean_kernel<<<dim3((maxNF + 127) / 128, maxT, n), 128, 0, st>>>(work, b);
  BW_CUDA_OK(cudaGetLastError());
  ts_dtw_kernel<<<n, 512, 3 * (Tcap + 2) * sizeof(float), st>>>(work, b, time_precision, out_dev);
  BW_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace bw
