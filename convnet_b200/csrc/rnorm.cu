// rnorm.cu — cross-map response normalisation, forward and backward. HBM-bound.
//
// Replaces kCrossMapRNorm / kCrossMapDenoms / kCrossMapRNormUndo
// (cudamat_conv_gemm.cu:438-543) and kFCNorm / kFRNormUndo2
// (cudamat_conv_others.cu:1158-1300,1562-1660).
//
// One thread owns one location (n, x, y) and walks the channels at stride num_locs,
// so a warp reads/writes one full 128-byte line per channel.  The sliding window
// lives in a per-thread shared-memory ring (k entries), which makes both passes
// single-read / single-write per element:
//   forward : reads x once, writes y once                    (2 floats/element)
//   backward: reads x and dy once, writes dx once            (3 floats/element)
// The reference's backward needs a cudaMalloc'd `denoms` scratch and two kernels per
// 4096-location batch (gemm.cu:1365-1399); here it is one launch with no scratch.
// For windows too large for shared memory the ring spills to library workspace.
//
//   forward window of channel j : [j - a, j + b]  with a = k/2, b = k - k/2 - 1   (gemm.cu:475-477)
//   inverse window of channel j : [j - b, j + a]                                 (gemm.cu:528-530)
//   blocked: both are the block [ (j/k)*k, (j/k)*k + k ).
#include <algorithm>

#include "conv_kernels.h"

namespace cnb {

constexpr int RN_THREADS = 128;

// ring element (slot, thread): ring[slot * ring_stride + lane].
// blockIdx.y selects a channel SEGMENT [f0, f1) (host: only when there are too few locations to fill the GPU); a
// segment re-reads the k-1 (forward) / 2(k-1) (backward) halo channels of its neighbours instead of waiting for them.
template <bool BLOCKED>
__global__ void __launch_bounds__(RN_THREADS) rnorm_fwd_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                                long long L, int F, int k, float alpha, float beta,
                                                                float* gring, long long gstride, int seg) {
  extern __shared__ float sring[];
  const long long loc = blockIdx.x * (long long)RN_THREADS + threadIdx.x;
  if (loc >= L) return;
  float* ring = gring ? gring + loc + (long long)blockIdx.y * k * gstride : sring + threadIdx.x;
  const long long rs = gring ? gstride : RN_THREADS;
  x += loc; y += loc;
  const int f0 = blockIdx.y * seg, f1 = min(F, f0 + seg);
  if (BLOCKED) {                                        // host guarantees seg % k == 0
    for (int s = f0; s < f1; s += k) {
      const int e = min(F, s + k);
      float sum = 0.f;
      for (int i = s; i < e; i++) { const float v = __ldg(x + (long long)i * L); ring[(i - s) * rs] = v; sum += v * v; }
      const float sc = __powf(1.f + alpha * sum, -beta);
      for (int i = s; i < e; i++) y[(long long)i * L] = ring[(i - s) * rs] * sc;
    }
    return;
  }
  const int a = k / 2, b = k - a - 1;
  float sum = 0.f;
  // q = entering channel; output channel j = q - b; window [j-a, j+b] = [q-k+1, q].
  // Loads are hoisted eight steps ahead of the (serial) ring updates to keep HBM requests in flight.
  const int q0 = max(0, f0 - a), q1 = f1 + b;      // channels >= F enter as zeros
  int slot_q = q0 % k, slot_j = ((q0 - b) % k + k) % k;           // ring slots of q and of j = q - b, advanced with wrap
  for (int qb = q0; qb < q1; qb += 8) {
    float xv[8];
#pragma unroll
    for (int u = 0; u < 8; u++) xv[u] = (qb + u < F) ? __ldg(x + (long long)(qb + u) * L) : 0.f;
#pragma unroll
    for (int u = 0; u < 8; u++) {
      const int q = qb + u;
      if (q >= q1) break;
      const int slot = slot_q;
      const float v = xv[u];
      float old = 0.f;
      if (q - q0 >= k) old = ring[slot * rs];
      if (q < F) ring[slot * rs] = v;
      sum += v * v - old * old;
      const int j = q - b;
      if (j >= f0 && j < f1) {
        const float xj = (j == q) ? v : ring[slot_j * rs];
        y[(long long)j * L] = xj * __powf(1.f + alpha * sum, -beta);
      }
      if (++slot_q == k) slot_q = 0;
      if (++slot_j == k) slot_j = 0;
    }
  }
}

template <bool BLOCKED>
__global__ void __launch_bounds__(RN_THREADS) rnorm_undo_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                                 float* __restrict__ dx, long long L, int F, int k,
                                                                 float alpha, float beta, float* gring, long long gstride,
                                                                 int seg) {
  // three rings of k entries per thread: x, t = dy*x*denom, p = dy*denom^(beta/(beta+1))
  extern __shared__ float sring[];
  const long long loc = blockIdx.x * (long long)RN_THREADS + threadIdx.x;
  if (loc >= L) return;
  const long long rs = gring ? gstride : RN_THREADS;
  float* rx = gring ? gring + loc + (long long)blockIdx.y * 3 * k * gstride : sring + threadIdx.x;
  float* rt = rx + (long long)k * rs;
  float* rp = rt + (long long)k * rs;
  x += loc; dy += loc; dx += loc;
  const float c2 = 2.f * alpha * beta;
  const int f0 = blockIdx.y * seg, f1 = min(F, f0 + seg);
  if (BLOCKED) {
    for (int s = f0; s < f1; s += k) {
      const int e = min(F, s + k);
      float sum = 0.f;
      for (int i = s; i < e; i++) { const float v = __ldg(x + (long long)i * L); rx[(i - s) * rs] = v; sum += v * v; }
      const float base = 1.f + alpha * sum;
      const float denom = __powf(base, -beta - 1.f), pw = __powf(base, -beta);
      float st = 0.f;
      for (int i = s; i < e; i++) { const float g = __ldg(dy + (long long)i * L); rt[(i - s) * rs] = g; st += g * rx[(i - s) * rs] * denom; }
      for (int i = s; i < e; i++) dx[(long long)i * L] = rt[(i - s) * rs] * pw - c2 * rx[(i - s) * rs] * st;
    }
    return;
  }
  const int a = k / 2, b = k - a - 1;
  float sumsq = 0.f, sumt = 0.f;
  // stage 1: entering channel q; channel i = q - b gets its forward sum, t_i and p_i
  // stage 2: output channel j = i - a gets sum of t over [j-b, j+a] = [i-k+1, i]
  const int q0 = max(0, f0 - (k - 1));                   // first x needed: (f0 - b) - a
  const int i0 = max(0, f0 - b);                         // first t needed
  const int Qend = f1 + a + b;                           // last output f1-1 needs t up to f1-1+a, i.e. q up to f1-1+a+b
  // ring slots of q, i = q - b and j = i - a, advanced with wrap instead of three `% k` per channel
  int slot_q = q0 % k, slot_i = ((q0 - b) % k + k) % k, slot_j = ((q0 - b - a) % k + k) % k;
  for (int qb = q0; qb < Qend; qb += 4) {
    float xv[4], gv[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int q = qb + u, i = q - b;
      xv[u] = (q < F) ? __ldg(x + (long long)q * L) : 0.f;
      gv[u] = (i >= i0 && i < F) ? __ldg(dy + (long long)i * L) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int q = qb + u;
      if (q >= Qend) break;
      const int sq = slot_q, si = slot_i, sj = slot_j;
      if (++slot_q == k) slot_q = 0;
      if (++slot_i == k) slot_i = 0;
      if (++slot_j == k) slot_j = 0;
      {
        const int slot = sq;
        const float v = xv[u];
        float old = 0.f;
        if (q - q0 >= k) old = rx[slot * rs];
        rx[slot * rs] = v;                              // zeros once q >= F
        sumsq += v * v - old * old;
      }
      const int i = q - b;
      if (i < i0) continue;
      {
        const int slot = si;
        float told = 0.f, t = 0.f;
        if (i - i0 >= k) told = rt[slot * rs];
        if (i < F) {
          const float base = 1.f + alpha * sumsq;
          const float g = gv[u];
          const float denom = __powf(base, -beta - 1.f);
          t = g * rx[slot * rs] * denom;
          rp[slot * rs] = g * denom * base;          // = g * base^(-beta)  (== denom^(beta/(beta+1)), gemm.cu:538)
        }
        rt[slot * rs] = t;
        sumt += t - told;
      }
      const int j = i - a;
      if (j >= f0 && j < f1) {
        const int slot = sj;
        dx[(long long)j * L] = rp[slot * rs] - c2 * rx[slot * rs] * sumt;
      }
    }
  }
}

static constexpr size_t kMaxRingSmem = 160 * 1024;

// channel segments per location: 1 unless the location count cannot fill the GPU
static int pick_segments(long long L, int F, int k, bool blocked) {
  const long long blocks = ceil_div<long long>(L, RN_THREADS);
  const long long want = ceil_div<long long>(2LL * num_sms(), blocks);
  int segs = (int)std::min<long long>(want, std::max(1, F / std::max(k, 1)));      // segment >= k: halo <= 2x / 3x reads
  if (segs < 1) segs = 1;
  (void)blocked;
  return segs;
}

void rnorm_forward(const float* images, float* targets, long long L, int F, int k, float alpha, float beta,
                   bool blocked) {
  CNB_REQUIRE(k >= 1 && F >= 1, "ResponseNormCrossMap");
  const int blocks = (int)ceil_div<long long>(L, RN_THREADS);
  int segs = pick_segments(L, F, k, blocked);
  int seg = ceil_div(F, segs);
  if (blocked) seg = ceil_div(seg, k) * k;
  segs = ceil_div(F, seg);
  size_t smem = sizeof(float) * (size_t)k * RN_THREADS;
  float* gring = nullptr;
  if (smem > kMaxRingSmem) { gring = (float*)workspace(sizeof(float) * (size_t)k * L * segs); smem = 0; }
  auto kern = blocked ? rnorm_fwd_kernel<true> : rnorm_fwd_kernel<false>;
  if (smem > 48 * 1024) CNB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<dim3(blocks, segs), RN_THREADS, smem, state().stream>>>(images, targets, L, F, k, alpha, beta, gring, L, seg);
  count_launch();
  CNB_LAUNCH_CHECK("rnorm_forward");
}

void rnorm_undo(const float* outGrads, const float* inputs, float* targets, long long L, int F, int k,
                float alpha, float beta, bool blocked) {
  CNB_REQUIRE(k >= 1 && F >= 1, "ResponseNormCrossMapUndo");
  const int blocks = (int)ceil_div<long long>(L, RN_THREADS);
  int segs = pick_segments(L, F, k, blocked);
  int seg = ceil_div(F, segs);
  if (blocked) seg = ceil_div(seg, k) * k;
  segs = ceil_div(F, seg);
  size_t smem = sizeof(float) * 3 * (size_t)k * RN_THREADS;
  float* gring = nullptr;
  if (smem > kMaxRingSmem) { gring = (float*)workspace(sizeof(float) * 3 * (size_t)k * L * segs); smem = 0; }
  auto kern = blocked ? rnorm_undo_kernel<true> : rnorm_undo_kernel<false>;
  if (smem > 48 * 1024) CNB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<dim3(blocks, segs), RN_THREADS, smem, state().stream>>>(outGrads, inputs, targets, L, F, k, alpha, beta, gring, L, seg);
  count_launch();
  CNB_LAUNCH_CHECK("rnorm_undo");
}

}  // namespace cnb
