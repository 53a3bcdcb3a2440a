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

// VEC consecutive locations per thread (16-byte accesses when VEC == 4): the scalar version spent ~70 instructions per
// channel step on one float (ncu: issue-bound at 76 % of the issue slots, 2.4 TB/s); the vector version shares the ring
// bookkeeping, branches and address arithmetic between four values.
template <int VEC> __device__ __forceinline__ void rld(const float* p, float (&v)[VEC]);
template <> __device__ __forceinline__ void rld<4>(const float* p, float (&v)[4]) {
  const float4 t = __ldg(reinterpret_cast<const float4*>(p)); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <> __device__ __forceinline__ void rld<1>(const float* p, float (&v)[1]) { v[0] = __ldg(p); }
template <int VEC> __device__ __forceinline__ void rget(const float* p, float (&v)[VEC]);      // ring (smem / scratch) read
template <> __device__ __forceinline__ void rget<4>(const float* p, float (&v)[4]) {
  const float4 t = *reinterpret_cast<const float4*>(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <> __device__ __forceinline__ void rget<1>(const float* p, float (&v)[1]) { v[0] = *p; }
template <int VEC> __device__ __forceinline__ void rput(float* p, const float (&v)[VEC]);
template <> __device__ __forceinline__ void rput<4>(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
template <> __device__ __forceinline__ void rput<1>(float* p, const float (&v)[1]) { *p = v[0]; }
#define RN_V _Pragma("unroll") for (int v = 0; v < VEC; v++)

// ring element (slot, thread): ring[slot * ring_stride + lane].
// blockIdx.y selects a channel SEGMENT [f0, f1) (host: only when there are too few locations to fill the GPU); a
// segment re-reads the k-1 (forward) / 2(k-1) (backward) halo channels of its neighbours instead of waiting for them.
template <bool BLOCKED, int VEC>
__global__ void __launch_bounds__(RN_THREADS) rnorm_fwd_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                                long long L, int F, int k, float alpha, float beta,
                                                                float* gring, long long gstride, int seg) {
  // L, gstride and all ring strides are in FLOATS; a thread owns floats [loc, loc + VEC)
  extern __shared__ __align__(16) float sring[];
  const long long loc = (blockIdx.x * (long long)RN_THREADS + threadIdx.x) * VEC;
  if (loc >= L) return;
  float* ring = gring ? gring + loc + (long long)blockIdx.y * k * gstride : sring + threadIdx.x * VEC;
  const long long rs = gring ? gstride : RN_THREADS * VEC;
  x += loc; y += loc;
  const int f0 = blockIdx.y * seg, f1 = min(F, f0 + seg);
  if (BLOCKED) {                                        // host guarantees seg % k == 0
    for (int s = f0; s < f1; s += k) {
      const int e = min(F, s + k);
      float sum[VEC];
      RN_V sum[v] = 0.f;
      for (int i = s; i < e; i++) {
        float t[VEC]; rld<VEC>(x + (long long)i * L, t); rput<VEC>(ring + (i - s) * rs, t);
        RN_V sum[v] += t[v] * t[v];
      }
      float sc[VEC];
      RN_V sc[v] = __powf(1.f + alpha * sum[v], -beta);
      for (int i = s; i < e; i++) {
        float t[VEC]; rget<VEC>(ring + (i - s) * rs, t);
        RN_V t[v] *= sc[v];
        rput<VEC>(y + (long long)i * L, t);
      }
    }
    return;
  }
  const int a = k / 2, b = k - a - 1;
  float sum[VEC];
  RN_V sum[v] = 0.f;
  // q = entering channel; output channel j = q - b; window [j-a, j+b] = [q-k+1, q].
  // Loads are hoisted four steps ahead of the (serial) ring updates to keep HBM requests in flight.
  constexpr int U = VEC == 4 ? 4 : 8;
  const int q0 = max(0, f0 - a), q1 = f1 + b;      // channels >= F enter as zeros
  int slot_q = q0 % k, slot_j = ((q0 - b) % k + k) % k;           // ring slots of q and of j = q - b, advanced with wrap
  for (int qb = q0; qb < q1; qb += U) {
    float xv[U][VEC];
#pragma unroll
    for (int u = 0; u < U; u++) {
      if (qb + u < F) rld<VEC>(x + (long long)(qb + u) * L, xv[u]);
      else { RN_V xv[u][v] = 0.f; }
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      const int q = qb + u;
      if (q >= q1) break;
      const int slot = slot_q;
      float old[VEC];
      RN_V old[v] = 0.f;
      if (q - q0 >= k) rget<VEC>(ring + slot * rs, old);
      if (q < F) rput<VEC>(ring + slot * rs, xv[u]);
      RN_V sum[v] += xv[u][v] * xv[u][v] - old[v] * old[v];
      const int j = q - b;
      if (j >= f0 && j < f1) {
        float xj[VEC];
        if (j == q) { RN_V xj[v] = xv[u][v]; } else rget<VEC>(ring + slot_j * rs, xj);
        RN_V xj[v] *= __powf(1.f + alpha * sum[v], -beta);
        rput<VEC>(y + (long long)j * L, xj);
      }
      if (++slot_q == k) slot_q = 0;
      if (++slot_j == k) slot_j = 0;
    }
  }
}

template <bool BLOCKED, int VEC>
__global__ void __launch_bounds__(RN_THREADS) rnorm_undo_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                                 float* __restrict__ dx, long long L, int F, int k,
                                                                 float alpha, float beta, float* gring, long long gstride,
                                                                 int seg) {
  // three rings of k entries per thread: x, t = dy*x*denom, p = dy*denom^(beta/(beta+1))
  extern __shared__ __align__(16) float sring[];
  const long long loc = (blockIdx.x * (long long)RN_THREADS + threadIdx.x) * VEC;
  if (loc >= L) return;
  const long long rs = gring ? gstride : RN_THREADS * VEC;
  float* rx = gring ? gring + loc + (long long)blockIdx.y * 3 * k * gstride : sring + threadIdx.x * VEC;
  float* rt = rx + (long long)k * rs;
  float* rp = rt + (long long)k * rs;
  x += loc; dy += loc; dx += loc;
  const float c2 = 2.f * alpha * beta;
  const int f0 = blockIdx.y * seg, f1 = min(F, f0 + seg);
  if (BLOCKED) {
    for (int s = f0; s < f1; s += k) {
      const int e = min(F, s + k);
      float sum[VEC];
      RN_V sum[v] = 0.f;
      for (int i = s; i < e; i++) {
        float t[VEC]; rld<VEC>(x + (long long)i * L, t); rput<VEC>(rx + (i - s) * rs, t);
        RN_V sum[v] += t[v] * t[v];
      }
      float denom[VEC], pw[VEC], st[VEC];
      RN_V { const float base = 1.f + alpha * sum[v]; denom[v] = __powf(base, -beta - 1.f); pw[v] = __powf(base, -beta); st[v] = 0.f; }
      for (int i = s; i < e; i++) {
        float g[VEC], xi[VEC]; rld<VEC>(dy + (long long)i * L, g); rget<VEC>(rx + (i - s) * rs, xi);
        rput<VEC>(rt + (i - s) * rs, g);
        RN_V st[v] += g[v] * xi[v] * denom[v];
      }
      for (int i = s; i < e; i++) {
        float g[VEC], xi[VEC]; rget<VEC>(rt + (i - s) * rs, g); rget<VEC>(rx + (i - s) * rs, xi);
        RN_V g[v] = g[v] * pw[v] - c2 * xi[v] * st[v];
        rput<VEC>(dx + (long long)i * L, g);
      }
    }
    return;
  }
  const int a = k / 2, b = k - a - 1;
  float sumsq[VEC], sumt[VEC];
  RN_V { sumsq[v] = 0.f; sumt[v] = 0.f; }
  // stage 1: entering channel q; channel i = q - b gets its forward sum, t_i and p_i
  // stage 2: output channel j = i - a gets sum of t over [j-b, j+a] = [i-k+1, i]
  const int q0 = max(0, f0 - (k - 1));                   // first x needed: (f0 - b) - a
  const int i0 = max(0, f0 - b);                         // first t needed
  const int Qend = f1 + a + b;                           // last output f1-1 needs t up to f1-1+a, i.e. q up to f1-1+a+b
  // ring slots of q, i = q - b and j = i - a, advanced with wrap instead of three `% k` per channel
  int slot_q = q0 % k, slot_i = ((q0 - b) % k + k) % k, slot_j = ((q0 - b - a) % k + k) % k;
  constexpr int U = VEC == 4 ? 2 : 4;
  for (int qb = q0; qb < Qend; qb += U) {
    float xv[U][VEC], gv[U][VEC];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const int q = qb + u, i = q - b;
      if (q < F) rld<VEC>(x + (long long)q * L, xv[u]); else { RN_V xv[u][v] = 0.f; }
      if (i >= i0 && i < F) rld<VEC>(dy + (long long)i * L, gv[u]); else { RN_V gv[u][v] = 0.f; }
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      const int q = qb + u;
      if (q >= Qend) break;
      const int sq = slot_q, si = slot_i, sj = slot_j;
      if (++slot_q == k) slot_q = 0;
      if (++slot_i == k) slot_i = 0;
      if (++slot_j == k) slot_j = 0;
      {
        float old[VEC];
        RN_V old[v] = 0.f;
        if (q - q0 >= k) rget<VEC>(rx + sq * rs, old);
        rput<VEC>(rx + sq * rs, xv[u]);                 // zeros once q >= F
        RN_V sumsq[v] += xv[u][v] * xv[u][v] - old[v] * old[v];
      }
      const int i = q - b;
      if (i < i0) continue;
      {
        float told[VEC], t[VEC];
        RN_V { told[v] = 0.f; t[v] = 0.f; }
        if (i - i0 >= k) rget<VEC>(rt + si * rs, told);
        if (i < F) {
          float xi[VEC], pv[VEC];
          rget<VEC>(rx + si * rs, xi);
          RN_V {
            const float base = 1.f + alpha * sumsq[v];
            const float denom = __powf(base, -beta - 1.f);
            t[v] = gv[u][v] * xi[v] * denom;
            pv[v] = gv[u][v] * denom * base;           // = g * base^(-beta)  (== denom^(beta/(beta+1)), gemm.cu:538)
          }
          rput<VEC>(rp + si * rs, pv);
        }
        rput<VEC>(rt + si * rs, t);
        RN_V sumt[v] += t[v] - told[v];
      }
      const int j = i - a;
      if (j >= f0 && j < f1) {
        float pj[VEC], xj[VEC];
        rget<VEC>(rp + sj * rs, pj); rget<VEC>(rx + sj * rs, xj);
        RN_V pj[v] -= c2 * xj[v] * sumt[v];
        rput<VEC>(dx + (long long)j * L, pj);
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

static inline bool rn_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <int VEC>
static void launch_fwd(const float* images, float* targets, long long L, int F, int k, float alpha, float beta, bool blocked) {
  const long long owners = L / VEC;                      // threads needed
  const int blocks = (int)ceil_div<long long>(owners, RN_THREADS);
  int segs = pick_segments(owners, F, k, blocked);
  int seg = ceil_div(F, segs);
  if (blocked) seg = ceil_div(seg, k) * k;
  segs = ceil_div(F, seg);
  size_t smem = sizeof(float) * (size_t)k * RN_THREADS * VEC;
  float* gring = nullptr;
  if (smem > kMaxRingSmem) { gring = (float*)workspace(sizeof(float) * (size_t)k * L * segs); smem = 0; }
  auto kern = blocked ? rnorm_fwd_kernel<true, VEC> : rnorm_fwd_kernel<false, VEC>;
  if (smem > 48 * 1024) CNB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<dim3(blocks, segs), RN_THREADS, smem, state().stream>>>(images, targets, L, F, k, alpha, beta, gring, L, seg);
}

void rnorm_forward(const float* images, float* targets, long long L, int F, int k, float alpha, float beta,
                   bool blocked) {
  CNB_REQUIRE(k >= 1 && F >= 1, "ResponseNormCrossMap");
  // four locations per thread only when that still leaves >= 4 blocks per SM: the channel walk is a serial dependency
  // chain, so small problems need the thread count more than the shorter instruction stream (measured: 105 -> 75 us on
  // 96 x 55 x 55 x 128, but 39 -> 47 us on 256 x 14 x 14 x 128)
  const bool wide = L % 4 == 0 && rn_aligned16(images) && rn_aligned16(targets) && L / 4 / RN_THREADS >= 4LL * num_sms();
  if (wide) launch_fwd<4>(images, targets, L, F, k, alpha, beta, blocked);
  else launch_fwd<1>(images, targets, L, F, k, alpha, beta, blocked);
  count_launch();
  CNB_LAUNCH_CHECK("rnorm_forward");
}

template <int VEC>
static void launch_undo(const float* outGrads, const float* inputs, float* targets, long long L, int F, int k, float alpha,
                        float beta, bool blocked) {
  const long long owners = L / VEC;
  const int blocks = (int)ceil_div<long long>(owners, RN_THREADS);
  int segs = pick_segments(owners, F, k, blocked);
  int seg = ceil_div(F, segs);
  if (blocked) seg = ceil_div(seg, k) * k;
  segs = ceil_div(F, seg);
  size_t smem = sizeof(float) * 3 * (size_t)k * RN_THREADS * VEC;
  float* gring = nullptr;
  if (smem > kMaxRingSmem) { gring = (float*)workspace(sizeof(float) * 3 * (size_t)k * L * segs); smem = 0; }
  auto kern = blocked ? rnorm_undo_kernel<true, VEC> : rnorm_undo_kernel<false, VEC>;
  if (smem > 48 * 1024) CNB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<dim3(blocks, segs), RN_THREADS, smem, state().stream>>>(outGrads, inputs, targets, L, F, k, alpha, beta, gring, L, seg);
}

void rnorm_undo(const float* outGrads, const float* inputs, float* targets, long long L, int F, int k,
                float alpha, float beta, bool blocked) {
  CNB_REQUIRE(k >= 1 && F >= 1, "ResponseNormCrossMapUndo");
  // the backward walk carries three rings and two dependent stages per channel: it is latency-bound, and the vector
  // version (a quarter of the threads, 4x the shared memory per block) measured SLOWER (210 -> 333 us); opt-in only
  static const bool wide_undo = getenv("CONVNET_B200_RNORM_UNDO_VEC4") && getenv("CONVNET_B200_RNORM_UNDO_VEC4")[0] == '1';
  if (wide_undo && L % 4 == 0 && rn_aligned16(outGrads) && rn_aligned16(inputs) && rn_aligned16(targets))
    launch_undo<4>(outGrads, inputs, targets, L, F, k, alpha, beta, blocked);
  else launch_undo<1>(outGrads, inputs, targets, L, F, k, alpha, beta, blocked);
  count_launch();
  CNB_LAUNCH_CHECK("rnorm_undo");
}

}  // namespace cnb
