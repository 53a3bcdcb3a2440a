// pool.cu — max / average pooling forward and backward (2-D and 3-D), HBM-bound.
//
// Replaces kPool / kMaxPoolUndo / kAvgPoolUndo (cudamat_conv_gemm.cu:153-300) and
// kLocalPool* / kLocalMaxUndo / kLocalAvgUndo (cudamat_conv_others.cu:1667-1864,3114-3437).
// Layout: images (N, W, H, C, T) with N fastest, so a thread
// owns VEC consecutive images of one (pixel, channel) and every load/store is a
// fully coalesced 16-byte access.  Backward passes are GATHERS over the windows that
// cover an input element: no atomics, deterministic (the reference scatters with
// atomicAdd + __syncthreads per tap).
#include <algorithm>

#include "conv_kernels.h"

namespace cnb {

template <int VEC> struct Vec;
template <> struct Vec<4> { using T = float4; };
template <> struct Vec<1> { using T = float; };

template <int VEC> __device__ __forceinline__ void vload(const float* p, float (&v)[VEC]);
template <> __device__ __forceinline__ void vload<4>(const float* p, float (&v)[4]) {
  const float4 t = __ldg(reinterpret_cast<const float4*>(p)); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <> __device__ __forceinline__ void vload<1>(const float* p, float (&v)[1]) { v[0] = __ldg(p); }
template <int VEC> __device__ __forceinline__ void vstore(float* p, const float (&v)[VEC]);
template <> __device__ __forceinline__ void vstore<4>(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
template <> __device__ __forceinline__ void vstore<1>(float* p, const float (&v)[1]) { *p = v[0]; }

// ---- forward -------------------------------------------------------------------------
template <int VEC, bool MAX>
__global__ void __launch_bounds__(256) pool_fwd_kernel(PoolGeom g, const float* __restrict__ images,
                                                        float* __restrict__ targets, float so, long long total) {
  const int NV = g.N / VEC;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int nv = (int)(idx % NV);
    long long r = idx / NV;
    const int mx = (int)(r % g.modX); r /= g.modX;
    const int my = (int)(r % g.modY); r /= g.modY;
    const int c = (int)(r % g.C);
    const int mt = (int)(r / g.C);
    int sX = mx * g.sx + g.px, sY = my * g.sy + g.py, sT = mt * g.st + g.pt;
    const int eX = min(sX + g.kx, g.W), eY = min(sY + g.ky, g.H), eT = min(sT + g.kt, g.T);
    sX = max(sX, 0); sY = max(sY, 0); sT = max(sT, 0);
    float acc[VEC];
#pragma unroll
    for (int v = 0; v < VEC; v++) acc[v] = MAX ? -2e38f : 0.f;     // base value: gemm.cu:71
    for (int T = sT; T < eT; T++)
      for (int Y = sY; Y < eY; Y++) {
        const float* row = images + (long long)g.N * ((long long)g.W * (Y + (long long)g.H * (c + (long long)g.C * T))) + nv * VEC;
        for (int X = sX; X < eX; X++) {
          float a[VEC];
          vload<VEC>(row + (long long)g.N * X, a);
#pragma unroll
          for (int v = 0; v < VEC; v++) acc[v] = MAX ? fmaxf(acc[v], a[v]) : acc[v] + a[v];
        }
      }
    if (!MAX) {
      const int region = (eX - sX) * (eY - sY) * (eT - sT);        // CLIPPED count: gemm.cu:185
#pragma unroll
      for (int v = 0; v < VEC; v++) acc[v] = acc[v] / region;
    }
#pragma unroll
    for (int v = 0; v < VEC; v++) acc[v] = so * acc[v];
    vstore<VEC>(targets + idx * VEC, acc);
  }
}

// ---- backward (gather) -----------------------------------------------------------------
// windows covering input coordinate X: m*s + p <= X < m*s + p + k
__device__ __forceinline__ void cover(int X, int s, int p, int k, int mods, int& lo, int& hi) {
  const int a = X - p - k + 1;                 // m*s >= a
  lo = a <= 0 ? 0 : (a + s - 1) / s;
  const int b = X - p;                         // m*s <= b   (b >= 0 whenever a window can cover X)
  hi = b < 0 ? -1 : min(b / s, mods - 1);
}

template <int VEC, bool MAX>
__global__ void __launch_bounds__(256) pool_undo_kernel(PoolGeom g, const float* __restrict__ images,
                                                         const float* __restrict__ grads,
                                                         const float* __restrict__ acts, float* targets,
                                                         float st, float so, long long total) {
  const int NV = g.N / VEC;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int nv = (int)(idx % NV);
    long long r = idx / NV;
    const int X = (int)(r % g.W); r /= g.W;
    const int Y = (int)(r % g.H); r /= g.H;
    const int c = (int)(r % g.C);
    const int T = (int)(r / g.C);
    int x0, x1, y0, y1, t0, t1;
    cover(X, g.sx, g.px, g.kx, g.modX, x0, x1);
    cover(Y, g.sy, g.py, g.ky, g.modY, y0, y1);
    cover(T, g.st, g.pt, g.kt, g.modT, t0, t1);
    float img[VEC], acc[VEC];
#pragma unroll
    for (int v = 0; v < VEC; v++) acc[v] = 0.f;
    if (MAX) vload<VEC>(images + idx * VEC, img);
    for (int mt = t0; mt <= t1; mt++)
      for (int my = y0; my <= y1; my++)
        for (int mx = x0; mx <= x1; mx++) {
          const long long off = (long long)g.N * (mx + (long long)g.modX * (my + (long long)g.modY * (c + (long long)g.C * mt))) + nv * VEC;
          float gr[VEC];
          vload<VEC>(grads + off, gr);
          if (MAX) {
            float a[VEC];
            vload<VEC>(acts + off, a);
#pragma unroll
            for (int v = 0; v < VEC; v++) acc[v] += (img[v] == a[v]) ? so * gr[v] : 0.f;   // ties duplicate: gemm.cu:291
          } else {
            int sX = mx * g.sx + g.px, sY = my * g.sy + g.py, sT = mt * g.st + g.pt;
            const int eX = min(sX + g.kx, g.W), eY = min(sY + g.ky, g.H), eT = min(sT + g.kt, g.T);
            sX = max(sX, 0); sY = max(sY, 0); sT = max(sT, 0);
            const int region = (eX - sX) * (eY - sY) * (eT - sT);
#pragma unroll
            for (int v = 0; v < VEC; v++) acc[v] += so * gr[v] / region;                  // gemm.cu:237
          }
        }
    if (st != 0.f) {
      float t[VEC];
      vload<VEC>(targets + idx * VEC, t);
#pragma unroll
      for (int v = 0; v < VEC; v++) acc[v] += st * t[v];
    }
    vstore<VEC>(targets + idx * VEC, acc);
  }
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

static int grid_for(long long total) {
  const long long want = ceil_div<long long>(total, 256);
  return (int)std::min<long long>(want, (long long)num_sms() * 16);   // multiple of the SM count when large
}

void pool_forward(const PoolGeom& g, bool is_max, const float* images, float* targets, float so) {
  const bool v4 = (g.N % 4 == 0) && aligned16(images) && aligned16(targets);
  const long long outs = (long long)g.modX * g.modY * g.C * g.modT;
  cudaStream_t s = state().stream;
  if (v4) {
    const long long total = outs * (g.N / 4);
    if (is_max) pool_fwd_kernel<4, true><<<grid_for(total), 256, 0, s>>>(g, images, targets, so, total);
    else pool_fwd_kernel<4, false><<<grid_for(total), 256, 0, s>>>(g, images, targets, so, total);
  } else {
    const long long total = outs * g.N;
    if (is_max) pool_fwd_kernel<1, true><<<grid_for(total), 256, 0, s>>>(g, images, targets, so, total);
    else pool_fwd_kernel<1, false><<<grid_for(total), 256, 0, s>>>(g, images, targets, so, total);
  }
  count_launch();
  CNB_LAUNCH_CHECK("pool_forward");
}

static void undo(const PoolGeom& g, bool is_max, const float* images, const float* grads, const float* acts,
                 float* targets, float st, float so) {
  const bool v4 = (g.N % 4 == 0) && aligned16(grads) && aligned16(targets) &&
                  (!is_max || (aligned16(images) && aligned16(acts)));
  const long long ins = (long long)g.W * g.H * g.C * g.T;
  cudaStream_t s = state().stream;
  if (v4) {
    const long long total = ins * (g.N / 4);
    if (is_max) pool_undo_kernel<4, true><<<grid_for(total), 256, 0, s>>>(g, images, grads, acts, targets, st, so, total);
    else pool_undo_kernel<4, false><<<grid_for(total), 256, 0, s>>>(g, images, grads, acts, targets, st, so, total);
  } else {
    const long long total = ins * g.N;
    if (is_max) pool_undo_kernel<1, true><<<grid_for(total), 256, 0, s>>>(g, images, grads, acts, targets, st, so, total);
    else pool_undo_kernel<1, false><<<grid_for(total), 256, 0, s>>>(g, images, grads, acts, targets, st, so, total);
  }
  count_launch();
  CNB_LAUNCH_CHECK("pool_undo");
}

void max_pool_undo(const PoolGeom& g, const float* images, const float* maxGrads, const float* maxActs,
                   float* targets, float st, float so) {
  undo(g, true, images, maxGrads, maxActs, targets, st, so);
}

void avg_pool_undo(const PoolGeom& g, const float* avgGrads, float* targets, float st, float so) {
  undo(g, false, nullptr, avgGrads, nullptr, targets, st, so);
}

}  // namespace cnb
