// elementwise.cu — the memory-bound steps the Edge layer runs either side of the conv ops
// (bias add, bias gradient, ReLU, SGD).  See include/convnet_b200_ext.h for the reference
// call sites they correspond to.  All are single-pass, float4-vectorised where aligned.
#include <algorithm>

#include <cuda_bf16.h>

#include "../../include/convnet_b200_ext.h"
#include "conv_kernels.h"

namespace cnb {

// optional bf16 twin of a float4 result (convnet_b200_emit_bf16_next): 8 more bytes per thread, no extra pass
__device__ __forceinline__ void emit4(__nv_bfloat16* out16, long long i4, const float4& v) {
  if (!out16) return;
  const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
  uint2 o; o.x = *reinterpret_cast<const uint32_t*>(&lo); o.y = *reinterpret_cast<const uint32_t*>(&hi);
  reinterpret_cast<uint2*>(out16)[i4] = o;
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
static int blocks_for(long long work, int threads) {
  return (int)std::max<long long>(1, std::min<long long>(ceil_div<long long>(work, threads), (long long)num_sms() * 16));
}

template <bool RELU>
__global__ void bias_kernel(float* acts, const float* __restrict__ bias, long long rows, int cols) {
  // rows % 4 == 0 guaranteed by caller for the vector path
  const long long rv = rows / 4, total = rv * cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const float b = __ldg(bias + i / rv);
    float4 v = reinterpret_cast<float4*>(acts)[i];
    v.x += b; v.y += b; v.z += b; v.w += b;
    if (RELU) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
    reinterpret_cast<float4*>(acts)[i] = v;
  }
}
template <bool RELU>
__global__ void bias_kernel_scalar(float* acts, const float* __restrict__ bias, long long rows, int cols) {
  const long long total = rows * cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    float v = acts[i] + __ldg(bias + i / rows);
    acts[i] = RELU ? fmaxf(v, 0.f) : v;
  }
}

template <bool RELU>
static void bias_launch(float* acts, const float* bias, long long rows, int cols) {
  if (rows <= 0 || cols <= 0) return;
  if (rows % 4 == 0 && aligned16(acts)) {
    bias_kernel<RELU><<<blocks_for(rows / 4 * cols, 256), 256, 0, state().stream>>>(acts, bias, rows, cols);
  } else {
    bias_kernel_scalar<RELU><<<blocks_for(rows * cols, 256), 256, 0, state().stream>>>(acts, bias, rows, cols);
  }
  count_launch();
  CNB_LAUNCH_CHECK("add_channel_bias");
}

// column sums of a column-major [rows x cols] matrix; one block per (column, row-slice)
__global__ void colsum_partial_kernel(const float* __restrict__ a, float* __restrict__ part, long long rows, int cols,
                                      int slices) {
  const int col = blockIdx.x, slice = blockIdx.y;
  const long long per = ceil_div<long long>(rows, slices);
  const long long r0 = slice * per, r1 = min(rows, r0 + per);
  const float* p = a + (long long)col * rows;
  float s = 0.f;
  for (long long r = r0 + threadIdx.x; r < r1; r += blockDim.x) s += __ldg(p + r);
  __shared__ float sh[32];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.f;
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) part[(long long)slice * cols + col] = s;
  }
}
__global__ void colsum_final_kernel(const float* __restrict__ part, float* out, int cols, int slices, float st, float so) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  float s = 0.f;
  for (int j = 0; j < slices; j++) s += part[(long long)j * cols + c];
  out[c] = (st == 0.f) ? so * s : st * out[c] + so * s;
}

// n4 = number of float4 groups (vector body); the scalar tail [4*n4, n) is handled by the same launch
__global__ void relu_kernel(float* x, long long n, long long n4, __nv_bfloat16* out16) {
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x, nt = (long long)gridDim.x * blockDim.x;
  for (long long i = tid; i < n4; i += nt) {
    float4 v = reinterpret_cast<float4*>(x)[i];
    v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
    reinterpret_cast<float4*>(x)[i] = v;
    emit4(out16, i, v);
  }
  for (long long i = 4 * n4 + tid; i < n; i += nt) { const float v = fmaxf(x[i], 0.f); x[i] = v; if (out16) out16[i] = __float2bfloat16_rn(v); }
}
__global__ void relu_deriv_kernel(float* dx, const float* __restrict__ y, long long n, long long n4) {
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x, nt = (long long)gridDim.x * blockDim.x;
  for (long long i = tid; i < n4; i += nt) {
    float4 d = reinterpret_cast<float4*>(dx)[i];
    const float4 s = __ldg(reinterpret_cast<const float4*>(y) + i);
    d.x = s.x > 0.f ? d.x : 0.f; d.y = s.y > 0.f ? d.y : 0.f; d.z = s.z > 0.f ? d.z : 0.f; d.w = s.w > 0.f ? d.w : 0.f;
    reinterpret_cast<float4*>(dx)[i] = d;
  }
  for (long long i = 4 * n4 + tid; i < n; i += nt) dx[i] = y[i] > 0.f ? dx[i] : 0.f;
}
// Multi-tensor SGD with momentum and L2 decay: ONE launch updates every tensor of a batch (an all-reduce bucket, or the
// whole net).  A block owns kSgdChunk consecutive elements of one tensor; the block -> tensor map is a prefix table that
// travels in the kernel parameters (no device-side descriptor to keep coherent).  Tensors whose staged bf16 copy exists
// (conv weights in bf16 mode) get it refreshed from the same registers.
constexpr int kSgdMaxTensors = 48, kSgdChunk = 4096;
struct SgdItem { float* w; float* h; const float* g; __nv_bfloat16* w16; long long n; float lr, mom, l2; int vec; };
struct SgdBatch { int count; int first_block[kSgdMaxTensors + 1]; SgdItem t[kSgdMaxTensors]; };

__global__ void __launch_bounds__(256) sgd_multi_kernel(const __grid_constant__ SgdBatch b) {
  int ti = 0;
  while (ti + 1 < b.count && (int)blockIdx.x >= b.first_block[ti + 1]) ti++;      // <= 48 uniform steps
  const SgdItem& t = b.t[ti];
  const long long e0 = (long long)((int)blockIdx.x - b.first_block[ti]) * kSgdChunk;
  const long long e1 = min(t.n, e0 + kSgdChunk);
  if (t.vec) {                                                                     // all pointers 16-byte aligned
    for (long long i = e0 + 4 * threadIdx.x; i < e1; i += 4 * 256) {
      if (i + 4 <= e1) {
        float4 w = *reinterpret_cast<const float4*>(t.w + i), h = *reinterpret_cast<const float4*>(t.h + i);
        const float4 g = __ldg(reinterpret_cast<const float4*>(t.g + i));
        h.x = t.mom * h.x + t.lr * (g.x + t.l2 * w.x); w.x -= h.x;
        h.y = t.mom * h.y + t.lr * (g.y + t.l2 * w.y); w.y -= h.y;
        h.z = t.mom * h.z + t.lr * (g.z + t.l2 * w.z); w.z -= h.z;
        h.w = t.mom * h.w + t.lr * (g.w + t.l2 * w.w); w.w -= h.w;
        *reinterpret_cast<float4*>(t.h + i) = h;
        *reinterpret_cast<float4*>(t.w + i) = w;
        emit4(t.w16, i >> 2, w);
      } else {
        for (long long j = i; j < e1; j++) {
          const float wi = t.w[j], hi = t.mom * t.h[j] + t.lr * (t.g[j] + t.l2 * wi);
          t.h[j] = hi; t.w[j] = wi - hi;
          if (t.w16) t.w16[j] = __float2bfloat16_rn(wi - hi);
        }
      }
    }
  } else {
    for (long long i = e0 + threadIdx.x; i < e1; i += 256) {
      const float wi = t.w[i], hi = t.mom * t.h[i] + t.lr * (t.g[i] + t.l2 * wi);
      t.h[i] = hi; t.w[i] = wi - hi;
      if (t.w16) t.w16[i] = __float2bfloat16_rn(wi - hi);
    }
  }
}

// n4 float4 groups (vector body) + scalar tail, like relu_kernel; element i always draws from hash(seed + i)
__global__ void dropout_kernel(float* x, float* mask, long long n, long long n4, float dropprob, float scale,
                               unsigned long long seed, __nv_bfloat16* out16) {
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x, nt = (long long)gridDim.x * blockDim.x;
  for (long long i = tid; i < n4; i += nt) {
    float4 v = reinterpret_cast<float4*>(x)[i], m;
    const unsigned long long b = seed + 4ULL * (unsigned long long)i;
    m.x = hash_u32(b) * (1.0f / 4294967296.0f) >= dropprob ? scale : 0.f;
    m.y = hash_u32(b + 1) * (1.0f / 4294967296.0f) >= dropprob ? scale : 0.f;
    m.z = hash_u32(b + 2) * (1.0f / 4294967296.0f) >= dropprob ? scale : 0.f;
    m.w = hash_u32(b + 3) * (1.0f / 4294967296.0f) >= dropprob ? scale : 0.f;
    v.x *= m.x; v.y *= m.y; v.z *= m.z; v.w *= m.w;
    if (mask) reinterpret_cast<float4*>(mask)[i] = m;
    reinterpret_cast<float4*>(x)[i] = v;
    emit4(out16, i, v);
  }
  for (long long i = 4 * n4 + tid; i < n; i += nt) {
    const float u = hash_u32(seed + (unsigned long long)i) * (1.0f / 4294967296.0f);
    const float m = u >= dropprob ? scale : 0.f;
    if (mask) mask[i] = m;
    const float v = x[i] * m;
    x[i] = v;
    if (out16) out16[i] = __float2bfloat16_rn(v);
  }
}
__global__ void mult_kernel(float* a, const float* __restrict__ b, long long n, long long n4, __nv_bfloat16* out16) {
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x, nt = (long long)gridDim.x * blockDim.x;
  for (long long i = tid; i < n4; i += nt) {
    float4 x = reinterpret_cast<float4*>(a)[i];
    const float4 m = __ldg(reinterpret_cast<const float4*>(b) + i);
    x.x *= m.x; x.y *= m.y; x.z *= m.z; x.w *= m.w;
    reinterpret_cast<float4*>(a)[i] = x;
    emit4(out16, i, x);
  }
  for (long long i = 4 * n4 + tid; i < n; i += nt) { const float v = a[i] * b[i]; a[i] = v; if (out16) out16[i] = __float2bfloat16_rn(v); }
}

// softmax over classes of a column-major [rows x cols] matrix: one block per 32 images; lane = image (coalesced
// along rows), the block's warps split the classes; max and sum combined through shared memory.
__global__ void __launch_bounds__(256) softmax_kernel(float* x, int rows, int cols) {
  __shared__ float red[8][33];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + lane;
  const bool ok = n < rows;
  float m = -INFINITY;
  if (ok) for (int c = w; c < cols; c += 8) m = fmaxf(m, x[n + (long long)rows * c]);
  red[w][lane] = m;
  __syncthreads();
  for (int k = 0; k < 8; k++) m = fmaxf(m, red[k][lane]);
  __syncthreads();
  float s = 0.f;
  if (ok) for (int c = w; c < cols; c += 8) { const float e = expf(x[n + (long long)rows * c] - m); x[n + (long long)rows * c] = e; s += e; }
  red[w][lane] = s;
  __syncthreads();
  s = 0.f;
  for (int k = 0; k < 8; k++) s += red[k][lane];
  const float inv = 1.f / s;
  if (ok) for (int c = w; c < cols; c += 8) x[n + (long long)rows * c] *= inv;
}
__global__ void softmax_ce_deriv_kernel(const float* __restrict__ p, const int* __restrict__ labels, float* deriv,
                                        float* loss, int rows, int cols) {
  // blockIdx.y walks the classes: one thread per (image, class slice) instead of one per image (1 block for batch 128)
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= rows) return;
  const int lab = labels[n];
  for (int c = blockIdx.y; c < cols; c += gridDim.y)
    deriv[n + (long long)rows * c] = p[n + (long long)rows * c] - (c == lab ? 1.f : 0.f);
  if (loss && blockIdx.y == 0) loss[n] = -logf(fmaxf(p[n + (long long)rows * lab], 1e-30f));
}
__global__ void sum_kernel(const float* __restrict__ a, float* out, int n) {   // single block
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += a[i];
  __shared__ float sh[32];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.f;
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) *out = s;
  }
}

void colsum_finish(const float* part, float* grad_bias, int cols, int slices, float st, float so) {
  colsum_final_kernel<<<ceil_div(cols, 128), 128, 0, state().stream>>>(part, grad_bias, cols, slices, st, so);
  count_launch();
  CNB_LAUNCH_CHECK("colsum_finish");
}

// dropout without a mask tensor, writing the bf16 twin when given: the fallback of convnet_b200_fuse_next_dropout for calls
// whose kernel cannot apply it in the epilogue (the caller owns the staging bookkeeping of x)
void dropout_apply(float* x, long long n, float dropprob, float scale, unsigned long long seed, __nv_bfloat16* out16) {
  if (n <= 0) return;
  const long long n4 = aligned16(x) ? n / 4 : 0;
  dropout_kernel<<<blocks_for(std::max(n4, n - 4 * n4), 256), 256, 0, state().stream>>>(x, nullptr, n, n4, dropprob, scale, seed, out16);
  count_launch(); CNB_LAUNCH_CHECK("dropout_apply");
}

// crop + mirror + transpose of a minibatch out of an image-major chunk (convnet_b200_extract_patches).  A 32 x 32 tile of
// (image, patch column) for one (patch row, colour) goes through shared memory: the reads run along a source row (128
// contiguous bytes per image, reversed when mirrored), the writes along the images (the fastest axis of the layer state).
// The reference's kernel maps threads to patch columns and so stores with a stride of N floats (cudamat_kernels.cu:1655).
__global__ void __launch_bounds__(256) extract_patches_kernel(const float* __restrict__ images, float* __restrict__ patches,
                                                              const float* __restrict__ width_offset,
                                                              const float* __restrict__ height_offset,
                                                              const float* __restrict__ flip, int N, int W, int H, int pw, int ph,
                                                              int C) {
  __shared__ float tile[32][33];
  const int row = (int)(blockIdx.z % (unsigned)ph), color = (int)(blockIdx.z / (unsigned)ph);
  const int n0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  for (int k = threadIdx.y; k < 32; k += 8) {
    const int n = n0 + k, dc = c0 + (int)threadIdx.x;
    if (n < N && dc < pw) {
      int sr = (int)height_offset[n] + row, sc = (int)width_offset[n] + dc;
      if (flip[n] > 0.5f) sc = W - sc - 1;
      sr = min(max(sr, 0), H - 1); sc = min(max(sc, 0), W - 1);
      tile[k][threadIdx.x] = __ldg(images + sc + (size_t)W * (sr + (size_t)H * (color + (size_t)C * n)));
    }
  }
  __syncthreads();
  for (int k = threadIdx.y; k < 32; k += 8) {
    const int dc = c0 + k, n = n0 + (int)threadIdx.x;
    if (n < N && dc < pw) patches[n + (size_t)N * (dc + (size_t)pw * (row + (size_t)ph * color))] = tile[threadIdx.x][k];
  }
}
int extract_patches(const float* images, float* patches, const float* width_offset, const float* height_offset,
                    const float* flip, int N, int W, int H, int pw, int ph, int C) {
  if (N <= 0 || pw <= 0 || ph <= 0 || C <= 0) return 0;
  if ((long long)ph * C > 65535 || ceil_div(N, 32) > 65535) return -1;
  bf16_note_write(patches, (long long)N * pw * ph * C);
  const dim3 grid((unsigned)ceil_div(pw, 32), (unsigned)ceil_div(N, 32), (unsigned)(ph * C));
  extract_patches_kernel<<<grid, dim3(32, 8), 0, state().stream>>>(images, patches, width_offset, height_offset, flip, N, W, H, pw, ph, C);
  count_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : -3;
}
}  // namespace cnb

using namespace cnb;

extern "C" {

void cnb_add_channel_bias(float* acts, const float* bias, long long rows, int cols) {
  const bool emit = take_fuse().emit_bf16 != 0;
  begin_write(acts, rows * cols, emit, false);
  bias_launch<false>(acts, bias, rows, cols);
  end_write(acts, rows * cols, emit, nullptr);
}
void cnb_add_channel_bias_relu(float* acts, const float* bias, long long rows, int cols) {
  const bool emit = take_fuse().emit_bf16 != 0;
  begin_write(acts, rows * cols, emit, false);
  bias_launch<true>(acts, bias, rows, cols);
  end_write(acts, rows * cols, emit, nullptr);
}
// partial sums of the bias gradient live in their OWN scratch, not in the shared workspace: a host may run this pass on a
// side stream beside conv kernels that are using the workspace (host/convnet.cc does)
static float* colsum_scratch(size_t floats) {
  static float* buf = nullptr; static size_t cap = 0; static int dev = -1;
  const int cur = current_device();
  if (buf && (cap < floats || dev != cur)) {
    CNB_CUDA_CHECK(cudaDeviceSynchronize());
    if (dev != cur) CNB_CUDA_CHECK(cudaSetDevice(dev));
    CNB_CUDA_CHECK(cudaFree(buf));
    if (dev != cur) CNB_CUDA_CHECK(cudaSetDevice(cur));
    buf = nullptr; cap = 0;
  }
  if (!buf) {
    const size_t want = std::max<size_t>(floats, (size_t)1 << 18);
    CNB_CUDA_CHECK(cudaMalloc((void**)&buf, want * sizeof(float)));
    cap = want; dev = cur;
  }
  return buf;
}
void cnb_channel_bias_grad(const float* derivs, float* grad_bias, long long rows, int cols, float st, float so) {
  if (cols <= 0) return;
  int slices = (int)std::max<long long>(1, std::min<long long>(64, (4LL * num_sms()) / cols));
  slices = (int)std::min<long long>(slices, std::max<long long>(1, rows / 1024));
  float* part = colsum_scratch((size_t)slices * cols);
  colsum_partial_kernel<<<dim3(cols, slices), 256, 0, state().stream>>>(derivs, part, rows, cols, slices);
  colsum_final_kernel<<<ceil_div(cols, 128), 128, 0, state().stream>>>(part, grad_bias, cols, slices, st, so);
  count_launch(2);
  CNB_LAUNCH_CHECK("channel_bias_grad");
}
void cnb_relu(float* x, long long n) {
  const bool emit = take_fuse().emit_bf16 != 0;
  if (n <= 0) return;
  const long long n4 = aligned16(x) ? n / 4 : 0;
  __nv_bfloat16* o16 = begin_write(x, n, emit, true);
  relu_kernel<<<blocks_for(std::max(n4, n - 4 * n4), 256), 256, 0, state().stream>>>(x, n, n4, o16);
  count_launch(); CNB_LAUNCH_CHECK("relu");
  end_write(x, n, emit, o16);
}
void cnb_relu_deriv(float* dx, const float* y, long long n) {
  const bool emit = take_fuse().emit_bf16 != 0;
  if (n <= 0) return;
  begin_write(dx, n, emit, false);
  const long long n4 = (aligned16(dx) && aligned16(y)) ? n / 4 : 0;
  relu_deriv_kernel<<<blocks_for(std::max(n4, n - 4 * n4), 256), 256, 0, state().stream>>>(dx, y, n, n4);
  count_launch(); CNB_LAUNCH_CHECK("relu_deriv");
  end_write(dx, n, emit, nullptr);
}
void cnb_dropout(float* x, float* mask, long long n, float dropprob, float scale, unsigned long long seed) {
  const bool emit = take_fuse().emit_bf16 != 0;
  if (n <= 0) return;
  const long long n4 = (aligned16(x) && aligned16(mask)) ? n / 4 : 0;
  __nv_bfloat16* o16 = begin_write(x, n, emit, true);
  bf16_note_write(mask, n);
  dropout_kernel<<<blocks_for(std::max(n4, n - 4 * n4), 256), 256, 0, state().stream>>>(x, mask, n, n4, dropprob, scale, seed, o16);
  count_launch(); CNB_LAUNCH_CHECK("dropout");
  end_write(x, n, emit, o16);
}
void cnb_mult(float* a, const float* b, long long n) {
  const bool emit = take_fuse().emit_bf16 != 0;
  if (n <= 0) return;
  const long long n4 = (aligned16(a) && aligned16(b)) ? n / 4 : 0;
  __nv_bfloat16* o16 = begin_write(a, n, emit, true);
  mult_kernel<<<blocks_for(std::max(n4, n - 4 * n4), 256), 256, 0, state().stream>>>(a, b, n, n4, o16);
  count_launch(); CNB_LAUNCH_CHECK("mult");
  end_write(a, n, emit, o16);
}
void cnb_softmax(float* x, int rows, int cols) {
  if (rows <= 0) return;
  bf16_note_write(x, (long long)rows * cols);
  softmax_kernel<<<ceil_div(rows, 32), 256, 0, state().stream>>>(x, rows, cols);
  count_launch(); CNB_LAUNCH_CHECK("softmax");
}
void cnb_softmax_ce_deriv(const float* probs, const int* labels, float* deriv, float* loss_per_image, int rows, int cols) {
  if (rows <= 0) return;
  bf16_note_write(deriv, (long long)rows * cols);
  const dim3 grid(ceil_div(rows, 128), std::max(1, std::min(cols, 4 * num_sms() / std::max(1, ceil_div(rows, 128)))));
  softmax_ce_deriv_kernel<<<grid, 128, 0, state().stream>>>(probs, labels, deriv, loss_per_image, rows, cols);
  count_launch(); CNB_LAUNCH_CHECK("softmax_ce_deriv");
}
void cnb_sum(const float* a, float* out, int n) {
  sum_kernel<<<1, 256, 0, state().stream>>>(a, out, n);
  count_launch(); CNB_LAUNCH_CHECK("sum");
}
void cnb_sgd_momentum_multi(const CnbSgdTensor* tensors, int count) {
  for (int base = 0; base < count; base += kSgdMaxTensors) {
    SgdBatch b;
    b.count = 0;
    int blocks = 0;
    for (int i = base; i < count && b.count < kSgdMaxTensors; i++) {
      const CnbSgdTensor& s = tensors[i];
      if (s.n <= 0) continue;
      SgdItem& t = b.t[b.count];
      t.w = s.w; t.h = s.hist; t.g = s.grad; t.n = s.n; t.lr = s.lr; t.mom = s.momentum; t.l2 = s.l2;
      t.vec = (aligned16(s.w) && aligned16(s.hist) && aligned16(s.grad)) ? 1 : 0;
      // the weights change: a staged bf16 copy of exactly this tensor is refreshed in the same pass, any other overlap dropped
      const bool had_copy = bf16_staged(s.w, s.n) != nullptr;      // only a copy somebody keeps valid is worth refreshing
      bf16_note_write(s.w, s.n);                    // every derived copy (bf16 twin, dgrad banks) goes stale ...
      t.w16 = had_copy ? bf16_refresh_slot(s.w, s.n) : nullptr;    // ... and the bf16 twin is rewritten by this kernel
      b.first_block[b.count] = blocks;
      blocks += (int)ceil_div<long long>(s.n, kSgdChunk);
      b.count++;
    }
    if (b.count == 0) continue;
    b.first_block[b.count] = blocks;
    sgd_multi_kernel<<<blocks, 256, 0, state().stream>>>(b);
    count_launch(); CNB_LAUNCH_CHECK("sgd_momentum_multi");
  }
}
void cnb_sgd_momentum(float* w, float* hist, const float* grad, long long n, float lr, float momentum, float l2) {
  CnbSgdTensor t = {w, hist, grad, n, lr, momentum, l2};
  cnb_sgd_momentum_multi(&t, 1);
}

}  // extern "C"
