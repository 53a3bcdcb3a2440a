/* conv_oracle.c — see conv_oracle.h.  TEST INFRASTRUCTURE ONLY (CPU, plain C).
 *
 * Every function cites the reference lines it restates.  Arithmetic is float,
 * accumulated in the reference's order (module-by-module, K ascending inside
 * the naive sgemm of eigenmat/eigenmat.cc:2284-2296), so conv results match
 * eigenmat/cpumat_conv.cc bit for bit on finite inputs.  Deliberate, documented
 * deviations:
 *   - scaleTargets == 0 never reads the target (GPU semantics,
 *     cudamat_conv_gemm.cu:391-404,44-50); the CPU reference multiplies by 0.
 *   - untied ("local") filters: module i uses filter block i.  The reference
 *     advances the filter pointer BEFORE the first GEMM (cpumat_conv.cc:195-198,
 *     cudamat_conv_gemm.cu:657), i.e. block i+1, running off the end of the
 *     buffer at the last module; that is undefined behaviour, not a spec.
 */
#include "conv_oracle.h"

#include <assert.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define OMIN(a, b) ((a) < (b) ? (a) : (b))
#define OMAX(a, b) ((a) > (b) ? (a) : (b))

typedef struct {
  int N, W, H, modX, modY, Cin, Cout;     /* Cin/Cout = sub-range sizes */
  int ky, kx, sy, sx, py, px;             /* py/px are the (<=0) start offsets */
  int cin_begin, cout_begin;
  int K, modules;
} Geo;

static Geo make_geo(const Shape4D* img, const Shape4D* out, ConvDesc d) {
  Geo g;
  int ic_end = d.input_channel_end == 0 ? d.num_input_channels : d.input_channel_end;
  int oc_end = d.output_channel_end == 0 ? d.num_output_channels : d.output_channel_end;
  g.N = img->shape[0]; g.W = img->shape[1]; g.H = img->shape[2];
  g.modX = out->shape[1]; g.modY = out->shape[2];
  assert(img->shape[0] == out->shape[0]);
  assert(img->shape[3] == d.num_input_channels);
  assert(out->shape[3] == d.num_output_channels);
  assert(d.num_groups == 1);
  g.cin_begin = d.input_channel_begin; g.cout_begin = d.output_channel_begin;
  g.Cin = ic_end - d.input_channel_begin; g.Cout = oc_end - d.output_channel_begin;
  assert(g.Cin > 0 && g.Cout > 0);
  g.ky = d.kernel_size_y; g.kx = d.kernel_size_x;
  g.sy = d.stride_y; g.sx = d.stride_x; g.py = d.padding_y; g.px = d.padding_x;
  g.K = g.ky * g.kx * g.Cin; g.modules = g.modX * g.modY;
  return g;
}

/* cpumat_conv.cc:109-220 */
void oracle_convUp(const float* images, const float* filters, float* targets,
                   const Shape4D* is, const Shape4D* fs, const Shape4D* ts,
                   ConvDesc d, float scaleTargets, float scaleOutput, int conv) {
  Geo g = make_geo(is, ts, d);
  assert(fs->shape[0] == g.Cout && fs->shape[1] == g.kx && fs->shape[2] == g.ky);
  const size_t N = (size_t)g.N;
  const float* img = images + (size_t)g.cin_begin * g.H * g.W * N;
  float* tgt = targets + (size_t)g.cout_begin * g.modules * N;
  float* acc = (float*)malloc(sizeof(float) * N);
  for (int m = 0; m < g.modules; m++) {
    const int startX = (m % g.modX) * g.sx + g.px;
    const int startY = (m / g.modX) * g.sy + g.py;
    const float* w = filters + (conv ? 0 : (size_t)m * g.Cout * g.K);
    for (int o = 0; o < g.Cout; o++) {
      for (size_t n = 0; n < N; n++) acc[n] = 0.f;
      for (int c = 0; c < g.Cin; c++)
        for (int y = 0; y < g.ky; y++) {
          const int Y = startY + y;
          if (Y < 0 || Y >= g.H) continue;
          for (int x = 0; x < g.kx; x++) {
            const int X = startX + x;
            if (X < 0 || X >= g.W) continue;   /* expand() wrote 0 here: adds +-0 */
            const float wv = w[o + (size_t)g.Cout * (x + g.kx * (y + g.ky * c))];
            const float* src = img + N * (X + (size_t)g.W * (Y + (size_t)g.H * c));
            for (size_t n = 0; n < N; n++) acc[n] += src[n] * wv;
          }
        }
      float* t = tgt + N * (m + (size_t)g.modules * o);
      if (scaleTargets == 0.f) for (size_t n = 0; n < N; n++) t[n] = scaleOutput * acc[n];
      else for (size_t n = 0; n < N; n++) t[n] = scaleTargets * t[n] + scaleOutput * acc[n];
    }
  }
  free(acc);
}

static void scale_all(float* a, size_t n, float s) {
  if (s == 0.f) memset(a, 0, n * sizeof(float));
  else if (s != 1.f) for (size_t i = 0; i < n; i++) a[i] *= s;
}

/* cpumat_conv.cc:222-338 */
void oracle_convDown(const float* derivs, const float* filters, float* targets,
                     const Shape4D* ds, const Shape4D* fs, const Shape4D* ts,
                     ConvDesc d, float scaleTargets, float scaleOutput, int conv) {
  Geo g = make_geo(ts, ds, d);
  assert(fs->shape[0] == g.Cout && fs->shape[1] == g.kx && fs->shape[2] == g.ky);
  const size_t N = (size_t)g.N;
  const float* der = derivs + (size_t)g.cout_begin * g.modules * N;
  float* tgt = targets + (size_t)g.cin_begin * g.H * g.W * N;
  /* the reference scales the WHOLE target matrix (:304-307) */
  scale_all(targets, N * g.W * g.H * (size_t)d.num_input_channels, scaleTargets);
  float* acc = (float*)malloc(sizeof(float) * N);
  for (int m = 0; m < g.modules; m++) {
    const int startX = (m % g.modX) * g.sx + g.px;
    const int startY = (m / g.modX) * g.sy + g.py;
    const float* w = filters + (conv ? 0 : (size_t)m * g.Cout * g.K);
    for (int c = 0; c < g.Cin; c++)
      for (int y = 0; y < g.ky; y++) {
        const int Y = startY + y;
        if (Y < 0 || Y >= g.H) continue;
        for (int x = 0; x < g.kx; x++) {
          const int X = startX + x;
          if (X < 0 || X >= g.W) continue;
          const size_t k = x + g.kx * (y + g.ky * c);
          for (size_t n = 0; n < N; n++) acc[n] = 0.f;
          for (int o = 0; o < g.Cout; o++) {
            const float wv = w[o + (size_t)g.Cout * k];
            const float* src = der + N * (m + (size_t)g.modules * o);
            for (size_t n = 0; n < N; n++) acc[n] += src[n] * wv;
          }
          float* t = tgt + N * (X + (size_t)g.W * (Y + (size_t)g.H * c));
          for (size_t n = 0; n < N; n++) t[n] += scaleOutput * acc[n];
        }
      }
  }
  free(acc);
}

/* one module's contribution: dw[o + Cout*k] += scaleOutput * sum_n derivs * expanded */
static void outp_module(const Geo* g, const float* img, const float* der, float* dw,
                        int m, float scaleOutput) {
  const size_t N = (size_t)g->N;
  const int startX = (m % g->modX) * g->sx + g->px;
  const int startY = (m / g->modX) * g->sy + g->py;
  for (int c = 0; c < g->Cin; c++)
    for (int y = 0; y < g->ky; y++) {
      const int Y = startY + y;
      for (int x = 0; x < g->kx; x++) {
        const int X = startX + x;
        const size_t k = x + g->kx * (y + g->ky * c);
        if (Y < 0 || Y >= g->H || X < 0 || X >= g->W) continue; /* res = 0: dw unchanged */
        const float* src = img + N * (X + (size_t)g->W * (Y + (size_t)g->H * c));
        for (int o = 0; o < g->Cout; o++) {
          const float* dv = der + N * (m + (size_t)g->modules * o);
          float res = 0.f;
          for (size_t n = 0; n < N; n++) res += dv[n] * src[n];
          dw[o + (size_t)g->Cout * k] = dw[o + (size_t)g->Cout * k] + scaleOutput * res;
        }
      }
    }
}

/* cpumat_conv.cc:340-460 */
void oracle_convOutp(const float* images, const float* derivs, float* targets,
                     const Shape4D* is, const Shape4D* ds, const Shape4D* ts,
                     ConvDesc d, float scaleTargets, float scaleOutput, int conv) {
  Geo g = make_geo(is, ds, d);
  assert(ts->shape[0] == g.Cout && ts->shape[1] == g.kx && ts->shape[2] == g.ky);
  const size_t N = (size_t)g.N;
  const float* img = images + (size_t)g.cin_begin * g.H * g.W * N;
  const float* der = derivs + (size_t)g.cout_begin * g.modules * N;
  scale_all(targets, (size_t)g.Cout * g.K * (conv ? 1 : g.modules), scaleTargets);
  for (int m = 0; m < g.modules; m++)
    outp_module(&g, img, der, targets + (conv ? 0 : (size_t)m * g.Cout * g.K), m, scaleOutput);
}

/* py/conv_cpu.py:78-136 (output_psums) ; weightacts.cu:3126-3170 */
void oracle_convOutpPartial(const float* images, const float* derivs, float* targets,
                            const Shape4D* is, const Shape4D* ds, const Shape4D* ts,
                            ConvDesc d, int psY, int psX, float scaleTargets,
                            float scaleOutput) {
  Geo g = make_geo(is, ds, d);
  if (psY <= 0) psY = g.modY;
  if (psX <= 0) psX = g.modX;
  const int chX = (g.modX + psX - 1) / psX, chY = (g.modY + psY - 1) / psY;
  assert(ts->shape[0] == g.Cout && ts->shape[3] == g.Cin * chX * chY);
  scale_all(targets, (size_t)g.Cout * g.K * chX * chY, scaleTargets);
  for (int m = 0; m < g.modules; m++) {
    const int id = ((m / g.modX) / psY) * chX + (m % g.modX) / psX;
    outp_module(&g, images, derivs, targets + (size_t)id * g.Cout * g.K, m, scaleOutput);
  }
}

/* ---- 3-D: cudamat_conv3d_gemm.cu:13-165 --------------------------------- */
typedef struct { ConvDesc d2; Shape4D a2, b2; int modT; size_t in_frame, out_frame; } Geo3;

static Geo3 make_geo3(const Shape4D* img, const Shape4D* out, ConvDesc d) {
  Geo3 g;
  assert(d.padding_t == 0);                                     /* :23 */
  g.modT = out->shape[3] / d.num_output_channels;
  g.in_frame = (size_t)img->shape[1] * img->shape[2] * d.num_input_channels;
  g.out_frame = (size_t)out->shape[1] * out->shape[2] * d.num_output_channels;
  g.d2 = d;
  g.d2.kernel_size_t = 1;
  g.d2.num_input_channels *= d.kernel_size_t;
  g.d2.input_channel_end *= d.kernel_size_t;
  g.a2 = *img; g.b2 = *out;
  g.a2.shape[3] = d.num_input_channels * d.kernel_size_t;
  g.b2.shape[3] = d.num_output_channels;
  return g;
}

void oracle_convUp3D(const float* images, const float* filters, float* targets,
                     const Shape4D* is, const Shape4D* fs, const Shape4D* ts,
                     ConvDesc d, float scaleTargets) {
  Geo3 g = make_geo3(is, ts, d);
  const size_t N = (size_t)is->shape[0];
  for (int t = 0; t < g.modT; t++)
    oracle_convUp(images + g.in_frame * N * d.stride_t * t, filters,
                  targets + g.out_frame * N * t, &g.a2, fs, &g.b2, g.d2,
                  scaleTargets, 1.f, 1);
}

void oracle_convDown3D(const float* derivs, const float* filters, float* targets,
                       const Shape4D* ds, const Shape4D* fs, const Shape4D* ts,
                       ConvDesc d, float scaleTargets) {
  Geo3 g = make_geo3(ts, ds, d);
  const size_t N = (size_t)ds->shape[0];
  scale_all(targets, N * ts->shape[1] * ts->shape[2] * (size_t)ts->shape[3], scaleTargets);
  for (int t = 0; t < g.modT; t++)
    oracle_convDown(derivs + g.out_frame * N * t, filters,
                    targets + g.in_frame * N * d.stride_t * t, &g.b2, fs, &g.a2,
                    g.d2, 1.f, 1.f, 1);
}

void oracle_convOutp3D(const float* images, const float* derivs, float* targets,
                       const Shape4D* is, const Shape4D* ds, const Shape4D* ts,
                       ConvDesc d, float scaleTargets, float scaleOutput) {
  Geo3 g = make_geo3(is, ds, d);
  const size_t N = (size_t)is->shape[0];
  scale_all(targets, (size_t)ts->shape[0] * ts->shape[1] * ts->shape[2] * ts->shape[3],
            scaleTargets);
  for (int t = 0; t < g.modT; t++)
    oracle_convOutp(images + g.in_frame * N * d.stride_t * t,
                    derivs + g.out_frame * N * t, targets, &g.a2, &g.b2, ts, g.d2,
                    1.f, scaleOutput, 1);
}

/* ---- pooling: cudamat_conv_gemm.cu:153-300 -------------------------------- */
typedef struct {
  int N, W, H, T, C, modX, modY, modT;
  int kx, ky, kt, sx, sy, st, px, py, pt;
} PGeo;

static PGeo make_pgeo(const Shape4D* img, const Shape4D* out, ConvDesc d) {
  PGeo g;
  g.N = img->shape[0]; g.W = img->shape[1]; g.H = img->shape[2];
  g.C = d.num_input_channels;
  g.T = img->shape[3] / d.num_input_channels;                 /* :1155 */
  g.modX = out->shape[1]; g.modY = out->shape[2];
  g.modT = out->shape[3] / d.num_output_channels;             /* :1156 */
  assert(img->shape[0] == out->shape[0]);
  assert(d.num_input_channels == d.num_output_channels);
  g.kx = d.kernel_size_x; g.ky = d.kernel_size_y; g.kt = d.kernel_size_t;
  g.sx = d.stride_x; g.sy = d.stride_y; g.st = d.stride_t;
  g.px = d.padding_x; g.py = d.padding_y; g.pt = d.padding_t;
  return g;
}

#define IMG_IDX(g, X, Y, c, Tt) \
  ((size_t)(g).N * ((X) + (size_t)(g).W * ((Y) + (size_t)(g).H * ((c) + (size_t)(g).C * (Tt)))))
#define MOD_IDX(g, mx, my, c, mt) \
  ((size_t)(g).N * ((mx) + (size_t)(g).modX * ((my) + (size_t)(g).modY * ((c) + (size_t)(g).C * (mt)))))

void oracle_pool(int is_max, const float* images, float* targets,
                 const Shape4D* is, const Shape4D* ts, ConvDesc d, float scaleOutput) {
  PGeo g = make_pgeo(is, ts, d);
  for (int c = 0; c < g.C; c++)
    for (int mt = 0; mt < g.modT; mt++)
      for (int my = 0; my < g.modY; my++)
        for (int mx = 0; mx < g.modX; mx++) {
          int sX = mx * g.sx + g.px, sY = my * g.sy + g.py, sT = mt * g.st + g.pt;
          int eX = OMIN(sX + g.kx, g.W), eY = OMIN(sY + g.ky, g.H), eT = OMIN(sT + g.kt, g.T);
          sX = OMAX(sX, 0); sY = OMAX(sY, 0); sT = OMAX(sT, 0);
          const int region = (eX - sX) * (eY - sY) * (eT - sT);
          float* t = targets + MOD_IDX(g, mx, my, c, mt);
          for (int n = 0; n < g.N; n++) {
            float v = is_max ? -2e38f : 0.f;
            for (int T = sT; T < eT; T++)
              for (int Y = sY; Y < eY; Y++)
                for (int X = sX; X < eX; X++) {
                  const float a = images[IMG_IDX(g, X, Y, c, T) + n];
                  v = is_max ? fmaxf(v, a) : v + a;
                }
            t[n] = scaleOutput * (is_max ? v : v / region);
          }
        }
}

void oracle_maxPoolUndo(const float* images, const float* maxGrads, const float* maxActs,
                        float* targets, const Shape4D* is, const Shape4D* gs,
                        ConvDesc d, float scaleTargets) {
  PGeo g = make_pgeo(is, gs, d);
  scale_all(targets, (size_t)g.N * g.W * g.H * g.C * g.T, scaleTargets);
  for (int c = 0; c < g.C; c++)
    for (int mt = 0; mt < g.modT; mt++)
      for (int my = 0; my < g.modY; my++)
        for (int mx = 0; mx < g.modX; mx++) {
          int sX = mx * g.sx + g.px, sY = my * g.sy + g.py, sT = mt * g.st + g.pt;
          int eX = OMIN(sX + g.kx, g.W), eY = OMIN(sY + g.ky, g.H), eT = OMIN(sT + g.kt, g.T);
          sX = OMAX(sX, 0); sY = OMAX(sY, 0); sT = OMAX(sT, 0);
          const size_t mi = MOD_IDX(g, mx, my, c, mt);
          for (int T = sT; T < eT; T++)
            for (int Y = sY; Y < eY; Y++)
              for (int X = sX; X < eX; X++) {
                const size_t ii = IMG_IDX(g, X, Y, c, T);
                for (int n = 0; n < g.N; n++)
                  if (images[ii + n] == maxActs[mi + n]) targets[ii + n] += maxGrads[mi + n];
              }
        }
}

void oracle_avgPoolUndo(const float* avgGrads, float* targets, const Shape4D* gs,
                        const Shape4D* ts, ConvDesc d, float scaleTargets,
                        float scaleOutput) {
  PGeo g = make_pgeo(ts, gs, d);
  scale_all(targets, (size_t)g.N * g.W * g.H * g.C * g.T, scaleTargets);
  for (int c = 0; c < g.C; c++)
    for (int mt = 0; mt < g.modT; mt++)
      for (int my = 0; my < g.modY; my++)
        for (int mx = 0; mx < g.modX; mx++) {
          int sX = mx * g.sx + g.px, sY = my * g.sy + g.py, sT = mt * g.st + g.pt;
          int eX = OMIN(sX + g.kx, g.W), eY = OMIN(sY + g.ky, g.H), eT = OMIN(sT + g.kt, g.T);
          sX = OMAX(sX, 0); sY = OMAX(sY, 0); sT = OMAX(sT, 0);
          const int region = (eX - sX) * (eY - sY) * (eT - sT);
          const size_t mi = MOD_IDX(g, mx, my, c, mt);
          for (int T = sT; T < eT; T++)
            for (int Y = sY; Y < eY; Y++)
              for (int X = sX; X < eX; X++) {
                const size_t ii = IMG_IDX(g, X, Y, c, T);
                for (int n = 0; n < g.N; n++)
                  targets[ii + n] += scaleOutput * avgGrads[mi + n] / region;
              }
        }
}

/* ---- cross-map response norm: cpumat_conv.cc:462-560 ---------------------- */
void oracle_rnorm(const float* data, float* target, long num_els, int F, int sizeF,
                  float addScale, float powScale, int blocked) {
  const long L = num_els / F;
  for (long loc = 0; loc < L; loc++) {
    float sum = 0;
    int prev_start = 0, prev_end = 0, start, end;
    for (int j = 0; j < F; j++) {
      start = blocked ? (j / sizeF) * sizeF : -sizeF / 2 + j;
      end = OMIN(F, start + sizeF);
      start = OMAX(0, start);
      for (int i = prev_start; i < start; i++) { float v = data[i * L + loc]; sum -= v * v; }
      for (int i = prev_end; i < end; i++) { float v = data[i * L + loc]; sum += v * v; }
      target[j * L + loc] = data[j * L + loc] * powf(1 + addScale * sum, -powScale);
      prev_start = start; prev_end = end;
    }
  }
}

void oracle_rnormUndo(const float* deriv, const float* data, float* target, long num_els,
                      int F, int sizeF, float addScale, float powScale, int blocked) {
  const long L = num_els / F;
  float* denoms = (float*)malloc(sizeof(float) * (size_t)L * F);
  for (long loc = 0; loc < L; loc++) {
    float sum = 0;
    int prev_start = 0, prev_end = 0, start, end;
    for (int j = 0; j < F; j++) {
      start = blocked ? (j / sizeF) * sizeF : -sizeF / 2 + j;
      end = OMIN(F, start + sizeF);
      start = OMAX(0, start);
      for (int i = prev_start; i < start; i++) { float v = data[i * L + loc]; sum -= v * v; }
      for (int i = prev_end; i < end; i++) { float v = data[i * L + loc]; sum += v * v; }
      denoms[j * L + loc] = powf(1 + addScale * sum, -powScale - 1);
      prev_start = start; prev_end = end;
    }
  }
  for (long loc = 0; loc < L; loc++) {
    float sum = 0;
    int prev_start = 0, prev_end = 0, start, end;
    for (int j = 0; j < F; j++) {
      start = blocked ? (j / sizeF) * sizeF : -sizeF + sizeF / 2 + j + 1;
      end = OMIN(F, start + sizeF);
      start = OMAX(0, start);
      for (int i = prev_start; i < start; i++) {
        long idx = i * L + loc; sum -= deriv[idx] * data[idx] * denoms[idx];
      }
      for (int i = prev_end; i < end; i++) {
        long idx = i * L + loc; sum += deriv[idx] * data[idx] * denoms[idx];
      }
      long idx = j * L + loc;
      target[idx] = deriv[idx] * powf(denoms[idx], powScale / (powScale + 1)) -
                    2 * addScale * powScale * data[idx] * sum;
      prev_start = start; prev_end = end;
    }
  }
  free(denoms);
}

/* ---- input pipeline: minibatch crop / mirror (eigenmat/eigenmat.cc:2046-2090) ---------------------------------------- */
int oracle_extract_patches(const float* images, float* patches, const float* width_offset, const float* height_offset,
                           const float* flip, int num_images, int num_colors, int img_width, int img_height,
                           int patch_width, int patch_height) {
  for (long n = 0; n < num_images; n++) {
    const int x0 = (int)width_offset[n], y0 = (int)height_offset[n];
    const int mirror = flip[n] > 0.5f;
    for (long c = 0; c < num_colors; c++)
      for (long y = 0; y < patch_height; y++)
        for (long x = 0; x < patch_width; x++) {
          long sx = x0 + x;
          if (mirror) sx = img_width - sx - 1;
          const long sy = y0 + y;
          patches[n + num_images * (x + patch_width * (y + patch_height * c))] =
              images[sx + img_width * (sy + img_height * (c + num_colors * n))];
        }
  }
  return 0;
}
