/* ref_shim.cc — extern "C" doors onto the REFERENCE's own CPU implementation
 * (eigenmat/cpumat_conv.cc), compiled from the sources where they lie under
 * /root/reference by oracle/Makefile into oracle/_ref/libeigenmat_ref.so.
 * TEST INFRASTRUCTURE ONLY: used to pin oracle/conv_oracle.c and as the timed
 * CPU baseline (bench.py cpu_baseline.kind == "reference").  No reference
 * source is copied into this repo; this file only calls the reference's API
 * (eigenmat/cpumat_conv.h:6-22) through plain pointers.
 */
#include "cpumat_conv.h"   // -I/root/reference/eigenmat at build time
#include "eigenmat.h"      // extract_patches (eigenmat.cc:2046)

#include <cstdlib>
#include <new>

// The reference's convUp/convDown run their naive sgemm with beta = 0 over scratch
// obtained from `new float[]` (cpumat_conv.cc:184-185,301-302; eigenmat.cc:2293:
// C = beta*C + alpha*res), i.e. they compute 0 * <uninitialised memory>, which is NaN
// whenever the recycled heap block holds a NaN/Inf bit pattern.  To make the
// reference deterministic WITHOUT touching its sources, array allocations made by
// THIS shared object are zero-filled (these definitions bind only inside this .so).
void* operator new[](std::size_t n) {
  void* p = std::calloc(n ? n : 1, 1);
  if (!p) throw std::bad_alloc();
  return p;
}
void operator delete[](void* p) noexcept { std::free(p); }
void operator delete[](void* p, std::size_t) noexcept { std::free(p); }

static eigenmat wrap(float* p, int rows, int cols) {
  eigenmat m; m.data = p; m.size[0] = rows; m.size[1] = cols; m.is_trans = 0; m.owns_data = 0;
  return m;
}
static int cols_of(const int* s) { return s[1] * s[2] * s[3]; }

extern "C" {

void ref_convUp(float* images, float* filters, float* targets, int* is, int* fs, int* ts,
                ConvDesc d, float scaleTargets, float scaleOutput, int conv) {
  Shape4D a, b, c; for (int i = 0; i < 4; i++) { a.shape[i] = is[i]; b.shape[i] = fs[i]; c.shape[i] = ts[i]; }
  eigenmat A = wrap(images, is[0], cols_of(is)), B = wrap(filters, fs[0], cols_of(fs)),
           C = wrap(targets, ts[0], cols_of(ts));
  convUp(&A, &B, &C, a, b, c, d, scaleTargets, scaleOutput, conv != 0);
}

void ref_convDown(float* derivs, float* filters, float* targets, int* ds, int* fs, int* ts,
                  ConvDesc d, float scaleTargets, float scaleOutput, int conv) {
  Shape4D a, b, c; for (int i = 0; i < 4; i++) { a.shape[i] = ds[i]; b.shape[i] = fs[i]; c.shape[i] = ts[i]; }
  eigenmat A = wrap(derivs, ds[0], cols_of(ds)), B = wrap(filters, fs[0], cols_of(fs)),
           C = wrap(targets, ts[0], cols_of(ts));
  convDown(&A, &B, &C, a, b, c, d, scaleTargets, scaleOutput, conv != 0);
}

void ref_convOutp(float* images, float* derivs, float* targets, int* is, int* ds, int* ts,
                  ConvDesc d, float scaleTargets, float scaleOutput, int conv) {
  Shape4D a, b, c; for (int i = 0; i < 4; i++) { a.shape[i] = is[i]; b.shape[i] = ds[i]; c.shape[i] = ts[i]; }
  eigenmat A = wrap(images, is[0], cols_of(is)), B = wrap(derivs, ds[0], cols_of(ds)),
           C = wrap(targets, ts[0], cols_of(ts));
  convOutp(&A, &B, &C, a, b, c, d, scaleTargets, scaleOutput, conv != 0);
}

void ref_rnorm(float* images, float* targets, int rows, int cols, int numFilters, int sizeF,
               float addScale, float powScale, int blocked) {
  eigenmat A = wrap(images, rows, cols), C = wrap(targets, rows, cols);
  ResponseNormCrossMap(&A, &C, numFilters, sizeF, addScale, powScale, blocked != 0);
}

void ref_rnormUndo(float* outGrads, float* inputs, float* targets, int rows, int cols,
                   int numFilters, int sizeF, float addScale, float powScale, int blocked) {
  eigenmat G = wrap(outGrads, rows, cols), A = wrap(inputs, rows, cols), C = wrap(targets, rows, cols);
  ResponseNormCrossMapUndo(&G, &A, &C, numFilters, sizeF, addScale, powScale, blocked != 0);
}

// the reference's CPU crop / mirror of a minibatch (eigenmat/eigenmat.cc:2046-2090); images: (C*W*H) x N, patches: N x (C*pw*ph)
int ref_extract_patches(float* images, float* patches, float* width_offset, float* height_offset, float* flip,
                        int num_images, int num_colors, int img_width, int img_height, int patch_width, int patch_height) {
  eigenmat I = wrap(images, num_colors * img_width * img_height, num_images),
           P = wrap(patches, num_images, num_colors * patch_width * patch_height),
           X = wrap(width_offset, 1, num_images), Y = wrap(height_offset, 1, num_images), F = wrap(flip, 1, num_images);
  return extract_patches(&I, &P, &X, &Y, &F, img_width, img_height, patch_width, patch_height);
}

}  // extern "C"
