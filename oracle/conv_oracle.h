/* conv_oracle.h — CPU oracle for the conv / pool / response-norm hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is linked into, imported by
 * or called from the product (convnet_b200/, include/).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load it — as the checker, never as the thing shipped or measured as ours.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_reference.py checks every function
 * against (a) golden vectors generated here by importing the reference's own
 * py/conv_cpu.py (tests/golden/*.npz, generator tools/gen_golden.py) and (b)
 * the reference's compiled CPU library (oracle/_ref/libeigenmat_ref.so, built
 * from /root/reference/eigenmat by oracle/Makefile).
 *
 * Plain C, float arithmetic in the reference's own summation order, so that
 * conv results are bit-identical to eigenmat/cpumat_conv.cc where that file
 * defines the op.  All tensors are host arrays in the reference layout
 * (SURVEY.md Appendix A): element (n, f) of an N x F matrix at a[n + N*f].
 */
#ifndef CONV_ORACLE_H_
#define CONV_ORACLE_H_
#include "../include/cudamat_abi.h"

#ifdef __cplusplus
extern "C" {
#endif

/* follows eigenmat/cpumat_conv.cc:109-220 (convUp; expand :14-61, naive sgemm eigenmat.cc:2225-2296). conv=0 -> untied filters. */
void oracle_convUp(const float* images, const float* filters, float* targets,
                   const Shape4D* images_shape, const Shape4D* filters_shape,
                   const Shape4D* targets_shape, ConvDesc d,
                   float scaleTargets, float scaleOutput, int conv);
/* follows eigenmat/cpumat_conv.cc:222-338 (convDown; contract :63-107). */
void oracle_convDown(const float* derivs, const float* filters, float* targets,
                     const Shape4D* derivs_shape, const Shape4D* filters_shape,
                     const Shape4D* targets_shape, ConvDesc d,
                     float scaleTargets, float scaleOutput, int conv);
/* follows eigenmat/cpumat_conv.cc:340-460 (convOutp). */
void oracle_convOutp(const float* images, const float* derivs, float* targets,
                     const Shape4D* images_shape, const Shape4D* derivs_shape,
                     const Shape4D* targets_shape, ConvDesc d,
                     float scaleTargets, float scaleOutput, int conv);
/* wgrad with cuda-convnet partial sums; follows py/conv_cpu.py:78-136 and
 * cudamat/cudamat_conv_weightacts.cu:3126-3170 (target layout). */
void oracle_convOutpPartial(const float* images, const float* derivs, float* targets,
                            const Shape4D* images_shape, const Shape4D* derivs_shape,
                            const Shape4D* targets_shape, ConvDesc d,
                            int partialSumY, int partialSumX,
                            float scaleTargets, float scaleOutput);

/* 3-D: follow the frame loops of cudamat/cudamat_conv3d_gemm.cu:13-165 over the 2-D ops above
 * (== py/conv_cpu.py:523-647). */
void oracle_convUp3D(const float* images, const float* filters, float* targets,
                     const Shape4D* images_shape, const Shape4D* filters_shape,
                     const Shape4D* targets_shape, ConvDesc d, float scaleTargets);
void oracle_convDown3D(const float* derivs, const float* filters, float* targets,
                       const Shape4D* derivs_shape, const Shape4D* filters_shape,
                       const Shape4D* targets_shape, ConvDesc d, float scaleTargets);
void oracle_convOutp3D(const float* images, const float* derivs, float* targets,
                       const Shape4D* images_shape, const Shape4D* derivs_shape,
                       const Shape4D* targets_shape, ConvDesc d,
                       float scaleTargets, float scaleOutput);

/* follows cudamat/cudamat_conv_gemm.cu:153-200 (kPool) == src/CPUMatrix.cc:574-637,702-765;
 * is_max: 1 max (base -2e38), 0 avg (divide by clipped count). Overwrites targets. */
void oracle_pool(int is_max, const float* images, float* targets,
                 const Shape4D* images_shape, const Shape4D* targets_shape,
                 ConvDesc d, float scaleOutput);
/* follows cudamat/cudamat_conv_gemm.cu:252-300 (kMaxPoolUndo) == src/CPUMatrix.cc:640-700. */
void oracle_maxPoolUndo(const float* images, const float* maxGrads, const float* maxActs,
                        float* targets, const Shape4D* images_shape,
                        const Shape4D* maxGrads_shape, ConvDesc d, float scaleTargets);
/* follows cudamat/cudamat_conv_gemm.cu:203-249 (kAvgPoolUndo) == src/CPUMatrix.cc:768-827. */
void oracle_avgPoolUndo(const float* avgGrads, float* targets,
                        const Shape4D* avgGrads_shape, const Shape4D* targets_shape,
                        ConvDesc d, float scaleTargets, float scaleOutput);

/* follows eigenmat/cpumat_conv.cc:462-495; num_els = total elements of the matrix. */
void oracle_rnorm(const float* images, float* targets, long num_els, int numFilters,
                  int sizeF, float addScale, float powScale, int blocked);
/* follows eigenmat/cpumat_conv.cc:497-560. */
void oracle_rnormUndo(const float* outGrads, const float* inputs, float* targets,
                      long num_els, int numFilters, int sizeF, float addScale,
                      float powScale, int blocked);

/* crop + mirror + transpose of a minibatch out of an image-major chunk; follows eigenmat/eigenmat.cc:2046-2090
 * (= cudamat/cudamat_kernels.cu:1655-1669).  images: one image per column, pixel = col + W*(row + H*color);
 * patches: image fastest, n + N*(x + pw*(y + ph*c)).  Returns 0. */
int oracle_extract_patches(const float* images, float* patches, const float* width_offset, const float* height_offset,
                           const float* flip, int num_images, int num_colors, int img_width, int img_height,
                           int patch_width, int patch_height);

#ifdef __cplusplus
}
#endif
#endif
