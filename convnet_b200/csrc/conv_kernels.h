// conv_kernels.h — host-side entry points of the kernel translation units.
#pragma once
#include "common.cuh"
#include "geom.h"

namespace cnb {

// conv_simt.cu — fp32 CUDA-core implicit GEMM (exact mode + shapes the tensor path skips)
void simt_conv_up(const ConvGeom& g, const float* images, const float* filters, float* targets,
                  float scaleTargets, float scaleOutput, const Fuse& fuse);
void simt_conv_down(const ConvGeom& g, const float* derivs, const float* filters, float* targets,
                    float scaleTargets, float scaleOutput, const Fuse& fuse);
// rectH x rectW: module rectangle per reduction chunk; keep_partials: one output block per chunk
void simt_conv_outp(const ConvGeom& g, const float* images, const float* derivs, float* targets,
                    int rectH, int rectW, bool keep_partials, float scaleTargets, float scaleOutput);
void scale_buffer(float* a, long long n, float s);
void reduce_partials(const float* part, float* out, long long elems, int groups, int per, float st, float so);

// conv_tc.cu — tcgen05 / TMA implicit GEMM (sm_100a tensor cores)
bool tc_conv_up(const ConvGeom& g, const float* images, const float* filters, float* targets,
                float scaleTargets, float scaleOutput, const Fuse& fuse);
bool tc_conv_down(const ConvGeom& g, const float* derivs, const float* filters, float* targets,
                  float scaleTargets, float scaleOutput, const Fuse& fuse);
bool tc_conv_outp(const ConvGeom& g, const float* images, const float* derivs, float* targets,
                  float scaleTargets, float scaleOutput);

// bf16 operand staging (convnet_b200_bf16_stage / _invalidate); release drops the buffers too
void bf16_stage(const float* ptr, long long n);
void bf16_invalidate(const float* ptr);
void bf16_release();

// pool.cu
void pool_forward(const PoolGeom& g, bool is_max, const float* images, float* targets, float scaleOutput);
void max_pool_undo(const PoolGeom& g, const float* images, const float* maxGrads, const float* maxActs,
                   float* targets, float scaleTargets, float scaleOutput, const float* relu_mask);
void avg_pool_undo(const PoolGeom& g, const float* avgGrads, float* targets, float scaleTargets,
                   float scaleOutput, const float* relu_mask);

// rnorm.cu
void rnorm_forward(const float* images, float* targets, long long num_locs, int numFilters, int sizeF,
                   float addScale, float powScale, bool blocked);
void rnorm_undo(const float* outGrads, const float* inputs, float* targets, long long num_locs,
                int numFilters, int sizeF, float addScale, float powScale, bool blocked);

}  // namespace cnb
