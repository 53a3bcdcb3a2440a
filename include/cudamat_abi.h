/* cudamat_abi.h — the three POD types that cross the reference's C ABI for the
 * conv / pool / response-norm path.
 *
 * These restate the binary layout the reference fixes in
 *   cudamat/cudamat.cuh:28-37   (struct cudamat, 48 bytes on x86-64)
 *   cudamat/cudamat.cuh:86-88   (Shape4D, 16 bytes)
 *   cudamat/cudamat.cuh:90-107  (ConvDesc, 64 bytes, passed BY VALUE)
 * and that cudamat/cudamat.py:127-157 mirrors for ctypes.  A caller compiled
 * against the reference's own cudamat.cuh can pass its structs to this library
 * unchanged; nothing here depends on CUDA headers, so plain C hosts can include it.
 *
 * Conventions:
 *   - cudamat is column-major rows x cols = size[0] x size[1] = images x features;
 *     element (n, f) lives at data_device[n + size[0] * f].
 *   - Shape4D = {N, W(x), H(y), C*T}: x BEFORE y.
 *   - ConvDesc.padding_* is the NEGATED config padding (<= 0 is "pad by -p");
 *     *_channel_end == 0 means "up to num_*_channels".
 */
#ifndef CONVNET_B200_CUDAMAT_ABI_H_
#define CONVNET_B200_CUDAMAT_ABI_H_

#ifdef __cplusplus
extern "C" {
#endif

/* If the reference's cudamat.cuh was included first, use its definitions. */
#ifndef _CUDAMAT_CUH

struct cudamat {
  float* data_host;
  float* data_device;
  int on_device;
  int on_host;
  int size[2];              /* {rows = num_images, cols = features} */
  int is_trans;             /* must be 0 on this path */
  int owns_data;
  unsigned long long tex_obj; /* cudaTextureObject_t in the reference; unused here */
};

typedef struct Shape4D {
  int shape[4];
} Shape4D;

typedef struct ConvDesc {
  int num_input_channels;
  int num_output_channels;
  int kernel_size_y;
  int kernel_size_x;
  int kernel_size_t;
  int stride_y;
  int stride_x;
  int stride_t;
  int padding_y;
  int padding_x;
  int padding_t;
  int input_channel_begin;
  int input_channel_end;
  int output_channel_begin;
  int output_channel_end;
  int num_groups;
} ConvDesc;

#endif /* _CUDAMAT_CUH */

#ifdef __cplusplus
}  /* extern "C" */
typedef struct cudamat cudamat;
#else
typedef struct cudamat cudamat;
#endif

#endif  /* CONVNET_B200_CUDAMAT_ABI_H_ */
