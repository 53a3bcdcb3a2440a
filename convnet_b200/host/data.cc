// data.cc — see data.h.
#include "data.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace cnbhost {

#define DATA_CUDA_CHECK(expr)                                                                          \
  do {                                                                                                 \
    cudaError_t _e = (expr);                                                                           \
    if (_e != cudaSuccess) { fprintf(stderr, "%s(%d): %s: %s\n", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); exit(1); } \
  } while (0)

DataIterator::DataIterator(int chunk_size, int channels, int image_size_y, int image_size_x, int gpu_image_size_y,
                           int gpu_image_size_x, bool translate, bool flip, uint64_t seed)
    : chunk_size_(chunk_size), channels_(channels), image_size_y_(image_size_y), image_size_x_(image_size_x),
      gpu_image_size_y_(gpu_image_size_y), gpu_image_size_x_(gpu_image_size_x), translate_(translate), flip_(flip),
      rng_(seed * 0x9E3779B97F4A7C15ULL + 0xD1B54A32D192ED03ULL) {
  if (gpu_image_size_y > image_size_y || gpu_image_size_x > image_size_x || chunk_size <= 0 || channels <= 0) {
    fprintf(stderr, "DataIterator: the crop must fit the image\n"); exit(1);
  }
  data_.AllocateGPUMemory(NumDims(), chunk_size);          // one image per column (src/datahandler.cc:60-75)
}
DataIterator::~DataIterator() {
  if (pinned_) {
    cudaStreamSynchronize(Matrix::Stream());                 // the last minibatch's offset copies may still read the block
    cudaFreeHost(pinned_);
  }
}

uint64_t DataIterator::NextRand() {                         // splitmix64
  uint64_t z = (rng_ += 0x9E3779B97F4A7C15ULL);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL; z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
float DataIterator::Uniform() { return (float)(NextRand() >> 40) * (1.0f / 16777216.0f); }

void DataIterator::Upload(const float* host, int first, int count) {
  if (first < 0 || count < 0 || first + count > chunk_size_) { fprintf(stderr, "DataIterator::Upload: out of range\n"); exit(1); }
  DATA_CUDA_CHECK(cudaMemcpyAsync(data_.GetDevData() + (size_t)first * NumDims(), host, sizeof(float) * (size_t)count * NumDims(),
                                  cudaMemcpyHostToDevice, Matrix::Stream()));
}

void DataIterator::ViewOffset(int multiplicity_id, int max_offset_x, int max_offset_y, int* w, int* h) {
  // position of view k = multiplicity_id % 5 along (x, y), in halves of the free range: 1 = centred, 0 / 2 = the two ends
  static const int kView[5][2] = {{1, 1}, {0, 0}, {2, 0}, {2, 2}, {0, 2}};
  auto place = [](int half_steps, int max_offset) { return half_steps == 1 ? max_offset / 2 : (half_steps == 2 ? max_offset : 0); };
  const int view = multiplicity_id % 5;
  *w = place(kView[view][0], max_offset_x);
  *h = place(kView[view][1], max_offset_y);
}

void DataIterator::SampleNoise(int batch_size, int multiplicity_id) {
  const int max_offset_y = image_size_y_ - gpu_image_size_y_, max_offset_x = image_size_x_ - gpu_image_size_x_;
  if (width_offset_.GetCols() != batch_size || width_offset_.GetDevData() == nullptr) {
    width_offset_.AllocateGPUMemory(1, batch_size);
    height_offset_.AllocateGPUMemory(1, batch_size);
    flip_bit_.AllocateGPUMemory(1, batch_size);
  }
  h_wo_.assign(batch_size, 0.f); h_ho_.assign(batch_size, 0.f); h_flip_.assign(batch_size, 0.f);
  if (translate_) {                                          // random jitter: uniform * (max + 1), rounded down
    for (int i = 0; i < batch_size; i++) {
      const int oy = (int)(Uniform() * (max_offset_y + 1)), ox = (int)(Uniform() * (max_offset_x + 1));
      h_ho_[i] = (float)(oy > max_offset_y ? max_offset_y : oy);      // (the product can round up to max + 1 in fp32)
      h_wo_[i] = (float)(ox > max_offset_x ? max_offset_x : ox);
    }
  } else {                                                   // deterministic views: the centre, then the four corners
    int wi, hi;
    ViewOffset(multiplicity_id, max_offset_x, max_offset_y, &wi, &hi);
    const float w = (float)wi, h = (float)hi;
    for (int i = 0; i < batch_size; i++) { h_wo_[i] = w; h_ho_[i] = h; }
  }
  for (int i = 0; i < batch_size; i++) h_flip_[i] = flip_ ? Uniform() : (float)(multiplicity_id / 5);   // mirrored if > 0.5
  // one pinned staging block: the three vectors travel behind whatever the stream is doing
  if (pinned_cap_ < 3 * batch_size) {
    if (pinned_) { DATA_CUDA_CHECK(cudaStreamSynchronize(Matrix::Stream())); cudaFreeHost(pinned_); }
    DATA_CUDA_CHECK(cudaMallocHost((void**)&pinned_, sizeof(float) * 3 * (size_t)batch_size));
    pinned_cap_ = 3 * batch_size;
  } else {
    DATA_CUDA_CHECK(cudaStreamSynchronize(Matrix::Stream()));       // the previous minibatch's copies have left the block
  }
  memcpy(pinned_, h_wo_.data(), sizeof(float) * batch_size);
  memcpy(pinned_ + batch_size, h_ho_.data(), sizeof(float) * batch_size);
  memcpy(pinned_ + 2 * batch_size, h_flip_.data(), sizeof(float) * batch_size);
  width_offset_.CopyFromHost(pinned_, batch_size);
  height_offset_.CopyFromHost(pinned_ + batch_size, batch_size);
  flip_bit_.CopyFromHost(pinned_ + 2 * batch_size, batch_size);
}

void DataIterator::AddNoise(int start, Matrix& dest) {
  const int batch_size = dest.GetRows();
  if (start < 0 || start + batch_size > chunk_size_ || width_offset_.GetCols() != batch_size) {
    fprintf(stderr, "DataIterator::AddNoise: slice out of range, or SampleNoise was not called for this batch size\n"); exit(1);
  }
  Matrix data_slice;
  data_.GetSlice(data_slice, start, start + batch_size);
  // (the reference copies with CopyTranspose when there is neither crop nor mirror; the same kernel covers that case)
  Matrix::ExtractPatches(data_slice, dest, width_offset_, height_offset_, flip_bit_, image_size_y_, image_size_x_,
                         gpu_image_size_y_, gpu_image_size_x_);
}

}  // namespace cnbhost
