// data.h — the device side of the reference's input pipeline: a chunk of the data set resident on the GPU,
// one image per column, and per minibatch a crop + mirror + transpose into the input layer (src/datahandler.cc:146-200
// DataHandler::GetBatch, :520-531 DataIterator::AddNoise, :533-568 DataIterator::SampleNoise).  Reading the data set from
// disk (HDF5 / image lists) stays with the caller: it hands over float pixels in the reference's (colour, row, column)
// order through Upload().
#pragma once
#include <cstdint>
#include <vector>

#include "matrix.h"

namespace cnbhost {

class DataIterator {
 public:
  // images of image_size_y x image_size_x x channels, `chunk_size` of them on the GPU; the net sees gpu_image_size_* crops
  DataIterator(int chunk_size, int channels, int image_size_y, int image_size_x, int gpu_image_size_y, int gpu_image_size_x,
               bool translate, bool flip, uint64_t seed);
  ~DataIterator();
  int ChunkSize() const { return chunk_size_; }
  int NumDims() const { return channels_ * image_size_y_ * image_size_x_; }
  // host pixels of images [first, first + count) of the chunk, image-major, each image (colour, row, column); async
  void Upload(const float* host, int first, int count);
  // :533-568 — the jitter of one minibatch: random offsets when `translate`, else the centre / corner crop number
  // multiplicity_id % 5; random mirror bits when `flip`, else multiplicity_id / 5
  void SampleNoise(int batch_size, int multiplicity_id);
  // the deterministic (translate == false) views of :547-556: centre, top-left, top-right, bottom-right, bottom-left
  static void ViewOffset(int multiplicity_id, int max_offset_x, int max_offset_y, int* w, int* h);
  // :520-531 + GetBatch's slice: images [start, start + batch) of the chunk -> dest (batch x C*gy*gx, image fastest)
  void AddNoise(int start, Matrix& dest);
  const std::vector<float>& LastWidthOffsets() const { return h_wo_; }
  const std::vector<float>& LastHeightOffsets() const { return h_ho_; }
  const std::vector<float>& LastFlipBits() const { return h_flip_; }

 private:
  int chunk_size_, channels_, image_size_y_, image_size_x_, gpu_image_size_y_, gpu_image_size_x_;
  bool translate_, flip_;
  uint64_t rng_;
  Matrix data_, width_offset_, height_offset_, flip_bit_;
  std::vector<float> h_wo_, h_ho_, h_flip_;
  float* pinned_ = nullptr;                       // 3 x batch floats, the staging area of the three vectors
  int pinned_cap_ = 0;
  uint64_t NextRand();
  float Uniform();                                // [0, 1)
};

}  // namespace cnbhost
