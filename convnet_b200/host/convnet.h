// convnet.h — the slice of the reference's ConvNet that sequences the hot path
// (src/convnet.{h,cc}: BuildNet :150, AllocateEdgeMemory :272-298, Fprop :377, Bprop :390,
// UpdateWeights :440-450, TrainOneBatch :475-485), its GradChecker (src/grad_check.cc) and the
// data-parallel gradient sync that replaces the reference's host-staged MPI
// Accumulate + Broadcast (src/convnet.cc:407-438) with in-place NCCL all-reduce over NVLink.
//
// Layers form a chain (every BASELINE net is one: Appendix B of SURVEY.md).  The model comes
// from a ModelConfig struct (protobuf / pbtxt parsing is out of scope);
// builders for the BASELINE configs live in models.cc.
#pragma once
#include <memory>
#include <string>
#include <vector>

#include "edge.h"

namespace cnbhost {

enum Activation { LINEAR, RECTIFIED_LINEAR, SOFTMAX };

struct LayerConfig {
  std::string name;
  int num_channels = 0;
  bool is_input = false, is_output = false;
  Activation activation = LINEAR;
  int image_size_y = 0, image_size_x = 0, image_size_t = 1;   // input layer only
  float dropprob = 0.f;
};

struct ModelConfig {
  std::string name;
  std::vector<LayerConfig> layer;
  std::vector<EdgeConfig> edge;
  unsigned seed = 42;
};

class Layer {                                   // src/layer.{h,cc}, reduced to state/deriv + activation
 public:
  explicit Layer(const LayerConfig& c) : config_(c), image_size_y_(0), image_size_x_(0), image_size_t_(1) {}
  void SetSize(int y, int x, int t) { image_size_y_ = y; image_size_x_ = x; image_size_t_ = t; }
  void AllocateMemory(int batch_size);
  void ApplyActivation(bool emit_bf16 = false);               // layer.cc:545-560
  void ApplyDerivativeOfActivation(bool emit_bf16 = false);   // layer.cc:562-580
  void ApplyDropout(bool train, unsigned long long step, unsigned long long salt, bool emit_bf16 = false);   // layer.cc:  mask = rand > dropprob ; state *= mask
  void ApplyDerivativeofDropout(bool emit_bf16 = false);
  bool HasDropout() const { return config_.dropprob > 0 && !config_.is_input; }
  float DropoutScale() const { return 1.0f / (1.0f - config_.dropprob); }
  float DropoutProb() const { return config_.dropprob; }
  unsigned long long DropoutSeed(unsigned long long step, unsigned long long salt) const;
  void SetDropoutDerivFolded(bool v) { dropout_deriv_folded_ = v; }   // this step: the edge above scaled the derivative instead
  bool HasSeparateActivationPass() const { return config_.activation == RECTIFIED_LINEAR && !activation_fused_; }
  bool HasSeparateDerivPass() const { return config_.activation == RECTIFIED_LINEAR && !deriv_fused_; }
  void ComputeDeriv();                          // softmax + cross-entropy: deriv = p - onehot   (loss_functions.cc)
  Matrix& GetState() { return state_; }
  Matrix& GetDeriv() { return deriv_; }
  int* GetLabels() { return labels_; }
  float* GetLossPerImage() { return loss_per_image_.GetDevData(); }
  bool IsInput() const { return config_.is_input; }
  bool IsOutput() const { return config_.is_output; }
  int GetNumChannels() const { return config_.num_channels; }
  int GetSizeY() const { return image_size_y_; }
  int GetSizeX() const { return image_size_x_; }
  int GetSizeT() const { return image_size_t_; }
  const std::string& GetName() const { return config_.name; }
  Activation GetActivation() const { return config_.activation; }
  void SetActivationFused(bool v) { activation_fused_ = v; }      // the incoming edge applies the ReLU in its epilogue
  void SetDerivFused(bool v) { deriv_fused_ = v; }                // the outgoing edge applies ReLU' in its epilogue
  ~Layer();

 private:
  LayerConfig config_;
  int image_size_y_, image_size_x_, image_size_t_;
  Matrix state_, deriv_, loss_per_image_, dropout_mask_;
  int* labels_ = nullptr;
  bool activation_fused_ = false, deriv_fused_ = false, dropout_deriv_folded_ = false;
};

// NCCL all-reduce of the flat gradient buffer, bucketed along edge boundaries and launched on a side
// stream as soon as the gradients of a bucket are final, so the exchange hides under back-propagation.
class DataParallelSync {
 public:
  DataParallelSync();
  ~DataParallelSync();
  static bool GetUniqueId(char out[128]);                       // rank 0; bytes are broadcast by the launcher
  bool Init(int rank, int world, const char id[128]);
  int world() const { return world_; }
  int rank() const { return rank_; }
  void Broadcast(float* buf, size_t count);                     // initial parameters from rank 0 (convnet.cc:300-309)
  // average buf[offset, offset+count) over ranks on `comm` (the caller orders `comm` after the producers of the gradients)
  void AllReduceAverageAsync(float* buf, size_t offset, size_t count, cudaStream_t comm);
  int reserved_sms() const { return nccl_ctas_; }               // SMs the collective's CTAs need while it is in flight
 private:
  void* comm_ = nullptr;
  cudaStream_t comm_stream_ = nullptr;                          // broadcast only; the all-reduces ride on ConvNet's side stream
  cudaEvent_t ready_ = nullptr, done_ = nullptr;
  int rank_ = 0, world_ = 1, nccl_ctas_ = 0;
};

// Gradient buckets: the unit of the overlapped all-reduce AND of the optimizer step.  Back-propagation finalises edge
// gradients from the LAST edge to the first, and edge slices are adjacent in the flat buffer (128-float padded), so a
// bucket is the contiguous range [lo, hi) — the edges [trigger, last] — that becomes final when edge `trigger` has run
// ComputeOuter.  Buckets are closed once they hold >= bucket_floats; the FIRST weighted edge of the net always gets a
// bucket of its own (its gradient is the last to appear: only that small exchange stays exposed at the end of the step);
// every parameter belongs to exactly one bucket.
struct Bucket { size_t lo, hi; int trigger, last; };
std::vector<Bucket> PlanBuckets(const std::vector<size_t>& edge_offset, const std::vector<size_t>& edge_size,
                                size_t bucket_floats);

class ConvNet {
 public:
  ConvNet(const ModelConfig& model, int batch_size);
  virtual ~ConvNet();
  void AllocateMemory();                                        // convnet.cc:272-298: ONE flat parameter / gradient buffer
  virtual void Fprop(bool train);                               // convnet.cc:377-388
  virtual void Bprop();                                         // convnet.cc:390-405
  virtual void UpdateWeights();                                 // convnet.cc:440-450
  void ComputeDeriv();
  void TrainOneBatch(float* loss_out);                          // convnet.cc:475-485
  float GetLoss();                                              // sum of per-image CE (synchronises)
  void SetDataParallel(DataParallelSync* dp, size_t bucket_floats);
  void SetBucketFloats(size_t bucket_floats);                   // re-plan the buckets (also used without data parallelism)
  void BroadcastParameters();
  void InvalidateStaging();                                     // after any write to the parameters from outside UpdateWeights

  Layer& InputLayer() { return *layers_.front(); }
  Layer& OutputLayer() { return *layers_.back(); }
  std::vector<Edge*>& Edges() { return edges_; }
  std::vector<Layer*>& Layers() { return layers_; }
  Matrix& Parameters() { return parameters_; }
  Matrix& GradParameters() { return grad_parameters_; }
  size_t NumParameters() const { return num_params_; }
  int BatchSize() const { return batch_size_; }
  double FlopsFprop() const;
  double FlopsTrainStep() const;                                // fprop + wgrad for every weighted edge + dgrad except into the input
  const std::vector<size_t>& EdgeOffsets() const { return edge_offset_; }
  const std::vector<size_t>& EdgeSizes() const { return edge_size_; }
  float* DeviceLoss() { return loss_sum_.GetDevData(); }
  // One traced TrainOneBatch: device times (ms since the step began) of the pipeline's milestones, for the scaling report:
  // {fprop_end, bprop_compute_end, step_end, n_buckets, then per bucket {MB, exchange_begin, exchange_end, sgd_end}}.
  // exchange_* are -1 on one rank.  The events cost a few microseconds; the ordinary step records none of them.
  std::vector<float> TraceStep();

 protected:
  ModelConfig model_;
  int batch_size_;
  std::vector<Layer*> layers_;
  std::vector<Edge*> edges_;                    // edges_[i]: layers_[i] -> layers_[i+1]
  Matrix parameters_, grad_parameters_, history_, loss_sum_;
  std::vector<size_t> edge_offset_, edge_size_;
  size_t num_params_ = 0;
  // Side-stream pipeline of TrainOneBatch: as soon as a bucket's gradients are final its all-reduce (data parallel) is
  // enqueued on side_, and once the bucket's edges have finished their dgrad the multi-tensor SGD step of that bucket
  // follows on the same stream — the exchange and the update of the FC layers hide under the conv back-propagation.
  void IssueBucketUpdate(const Bucket& b);
  // layers_[i] is a ReLU layer with dropout whose derivative is written by a dgrad that can apply relu'(state) * 1/(1-p)
  // itself: the backward pass needs no mask tensor, and the forward pass may fuse the dropout into the edge below
  bool DropoutFolds(size_t i) const;
  void WaitSide();
  DataParallelSync* dp_ = nullptr;
  std::vector<Bucket> buckets_;
  // side_: bias-gradient passes; comm_: the NCCL all-reduces only; opt_: the per-bucket SGD steps.  Three streams so that a
  // long exchange (fc6: 302 MB) never delays the column sums, and an all-reduce (which waits for the side stream's bias
  // gradients) never queues behind the optimizer step of an earlier bucket
  cudaStream_t side_ = nullptr, comm_ = nullptr, opt_ = nullptr;
  cudaEvent_t ev_main_ = nullptr, ev_side_ = nullptr, ev_comm_ = nullptr, ev_opt_ = nullptr;
  std::vector<cudaEvent_t> ev_reduced_;         // per bucket: its all-reduce has finished (comm_ -> side_ / main)
  bool comm_pending_ = false;
  SideLane lane_;                               // what the edges see of the side stream (bias-gradient passes)
  bool eager_update_ = false, side_pending_ = false, opt_pending_ = false, updated_in_bprop_ = false;
  bool dropout_active_ = false;                 // the last Fprop applied dropout (train == true): states hold relu(x) * mask
  struct Trace {
    bool on = false;
    cudaEvent_t t0 = nullptr, fwd = nullptr, bwd = nullptr, end = nullptr;
    std::vector<cudaEvent_t> c0, c1, s1;        // per bucket: exchange begin / end (comm_), optimizer step end (side_)
  } trace_;
  unsigned long long step_ = 0;
  unsigned long long dropout_salt_ = 0xD1B54A32D192ED03ULL;      // model seed and data-parallel rank, see SetDataParallel
};

// src/grad_check.{h,cc}: finite-difference check of dLoss/dparam for the first k weights and biases
// of every edge flagged grad_check, through the whole net.
struct GradCheckResult {
  std::string edge;
  float epsilon;
  float mean_scaled_diff_w, mean_scaled_diff_b;     // pass if < 0.01 (grad_check.cc:61)
};
class GradChecker : public ConvNet {
 public:
  GradChecker(const ModelConfig& model, int batch_size) : ConvNet(model, batch_size) {}
  std::vector<GradCheckResult> Run(unsigned seed);
 private:
  float LossAt(Matrix& w, size_t index, float value);
  double LossAtD(Matrix& w, size_t index, float value);
};

// models.cc
ModelConfig BuildAlexNet();     // examples/imagenet/CLS_net_20140801232522.pbtxt
ModelConfig BuildLeNet();       // examples/mnist-conv/net.pbtxt
ModelConfig BuildC3D();
ModelConfig BuildTinyNet();     // small conv+pool+rnorm+1x1+fc net for tests / grad check
ModelConfig BuildModel(const std::string& name);

}  // namespace cnbhost
