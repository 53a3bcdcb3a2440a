// edge.h — the Edge operator API of the reference (src/edge.h:20-190, src/edge_with_weight.h:10-58)
// and the six edge types on the BASELINE configs' path, on top of the host Matrix facade.
//
//   ConvEdge            src/conv_edge.{h,cc}            conv -> shared bias ; wgrad -> bias grad
//   MaxPoolEdge         src/maxpool_edge.{h,cc}
//   AvgPoolEdge         src/avgpool_edge.{h,cc}
//   ResponseNormEdge    src/response_norm_edge.{h,cc}
//   FCEdge              src/fc_edge.{h,cc}              (reference: Matrix::Dot / cublasSgemm)
//   ConvOneToOneEdge    src/conv_onetoone_edge.{h,cc}   (reference: Matrix::Dot / cublasSgemm)
// FC and 1x1 edges run on the same implicit-GEMM conv kernels (a 1x1 convolution IS that GEMM),
// SURVEY.md §8(f) rank 1.  The protobuf `config::Edge` is replaced by the plain EdgeConfig struct
// (protobuf is not in the image); field names follow proto/convnet_config.proto:120-221.
#pragma once
#include <string>
#include <vector>

#include "matrix.h"

namespace cnbhost {

class Layer;

enum EdgeType { FC, CONVOLUTIONAL, MAXPOOL, AVGPOOL, RESPONSE_NORM, CONV_ONETOONE };

struct OptimizerConfig {             // proto/convnet_config.proto Optimizer (SGD subset, src/optimizer.cc:174-200)
  float epsilon = 0.01f;
  float momentum = 0.9f;
  float l2_decay = 0.f;
};

struct EdgeConfig {
  std::string name, source, dest;
  EdgeType edge_type = FC;
  int kernel_size = 1, stride = 1, padding = 0;
  int kernel_size_y = 0, kernel_size_x = 0, stride_y = 0, stride_x = 0, padding_y = -1, padding_x = -1;   // 0/-1: unset
  int kernel_size_t = 1, stride_t = 1, padding_t = 0;
  bool shared_bias = true, has_no_bias = false;
  float add_scale = 0.0005f, pow_scale = 0.75f, frac_of_filters_response_norm = 0.25f;
  bool response_norm_in_blocks = false;
  float scale_gradients = 1.f;
  float init_wt = 0.f;               // 0: DENSE_UNIFORM_SQRT_FAN_IN (edge_with_weight.cc:120-128)
  OptimizerConfig weight_optimizer, bias_optimizer;
  bool grad_check = false;
  int grad_check_num_params = 10;
  std::vector<float> grad_check_epsilon;
};

class Edge {
 public:
  explicit Edge(const EdgeConfig& c);
  virtual ~Edge() {}
  static Edge* ChooseEdgeClass(const EdgeConfig& c);                     // src/edge.cc:17-60
  static ConvDesc GetConvDesc(const EdgeConfig& c);                      // src/edge.cc:87-106 (negates padding)
  static void GetNumModules(const ConvDesc d, int image_size_y, int image_size_x, int image_size_t,
                            int& num_modules_y, int& num_modules_x, int& num_modules_t);   // src/edge.cc:108-114

  virtual void ComputeUp(Matrix& input, Matrix& output, bool overwrite, bool train) = 0;
  virtual void ComputeDown(Matrix& deriv_output, Matrix& input, Matrix& output, Matrix& deriv_input,
                           bool overwrite) = 0;
  virtual void ComputeOuter(Matrix& input, Matrix& deriv_output) {}
  virtual void UpdateWeights() {}
  virtual void SetMemory(Matrix& p) {}
  virtual void SetGradMemory(Matrix& p) {}
  virtual void SetHistoryMemory(Matrix& p) {}
  virtual size_t GetParameterMemoryRequirement() { return 0; }
  virtual void Initialize(unsigned seed) {}
  virtual bool HasNoParameters() const { return true; }
  virtual void SetImageSize(int image_size_y, int image_size_x, int image_size_t);
  virtual double FlopsUp() const { return 0; }                           // 2*MACs per batch

  int GetNumModulesY() const { return num_modules_y_; }
  int GetNumModulesX() const { return num_modules_x_; }
  int GetNumModulesT() const { return num_modules_t_; }
  void SetInputChannels(int a) { num_input_channels_ = a; }
  void SetOutputChannels(int a) { num_output_channels_ = a; }
  int GetNumOutputChannels() const { return num_output_channels_; }
  const std::string& GetName() const { return name_; }
  const EdgeConfig& Config() const { return config_; }
  Layer* GetSource() { return source_; }
  Layer* GetDest() { return dest_; }
  void SetSource(Layer* l) { source_ = l; }
  void SetDest(Layer* l) { dest_ = l; }
  void SetBatchSize(int n) { batch_size_ = n; }
  // Epilogue fusion (convnet_b200_fuse_next): the ReLU of the destination layer rides in ComputeUp's conv epilogue
  // (together with the shared bias), the ReLU derivative of the source layer in ComputeDown's. ConvNet decides.
  virtual bool CanFuseReLU() const { return false; }
  virtual bool CanFuseMask() const { return false; }
  void SetFuseReLU(bool v) { fuse_relu_ = v; }
  void SetFuseMask(bool v) { fuse_mask_ = v; }
  bool WantsFuseReLU() const { return fuse_relu_; }
  bool WantsFuseMask() const { return fuse_mask_; }
  // bf16 mode: the kernel that writes a tensor LAST also leaves its bf16 copy for the conv edge that reads it next
  // (convnet_b200_emit_bf16_next) instead of that edge running a conversion pass.  ConvNet sets these before each call:
  // emit_up: ComputeUp is the last writer of the destination state and the next edge multiplies in bf16;
  // emit_down: ComputeDown is the last writer of the source layer's derivative and the edge below multiplies in bf16.
  void SetEmitUp(bool v) { emit_up_ = v; }
  void SetEmitDown(bool v) { emit_down_ = v; }
  // Fused bias gradient: the edge ABOVE writes this edge's output derivative last, and its kernel can sum the channels
  // while it stores them (convnet_b200_fuse_next_bias_grad).  ConvNet asks the lower edge for its target (which makes that
  // edge skip its own SumRows in ComputeOuter) and hands it to the upper edge's ComputeDown.
  struct BiasGradTarget { float* grad_bias = nullptr; float st = 0.f, so = 1.f; };
  virtual bool OfferFusedBiasGrad(BiasGradTarget*) { return false; }
  void SetBiasGradRequest(const BiasGradTarget& t) { bg_request_ = t; }
  // ComputeDown multiplies the derivative it writes by this factor (1 = none): the dropout derivative of a ReLU layer folded
  // into the dgrad epilogue (convnet_b200_fuse_next_scale); only edges whose ComputeDown is a conv dgrad with the mask fused
  void SetDerivScale(float s) { deriv_scale_ = s; }
  virtual bool CanScaleDeriv() const { return false; }
  // The dropout of the destination layer rides in ComputeUp's conv epilogue behind the fused bias + ReLU
  // (convnet_b200_fuse_next_dropout: no mask tensor — ConvNet asks only when the backward pass folds the dropout derivative
  // into the dgrad above, see ConvNet::DropoutFolds).  One-shot: the next ComputeUp consumes it.
  virtual bool CanFuseDropout() const { return false; }
  void SetDropoutRequest(float prob, float scale, unsigned long long seed) { drop_prob_ = prob; drop_scale_ = scale; drop_seed_ = seed; }
  virtual bool CanProduceBiasGrad() const { return false; }   // ComputeDown kernels that take the request
  virtual bool WantsBf16Input() const { return false; }      // this edge reads its input (fprop / wgrad) as bf16
  virtual bool WantsBf16Deriv() const { return false; }      // this edge reads its output derivative (wgrad / dgrad) as bf16

 protected:
  EdgeConfig config_;
  std::string name_;
  Layer *source_, *dest_;
  int num_input_channels_, num_output_channels_;
  int image_size_y_, image_size_x_, image_size_t_;
  int num_modules_y_, num_modules_x_, num_modules_t_;
  int batch_size_;
  bool fuse_relu_ = false, fuse_mask_ = false;
  bool emit_up_ = false, emit_down_ = false;
  BiasGradTarget bg_request_;
  float deriv_scale_ = 1.f;
  float drop_prob_ = 0.f, drop_scale_ = 0.f;
  unsigned long long drop_seed_ = 0;
  void ApplyDropoutRequest(bool fused_epilogue) {   // call right before the ComputeUp kernel (after convnet_b200_fuse_next)
    if (drop_scale_ != 0.f && fused_epilogue) convnet_b200_fuse_next_dropout(drop_prob_, drop_scale_, drop_seed_);
    drop_scale_ = 0.f;
  }
  void ApplyBiasGradRequest() {            // call right before the ComputeDown kernel
    if (bg_request_.grad_bias) convnet_b200_fuse_next_bias_grad(bg_request_.grad_bias, bg_request_.st, bg_request_.so);
    bg_request_ = BiasGradTarget();
    if (deriv_scale_ != 1.f) convnet_b200_fuse_next_scale(deriv_scale_);
    deriv_scale_ = 1.f;
  }
};

// The side stream of ConvNet::TrainOneBatch (all-reduce + optimizer, see convnet.h): bias-gradient column sums are
// memory-bound passes over a derivative that is already final, so they run there, beside the tensor-bound wgrad / dgrad
// kernels of the main stream.  Null stream: everything stays on the main stream.
struct SideLane { cudaStream_t stream = nullptr; cudaEvent_t ready = nullptr; bool used = false; };

class EdgeWithWeight : public Edge {
 public:
  void SetSideLane(SideLane* s) { side_ = s; }
  // after this edge's optimizer step, on the optimizer's stream: rebuild what the next ComputeDown derives from the
  // weights alone (convnet_b200_prestage_next), off the next step's critical path.  Default: nothing to prepare.
  virtual void PrestageDown() {}
  explicit EdgeWithWeight(const EdgeConfig& c) : Edge(c), has_no_bias_(c.has_no_bias), scale_gradients_(c.scale_gradients), num_grads_received_(0) {}
  bool HasNoParameters() const override { return false; }
  void UpdateWeights() override;                                         // src/edge_with_weight.cc:96-118
  void SetHistoryMemory(Matrix& p) override;
  void Initialize(unsigned seed) override;
  Matrix& GetWeight() { return weights_; }
  Matrix& GetGradWeight() { return grad_weights_; }
  Matrix& GetBias() { return bias_; }
  Matrix& GetGradBias() { return grad_bias_; }
  int GetNumGradsReceived() const { return num_grads_received_; }
  void IncrementNumGradsReceived() { num_grads_received_++; }
  void NotifyStart() { num_grads_received_ = 0; }
  virtual int FanIn() const = 0;
  // bf16 mode (convnet_b200_set_conv_precision(2)): each tensor this edge feeds to two conv calls of a step has ONE bf16
  // copy — the input (fprop + wgrad), the output derivative (wgrad + dgrad) and the weights (fprop + dgrad).  The copy is
  // normally written by the kernel that produced the tensor (emit_up / emit_down of the neighbouring edges, the dropout
  // and SGD kernels); convnet_b200_bf16_ensure converts only when no valid copy exists.  Which calls really run in bf16 is
  // learnt from convnet_b200_last_conv_path() during the first step (FC-shaped calls stay on tf32 and are not staged).
  bool WantsBf16Input() const override { return bf_up_ == 1 || bf_outer_ == 1; }
  bool WantsBf16Deriv() const override { return bf_outer_ == 1 || bf_down_ == 1; }
  void AppendSgdTensors(std::vector<CnbSgdTensor>& out);                 // weights (+ bias) of this edge for one multi-tensor update
  bool OfferFusedBiasGrad(BiasGradTarget* t) override;
  virtual bool BiasIsPerChannel2D() const { return !has_no_bias_; }       // one bias per output channel, 2-D layer
  // (a conv dgrad would only run the column-sum pass inside the library call, on the main stream; leaving it to the edge
  //  below puts it on the side lane instead — so only the pooling edges, whose kernels really fuse it, take the request)
  bool CanScaleDeriv() const override { return fuse_mask_; }               // (3-D ConvEdge: fuse_mask_ is off, CanFuseMask)

 protected:
  void StageForUp(Matrix& input);
  void StageForBprop(Matrix& deriv_output);
  void SumBiasRows(Matrix& deriv_output, float scale_targets, float scale);      // SumRows on the side lane when there is one
  SideLane* side_ = nullptr;
  void NoteUp();
  void NoteDown();
  void NoteOuter();
  Matrix weights_, grad_weights_, bias_, grad_bias_, hist_weights_, hist_bias_;
  bool has_no_bias_;
  float scale_gradients_;
  int num_grads_received_;
  int bf_up_ = -1, bf_down_ = -1, bf_outer_ = -1;        // -1 unknown, 0 tf32 / fp32 path, 1 bf16 path
  // the tensors of the last ComputeDown that took the bf16 path (layer-owned, stable): PrestageDown re-describes that call
  Matrix* down_out_ = nullptr;
  Matrix* down_in_ = nullptr;
  void RememberDown(Matrix& deriv_output, Matrix& deriv_input) {
    down_out_ = bf_down_ == 1 ? &deriv_output : nullptr;
    down_in_ = bf_down_ == 1 ? &deriv_input : nullptr;
  }
  bool bias_grad_fused_ = false;                         // this step's bias gradient comes from the edge above (ComputeOuter skips SumRows)
};

class ConvEdge : public EdgeWithWeight {
 public:
  explicit ConvEdge(const EdgeConfig& c);
  void SetImageSize(int y, int x, int t) override;
  size_t GetParameterMemoryRequirement() override;
  void SetMemory(Matrix& p) override;
  void SetGradMemory(Matrix& p) override;
  void ComputeUp(Matrix& input, Matrix& output, bool overwrite, bool train) override;
  void ComputeDown(Matrix& deriv_output, Matrix& input, Matrix& output, Matrix& deriv_input, bool overwrite) override;
  void PrestageDown() override;
  void ComputeOuter(Matrix& input, Matrix& deriv_output) override;
  double FlopsUp() const override;
  int FanIn() const override;
  ConvDesc GetConvDesc() const { return conv_desc_; }
  bool CanFuseReLU() const override { return !has_no_bias_ && shared_bias_ && image_size_t_ == 1; }
  bool CanFuseDropout() const override { return fuse_relu_ && CanFuseReLU(); }
  bool CanFuseMask() const override { return image_size_t_ == 1; }
  bool BiasIsPerChannel2D() const override { return !has_no_bias_ && shared_bias_ && image_size_t_ == 1; }

 private:
  ConvDesc conv_desc_;
  int partial_sum_y_, partial_sum_x_;
  bool shared_bias_;
};

class FCEdge : public EdgeWithWeight {          // weights [Cout x K] column-major, like a conv filter bank
 public:
  explicit FCEdge(const EdgeConfig& c) : EdgeWithWeight(c) {}
  void SetImageSize(int y, int x, int t) override;
  size_t GetParameterMemoryRequirement() override;
  void SetMemory(Matrix& p) override;
  void SetGradMemory(Matrix& p) override;
  void ComputeUp(Matrix& input, Matrix& output, bool overwrite, bool train) override;
  void ComputeDown(Matrix& deriv_output, Matrix& input, Matrix& output, Matrix& deriv_input, bool overwrite) override;
  void ComputeOuter(Matrix& input, Matrix& deriv_output) override;
  double FlopsUp() const override;
  int FanIn() const override { return num_inputs_; }
  bool CanFuseReLU() const override { return !has_no_bias_; }
  bool CanFuseDropout() const override { return fuse_relu_ && CanFuseReLU(); }
  bool CanFuseMask() const override { return true; }


 private:
  void View(Matrix& in, Matrix& out);
  int num_inputs_ = 0;
  ConvDesc desc_;
};

class ConvOneToOneEdge : public EdgeWithWeight {
 public:
  explicit ConvOneToOneEdge(const EdgeConfig& c) : EdgeWithWeight(c) {}
  void SetImageSize(int y, int x, int t) override;
  size_t GetParameterMemoryRequirement() override;
  void SetMemory(Matrix& p) override;
  void SetGradMemory(Matrix& p) override;
  void ComputeUp(Matrix& input, Matrix& output, bool overwrite, bool train) override;
  void ComputeDown(Matrix& deriv_output, Matrix& input, Matrix& output, Matrix& deriv_input, bool overwrite) override;
  void PrestageDown() override;
  void ComputeOuter(Matrix& input, Matrix& deriv_output) override;
  double FlopsUp() const override;
  int FanIn() const override { return num_input_channels_; }
  bool CanFuseReLU() const override { return !has_no_bias_; }
  bool CanFuseDropout() const override { return fuse_relu_ && CanFuseReLU(); }
  bool CanFuseMask() const override { return true; }

 private:
  ConvDesc desc_;
};

class MaxPoolEdge : public Edge {
 public:
  explicit MaxPoolEdge(const EdgeConfig& c) : Edge(c), conv_desc_(Edge::GetConvDesc(c)) {}
  void SetImageSize(int y, int x, int t) override;
  void ComputeUp(Matrix& input, Matrix& output, bool overwrite, bool train) override;
  void ComputeDown(Matrix& deriv_output, Matrix& input, Matrix& output, Matrix& deriv_input, bool overwrite) override;
  bool CanFuseMask() const override { return true; }
  bool CanProduceBiasGrad() const override { return image_size_t_ == 1; }

 protected:
  ConvDesc conv_desc_;
};

class AvgPoolEdge : public MaxPoolEdge {
 public:
  explicit AvgPoolEdge(const EdgeConfig& c) : MaxPoolEdge(c) {}
  void ComputeUp(Matrix& input, Matrix& output, bool overwrite, bool train) override;
  void ComputeDown(Matrix& deriv_output, Matrix& input, Matrix& output, Matrix& deriv_input, bool overwrite) override;
};

class ResponseNormEdge : public Edge {
 public:
  bool CanFuseReLU() const override { return image_size_t_ == 1; }      // max(., 0) rides in the rnorm kernel's store
  explicit ResponseNormEdge(const EdgeConfig& c)
      : Edge(c), num_filters_response_norm_(0), blocked_(c.response_norm_in_blocks), add_scale_(c.add_scale),
        pow_scale_(c.pow_scale), frac_of_filters_response_norm_(c.frac_of_filters_response_norm) {}
  void SetImageSize(int y, int x, int t) override;
  void ComputeUp(Matrix& input, Matrix& output, bool overwrite, bool train) override;
  void ComputeDown(Matrix& deriv_output, Matrix& input, Matrix& output, Matrix& deriv_input, bool overwrite) override;

 private:
  int num_filters_response_norm_;
  bool blocked_;
  float add_scale_, pow_scale_, frac_of_filters_response_norm_;
};

}  // namespace cnbhost
