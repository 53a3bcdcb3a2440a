// models.cc — ModelConfig builders for the BASELINE configs (the reference reads these from pbtxt).
#include <cstdio>
#include <cstdlib>

#include "convnet.h"

namespace cnbhost {

namespace {
LayerConfig L(const char* name, int ch, Activation act = LINEAR, float dropprob = 0.f) {
  LayerConfig l; l.name = name; l.num_channels = ch; l.activation = act; l.dropprob = dropprob; return l;
}
EdgeConfig E(EdgeType t, int k = 1, int s = 1, int p = 0) {
  EdgeConfig e; e.edge_type = t; e.kernel_size = k; e.stride = s; e.padding = p;
  e.weight_optimizer.epsilon = 0.01f; e.weight_optimizer.momentum = 0.9f;
  e.bias_optimizer.epsilon = 0.01f; e.bias_optimizer.momentum = 0.9f;
  return e;
}
EdgeConfig Conv(int k, int s, int p, float l2 = 0.f) { EdgeConfig e = E(CONVOLUTIONAL, k, s, p); e.weight_optimizer.l2_decay = l2; return e; }
EdgeConfig Pool(int k, int s, int p) { return E(MAXPOOL, k, s, p); }
EdgeConfig RNorm(float add = 0.0005f, float pow = 0.75f, float frac = 0.25f) {
  EdgeConfig e = E(RESPONSE_NORM); e.add_scale = add; e.pow_scale = pow; e.frac_of_filters_response_norm = frac; return e;
}
void finish(ModelConfig& m) {
  for (size_t i = 0; i < m.edge.size(); i++) {
    m.edge[i].source = m.layer[i].name; m.edge[i].dest = m.layer[i + 1].name;
    m.edge[i].name = m.layer[i].name + ":" + m.layer[i + 1].name;
  }
}
}  // namespace

// examples/imagenet/CLS_net_20140801232522.pbtxt (net A)
ModelConfig BuildAlexNet() {
  ModelConfig m; m.name = "CLS_net_20140801232522";
  LayerConfig in = L("input", 3); in.is_input = true; in.image_size_y = in.image_size_x = 224; in.image_size_t = 1;
  m.layer = {in,
             L("hidden1_conv", 96, RECTIFIED_LINEAR), L("hidden1_maxpool", 96), L("hidden1_rnorm", 96, RECTIFIED_LINEAR),
             L("hidden2_conv", 256, RECTIFIED_LINEAR), L("hidden2_conv_nin1", 256, RECTIFIED_LINEAR),
             L("hidden2_maxpool", 256), L("hidden2_rnorm", 256, RECTIFIED_LINEAR),
             L("hidden3_conv", 384, RECTIFIED_LINEAR), L("hidden3_conv_nin1", 768, RECTIFIED_LINEAR),
             L("hidden4_conv", 384, RECTIFIED_LINEAR), L("hidden4_conv_nin1", 768, RECTIFIED_LINEAR, 0.1f),
             L("hidden4_conv_nin2", 384, RECTIFIED_LINEAR),
             L("hidden5_conv", 512, RECTIFIED_LINEAR), L("hidden5_conv_nin1", 1024, RECTIFIED_LINEAR, 0.3f),
             L("hidden5_conv_nin2", 512, RECTIFIED_LINEAR), L("hidden5_maxpool", 512),
             L("hidden6", 4096, RECTIFIED_LINEAR, 0.5f), L("hidden7", 4096, RECTIFIED_LINEAR, 0.5f),
             L("output", 1000, SOFTMAX)};
  m.layer.back().is_output = true;
  m.edge = {Conv(7, 2, 1), Pool(3, 2, 1), RNorm(),
            Conv(5, 2, 1), E(CONV_ONETOONE), Pool(3, 2, 1), RNorm(),
            Conv(3, 1, 1, 0.0005f), E(CONV_ONETOONE),
            Conv(3, 1, 1, 0.0005f), E(CONV_ONETOONE), E(CONV_ONETOONE),
            Conv(3, 1, 0, 0.0005f), E(CONV_ONETOONE), E(CONV_ONETOONE), Pool(3, 2, 1),
            E(FC), E(FC), E(FC)};
  finish(m);
  return m;
}

// examples/mnist-conv/net.pbtxt (net M): 28x28x1 -conv4x4-> 25x25x48 -pool4/2-> 11x11x48 -conv4x4-> 8x8x128 -pool4/2-> 3x3x128 -fc-> 10
ModelConfig BuildLeNet() {
  ModelConfig m; m.name = "mnist-conv";
  LayerConfig in = L("input", 1); in.is_input = true; in.image_size_y = in.image_size_x = 28;
  m.layer = {in, L("hidden1_conv", 48, RECTIFIED_LINEAR), L("hidden1_maxpool", 48),
             L("hidden2_conv", 128, RECTIFIED_LINEAR), L("hidden2_maxpool", 128), L("output", 10, SOFTMAX)};
  m.layer.back().is_output = true;
  m.edge = {Conv(4, 1, 0, 0.0005f), Pool(4, 2, 0), Conv(4, 1, 0, 0.0005f), Pool(4, 2, 0), E(FC)};
  for (EdgeConfig& e : m.edge) { e.weight_optimizer.momentum = 0.95f; e.bias_optimizer.momentum = 0.95f; }
  finish(m);
  return m;
}

// C3D-style video net: 16 x 112 x 112 x 3 clips, 3x3x3 kernels, pad y/x 1, pad t 0
ModelConfig BuildC3D() {
  ModelConfig m; m.name = "c3d";
  LayerConfig in = L("input", 3); in.is_input = true; in.image_size_y = in.image_size_x = 112; in.image_size_t = 16;
  m.layer = {in, L("conv1a", 64, RECTIFIED_LINEAR), L("pool1", 64), L("conv2a", 128, RECTIFIED_LINEAR), L("pool2", 128),
             L("conv3a", 256, RECTIFIED_LINEAR), L("pool3", 256), L("output", 101, SOFTMAX)};
  m.layer.back().is_output = true;
  auto c3 = []() { EdgeConfig e = Conv(3, 1, 1); e.kernel_size_t = 3; e.stride_t = 1; e.padding_t = 0; return e; };
  auto p3 = [](int kt) { EdgeConfig e = Pool(2, 2, 0); e.kernel_size_t = kt; e.stride_t = kt; e.padding_t = 0; return e; };
  EdgeConfig gp = Pool(0, 1, 0); gp.kernel_size_t = 0; gp.stride_t = 1;     // global pooling before the classifier
  m.edge = {c3(), p3(1), c3(), p3(2), c3(), gp, E(FC)};
  finish(m);
  return m;
}

// small net touching every edge type, used by tests and the grad check
ModelConfig BuildTinyNet() {
  ModelConfig m; m.name = "tiny";
  LayerConfig in = L("input", 8); in.is_input = true; in.image_size_y = in.image_size_x = 12;
  m.layer = {in, L("conv1", 16, RECTIFIED_LINEAR), L("pool1", 16), L("rnorm1", 16, RECTIFIED_LINEAR),
             L("nin1", 24, RECTIFIED_LINEAR), L("conv2", 16, RECTIFIED_LINEAR), L("avgpool", 16), L("output", 10, SOFTMAX)};
  m.layer.back().is_output = true;
  EdgeConfig ap = E(AVGPOOL, 2, 2, 0);
  m.edge = {Conv(3, 1, 1), Pool(3, 2, 1), RNorm(0.01f, 0.75f, 0.5f), E(CONV_ONETOONE), Conv(3, 2, 1), ap, E(FC)};
  for (EdgeConfig& e : m.edge) { e.grad_check = true; e.grad_check_num_params = 8; e.grad_check_epsilon = {1e-2f, 3e-3f, 1e-3f}; }
  finish(m);
  return m;
}

// net for run_grad_check: one edge of every weighted type around pooling and response-norm, with SMOOTH
// activations (linear units, average pooling).  Finite differences are only meaningful away from kinks: with
// ReLU / max-pool a unit that sits within epsilon of its kink makes the central difference the AVERAGE of two
// one-sided slopes at every epsilon (observed: float64 finite differences show the same), so the reference's 1 %
// criterion (grad_check.cc:61) is a data lottery there.  The ReLU / max-pool backward ops are verified against
// float64 autograd instead (tests/test_gpu_net.py::test_backprop_matches_float64_autograd).
ModelConfig BuildGradCheckNet() {
  ModelConfig m; m.name = "gradcheck";
  LayerConfig in = L("input", 4); in.is_input = true; in.image_size_y = in.image_size_x = 8;
  m.layer = {in, L("conv1", 8), L("pool1", 8), L("rnorm1", 8), L("nin1", 12), L("output", 5, SOFTMAX)};
  m.layer.back().is_output = true;
  m.edge = {Conv(3, 1, 1), E(AVGPOOL, 3, 2, 1), RNorm(0.01f, 0.75f, 0.5f), E(CONV_ONETOONE), E(FC)};
  for (EdgeConfig& e : m.edge) { e.grad_check = true; e.grad_check_num_params = 10; e.grad_check_epsilon = {1e-2f, 3e-3f, 1e-3f}; }
  finish(m);
  return m;
}

ModelConfig BuildModel(const std::string& name) {
  // "<model>+gradcheck": the model with run_grad_check's edge flags as BASELINE config 1 states them
  // (grad_check_num_params: 10, grad_check_epsilon: [1e-2, 1e-3, 1e-4]; src/grad_check.cc:20-61)
  const std::string suffix = "+gradcheck";
  if (name.size() > suffix.size() && name.compare(name.size() - suffix.size(), suffix.size(), suffix) == 0) {
    ModelConfig m = BuildModel(name.substr(0, name.size() - suffix.size()));
    for (EdgeConfig& e : m.edge) { e.grad_check = true; e.grad_check_num_params = 10; e.grad_check_epsilon = {1e-2f, 1e-3f, 1e-4f}; }
    return m;
  }
  if (name == "gradcheck") return BuildGradCheckNet();
  if (name == "alexnet") return BuildAlexNet();
  if (name == "lenet") return BuildLeNet();
  if (name == "c3d") return BuildC3D();
  if (name == "tiny") return BuildTinyNet();
  fprintf(stderr, "unknown model '%s'\n", name.c_str());
  exit(1);
}

}  // namespace cnbhost
