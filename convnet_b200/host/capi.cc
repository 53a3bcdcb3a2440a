// capi.cc — plain C doors onto the host C++ (ConvNet / GradChecker / DataParallelSync) for ctypes.
#include <cstring>

#include "convnet.h"
#include "data.h"

using namespace cnbhost;

#define API extern "C" __attribute__((visibility("default")))

struct NetHandle {
  ConvNet* net = nullptr;
  GradChecker* checker = nullptr;
  DataParallelSync* dp = nullptr;
};

API void* cnb_net_create(const char* model, int batch_size, unsigned seed, int grad_checker) {
  ModelConfig m = BuildModel(model);
  m.seed = seed;
  NetHandle* h = new NetHandle;
  if (grad_checker) { h->checker = new GradChecker(m, batch_size); h->net = h->checker; }
  else h->net = new ConvNet(m, batch_size);
  h->net->AllocateMemory();
  return h;
}
API void cnb_net_destroy(void* p) {
  NetHandle* h = (NetHandle*)p;
  delete h->net; delete h->dp; delete h;
}
API long long cnb_net_num_params(void* p) { return (long long)((NetHandle*)p)->net->NumParameters(); }
API int cnb_net_num_edges(void* p) { return (int)((NetHandle*)p)->net->Edges().size(); }
API const char* cnb_net_edge_name(void* p, int i) { return ((NetHandle*)p)->net->Edges()[i]->GetName().c_str(); }
API double cnb_net_edge_flops(void* p, int i) { return ((NetHandle*)p)->net->Edges()[i]->FlopsUp(); }
API long long cnb_net_edge_offset(void* p, int i) { return (long long)((NetHandle*)p)->net->EdgeOffsets()[i]; }
API long long cnb_net_edge_size(void* p, int i) { return (long long)((NetHandle*)p)->net->EdgeSizes()[i]; }
API double cnb_net_flops_fprop(void* p) { return ((NetHandle*)p)->net->FlopsFprop(); }
API double cnb_net_flops_train(void* p) { return ((NetHandle*)p)->net->FlopsTrainStep(); }
API float* cnb_net_input(void* p) { return ((NetHandle*)p)->net->InputLayer().GetState().GetDevData(); }
API long long cnb_net_input_floats(void* p) { return (long long)((NetHandle*)p)->net->InputLayer().GetState().GetNumEls(); }
API int* cnb_net_labels(void* p) { return ((NetHandle*)p)->net->OutputLayer().GetLabels(); }
API float* cnb_net_output(void* p) { return ((NetHandle*)p)->net->OutputLayer().GetState().GetDevData(); }
API int cnb_net_num_classes(void* p) { return ((NetHandle*)p)->net->OutputLayer().GetState().GetCols(); }
// the caller may write through this pointer: staged bf16 copies of the weights are dropped
API float* cnb_net_params(void* p) { ((NetHandle*)p)->net->InvalidateStaging(); return ((NetHandle*)p)->net->Parameters().GetDevData(); }
API float* cnb_net_grads(void* p) { return ((NetHandle*)p)->net->GradParameters().GetDevData(); }
API float* cnb_net_layer_state(void* p, int i) { return ((NetHandle*)p)->net->Layers()[i]->GetState().GetDevData(); }
API long long cnb_net_layer_floats(void* p, int i) { return (long long)((NetHandle*)p)->net->Layers()[i]->GetState().GetNumEls(); }
API int cnb_net_num_layers(void* p) { return (int)((NetHandle*)p)->net->Layers().size(); }
API float* cnb_net_device_loss(void* p) { return ((NetHandle*)p)->net->DeviceLoss(); }

API void cnb_net_fprop(void* p, int train) { ((NetHandle*)p)->net->Fprop(train != 0); }
API void cnb_net_bprop(void* p) { ((NetHandle*)p)->net->ComputeDeriv(); ((NetHandle*)p)->net->Bprop(); }
API void cnb_net_update(void* p) { ((NetHandle*)p)->net->UpdateWeights(); }
API float cnb_net_loss(void* p) { return ((NetHandle*)p)->net->GetLoss(); }
// one training step; *loss (may be NULL) receives the summed cross-entropy of the batch (one scalar D2H, like GetLoss)
API void cnb_net_train_step(void* p, float* loss) { ((NetHandle*)p)->net->TrainOneBatch(loss); }
// one traced training step (ConvNet::TraceStep); returns the number of floats the full record has, writes min(cap, that)
API int cnb_net_trace_step(void* p, float* out, int cap) {
  const std::vector<float> t = ((NetHandle*)p)->net->TraceStep();
  for (int i = 0; i < cap && i < (int)t.size(); i++) out[i] = t[i];
  return (int)t.size();
}

// data parallel: rank 0 calls cnb_dp_unique_id, the launcher broadcasts the 128 bytes, every rank calls cnb_net_dp_init
API int cnb_dp_unique_id(char* out128) { return DataParallelSync::GetUniqueId(out128) ? 0 : -1; }
API int cnb_net_dp_init(void* p, int rank, int world, const char* id128, long long bucket_floats) {
  NetHandle* h = (NetHandle*)p;
  h->dp = new DataParallelSync();
  if (!h->dp->Init(rank, world, id128)) return -1;
  h->net->SetDataParallel(h->dp, (size_t)bucket_floats);
  h->net->BroadcastParameters();
  return 0;
}

// grad check: fills up to `cap` results; returns the number of checked edges
API int cnb_net_grad_check(void* p, unsigned seed, int cap, char* names /*cap x 64*/, float* eps, float* diff_w, float* diff_b) {
  NetHandle* h = (NetHandle*)p;
  if (!h->checker) return -1;
  std::vector<GradCheckResult> r = h->checker->Run(seed);
  int n = 0;
  for (const GradCheckResult& g : r) {
    if (n >= cap) break;
    strncpy(names + 64 * n, g.edge.c_str(), 63); names[64 * n + 63] = 0;
    eps[n] = g.epsilon; diff_w[n] = g.mean_scaled_diff_w; diff_b[n] = g.mean_scaled_diff_b;
    n++;
  }
  return n;
}

// the gradient-bucket plan (pure host logic; testable without a GPU). Returns the number of buckets (<= cap).
API int cnb_plan_buckets(int n_edges, const long long* offsets, const long long* sizes, long long bucket_floats, int cap,
                         long long* lo, long long* hi, int* trigger) {
  std::vector<size_t> off(offsets, offsets + n_edges), sz(sizes, sizes + n_edges);
  std::vector<Bucket> b = PlanBuckets(off, sz, (size_t)bucket_floats);
  int n = 0;
  for (const Bucket& k : b) { if (n >= cap) break; lo[n] = (long long)k.lo; hi[n] = (long long)k.hi; trigger[n] = k.trigger; n++; }
  return n;
}
// static description of a model (no device memory): per-edge parameter count, for planning / reporting
API int cnb_model_edge_params(const char* model, int batch, int cap, long long* sizes) {
  ModelConfig m = BuildModel(model);
  ConvNet net(m, batch);
  int n = 0;
  for (Edge* e : net.Edges()) { if (n >= cap) break; sizes[n++] = (long long)e->GetParameterMemoryRequirement(); }
  return n;
}

// ---- the device side of the input pipeline (data.h): a GPU-resident chunk + per-minibatch crop / mirror into the net's input
API void* cnb_data_create(int chunk_size, int channels, int image_size_y, int image_size_x, int gpu_image_size_y,
                          int gpu_image_size_x, int translate, int flip, unsigned long long seed) {
  return new DataIterator(chunk_size, channels, image_size_y, image_size_x, gpu_image_size_y, gpu_image_size_x,
                          translate != 0, flip != 0, seed);
}
API void cnb_data_destroy(void* d) { delete (DataIterator*)d; }
API void cnb_data_upload(void* d, const float* host, int first, int count) { ((DataIterator*)d)->Upload(host, first, count); }
// DataHandler::GetBatch for the input layer of `net`: sample the jitter, then cut images [start, start + batch) into it
API void cnb_data_get_batch(void* d, void* net, int start, int multiplicity_id) {
  DataIterator* it = (DataIterator*)d;
  Matrix& dest = ((NetHandle*)net)->net->InputLayer().GetState();
  it->SampleNoise(dest.GetRows(), multiplicity_id);
  it->AddNoise(start, dest);
}
// the jitter of the last minibatch (host copies): out = {width offsets, height offsets, mirror bits}, 3 x batch floats
API int cnb_data_last_noise(void* d, float* out, int cap) {
  DataIterator* it = (DataIterator*)d;
  const int n = (int)it->LastWidthOffsets().size();
  if (cap < 3 * n) return -1;
  memcpy(out, it->LastWidthOffsets().data(), sizeof(float) * n);
  memcpy(out + n, it->LastHeightOffsets().data(), sizeof(float) * n);
  memcpy(out + 2 * n, it->LastFlipBits().data(), sizeof(float) * n);
  return n;
}
// pure host logic (no GPU): offset of deterministic view `multiplicity_id` for a free range of (max_x, max_y) pixels
API void cnb_data_view_offset(int multiplicity_id, int max_offset_x, int max_offset_y, int* w, int* h) {
  DataIterator::ViewOffset(multiplicity_id, max_offset_x, max_offset_y, w, h);
}
