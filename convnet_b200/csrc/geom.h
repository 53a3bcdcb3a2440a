// geom.h — validated geometry of one conv / pool call, built from the ABI arguments.
// The checks restate the reference's assert() lists
// (cudamat_conv_gemm.cu:576-613 fprop, :715-750 dgrad, :858-893 wgrad, :1151-1166 pool).
#pragma once
#include "common.cuh"

namespace cnb {

struct ConvGeom {
  int N, W, H;              // images per GPU, image width (x), height (y)
  int CinT, CoutT;          // channel counts of the image / output tensors (per frame window)
  int cin0, Cin;            // input-channel sub-range [cin0, cin0+Cin)
  int cout0, Cout;          // output-channel sub-range
  int modX, modY, modules;
  int ky, kx, sy, sx, py, px;   // py/px: (<=0) window start offsets, as in ConvDesc
  int K;                    // ky*kx*Cin
  int conv;                 // 1 tied filters, 0 untied (one bank per module)
  // 3-D: `frames` output frames; frame f reads the image window starting at
  // f*in_frame_step floats and writes at f*out_frame_step floats.
  int frames;
  long long in_frame_step, out_frame_step;
  long long img_total, out_total;   // floats in the whole image / output matrices
};

inline ConvGeom conv_geom(const Shape4D& img, const Shape4D& flt, const Shape4D& out,
                          const cudamat* img_m, const cudamat* flt_m, const cudamat* out_m,
                          ConvDesc d, bool conv, const char* what) {
  ConvGeom g;
  const int kt = d.kernel_size_t > 0 ? d.kernel_size_t : 1;
  int ic_end = d.input_channel_end == 0 ? d.num_input_channels : d.input_channel_end;
  int oc_end = d.output_channel_end == 0 ? d.num_output_channels : d.output_channel_end;
  g.N = img.shape[0]; g.W = img.shape[1]; g.H = img.shape[2];
  g.modX = out.shape[1]; g.modY = out.shape[2];
  g.modules = g.modX * g.modY;
  g.conv = conv ? 1 : 0;
  const int fmm = conv ? 1 : g.modules;        // filterModuleMult
  CNB_REQUIRE(d.num_groups == 1, what);
  CNB_REQUIRE(img.shape[0] == out.shape[0], what);
  CNB_REQUIRE(d.num_input_channels > 0 && d.num_output_channels > 0, what);
  CNB_REQUIRE(img.shape[3] % d.num_input_channels == 0, what);
  CNB_REQUIRE(out.shape[3] % d.num_output_channels == 0, what);
  const int T = img.shape[3] / d.num_input_channels;      // image frames
  g.frames = out.shape[3] / d.num_output_channels;        // output frames (modT)
  if (T > 1 || kt > 1 || g.frames > 1) {
    // 3-D (cudamat_conv3d_gemm.cu:23,37): kt frames are folded into channels.
    CNB_REQUIRE(d.padding_t == 0, what);
    CNB_REQUIRE(kt == 1 || (ic_end == d.num_input_channels && d.input_channel_begin == 0), what);
    CNB_REQUIRE(d.stride_t >= 1, what);
    CNB_REQUIRE((T - kt) / d.stride_t + 1 == g.frames, what);
    CNB_REQUIRE(conv, what);
  } else {
    CNB_REQUIRE(g.frames == 1, what);
  }
  g.CinT = d.num_input_channels * kt;
  g.CoutT = d.num_output_channels;
  ic_end *= kt;
  g.cin0 = d.input_channel_begin; g.Cin = ic_end - d.input_channel_begin;
  g.cout0 = d.output_channel_begin; g.Cout = oc_end - d.output_channel_begin;
  CNB_REQUIRE(g.cin0 >= 0 && g.cout0 >= 0 && g.Cin > 0 && g.Cout > 0, what);
  CNB_REQUIRE(ic_end <= g.CinT && oc_end <= g.CoutT, what);
  g.ky = d.kernel_size_y; g.kx = d.kernel_size_x;
  g.sy = d.stride_y; g.sx = d.stride_x; g.py = d.padding_y; g.px = d.padding_x;
  CNB_REQUIRE(g.ky > 0 && g.kx > 0 && g.sy > 0 && g.sx > 0, what);
  g.K = g.ky * g.kx * g.Cin;
  CNB_REQUIRE(flt.shape[0] == g.Cout, what);
  CNB_REQUIRE(flt.shape[1] == g.kx && flt.shape[2] == g.ky, what);
  CNB_REQUIRE(flt.shape[3] == g.Cin * fmm, what);
  g.in_frame_step = (long long)g.W * g.H * d.num_input_channels * g.N * d.stride_t;
  g.out_frame_step = (long long)g.modules * g.CoutT * g.N;
  g.img_total = (long long)g.N * g.W * g.H * img.shape[3];
  g.out_total = (long long)g.N * g.modules * out.shape[3];
  if (img_m) {
    CNB_REQUIRE(img_m->size[0] == g.N, what);
    CNB_REQUIRE((long long)img_m->size[1] == (long long)g.W * g.H * img.shape[3], what);
    CNB_REQUIRE(!img_m->is_trans, what);
  }
  if (out_m) {
    CNB_REQUIRE(out_m->size[0] == g.N, what);
    CNB_REQUIRE((long long)out_m->size[1] == (long long)g.modules * out.shape[3], what);
    CNB_REQUIRE(!out_m->is_trans, what);
  }
  if (flt_m) {
    CNB_REQUIRE(flt_m->size[0] == g.Cout, what);
    CNB_REQUIRE((long long)flt_m->size[1] == (long long)g.K * fmm, what);
    CNB_REQUIRE(!flt_m->is_trans, what);
  }
  return g;
}

struct PoolGeom {
  int N, W, H, T, C, modX, modY, modT;
  int kx, ky, kt, sx, sy, st, px, py, pt;
};

inline PoolGeom pool_geom(const Shape4D& img, const Shape4D& out, const cudamat* img_m,
                          const cudamat* out_m, ConvDesc d, const char* what) {
  PoolGeom g;
  CNB_REQUIRE(d.num_input_channels > 0 && d.num_output_channels > 0, what);
  g.N = img.shape[0]; g.W = img.shape[1]; g.H = img.shape[2];
  g.C = d.num_input_channels;
  g.T = img.shape[3] / d.num_input_channels;
  g.modX = out.shape[1]; g.modY = out.shape[2];
  g.modT = out.shape[3] / d.num_output_channels;
  CNB_REQUIRE(img.shape[0] == out.shape[0], what);
  CNB_REQUIRE(img.shape[3] % d.num_input_channels == 0, what);
  CNB_REQUIRE(out.shape[3] % d.num_output_channels == 0, what);
  CNB_REQUIRE(d.num_input_channels == d.num_output_channels, what);
  g.kx = d.kernel_size_x; g.ky = d.kernel_size_y; g.kt = d.kernel_size_t > 0 ? d.kernel_size_t : 1;
  g.sx = d.stride_x; g.sy = d.stride_y; g.st = d.stride_t > 0 ? d.stride_t : 1;
  g.px = d.padding_x; g.py = d.padding_y; g.pt = d.padding_t;
  CNB_REQUIRE(g.kx > 0 && g.ky > 0 && g.sx > 0 && g.sy > 0, what);
  if (img_m) {
    CNB_REQUIRE(img_m->size[0] == g.N, what);
    CNB_REQUIRE((long long)img_m->size[1] == (long long)g.W * g.H * g.C * g.T, what);
  }
  if (out_m) {
    CNB_REQUIRE(out_m->size[0] == g.N, what);
    CNB_REQUIRE((long long)out_m->size[1] == (long long)g.modX * g.modY * g.C * g.modT, what);
  }
  return g;
}

}  // namespace cnb
