// The decoders' 1x1 tail, TRAIN:285-290 (albedo) / 345-350 (depth): conv_c2_2 + BN + LeakyReLU, conv_c2_3 + BN + LeakyReLU,
// conv_c2_o (+ sigmoid, or x100 for the depth) on one pixel's 16 channels.  Shared by the stand-alone kernel
// (cnn_aux.cu: head_1x1_kernel) and the fused epilogue of the last 3x3 layer (conv_p16.cu, HEAD = true).  The weights travel as
// a __grid_constant__ kernel parameter: every multiply-add takes its weight straight from the constant bank.
#pragma once

namespace gfr_head {

struct HeadWeights { float w2[16][16]; float b2[16]; float w3[16][16]; float b3[16]; float wo[3][16]; float bo[3]; };   // [co][ci]

// x: the 16 input channels (clobbered).  out[o], o < n_out <= 3.  act: 0 none, 2 sigmoid.
__device__ __forceinline__ void apply(const HeadWeights& wt, float (&x)[16], int n_out, int act, float scale, float (&out)[3]) {
  float h[16];
#pragma unroll
  for (int o = 0; o < 16; ++o) {
    float s = wt.b2[o];
#pragma unroll
    for (int c = 0; c < 16; ++c) s = fmaf(wt.w2[o][c], x[c], s);
    h[o] = s > 0.f ? s : 0.2f * s;
  }
#pragma unroll
  for (int o = 0; o < 16; ++o) {
    float s = wt.b3[o];
#pragma unroll
    for (int c = 0; c < 16; ++c) s = fmaf(wt.w3[o][c], h[c], s);
    x[o] = s > 0.f ? s : 0.2f * s;
  }
#pragma unroll
  for (int o = 0; o < 3; ++o) {
    float s = wt.bo[o];
#pragma unroll
    for (int c = 0; c < 16; ++c) s = fmaf(wt.wo[o][c], x[c], s);
    if (act == 2) s = 1.0f / (1.0f + expf(-s));
    out[o] = o < n_out ? s * scale : 0.f;
  }
}

// host side: the six host arrays -> the parameter block
inline void fill(HeadWeights& wt, const float* w2_host, const float* b2_host, const float* w3_host, const float* b3_host,
                 const float* wo_host, const float* bo_host, int n_out) {
  for (int o = 0; o < 16; ++o) {
    wt.b2[o] = b2_host[o]; wt.b3[o] = b3_host[o];
    for (int c = 0; c < 16; ++c) { wt.w2[o][c] = w2_host[o * 16 + c]; wt.w3[o][c] = w3_host[o * 16 + c]; }
  }
  for (int o = 0; o < 3; ++o) {
    wt.bo[o] = o < n_out ? bo_host[o] : 0.f;
    for (int c = 0; c < 16; ++c) wt.wo[o][c] = o < n_out ? wo_host[o * 16 + c] : 0.f;
  }
}

}  // namespace gfr_head
