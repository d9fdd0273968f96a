// Coordinate arithmetic shared by the trilinear warping kernels (warp.cu, grid_sample_brick.cu).  It follows ATen
// (GridSampler.h:27-36,58-60; UpSample.h area_pixel_compute_source_index; RangeFactories linspace) operation by
// operation so that results agree with the CPU oracle to ~1e-6.
#pragma once
#include "common.cuh"

// ------------------------------------------------------------------------------------------------ shared math
struct Taps {
  int off[8];    // element offsets inside one (sample, channel) volume, or -1 when skipped
  float w[8];
};

// pix coordinates are already clamped to [0, size-1] (padding_mode='border')
__device__ __forceinline__ void make_taps(float ix, float iy, float iz, int D, int H, int W, Taps& t) {
  float fx = floorf(ix), fy = floorf(iy), fz = floorf(iz);
  int x0 = (int)fx, y0 = (int)fy, z0 = (int)fz;
  float tx = ix - fx, ty = iy - fy, tz = iz - fz;   // weight of the +1 corner
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    int dx = k & 1, dy = (k >> 1) & 1, dz = k >> 2;
    int x = x0 + dx, y = y0 + dy, z = z0 + dz;
    float w = (dx ? tx : 1.f - tx) * (dy ? ty : 1.f - ty) * (dz ? tz : 1.f - tz);
    bool ok = x >= 0 && x < W && y >= 0 && y < H && z >= 0 && z < D;   // within_bounds_3d
    t.off[k] = ok ? (z * H + y) * W + x : -1;
    t.w[k] = ok ? w : 0.f;
  }
}

__device__ __forceinline__ float unnormalize_clip(float g, int size) {
  float p = ((g + 1.f) / 2.f) * (float)(size - 1);            // grid_sampler_unnormalize, align_corners=True
  return fminf((float)(size - 1), fmaxf(p, 0.f));            // clip_coordinates
}

// torch.linspace(-1, 1, steps)[i] (RangeFactories: symmetric evaluation around the midpoint)
__device__ __forceinline__ float linspace_m1_1(int i, int steps) {
  if (steps == 1) return -1.f;
  float step = 2.f / (float)(steps - 1);
  return (i < steps / 2) ? (-1.f + step * (float)i) : (1.f - step * (float)(steps - i - 1));
}

// align_corners=True source index (UpSample.h)
__device__ __forceinline__ void src_ac_true(int dst, int in_size, int out_size, int& i0, int& i1, float& l1) {
  float scale = out_size > 1 ? (float)(in_size - 1) / (float)(out_size - 1) : 0.f;
  float src = scale * (float)dst;
  i0 = (int)src;
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  l1 = src - (float)i0;
}
// align_corners=False source index (UpSample.h: scale*(dst+0.5)-0.5 clamped at 0)
__device__ __forceinline__ void src_ac_false(int dst, int in_size, int out_size, int& i0, int& i1, float& l1) {
  float scale = (float)in_size / (float)out_size;
  float src = scale * ((float)dst + 0.5f) - 0.5f;
  if (src < 0.f) src = 0.f;
  i0 = (int)src;
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  l1 = src - (float)i0;
}

// One thread = one output voxel (x fastest => coalesced grid reads and output writes) x a chunk of channels.
constexpr int GS_CCHUNK = 8;

// Gather `c1 - c0` channels for one output voxel.  Channels are processed four at a time with all 32 tap loads
// issued before the first use, so each thread keeps 32 independent requests in flight (the kernel is latency-bound,
// not bandwidth-bound, when written one channel at a time).
__device__ __forceinline__ void gather_channels(const float* __restrict__ v, float* __restrict__ out, const Taps& t,
                                                int c0, int c1, int64_t in_cs, int64_t out_cs) {
  int off[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) off[k] = t.off[k] >= 0 ? t.off[k] : t.off[0];   // skipped corners: weight 0, valid address
  int c = c0;
  for (; c + 4 <= c1; c += 4) {
    float x[4][8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float* p = v + (int64_t)(c + j) * in_cs;
#pragma unroll
      for (int k = 0; k < 8; ++k) x[j][k] = __ldg(p + off[k]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) acc = fmaf(x[j][k], t.w[k], acc);
      out[(int64_t)(c + j) * out_cs] = acc;
    }
  }
  for (; c < c1; ++c) {
    const float* p = v + (int64_t)c * in_cs;
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) acc = fmaf(__ldg(p + off[k]), t.w[k], acc);
    out[(int64_t)c * out_cs] = acc;
  }
}

// flow value of channel k at output voxel (d,h,w): trilinear align_corners=True resample of wf [3,Df,Hf,Wf]
__device__ __forceinline__ float resample_flow(const float* __restrict__ wf, int Df, int Hf, int Wf, int d0, int d1,
                                               float ld, int h0, int h1, float lh, int w0, int w1, float lw) {
  auto at = [&](int z, int y, int x) { return __ldg(wf + ((int64_t)z * Hf + y) * Wf + x); };
  float a = (1.f - lh) * ((1.f - lw) * at(d0, h0, w0) + lw * at(d0, h0, w1)) +
            lh * ((1.f - lw) * at(d0, h1, w0) + lw * at(d0, h1, w1));
  float b = (1.f - lh) * ((1.f - lw) * at(d1, h0, w0) + lw * at(d1, h0, w1)) +
            lh * ((1.f - lw) * at(d1, h1, w0) + lw * at(d1, h1, w1));
  return (1.f - ld) * a + ld * b;
}

