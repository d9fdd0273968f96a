// G2d output head in ONE pass (reference model.py:748-751, 763):
//     GroupNorm(32, 64) -> ReLU -> Conv2d(64, 3, 3, padding=1) -> Sigmoid
// x is the fp32 channels-last output of the last decoder block [N,H,W,64]; the GroupNorm arrives as the per-(sample,
// channel) scale/shift pairs of mp_gn_finalize.  The un-fused chain (affine pass writing split planes, a 64->27 tensor-core
// GEMM, a shift-and-add pass) moved 2.1 GB four times; here x is read once and only the 3-channel image is written.
//
// CTA = 8 x 32 output pixels; 256 threads = two channel halves (32 input channels each, partial sums merged through
// shared memory) x 128 pixel-pair owners (two vertically adjacent pixels per thread).  The normalised, rectified
// 10 x 34 halo tile is staged channel-major in shared memory (odd pitch: conflict-free transposed fill, conflict-free
// row reads); the 64*9*3 weights travel as a __grid_constant__ kernel parameter, so every FFMA takes its weight straight
// from the constant bank.  fp32 FMA throughout.  Two CTAs per SM: one fills while the other computes.
#include "common.cuh"

namespace {

constexpr int HC = 64;             // input channels
constexpr int HO = 3;              // output channels
constexpr int TH = 8, TW = 32;     // output tile
constexpr int PH = TH + 2, PW = TW + 2;
constexpr int PITCH = PH * PW + 1; // 341 floats per channel (odd)
constexpr int HEAD_THREADS = 256;   // two channel halves x (4 row pairs x 32 columns)

struct HeadWeights {
  float w[HC][9][HO];              // [cin][kh*3+kw][cout]
  float bias[HO];
};

__global__ void __launch_bounds__(HEAD_THREADS, 2)
k_gn_relu_conv3x3_head(const float* __restrict__ x, const float* __restrict__ ab, float* __restrict__ out, int H, int W,
                       int act, const __grid_constant__ HeadWeights wt) {
  extern __shared__ float s[];     // [HC][PITCH]
  const int n = blockIdx.z, y0 = blockIdx.y * TH, x0 = blockIdx.x * TW;
  const int tid = threadIdx.x;
  // ---- fill: (pixel, 4-channel vector) items; a thread always owns the same four channels
  {
    const int c4 = tid & 15;
    const float4 a = make_float4(ab[((int64_t)n * HC + c4 * 4 + 0) * 2], ab[((int64_t)n * HC + c4 * 4 + 1) * 2],
                                 ab[((int64_t)n * HC + c4 * 4 + 2) * 2], ab[((int64_t)n * HC + c4 * 4 + 3) * 2]);
    const float4 b = make_float4(ab[((int64_t)n * HC + c4 * 4 + 0) * 2 + 1], ab[((int64_t)n * HC + c4 * 4 + 1) * 2 + 1],
                                 ab[((int64_t)n * HC + c4 * 4 + 2) * 2 + 1], ab[((int64_t)n * HC + c4 * 4 + 3) * 2 + 1]);
    // batches of FB independent 16-byte loads in flight per thread before the first dependent use: the fill is
    // latency-bound otherwise (8 warps per SM)
    constexpr int FB = 8, STEP = HEAD_THREADS / 16;
    for (int p0 = tid >> 4; p0 < PH * PW; p0 += FB * STEP) {
      float4 q[FB];
      bool ok[FB];
#pragma unroll
      for (int j = 0; j < FB; ++j) {
        const int pix = p0 + j * STEP;
        const int py = pix / PW, px = pix - py * PW;
        const int yy = y0 + py - 1, xx = x0 + px - 1;
        ok[j] = pix < PH * PW && yy >= 0 && yy < H && xx >= 0 && xx < W;
        q[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ok[j]) q[j] = __ldg(reinterpret_cast<const float4*>(x + (((int64_t)n * H + yy) * W + xx) * HC + c4 * 4));
      }
#pragma unroll
      for (int j = 0; j < FB; ++j) {
        const int pix = p0 + j * STEP;
        if (pix < PH * PW) {
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);      // the convolution zero-pads its (rectified) input
          if (ok[j]) {
            v.x = fmaxf(fmaf(q[j].x, a.x, b.x), 0.f); v.y = fmaxf(fmaf(q[j].y, a.y, b.y), 0.f);
            v.z = fmaxf(fmaf(q[j].z, a.z, b.z), 0.f); v.w = fmaxf(fmaf(q[j].w, a.w, b.w), 0.f);
          }
          float* d = s + (c4 * 4) * PITCH + pix;
          d[0] = v.x; d[PITCH] = v.y; d[2 * PITCH] = v.z; d[3 * PITCH] = v.w;
        }
      }
    }
  }
  __syncthreads();
  // ---- compute: thread = column tx, rows 2*ty and 2*ty + 1 of the tile
  const int tx = tid & 31, ty = (tid >> 5) & 3, chalf = tid >> 7;
  float acc0[HO], acc1[HO];
#pragma unroll
  for (int o = 0; o < HO; ++o) acc0[o] = acc1[o] = chalf ? 0.f : wt.bias[o];
  const float* sp = s + (2 * ty) * PW + tx;
  if (chalf == 0) {
    // partially unrolled: the fully unrolled body (3 500 FFMAs, ~70 KB of SASS) thrashed the instruction cache (ncu:
    // "no instruction" was the top stall); with a rolled loop the weights are fetched by uniform-indexed LDCU
#pragma unroll 2
    for (int c = 0; c < HC / 2; ++c) {
      float v[4][3];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int k = 0; k < 3; ++k) v[r][k] = sp[c * PITCH + r * PW + k];
#pragma unroll
      for (int kh = 0; kh < 3; ++kh)
#pragma unroll
        for (int kw = 0; kw < 3; ++kw)
#pragma unroll
          for (int o = 0; o < HO; ++o) {
            acc0[o] = fmaf(v[kh][kw], wt.w[c][kh * 3 + kw][o], acc0[o]);
            acc1[o] = fmaf(v[kh + 1][kw], wt.w[c][kh * 3 + kw][o], acc1[o]);
          }
    }
  } else {
#pragma unroll 2
  for (int c = HC / 2; c < HC; ++c) {
    float v[4][3];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int k = 0; k < 3; ++k) v[r][k] = sp[c * PITCH + r * PW + k];
#pragma unroll
    for (int kh = 0; kh < 3; ++kh)
#pragma unroll
      for (int kw = 0; kw < 3; ++kw)
#pragma unroll
        for (int o = 0; o < HO; ++o) {
          acc0[o] = fmaf(v[kh][kw], wt.w[c][kh * 3 + kw][o], acc0[o]);
          acc1[o] = fmaf(v[kh + 1][kw], wt.w[c][kh * 3 + kw][o], acc1[o]);
        }
  }
  }
  // merge the two channel halves (fixed order: lower half + upper half), then the lower-half threads store
  __syncthreads();
  float* red = s;                  // the tile is dead: reuse [6][128] floats
  if (chalf) {
#pragma unroll
    for (int o = 0; o < HO; ++o) {
      red[(2 * o) * 128 + (tid & 127)] = acc0[o];
      red[(2 * o + 1) * 128 + (tid & 127)] = acc1[o];
    }
  }
  __syncthreads();
  if (chalf) return;
#pragma unroll
  for (int o = 0; o < HO; ++o) {
    acc0[o] += red[(2 * o) * 128 + tid];
    acc1[o] += red[(2 * o + 1) * 128 + tid];
  }
  const int ya = y0 + 2 * ty, xa = x0 + tx;
  if (xa < W) {
#pragma unroll
    for (int o = 0; o < HO; ++o) {
      float* po = out + (((int64_t)n * HO + o) * H) * W + xa;
      if (ya < H) po[(int64_t)ya * W] = mp_apply_act(acc0[o], act);
      if (ya + 1 < H) po[(int64_t)(ya + 1) * W] = mp_apply_act(acc1[o], act);
    }
  }
}

}  // namespace

extern "C" int mp_gn_relu_conv3x3_head(const float* x, const float* ab, const float* weight_host, const float* bias_host,
                                       float* out, int N, int H, int W, int Cin, int Cout, int act, void* stream) {
  MP_REQUIRE(x && ab && weight_host && out, "mp_gn_relu_conv3x3_head: null pointer");
  MP_REQUIRE(Cin == HC && Cout == HO, "mp_gn_relu_conv3x3_head: only the 64 -> 3 head is instantiated");
  MP_REQUIRE(N > 0 && N <= 65535 && H > 0 && W > 0 && (H + TH - 1) / TH <= 65535, "mp_gn_relu_conv3x3_head: bad dims");
  MP_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, "mp_gn_relu_conv3x3_head: x must be 16-byte aligned");
  HeadWeights wt;
  for (int o = 0; o < HO; ++o) {
    wt.bias[o] = bias_host ? bias_host[o] : 0.f;
    for (int c = 0; c < HC; ++c)
      for (int t = 0; t < 9; ++t) wt.w[c][t][o] = weight_host[(o * HC + c) * 9 + t];     // OIHW
  }
  const size_t smem = (size_t)HC * PITCH * sizeof(float);
  static bool attr_done[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !attr_done[dev]) {
    cudaError_t e = cudaFuncSetAttribute(k_gn_relu_conv3x3_head, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    MP_REQUIRE(e == cudaSuccess, "mp_gn_relu_conv3x3_head: cannot opt in to %zu B shared memory: %s", smem,
               cudaGetErrorString(e));
    if (dev >= 0 && dev < 64) attr_done[dev] = true;
  }
  dim3 grid((unsigned)((W + TW - 1) / TW), (unsigned)((H + TH - 1) / TH), (unsigned)N);
  k_gn_relu_conv3x3_head<<<grid, HEAD_THREADS, smem, mp_stream(stream)>>>(x, ab, out, H, W, act, wt);
  MP_LAUNCH_CHECK("mp_gn_relu_conv3x3_head");
  return 0;
}
