// fp16 single-plane operators of the pooled motion-encoder trunks (Emtn, reference model.py:869-907): the HBM-bound
// kernels that sit between the MP_PREC_F16X2 tensor-core convolutions.  All tensors are channels-last.
#include <math.h>

#include "common.cuh"

namespace {

// ------------------------------------------------------------------------------------------------ RGB stem patches
// One thread = one output position: 9*C plane reads (coalesced across the warp: neighbouring threads read neighbouring
// pixels of the same plane) -> one 64-byte row of 32 fp16 patch channels, channel (kh*3+kw)*C + c.
template <int C>
__global__ void __launch_bounds__(256)
k_im2col3x3_f16(const float* __restrict__ in, f16* __restrict__ out, int H, int W, int stride) {
  const int Ho = H / stride, Wo = W / stride;
  const int64_t So = (int64_t)Ho * Wo;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= So) return;
  const int n = blockIdx.y;
  const int wo = (int)(t % Wo), ho = (int)(t / Wo);
  const float* base = in + (int64_t)n * C * H * W;
  f16x8 row[4];
#pragma unroll
  for (int i = 0; i < 32; ++i) row[i >> 3].v[i & 7] = __float2half_rn(0.f);
#pragma unroll
  for (int kh = 0; kh < 3; ++kh) {
    const int y = ho * stride + kh - 1;
#pragma unroll
    for (int kw = 0; kw < 3; ++kw) {
      const int x = wo * stride + kw - 1;
      const bool ok = y >= 0 && y < H && x >= 0 && x < W;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const int ch = (kh * 3 + kw) * C + c;
        const float v = ok ? __ldg(base + ((int64_t)c * H + y) * W + x) : 0.f;
        row[ch >> 3].v[ch & 7] = mp_to_f16(v);
      }
    }
  }
  f16x8* o = reinterpret_cast<f16x8*>(out + ((int64_t)n * So + t) * 32);
#pragma unroll
  for (int i = 0; i < 4; ++i) o[i] = row[i];
}

// ------------------------------------------------------------------------------------------------ MaxPool2d(3, 2, 1)
// One thread = one output pixel x 8 channels (16 bytes); fp16 max is exact.
__global__ void __launch_bounds__(256)
k_maxpool3x3s2_cl_f16(const f16* __restrict__ in, f16* __restrict__ out, int H, int W, int C) {
  const int C8 = C >> 3, Ho = H >> 1, Wo = W >> 1;
  const int idx = blockIdx.x * 256 + threadIdx.x;
  if (idx >= Wo * C8) return;
  const int wo = idx / C8, c8 = idx - wo * C8;
  const int ho = blockIdx.y, n = blockIdx.z;
  __half2 m[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) m[i] = __float2half2_rn(-65504.f);
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const int y = ho * 2 + a - 1;
    if (y < 0 || y >= H) continue;
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      const int x = wo * 2 + b - 1;
      if (x < 0 || x >= W) continue;
      const uint4 raw = *reinterpret_cast<const uint4*>(in + (((int64_t)n * H + y) * W + x) * C + c8 * 8);
      const __half2* h = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
      for (int k = 0; k < 4; ++k) m[k] = __hmax2(m[k], h[k]);
    }
  }
  uint4 o;
  __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
  for (int k = 0; k < 4; ++k) oh[k] = m[k];
  *reinterpret_cast<uint4*>(out + ((((int64_t)n * Ho + ho) * Wo + wo) * C8 + c8) * 8) = o;
}

// ------------------------------------------------------------------------------------------------ AdaptiveAvgPool2d(1)
// grid (C/64 column groups, N); block 32 x 8: a thread owns 2 adjacent channels, 8 row-walkers per column pair,
// fp32 accumulation, fixed-order shared-memory reduction (bit-reproducible).
__global__ void k_global_avgpool_cl_f16(const f16* __restrict__ in, float* __restrict__ out, int64_t S, int C) {
  __shared__ float red[8][65];
  const int n = blockIdx.y;
  const int c = blockIdx.x * 64 + threadIdx.x * 2;
  float a0 = 0.f, a1 = 0.f;
  if (c < C) {
    for (int64_t s = threadIdx.y; s < S; s += 8) {
      const __half2 v = *reinterpret_cast<const __half2*>(in + ((int64_t)n * S + s) * C + c);
      const float2 f = __half22float2(v);
      a0 += f.x; a1 += f.y;
    }
  }
  red[threadIdx.y][threadIdx.x * 2] = a0;
  red[threadIdx.y][threadIdx.x * 2 + 1] = a1;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    float t0 = 0.f, t1 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { t0 += red[i][threadIdx.x * 2]; t1 += red[i][threadIdx.x * 2 + 1]; }
    out[(int64_t)n * C + c] = t0 / (float)S;
    out[(int64_t)n * C + c + 1] = t1 / (float)S;
  }
}

}  // namespace

extern "C" int mp_im2col3x3_f16(const float* in, void* out_h, int N, int C, int H, int W, int stride, void* stream) {
  MP_REQUIRE(in && out_h, "mp_im2col3x3_f16: null pointer");
  MP_REQUIRE(N > 0 && N <= 65535 && H > 0 && W > 0 && (stride == 1 || stride == 2) && H % stride == 0 && W % stride == 0,
             "mp_im2col3x3_f16: bad dims");
  MP_REQUIRE(C == 3, "mp_im2col3x3_f16: only C = 3 (RGB frames) is instantiated");
  const int64_t So = (int64_t)(H / stride) * (W / stride);
  dim3 grid((unsigned)((So + 255) / 256), (unsigned)N);
  k_im2col3x3_f16<3><<<grid, 256, 0, mp_stream(stream)>>>(in, (f16*)out_h, H, W, stride);
  MP_LAUNCH_CHECK("mp_im2col3x3_f16");
  return 0;
}

extern "C" int mp_maxpool3x3s2_cl_f16(const void* in_h, void* out_h, int N, int H, int W, int C, void* stream) {
  MP_REQUIRE(in_h && out_h, "mp_maxpool3x3s2_cl_f16: null pointer");
  MP_REQUIRE(N > 0 && N <= 65535 && C % 8 == 0 && H % 2 == 0 && W % 2 == 0 && H / 2 <= 65535,
             "mp_maxpool3x3s2_cl_f16: bad dims");
  dim3 grid((unsigned)(((W / 2) * (C / 8) + 255) / 256), (unsigned)(H / 2), (unsigned)N);
  k_maxpool3x3s2_cl_f16<<<grid, 256, 0, mp_stream(stream)>>>((const f16*)in_h, (f16*)out_h, H, W, C);
  MP_LAUNCH_CHECK("mp_maxpool3x3s2_cl_f16");
  return 0;
}

extern "C" int mp_global_avgpool_cl_f16(const void* in_h, float* out, int N, int64_t S, int C, void* stream) {
  MP_REQUIRE(in_h && out, "mp_global_avgpool_cl_f16: null pointer");
  MP_REQUIRE(N > 0 && N <= 65535 && S > 0 && C > 0 && C % 2 == 0, "mp_global_avgpool_cl_f16: bad dims");
  dim3 grid((C + 63) / 64, N), block(32, 8);
  k_global_avgpool_cl_f16<<<grid, block, 0, mp_stream(stream)>>>((const f16*)in_h, out, S, C);
  MP_LAUNCH_CHECK("mp_global_avgpool_cl_f16");
  return 0;
}
