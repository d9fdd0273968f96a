// Shared helpers for libmpb200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mpb200.h"

int mp_set_error(const char* fmt, ...);

#define MP_REQUIRE(cond, ...)                  \
  do {                                         \
    if (!(cond)) return mp_set_error(__VA_ARGS__); \
  } while (0)

#define MP_LAUNCH_CHECK(name)                                                          \
  do {                                                                                 \
    cudaError_t e__ = cudaGetLastError();                                              \
    if (e__ != cudaSuccess) return mp_set_error("%s: %s", name, cudaGetErrorString(e__)); \
  } while (0)

static inline cudaStream_t mp_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

typedef __nv_bfloat16 bf16;

// x ~= hi + lo with 16 mantissa bits in total (both round-to-nearest-even).
__device__ __forceinline__ void mp_split2(float x, bf16& hi, bf16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}
__device__ __forceinline__ float mp_join(bf16 hi, bf16 lo) { return __bfloat162float(hi) + __bfloat162float(lo); }

__device__ __forceinline__ float mp_apply_act(float v, int act) {
  switch (act) {
    case MP_ACT_RELU: return fmaxf(v, 0.f);
    case MP_ACT_RELU_TANH: return tanhf(fmaxf(v, 0.f));
    case MP_ACT_SIGMOID: return 1.f / (1.f + expf(-v));
    case MP_ACT_TANH: return tanhf(v);
    default: return v;
  }
}

// fp16 single-plane activations (MP_PREC_F16X2 convolutions): round-to-nearest-even, saturating at the fp16 range.
typedef __half f16;
__device__ __forceinline__ f16 mp_to_f16(float x) { return __float2half_rn(fminf(fmaxf(x, -65504.f), 65504.f)); }
struct __align__(16) f16x8 {
  f16 v[8];
};
struct __align__(8) f16x4 {
  f16 v[4];
};

// "F16_Q8" operand format (MP_PREC_F16_Q8): fp16 plane + byte plane; per 64-channel group the byte plane holds
// [64 x e4m3(x)] [64 x e4m3((x - fp16(x)) * 2048)].  Byte offset of channel ch inside a position's 2*C bytes:
__device__ __forceinline__ int mp_q8_off(int ch) { return ((ch >> 6) << 7) + (ch & 63); }
__device__ __forceinline__ uint16_t mp_e4m3x2(float a, float b) {
  return (uint16_t)__nv_cvt_float2_to_fp8x2(make_float2(a, b), __NV_SATFINITE, __NV_E4M3);
}
__device__ __forceinline__ float mp_e4m3_to_float(uint8_t v) {
  return __half2float(__half(__nv_cvt_fp8_to_halfraw((__nv_fp8_storage_t)v, __NV_E4M3)));
}
// 8 consecutive channels -> (fp16 x 8, e4m3(x) x 8, e4m3((x - fp16 x) * 2048) x 8)
// `s` = the per-tensor power-of-two scale of the byte plane (exact: it only shifts exponents)
__device__ __forceinline__ void mp_hq_pack8(const float* x, f16x8& h, uint2& a8, uint2& al8, float s = 1.f) {
  float l[8], xs[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    h.v[i] = mp_to_f16(x[i]);
    l[i] = (x[i] - __half2float(h.v[i])) * (2048.f * s);
    xs[i] = x[i] * s;
  }
  a8.x = (uint32_t)mp_e4m3x2(xs[0], xs[1]) | ((uint32_t)mp_e4m3x2(xs[2], xs[3]) << 16);
  a8.y = (uint32_t)mp_e4m3x2(xs[4], xs[5]) | ((uint32_t)mp_e4m3x2(xs[6], xs[7]) << 16);
  al8.x = (uint32_t)mp_e4m3x2(l[0], l[1]) | ((uint32_t)mp_e4m3x2(l[2], l[3]) << 16);
  al8.y = (uint32_t)mp_e4m3x2(l[4], l[5]) | ((uint32_t)mp_e4m3x2(l[6], l[7]) << 16);
}

// 4 consecutive channels <-> 8-byte bf16x4 vectors
struct __align__(8) bf16x4 {
  bf16 v[4];
};

__device__ __forceinline__ void mp_store_split4(bf16* hi, bf16* lo, int64_t idx, float4 v) {
  bf16x4 h, l;
  mp_split2(v.x, h.v[0], l.v[0]);
  mp_split2(v.y, h.v[1], l.v[1]);
  mp_split2(v.z, h.v[2], l.v[2]);
  mp_split2(v.w, h.v[3], l.v[3]);
  *reinterpret_cast<bf16x4*>(hi + idx) = h;
  *reinterpret_cast<bf16x4*>(lo + idx) = l;
}
__device__ __forceinline__ float4 mp_load_split4(const bf16* hi, const bf16* lo, int64_t idx) {
  bf16x4 h = *reinterpret_cast<const bf16x4*>(hi + idx);
  bf16x4 l = *reinterpret_cast<const bf16x4*>(lo + idx);
  return make_float4(mp_join(h.v[0], l.v[0]), mp_join(h.v[1], l.v[1]), mp_join(h.v[2], l.v[2]),
                     mp_join(h.v[3], l.v[3]));
}
