// Layout conversion, pooling and resampling kernels on channels-last (CL) tensors.  All are HBM-bound:
// one read + one write per element, float4 / bf16x4 vector accesses along the channel axis.
#include <stdarg.h>
#include <stdio.h>

#include <stdlib.h>

#include "common.cuh"

// ------------------------------------------------------------------------------------------------ errors / misc
static thread_local char g_err[512] = "";

int mp_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return 1;
}

extern "C" const char* mp_last_error(void) { return g_err; }
extern "C" int mp_abi_version(void) { return MPB200_ABI_VERSION; }
extern "C" int mp_device_supported(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return -1;
  return major == 10 ? 1 : 0;
}

// ------------------------------------------------------------------------------------------------ NCDHW <-> CL
// in [N][C][S] -> out [N][S][C]; 32x32 tile through shared memory, coalesced on both sides.
__global__ void k_nchw_to_cl(const float* __restrict__ in, float* __restrict__ out_f32, bf16* __restrict__ out_hi,
                             bf16* __restrict__ out_lo, int C, int64_t S) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int64_t s0 = (int64_t)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const float* src = in + (int64_t)n * C * S;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int c = c0 + j;
    int64_t s = s0 + threadIdx.x;
    tile[j][threadIdx.x] = (c < C && s < S) ? src[(int64_t)c * S + s] : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int64_t s = s0 + j;
    int c = c0 + threadIdx.x;
    if (s < S && c < C) {
      float v = tile[threadIdx.x][j];
      int64_t o = ((int64_t)n * S + s) * C + c;
      if (out_f32) out_f32[o] = v;
      if (out_hi) {
        bf16 h, l;
        mp_split2(v, h, l);
        out_hi[o] = h;
        out_lo[o] = l;
      }
    }
  }
}

extern "C" int mp_nchw_to_cl(const float* in, float* out_f32, void* out_hi, void* out_lo, int N, int C, int64_t S,
                             void* stream) {
  MP_REQUIRE(in && (out_f32 || (out_hi && out_lo)), "mp_nchw_to_cl: null pointer");
  MP_REQUIRE(N > 0 && C > 0 && S > 0 && N <= 65535, "mp_nchw_to_cl: bad dims");
  dim3 grid((unsigned)((S + 31) / 32), (C + 31) / 32, N), block(32, 8);
  k_nchw_to_cl<<<grid, block, 0, mp_stream(stream)>>>(in, out_f32, (bf16*)out_hi, (bf16*)out_lo, C, S);
  MP_LAUNCH_CHECK("mp_nchw_to_cl");
  return 0;
}

__global__ void k_cl_to_nchw(const float* __restrict__ in_f32, const bf16* __restrict__ in_hi,
                             const bf16* __restrict__ in_lo, float* __restrict__ out, int C, int64_t S) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int64_t s0 = (int64_t)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int64_t s = s0 + j;
    int c = c0 + threadIdx.x;
    float v = 0.f;
    if (s < S && c < C) {
      int64_t o = ((int64_t)n * S + s) * C + c;
      v = in_f32 ? in_f32[o] : mp_join(in_hi[o], in_lo[o]);
    }
    tile[j][threadIdx.x] = v;
  }
  __syncthreads();
  float* dst = out + (int64_t)n * C * S;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int c = c0 + j;
    int64_t s = s0 + threadIdx.x;
    if (c < C && s < S) dst[(int64_t)c * S + s] = tile[threadIdx.x][j];
  }
}

extern "C" int mp_cl_to_nchw(const float* in_f32, const void* in_hi, const void* in_lo, float* out, int N, int C,
                             int64_t S, void* stream) {
  MP_REQUIRE(out && (in_f32 || (in_hi && in_lo)), "mp_cl_to_nchw: null pointer");
  MP_REQUIRE(N > 0 && C > 0 && S > 0 && N <= 65535, "mp_cl_to_nchw: bad dims");
  dim3 grid((unsigned)((S + 31) / 32), (C + 31) / 32, N), block(32, 8);
  k_cl_to_nchw<<<grid, block, 0, mp_stream(stream)>>>(in_f32, (const bf16*)in_hi, (const bf16*)in_lo, out, C, S);
  MP_LAUNCH_CHECK("mp_cl_to_nchw");
  return 0;
}

__global__ void k_split(const float* __restrict__ in, bf16* __restrict__ hi, bf16* __restrict__ lo, int64_t n) {
  int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    float4 v = *reinterpret_cast<const float4*>(in + i);
    mp_store_split4(hi, lo, i, v);
  } else {
    for (; i < n; ++i) mp_split2(in[i], hi[i], lo[i]);
  }
}

extern "C" int mp_split(const float* in, void* out_hi, void* out_lo, int64_t n, void* stream) {
  MP_REQUIRE(in && out_hi && out_lo && n > 0, "mp_split: bad args");
  int64_t q = (n + 3) / 4;
  k_split<<<(unsigned)((q + 255) / 256), 256, 0, mp_stream(stream)>>>(in, (bf16*)out_hi, (bf16*)out_lo, n);
  MP_LAUNCH_CHECK("mp_split");
  return 0;
}

// ------------------------------------------------------------------------------------------------ AvgPool(2)
// One thread = one output position x 4 channels.
__global__ void k_avgpool2_cl(const float* __restrict__ in, float* __restrict__ out_f32, bf16* __restrict__ out_hi,
                              bf16* __restrict__ out_lo, int N, int D, int H, int W, int C, int pd) {
  const int C4 = C >> 2;
  const int Do = D / pd, Ho = H >> 1, Wo = W >> 1;
  int64_t total = (int64_t)N * Do * Ho * Wo * C4;
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  int c4 = (int)(t % C4);
  int64_t p = t / C4;
  int wo = (int)(p % Wo); p /= Wo;
  int ho = (int)(p % Ho); p /= Ho;
  int d_o = (int)(p % Do);
  int n = (int)(p / Do);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int dd = 0; dd < pd; ++dd)
    for (int hh = 0; hh < 2; ++hh)
      for (int ww = 0; ww < 2; ++ww) {
        int64_t idx = ((((int64_t)n * D + (d_o * pd + dd)) * H + (ho * 2 + hh)) * W + (wo * 2 + ww)) * C + c4 * 4;
        float4 v = *reinterpret_cast<const float4*>(in + idx);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
  const float inv = 1.f / (4.f * pd);
  acc.x *= inv; acc.y *= inv; acc.z *= inv; acc.w *= inv;
  int64_t o = t * 4;
  if (out_f32) *reinterpret_cast<float4*>(out_f32 + o) = acc;
  if (out_hi) mp_store_split4(out_hi, out_lo, o, acc);
}

extern "C" int mp_avgpool2_cl(const float* in, float* out_f32, void* out_hi, void* out_lo, int N, int D, int H,
                              int W, int C, int pool_d, void* stream) {
  MP_REQUIRE(in && (out_f32 || (out_hi && out_lo)), "mp_avgpool2_cl: null pointer");
  MP_REQUIRE(C % 4 == 0 && H % 2 == 0 && W % 2 == 0 && (pool_d == 1 || (pool_d == 2 && D % 2 == 0)),
             "mp_avgpool2_cl: bad dims C=%d D=%d H=%d W=%d pool_d=%d", C, D, H, W, pool_d);
  int64_t total = (int64_t)N * (D / pool_d) * (H / 2) * (W / 2) * (C / 4);
  k_avgpool2_cl<<<(unsigned)((total + 255) / 256), 256, 0, mp_stream(stream)>>>(in, out_f32, (bf16*)out_hi,
                                                                                (bf16*)out_lo, N, D, H, W, C, pool_d);
  MP_LAUNCH_CHECK("mp_avgpool2_cl");
  return 0;
}

// ------------------------------------------------------------------------------------------------ linear x2, ac=True
// ATen's align_corners=True rule (UpSample.h area_pixel_compute_scale / compute_source_index):
//   scale = (in-1)/(out-1) (float), src = scale * dst, i0 = (int)src, l1 = src - i0, i1 = i0 + (i0 < in-1).
__device__ __forceinline__ void lin_src(int dst, int in_size, int out_size, int& i0, int& i1, float& l1) {
  if (out_size <= 1 || in_size == out_size) {
    i0 = i1 = dst;
    l1 = 0.f;
    if (in_size != out_size) i0 = i1 = 0;
    return;
  }
  float scale = (float)(in_size - 1) / (float)(out_size - 1);
  float src = scale * (float)dst;
  i0 = (int)src;
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  l1 = src - (float)i0;
}

template <bool SPLIT_IN>
__global__ void k_upsample2x_linear_cl(const float* __restrict__ in_f32, const bf16* __restrict__ in_hi,
                                       const bf16* __restrict__ in_lo, float* __restrict__ out_f32,
                                       bf16* __restrict__ out_hi, bf16* __restrict__ out_lo, int N, int D, int H,
                                       int W, int C, int ud) {
  const int C4 = C >> 2;
  const int Do = D * ud, Ho = H * 2, Wo = W * 2;
  int64_t total = (int64_t)N * Do * Ho * Wo * C4;
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  int c4 = (int)(t % C4);
  int64_t p = t / C4;
  int wo = (int)(p % Wo); p /= Wo;
  int ho = (int)(p % Ho); p /= Ho;
  int d_o = (int)(p % Do);
  int n = (int)(p / Do);
  int d0, d1, h0, h1, w0, w1;
  float ld, lh, lw;
  lin_src(d_o, D, Do, d0, d1, ld);
  lin_src(ho, H, Ho, h0, h1, lh);
  lin_src(wo, W, Wo, w0, w1, lw);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int a = 0; a < 2; ++a) {
    if (a == 1 && ud == 1) break;
    float wd = (ud == 1) ? 1.f : (a ? ld : 1.f - ld);
    int dz = a ? d1 : d0;
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      float wh = b ? lh : 1.f - lh;
      int hy = b ? h1 : h0;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        float ww = c ? lw : 1.f - lw;
        int wx = c ? w1 : w0;
        int64_t idx = ((((int64_t)n * D + dz) * H + hy) * W + wx) * C + c4 * 4;
        float4 v;
        if (SPLIT_IN) v = mp_load_split4(in_hi, in_lo, idx);
        else v = *reinterpret_cast<const float4*>(in_f32 + idx);
        // ATen accumulates w_d*w_h*w_w * v (UpSampleKernel.cpp / upsample_trilinear3d)
        float wgt = wd * wh * ww;
        acc.x += wgt * v.x; acc.y += wgt * v.y; acc.z += wgt * v.z; acc.w += wgt * v.w;
      }
    }
  }
  int64_t o = t * 4;
  if (out_f32) *reinterpret_cast<float4*>(out_f32 + o) = acc;
  if (out_hi) mp_store_split4(out_hi, out_lo, o, acc);
}

// Split-in / split-out, 8 channels (16 bytes per plane) per thread: the G2d decoder's bilinear upsamples.
struct __align__(16) bf16x8v {
  bf16 v[8];
};
__global__ void __launch_bounds__(256)
k_upsample2x_bilinear_split8(const bf16* __restrict__ in_hi, const bf16* __restrict__ in_lo, bf16* __restrict__ out_hi,
                             bf16* __restrict__ out_lo, int H, int W, int C) {
  // grid: x = chunks of (wo, c8) pairs of one output row, y = output row, z = sample  (no 64-bit divisions)
  const int C8 = C >> 3;
  const int Ho = H * 2, Wo = W * 2;
  const int idx = blockIdx.x * 256 + threadIdx.x;
  if (idx >= Wo * C8) return;
  const int wo = idx / C8, c8 = idx - wo * C8;
  const int ho = blockIdx.y, n = blockIdx.z;
  int h0, h1, w0, w1;
  float lh, lw;
  lin_src(ho, H, Ho, h0, h1, lh);
  lin_src(wo, W, Wo, w0, w1, lw);
  const int64_t row0 = ((int64_t)n * H + h0) * W, row1 = ((int64_t)n * H + h1) * W;
  const int64_t i00 = (row0 + w0) * C + c8 * 8, i01 = (row0 + w1) * C + c8 * 8;
  const int64_t i10 = (row1 + w0) * C + c8 * 8, i11 = (row1 + w1) * C + c8 * 8;
  // all eight 16-byte loads are issued before the first use
  const bf16x8v h00 = *reinterpret_cast<const bf16x8v*>(in_hi + i00), l00 = *reinterpret_cast<const bf16x8v*>(in_lo + i00);
  const bf16x8v h01 = *reinterpret_cast<const bf16x8v*>(in_hi + i01), l01 = *reinterpret_cast<const bf16x8v*>(in_lo + i01);
  const bf16x8v h10 = *reinterpret_cast<const bf16x8v*>(in_hi + i10), l10 = *reinterpret_cast<const bf16x8v*>(in_lo + i10);
  const bf16x8v h11 = *reinterpret_cast<const bf16x8v*>(in_hi + i11), l11 = *reinterpret_cast<const bf16x8v*>(in_lo + i11);
  const float w00 = (1.f - lh) * (1.f - lw), w01 = (1.f - lh) * lw, w10 = lh * (1.f - lw), w11 = lh * lw;
  bf16x8v oh, ol;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    // same accumulation order as the generic kernel: taps (h0,w0), (h0,w1), (h1,w0), (h1,w1)
    float acc = w00 * mp_join(h00.v[i], l00.v[i]);
    acc += w01 * mp_join(h01.v[i], l01.v[i]);
    acc += w10 * mp_join(h10.v[i], l10.v[i]);
    acc += w11 * mp_join(h11.v[i], l11.v[i]);
    mp_split2(acc, oh.v[i], ol.v[i]);
  }
  const int64_t o = ((((int64_t)n * Ho + ho) * Wo + wo) * C8 + c8) * 8;
  *reinterpret_cast<bf16x8v*>(out_hi + o) = oh;
  *reinterpret_cast<bf16x8v*>(out_lo + o) = ol;
}

// Same operator, one thread = a 2 x 2 block of output pixels x 8 channels.  With align_corners=True the source step is
// (in-1)/(out-1) < 1/2, so the four outputs draw their taps from a 3 x 3 input patch: 18 loads instead of 32, and the
// bf16 pair -> fp32 joins (the bulk of the instruction stream: the one-output-per-thread kernel is issue-bound, ncu
// issue-active 81 %) are done once per patch element.  Tap order and weights are exactly those of the kernel above.
// HQ = true: the outputs are the F16_Q8 plane pair instead (fp16 plane `out_hi`, FP8 byte plane `out_lo` written with the
// per-tensor power-of-two `q8_scale`), feeding the fp16 + FP8 cross-term convolutions of the G2d up-blocks.
// Instruction diet (round 2; the kernel was issue-bound at 37 instructions per output value, ncu issue-active 68 %): (1) interior
// threads (i >= 1 and j >= 1: the four outputs always read patch rows / columns (0,1) and (1,2)) take a path without a single
// select; only the first row / column of a frame runs the generic select path; (2) bf16 pairs are joined and split two at a
// time on 32-bit words (one shift / mask per value, one F2FP per pair).  Same taps, weights and summation order: bit-identical.
__device__ __forceinline__ void split2_pair(float a, float b, uint32_t& hi2, uint32_t& lo2) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);                     // .x = a (low half)
  hi2 = *reinterpret_cast<const uint32_t*>(&h);
  const float ah = __uint_as_float(hi2 << 16), bh = __uint_as_float(hi2 & 0xffff0000u);
  const __nv_bfloat162 l = __floats2bfloat162_rn(a - ah, b - bh);
  lo2 = *reinterpret_cast<const uint32_t*>(&l);
}

// One output pixel x 8 channels from the 3 x 3 patch.  R0 / C0 >= 0: compile-time first patch row / column (interior
// threads); R0 < 0: the run-time flags sh / sw pick rows (0,1) or (1,2) and columns (0,1) or (1,2) (first row / column of a frame).
template <bool HQ, int R0, int C0>
__device__ __forceinline__ void upsample_emit(const float (&P)[3][3][8], float lh, float lw, int64_t pos, int c8, int C,
                                              void* out_hi_v, void* out_lo_v, float q8_scale, bool sh = false, bool sw = false) {
  // taps (h0,w0), (h0,w1), (h1,w0), (h1,w1) in this order, weights (1-lh)(1-lw), (1-lh)lw, lh(1-lw), lh*lw
  const float w00 = (1.f - lh) * (1.f - lw), w01 = (1.f - lh) * lw, w10 = lh * (1.f - lw), w11 = lh * lw;
  float accs[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    float v00, v01, v10, v11;
    if (R0 >= 0) {
      constexpr int r = R0 < 0 ? 0 : R0, c = C0 < 0 ? 0 : C0;
      v00 = P[r][c][k]; v01 = P[r][c + 1][k]; v10 = P[r + 1][c][k]; v11 = P[r + 1][c + 1][k];
    } else {
      const float t0 = sh ? P[1][0][k] : P[0][0][k], t1 = sh ? P[1][1][k] : P[0][1][k], t2 = sh ? P[1][2][k] : P[0][2][k];
      const float b0 = sh ? P[2][0][k] : P[1][0][k], b1 = sh ? P[2][1][k] : P[1][1][k], b2 = sh ? P[2][2][k] : P[1][2][k];
      v00 = sw ? t1 : t0; v01 = sw ? t2 : t1; v10 = sw ? b1 : b0; v11 = sw ? b2 : b1;
    }
    float acc = w00 * v00;
    acc += w01 * v01;
    acc += w10 * v10;
    acc += w11 * v11;
    accs[k] = acc;
  }
  const int64_t o = (pos * (C >> 3) + c8) * 8;
  if (HQ) {
    f16x8 h;
    uint2 a8, al8;
    mp_hq_pack8(accs, h, a8, al8, q8_scale);
    *reinterpret_cast<f16x8*>(reinterpret_cast<f16*>(out_hi_v) + o) = h;
    uint8_t* q = reinterpret_cast<uint8_t*>(out_lo_v) + 2 * pos * C + mp_q8_off(c8 * 8);
    *reinterpret_cast<uint2*>(q) = a8;
    *reinterpret_cast<uint2*>(q + 64) = al8;
  } else {
    uint4 oh, ol;
    split2_pair(accs[0], accs[1], oh.x, ol.x);
    split2_pair(accs[2], accs[3], oh.y, ol.y);
    split2_pair(accs[4], accs[5], oh.z, ol.z);
    split2_pair(accs[6], accs[7], oh.w, ol.w);
    *reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(out_hi_v) + o) = oh;
    *reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(out_lo_v) + o) = ol;
  }
}

template <bool HQ>
__global__ void __launch_bounds__(256)
k_upsample2x_bilinear_split8_2x2(const bf16* __restrict__ in_hi, const bf16* __restrict__ in_lo, void* __restrict__ out_hi_v,
                                 void* __restrict__ out_lo_v, int H, int W, int C, float q8_scale) {
  // grid: x = chunks of (input column j, c8) pairs, y = input row i (output rows 2i, 2i+1), z = sample
  const int C8 = C >> 3;
  const int Ho = H * 2, Wo = W * 2;
  const int idx = blockIdx.x * 256 + threadIdx.x;
  if (idx >= W * C8) return;
  const int j = idx / C8, c8 = idx - j * C8;
  const int i = blockIdx.y, n = blockIdx.z;
  int h0a, h1a, h0b, h1b, w0a, w1a, w0b, w1b;
  float lha, lhb, lwa, lwb;
  lin_src(2 * i, H, Ho, h0a, h1a, lha);
  lin_src(2 * i + 1, H, Ho, h0b, h1b, lhb);
  lin_src(2 * j, W, Wo, w0a, w1a, lwa);
  lin_src(2 * j + 1, W, Wo, w0b, w1b, lwb);
  const bool dh = h0b != h0a, dw = w0b != w0a;
  const int rr[3] = {h0a, min(h0a + 1, H - 1), min(h0a + 2, H - 1)};
  const int cc[3] = {w0a, min(w0a + 1, W - 1), min(w0a + 2, W - 1)};
  float P[3][3][8];
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      const int64_t o = ((((int64_t)n * H + rr[a]) * W + cc[b]) * C8 + c8) * 8;
      const uint4 h = *reinterpret_cast<const uint4*>(in_hi + o), l = *reinterpret_cast<const uint4*>(in_lo + o);
      const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw4[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {     // hi + lo is exact (the planes are the split of one fp32 value)
        P[a][b][2 * k] = __uint_as_float(hw[k] << 16) + __uint_as_float(lw4[k] << 16);
        P[a][b][2 * k + 1] = __uint_as_float(hw[k] & 0xffff0000u) + __uint_as_float(lw4[k] & 0xffff0000u);
      }
    }
  const int64_t pos00 = ((int64_t)n * Ho + 2 * i) * Wo + 2 * j, pos10 = pos00 + Wo;
  if (dh && dw) {      // interior thread (i >= 1 and j >= 1): no selects
    upsample_emit<HQ, 0, 0>(P, lha, lwa, pos00, c8, C, out_hi_v, out_lo_v, q8_scale);
    upsample_emit<HQ, 0, 1>(P, lha, lwb, pos00 + 1, c8, C, out_hi_v, out_lo_v, q8_scale);
    upsample_emit<HQ, 1, 0>(P, lhb, lwa, pos10, c8, C, out_hi_v, out_lo_v, q8_scale);
    upsample_emit<HQ, 1, 1>(P, lhb, lwb, pos10 + 1, c8, C, out_hi_v, out_lo_v, q8_scale);
  } else {
    upsample_emit<HQ, -1, -1>(P, lha, lwa, pos00, c8, C, out_hi_v, out_lo_v, q8_scale, false, false);
    upsample_emit<HQ, -1, -1>(P, lha, lwb, pos00 + 1, c8, C, out_hi_v, out_lo_v, q8_scale, false, dw);
    upsample_emit<HQ, -1, -1>(P, lhb, lwa, pos10, c8, C, out_hi_v, out_lo_v, q8_scale, dh, false);
    upsample_emit<HQ, -1, -1>(P, lhb, lwb, pos10 + 1, c8, C, out_hi_v, out_lo_v, q8_scale, dh, dw);
  }
}

extern "C" int mp_upsample2x_linear_cl(const float* in_f32, const void* in_hi, const void* in_lo, float* out_f32,
                                       void* out_hi, void* out_lo, int N, int D, int H, int W, int C, int up_d,
                                       void* stream) {
  MP_REQUIRE((in_f32 || (in_hi && in_lo)) && (out_f32 || (out_hi && out_lo)), "mp_upsample2x_linear_cl: null pointer");
  MP_REQUIRE(C % 4 == 0 && (up_d == 1 || up_d == 2), "mp_upsample2x_linear_cl: bad dims");
  if (!in_f32 && !out_f32 && D == 1 && up_d == 1 && C % 8 == 0 && H * 2 <= 65535 && N <= 65535) {
    static int one_px = [] { const char* e = getenv("MPB200_UPSAMPLE_1PX"); return (e && atoi(e)) ? 1 : 0; }();
    if (one_px) {
      dim3 g8((unsigned)((W * 2 * (C / 8) + 255) / 256), (unsigned)(H * 2), (unsigned)N);
      k_upsample2x_bilinear_split8<<<g8, 256, 0, mp_stream(stream)>>>((const bf16*)in_hi, (const bf16*)in_lo,
                                                                       (bf16*)out_hi, (bf16*)out_lo, H, W, C);
    } else {
      dim3 g4((unsigned)((W * (C / 8) + 255) / 256), (unsigned)H, (unsigned)N);
      k_upsample2x_bilinear_split8_2x2<false><<<g4, 256, 0, mp_stream(stream)>>>((const bf16*)in_hi, (const bf16*)in_lo,
                                                                                  out_hi, out_lo, H, W, C, 1.f);
    }
    MP_LAUNCH_CHECK("mp_upsample2x_linear_cl");
    return 0;
  }
  int64_t total = (int64_t)N * D * up_d * H * 2 * W * 2 * (C / 4);
  unsigned grid = (unsigned)((total + 255) / 256);
  if (in_f32)
    k_upsample2x_linear_cl<false><<<grid, 256, 0, mp_stream(stream)>>>(in_f32, nullptr, nullptr, out_f32,
                                                                       (bf16*)out_hi, (bf16*)out_lo, N, D, H, W, C, up_d);
  else
    k_upsample2x_linear_cl<true><<<grid, 256, 0, mp_stream(stream)>>>(nullptr, (const bf16*)in_hi, (const bf16*)in_lo,
                                                                      out_f32, (bf16*)out_hi, (bf16*)out_lo, N, D, H, W,
                                                                      C, up_d);
  MP_LAUNCH_CHECK("mp_upsample2x_linear_cl");
  return 0;
}

// ---- backward of the x2 linear upsample (row f-2): the adjoint as a GATHER, so the result is deterministic.  Input voxel i of an
// axis is read by the outputs o with i0(o) == i (weight 1 - l) or i1(o) == i (weight l) under the forward's own `lin_src`; for a
// x2 resize that is at most 5 outputs per axis.  One thread = one input position x 4 channels.
__device__ __forceinline__ int lin_bwd_taps(int i, int in_size, int out_size, int (&oo)[8], float (&ww)[8]) {
  int lo = 0, hi = out_size - 1;
  if (in_size > 1 && out_size > 1 && in_size != out_size) {
    const float inv = (float)(out_size - 1) / (float)(in_size - 1);
    lo = max((int)floorf((float)(i - 1) * inv) - 1, 0);
    hi = min((int)ceilf((float)(i + 1) * inv) + 1, out_size - 1);
  } else if (in_size == out_size) {
    lo = hi = i;
  }
  int n = 0;
  for (int o = lo; o <= hi && n < 8; ++o) {
    int i0, i1;
    float l;
    lin_src(o, in_size, out_size, i0, i1, l);
    const float w = (i0 == i ? 1.f - l : 0.f) + (i1 == i ? l : 0.f);
    if (w != 0.f) { oo[n] = o; ww[n] = w; ++n; }
  }
  return n;
}

__global__ void k_upsample2x_linear_bwd_cl(const float* __restrict__ gout, float* __restrict__ gin, int N, int D, int H, int W,
                                           int C, int ud) {
  const int C4 = C >> 2;
  const int Do = D * ud, Ho = H * 2, Wo = W * 2;
  const int64_t total = (int64_t)N * D * H * W * C4;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int c4 = (int)(t % C4);
  int64_t p = t / C4;
  const int w = (int)(p % W); p /= W;
  const int h = (int)(p % H); p /= H;
  const int d = (int)(p % D);
  const int n = (int)(p / D);
  int od[8], oh[8], ow[8];
  float wd[8], wh[8], ww[8];
  const int nd = lin_bwd_taps(d, D, Do, od, wd), nh = lin_bwd_taps(h, H, Ho, oh, wh), nw = lin_bwd_taps(w, W, Wo, ow, ww);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int a = 0; a < nd; ++a)
    for (int b = 0; b < nh; ++b) {
      const float wab = wd[a] * wh[b];
      const int64_t row = (((int64_t)n * Do + od[a]) * Ho + oh[b]) * Wo;
      for (int c = 0; c < nw; ++c) {
        const float4 g = *reinterpret_cast<const float4*>(gout + (row + ow[c]) * C + c4 * 4);
        const float wgt = wab * ww[c];
        acc.x += wgt * g.x; acc.y += wgt * g.y; acc.z += wgt * g.z; acc.w += wgt * g.w;
      }
    }
  *reinterpret_cast<float4*>(gin + t * 4) = acc;
}

extern "C" int mp_upsample2x_linear_backward_cl(const float* grad_out, float* grad_in, int N, int D, int H, int W, int C, int up_d,
                                                void* stream) {
  MP_REQUIRE(grad_out && grad_in, "mp_upsample2x_linear_backward_cl: null pointer");
  MP_REQUIRE(N > 0 && D > 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0 && (up_d == 1 || up_d == 2),
             "mp_upsample2x_linear_backward_cl: bad dims");
  const int64_t total = (int64_t)N * D * H * W * (C / 4);
  k_upsample2x_linear_bwd_cl<<<(unsigned)((total + 255) / 256), 256, 0, mp_stream(stream)>>>(grad_out, grad_in, N, D, H, W, C, up_d);
  MP_LAUNCH_CHECK("mp_upsample2x_linear_backward_cl");
  return 0;
}

// nn.Upsample(x2, bilinear, align_corners=True) from split-bf16 planes to the F16_Q8 plane pair (G2d up-blocks)
extern "C" int mp_upsample2x_bilinear_hq(const void* in_hi, const void* in_lo, void* out_h16, void* out_q8, int N, int H, int W,
                                         int C, float q8_scale, void* stream) {
  MP_REQUIRE(in_hi && in_lo && out_h16 && out_q8, "mp_upsample2x_bilinear_hq: null pointer");
  MP_REQUIRE(N > 0 && N <= 65535 && H > 0 && H * 2 <= 65535 && W > 0 && C > 0 && C % 64 == 0 && q8_scale > 0.f,
             "mp_upsample2x_bilinear_hq: bad dims (channels in multiples of 64)");
  dim3 g4((unsigned)((W * (C / 8) + 255) / 256), (unsigned)H, (unsigned)N);
  k_upsample2x_bilinear_split8_2x2<true><<<g4, 256, 0, mp_stream(stream)>>>((const bf16*)in_hi, (const bf16*)in_lo, out_h16, out_q8,
                                                                               H, W, C, q8_scale);
  MP_LAUNCH_CHECK("mp_upsample2x_bilinear_hq");
  return 0;
}

// ------------------------------------------------------------------------------------------------ nearest
__global__ void k_upsample_nearest_cl(const float* __restrict__ in, float* __restrict__ out_f32,
                                      bf16* __restrict__ out_hi, bf16* __restrict__ out_lo, int N, int D, int H, int W,
                                      int C, int sd, int sh, int sw) {
  const int C4 = C >> 2;
  const int Do = D * sd, Ho = H * sh, Wo = W * sw;
  int64_t total = (int64_t)N * Do * Ho * Wo * C4;
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  int c4 = (int)(t % C4);
  int64_t p = t / C4;
  int wo = (int)(p % Wo); p /= Wo;
  int ho = (int)(p % Ho); p /= Ho;
  int d_o = (int)(p % Do);
  int n = (int)(p / Do);
  int64_t idx = ((((int64_t)n * D + d_o / sd) * H + ho / sh) * W + wo / sw) * C + c4 * 4;
  float4 v = *reinterpret_cast<const float4*>(in + idx);
  int64_t o = t * 4;
  if (out_f32) *reinterpret_cast<float4*>(out_f32 + o) = v;
  if (out_hi) mp_store_split4(out_hi, out_lo, o, v);
}

extern "C" int mp_upsample_nearest_cl(const float* in, float* out_f32, void* out_hi, void* out_lo, int N, int D,
                                      int H, int W, int C, int sd, int sh, int sw, void* stream) {
  MP_REQUIRE(in && (out_f32 || (out_hi && out_lo)), "mp_upsample_nearest_cl: null pointer");
  MP_REQUIRE(C % 4 == 0 && sd >= 1 && sh >= 1 && sw >= 1, "mp_upsample_nearest_cl: bad dims");
  int64_t total = (int64_t)N * D * sd * H * sh * W * sw * (C / 4);
  k_upsample_nearest_cl<<<(unsigned)((total + 255) / 256), 256, 0, mp_stream(stream)>>>(
      in, out_f32, (bf16*)out_hi, (bf16*)out_lo, N, D, H, W, C, sd, sh, sw);
  MP_LAUNCH_CHECK("mp_upsample_nearest_cl");
  return 0;
}

// ------------------------------------------------------------------------------------------------ image pyramid
__global__ void k_blur_subsample(const float* __restrict__ x, const float* __restrict__ kern, float* __restrict__ out,
                                 int NC, int H, int W, int ks, int step) {
  const int Ho = H / step, Wo = W / step;
  int64_t total = (int64_t)NC * Ho * Wo;
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  int wo = (int)(t % Wo);
  int ho = (int)((t / Wo) % Ho);
  int nc = (int)(t / ((int64_t)Wo * Ho));
  const int ka = ks / 2;
  const float* src = x + (int64_t)nc * H * W;
  float acc = 0.f;
  for (int i = 0; i < ks; ++i) {
    int y = ho * step + i - ka;
    if (y < 0 || y >= H) continue;
    for (int j = 0; j < ks; ++j) {
      int xx = wo * step + j - ka;
      if (xx < 0 || xx >= W) continue;
      acc += src[(int64_t)y * W + xx] * kern[i * ks + j];
    }
  }
  out[t] = acc;
}

extern "C" int mp_blur_subsample(const float* x, const float* kernel, float* out, int N, int C, int H, int W, int ks,
                                 int step, void* stream) {
  MP_REQUIRE(x && kernel && out, "mp_blur_subsample: null pointer");
  MP_REQUIRE(ks % 2 == 1 && step >= 1 && H % step == 0 && W % step == 0, "mp_blur_subsample: bad dims");
  int64_t total = (int64_t)N * C * (H / step) * (W / step);
  k_blur_subsample<<<(unsigned)((total + 255) / 256), 256, 0, mp_stream(stream)>>>(x, kernel, out, N * C, H, W, ks,
                                                                                   step);
  MP_LAUNCH_CHECK("mp_blur_subsample");
  return 0;
}

// ------------------------------------------------------------------------------------------------ motion-encoder pools
// MaxPool2d(3, stride 2, padding 1) on split CL tensors: one thread = one output pixel x 4 channels.
__global__ void __launch_bounds__(256)
k_maxpool3x3s2_cl(const bf16* __restrict__ in_hi, const bf16* __restrict__ in_lo, bf16* __restrict__ out_hi,
                  bf16* __restrict__ out_lo, int H, int W, int C) {
  // grid: x = chunks of (wo, c8) pairs of one output row, y = output row, z = sample; 8 channels (16 B) per thread
  const int C8 = C >> 3, Ho = H >> 1, Wo = W >> 1;
  const int idx = blockIdx.x * 256 + threadIdx.x;
  if (idx >= Wo * C8) return;
  const int wo = idx / C8, c8 = idx - wo * C8;
  const int ho = blockIdx.y, n = blockIdx.z;
  float m[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) m[i] = -INFINITY;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const int y = ho * 2 + a - 1;
    if (y < 0 || y >= H) continue;
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      const int x = wo * 2 + b - 1;
      if (x < 0 || x >= W) continue;
      const int64_t i = (((int64_t)n * H + y) * W + x) * C + c8 * 8;
      const bf16x8v h = *reinterpret_cast<const bf16x8v*>(in_hi + i);
      const bf16x8v l = *reinterpret_cast<const bf16x8v*>(in_lo + i);
#pragma unroll
      for (int k = 0; k < 8; ++k) m[k] = fmaxf(m[k], mp_join(h.v[k], l.v[k]));
    }
  }
  bf16x8v oh, ol;
#pragma unroll
  for (int k = 0; k < 8; ++k) mp_split2(m[k], oh.v[k], ol.v[k]);
  const int64_t o = ((((int64_t)n * Ho + ho) * Wo + wo) * C8 + c8) * 8;
  *reinterpret_cast<bf16x8v*>(out_hi + o) = oh;
  *reinterpret_cast<bf16x8v*>(out_lo + o) = ol;
}

extern "C" int mp_maxpool3x3s2_cl(const void* in_hi, const void* in_lo, void* out_hi, void* out_lo, int N, int H, int W,
                                  int C, void* stream) {
  MP_REQUIRE(in_hi && in_lo && out_hi && out_lo, "mp_maxpool3x3s2_cl: null pointer");
  MP_REQUIRE(N > 0 && N <= 65535 && C % 8 == 0 && H % 2 == 0 && W % 2 == 0 && H / 2 <= 65535,
             "mp_maxpool3x3s2_cl: bad dims");
  dim3 grid((unsigned)(((W / 2) * (C / 8) + 255) / 256), (unsigned)(H / 2), (unsigned)N);
  k_maxpool3x3s2_cl<<<grid, 256, 0, mp_stream(stream)>>>((const bf16*)in_hi, (const bf16*)in_lo, (bf16*)out_hi,
                                                         (bf16*)out_lo, H, W, C);
  MP_LAUNCH_CHECK("mp_maxpool3x3s2_cl");
  return 0;
}

// AdaptiveAvgPool2d(1): grid (C/32 column groups, N); block 32 x 8: 8 row-walkers per channel, smem reduce.
__global__ void k_global_avgpool_cl(const float* __restrict__ in_f32, const bf16* __restrict__ in_hi,
                                    const bf16* __restrict__ in_lo, float* __restrict__ out, int64_t S, int C) {
  __shared__ float red[8][33];
  const int n = blockIdx.y;
  const int c = blockIdx.x * 32 + threadIdx.x;
  float acc = 0.f;
  if (c < C) {
    for (int64_t s = threadIdx.y; s < S; s += 8) {
      int64_t o = ((int64_t)n * S + s) * C + c;
      acc += in_f32 ? in_f32[o] : mp_join(in_hi[o], in_lo[o]);
    }
  }
  red[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x];
    out[(int64_t)n * C + c] = t / (float)S;
  }
}

extern "C" int mp_global_avgpool_cl(const float* in_f32, const void* in_hi, const void* in_lo, float* out, int N,
                                    int64_t S, int C, void* stream) {
  MP_REQUIRE(out && (in_f32 || (in_hi && in_lo)), "mp_global_avgpool_cl: null pointer");
  MP_REQUIRE(N > 0 && N <= 65535 && S > 0 && C > 0, "mp_global_avgpool_cl: bad dims");
  dim3 grid((C + 31) / 32, N), block(32, 8);
  k_global_avgpool_cl<<<grid, block, 0, mp_stream(stream)>>>(in_f32, (const bf16*)in_hi, (const bf16*)in_lo, out, S, C);
  MP_LAUNCH_CHECK("mp_global_avgpool_cl");
  return 0;
}

// ------------------------------------------------------------------------------------------------ RGB stem input
// NCHW fp32 [N,C,S] (C <= 16, the RGB frames) -> split channels-last [N,S,16] with the channel axis zero-padded to 16:
// the operand of the tensor-core stems (one thread per position, coalesced plane reads, two 32-byte row writes).
struct __align__(16) bf16x8p {
  bf16 v[8];
};
__global__ void k_nchw_to_cl_pad16(const float* __restrict__ in, bf16* __restrict__ hi, bf16* __restrict__ lo, int C,
                                   int64_t S) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  const int n = blockIdx.y;
  bf16x8p h[2], l[2];
#pragma unroll
  for (int c = 0; c < 16; ++c) {
    const float v = c < C ? in[((int64_t)n * C + c) * S + s] : 0.f;
    mp_split2(v, h[c >> 3].v[c & 7], l[c >> 3].v[c & 7]);
  }
  const int64_t o = ((int64_t)n * S + s) * 16;
  *reinterpret_cast<bf16x8p*>(hi + o) = h[0];
  *reinterpret_cast<bf16x8p*>(hi + o + 8) = h[1];
  *reinterpret_cast<bf16x8p*>(lo + o) = l[0];
  *reinterpret_cast<bf16x8p*>(lo + o + 8) = l[1];
}

extern "C" int mp_nchw_to_cl_pad16(const float* in, void* out_hi, void* out_lo, int N, int C, int64_t S, void* stream) {
  MP_REQUIRE(in && out_hi && out_lo, "mp_nchw_to_cl_pad16: null pointer");
  MP_REQUIRE(N > 0 && N <= 65535 && C > 0 && C <= 16 && S > 0, "mp_nchw_to_cl_pad16: bad dims");
  dim3 grid((unsigned)((S + 255) / 256), N);
  k_nchw_to_cl_pad16<<<grid, 256, 0, mp_stream(stream)>>>(in, (bf16*)out_hi, (bf16*)out_lo, C, S);
  MP_LAUNCH_CHECK("mp_nchw_to_cl_pad16");
  return 0;
}
