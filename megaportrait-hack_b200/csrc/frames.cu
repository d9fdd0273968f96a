// Frame I/O on the device: the pre- and post-processing that the reference's inference.py does on the
// host with torchvision / numpy (inference.py:16-20, 36-43), as two HBM-bound kernels, so that a serving loop moves
// uint8 frames (0.75 MB per 512x512 frame) over PCIe instead of fp32 tensors (3 MB).
#include "common.cuh"

namespace {

// uint8 HWC [N,H,W,3] -> fp32 NCHW [N,3,H,W]: transforms.ToTensor() (x / 255) then Normalize(mean, std) per channel.
__global__ void k_u8hwc_to_nchw(const uint8_t* __restrict__ in, float* __restrict__ out, int64_t HW, float mean, float inv_std) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= HW) return;
  const int n = blockIdx.y;
  const uint8_t* p = in + ((int64_t)n * HW + t) * 3;
  float* o = out + (int64_t)n * 3 * HW + t;
#pragma unroll
  for (int c = 0; c < 3; ++c) o[c * HW] = ((float)p[c] / 255.0f - mean) * inv_std;
}

// fp32 NCHW [N,3,H,W] -> uint8 HWC [N,H,W,3]: ((x + shift) * scale * 255) truncated like numpy's astype(np.uint8) on the
// in-range values (clamped to [0, 255] outside), optional channel reversal (cv2.cvtColor(.., COLOR_BGR2RGB)).
__global__ void k_nchw_to_u8hwc(const float* __restrict__ in, uint8_t* __restrict__ out, int64_t HW, float shift, float scale,
                                int reverse) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= HW) return;
  const int n = blockIdx.y;
  const float* p = in + (int64_t)n * 3 * HW + t;
  uint8_t* o = out + ((int64_t)n * HW + t) * 3;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float v = (p[c * HW] + shift) * scale * 255.0f;
    o[reverse ? 2 - c : c] = (uint8_t)fminf(fmaxf(v, 0.f), 255.f);
  }
}

}  // namespace

extern "C" int mp_frames_u8_to_f32(const void* in_u8, float* out, int N, int H, int W, float mean, float std, void* stream) {
  MP_REQUIRE(in_u8 && out, "mp_frames_u8_to_f32: null pointer");
  MP_REQUIRE(N > 0 && N <= 65535 && H > 0 && W > 0 && std != 0.f, "mp_frames_u8_to_f32: bad arguments");
  const int64_t HW = (int64_t)H * W;
  dim3 grid((unsigned)((HW + 255) / 256), (unsigned)N);
  k_u8hwc_to_nchw<<<grid, 256, 0, mp_stream(stream)>>>((const uint8_t*)in_u8, out, HW, mean, 1.0f / std);
  MP_LAUNCH_CHECK("mp_frames_u8_to_f32");
  return 0;
}

extern "C" int mp_frames_f32_to_u8(const float* in, void* out_u8, int N, int H, int W, float shift, float scale, int reverse_channels,
                                   void* stream) {
  MP_REQUIRE(in && out_u8, "mp_frames_f32_to_u8: null pointer");
  MP_REQUIRE(N > 0 && N <= 65535 && H > 0 && W > 0, "mp_frames_f32_to_u8: bad arguments");
  const int64_t HW = (int64_t)H * W;
  dim3 grid((unsigned)((HW + 255) / 256), (unsigned)N);
  k_nchw_to_u8hwc<<<grid, 256, 0, mp_stream(stream)>>>(in, (uint8_t*)out_u8, HW, shift, scale, reverse_channels);
  MP_LAUNCH_CHECK("mp_frames_f32_to_u8");
  return 0;
}
