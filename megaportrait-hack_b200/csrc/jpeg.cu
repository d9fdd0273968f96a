// JPEG frames -> uint8 HWC on the device through nvJPEG (reference inference.py:10-13 `Image.open(...).convert("RGB")` and
// the per-frame decode of EmoDataset.py:180-247), so that a compressed frame crosses PCIe instead of 3 bytes per pixel
// and the decode result feeds mp_frames_u8_to_f32 without touching the host again.
//
// nvJPEG is a CUDA-toolkit library (like cuBLAS); it is loaded with dlopen at first use so that libmpb200.so has no
// link-time dependency on it -- if it is missing the two entry points return an error and nothing else is affected.
#include <dlfcn.h>
#include <nvjpeg.h>

#include <mutex>

#include "common.cuh"

namespace {

struct Api {
  void* lib = nullptr;
  nvjpegStatus_t (*CreateSimple)(nvjpegHandle_t*) = nullptr;
  nvjpegStatus_t (*JpegStateCreate)(nvjpegHandle_t, nvjpegJpegState_t*) = nullptr;
  nvjpegStatus_t (*GetImageInfo)(nvjpegHandle_t, const unsigned char*, size_t, int*, nvjpegChromaSubsampling_t*, int*, int*) = nullptr;
  nvjpegStatus_t (*Decode)(nvjpegHandle_t, nvjpegJpegState_t, const unsigned char*, size_t, nvjpegOutputFormat_t, nvjpegImage_t*,
                           cudaStream_t) = nullptr;
  nvjpegHandle_t handle = nullptr;
  nvjpegJpegState_t state[64] = {nullptr};   // one decoder state per device
  bool ok = false;
};

Api& api() {
  static Api a;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {"libnvjpeg.so.12", "/usr/local/cuda/lib64/libnvjpeg.so.12", "libnvjpeg.so"};
    for (const char* n : names) {
      a.lib = dlopen(n, RTLD_NOW | RTLD_LOCAL);
      if (a.lib) break;
    }
    if (!a.lib) return;
    a.CreateSimple = reinterpret_cast<decltype(a.CreateSimple)>(dlsym(a.lib, "nvjpegCreateSimple"));
    a.JpegStateCreate = reinterpret_cast<decltype(a.JpegStateCreate)>(dlsym(a.lib, "nvjpegJpegStateCreate"));
    a.GetImageInfo = reinterpret_cast<decltype(a.GetImageInfo)>(dlsym(a.lib, "nvjpegGetImageInfo"));
    a.Decode = reinterpret_cast<decltype(a.Decode)>(dlsym(a.lib, "nvjpegDecode"));
    if (!a.CreateSimple || !a.JpegStateCreate || !a.GetImageInfo || !a.Decode) return;
    if (a.CreateSimple(&a.handle) != NVJPEG_STATUS_SUCCESS) return;
    a.ok = true;
  });
  return a;
}

std::mutex g_jpeg_mu;

}  // namespace

// Width / height of a JPEG bitstream in HOST memory (no decode).
extern "C" int mp_jpeg_info(const unsigned char* jpeg_host, size_t nbytes, int* width, int* height) {
  MP_REQUIRE(jpeg_host && nbytes > 0 && width && height, "mp_jpeg_info: null pointer");
  Api& a = api();
  MP_REQUIRE(a.ok, "mp_jpeg_info: nvJPEG (libnvjpeg.so.12) is not available");
  int comps = 0, w[NVJPEG_MAX_COMPONENT] = {0}, h[NVJPEG_MAX_COMPONENT] = {0};
  nvjpegChromaSubsampling_t ss;
  std::lock_guard<std::mutex> lock(g_jpeg_mu);
  nvjpegStatus_t st = a.GetImageInfo(a.handle, jpeg_host, nbytes, &comps, &ss, w, h);
  MP_REQUIRE(st == NVJPEG_STATUS_SUCCESS, "mp_jpeg_info: not a decodable JPEG (nvjpeg status %d)", (int)st);
  *width = w[0];
  *height = h[0];
  return 0;
}

// Decodes `n` JPEG bitstreams (HOST pointers `jpeg_host[i]`, `nbytes[i]`), all H x W, into out_u8 [n, H, W, 3] (DEVICE,
// RGB interleaved: the layout mp_frames_u8_to_f32 reads).  Stream-ordered on `stream`; the bitstreams must stay valid
// until the stream has passed this call.
extern "C" int mp_decode_jpeg_frames(const unsigned char* const* jpeg_host, const size_t* nbytes, int n, unsigned char* out_u8,
                                     int H, int W, void* stream) {
  MP_REQUIRE(jpeg_host && nbytes && out_u8 && n > 0 && H > 0 && W > 0, "mp_decode_jpeg_frames: bad arguments");
  Api& a = api();
  MP_REQUIRE(a.ok, "mp_decode_jpeg_frames: nvJPEG (libnvjpeg.so.12) is not available");
  int dev = 0;
  cudaGetDevice(&dev);
  MP_REQUIRE(dev >= 0 && dev < 64, "mp_decode_jpeg_frames: bad device");
  std::lock_guard<std::mutex> lock(g_jpeg_mu);
  if (!a.state[dev]) {
    nvjpegStatus_t st = a.JpegStateCreate(a.handle, &a.state[dev]);
    MP_REQUIRE(st == NVJPEG_STATUS_SUCCESS, "mp_decode_jpeg_frames: nvjpegJpegStateCreate failed (%d)", (int)st);
  }
  for (int i = 0; i < n; ++i) {
    int comps = 0, w[NVJPEG_MAX_COMPONENT] = {0}, h[NVJPEG_MAX_COMPONENT] = {0};
    nvjpegChromaSubsampling_t ss;
    nvjpegStatus_t st = a.GetImageInfo(a.handle, jpeg_host[i], nbytes[i], &comps, &ss, w, h);
    MP_REQUIRE(st == NVJPEG_STATUS_SUCCESS, "mp_decode_jpeg_frames: frame %d is not a decodable JPEG (%d)", i, (int)st);
    MP_REQUIRE(w[0] == W && h[0] == H, "mp_decode_jpeg_frames: frame %d is %d x %d, expected %d x %d", i, w[0], h[0], W, H);
    nvjpegImage_t img;
    for (int c = 0; c < NVJPEG_MAX_COMPONENT; ++c) { img.channel[c] = nullptr; img.pitch[c] = 0; }
    img.channel[0] = out_u8 + (size_t)i * H * W * 3;
    img.pitch[0] = (size_t)W * 3;
    st = a.Decode(a.handle, a.state[dev], jpeg_host[i], nbytes[i], NVJPEG_OUTPUT_RGBI, &img, mp_stream(stream));
    MP_REQUIRE(st == NVJPEG_STATUS_SUCCESS, "mp_decode_jpeg_frames: nvjpegDecode failed on frame %d (%d)", i, (int)st);
  }
  return 0;
}
