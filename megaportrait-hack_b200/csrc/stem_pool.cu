// Fused RGB stem of the CIFAR-style ResNet-18 trunks of the motion encoder (resnet.py:192-197, 271-273):
//     nn.Conv2d(3, C, 3, 1, 1) + BatchNorm (folded) + ReLU + nn.MaxPool2d(3, 2, 1)
// NCHW fp32 frames [N,3,H,W] -> fp16 channels-last [N, H/2, W/2, C].
//
// Before this kernel the stage ran as im2col (patch tensor, 64 B / pixel) -> 1x1 tensor-core convolution (C fp16 channels at
// FULL resolution: 2.1 GB at batch 32, C = 128) -> max-pool reading it back: 1.8 ms of HBM traffic for 58 GFLOP.  Here the
// full-resolution tensor only ever exists in shared memory: a CTA owns an 8 x 8 tile of POOLED pixels of one frame, i.e. a
// 17 x 17 tile of convolution outputs (one row / column of halo above / left: pool windows are rows 2i-1 .. 2i+1), computed
// from a 19 x 19 x 3 input patch.
//
//   phase 0   input patch -> shared memory as fp16 (the same rounding as mp_im2col3x3_f16), zero outside the frame.  A CTA walks
//             up to four consecutive tiles of a tile row and requests the NEXT tile's patch (5 loads per thread, all in flight)
//             before it runs phases 1-3 of the current one: only the first tile of a CTA waits for HBM / L2.
//   phase 1   patch matrix P[304 x 32] fp16 (K = 27 taps x channels, zero-padded to 32; 289 rows used), 80-byte row pitch;
//             a thread gathers 8 consecutive k of one row (loop-invariant tap offsets) and stores 16 bytes
//   phase 2   tensor cores, legacy path (mma.sync m16n8k16, fp16 x fp16 -> fp32; SASS HMMA.16816.F32): K = 32 is far too short
//             to amortise a tcgen05 pipeline (TMEM round trip per 128 x 128 x 32 tile) and the accumulators are consumed by a
//             max over NEIGHBOURING rows, which wants them in shared memory anyway.  Same operand format as MP_PREC_F16X2:
//             activations fp16, weights fp16 hi + fp16 lo * 2048 (two accumulators, joined in fp32).  Warp w owns 32 output
//             channels (w & 3) and every other 16-row block (w >> 2); its weight fragments stay in registers for all tiles.
//             Epilogue: hi * s + (lo * s / 2048 + bias), then ReLU + saturation + fp16 pack in ONE F2FP instruction.
//             The halo row / column above / left of the frame is zeroed afterwards on border tiles (ReLU >= 0, so 0 is
//             neutral for the max and equals PyTorch's -inf padding whenever a window holds one real pixel -- always).
//   phase 3   3 x 3 / stride 2 max over the tile in shared memory, 16-byte coalesced channels-last stores.
//
// 107 KB of shared memory -> two CTAs per SM, so one CTA's tensor phase overlaps the other's shared-memory phases.
// Measured (B200, 32 x 512 x 512 frames, C = 128): 0.75 ms against 1.77 ms for the three kernels it replaces; ncu: HMMA pipe
// 39 %, shared-memory wavefronts 54 %, 16 warps per SM (latency-bound: `profiles/round2_stem_pool_full.txt`).
#include "common.cuh"

namespace mpb200 {

constexpr int SP_PT = 8;                     // pooled tile edge
constexpr int SP_CT = 2 * SP_PT + 1;         // convolution tile edge (17)
constexpr int SP_IT = SP_CT + 2;             // input tile edge (19)
constexpr int SP_IPITCH = 20;                // halves per input row
constexpr int SP_ROWS = SP_CT * SP_CT;       // 289 convolution pixels
constexpr int SP_MT = (SP_ROWS + 15) / 16;   // 19 blocks of 16 rows
constexpr int SP_PPITCH = 40;                // halves per patch-matrix row (80 B: conflict-free ldmatrix)
constexpr int SP_THREADS = 256;

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void hmma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// relu + saturate to the fp16 range + pack: one F2FP.SATFINITE.RELU.F16.F32.PACK_AB
__device__ __forceinline__ uint32_t relu_pack_f16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;\n" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

template <int C>
__global__ void __launch_bounds__(SP_THREADS, 2)
k_stem3x3_relu_maxpool_f16(const float* __restrict__ x, const __half* __restrict__ w_hi, const __half* __restrict__ w_lo,
                           const float* __restrict__ bias, float acc_scale, __half* __restrict__ out, int H, int W, int tiles_per_cta) {
  constexpr int CPITCH = C + 8;              // halves per convolution-tile row (+16 B: conflict-free fragment stores)
  constexpr int NG = C / 32;                 // channel groups of 32
  constexpr int MSPLIT = 8 / NG;             // warps that share a channel group split the row blocks
  extern __shared__ __align__(16) unsigned char smem[];
  __half* s_in = reinterpret_cast<__half*>(smem);                          // [3][19][20]
  __half* s_p = s_in + 3 * SP_IT * SP_IPITCH + 4;                          // [304][40]   (+4 halves: 16-byte aligned)
  __half* s_c = s_p + SP_MT * 16 * SP_PPITCH;                              // [304][C + 8]
  static_assert((3 * SP_IT * SP_IPITCH + 4) % 8 == 0, "patch matrix must start 16-byte aligned");

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = blockIdx.z, pi0 = blockIdx.y * SP_PT;
  const int cr0 = 2 * pi0 - 1;                            // frame row of convolution-tile row 0
  const float* xn = x + (size_t)n * 3 * H * W;

  // ---- weight fragments (registers, for all tiles of this CTA): B[k][n] = w[n][k], "col" operand, two consecutive k per
  //      register; acc_scale is folded into the epilogue constants: out = hi * s + (lo * s / 2048 + bias)
  const int g = lane >> 2, t = lane & 3;
  const int cg = warp % NG, ms = warp / NG;
  uint32_t bh[4][2][2], bl[4][2][2];
  float bs[4][2];
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    const int co = cg * 32 + nt * 8 + g;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      bh[nt][ks][0] = __ldg(reinterpret_cast<const uint32_t*>(w_hi + co * 32 + ks * 16 + 2 * t));
      bh[nt][ks][1] = __ldg(reinterpret_cast<const uint32_t*>(w_hi + co * 32 + ks * 16 + 2 * t + 8));
      bl[nt][ks][0] = __ldg(reinterpret_cast<const uint32_t*>(w_lo + co * 32 + ks * 16 + 2 * t));
      bl[nt][ks][1] = __ldg(reinterpret_cast<const uint32_t*>(w_lo + co * 32 + ks * 16 + 2 * t + 8));
    }
    bs[nt][0] = bias ? __ldg(bias + cg * 32 + nt * 8 + 2 * t) : 0.f;
    bs[nt][1] = bias ? __ldg(bias + cg * 32 + nt * 8 + 2 * t + 1) : 0.f;
  }
  const float s_lo = acc_scale * (1.0f / 2048.0f);

  // ---- phase 0 (loads): a thread's share of the 3 x 19 x 19 input patch, all loads in flight at once; the NEXT tile's patch
  //      is requested before this tile's phases 1-3 run, so only the first tile of a CTA waits for HBM / L2
  constexpr int IT0 = (3 * SP_IT * SP_IT + SP_THREADS - 1) / SP_THREADS;
  int goff[IT0], dc[IT0];        // frame offset of (channel, row, column 0) or -1 (row outside the frame); smem index | column << 16
#pragma unroll
  for (int it = 0; it < IT0; ++it) {
    const int i = tid + it * SP_THREADS;
    const int ch = i / (SP_IT * SP_IT), rem = i - ch * (SP_IT * SP_IT), r = rem / SP_IT, c = rem - r * SP_IT;
    const int fr = cr0 - 1 + r;
    const bool live = i < 3 * SP_IT * SP_IT;
    goff[it] = (live && fr >= 0 && fr < H) ? (ch * H + fr) * W : -1;
    dc[it] = live ? (((ch * SP_IT + r) * SP_IPITCH + c) | (c << 16)) : -1;
  }
  float v[IT0];
  auto fetch = [&](int pj0) {
#pragma unroll
    for (int it = 0; it < IT0; ++it) {
      const int fc = 2 * pj0 - 2 + (dc[it] >> 16);
      v[it] = 0.f;
      if (goff[it] >= 0 && fc >= 0 && fc < W) v[it] = __ldg(xn + goff[it] + fc);
    }
  };
  const int tile0 = blockIdx.x * tiles_per_cta;
  fetch(tile0 * SP_PT);

  // loop invariants of phase 1: a thread fills 8 consecutive k (one 16-byte store) of the rows (tid >> 2) + 64 j; the columns
  // k >= 27 meet zero weights, so they may hold any finite value (they re-read tap 0)
  int koff[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int k = (tid & 3) * 8 + e;
    const int tp = k / 3, ch = k - 3 * tp, kh = tp / 3, kw = tp - 3 * kh;
    koff[e] = k < 27 ? (ch * SP_IT + kh) * SP_IPITCH + kw : 0;
  }
  const uint32_t p_base = (uint32_t)__cvta_generic_to_shared(s_p) +
                          (uint32_t)(((lane & 7) + ((lane >> 3) & 1) * 8) * SP_PPITCH + (lane >> 4) * 8) * 2u;
  const int Ho = H >> 1, Wo = W >> 1;

  for (int tt = 0; tt < tiles_per_cta; ++tt) {
    const int pj0 = (tile0 + tt) * SP_PT;
    // ---- phase 0 (stores): fp16 with the rounding of mp_im2col3x3_f16, zero outside the frame
#pragma unroll
    for (int it = 0; it < IT0; ++it)
      if (dc[it] >= 0) s_in[dc[it] & 0xffff] = mp_to_f16(v[it]);
    __syncthreads();
    if (tt + 1 < tiles_per_cta) fetch(pj0 + SP_PT);

    // ---- phase 1: patch matrix; k = (kh * 3 + kw) * 3 + ch, as mp_im2col3x3_f16 / pack_stem3x3_f16 order it (rows 289..303
    //      are padding of the last 16-row block: they repeat row 288 and are never read by the pool)
#pragma unroll
    for (int j = 0; j < (SP_MT * 16 + 63) / 64; ++j) {
      const int p = (tid >> 2) + 64 * j;
      if (p < SP_MT * 16) {
        const int pc = min(p, SP_ROWS - 1), r = pc / SP_CT, c = pc - r * SP_CT;
        const __half* base = s_in + r * SP_IPITCH + c;
        uint32_t w4[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const uint32_t lo = *reinterpret_cast<const uint16_t*>(base + koff[2 * e]);
          const uint32_t hi = *reinterpret_cast<const uint16_t*>(base + koff[2 * e + 1]);
          w4[e] = lo | (hi << 16);
        }
        *reinterpret_cast<uint4*>(s_p + p * SP_PPITCH + (tid & 3) * 8) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
      }
    }
    __syncthreads();

    // ---- phase 2: P[304 x 32] x W^T[32 x C] on the tensor cores
    for (int mt = ms; mt < SP_MT; mt += MSPLIT) {
      uint32_t a0[4], a1[4];
      ldmatrix_x4(a0, p_base + (uint32_t)(mt * 16 * SP_PPITCH) * 2u);
      ldmatrix_x4(a1, p_base + (uint32_t)(mt * 16 * SP_PPITCH + 16) * 2u);
      __half* crow = s_c + (mt * 16 + g) * CPITCH + cg * 32 + 2 * t;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        float ch_[4] = {0.f, 0.f, 0.f, 0.f}, cl_[4] = {0.f, 0.f, 0.f, 0.f};
        hmma16816(ch_, a0, bh[nt][0][0], bh[nt][0][1]);
        hmma16816(cl_, a0, bl[nt][0][0], bl[nt][0][1]);
        hmma16816(ch_, a1, bh[nt][1][0], bh[nt][1][1]);
        hmma16816(cl_, a1, bl[nt][1][0], bl[nt][1][1]);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const float v0 = fmaf(ch_[2 * h], acc_scale, fmaf(cl_[2 * h], s_lo, bs[nt][0]));
          const float v1 = fmaf(ch_[2 * h + 1], acc_scale, fmaf(cl_[2 * h + 1], s_lo, bs[nt][1]));
          *reinterpret_cast<uint32_t*>(crow + h * 8 * CPITCH + nt * 8) = relu_pack_f16x2(v0, v1);
        }
      }
    }
    __syncthreads();
    // halo row / column above / left of the frame (tiles on the top / left border only): their "convolution outputs" were
    // computed from zero padding and must not take part in the max; 0 is neutral after ReLU (the rows 289..303 of the last
    // 16-row block are never read)
    if (pi0 == 0 || pj0 == 0) {
      if (pi0 == 0)
        for (int i = tid; i < SP_CT * (C / 2); i += SP_THREADS)
          *reinterpret_cast<uint32_t*>(s_c + (i / (C / 2)) * CPITCH + 2 * (i % (C / 2))) = 0u;
      if (pj0 == 0)
        for (int i = tid; i < SP_CT * (C / 2); i += SP_THREADS)
          *reinterpret_cast<uint32_t*>(s_c + (i / (C / 2)) * SP_CT * CPITCH + 2 * (i % (C / 2))) = 0u;
      __syncthreads();
    }

    // ---- phase 3: 3x3 stride-2 max; 8 channels (16 bytes) per thread
    constexpr int VEC = C / 8;                              // 16-byte vectors per pixel
    for (int i = tid; i < SP_PT * SP_PT * VEC; i += SP_THREADS) {
      const int vv = i % VEC, pp = i / VEC, pj = pp % SP_PT, pi = pp / SP_PT;
      const __half* src = s_c + ((2 * pi) * SP_CT + 2 * pj) * CPITCH + vv * 8;
      __half2 m[4];
      {
        const uint4 u = *reinterpret_cast<const uint4*>(src);
        m[0] = *reinterpret_cast<const __half2*>(&u.x); m[1] = *reinterpret_cast<const __half2*>(&u.y);
        m[2] = *reinterpret_cast<const __half2*>(&u.z); m[3] = *reinterpret_cast<const __half2*>(&u.w);
      }
#pragma unroll
      for (int d = 1; d < 9; ++d) {
        const int dr = d / 3, dc = d - 3 * dr;
        const uint4 u = *reinterpret_cast<const uint4*>(src + (dr * SP_CT + dc) * CPITCH);
        m[0] = __hmax2(m[0], *reinterpret_cast<const __half2*>(&u.x));
        m[1] = __hmax2(m[1], *reinterpret_cast<const __half2*>(&u.y));
        m[2] = __hmax2(m[2], *reinterpret_cast<const __half2*>(&u.z));
        m[3] = __hmax2(m[3], *reinterpret_cast<const __half2*>(&u.w));
      }
      uint4 o;
      o.x = *reinterpret_cast<const uint32_t*>(&m[0]); o.y = *reinterpret_cast<const uint32_t*>(&m[1]);
      o.z = *reinterpret_cast<const uint32_t*>(&m[2]); o.w = *reinterpret_cast<const uint32_t*>(&m[3]);
      *reinterpret_cast<uint4*>(out + (((size_t)n * Ho + pi0 + pi) * Wo + pj0 + pj) * C + vv * 8) = o;
    }
    // (no barrier here: the next iteration writes s_in, last read in phase 1; s_p and s_c are rewritten only after the next
    //  iteration's barriers)
  }
}

template <int C>
static int launch_stem(const float* x, const void* w_hi, const void* w_lo, const float* bias, float acc_scale, void* out,
                       int N, int H, int W, void* stream) {
  const size_t smem = (size_t)(3 * SP_IT * SP_IPITCH + 4 + SP_MT * 16 * SP_PPITCH + SP_MT * 16 * (C + 8)) * sizeof(__half);
  static bool attr_done[64] = {false};  // per device (the attribute is per device; idempotent, so a race only repeats the call)
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !attr_done[dev]) {
    cudaError_t e = cudaFuncSetAttribute(k_stem3x3_relu_maxpool_f16<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return mp_set_error("mp_stem3x3_relu_maxpool_f16: %s", cudaGetErrorString(e));
    if (dev >= 0 && dev < 64) attr_done[dev] = true;
  }
  const int tiles_x = W / (2 * SP_PT);
  const int per_cta = tiles_x % 4 == 0 ? 4 : tiles_x % 2 == 0 ? 2 : 1;     // consecutive tiles of one CTA (input prefetch)
  dim3 grid(tiles_x / per_cta, H / (2 * SP_PT), N);
  k_stem3x3_relu_maxpool_f16<C><<<grid, SP_THREADS, smem, mp_stream(stream)>>>(
      x, (const __half*)w_hi, (const __half*)w_lo, bias, acc_scale, (__half*)out, H, W, per_cta);
  MP_LAUNCH_CHECK("mp_stem3x3_relu_maxpool_f16");
  return 0;
}

}  // namespace mpb200

extern "C" int mp_stem3x3_relu_maxpool_f16(const float* x, const void* w_hi, const void* w_lo, const float* bias,
                                           float acc_scale, void* out_h, int N, int H, int W, int C, void* stream) {
  MP_REQUIRE(x && w_hi && w_lo && out_h, "mp_stem3x3_relu_maxpool_f16: null pointer");
  MP_REQUIRE(N > 0 && N <= 65535 && H > 0 && W > 0 && H % 16 == 0 && W % 16 == 0,
             "mp_stem3x3_relu_maxpool_f16: H and W must be multiples of 16 (8x8 pooled tiles), got %dx%dx%d", N, H, W);
  MP_REQUIRE(H / 16 <= 65535, "mp_stem3x3_relu_maxpool_f16: frame too tall");
  if (C == 128) return mpb200::launch_stem<128>(x, w_hi, w_lo, bias, acc_scale > 0.f ? acc_scale : 1.f, out_h, N, H, W, stream);
  if (C == 64) return mpb200::launch_stem<64>(x, w_hi, w_lo, bias, acc_scale > 0.f ? acc_scale : 1.f, out_h, N, H, W, stream);
  return mp_set_error("mp_stem3x3_relu_maxpool_f16: C must be 64 or 128, got %d", C);
}
