// Row f-2 (training step), second slice: the gradients that the forward operators of the path still lacked.
//
//   mp_conv_wgrad              dL/dW of a stride-1 "same" convolution (nn.Conv2d / nn.Conv3d, model.py:61-86, 439-471):
//                              dW[co][tap][ci] = sum_p dY[p][co] * X[p + off(tap)][ci]   (zero outside the volume)
//   mp_bias_grad               dL/db = column sums of dY
//   mp_group_norm_backward     nn.GroupNorm backward (dX, dgamma, dbeta) from the forward's (sum, sum of squares) statistics
//
// mp_conv_wgrad: a GEMM per filter tap with K = positions (the long axis), M = Cout, N = Cin.  Both operands arrive
// channels-last, i.e. K-outer / M-(N-)contiguous -- "transposed" for the tensor core -- so the fragments are fetched with
// ldmatrix.trans from [position][channel] tiles; the products run as three bf16 passes (hi*hi + hi*lo + lo*hi, the split of
// each fp32 value made while it is staged) with fp32 accumulation: fp32-grade gradients, like the forward and the data
// gradient (ops.conv_input_grad on the tcgen05 kernel).  CTA = 64 x 64 tile of one tap over a slice of the positions; the
// slices are summed with fp32 RED into dW (split-K).  This slice uses the legacy mma.sync path: the operand it needs (MN-major
// A and B from channels-last tensors with a per-tap position shift and border mask) has no TMA box, so the tiles are built by
// the threads anyway.  (Superseded on supported shapes by mp_conv_wgrad_tc in conv_tc.cu: the same GEMM on tcgen05 with
// MN-major UMMA descriptors over TMA boxes; this kernel stays as the fallback and as its cross-check.)
//
//   mp_pack_conv_weights       OIDHW fp32 weights -> the split-bf16 K-major operand planes of mp_conv_tc in ONE launch, for the
//                              forward convolution or (transposed + tap-flipped) for its data gradient: the training step repacks
//                              every weight twice per iteration, which as ATen expressions cost ~10 launches each
#include "common.cuh"

namespace mpb200 {

constexpr int WG_TM = 64, WG_TN = 64, WG_TK = 32;       // Cout tile, Cin tile, positions per stage
constexpr int WG_PITCH = WG_TM + 8;                     // bf16 per smem row (144 B: conflict-free ldmatrix)
constexpr int WG_THREADS = 128;

__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void split4(const float4 v, uint2& hi, uint2& lo) {
  const __nv_bfloat162 h0 = __floats2bfloat162_rn(v.x, v.y), h1 = __floats2bfloat162_rn(v.z, v.w);
  hi.x = *reinterpret_cast<const uint32_t*>(&h0);
  hi.y = *reinterpret_cast<const uint32_t*>(&h1);
  const __nv_bfloat162 l0 = __floats2bfloat162_rn(v.x - __uint_as_float(hi.x << 16), v.y - __uint_as_float(hi.x & 0xffff0000u));
  const __nv_bfloat162 l1 = __floats2bfloat162_rn(v.z - __uint_as_float(hi.y << 16), v.w - __uint_as_float(hi.y & 0xffff0000u));
  lo.x = *reinterpret_cast<const uint32_t*>(&l0);
  lo.y = *reinterpret_cast<const uint32_t*>(&l1);
}

struct WgradParams {
  const float* x;      // [N, D, H, W, Cin]
  const float* dy;     // [N, D, H, W, Cout]
  float* dw;           // [Cout, KD*KH*KW, Cin], zeroed by the caller
  int N, D, H, W, Cin, Cout, KD, KH, KW;
  int tiles_n;         // Cin tiles
  int64_t P, chunk;    // positions, positions per K slice (multiple of WG_TK)
};

__global__ void __launch_bounds__(WG_THREADS)
k_conv_wgrad(const WgradParams p) {
  __shared__ __align__(16) bf16 sA[2][2][WG_TK][WG_PITCH];      // [stage][hi, lo][position][co]
  __shared__ __align__(16) bf16 sB[2][2][WG_TK][WG_PITCH];      // [stage][hi, lo][position][ci]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int co0 = (blockIdx.x / p.tiles_n) * WG_TM, ci0 = (blockIdx.x % p.tiles_n) * WG_TN;
  const int tap = blockIdx.y;
  const int kw = tap % p.KW, kh = (tap / p.KW) % p.KH, kd = tap / (p.KW * p.KH);
  const int dz = kd - p.KD / 2, dyo = kh - p.KH / 2, dx = kw - p.KW / 2;
  const int64_t k_begin = (int64_t)blockIdx.z * p.chunk;
  const int64_t k_end = min(k_begin + p.chunk, p.P);
  if (k_begin >= k_end) return;

  // a thread stages 4 float4 of A (dY) and 4 of B (X) per 32-position stage: item f = tid + 128 j -> row f / 16, column 4 (f % 16)
  const int col4 = (tid & 15) * 4, row_base = tid >> 4;          // rows row_base + 8 j
  // position -> (z, y, x) inside its sample with 32-bit arithmetic: `vbase` = (first position of the stage) mod D*H*W is carried
  // from stage to stage, a row adds its offset (< 32) and wraps once (volumes of fewer than 32 positions take the 64-bit path)
  const uint32_t HW = (uint32_t)p.H * p.W, DHW = (uint32_t)p.D * HW;
  const bool small = DHW < (uint32_t)WG_TK;
  uint32_t vbase = (uint32_t)(k_begin % DHW);
  float4 ra[4], rb[4];
  auto fetch = [&](int64_t k0) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = row_base + 8 * j;
      const int64_t pos = k0 + r;
      ra[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      rb[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (pos < k_end) {
        if (co0 + col4 < p.Cout) ra[j] = __ldg(reinterpret_cast<const float4*>(p.dy + pos * p.Cout + co0 + col4));
        // the tap reads x at (z + dz, y + dy, x + dx): zero outside the volume
        uint32_t q = small ? (uint32_t)(pos % DHW) : vbase + (uint32_t)r;
        if (!small && q >= DHW) q -= DHW;
        const uint32_t zz = q / HW, rem = q - zz * HW, yy = rem / (uint32_t)p.W, xx = rem - yy * (uint32_t)p.W;
        const int zs = (int)zz + dz, ys = (int)yy + dyo, xs = (int)xx + dx;
        if (zs >= 0 && zs < p.D && ys >= 0 && ys < p.H && xs >= 0 && xs < p.W && ci0 + col4 < p.Cin) {
          const int64_t src = pos + ((int64_t)dz * p.H + dyo) * p.W + dx;
          rb[j] = __ldg(reinterpret_cast<const float4*>(p.x + src * p.Cin + ci0 + col4));
        }
      }
    }
    if (!small) {                        // advance to the next stage
      vbase += WG_TK;
      if (vbase >= DHW) vbase -= DHW;
    }
  };
  auto stage = [&](int s) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = row_base + 8 * j;
      uint2 hi, lo;
      split4(ra[j], hi, lo);
      *reinterpret_cast<uint2*>(&sA[s][0][r][col4]) = hi;
      *reinterpret_cast<uint2*>(&sA[s][1][r][col4]) = lo;
      split4(rb[j], hi, lo);
      *reinterpret_cast<uint2*>(&sB[s][0][r][col4]) = hi;
      *reinterpret_cast<uint2*>(&sB[s][1][r][col4]) = lo;
    }
  };

  // warp tile: 32 (co) x 32 (ci) = 2 x 4 m16n8 accumulators
  const int wm = (warp >> 1) * 32, wn = (warp & 1) * 32;
  float acc[2][4][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[i][j][e] = 0.f;
  // ldmatrix.trans lane addresses (see the header): A^T tile [k][m], B tile [k][n]
  const int a_row = (lane & 7) + ((lane >> 4) & 1) * 8, a_col = ((lane >> 3) & 1) * 8;
  const int b_row = (lane & 7) + ((lane >> 3) & 1) * 8, b_col = ((lane >> 4) & 1) * 8;

  fetch(k_begin);
  int s = 0;
  for (int64_t k0 = k_begin; k0 < k_end; k0 += WG_TK, s ^= 1) {
    stage(s);
    __syncthreads();
    if (k0 + WG_TK < k_end) fetch(k0 + WG_TK);          // next stage's global loads fly under this stage's MMAs
#pragma unroll
    for (int ks = 0; ks < WG_TK; ks += 16) {
      uint32_t ah[2][4], al[2][4], bh[2][4], bl[2][4];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        ldsm_x4_t(ah[i], (uint32_t)__cvta_generic_to_shared(&sA[s][0][ks + a_row][wm + 16 * i + a_col]));
        ldsm_x4_t(al[i], (uint32_t)__cvta_generic_to_shared(&sA[s][1][ks + a_row][wm + 16 * i + a_col]));
        ldsm_x4_t(bh[i], (uint32_t)__cvta_generic_to_shared(&sB[s][0][ks + b_row][wn + 16 * i + b_col]));
        ldsm_x4_t(bl[i], (uint32_t)__cvta_generic_to_shared(&sB[s][1][ks + b_row][wn + 16 * i + b_col]));
      }
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t bh0 = bh[j >> 1][(j & 1) * 2], bh1 = bh[j >> 1][(j & 1) * 2 + 1];
          const uint32_t bl0 = bl[j >> 1][(j & 1) * 2], bl1 = bl[j >> 1][(j & 1) * 2 + 1];
          mma_bf16(acc[i][j], al[i], bh0, bh1);           // small terms first
          mma_bf16(acc[i][j], ah[i], bl0, bl1);
          mma_bf16(acc[i][j], ah[i], bh0, bh1);
        }
    }
    // (two stages: the next iteration writes the other buffer; the barrier at its top orders this stage's reads before the
    //  iteration after next overwrites it)
  }
  // split-K reduction: fp32 RED into dW[co][tap][ci]
  const int g = lane >> 2, t = lane & 3;
  const int T = p.KD * p.KH * p.KW;
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int co = co0 + wm + 16 * i + g + (e >> 1) * 8, ci = ci0 + wn + 8 * j + 2 * t + (e & 1);
        if (co < p.Cout && ci < p.Cin) atomicAdd(p.dw + ((int64_t)co * T + tap) * p.Cin + ci, acc[i][j][e]);
      }
}

// column sums of dY [P, C] -> db [C] (fp32 RED; db zeroed by the caller).  Block = R rows x CV channel-vectors (16-byte loads);
// a thread owns VEC fixed channels and walks rows; the R row-lanes meet in shared memory before one RED per channel and block.
template <int VEC>
__global__ void __launch_bounds__(256)
k_bias_grad(const float* __restrict__ dy, float* __restrict__ db, int64_t P, int C, int Cc, int64_t rows_per_block) {
  // grid.y walks chunks of Cc <= 1024 channels (c0 = blockIdx.y * Cc) of the C-wide rows
  __shared__ float sh[1024];
  const int c0 = blockIdx.y * Cc;
  const int cw = min(Cc, C - c0);
  const int CV = Cc / VEC, R = blockDim.x / CV;
  const int r = threadIdx.x / CV, cv = threadIdx.x % CV;
  for (int i = threadIdx.x; i < Cc; i += blockDim.x) sh[i] = 0.f;
  __syncthreads();
  if (r < R && cv * VEC < cw) {
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_block, r1 = min(r0 + rows_per_block, P);
    float s[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) s[k] = 0.f;
    for (int64_t row = r0 + r; row < r1; row += R) {
      if (VEC == 4) {
        const float4 v = *reinterpret_cast<const float4*>(dy + row * C + c0 + cv * 4);
        s[0] += v.x; s[1 % VEC] += v.y; s[2 % VEC] += v.z; s[3 % VEC] += v.w;
      } else {
        s[0] += dy[row * C + c0 + cv];
      }
    }
#pragma unroll
    for (int k = 0; k < VEC; ++k) atomicAdd(&sh[cv * VEC + k], s[k]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < cw; i += blockDim.x) atomicAdd(db + c0 + i, sh[i]);
}

// ---- GroupNorm backward.  y = (x - mean_g) * rstd_g * gamma_c + beta_c over the group's S * C/G elements of one sample.
//   dgamma_c = sum_{n,s} dy * xhat      dbeta_c = sum_{n,s} dy
//   dx = rstd * (dy*gamma - mean_g(dy*gamma) - xhat * mean_g(dy*gamma*xhat))
// pass 1: per (sample, group) the two group sums (double), per channel dgamma / dbeta (double RED); pass 2: elementwise.
// grid (row blocks, N); block = R rows x CV channel-vectors.  A thread owns VEC fixed channels (their group statistics live in
// registers) and walks rows with 16-byte loads; fp32 partial sums over <= 16 rows, promoted to double where the R row-lanes meet
// in shared memory; then one double RED per channel (dgamma, dbeta) and per group (the two group sums) and block.
template <int VEC>
__global__ void __launch_bounds__(256)
k_gn_bwd_reduce(const float* __restrict__ x, const float* __restrict__ dy, const double* __restrict__ stats,
                const float* __restrict__ gamma, double* __restrict__ gsum, double* __restrict__ dgamma,
                double* __restrict__ dbeta, int64_t S, int C, int G, float eps, int64_t rows_per_block) {
  extern __shared__ double shd[];          // [C][2] per-channel sums, then [G][2] group sums
  const int n = blockIdx.y;
  const int cpg = C / G;
  const double cnt = (double)S * cpg;
  const int CV = C / VEC, R = blockDim.x / CV;
  const int r = threadIdx.x / CV, cv = threadIdx.x % CV;
  for (int i = threadIdx.x; i < 2 * C + 2 * G; i += blockDim.x) shd[i] = 0.0;
  __syncthreads();
  if (r < R) {
    float mu[VEC], rstd[VEC], s_dy[VEC], s_dyx[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      const int g = (cv * VEC + k) / cpg;
      const double mean = stats[((int64_t)n * G + g) * 2] / cnt;
      const double var = fmax(stats[((int64_t)n * G + g) * 2 + 1] / cnt - mean * mean, 0.0);
      rstd[k] = (float)(1.0 / sqrt(var + (double)eps));
      mu[k] = (float)mean;
      s_dy[k] = s_dyx[k] = 0.f;
    }
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_block, r1 = min(r0 + rows_per_block, S);
#pragma unroll 4
    for (int64_t row = r0 + r; row < r1; row += R) {
      const int64_t o = ((int64_t)n * S + row) * C + cv * VEC;
      if (VEC == 4) {
        const float4 xv = *reinterpret_cast<const float4*>(x + o), dv = *reinterpret_cast<const float4*>(dy + o);
        s_dy[0] += dv.x; s_dyx[0] += dv.x * ((xv.x - mu[0]) * rstd[0]);
        s_dy[1 % VEC] += dv.y; s_dyx[1 % VEC] += dv.y * ((xv.y - mu[1 % VEC]) * rstd[1 % VEC]);
        s_dy[2 % VEC] += dv.z; s_dyx[2 % VEC] += dv.z * ((xv.z - mu[2 % VEC]) * rstd[2 % VEC]);
        s_dy[3 % VEC] += dv.w; s_dyx[3 % VEC] += dv.w * ((xv.w - mu[3 % VEC]) * rstd[3 % VEC]);
      } else {
        const float d = dy[o];
        s_dy[0] += d; s_dyx[0] += d * ((x[o] - mu[0]) * rstd[0]);
      }
    }
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      atomicAdd(&shd[2 * (cv * VEC + k)], (double)s_dy[k]);
      atomicAdd(&shd[2 * (cv * VEC + k) + 1], (double)s_dyx[k]);
    }
  }
  __syncthreads();
  double* shg = shd + 2 * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const double a = shd[2 * c], b = shd[2 * c + 1], gm = gamma ? (double)gamma[c] : 1.0;
    atomicAdd(dbeta + c, a);
    atomicAdd(dgamma + c, b);
    atomicAdd(&shg[2 * (c / cpg)], a * gm);
    atomicAdd(&shg[2 * (c / cpg) + 1], b * gm);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * G; i += blockDim.x) atomicAdd(gsum + (int64_t)n * G * 2 + i, shg[i]);
}

// dx = rstd * (dy * gamma - m1 - xhat * m2) is affine in (dy, x) per (sample, channel): dx = a * dy + b * x + c with
//   a = rstd * gamma,  b = -rstd^2 * m2,  c = rstd * (mu * rstd * m2 - m1).
// The double-precision part runs once per (sample, channel) here; the elementwise pass below is three FMAs on 16-byte vectors.
__global__ void k_gn_bwd_coef(const double* __restrict__ stats, const float* __restrict__ gamma, const double* __restrict__ gsum,
                              float* __restrict__ coef, int64_t S, int C, int G, float eps) {
  const int n = blockIdx.x;
  const int cpg = C / G;
  const double cnt = (double)S * cpg;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = c / cpg;
    const double mean = stats[((int64_t)n * G + g) * 2] / cnt;
    const double var = fmax(stats[((int64_t)n * G + g) * 2 + 1] / cnt - mean * mean, 0.0);
    const double rstd = 1.0 / sqrt(var + (double)eps);
    const double m1 = gsum[((int64_t)n * G + g) * 2] / cnt, m2 = gsum[((int64_t)n * G + g) * 2 + 1] / cnt;
    const double gm = gamma ? (double)gamma[c] : 1.0;
    float* o = coef + ((int64_t)n * 3) * C + c;
    o[0] = (float)(rstd * gm);
    o[C] = (float)(-rstd * rstd * m2);
    o[2 * C] = (float)(rstd * (mean * rstd * m2 - m1));
  }
}

template <int VEC>
__global__ void __launch_bounds__(256)
k_gn_bwd_apply(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ coef, float* __restrict__ dx,
               int64_t S, int C, int64_t total_vec) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total_vec) return;
  const int CV = C / VEC;
  const int cv = (int)(t % CV);
  const int64_t n = (t / CV) / S;
  const float* k = coef + n * 3 * C + cv * VEC;
  const int64_t o = t * VEC;
  if (VEC == 4) {
    const float4 a = *reinterpret_cast<const float4*>(k), b = *reinterpret_cast<const float4*>(k + C),
                 c = *reinterpret_cast<const float4*>(k + 2 * C);
    const float4 xv = *reinterpret_cast<const float4*>(x + o), dv = *reinterpret_cast<const float4*>(dy + o);
    float4 r;
    r.x = fmaf(a.x, dv.x, fmaf(b.x, xv.x, c.x));
    r.y = fmaf(a.y, dv.y, fmaf(b.y, xv.y, c.y));
    r.z = fmaf(a.z, dv.z, fmaf(b.z, xv.z, c.z));
    r.w = fmaf(a.w, dv.w, fmaf(b.w, xv.w, c.w));
    *reinterpret_cast<float4*>(dx + o) = r;
  } else {
    dx[o] = fmaf(k[0], dy[o], fmaf(k[C], x[o], k[2 * C]));
  }
}

}  // namespace mpb200

namespace mpb200 {
// out planes [rows_pad][K] bf16 (hi, lo).  dgrad = 0: rows = Cout, K index = tap * Cin + ci, value w[co][ci][tap].
// dgrad = 1: rows = Cin, K index = tap' * Cout + co, value w[co][ci][T - 1 - tap'] (all spatial axes flipped).
__global__ void k_pack_conv_weights(const float* __restrict__ w, bf16* __restrict__ hi, bf16* __restrict__ lo, int Cout, int Cin,
                                    int T, int rows, int rows_pad, int dgrad) {
  const int64_t K = (int64_t)T * (dgrad ? Cout : Cin);
  const int64_t total = (int64_t)rows_pad * K;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / K);
    const int64_t kk = i - (int64_t)r * K;
    float v = 0.f;
    if (r < rows) {
      if (!dgrad) {
        const int tap = (int)(kk / Cin), ci = (int)(kk % Cin);
        v = w[((int64_t)r * Cin + ci) * T + tap];
      } else {
        const int tap = (int)(kk / Cout), co = (int)(kk % Cout);
        v = w[((int64_t)co * Cin + r) * T + (T - 1 - tap)];
      }
    }
    mp_split2(v, hi[i], lo[i]);
  }
}
}  // namespace mpb200

namespace mpb200 {
// RGB patches for the weight gradient of a stem convolution (row f-2): x NCHW fp32 [N, C <= 4, H, W] -> split-bf16 rows
// [N, Ho, Wo, Kpad], element (kh * KW + kw) * C + c = x[n, c, ho * stride + kh - KH/2, wo * stride + kw - KW/2] (zero outside the
// image and for the padding up to Kpad).  dW of the stem is then ONE K = positions GEMM with Kpad "input channels"
// (mp_conv_wgrad_tc with a 1x1 filter) instead of KH*KW tap GEMMs over 16 zero-padded channels.  One thread = 8 patch elements.
struct __align__(16) bf16x8 {
  bf16 v[8];
};
__global__ void __launch_bounds__(256)
k_im2col_rgb_split(const float* __restrict__ x, bf16* __restrict__ hi, bf16* __restrict__ lo, int C, int H, int W, int KH, int KW,
                   int stride, int Kpad, int64_t total_chunks) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total_chunks) return;
  const int Ho = H / stride, Wo = W / stride, chunks = Kpad >> 3;
  const int j = (int)(t % chunks);
  int64_t p = t / chunks;
  const int wo = (int)(p % Wo); p /= Wo;
  const int ho = (int)(p % Ho);
  const int n = (int)(p / Ho);
  const int Kreal = KH * KW * C;
  const float* base = x + (int64_t)n * C * H * W;
  bf16x8 h, l;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int k = j * 8 + e;
    float v = 0.f;
    if (k < Kreal) {
      const int tap = k / C, c = k - tap * C;
      const int kh = tap / KW, kw = tap - kh * KW;
      const int y = ho * stride + kh - KH / 2, xx = wo * stride + kw - KW / 2;
      if (y >= 0 && y < H && xx >= 0 && xx < W) v = __ldg(base + ((int64_t)c * H + y) * W + xx);
    }
    mp_split2(v, h.v[e], l.v[e]);
  }
  *reinterpret_cast<bf16x8*>(hi + t * 8) = h;
  *reinterpret_cast<bf16x8*>(lo + t * 8) = l;
}
}  // namespace mpb200

extern "C" int mp_im2col_rgb_split(const float* x, void* out_hi, void* out_lo, int N, int C, int H, int W, int KH, int KW, int stride,
                                   int Kpad, void* stream) {
  MP_REQUIRE(x && out_hi && out_lo, "mp_im2col_rgb_split: null pointer");
  MP_REQUIRE(N > 0 && C > 0 && C <= 4 && H > 0 && W > 0 && KH % 2 == 1 && KW % 2 == 1 && (stride == 1 || stride == 2) &&
             H % stride == 0 && W % stride == 0 && Kpad % 8 == 0 && Kpad >= KH * KW * C,
             "mp_im2col_rgb_split: bad arguments");
  const int64_t total = (int64_t)N * (H / stride) * (W / stride) * (Kpad / 8);
  mpb200::k_im2col_rgb_split<<<(unsigned)((total + 255) / 256), 256, 0, mp_stream(stream)>>>(x, (bf16*)out_hi, (bf16*)out_lo, C, H, W,
                                                                                             KH, KW, stride, Kpad, total);
  MP_LAUNCH_CHECK("mp_im2col_rgb_split");
  return 0;
}

namespace mpb200 {
// ---- nn.MaxPool2d(3, stride 2, padding 1) for the training path (row f-2; the stems of the three ResNets), channels-last fp32.
// Forward: the maximum and the window position (kh * 3 + kw, 0..8) of the FIRST maximum in ATen's scan order (kh outer, kw inner,
// strict >; a NaN wins, as in ATen).  Backward: a gather -- an input pixel lies in at most 2 x 2 windows; it receives a window's
// gradient iff the stored position is its own (deterministic, no atomics).  One thread = one pixel x 4 channels.
__global__ void __launch_bounds__(256)
k_maxpool3x3s2_fwd_idx(const float* __restrict__ in, float* __restrict__ out, uint8_t* __restrict__ idx, int H, int W, int C,
                       int64_t total) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int C4 = C >> 2, Ho = H >> 1, Wo = W >> 1;
  const int c4 = (int)(t % C4);
  int64_t p = t / C4;
  const int wo = (int)(p % Wo); p /= Wo;
  const int ho = (int)(p % Ho);
  const int64_t n = p / Ho;
  float m[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
  int at[4] = {-1, -1, -1, -1};
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const int y = ho * 2 + a - 1;
    if (y < 0 || y >= H) continue;
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      const int x = wo * 2 + b - 1;
      if (x < 0 || x >= W) continue;
      const float4 v = *reinterpret_cast<const float4*>(in + ((n * H + y) * W + x) * C + c4 * 4);
      const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (at[k] < 0 || e[k] > m[k] || e[k] != e[k]) { m[k] = e[k]; at[k] = a * 3 + b; }
    }
  }
  *reinterpret_cast<float4*>(out + t * 4) = make_float4(m[0], m[1], m[2], m[3]);
  *reinterpret_cast<uchar4*>(idx + t * 4) = make_uchar4((unsigned char)at[0], (unsigned char)at[1], (unsigned char)at[2],
                                                        (unsigned char)at[3]);
}

__global__ void __launch_bounds__(256)
k_maxpool3x3s2_bwd(const float* __restrict__ gout, const uint8_t* __restrict__ idx, float* __restrict__ gin, int H, int W, int C,
                   int64_t total) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int C4 = C >> 2, Ho = H >> 1, Wo = W >> 1;
  const int c4 = (int)(t % C4);
  int64_t p = t / C4;
  const int x = (int)(p % W); p /= W;
  const int y = (int)(p % H);
  const int64_t n = p / H;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  const int ho0 = y >> 1, ho1 = (y + 1) >> 1, wo0 = x >> 1, wo1 = (x + 1) >> 1;
#pragma unroll
  for (int a = 0; a < 2; ++a) {
    const int ho = a ? ho1 : ho0;
    if ((a && ho1 == ho0) || ho >= Ho) continue;
    const int kh = y - (2 * ho - 1);
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const int wo = b ? wo1 : wo0;
      if ((b && wo1 == wo0) || wo >= Wo) continue;
      const int me = kh * 3 + (x - (2 * wo - 1));
      const int64_t o = ((n * Ho + ho) * Wo + wo) * C + c4 * 4;
      const uchar4 at = *reinterpret_cast<const uchar4*>(idx + o);
      const float4 g = *reinterpret_cast<const float4*>(gout + o);
      if (at.x == me) acc[0] += g.x;
      if (at.y == me) acc[1] += g.y;
      if (at.z == me) acc[2] += g.z;
      if (at.w == me) acc[3] += g.w;
    }
  }
  *reinterpret_cast<float4*>(gin + t * 4) = make_float4(acc[0], acc[1], acc[2], acc[3]);
}
}  // namespace mpb200

extern "C" int mp_maxpool3x3s2_forward_idx(const float* in, float* out, void* idx, int N, int H, int W, int C, void* stream) {
  MP_REQUIRE(in && out && idx, "mp_maxpool3x3s2_forward_idx: null pointer");
  MP_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0 && H % 2 == 0 && W % 2 == 0, "mp_maxpool3x3s2_forward_idx: bad dims");
  const int64_t total = (int64_t)N * (H / 2) * (W / 2) * (C / 4);
  mpb200::k_maxpool3x3s2_fwd_idx<<<(unsigned)((total + 255) / 256), 256, 0, mp_stream(stream)>>>(in, out, (uint8_t*)idx, H, W, C, total);
  MP_LAUNCH_CHECK("mp_maxpool3x3s2_forward_idx");
  return 0;
}

extern "C" int mp_maxpool3x3s2_backward(const float* grad_out, const void* idx, float* grad_in, int N, int H, int W, int C,
                                        void* stream) {
  MP_REQUIRE(grad_out && idx && grad_in, "mp_maxpool3x3s2_backward: null pointer");
  MP_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0 && H % 2 == 0 && W % 2 == 0, "mp_maxpool3x3s2_backward: bad dims");
  const int64_t total = (int64_t)N * H * W * (C / 4);
  mpb200::k_maxpool3x3s2_bwd<<<(unsigned)((total + 255) / 256), 256, 0, mp_stream(stream)>>>(grad_out, (const uint8_t*)idx, grad_in, H,
                                                                                             W, C, total);
  MP_LAUNCH_CHECK("mp_maxpool3x3s2_backward");
  return 0;
}

extern "C" int mp_pack_conv_weights(const float* w, void* out_hi, void* out_lo, int Cout, int Cin, int T, int rows_pad, int dgrad,
                                    void* stream) {
  MP_REQUIRE(w && out_hi && out_lo && Cout > 0 && Cin > 0 && T > 0, "mp_pack_conv_weights: bad arguments");
  const int rows = dgrad ? Cin : Cout;
  MP_REQUIRE(rows_pad >= rows, "mp_pack_conv_weights: rows_pad %d < %d rows", rows_pad, rows);
  const int64_t total = (int64_t)rows_pad * T * (dgrad ? Cout : Cin);
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  mpb200::k_pack_conv_weights<<<(unsigned)blocks, 256, 0, mp_stream(stream)>>>(w, (bf16*)out_hi, (bf16*)out_lo, Cout, Cin, T, rows,
                                                                              rows_pad, dgrad);
  MP_LAUNCH_CHECK("mp_pack_conv_weights");
  return 0;
}

extern "C" int mp_conv_wgrad(const float* x, const float* dy, float* dw, int N, int D, int H, int W, int Cin, int Cout, int KD,
                             int KH, int KW, void* stream) {
  using namespace mpb200;
  MP_REQUIRE(x && dy && dw, "mp_conv_wgrad: null pointer");
  MP_REQUIRE(N > 0 && D > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0, "mp_conv_wgrad: bad dims");
  MP_REQUIRE(Cin % 4 == 0 && Cout % 4 == 0, "mp_conv_wgrad: channel counts must be multiples of 4 (16-byte rows), got %d -> %d",
             Cin, Cout);
  MP_REQUIRE(KD % 2 == 1 && KH % 2 == 1 && KW % 2 == 1 && KD * KH * KW <= 65535, "mp_conv_wgrad: odd kernel sizes only");
  WgradParams p;
  p.x = x; p.dy = dy; p.dw = dw;
  p.N = N; p.D = D; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.KD = KD; p.KH = KH; p.KW = KW;
  p.tiles_n = (Cin + WG_TN - 1) / WG_TN;
  p.P = (int64_t)N * D * H * W;
  const int T = KD * KH * KW;
  const int64_t tiles = (int64_t)((Cout + WG_TM - 1) / WG_TM) * p.tiles_n * T;
  // split K so that the grid is ~8 CTAs per SM deep, slices of at least 8 stages
  int64_t slices = (8 * 148 + tiles - 1) / tiles;
  const int64_t max_slices = (p.P + 8 * WG_TK - 1) / (8 * WG_TK);
  if (slices > max_slices) slices = max_slices;
  if (slices < 1) slices = 1;
  if (slices > 65535) slices = 65535;
  p.chunk = ((p.P + slices - 1) / slices + WG_TK - 1) / WG_TK * WG_TK;
  slices = (p.P + p.chunk - 1) / p.chunk;
  cudaError_t e = cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)Cout * T * Cin, mp_stream(stream));
  MP_REQUIRE(e == cudaSuccess, "mp_conv_wgrad: memset: %s", cudaGetErrorString(e));
  dim3 grid((unsigned)(tiles / T), (unsigned)T, (unsigned)slices);
  k_conv_wgrad<<<grid, WG_THREADS, 0, mp_stream(stream)>>>(p);
  MP_LAUNCH_CHECK("mp_conv_wgrad");
  return 0;
}

extern "C" int mp_bias_grad(const float* dy, float* db, int64_t P, int C, void* stream) {
  MP_REQUIRE(dy && db && P > 0 && C > 0, "mp_bias_grad: bad arguments");
  cudaError_t e = cudaMemsetAsync(db, 0, sizeof(float) * (size_t)C, mp_stream(stream));
  MP_REQUIRE(e == cudaSuccess, "mp_bias_grad: memset: %s", cudaGetErrorString(e));
  const int vec = (C % 4 == 0) ? 4 : 1;
  const int Cc = C <= 1024 ? C : 1024;                  // channels per block column (C % 4 == 0 keeps chunks 16-byte aligned)
  const int CV = Cc / vec;
  int R = 256 / CV;
  if (R < 1) R = 1;
  const int threads = R * CV;
  const int64_t rows = (int64_t)R * 64;
  dim3 grid((unsigned)((P + rows - 1) / rows), (unsigned)((C + Cc - 1) / Cc));
  MP_REQUIRE(grid.y <= 65535 && (vec == 4 || C <= 256), "mp_bias_grad: unsupported channel count %d", C);
  if (vec == 4) mpb200::k_bias_grad<4><<<grid, threads, 0, mp_stream(stream)>>>(dy, db, P, C, Cc, rows);
  else mpb200::k_bias_grad<1><<<grid, threads, 0, mp_stream(stream)>>>(dy, db, P, C, Cc, rows);
  MP_LAUNCH_CHECK("mp_bias_grad");
  return 0;
}

extern "C" int mp_group_norm_backward(const float* x, const float* dy, const double* stats, const float* gamma, float* dx,
                                      double* dgamma, double* dbeta, double* workspace, int N, int64_t S, int C, int G, float eps,
                                      void* stream) {
  MP_REQUIRE(x && dy && stats && dx && dgamma && dbeta && workspace, "mp_group_norm_backward: null pointer");
  MP_REQUIRE(N > 0 && N <= 65535 && S > 0 && C > 0 && G > 0 && C % G == 0, "mp_group_norm_backward: bad dims");
  cudaStream_t st = mp_stream(stream);
  cudaError_t e = cudaMemsetAsync(workspace, 0, sizeof(double) * (size_t)N * G * 2, st);
  if (e == cudaSuccess) e = cudaMemsetAsync(dgamma, 0, sizeof(double) * (size_t)C, st);
  if (e == cudaSuccess) e = cudaMemsetAsync(dbeta, 0, sizeof(double) * (size_t)C, st);
  MP_REQUIRE(e == cudaSuccess, "mp_group_norm_backward: memset: %s", cudaGetErrorString(e));
  MP_REQUIRE(C <= 1024 && (C % 4 == 0 || C <= 256), "mp_group_norm_backward: unsupported channel count %d", C);
  const int vec = (C % 4 == 0) ? 4 : 1, CV = C / vec;
  int R = 256 / CV;
  if (R < 1) R = 1;
  const int threads = R * CV;
  // 16 rows per thread: enough blocks (>= 8 per SM on the large tensors) to hide the latency of a streaming read
  const int64_t rows = (int64_t)R * 16;
  dim3 grid((unsigned)((S + rows - 1) / rows), (unsigned)N);
  const size_t shb = sizeof(double) * (size_t)(2 * C + 2 * G);
  if (vec == 4)
    mpb200::k_gn_bwd_reduce<4><<<grid, threads, shb, st>>>(x, dy, stats, gamma, workspace, dgamma, dbeta, S, C, G, eps, rows);
  else
    mpb200::k_gn_bwd_reduce<1><<<grid, threads, shb, st>>>(x, dy, stats, gamma, workspace, dgamma, dbeta, S, C, G, eps, rows);
  MP_LAUNCH_CHECK("mp_group_norm_backward (reduce)");
  const int64_t total = (int64_t)N * S * C;
  float* coef = reinterpret_cast<float*>(workspace + (size_t)N * G * 2);          // [N][3][C] floats after the group sums
  mpb200::k_gn_bwd_coef<<<N, 256, 0, st>>>(stats, gamma, workspace, coef, S, C, G, eps);
  MP_LAUNCH_CHECK("mp_group_norm_backward (coefficients)");
  const int64_t total_vec = total / vec;
  if (vec == 4)
    mpb200::k_gn_bwd_apply<4><<<(unsigned)((total_vec + 255) / 256), 256, 0, st>>>(x, dy, coef, dx, S, C, total_vec);
  else
    mpb200::k_gn_bwd_apply<1><<<(unsigned)((total_vec + 255) / 256), 256, 0, st>>>(x, dy, coef, dx, S, C, total_vec);
  MP_LAUNCH_CHECK("mp_group_norm_backward (apply)");
  return 0;
}
