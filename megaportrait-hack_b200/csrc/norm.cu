// GroupNorm on channels-last tensors, split into (1) statistics, (2) per-(sample,channel) affine finalisation,
// (3) a fused affine + residual + activation + bf16-split pass.  All HBM-bound: (1) reads x once, (3) reads x
// (and the residual) once and writes each requested output once.  The convolution epilogues can produce (1)
// for free, so a GroupNorm between two convolutions costs one read + one write of the activation.
#include "common.cuh"

// ------------------------------------------------------------------------------------------------ statistics
// grid (chunks, N); block = R rows x CV channel-vectors.  Each thread owns VEC fixed channels and walks rows.
template <int VEC>
__global__ void k_gn_stats(const float* __restrict__ x, double* __restrict__ stats, int64_t S, int C, int G,
                           int rows_per_block) {
  extern __shared__ double sh[];  // [G][2]
  const int n = blockIdx.y;
  const int CV = C / VEC;
  const int R = blockDim.x / CV;
  const int r = threadIdx.x / CV, cv = threadIdx.x % CV;
  for (int i = threadIdx.x; i < 2 * G; i += blockDim.x) sh[i] = 0.0;
  __syncthreads();
  float s[VEC], q[VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) s[k] = q[k] = 0.f;
  if (r < R) {
    int64_t row0 = (int64_t)blockIdx.x * rows_per_block;
    int64_t row1 = row0 + rows_per_block;
    if (row1 > S) row1 = S;
    const float* base = x + (int64_t)n * S * C + cv * VEC;
    // fp32 partial sums over <= rows_per_block/R rows per thread, promoted to double below
    for (int64_t row = row0 + r; row < row1; row += R) {
      if (VEC == 4) {
        float4 v = *reinterpret_cast<const float4*>(base + row * C);
        s[0] += v.x; q[0] += v.x * v.x;
        s[1 % VEC] += v.y; q[1 % VEC] += v.y * v.y;
        s[2 % VEC] += v.z; q[2 % VEC] += v.z * v.z;
        s[3 % VEC] += v.w; q[3 % VEC] += v.w * v.w;
      } else {
        float v = base[row * C];
        s[0] += v; q[0] += v * v;
      }
    }
    const int cpg = C / G;
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      int g = (cv * VEC + k) / cpg;
      atomicAdd(&sh[2 * g], (double)s[k]);
      atomicAdd(&sh[2 * g + 1], (double)q[k]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * G; i += blockDim.x) atomicAdd(&stats[(int64_t)n * 2 * G + i], sh[i]);
}

extern "C" int mp_gn_stats(const float* x, double* stats, int N, int64_t S, int C, int G, void* stream) {
  MP_REQUIRE(x && stats, "mp_gn_stats: null pointer");
  MP_REQUIRE(N > 0 && S > 0 && C > 0 && G > 0 && C % G == 0 && N <= 65535, "mp_gn_stats: bad dims");
  const int vec = (C % 4 == 0) ? 4 : 1;
  const int CV = C / vec;
  MP_REQUIRE(CV <= 1024, "mp_gn_stats: C too large");
  int R = 256 / CV;
  if (R < 1) R = 1;
  int threads = R * CV;
  // ~64 rows per thread keeps the fp32 partials short and the grid large
  int rows_per_block = R * 64;
  int64_t chunks = (S + rows_per_block - 1) / rows_per_block;
  dim3 grid((unsigned)chunks, N);
  size_t smem = 2 * G * sizeof(double);
  if (vec == 4)
    k_gn_stats<4><<<grid, threads, smem, mp_stream(stream)>>>(x, stats, S, C, G, rows_per_block);
  else
    k_gn_stats<1><<<grid, threads, smem, mp_stream(stream)>>>(x, stats, S, C, G, rows_per_block);
  MP_LAUNCH_CHECK("mp_gn_stats");
  return 0;
}

// ------------------------------------------------------------------------------------------------ finalize
__global__ void k_gn_finalize(const double* __restrict__ stats, const float* __restrict__ gamma,
                              const float* __restrict__ beta, const float* __restrict__ gamma2,
                              const float* __restrict__ beta2, float* __restrict__ ab, int64_t S, int C, int G,
                              float eps) {
  const int n = blockIdx.x;
  const int cpg = C / G;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    int g = c / cpg;
    double cnt = (double)S * cpg;
    double mean = stats[((int64_t)n * G + g) * 2] / cnt;
    double var = stats[((int64_t)n * G + g) * 2 + 1] / cnt - mean * mean;
    if (var < 0) var = 0;
    double rstd = 1.0 / sqrt(var + (double)eps);
    double a = rstd, b = -mean * rstd;
    if (gamma) { a *= gamma[c]; b *= gamma[c]; }
    if (beta) b += beta[c];
    if (gamma2) { a *= gamma2[c]; b *= gamma2[c]; }
    if (beta2) b += beta2[c];
    ab[((int64_t)n * C + c) * 2] = (float)a;
    ab[((int64_t)n * C + c) * 2 + 1] = (float)b;
  }
}

extern "C" int mp_gn_finalize(const double* stats, const float* gamma, const float* beta, const float* gamma2,
                              const float* beta2, float* ab, int N, int64_t S, int C, int G, float eps, void* stream) {
  MP_REQUIRE(stats && ab, "mp_gn_finalize: null pointer");
  MP_REQUIRE(N > 0 && C % G == 0, "mp_gn_finalize: bad dims");
  k_gn_finalize<<<N, 256, 0, mp_stream(stream)>>>(stats, gamma, beta, gamma2, beta2, ab, S, C, G, eps);
  MP_LAUNCH_CHECK("mp_gn_finalize");
  return 0;
}

// ------------------------------------------------------------------------------------------------ affine + act
template <int VEC>
__global__ void k_affine_act_cl(const float* __restrict__ x, const float* __restrict__ ab,
                                const float* __restrict__ res_f32, const bf16* __restrict__ res_hi,
                                const bf16* __restrict__ res_lo, float* __restrict__ out_f32,
                                bf16* __restrict__ out_hi, bf16* __restrict__ out_lo, int64_t S, int C, int act,
                                int64_t total_vec) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total_vec) return;
  const int CV = C / VEC;
  int cv = (int)(t % CV);
  int64_t row = t / CV;
  int n = (int)(row / S);
  int64_t o = t * VEC;
  float v[VEC];
  if (VEC == 4) {
    float4 q = *reinterpret_cast<const float4*>(x + o);
    v[0] = q.x; v[1 % VEC] = q.y; v[2 % VEC] = q.z; v[3 % VEC] = q.w;
  } else {
    v[0] = x[o];
  }
  if (ab) {
    const float* p = ab + ((int64_t)n * C + cv * VEC) * 2;
#pragma unroll
    for (int k = 0; k < VEC; ++k) v[k] = fmaf(v[k], p[2 * k], p[2 * k + 1]);
  }
  if (res_f32) {
#pragma unroll
    for (int k = 0; k < VEC; ++k) v[k] += res_f32[o + k];
  } else if (res_hi) {
#pragma unroll
    for (int k = 0; k < VEC; ++k) v[k] += mp_join(res_hi[o + k], res_lo[o + k]);
  }
#pragma unroll
  for (int k = 0; k < VEC; ++k) v[k] = mp_apply_act(v[k], act);
  if (out_f32) {
    if (VEC == 4) *reinterpret_cast<float4*>(out_f32 + o) = make_float4(v[0], v[1 % VEC], v[2 % VEC], v[3 % VEC]);
    else out_f32[o] = v[0];
  }
  if (out_hi) {
    if (VEC == 4) mp_store_split4(out_hi, out_lo, o, make_float4(v[0], v[1 % VEC], v[2 % VEC], v[3 % VEC]));
    else mp_split2(v[0], out_hi[o], out_lo[o]);
  }
}

extern "C" int mp_affine_act_cl(const float* x, const float* ab, const float* res_f32, const void* res_hi,
                                const void* res_lo, float* out_f32, void* out_hi, void* out_lo, int N, int64_t S,
                                int C, int act, void* stream) {
  MP_REQUIRE(x && (out_f32 || (out_hi && out_lo)), "mp_affine_act_cl: null pointer");
  MP_REQUIRE(N > 0 && S > 0 && C > 0, "mp_affine_act_cl: bad dims");
  const int vec = (C % 4 == 0) ? 4 : 1;
  int64_t total = (int64_t)N * S * C / vec;
  unsigned grid = (unsigned)((total + 255) / 256);
  if (vec == 4)
    k_affine_act_cl<4><<<grid, 256, 0, mp_stream(stream)>>>(x, ab, res_f32, (const bf16*)res_hi, (const bf16*)res_lo,
                                                            out_f32, (bf16*)out_hi, (bf16*)out_lo, S, C, act, total);
  else
    k_affine_act_cl<1><<<grid, 256, 0, mp_stream(stream)>>>(x, ab, res_f32, (const bf16*)res_hi, (const bf16*)res_lo,
                                                            out_f32, (bf16*)out_hi, (bf16*)out_lo, S, C, act, total);
  MP_LAUNCH_CHECK("mp_affine_act_cl");
  return 0;
}
