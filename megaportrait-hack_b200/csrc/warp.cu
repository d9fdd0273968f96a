// Trilinear warping kernels (memory-bound gathers, no tensor cores).
//
//   mp_grid_sample3d        F.grid_sample(5-D, bilinear, border, align_corners=True)           model.py:1062
//   mp_apply_warping_field  apply_warping_field(v, warp_field) in the reference NCDHW layout    model.py:1028-1065
//   mp_warp_field           compute_rt_warp + F.interpolate(w_em, 64^3) + add                   model.py:965-973
//   mp_warp_fused_cl        the pipeline's fused form on channels-last volumes (+ optional sum over D, :1171)
//
// Coordinate arithmetic follows ATen (GridSampler.h:27-36,58-60; UpSample.h area_pixel_compute_source_index;
// RangeFactories linspace) operation by operation so that results agree with the CPU oracle to ~1e-6.
#include "common.cuh"
#include "warp_common.cuh"

// ------------------------------------------------------------------------------------------------ NCDHW gather
__global__ void k_grid_sample3d(const float* __restrict__ v, const float* __restrict__ grid, float* __restrict__ out,
                                int C, int D, int H, int W, int64_t So) {
  int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= So) return;
  const int n = blockIdx.z;
  const float* g = grid + ((int64_t)n * So + s) * 3;
  float ix = unnormalize_clip(g[0], W), iy = unnormalize_clip(g[1], H), iz = unnormalize_clip(g[2], D);
  Taps t;
  make_taps(ix, iy, iz, D, H, W, t);
  int c0 = blockIdx.y * GS_CCHUNK, c1 = min(C, c0 + GS_CCHUNK);
  int64_t in_cs = (int64_t)D * H * W;
  gather_channels(v + (int64_t)n * C * in_cs, out + (int64_t)n * C * So + s, t, c0, c1, in_cs, So);
}

extern "C" int mp_grid_sample3d(const float* v, const float* grid, float* out, int N, int C, int D, int H, int W,
                                int Do, int Ho, int Wo, void* stream) {
  MP_REQUIRE(v && grid && out, "mp_grid_sample3d: null pointer");
  MP_REQUIRE(N > 0 && N <= 65535 && C > 0 && D > 0 && H > 0 && W > 0 && Do > 0 && Ho > 0 && Wo > 0,
             "mp_grid_sample3d: bad dims");
  int64_t So = (int64_t)Do * Ho * Wo;
  dim3 g((unsigned)((So + 127) / 128), (C + GS_CCHUNK - 1) / GS_CCHUNK, N);
  k_grid_sample3d<<<g, 128, 0, mp_stream(stream)>>>(v, grid, out, C, D, H, W, So);
  MP_LAUNCH_CHECK("mp_grid_sample3d");
  return 0;
}

__global__ void k_apply_warping_field(const float* __restrict__ v, const float* __restrict__ wf,
                                      float* __restrict__ out, int C, int D, int H, int W, int Df, int Hf, int Wf) {
  const int64_t S = (int64_t)D * H * W;
  int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  const int n = blockIdx.z;
  int w = (int)(s % W), h = (int)((s / W) % H), d = (int)(s / ((int64_t)W * H));
  int d0, d1, h0, h1, w0, w1;
  float ld, lh, lw;
  src_ac_true(d, Df, D, d0, d1, ld);
  src_ac_true(h, Hf, H, h0, h1, lh);
  src_ac_true(w, Wf, W, w0, w1, lw);
  const int64_t fs = (int64_t)Df * Hf * Wf;
  const float* f = wf + (int64_t)n * 3 * fs;
  float fx = resample_flow(f, Df, Hf, Wf, d0, d1, ld, h0, h1, lh, w0, w1, lw);
  float fy = resample_flow(f + fs, Df, Hf, Wf, d0, d1, ld, h0, h1, lh, w0, w1, lw);
  float fz = resample_flow(f + 2 * fs, Df, Hf, Wf, d0, d1, ld, h0, h1, lh, w0, w1, lw);
  // grid + flow, then the reference's "2*g/(size-1) - 1" (model.py:1052-1058), then ATen's un-normalisation
  float gx = 2.0f * (linspace_m1_1(w, W) + fx) / (float)(W - 1) - 1.0f;
  float gy = 2.0f * (linspace_m1_1(h, H) + fy) / (float)(H - 1) - 1.0f;
  float gz = 2.0f * (linspace_m1_1(d, D) + fz) / (float)(D - 1) - 1.0f;
  Taps t;
  make_taps(unnormalize_clip(gx, W), unnormalize_clip(gy, H), unnormalize_clip(gz, D), D, H, W, t);
  int c0 = blockIdx.y * GS_CCHUNK, c1 = min(C, c0 + GS_CCHUNK);
  gather_channels(v + (int64_t)n * C * S, out + (int64_t)n * C * S + s, t, c0, c1, S, S);
}

extern "C" int mp_apply_warping_field(const float* v, const float* warp_field, float* out, int N, int C, int D, int H,
                                      int W, int Df, int Hf, int Wf, void* stream) {
  MP_REQUIRE(v && warp_field && out, "mp_apply_warping_field: null pointer");
  MP_REQUIRE(N > 0 && N <= 65535 && C > 0 && D > 1 && H > 1 && W > 1 && Df > 0 && Hf > 0 && Wf > 0,
             "mp_apply_warping_field: bad dims");
  int64_t S = (int64_t)D * H * W;
  dim3 g((unsigned)((S + 127) / 128), (C + GS_CCHUNK - 1) / GS_CCHUNK, N);
  k_apply_warping_field<<<g, 128, 0, mp_stream(stream)>>>(v, warp_field, out, C, D, H, W, Df, Hf, Wf);
  MP_LAUNCH_CHECK("mp_apply_warping_field");
  return 0;
}

// ------------------------------------------------------------------------------------------------ 64^3 field
// value of channel k of the G^3 warp field at (z,y,x): affine_grid(align_corners=False) + trilinear(em, ac=False)
__device__ __forceinline__ void field_at(const float* __restrict__ em, const float* __restrict__ th, int E, int G,
                                         int z, int y, int x, float f[3]) {
  // affine_grid base coordinate: linspace(-1,1,G) * (G-1)/G  ==  (2i+1)/G - 1
  float cx = linspace_m1_1(x, G) * (float)(G - 1) / (float)G;
  float cy = linspace_m1_1(y, G) * (float)(G - 1) / (float)G;
  float cz = linspace_m1_1(z, G) * (float)(G - 1) / (float)G;
  int z0, z1, y0, y1, x0, x1;
  float lz, ly, lx;
  src_ac_false(z, E, G, z0, z1, lz);
  src_ac_false(y, E, G, y0, y1, ly);
  src_ac_false(x, E, G, x0, x1, lx);
  auto at = [&](int zz, int yy, int xx, int k) { return __ldg(em + (((int64_t)zz * E + yy) * E + xx) * 3 + k); };
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float a = (1.f - ly) * ((1.f - lx) * at(z0, y0, x0, k) + lx * at(z0, y0, x1, k)) +
              ly * ((1.f - lx) * at(z0, y1, x0, k) + lx * at(z0, y1, x1, k));
    float b = (1.f - ly) * ((1.f - lx) * at(z1, y0, x0, k) + lx * at(z1, y0, x1, k)) +
              ly * ((1.f - lx) * at(z1, y1, x0, k) + lx * at(z1, y1, x1, k));
    float emv = (1.f - lz) * a + lz * b;
    float rt = th[k * 4 + 0] * cx + th[k * 4 + 1] * cy + th[k * 4 + 2] * cz + th[k * 4 + 3];
    f[k] = rt + emv;
  }
}

__global__ void k_warp_field(const float* __restrict__ em, const float* __restrict__ theta, float* __restrict__ out,
                             int E, int G) {
  const int64_t S = (int64_t)G * G * G;
  int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  const int n = blockIdx.y;
  int x = (int)(s % G), y = (int)((s / G) % G), z = (int)(s / ((int64_t)G * G));
  float f[3];
  field_at(em + (int64_t)n * E * E * E * 3, theta + n * 12, E, G, z, y, x, f);
#pragma unroll
  for (int k = 0; k < 3; ++k) out[((int64_t)n * 3 + k) * S + s] = f[k];
}

extern "C" int mp_warp_field(const float* em_cl, const float* theta, float* out, int N, int E, int G, void* stream) {
  MP_REQUIRE(em_cl && theta && out, "mp_warp_field: null pointer");
  MP_REQUIRE(N > 0 && N <= 65535 && E > 0 && G > 0, "mp_warp_field: bad dims");
  int64_t S = (int64_t)G * G * G;
  dim3 g((unsigned)((S + 255) / 256), N);
  k_warp_field<<<g, 256, 0, mp_stream(stream)>>>(em_cl, theta, out, E, G);
  MP_LAUNCH_CHECK("mp_warp_field");
  return 0;
}

// ------------------------------------------------------------------------------------------------ fused CL warp
// Block = one (sample, h, 64-wide w segment); D is walked in slabs of WF_DSLAB depth slices.  Phase 1: one thread per
// voxel of the slab builds the sampling taps (closed form of the whole flow -> grid chain)
// into shared memory.  Phase 2: threads re-map to (voxel, 4-channel vector) and gather with 16-byte loads that are
// contiguous across the lanes of a warp (channels-last), then either store or accumulate the depth sum in registers.
constexpr int WF_WSEG = 64;
constexpr int WF_DSLAB = 4;
constexpr int WF_THREADS = WF_WSEG * WF_DSLAB;   // 256
constexpr int WF_MAXACC = 8;                     // ceil(WSEG * C/4 / THREADS) for C <= 128

template <bool SUM_D>
__global__ void __launch_bounds__(WF_THREADS)
k_warp_fused_cl(const float* __restrict__ v, const float* __restrict__ em, const float* __restrict__ theta,
                float* __restrict__ out_f32, bf16* __restrict__ out_hi, bf16* __restrict__ out_lo, int Nv, int C,
                int D, int H, int W, int E, int G) {
  __shared__ int s_off[WF_THREADS][8];
  __shared__ float s_w[WF_THREADS][8];
  const int n = blockIdx.z, h = blockIdx.y, wbase = blockIdx.x * WF_WSEG;
  const int C4 = C >> 2;
  const int items = WF_WSEG * C4;                 // (voxel-in-row, channel vector) work items per depth slice
  const float* vn = v + (Nv == 1 ? 0 : (int64_t)n * D * H * W * C);
  const float* emn = em + (int64_t)n * E * E * E * 3;
  const float* th = theta + n * 12;
  float4 acc[WF_MAXACC];
#pragma unroll
  for (int i = 0; i < WF_MAXACC; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);

  for (int dbase = 0; dbase < D; dbase += WF_DSLAB) {
    // ---- phase 1: taps for voxel (d, h, w)
    {
      int dl = threadIdx.x / WF_WSEG, wl = threadIdx.x % WF_WSEG;
      int d = dbase + dl, w = wbase + wl;
      Taps t;
      if (d < D && w < W) {
        // flow: G^3 field resampled to (D,H,W) with align_corners=True (model.py:1036)
        int z0, z1, y0, y1, x0, x1;
        float lz, ly, lx;
        src_ac_true(d, G, D, z0, z1, lz);
        src_ac_true(h, G, H, y0, y1, ly);
        src_ac_true(w, G, W, x0, x1, lx);
        float F[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int a = 0; a < 2; ++a) {
          float wz = a ? lz : 1.f - lz;
          if (wz == 0.f) continue;
#pragma unroll
          for (int b = 0; b < 2; ++b) {
            float wy = b ? ly : 1.f - ly;
            if (wy == 0.f) continue;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              float wx = c ? lx : 1.f - lx;
              if (wx == 0.f) continue;
              float f[3];
              field_at(emn, th, E, G, a ? z1 : z0, b ? y1 : y0, c ? x1 : x0, f);
              float wgt = wz * wy * wx;
              F[0] += wgt * f[0]; F[1] += wgt * f[1]; F[2] += wgt * f[2];
            }
          }
        }
        float gx = 2.0f * (linspace_m1_1(w, W) + F[0]) / (float)(W - 1) - 1.0f;
        float gy = 2.0f * (linspace_m1_1(h, H) + F[1]) / (float)(H - 1) - 1.0f;
        float gz = 2.0f * (linspace_m1_1(d, D) + F[2]) / (float)(D - 1) - 1.0f;
        make_taps(unnormalize_clip(gx, W), unnormalize_clip(gy, H), unnormalize_clip(gz, D), D, H, W, t);
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) { t.off[k] = -1; t.w[k] = 0.f; }
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) { s_off[threadIdx.x][k] = t.off[k]; s_w[threadIdx.x][k] = t.w[k]; }
    }
    __syncthreads();
    // ---- phase 2: gather
    for (int dl = 0; dl < WF_DSLAB && dbase + dl < D; ++dl) {
      int d = dbase + dl;
#pragma unroll
      for (int i = 0; i < WF_MAXACC; ++i) {
        int it = threadIdx.x + i * WF_THREADS;
        if (it >= items) break;
        int wl = it / C4, c4 = it % C4;
        if (wbase + wl >= W) continue;
        const int* po = s_off[dl * WF_WSEG + wl];
        const float* pw = s_w[dl * WF_WSEG + wl];
        float4 r = SUM_D ? acc[i] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          int off = po[k];
          if (off >= 0) {
            float4 q = __ldg(reinterpret_cast<const float4*>(vn + (int64_t)off * C + c4 * 4));
            float wk = pw[k];
            r.x = fmaf(q.x, wk, r.x); r.y = fmaf(q.y, wk, r.y); r.z = fmaf(q.z, wk, r.z); r.w = fmaf(q.w, wk, r.w);
          }
        }
        if (SUM_D) {
          acc[i] = r;
        } else {
          int64_t o = ((((int64_t)n * D + d) * H + h) * W + (wbase + wl)) * C + c4 * 4;
          if (out_f32) *reinterpret_cast<float4*>(out_f32 + o) = r;
          if (out_hi) mp_store_split4(out_hi, out_lo, o, r);
        }
      }
    }
    __syncthreads();
  }
  if (SUM_D) {
#pragma unroll
    for (int i = 0; i < WF_MAXACC; ++i) {
      int it = threadIdx.x + i * WF_THREADS;
      if (it >= items) break;
      int wl = it / C4, c4 = it % C4;
      if (wbase + wl >= W) continue;
      int64_t o = (((int64_t)n * H + h) * W + (wbase + wl)) * C + c4 * 4;
      if (out_f32) *reinterpret_cast<float4*>(out_f32 + o) = acc[i];
      if (out_hi) mp_store_split4(out_hi, out_lo, o, acc[i]);
    }
  }
}

extern "C" int mp_warp_fused_cl(const float* v, const float* em_cl, const float* theta, float* out_f32, void* out_hi,
                                void* out_lo, int N, int Nv, int C, int D, int H, int W, int E, int G, int sum_d,
                                void* stream) {
  MP_REQUIRE(v && em_cl && theta && (out_f32 || (out_hi && out_lo)), "mp_warp_fused_cl: null pointer");
  MP_REQUIRE(N > 0 && N <= 65535 && (Nv == 1 || Nv == N), "mp_warp_fused_cl: bad batch N=%d Nv=%d", N, Nv);
  MP_REQUIRE(C % 4 == 0 && WF_WSEG * (C / 4) <= WF_MAXACC * WF_THREADS, "mp_warp_fused_cl: unsupported C=%d", C);
  MP_REQUIRE(D > 1 && H > 1 && W > 1 && H <= 65535 && E > 0 && G > 0, "mp_warp_fused_cl: bad dims");
  dim3 g((W + WF_WSEG - 1) / WF_WSEG, H, N);
  if (sum_d)
    k_warp_fused_cl<true><<<g, WF_THREADS, 0, mp_stream(stream)>>>(v, em_cl, theta, out_f32, (bf16*)out_hi,
                                                                   (bf16*)out_lo, Nv, C, D, H, W, E, G);
  else
    k_warp_fused_cl<false><<<g, WF_THREADS, 0, mp_stream(stream)>>>(v, em_cl, theta, out_f32, (bf16*)out_hi,
                                                                    (bf16*)out_lo, Nv, C, D, H, W, E, G);
  MP_LAUNCH_CHECK("mp_warp_fused_cl");
  return 0;
}

// ------------------------------------------------------------------------------------------------ workspace variants
// NCDHW gathers with data-dependent (jittery or random) grids are bound by L1 wavefronts: every lane's 4-byte tap
// sits in a different 32-byte sector, and the other 28 bytes of that sector belong to the SAME channel, which the
// lane does not need.  In channels-last order one sector holds 8 channels of the tap, i.e. exactly what the other
// lanes of the warp want.  So: (1) transpose the volume to channels-last into a caller-provided workspace
// (coalesced on both sides), (2) gather 16-byte channel vectors from it -- contiguous across the lanes of a warp for
// ANY grid -- and transpose each 32-voxel x C result tile back through shared memory so that the NCDHW stores are
// 128-byte rows.  ncu (profiles/): 23 sectors per request for the direct kernel on the "spread" grid, 4 here.
constexpr int GW_VOX = 32;
constexpr int GW_THREADS = 256;

template <int MODE>   // 0: explicit grid [N,So,3]; 1: warp field [N,3,Df,Hf,Wf] (apply_warping_field semantics)
__global__ void __launch_bounds__(GW_THREADS)
k_gather_cl_to_ncdhw(const float* __restrict__ vcl, const float* __restrict__ aux, float* __restrict__ out, int C, int D,
                     int H, int W, int Do, int Ho, int Wo, int Df, int Hf, int Wf) {
  extern __shared__ float s_tile[];               // [C][GW_VOX + 1]
  __shared__ float s_pix[GW_VOX][3];
  __shared__ int s_off[GW_VOX][8];
  __shared__ float s_w[GW_VOX][8];
  const int n = blockIdx.y;
  const int64_t So = (int64_t)Do * Ho * Wo;
  const int64_t s0 = (int64_t)blockIdx.x * GW_VOX;
  const int C4 = C >> 2;
  // phase A1: one thread per voxel -> clamped pixel coordinates
  if (threadIdx.x < GW_VOX) {
    const int64_t s = s0 + threadIdx.x;
    float ix = 0.f, iy = 0.f, iz = 0.f;
    if (s < So) {
      if (MODE == 0) {
        const float* g = aux + ((int64_t)n * So + s) * 3;
        ix = unnormalize_clip(g[0], W); iy = unnormalize_clip(g[1], H); iz = unnormalize_clip(g[2], D);
      } else {
        const int w = (int)(s % Wo), h = (int)((s / Wo) % Ho), d = (int)(s / ((int64_t)Wo * Ho));
        int d0, d1, h0, h1, w0, w1;
        float ld, lh, lw;
        src_ac_true(d, Df, D, d0, d1, ld);
        src_ac_true(h, Hf, H, h0, h1, lh);
        src_ac_true(w, Wf, W, w0, w1, lw);
        const int64_t fs = (int64_t)Df * Hf * Wf;
        const float* f = aux + (int64_t)n * 3 * fs;
        const float fx = resample_flow(f, Df, Hf, Wf, d0, d1, ld, h0, h1, lh, w0, w1, lw);
        const float fy = resample_flow(f + fs, Df, Hf, Wf, d0, d1, ld, h0, h1, lh, w0, w1, lw);
        const float fz = resample_flow(f + 2 * fs, Df, Hf, Wf, d0, d1, ld, h0, h1, lh, w0, w1, lw);
        ix = unnormalize_clip(2.0f * (linspace_m1_1(w, W) + fx) / (float)(W - 1) - 1.0f, W);
        iy = unnormalize_clip(2.0f * (linspace_m1_1(h, H) + fy) / (float)(H - 1) - 1.0f, H);
        iz = unnormalize_clip(2.0f * (linspace_m1_1(d, D) + fz) / (float)(D - 1) - 1.0f, D);
      }
    }
    s_pix[threadIdx.x][0] = ix; s_pix[threadIdx.x][1] = iy; s_pix[threadIdx.x][2] = iz;
  }
  __syncthreads();
  // phase A2: one thread per (voxel, corner) -> offset + weight (same arithmetic as make_taps)
  for (int i = threadIdx.x; i < GW_VOX * 8; i += GW_THREADS) {
    const int vox = i >> 3, k = i & 7;
    const float ix = s_pix[vox][0], iy = s_pix[vox][1], iz = s_pix[vox][2];
    const float fx = floorf(ix), fy = floorf(iy), fz = floorf(iz);
    const float tx = ix - fx, ty = iy - fy, tz = iz - fz;
    const int dx = k & 1, dy = (k >> 1) & 1, dz = k >> 2;
    const int x = (int)fx + dx, y = (int)fy + dy, z = (int)fz + dz;
    const float w = (dx ? tx : 1.f - tx) * (dy ? ty : 1.f - ty) * (dz ? tz : 1.f - tz);
    const bool ok = (s0 + vox < So) && x >= 0 && x < W && y >= 0 && y < H && z >= 0 && z < D;
    s_off[vox][k] = ok ? (z * H + y) * W + x : -1;
    s_w[vox][k] = ok ? w : 0.f;
  }
  __syncthreads();
  // phase B: gather 16-byte channel vectors (contiguous across lanes for any grid), transpose through shared memory
  const float* vn = vcl + (int64_t)n * D * H * W * C;
  for (int it = threadIdx.x; it < GW_VOX * C4; it += GW_THREADS) {
    const int vox = it / C4, q = it % C4;
    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int off = s_off[vox][k];
      if (off >= 0) {
        const float4 x = __ldg(reinterpret_cast<const float4*>(vn + (int64_t)off * C + q * 4));
        const float wk = s_w[vox][k];
        r.x = fmaf(x.x, wk, r.x); r.y = fmaf(x.y, wk, r.y); r.z = fmaf(x.z, wk, r.z); r.w = fmaf(x.w, wk, r.w);
      }
    }
    float* tp = s_tile + (q * 4) * (GW_VOX + 1) + vox;
    tp[0] = r.x; tp[GW_VOX + 1] = r.y; tp[2 * (GW_VOX + 1)] = r.z; tp[3 * (GW_VOX + 1)] = r.w;
  }
  __syncthreads();
  // phase C: NCDHW stores, GW_VOX consecutive voxels (256 bytes) per channel row
  for (int i = threadIdx.x; i < C * GW_VOX; i += GW_THREADS) {
    const int c = i / GW_VOX, vox = i % GW_VOX;
    if (s0 + vox < So) out[((int64_t)n * C + c) * So + s0 + vox] = s_tile[c * (GW_VOX + 1) + vox];
  }
}

extern "C" size_t mp_gather_workspace_bytes(int N, int C, int D, int H, int W) {
  return (C % 4 == 0 && C <= 512) ? (size_t)N * C * D * H * W * sizeof(float) : 0;
}

static int gather_ws_common(int mode, const float* v, const float* aux, float* out, void* ws, size_t ws_bytes, int N, int C,
                            int D, int H, int W, int Do, int Ho, int Wo, int Df, int Hf, int Wf, void* stream) {
  const size_t need = mp_gather_workspace_bytes(N, C, D, H, W);
  MP_REQUIRE(need > 0, "mp_*_ws: C must be a multiple of 4 and <= 512 (use the direct entry point)");
  MP_REQUIRE(ws && ws_bytes >= need, "mp_*_ws: workspace too small (%zu < %zu)", ws_bytes, need);
  MP_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 15) == 0, "mp_*_ws: workspace must be 16-byte aligned");
  if (int e = mp_nchw_to_cl(v, (float*)ws, nullptr, nullptr, N, C, (int64_t)D * H * W, stream)) return e;
  const int64_t So = (int64_t)Do * Ho * Wo;
  dim3 grid((unsigned)((So + GW_VOX - 1) / GW_VOX), N);
  const size_t smem = (size_t)C * (GW_VOX + 1) * sizeof(float);
  if (mode == 0) {
    if (smem > 48 * 1024) cudaFuncSetAttribute(k_gather_cl_to_ncdhw<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_gather_cl_to_ncdhw<0><<<grid, GW_THREADS, smem, mp_stream(stream)>>>((const float*)ws, aux, out, C, D, H, W, Do, Ho,
                                                                         Wo, 0, 0, 0);
  } else {
    if (smem > 48 * 1024) cudaFuncSetAttribute(k_gather_cl_to_ncdhw<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_gather_cl_to_ncdhw<1><<<grid, GW_THREADS, smem, mp_stream(stream)>>>((const float*)ws, aux, out, C, D, H, W, Do, Ho,
                                                                         Wo, Df, Hf, Wf);
  }
  MP_LAUNCH_CHECK("mp_gather_ws");
  return 0;
}

extern "C" int mp_grid_sample3d_ws(const float* v, const float* grid, float* out, void* workspace, size_t workspace_bytes,
                                   int N, int C, int D, int H, int W, int Do, int Ho, int Wo, void* stream) {
  MP_REQUIRE(v && grid && out, "mp_grid_sample3d_ws: null pointer");
  MP_REQUIRE(N > 0 && N <= 65535 && C > 0 && D > 0 && H > 0 && W > 0 && Do > 0 && Ho > 0 && Wo > 0,
             "mp_grid_sample3d_ws: bad dims");
  return gather_ws_common(0, v, grid, out, workspace, workspace_bytes, N, C, D, H, W, Do, Ho, Wo, 0, 0, 0, stream);
}

extern "C" int mp_apply_warping_field_ws(const float* v, const float* warp_field, float* out, void* workspace,
                                         size_t workspace_bytes, int N, int C, int D, int H, int W, int Df, int Hf, int Wf,
                                         void* stream) {
  MP_REQUIRE(v && warp_field && out, "mp_apply_warping_field_ws: null pointer");
  MP_REQUIRE(N > 0 && N <= 65535 && C > 0 && D > 1 && H > 1 && W > 1 && Df > 0 && Hf > 0 && Wf > 0,
             "mp_apply_warping_field_ws: bad dims");
  return gather_ws_common(1, v, warp_field, out, workspace, workspace_bytes, N, C, D, H, W, D, H, W, Df, Hf, Wf, stream);
}

// ------------------------------------------------------------------------------------------------ backward (row f-2)
// Gradients of apply_warping_field(v, warp_field) (model.py:1028-1065), i.e. of F.interpolate(flow, trilinear,
// align_corners=True) -> identity grid + flow -> the reference's re-normalisation -> F.grid_sample(bilinear, border,
// align_corners=True), following ATen's grid_sampler_3d_backward (GridSampler.h: clip_coordinates_set_grad zeroes the
// coordinate gradient AT and beyond the border, out-of-range corners contribute nothing):
//   grad_v[n,c,tap]  += grad_out[n,c,s] * w_tap                                   (scatter, fp32 atomics)
//   d out / d ix      = sum_taps v[tap] * d w_tap / d ix  (and iy, iz); the chain pix = ((g+1)/2)(size-1),
//                       g = 2 (lin + flow) / (size-1) - 1 has d pix / d flow = 1, times the clip mask
//   grad_flow[n,k,:] += (d out / d pix_k) spread over the 8 flow voxels of the align_corners=True resample
// One thread = one output voxel x GS_CCHUNK channels; both outputs must be zero-initialised by the caller.
__device__ __forceinline__ float clip_grad_mask(float p, int size) {
  // p = un-normalised coordinate BEFORE clipping: gradient 1 strictly inside (0, size-1), else 0
  return (p <= 0.f || p >= (float)(size - 1)) ? 0.f : 1.f;
}

__global__ void k_apply_warping_field_bwd(const float* __restrict__ go, const float* __restrict__ v,
                                          const float* __restrict__ wf, float* __restrict__ gv, float* __restrict__ gwf,
                                          int C, int D, int H, int W, int Df, int Hf, int Wf) {
  const int64_t S = (int64_t)D * H * W;
  int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  const int n = blockIdx.z;
  int w = (int)(s % W), h = (int)((s / W) % H), d = (int)(s / ((int64_t)W * H));
  int d0, d1, h0, h1, w0, w1;
  float ld, lh, lw;
  src_ac_true(d, Df, D, d0, d1, ld);
  src_ac_true(h, Hf, H, h0, h1, lh);
  src_ac_true(w, Wf, W, w0, w1, lw);
  const int64_t fs = (int64_t)Df * Hf * Wf;
  const float* f = wf + (int64_t)n * 3 * fs;
  const float fx = resample_flow(f, Df, Hf, Wf, d0, d1, ld, h0, h1, lh, w0, w1, lw);
  const float fy = resample_flow(f + fs, Df, Hf, Wf, d0, d1, ld, h0, h1, lh, w0, w1, lw);
  const float fz = resample_flow(f + 2 * fs, Df, Hf, Wf, d0, d1, ld, h0, h1, lh, w0, w1, lw);
  const float gx = 2.0f * (linspace_m1_1(w, W) + fx) / (float)(W - 1) - 1.0f;
  const float gy = 2.0f * (linspace_m1_1(h, H) + fy) / (float)(H - 1) - 1.0f;
  const float gz = 2.0f * (linspace_m1_1(d, D) + fz) / (float)(D - 1) - 1.0f;
  const float px = ((gx + 1.f) / 2.f) * (float)(W - 1), py = ((gy + 1.f) / 2.f) * (float)(H - 1),
              pz = ((gz + 1.f) / 2.f) * (float)(D - 1);
  const float ix = fminf((float)(W - 1), fmaxf(px, 0.f)), iy = fminf((float)(H - 1), fmaxf(py, 0.f)),
              iz = fminf((float)(D - 1), fmaxf(pz, 0.f));
  const float flx = floorf(ix), fly = floorf(iy), flz = floorf(iz);
  const int x0 = (int)flx, y0 = (int)fly, z0 = (int)flz;
  const float tx = ix - flx, ty = iy - fly, tz = iz - flz;
  const int c0 = blockIdx.y * GS_CCHUNK, c1 = min(C, c0 + GS_CCHUNK);
  const float* vn = v + (int64_t)n * C * S;
  const float* gon = go + (int64_t)n * C * S + s;
  float* gvn = gv ? gv + (int64_t)n * C * S : nullptr;
  float gix = 0.f, giy = 0.f, giz = 0.f;
  for (int c = c0; c < c1; ++c) {
    const float g = __ldg(gon + (int64_t)c * S);
    const float* vc = vn + (int64_t)c * S;
    float* gvc = gvn ? gvn + (int64_t)c * S : nullptr;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int dx = k & 1, dy = (k >> 1) & 1, dz = k >> 2;
      const int x = x0 + dx, y = y0 + dy, z = z0 + dz;
      if (x >= W || y >= H || z >= D) continue;                 // within_bounds_3d (cells are never negative after the clip)
      const float wx = dx ? tx : 1.f - tx, wy = dy ? ty : 1.f - ty, wz = dz ? tz : 1.f - tz;
      const int64_t off = ((int64_t)z * H + y) * W + x;
      if (gvc) atomicAdd(gvc + off, g * (wx * wy * wz));
      if (gwf) {
        const float val = __ldg(vc + off) * g;
        gix += val * (dx ? 1.f : -1.f) * wy * wz;
        giy += val * (dy ? 1.f : -1.f) * wx * wz;
        giz += val * (dz ? 1.f : -1.f) * wx * wy;
      }
    }
  }
  if (!gwf) return;
  gix *= clip_grad_mask(px, W);
  giy *= clip_grad_mask(py, H);
  giz *= clip_grad_mask(pz, D);
  if (gix == 0.f && giy == 0.f && giz == 0.f) return;
  float* gf = gwf + (int64_t)n * 3 * fs;
  const float gk[3] = {gix, giy, giz};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    if (gk[k] == 0.f) continue;
    float* p = gf + k * fs;
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const float wgt = (a ? ld : 1.f - ld) * (b ? lh : 1.f - lh) * (e ? lw : 1.f - lw);
          if (wgt != 0.f) atomicAdd(p + ((int64_t)(a ? d1 : d0) * Hf + (b ? h1 : h0)) * Wf + (e ? w1 : w0), gk[k] * wgt);
        }
  }
}

extern "C" int mp_apply_warping_field_backward(const float* grad_out, const float* v, const float* warp_field, float* grad_v,
                                               float* grad_warp_field, int N, int C, int D, int H, int W, int Df, int Hf,
                                               int Wf, void* stream) {
  MP_REQUIRE(grad_out && v && warp_field && (grad_v || grad_warp_field), "mp_apply_warping_field_backward: null pointer");
  MP_REQUIRE(N > 0 && N <= 65535 && C > 0 && D > 1 && H > 1 && W > 1 && Df > 0 && Hf > 0 && Wf > 0,
             "mp_apply_warping_field_backward: bad dims");
  const int64_t S = (int64_t)D * H * W;
  dim3 g((unsigned)((S + 127) / 128), (C + GS_CCHUNK - 1) / GS_CCHUNK, N);
  k_apply_warping_field_bwd<<<g, 128, 0, mp_stream(stream)>>>(grad_out, v, warp_field, grad_v, grad_warp_field, C, D, H, W, Df, Hf,
                                                            Wf);
  MP_LAUNCH_CHECK("mp_apply_warping_field_backward");
  return 0;
}
