// Brick-staged trilinear grid_sample / apply_warping_field on the reference's NCDHW layout (model.py:1028-1065).
//
// The gather is memory-bound: 4 B read + 4 B written per (voxel, channel) and 12 B of grid per voxel.  What the plain
// gather kernels of warp.cu lose is NOT HBM bandwidth but the L2 -> L1 path: 8 taps per output re-read the volume
// through 32-byte sectors of which one float is used.  Here every voxel of the input travels HBM -> shared memory once
// per tile that needs it, as a TMA box, and the 8 taps are served from shared memory:
//
//   * CTA = one output tile (tz x ty x tx voxels, <= 4096) x one group of channels.
//   * Phase A (once per CTA): sampling coordinates of every voxel of the tile -> cell (x0, y0, z0) + fractions, kept in
//     shared memory (16 B per voxel) and shared by all channels; bounding box of the cells by warp reductions.
//   * If the bounding box (+1 for the far corner) fits the brick (BD x BH x BW floats): per channel ONE 4-D TMA box
//     load of the NCDHW tensor [W, H, D, N*C] into a 3-deep ring (out-of-bounds elements are zero-filled, which is what
//     ATen's `within_bounds` test does to the far corner at the border), interpolation from shared memory, result tile
//     staged in shared memory and written by ONE TMA box store (128-byte NCDHW rows) while the next channel is computed.
//   * Bank conflicts: a warp reads 32 taps of 32 different voxels.  With a jittery grid these land on random banks
//     (~3.4-way conflicts).  Because the tap pattern is the same for every channel, the tile's voxels are bucketed ONCE
//     by the bank of their base tap (one shared-memory atomic per voxel) and a warp takes one voxel per bucket: all 8
//     tap loads of every channel are then conflict-free.  Degenerate grids (everything in one bucket, e.g. the
//     reference's corner-only sampling) keep the natural order, where equal addresses are broadcasts.
//   * Tiles whose bounding box does not fit (adversarial / large-displacement grids) fall back, per tile, to direct
//     global gathers with the same arithmetic (the `_ws` entry points of warp.cu remain the better choice when the
//     caller knows the grid is a random permutation).
//
// Arithmetic is shared with warp.cu (warp_common.cuh), so results are bit-identical to k_grid_sample3d.
#include <cuda.h>
#include <stdlib.h>

#include <mutex>

#include "common.cuh"
#include "warp_common.cuh"

namespace {

constexpr int GB_MAX_THREADS = 1024;     // two builds: 512 threads (<= 128 registers) and 1024 threads (<= 64)
constexpr int GB_NB = 4;            // brick ring: two channel pairs
constexpr int GB_TILE_MAX = 4096;   // voxels per tile: each thread builds GB_TILE_MAX / T taps ...
constexpr int GB_ROWS_MAX = 160;    // ... and each warp keeps GB_ROWS_MAX / (T / 32) warp rows of tap records in registers
constexpr uint32_t GB_SMEM_LIMIT = 227 * 1024;

struct __align__(16) TapRec {
  int ov;         // brick offset (floats, < 65536) | voxel index inside the tile << 16; -1 = empty slot
  float tx, ty, tz;
};

struct GbParams {
  const float* v;
  const float* aux;     // MODE 0: grid [N, Do, Ho, Wo, 3]; MODE 1: warp field [N, 3, Df, Hf, Wf]
  float* out;
  unsigned char* fitmap;   // [N, tiles]: 1 = the tile was served from bricks; NULL = unfit tiles gather inside the kernel
  int C, D, H, W, Do, Ho, Wo, Df, Hf, Wf;
  int tz, ty, tx, tile_vox;
  int tiles_z, tiles_y, tiles_x;
  int BD, BH, BW;       // brick box (floats); BW % 4 == 0
  int ngroups;          // channel groups per tile
  int sort;             // bucket the voxels by the bank of their base tap
  int rows_cap;         // warp rows the slot table can hold (<= GB_RPW * warps)
  uint32_t box_bytes, brick_bytes, stage_bytes, off_stage, off_misc;   // box_bytes: one TMA box; brick_bytes: ring stride
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  long long t0 = clock64();
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) break;
    if (clock64() - t0 > 4000000000LL) __trap();   // a protocol bug must not hang the GPU box
  }
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

struct Misc {
  unsigned long long full[GB_NB];
  int bmin[3], bmax[3];
  int rows, sorted;
  unsigned short hist[GB_MAX_THREADS / 32][32];   // per-warp bucket counts, then exclusive prefix over the warps
};

// Shared-memory map (byte offsets from the 128-byte aligned base):
//   [0, 4 * brick_bytes)                     brick ring; before the first TMA load: the tile's raw grid (MODE 0)
//   [off_stage, off_stage + 4 * stage_bytes) two staging buffers x two channels; before the channel loop, together with
//                                            brick 3: the slot table (GB_RPW * warps * 32 tap records)
//   [off_misc, ...)                          mbarriers, bounding box, bucket histograms
// clamped pixel coordinates of output voxel (w, h, d) of sample n: from 3 grid values (MODE 0) or from the warp field
// (MODE 1: flow resample, identity grid, the reference's re-normalisation -- model.py:1036-1058)
template <int MODE>
__device__ __forceinline__ void gb_pix(const GbParams& p, int n, int w, int h, int d, const float* g, float& ix, float& iy,
                                       float& iz) {
  if (MODE == 0) {
    ix = unnormalize_clip(g[0], p.W); iy = unnormalize_clip(g[1], p.H); iz = unnormalize_clip(g[2], p.D);
  } else {
    int d0, d1, h0, h1, w0, w1;
    float ld, lh, lw;
    src_ac_true(d, p.Df, p.D, d0, d1, ld);
    src_ac_true(h, p.Hf, p.H, h0, h1, lh);
    src_ac_true(w, p.Wf, p.W, w0, w1, lw);
    const int64_t fs = (int64_t)p.Df * p.Hf * p.Wf;
    const float* f = p.aux + (int64_t)n * 3 * fs;
    const float fx = resample_flow(f, p.Df, p.Hf, p.Wf, d0, d1, ld, h0, h1, lh, w0, w1, lw);
    const float fy = resample_flow(f + fs, p.Df, p.Hf, p.Wf, d0, d1, ld, h0, h1, lh, w0, w1, lw);
    const float fz = resample_flow(f + 2 * fs, p.Df, p.Hf, p.Wf, d0, d1, ld, h0, h1, lh, w0, w1, lw);
    ix = unnormalize_clip(2.0f * (linspace_m1_1(w, p.W) + fx) / (float)(p.W - 1) - 1.0f, p.W);
    iy = unnormalize_clip(2.0f * (linspace_m1_1(h, p.H) + fy) / (float)(p.H - 1) - 1.0f, p.H);
    iz = unnormalize_clip(2.0f * (linspace_m1_1(d, p.D) + fz) / (float)(p.D - 1) - 1.0f, p.D);
  }
}

// taps of cell (x0, y0, z0) + fractions for direct global gathers (same arithmetic as make_taps)
__device__ __forceinline__ void gb_cell_taps(const GbParams& p, int x0, int y0, int z0, float ftx, float fty, float ftz, Taps& tp) {
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int dx = k & 1, dy = (k >> 1) & 1, dz = k >> 2;
    const int x = x0 + dx, y = y0 + dy, z = z0 + dz;
    const float wgt = (dx ? ftx : 1.f - ftx) * (dy ? fty : 1.f - fty) * (dz ? ftz : 1.f - ftz);
    const bool ok = x < p.W && y < p.H && z < p.D;
    tp.off[k] = ok ? (z * p.H + y) * p.W + x : -1;
    tp.w[k] = ok ? wgt : 0.f;
  }
}

// Second pass (only when the caller passed a fit map): tiles whose sampling region did not fit the brick, gathered
// straight from global memory at full occupancy.  One CTA = one tile x one share of the channels.
template <int MODE>
__global__ void __launch_bounds__(256)
k_gs_unfit(const GbParams p, int shares) {
  const int n = blockIdx.z;
  const int tiles = p.tiles_x * p.tiles_y * p.tiles_z;
  if (p.fitmap[(int64_t)n * tiles + blockIdx.x]) return;
  int t = blockIdx.x;
  const int tix = t % p.tiles_x; t /= p.tiles_x;
  const int tiy = t % p.tiles_y; t /= p.tiles_y;
  const int ox0 = tix * p.tx, oy0 = tiy * p.ty, oz0 = t * p.tz;
  const int cb = p.C / shares, ce = p.C % shares;
  const int c0 = blockIdx.y * cb + min((int)blockIdx.y, ce), nch = cb + ((int)blockIdx.y < ce ? 1 : 0);
  const int64_t S = (int64_t)p.D * p.H * p.W, So = (int64_t)p.Do * p.Ho * p.Wo;
  const float* vn = p.v + ((int64_t)n * p.C + c0) * S;
  float* on = p.out + ((int64_t)n * p.C + c0) * So;
  for (int vid = threadIdx.x; vid < p.tile_vox; vid += blockDim.x) {
    const int lx = vid % p.tx, ly = (vid / p.tx) % p.ty, lz = vid / (p.tx * p.ty);
    const int w = ox0 + lx, h = oy0 + ly, d = oz0 + lz;
    if (w >= p.Wo || h >= p.Ho || d >= p.Do) continue;
    const int64_t s = ((int64_t)d * p.Ho + h) * p.Wo + w;
    float ix, iy, iz;
    gb_pix<MODE>(p, n, w, h, d, p.aux + ((int64_t)n * So + s) * 3, ix, iy, iz);
    Taps tp;
    make_taps(ix, iy, iz, p.D, p.H, p.W, tp);
    gather_channels(vn, on + s, tp, 0, nch, S, So);
  }
}

// NS warp rows of one channel pair, branch-free (empty slots read offset 0 and skip the store), so that the loads of
// all NS slots can be in flight together.
template <int NS, int O1, int O2, int RPW>
__device__ __forceinline__ void gb_slots(const int (&m_ov)[RPW], const float (&m_tx)[RPW], const float (&m_ty)[RPW],
                                         const float (&m_tz)[RPW], const float* __restrict__ bka,
                                         const float* __restrict__ bkb, float* __restrict__ sta, float* __restrict__ stb,
                                         int o1r, int o2r) {
  const int o1 = O1 > 0 ? O1 : o1r, o2 = O1 > 0 ? O2 : o2r;     // brick row / slice pitch: immediates when known
#pragma unroll
  for (int j = 0; j < (NS < RPW ? NS : RPW); ++j) {
    const bool ok = m_ov[j] >= 0;
    const int off = ok ? (m_ov[j] & 0xffff) : 0, vid = m_ov[j] >> 16;
    const float* qa = bka + off;
    const float* qb = bkb + off;
    const float a000 = qa[0], a001 = qa[1], a010 = qa[o1], a011 = qa[o1 + 1];
    const float a100 = qa[o2], a101 = qa[o2 + 1], a110 = qa[o2 + o1], a111 = qa[o2 + o1 + 1];
    const float b000 = qb[0], b001 = qb[1], b010 = qb[o1], b011 = qb[o1 + 1];
    const float b100 = qb[o2], b101 = qb[o2 + 1], b110 = qb[o2 + o1], b111 = qb[o2 + o1 + 1];
    const float tx = m_tx[j], ty = m_ty[j], tz = m_tz[j];
    const float ax = 1.f - tx, ay = 1.f - ty, az = 1.f - tz;
    // same association as make_taps: ((wx * wy) * wz), accumulated in tap order k = dx + 2 dy + 4 dz
    const float w00 = ax * ay, w01 = tx * ay, w10 = ax * ty, w11 = tx * ty;
    const float w0 = w00 * az, w1 = w01 * az, w2 = w10 * az, w3 = w11 * az;
    const float w4 = w00 * tz, w5 = w01 * tz, w6 = w10 * tz, w7 = w11 * tz;
    float ra = 0.f, rb = 0.f;
    ra = fmaf(a000, w0, ra); rb = fmaf(b000, w0, rb);
    ra = fmaf(a001, w1, ra); rb = fmaf(b001, w1, rb);
    ra = fmaf(a010, w2, ra); rb = fmaf(b010, w2, rb);
    ra = fmaf(a011, w3, ra); rb = fmaf(b011, w3, rb);
    ra = fmaf(a100, w4, ra); rb = fmaf(b100, w4, rb);
    ra = fmaf(a101, w5, ra); rb = fmaf(b101, w5, rb);
    ra = fmaf(a110, w6, ra); rb = fmaf(b110, w6, rb);
    ra = fmaf(a111, w7, ra); rb = fmaf(b111, w7, rb);
    if (ok) {
      sta[vid] = ra;
      stb[vid] = rb;                     // without a second channel: stale brick data, never stored to global memory
    }
  }
}

template <int MODE, int O1, int O2, int T>
__global__ void __launch_bounds__(T, 1)
k_grid_sample_brick(const __grid_constant__ CUtensorMap map_in, const __grid_constant__ CUtensorMap map_out,
                    const GbParams p) {
  constexpr int GB_VPT = GB_TILE_MAX / T, GB_RPW = GB_ROWS_MAX / (T / 32);
  extern __shared__ uint8_t smem_raw[];
  // 128-byte alignment by an OFFSET into the shared array (an integer round trip would turn every access below into a
  // generic-address load with 64-bit address arithmetic)
  uint8_t* base = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
  float* bricks = reinterpret_cast<float*>(base);
  float* stage = reinterpret_cast<float*>(base + p.off_stage);
  TapRec* table = reinterpret_cast<TapRec*>(base + 3u * p.brick_bytes);
  Misc* misc = reinterpret_cast<Misc*>(base + p.off_misc);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int n = blockIdx.z;
  const int grp = blockIdx.y;
  int t = blockIdx.x;
  const int tix = t % p.tiles_x; t /= p.tiles_x;
  const int tiy = t % p.tiles_y; t /= p.tiles_y;
  const int tiz = t;
  const int ox0 = tix * p.tx, oy0 = tiy * p.ty, oz0 = tiz * p.tz;
  // channels of this group: balanced split of C over ngroups
  const int cbase = p.C / p.ngroups, cextra = p.C % p.ngroups;
  const int c0 = grp * cbase + min(grp, cextra);
  const int nch = cbase + (grp < cextra ? 1 : 0);
  const int64_t So = (int64_t)p.Do * p.Ho * p.Wo;
  const int rows_cap = p.rows_cap;

  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < GB_NB; ++i) mbar_init(smem_u32(&misc->full[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 3; ++i) { misc->bmin[i] = 0x7fffffff; misc->bmax[i] = -1; }
  }
  misc->hist[warp][lane] = 0;
  // ---------------------------------------------------------------- phase 0: the tile's grid values, coalesced
  if (MODE == 0) {
    // rows of 3 * tx floats, contiguous in the grid tensor and 16-byte aligned (tx % 4 == 0): 16-byte loads, all of a
    // thread's loads in flight before the first store
    float4* graw4 = reinterpret_cast<float4*>(bricks);
    const int row4 = (3 * p.tx) >> 2, nrows = p.tz * p.ty;
    const int valid4 = (3 * min(p.tx, p.Wo - ox0)) >> 2;
    const float* gn = p.aux + (int64_t)n * So * 3;
    const bool vec = ((reinterpret_cast<uintptr_t>(gn) & 15) == 0);
    constexpr int U = 6;
    for (int e0 = tid; e0 < nrows * row4; e0 += U * blockDim.x) {
      float4 val[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int e = e0 + u * blockDim.x;
        val[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (e < nrows * row4) {
          const int r = e / row4, c = e - r * row4;
          const int h = oy0 + r % p.ty, d = oz0 + r / p.ty;
          if (c < valid4 && h < p.Ho && d < p.Do) {
            const float* src = gn + (((int64_t)d * p.Ho + h) * p.Wo + ox0) * 3 + 4 * c;
            if (vec) val[u] = __ldg(reinterpret_cast<const float4*>(src));
            else val[u] = make_float4(__ldg(src), __ldg(src + 1), __ldg(src + 2), __ldg(src + 3));
          }
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int e = e0 + u * blockDim.x;
        if (e < nrows * row4) graw4[e] = val[u];
      }
    }
  }
  __syncthreads();

  // ---------------------------------------------------------------- phase A: cells + fractions (registers), bounding box
  int a_cell[GB_VPT];                // x0 | y0 << 10 | z0 << 20, or -1: voxel outside the output / beyond the tile
  float a_tx[GB_VPT], a_ty[GB_VPT], a_tz[GB_VPT];
  int mn[3] = {0x7fffffff, 0x7fffffff, 0x7fffffff}, mx[3] = {-1, -1, -1};
#pragma unroll
  for (int j = 0; j < GB_VPT; ++j) {
    const int vid = (warp * GB_VPT + j) * 32 + lane;
    a_cell[j] = -1; a_tx[j] = a_ty[j] = a_tz[j] = 0.f;
    if (vid >= p.tile_vox) continue;
    const int lx = vid % p.tx, ly = (vid / p.tx) % p.ty, lz = vid / (p.tx * p.ty);
    const int w = ox0 + lx, h = oy0 + ly, d = oz0 + lz;
    if (w >= p.Wo || h >= p.Ho || d >= p.Do) continue;
    float ix, iy, iz;
    gb_pix<MODE>(p, n, w, h, d, bricks + vid * 3, ix, iy, iz);
    const float fx = floorf(ix), fy = floorf(iy), fz = floorf(iz);
    int x0 = (int)fx, y0 = (int)fy, z0 = (int)fz;
    a_tx[j] = ix - fx; a_ty[j] = iy - fy; a_tz[j] = iz - fz;
    // the clip maps NaN coordinates to 0 (fmaxf returns the non-NaN operand), so cells are always inside the volume
    x0 = min(max(x0, 0), p.W - 1); y0 = min(max(y0, 0), p.H - 1); z0 = min(max(z0, 0), p.D - 1);
    a_cell[j] = x0 | (y0 << 10) | (z0 << 20);
    mn[0] = min(mn[0], x0); mn[1] = min(mn[1], y0); mn[2] = min(mn[2], z0);
    mx[0] = max(mx[0], x0); mx[1] = max(mx[1], y0); mx[2] = max(mx[2], z0);
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int a = __reduce_min_sync(0xffffffffu, mn[i]), b = __reduce_max_sync(0xffffffffu, mx[i]);
    if (lane == 0) { atomicMin(&misc->bmin[i], a); atomicMax(&misc->bmax[i], b); }
  }
  __syncthreads();
  if (misc->bmax[0] < 0) return;      // tile entirely outside the output (cannot happen with the host's grid)
  // brick origin: the whole axis when it fits together with its zero-filled far neighbour, else the box minimum
  // (x rounded down to 4 floats: TMA box rows start 16-byte aligned)
  const int xb = (p.W + 1 <= p.BW) ? 0 : (misc->bmin[0] & ~3);
  const int yb = (p.H + 1 <= p.BH) ? 0 : misc->bmin[1];
  const int zb = (p.D + 1 <= p.BD) ? 0 : misc->bmin[2];
  const bool fit = (misc->bmax[0] + 1 - xb < p.BW) && (misc->bmax[1] + 1 - yb < p.BH) && (misc->bmax[2] + 1 - zb < p.BD);

  if (p.fitmap) {
    if (tid == 0 && grp == 0) p.fitmap[(int64_t)n * (p.tiles_x * p.tiles_y * p.tiles_z) + blockIdx.x] = fit ? 1 : 0;
    if (!fit) return;                 // k_gs_unfit gathers this tile at full occupancy
  }
  if (!fit) {
    // ------------------------------------------------------------ fallback: direct global gathers for this tile
    const int64_t S = (int64_t)p.D * p.H * p.W;
    const float* vn = p.v + ((int64_t)n * p.C + c0) * S;
    float* on = p.out + ((int64_t)n * p.C + c0) * So;
#pragma unroll 1
    for (int j = 0; j < GB_VPT; ++j) {
      int cell = -1;
      float ftx = 0.f, fty = 0.f, ftz = 0.f;
#pragma unroll
      for (int q = 0; q < GB_VPT; ++q)
        if (q == j) { cell = a_cell[q]; ftx = a_tx[q]; fty = a_ty[q]; ftz = a_tz[q]; }
      if (cell < 0) continue;
      const int x0 = cell & 1023, y0 = (cell >> 10) & 1023, z0 = cell >> 20;
      Taps tp;
      gb_cell_taps(p, x0, y0, z0, ftx, fty, ftz, tp);
      const int vid = (warp * GB_VPT + j) * 32 + lane;
      const int lx = vid % p.tx, ly = (vid / p.tx) % p.ty, lz = vid / (p.tx * p.ty);
      const int64_t s = ((int64_t)(oz0 + lz) * p.Ho + (oy0 + ly)) * p.Wo + (ox0 + lx);
      gather_channels(vn, on + s, tp, 0, nch, S, So);
    }
    return;
  }

  // ---------------------------------------------------------------- first bricks in flight while the taps are ordered
  const int nc_in = n * p.C + c0;
  if (tid == 0) {
    fence_proxy_async();                 // the ring held generic-proxy data (raw grid) until the barrier above
    for (int i = 0; i < 3 && i < nch; ++i) {
      const uint32_t bar = smem_u32(&misc->full[i]);
      mbar_expect_tx(bar, p.box_bytes);
      tma_load_4d(smem_u32(bricks) + i * p.brick_bytes, &map_in, bar, xb, yb, zb, nc_in + i);
    }
  }
  // ---------------------------------------------------------------- phase B: brick offsets; rank inside the bank bucket
  // (warp-private histogram updated by match.any leaders: no shared-memory atomics)
  const int slice = p.BH * p.BW;
  int b_off[GB_VPT], b_rank[GB_VPT];
  for (int i = tid; i < rows_cap * 32; i += blockDim.x) table[i].ov = -1;
#pragma unroll
  for (int j = 0; j < GB_VPT; ++j) {
    const int cell = a_cell[j];
    const int x0 = cell & 1023, y0 = (cell >> 10) & 1023, z0 = cell >> 20;
    b_off[j] = cell < 0 ? -1 : (z0 - zb) * slice + (y0 - yb) * p.BW + (x0 - xb);
    b_rank[j] = 0;
    if (p.sort) {
      const int key = cell < 0 ? 32 + lane : (b_off[j] & 31);       // voxels outside the output: singleton groups
      const unsigned m = __match_any_sync(0xffffffffu, key);
      const int leader = __ffs(m) - 1;
      int cnt = 0;
      if (lane == leader && cell >= 0) cnt = misc->hist[warp][key];
      cnt = __shfl_sync(0xffffffffu, cnt, leader);
      b_rank[j] = cnt + __popc(m & ((1u << lane) - 1u));
      if (lane == leader && cell >= 0) misc->hist[warp][key] = (unsigned short)(cnt + __popc(m));
      __syncwarp();
    }
  }
  __syncthreads();
  if (warp == 0) {
    int rows = (p.tile_vox + 31) >> 5, sorted = 0;
    if (p.sort) {
      int run = 0;                                  // exclusive prefix over the warps for bucket `lane`
      for (int w = 0; w < nwarps; ++w) { const int c = misc->hist[w][lane]; misc->hist[w][lane] = (unsigned short)run; run += c; }
      const int cmax = __reduce_max_sync(0xffffffffu, run);
      if (cmax <= rows_cap) { sorted = 1; rows = cmax; }
    }
    if (lane == 0) { misc->rows = rows; misc->sorted = sorted; }
  }
  __syncthreads();
  const int rows = misc->rows;
  const bool sorted = misc->sorted != 0;
#pragma unroll
  for (int j = 0; j < GB_VPT; ++j) {
    if (b_off[j] < 0) continue;
    const int vid = (warp * GB_VPT + j) * 32 + lane;
    const int b = b_off[j] & 31;
    const int slot = sorted ? (misc->hist[warp][b] + b_rank[j]) * 32 + b : vid;
    TapRec r;
    r.ov = b_off[j] | (vid << 16); r.tx = a_tx[j]; r.ty = a_ty[j]; r.tz = a_tz[j];
    table[slot] = r;
  }
  __syncthreads();
  // my rows of the slot table -> registers (they serve every channel of the group)
  int m_ov[GB_RPW];
  float m_tx[GB_RPW], m_ty[GB_RPW], m_tz[GB_RPW];
#pragma unroll
  for (int j = 0; j < GB_RPW; ++j) {
    const int row = warp + j * nwarps;
    m_ov[j] = -1; m_tx[j] = m_ty[j] = m_tz[j] = 0.f;
    if (row < rows) {
      const TapRec r = table[row * 32 + lane];
      m_ov[j] = r.ov; m_tx[j] = r.tx; m_ty[j] = r.ty; m_tz[j] = r.tz;
    }
  }
  const int nslots = rows > warp ? (rows - warp + nwarps - 1) / nwarps : 0;     // warp-uniform, <= GB_RPW
  __syncthreads();                       // the slot table's shared memory becomes brick 3 + staging
  if (tid == 0 && nch > 3) {
    fence_proxy_async();
    const uint32_t bar = smem_u32(&misc->full[3]);
    mbar_expect_tx(bar, p.box_bytes);
    tma_load_4d(smem_u32(bricks) + 3 * p.brick_bytes, &map_in, bar, xb, yb, zb, nc_in + 3);
  }

  // ---------------------------------------------------------------- channel pairs
  const int o1 = p.BW, o2 = slice;
  const int brick_f = (int)(p.brick_bytes >> 2), stage_f = (int)(p.stage_bytes >> 2);
  for (int ca = 0; ca < nch; ca += 2) {
    const bool has2 = ca + 1 < nch;
    const int sa = ca & 3;
    const uint32_t par = (uint32_t)((ca >> 2) & 1);
    mbar_wait(smem_u32(&misc->full[sa]), par);
    if (has2) mbar_wait(smem_u32(&misc->full[sa + 1]), par);
    const float* bka = bricks + sa * brick_f;
    const float* bkb = bka + brick_f;
    float* sta = stage + ((ca >> 1) & 1) * 2 * stage_f;
    float* stb = sta + stage_f;
    switch (nslots) {
      case 10: gb_slots<10, O1, O2, GB_RPW>(m_ov, m_tx, m_ty, m_tz, bka, bkb, sta, stb, o1, o2); break;
      case 9: gb_slots<9, O1, O2, GB_RPW>(m_ov, m_tx, m_ty, m_tz, bka, bkb, sta, stb, o1, o2); break;
      case 8: gb_slots<8, O1, O2, GB_RPW>(m_ov, m_tx, m_ty, m_tz, bka, bkb, sta, stb, o1, o2); break;
      case 7: gb_slots<7, O1, O2, GB_RPW>(m_ov, m_tx, m_ty, m_tz, bka, bkb, sta, stb, o1, o2); break;
      case 6: gb_slots<6, O1, O2, GB_RPW>(m_ov, m_tx, m_ty, m_tz, bka, bkb, sta, stb, o1, o2); break;
      case 5: gb_slots<5, O1, O2, GB_RPW>(m_ov, m_tx, m_ty, m_tz, bka, bkb, sta, stb, o1, o2); break;
      case 4: gb_slots<4, O1, O2, GB_RPW>(m_ov, m_tx, m_ty, m_tz, bka, bkb, sta, stb, o1, o2); break;
      case 3: gb_slots<3, O1, O2, GB_RPW>(m_ov, m_tx, m_ty, m_tz, bka, bkb, sta, stb, o1, o2); break;
      case 2: gb_slots<2, O1, O2, GB_RPW>(m_ov, m_tx, m_ty, m_tz, bka, bkb, sta, stb, o1, o2); break;
      case 1: gb_slots<1, O1, O2, GB_RPW>(m_ov, m_tx, m_ty, m_tz, bka, bkb, sta, stb, o1, o2); break;
      default: break;
    }
    fence_proxy_async();                 // my staged values -> visible to the TMA store
    if (tid == 0) bulk_wait_read0();     // the stores of the previous pair have finished reading the other staging buffer
    __syncthreads();
    if (tid == 0) {
      tma_store_4d(&map_out, smem_u32(sta), ox0, oy0, oz0, nc_in + ca);
      if (has2) tma_store_4d(&map_out, smem_u32(stb), ox0, oy0, oz0, nc_in + ca + 1);
      bulk_commit();
#pragma unroll
      for (int k = 0; k < 2; ++k) {      // every warp is done with these two bricks: refill their ring slots
        if (ca + 4 + k < nch) {
          const uint32_t bar = smem_u32(&misc->full[sa + k]);
          mbar_expect_tx(bar, p.box_bytes);
          tma_load_4d(smem_u32(bricks) + (sa + k) * p.brick_bytes, &map_in, bar, xb, yb, zb, nc_in + ca + 4 + k);
        }
      }
    }
  }
  if (tid == 0) bulk_wait_read0();       // shared memory must outlive the last store's reads
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encoder() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// fp32 NCDHW tensor as a 4-D map [X, Y, Z, N*C] with a (bx, by, bz, 1) box, no swizzle, zero fill out of bounds
int encode_vol_map(CUtensorMap* m, const void* ptr, int64_t NC, int Z, int Y, int X, int bz, int by, int bx) {
  cuuint64_t dims[4] = {(cuuint64_t)X, (cuuint64_t)Y, (cuuint64_t)Z, (cuuint64_t)NC};
  cuuint64_t strides[3] = {(cuuint64_t)X * 4, (cuuint64_t)X * Y * 4, (cuuint64_t)X * Y * Z * 4};
  cuuint32_t box[4] = {(cuuint32_t)bx, (cuuint32_t)by, (cuuint32_t)bz, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = encoder()(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : mp_set_error("mp_grid_sample3d_brick: cuTensorMapEncodeTiled failed (%d)", (int)r);
}

// tuning overrides (mp_gs_brick_tune): {tz, ty, tx, BD, BH, BW, threads, ngroups}; 0 = automatic
int g_tune[8] = {0, 0, 0, 0, 0, 0, 0, 0};

int round_up(int v, int m) { return (v + m - 1) / m * m; }

void tile_shape(int Do, int Ho, int Wo, int threads, int& tz, int& ty, int& tx) {
  const int T = threads > 512 ? 1024 : 512;
  tx = g_tune[2] ? g_tune[2] : (Wo < 64 ? Wo : 64);
  ty = g_tune[1] ? g_tune[1] : (Ho < 16 ? Ho : 16);
  tz = g_tune[0] ? g_tune[0] : (Do < 4 ? Do : 4);
  while (tz > 1 && tz * ty * tx > GB_TILE_MAX / T * threads) --tz;
}

int launch_brick(int mode, const float* v, const float* aux, float* out, void* workspace, size_t workspace_bytes, int N, int C,
                 int D, int H, int W, int Do, int Ho, int Wo, int Df, int Hf, int Wf, int flags, void* stream, const char* who) {
  MP_REQUIRE(encoder() != nullptr, "%s: cuTensorMapEncodeTiled not available from the driver", who);
  MP_REQUIRE(W % 4 == 0 && Wo % 4 == 0, "%s: W and Wo must be multiples of 4 (16-byte TMA rows); use the direct entry point", who);
  MP_REQUIRE(D <= 1023 && H <= 1023 && W <= 1023, "%s: volume extents above 1023 are not supported", who);
  MP_REQUIRE(((reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(out)) & 15) == 0, "%s: v and out must be 16-byte aligned", who);
  GbParams p;
  p.v = v; p.aux = aux; p.out = out;
  p.C = C; p.D = D; p.H = H; p.W = W; p.Do = Do; p.Ho = Ho; p.Wo = Wo; p.Df = Df; p.Hf = Hf; p.Wf = Wf;
  // output tile: whole rows up to 64 wide, 16 rows, as many slices as 4096 voxels allow (at most 4)
  const int threads = g_tune[6] ? g_tune[6] : 512;
  tile_shape(Do, Ho, Wo, threads, p.tz, p.ty, p.tx);
  MP_REQUIRE(threads % 32 == 0 && threads >= 32 && threads <= GB_MAX_THREADS, "%s: bad thread count", who);
  const int T = threads > 512 ? 1024 : 512;                 // kernel build: registers per thread, taps per thread
  const int vpt = GB_TILE_MAX / T, rpw = GB_ROWS_MAX / (T / 32);
  MP_REQUIRE(p.tx % 4 == 0 && p.tx <= 256 && p.ty <= 256 && p.tz <= 256 && p.tz * p.ty * p.tx <= vpt * threads,
             "%s: bad tile (at most %d voxels)", who, vpt * threads);
  p.tile_vox = p.tz * p.ty * p.tx;
  p.tiles_x = (Wo + p.tx - 1) / p.tx; p.tiles_y = (Ho + p.ty - 1) / p.ty; p.tiles_z = (Do + p.tz - 1) / p.tz;
  // brick: tile extent + a halo of 4 cells (y, x) / 1 cell (z) on each side + the far corner, capped by the volume
  p.BW = g_tune[5] ? g_tune[5] : round_up(min(W + 1, p.tx + 9), 4);
  p.BH = g_tune[4] ? g_tune[4] : min(H + 1, p.ty + 8);
  p.BD = g_tune[3] ? g_tune[3] : min(D + 1, p.tz + 2);
  MP_REQUIRE(p.BW % 4 == 0 && p.BW <= 256 && p.BH <= 256 && p.BD <= 256, "%s: bad brick", who);
  p.box_bytes = (uint32_t)(p.BD * p.BH * p.BW * 4);
  p.brick_bytes = (uint32_t)round_up((int)p.box_bytes, 128);
  MP_REQUIRE(p.BD * p.BH * p.BW <= 65536, "%s: brick above 65536 floats", who);
  p.sort = (flags & 1) ? 0 : 1;
  p.stage_bytes = (uint32_t)round_up(p.tile_vox * 4, 128);
  p.off_stage = GB_NB * p.brick_bytes;
  p.off_misc = p.off_stage + 4u * p.stage_bytes;
  const uint32_t smem = p.off_misc + (uint32_t)sizeof(Misc) + 128u;
  // scratch that aliases the ring before the channel loop: raw grid over bricks 0-1.., slot table over brick 3 + staging
  MP_REQUIRE((uint32_t)p.tile_vox * 12u <= p.off_misc || mode != 0, "%s: ring too small for the grid scratch", who);
  p.rows_cap = min(rpw * (threads / 32), (int)((p.brick_bytes + 4u * p.stage_bytes) / 512u));
  MP_REQUIRE((p.tile_vox + 31) / 32 <= p.rows_cap, "%s: slot table does not fit", who);
  MP_REQUIRE(smem <= GB_SMEM_LIMIT, "%s: tile/brick configuration needs %u B of shared memory", who, smem);
  // channel groups: enough CTAs to fill the SMs, at least ~8 channels per CTA to amortise phase A
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t tiles = (int64_t)p.tiles_x * p.tiles_y * p.tiles_z * N;
  int ng = g_tune[7];
  if (!ng) {
    // smallest group count that gives >= 4 waves, else the count whose last wave is fullest
    double best = -1.0;
    const int ng_max = max(1, C / 8);
    for (int g = 1; g <= ng_max; ++g) {
      const int64_t ctas = tiles * g;
      const int64_t waves = (ctas + sms - 1) / sms;
      const double eff = (double)ctas / (double)(waves * sms);
      const double score = eff - 0.002 * g;          // prefer fewer groups (phase A is repeated per group)
      if (score > best + 1e-9) { best = score; ng = g; }
      if (waves >= 4 && eff > 0.95) break;
    }
  }
  MP_REQUIRE(ng >= 1 && ng <= C && ng <= 65535, "%s: bad channel-group count", who);
  p.ngroups = ng;
  MP_REQUIRE(tiles / N <= 0x7fffffff && N <= 65535, "%s: too many tiles", who);
  CUtensorMap mi, mo;
  if (int e = encode_vol_map(&mi, v, (int64_t)N * C, D, H, W, p.BD, p.BH, p.BW)) return e;
  if (int e = encode_vol_map(&mo, out, (int64_t)N * C, Do, Ho, Wo, p.tz, p.ty, p.tx)) return e;
  // the (16, 64, 64) volume's brick (BW = 68, BH = 24) gets its pitches as immediates
  const bool fixed = p.BW == 68 && p.BH == 24;
  typedef void (*KernelFn)(const CUtensorMap, const CUtensorMap, const GbParams);
  KernelFn fn;
  if (T == 512)
    fn = mode == 0 ? (fixed ? k_grid_sample_brick<0, 68, 68 * 24, 512> : k_grid_sample_brick<0, 0, 0, 512>)
                   : (fixed ? k_grid_sample_brick<1, 68, 68 * 24, 512> : k_grid_sample_brick<1, 0, 0, 512>);
  else
    fn = mode == 0 ? (fixed ? k_grid_sample_brick<0, 68, 68 * 24, 1024> : k_grid_sample_brick<0, 0, 0, 1024>)
                   : (fixed ? k_grid_sample_brick<1, 68, 68 * 24, 1024> : k_grid_sample_brick<1, 0, 0, 1024>);
  static bool attr_done[8] = {false, false, false, false, false, false, false, false};
  const int ai = mode * 4 + (fixed ? 2 : 0) + (T == 512 ? 0 : 1);
  if (!attr_done[ai]) {
    cudaError_t ae = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GB_SMEM_LIMIT);
    MP_REQUIRE(ae == cudaSuccess, "%s: cannot opt in to %u B shared memory: %s", who, GB_SMEM_LIMIT, cudaGetErrorString(ae));
    attr_done[ai] = true;
  }
  dim3 grid((unsigned)(tiles / N), (unsigned)ng, (unsigned)N);
  p.fitmap = nullptr;
  if (workspace) {
    MP_REQUIRE(workspace_bytes >= (size_t)tiles, "%s: workspace too small (%zu < %lld)", who, workspace_bytes, (long long)tiles);
    p.fitmap = static_cast<unsigned char*>(workspace);
  }
  fn<<<grid, threads, smem, mp_stream(stream)>>>(mi, mo, p);
  if (p.fitmap) {
    MP_LAUNCH_CHECK(who);
    // 8 channels per CTA like the plain gather kernel; CTAs of served tiles exit on their first load (one byte)
    const int shares = (C + GS_CCHUNK - 1) / GS_CCHUNK;
    dim3 g2((unsigned)(tiles / N), (unsigned)shares, (unsigned)N);
    if (mode == 0) k_gs_unfit<0><<<g2, 256, 0, mp_stream(stream)>>>(p, shares);
    else k_gs_unfit<1><<<g2, 256, 0, mp_stream(stream)>>>(p, shares);
  }
  MP_LAUNCH_CHECK(who);
  return 0;
}

}  // namespace

extern "C" int mp_gs_brick_tune(const int* cfg, int n) {
  for (int i = 0; i < 8; ++i) g_tune[i] = (cfg && i < n) ? cfg[i] : 0;
  return 0;
}

extern "C" size_t mp_gs_brick_workspace_bytes(int N, int Do, int Ho, int Wo) {
  int tz, ty, tx;
  tile_shape(Do, Ho, Wo, g_tune[6] ? g_tune[6] : 512, tz, ty, tx);
  return (size_t)N * ((Wo + tx - 1) / tx) * ((Ho + ty - 1) / ty) * ((Do + tz - 1) / tz);
}

extern "C" int mp_grid_sample3d_brick(const float* v, const float* grid, float* out, void* workspace, size_t workspace_bytes,
                                      int N, int C, int D, int H, int W, int Do, int Ho, int Wo, int flags, void* stream) {
  MP_REQUIRE(v && grid && out, "mp_grid_sample3d_brick: null pointer");
  MP_REQUIRE(N > 0 && N <= 65535 && C > 0 && D > 0 && H > 0 && W > 0 && Do > 0 && Ho > 0 && Wo > 0,
             "mp_grid_sample3d_brick: bad dims");
  return launch_brick(0, v, grid, out, workspace, workspace_bytes, N, C, D, H, W, Do, Ho, Wo, 0, 0, 0, flags, stream,
                      "mp_grid_sample3d_brick");
}

extern "C" int mp_apply_warping_field_brick(const float* v, const float* warp_field, float* out, void* workspace,
                                            size_t workspace_bytes, int N, int C, int D, int H, int W, int Df, int Hf, int Wf,
                                            int flags, void* stream) {
  MP_REQUIRE(v && warp_field && out, "mp_apply_warping_field_brick: null pointer");
  MP_REQUIRE(N > 0 && N <= 65535 && C > 0 && D > 1 && H > 1 && W > 1 && Df > 0 && Hf > 0 && Wf > 0,
             "mp_apply_warping_field_brick: bad dims");
  return launch_brick(1, v, warp_field, out, workspace, workspace_bytes, N, C, D, H, W, D, H, W, Df, Hf, Wf, flags, stream,
                      "mp_apply_warping_field_brick");
}
