// Implicit-GEMM convolution on 5th-gen tensor cores (tcgen05.mma, accumulators in TMEM) fed by TMA.  sm_100a only.
//
// GEMM view of a "same" convolution with stride 1 (nn.Conv2d / nn.Conv3d(k, padding=k//2)):
//     D[m, co] = sum_{tap, ci} A[m + shift(tap), ci] * Wt[co, tap, ci]        m = output position, fp32 accumulate
// Layout: activations channels-last [N, D, H, W, C] as TWO bf16 planes (hi, lo; x ~= hi + lo), weights packed
// [Cout_pad][taps*Cin] (K-major) as two bf16 planes.  fp32-grade products come from three bf16 MMAs per K step:
//     hi*hi + lo*hi + hi*lo      (the lo*lo term, ~2^-18 relative, is dropped)
//
// One CTA computes a 128-position x BN-channel tile.  The 128 positions are a (BD, BH, BW) box of the output
// grid, so for every filter tap the A operand is the same box shifted by (kd-pd, kh-ph, kw-pw): ONE 5-D TMA tile
// load per (tap, 64-channel chunk), with the zero padding supplied by TMA's out-of-bounds fill.  No im2col buffer
// ever exists in HBM.  The box lands in shared memory as 128 rows x (CCHUNK*2) bytes with the 128B/64B/32B swizzle
// that the UMMA shared-memory descriptor expects for a K-major operand.
//
// Warp roles (192 threads): warp 0 = TMA producer (one lane), warp 1 = TMEM allocator + MMA issuer (one lane),
// warps 2..5 = epilogue (TMEM -> registers -> bias/residual/activation/GroupNorm statistics -> HBM).
// Pipelines: STAGES-deep smem ring (full/empty mbarriers, tcgen05.commit frees a slot), one tmem_full barrier.
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>

#include <mutex>

#include "common.cuh"

int mp_conv_validate(const mp_conv_desc* d, const char* who);

namespace {

constexpr int TILE_M = 128;
constexpr int NUM_THREADS = 192;
constexpr uint32_t SMEM_LIMIT = 227 * 1024;

struct TcParams {
  int D, H, W, Cin, Cout;
  int KD, KH, KW;
  int BD, BH, BW;
  int tiles_d, tiles_h, tiles_w;
  int BN, CCHUNK, STAGES, num_cchunks;
  uint32_t layout_type;   // UMMA LayoutType: 2 = SWIZZLE_128B, 4 = SWIZZLE_64B, 6 = SWIZZLE_32B
  uint32_t sbo;           // bytes between 8-row groups
  uint32_t a_bytes, b_bytes, stage_bytes;
  uint32_t tmem_cols, idesc;
  const float* bias;
  const float* res_f32;
  const bf16 *res_hi, *res_lo;
  float* out_f32;
  bf16 *out_hi, *out_lo;
  double* stats;
  int gn_groups, act;
  int64_t S;
};

// ------------------------------------------------------------------------------------------------ PTX wrappers
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
    if (clock64() - t0 > 4000000000LL) {   // ~2 s: a protocol bug must not hang the GPU box
      printf("mp_conv_tc: mbarrier timeout (block %d,%d thread %d bar %u parity %u)\n", blockIdx.x, blockIdx.y,
             threadIdx.x, bar, parity);
      __trap();
    }
  }
}

__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}

__device__ __forceinline__ uint64_t make_smem_desc(uint32_t addr, uint32_t sbo, uint32_t layout_type) {
  // cute::UMMA::SmemDescriptor: start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48) | layout [61,64)
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;                       // LBO is ignored for swizzled K-major operands; canonical value 1
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout_type << 61;
  return d;
}

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float v[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// ------------------------------------------------------------------------------------------------ kernel
__global__ void __launch_bounds__(NUM_THREADS, 1)
k_conv_tc(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
          const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo, const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = smem_base + p.STAGES * p.stage_bytes;       // full[STAGES], empty[STAGES], tmem_full
  const uint32_t tmem_slot = bars + (2 * p.STAGES + 1) * 8;
  uint8_t* gen_base = smem_raw + (smem_base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen_base + (tmem_slot - smem_base));
  double* s_stats = reinterpret_cast<double*>(gen_base + (tmem_slot + 8 - smem_base));   // [BN][2]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  auto full_bar = [&](int s) { return bars + s * 8; };
  auto empty_bar = [&](int s) { return bars + (p.STAGES + s) * 8; };
  const uint32_t tmem_full_bar = bars + 2 * p.STAGES * 8;

  // tile coordinates
  int t = blockIdx.x;
  const int tw = t % p.tiles_w; t /= p.tiles_w;
  const int th = t % p.tiles_h; t /= p.tiles_h;
  const int td = t % p.tiles_d;
  const int n = t / p.tiles_d;
  const int w0 = tw * p.BW, h0 = th * p.BH, d0 = td * p.BD;
  const int n0 = blockIdx.y * p.BN;
  const int taps = p.KD * p.KH * p.KW;
  const int num_kb = taps * p.num_cchunks;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a_hi)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a_lo)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_b_hi)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_b_lo)) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (p.stats)
    for (int i = threadIdx.x; i < 2 * p.BN; i += NUM_THREADS) s_stats[i] = 0.0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (lane == 0) {
      const int pd = p.KD / 2, ph = p.KH / 2, pw = p.KW / 2;
      int kb = 0;
      for (int tap = 0; tap < taps; ++tap) {
        const int kw = tap % p.KW, kh = (tap / p.KW) % p.KH, kd = tap / (p.KW * p.KH);
        for (int cc = 0; cc < p.num_cchunks; ++cc, ++kb) {
          const int s = kb % p.STAGES;
          const uint32_t ph_bit = (kb / p.STAGES) & 1;
          mbar_wait(empty_bar(s), ph_bit ^ 1);
          const uint32_t sa = smem_base + s * p.stage_bytes;
          mbar_expect_tx(full_bar(s), 2 * p.a_bytes + 2 * p.b_bytes);
          const int c0 = cc * p.CCHUNK;
          tma_load_5d(sa, &map_a_hi, full_bar(s), c0, w0 + kw - pw, h0 + kh - ph, d0 + kd - pd, n);
          tma_load_5d(sa + p.a_bytes, &map_a_lo, full_bar(s), c0, w0 + kw - pw, h0 + kh - ph, d0 + kd - pd, n);
          tma_load_2d(sa + 2 * p.a_bytes, &map_b_hi, full_bar(s), tap * p.Cin + c0, n0);
          tma_load_2d(sa + 2 * p.a_bytes + p.b_bytes, &map_b_lo, full_bar(s), tap * p.Cin + c0, n0);
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    if (lane == 0) {
      const int ksteps = p.CCHUNK / 16;
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % p.STAGES;
        const uint32_t ph_bit = (kb / p.STAGES) & 1;
        mbar_wait(full_bar(s), ph_bit);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t sa = smem_base + s * p.stage_bytes;
        const uint32_t a_hi = sa, a_lo = sa + p.a_bytes, b_hi = sa + 2 * p.a_bytes, b_lo = b_hi + p.b_bytes;
        for (int k = 0; k < ksteps; ++k) {
          const uint32_t ko = k * 32;   // 16 bf16 along K inside the swizzle atom
          const uint64_t dah = make_smem_desc(a_hi + ko, p.sbo, p.layout_type);
          const uint64_t dal = make_smem_desc(a_lo + ko, p.sbo, p.layout_type);
          const uint64_t dbh = make_smem_desc(b_hi + ko, p.sbo, p.layout_type);
          const uint64_t dbl = make_smem_desc(b_lo + ko, p.sbo, p.layout_type);
          umma_bf16(tmem_base, dal, dbh, p.idesc, (kb > 0 || k > 0) ? 1u : 0u);
          umma_bf16(tmem_base, dah, dbl, p.idesc, 1u);
          umma_bf16(tmem_base, dah, dbh, p.idesc, 1u);
        }
        umma_commit(empty_bar(s));   // implicit tcgen05.fence::before_thread_sync
      }
      umma_commit(tmem_full_bar);
    }
  } else {
    // ===================================================================== epilogue (warps 2..5)
    const int q = warp & 3;                    // TMEM lane quarter this warp may access
    const int r = q * 32 + lane;               // accumulator row == position inside the tile
    const int ww = r % p.BW, hh = (r / p.BW) % p.BH, dd = r / (p.BW * p.BH);
    const int64_t pos = (((int64_t)n * p.D + d0 + dd) * p.H + h0 + hh) * p.W + w0 + ww;
    const int64_t obase = pos * p.Cout;
    const int cpg = p.gn_groups > 0 ? p.Cout / p.gn_groups : 1;
    const bool vec4 = (p.Cout % 4) == 0, vec8 = (p.Cout % 8) == 0;
    mbar_wait(tmem_full_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < p.BN; c0 += 16) {
      float v[16];
      tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
      const int co0 = n0 + c0;
      if (co0 >= p.Cout) break;                // padded output channels (warp-uniform)
      const bool full16 = co0 + 16 <= p.Cout;
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (p.bias && co0 + i < p.Cout) v[i] += __ldg(p.bias + co0 + i);
      if (p.res_f32) {
        if (full16 && vec4) {
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            float4 rr = *reinterpret_cast<const float4*>(p.res_f32 + obase + co0 + i);
            v[i] += rr.x; v[i + 1] += rr.y; v[i + 2] += rr.z; v[i + 3] += rr.w;
          }
        } else {
          for (int i = 0; i < 16; ++i)
            if (co0 + i < p.Cout) v[i] += p.res_f32[obase + co0 + i];
        }
      } else if (p.res_hi) {
        if (full16 && vec4) {
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            float4 rr = mp_load_split4(p.res_hi, p.res_lo, obase + co0 + i);
            v[i] += rr.x; v[i + 1] += rr.y; v[i + 2] += rr.z; v[i + 3] += rr.w;
          }
        } else {
          for (int i = 0; i < 16; ++i)
            if (co0 + i < p.Cout) v[i] += mp_join(p.res_hi[obase + co0 + i], p.res_lo[obase + co0 + i]);
        }
      }
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = mp_apply_act(v[i], p.act);
      if (p.out_f32) {
        if (full16 && vec4) {
#pragma unroll
          for (int i = 0; i < 16; i += 4)
            *reinterpret_cast<float4*>(p.out_f32 + obase + co0 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        } else {
          for (int i = 0; i < 16; ++i)
            if (co0 + i < p.Cout) p.out_f32[obase + co0 + i] = v[i];
        }
      }
      if (p.out_hi) {
        if (full16 && vec8) {
#pragma unroll
          for (int i = 0; i < 16; i += 4) mp_store_split4(p.out_hi, p.out_lo, obase + co0 + i, make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]));
        } else {
          for (int i = 0; i < 16; ++i)
            if (co0 + i < p.Cout) mp_split2(v[i], p.out_hi[obase + co0 + i], p.out_lo[obase + co0 + i]);
        }
      }
      if (p.stats) {
        // per-column sums over the 32 rows of this warp, then one shared-memory add per column
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float sm = (co0 + i < p.Cout) ? v[i] : 0.f;
          float sq = sm * sm;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            sm += __shfl_xor_sync(0xffffffffu, sm, o);
            sq += __shfl_xor_sync(0xffffffffu, sq, o);
          }
          if (lane == 0) {
            atomicAdd(&s_stats[2 * (c0 + i)], (double)sm);
            atomicAdd(&s_stats[2 * (c0 + i) + 1], (double)sq);
          }
        }
      }
    }
    if (p.stats) {
      asm volatile("bar.sync 1, 128;" ::: "memory");   // the four epilogue warps only
      const int et = threadIdx.x - 64;
      for (int c = et; c < p.BN; c += 128) {
        const int co = n0 + c;
        if (co < p.Cout) {
          double* st = p.stats + ((int64_t)n * p.gn_groups + co / cpg) * 2;
          atomicAdd(st, s_stats[2 * c]);
          atomicAdd(st + 1, s_stats[2 * c + 1]);
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encoder() {
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

struct Plan {
  TcParams p;
  int tiles_m, tiles_n;
  uint32_t smem_bytes;
  CUtensorMapSwizzle swz;
};

int next_pow2(int v) { int r = 32; while (r < v) r <<= 1; return r; }

// returns 0 and fills plan when the shape is supported
int make_plan(const mp_conv_desc* d, Plan& pl, bool report) {
  TcParams& p = pl.p;
  auto fail = [&](const char* why) { return report ? mp_set_error("mp_conv_tc: unsupported shape: %s", why) : 1; };
  if (d->Cin % 16 != 0) return fail("Cin % 16 != 0");
  if (d->Cout_pad % 16 != 0) return fail("Cout_pad % 16 != 0");
  p.CCHUNK = (d->Cin % 64 == 0) ? 64 : (d->Cin % 32 == 0) ? 32 : 16;
  p.layout_type = p.CCHUNK == 64 ? 2u : p.CCHUNK == 32 ? 4u : 6u;
  pl.swz = p.CCHUNK == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : p.CCHUNK == 32 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                                         : CU_TENSOR_MAP_SWIZZLE_32B;
  p.sbo = 8u * p.CCHUNK * 2u;
  p.num_cchunks = d->Cin / p.CCHUNK;
  // output tile box
  auto pow2_le = [](int v, int cap) { int r = 1; while (r * 2 <= v && r * 2 <= cap) r *= 2; return r; };
  p.BW = pow2_le(d->W, TILE_M);
  p.BH = pow2_le(d->H, TILE_M / p.BW);
  p.BD = pow2_le(d->D, TILE_M / (p.BW * p.BH));
  if (p.BW * p.BH * p.BD != TILE_M) return fail("fewer than 128 positions per sample box");
  if (d->W % p.BW || d->H % p.BH || d->D % p.BD) return fail("grid not divisible by the tile box");
  p.tiles_w = d->W / p.BW; p.tiles_h = d->H / p.BH; p.tiles_d = d->D / p.BD;
  pl.tiles_m = d->N * p.tiles_d * p.tiles_h * p.tiles_w;
  // N tile: largest multiple-of-16 divisor of Cout_pad up to the cap, shrunk while the grid under-fills the GPU
  static int bn_cap = [] { const char* e = getenv("MPB200_TC_BN_MAX"); int v = e ? atoi(e) : 128; return v < 16 ? 16 : (v > 256 ? 256 : v); }();
  const int cands[] = {256, 192, 128, 96, 64, 48, 32, 16};
  p.BN = 0;
  for (int c : cands)
    if (c <= bn_cap && d->Cout_pad % c == 0) { p.BN = c; break; }
  if (!p.BN) return fail("no N tile");
  while ((int64_t)pl.tiles_m * (d->Cout_pad / p.BN) < 148 && p.BN % 32 == 0 && p.BN > 32) p.BN /= 2;
  pl.tiles_n = d->Cout_pad / p.BN;
  p.a_bytes = TILE_M * p.CCHUNK * 2;
  p.b_bytes = p.BN * p.CCHUNK * 2;
  p.stage_bytes = 2 * p.a_bytes + 2 * p.b_bytes;
  const uint32_t fixed = 1024 /*align slack*/ + 1024 /*barriers, tmem slot*/ + 2 * 256 * sizeof(double);
  int stages = (int)((SMEM_LIMIT - fixed) / p.stage_bytes);
  if (stages > 8) stages = 8;
  if (stages < 2) return fail("tile does not fit shared memory");
  p.STAGES = stages;
  pl.smem_bytes = fixed + stages * p.stage_bytes;
  p.tmem_cols = next_pow2(p.BN);
  p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
  p.D = d->D; p.H = d->H; p.W = d->W; p.Cin = d->Cin; p.Cout = d->Cout;
  p.KD = d->KD; p.KH = d->KH; p.KW = d->KW;
  p.bias = d->bias; p.res_f32 = d->res_f32; p.res_hi = (const bf16*)d->res_hi; p.res_lo = (const bf16*)d->res_lo;
  p.out_f32 = d->out_f32; p.out_hi = (bf16*)d->out_hi; p.out_lo = (bf16*)d->out_lo;
  p.stats = d->stats; p.gn_groups = d->gn_groups; p.act = d->act;
  p.S = (int64_t)d->D * d->H * d->W;
  if (pl.tiles_m > 0x7fffffff) return fail("too many tiles");
  return 0;
}

int encode_act_map(CUtensorMap* m, const void* ptr, const mp_conv_desc* d, const Plan& pl) {
  cuuint64_t dims[5] = {(cuuint64_t)d->Cin, (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->D, (cuuint64_t)d->N};
  cuuint64_t strides[4] = {(cuuint64_t)d->Cin * 2, (cuuint64_t)d->Cin * 2 * d->W, (cuuint64_t)d->Cin * 2 * d->W * d->H,
                           (cuuint64_t)d->Cin * 2 * d->W * d->H * d->D};
  cuuint32_t box[5] = {(cuuint32_t)pl.p.CCHUNK, (cuuint32_t)pl.p.BW, (cuuint32_t)pl.p.BH, (cuuint32_t)pl.p.BD, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = get_encoder()(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(ptr), dims, strides, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, pl.swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : mp_set_error("mp_conv_tc: cuTensorMapEncodeTiled(activation) failed (%d)", (int)r);
}

int encode_w_map(CUtensorMap* m, const void* ptr, const mp_conv_desc* d, const Plan& pl) {
  const cuuint64_t ktot = (cuuint64_t)d->KD * d->KH * d->KW * d->Cin;
  cuuint64_t dims[2] = {ktot, (cuuint64_t)d->Cout_pad};
  cuuint64_t strides[1] = {ktot * 2};
  cuuint32_t box[2] = {(cuuint32_t)pl.p.CCHUNK, (cuuint32_t)pl.p.BN};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = get_encoder()(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, pl.swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : mp_set_error("mp_conv_tc: cuTensorMapEncodeTiled(weights) failed (%d)", (int)r);
}

}  // namespace

extern "C" int mp_conv_tc_supported(const mp_conv_desc* d) {
  if (!d) return 0;
  Plan pl;
  return make_plan(d, pl, false) == 0 ? 1 : 0;
}

extern "C" int mp_conv_tc(const mp_conv_desc* d, void* stream) {
  if (int e = mp_conv_validate(d, "mp_conv_tc")) return e;
  MP_REQUIRE(get_encoder() != nullptr, "mp_conv_tc: cuTensorMapEncodeTiled not available from the driver");
  Plan pl;
  if (int e = make_plan(d, pl, true)) return e;
  MP_REQUIRE((((uintptr_t)d->in_hi | (uintptr_t)d->in_lo | (uintptr_t)d->w_hi | (uintptr_t)d->w_lo) & 15) == 0,
             "mp_conv_tc: operands must be 16-byte aligned");
  CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo;
  if (int e = encode_act_map(&ma_hi, d->in_hi, d, pl)) return e;
  if (int e = encode_act_map(&ma_lo, d->in_lo, d, pl)) return e;
  if (int e = encode_w_map(&mb_hi, d->w_hi, d, pl)) return e;
  if (int e = encode_w_map(&mb_lo, d->w_lo, d, pl)) return e;
  {
    static bool attr_done[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !attr_done[dev]) {
      cudaError_t ae = cudaFuncSetAttribute(k_conv_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT);
      MP_REQUIRE(ae == cudaSuccess, "mp_conv_tc: cannot opt in to %u B shared memory: %s", SMEM_LIMIT,
                 cudaGetErrorString(ae));
      if (dev >= 0 && dev < 64) attr_done[dev] = true;
    }
  }
  dim3 grid((unsigned)pl.tiles_m, (unsigned)pl.tiles_n);
  k_conv_tc<<<grid, NUM_THREADS, pl.smem_bytes, mp_stream(stream)>>>(ma_hi, ma_lo, mb_hi, mb_lo, pl.p);
  MP_LAUNCH_CHECK("mp_conv_tc");
  return 0;
}
