// Implicit-GEMM convolution on 5th-gen tensor cores (tcgen05.mma, accumulators in TMEM) fed by TMA.  sm_100a only.
//
// GEMM view of a "same" convolution (nn.Conv2d / nn.Conv3d(k, padding=k//2), stride 1 or 2 on H/W):
//     D[m, co] = sum_{tap, ci} A[m + shift(tap), ci] * Wt[co, tap, ci]        m = output position, fp32 accumulate
// Layout: activations channels-last [N, D, H, W, C], weights packed [Cout_pad][taps*Cin (+ Cin2)] (K-major).  Three
// operand formats (mp_conv_desc.prec):
//     split-bf16  two bf16 planes (hi, lo); products hi*hi + lo*hi + hi*lo, three MMA passes (lo*lo ~2^-18 dropped);
//     fp16 x2     one fp16 activation plane, fp16 hi + scaled fp16 lo weights: ONE MMA A x [Wh | Wl'] per K step;
//     fp16 + q8   fp16 x fp16 main product (kind::f16) + both cross terms as e4m3 x e4m3 (kind::f8f6f4) in a second
//                 accumulator, issued by two warps.
//
// A CTA computes 128-position x BN-channel tiles.  The 128 positions are a (BD, BH, BW) box of the output grid, so for
// every filter tap the A operand is the same box shifted by (kd-pd, kh-ph, kw-pw): ONE 5-D TMA tile load per (tap,
// channel chunk), zero padding supplied by TMA's out-of-bounds fill, stride 2 by TMA element strides.  No im2col buffer
// ever exists in HBM.  The box lands in shared memory as 128 rows x (CCHUNK*2) bytes with the 128B/64B/32B swizzle that
// the UMMA shared-memory descriptor expects for a K-major operand.  An optional second source (fused 1x1 shortcut)
// appends Cin2/CCHUNK K-chunks whose A tiles come from another tensor map.
//
// Kernels: k_conv_tc2 (general: persistent, BN up to 256, smem ring of 2-8 stages, optional resident weights),
// k_conv_tc3 (slab: 3x3(x3) with Cout <= 128, one (MT*16+2) x 8 pixel slab serves the three vertical taps of MT
// accumulators), k_conv_q8_pair (cta_group::2 pairs for the fp16 + FP8 format).  Warp roles in the persistent kernels (352 threads):
// warp 0 = TMA producer + tile scheduler (tile ids drawn from a global atomic counter, published to the other roles
// through a shared-memory queue), warp 1 (+ warp 10 when a tile has two independently issued accumulators) = MMA issue,
// whole warp converged with one elected lane issuing, warps 2..9 = epilogue (TMEM -> registers -> bias / residual /
// activation / GroupNorm statistics -> shared-memory staging -> coalesced stores, or swizzled staging + TMA bulk stores
// for fp16 outputs).  TMEM holds two tiles' accumulators so the epilogue of tile i overlaps the MMAs of tile i+1.
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <unordered_map>

#include "common.cuh"

int mp_conv_validate(const mp_conv_desc* d, const char* who);

namespace {

constexpr int TILE_M = 128;
constexpr int NUM_THREADS2 = 320;   // v2: producer + MMA + 8 epilogue warps
constexpr int NUM_THREADS3 = 352;   // v3 (slab): + a second MMA-issuing warp (one accumulator each when MT = 2)
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
  int tiles_n, total_tiles;   // persistent kernel: tile id = m_tile * tiles_n + n_tile (n fastest => A shared in L2)
  uint32_t epi_off;           // byte offset (from the aligned smem base) of the epilogue staging area
  int stride, in_c_off, out_C, out_c_off;   // ABI v2: H/W stride, channel windows (D/H/W below are OUTPUT dims)
  int N, BNb;                 // batch size; samples per tile box (> 1 when one sample has fewer than 128 positions)
  unsigned long long tap_mask;   // bit t set: filter tap t overlaps the input for at least one position (<= 64 taps)
  int dualb, acc_w;           // dual-B mode: D[:, 0:BN] += Al*Bh ; D[:, 0:2BN] += Ah*[Bh|Bl]; acc_w = TMEM columns per acc
  uint32_t idesc2;            // instruction descriptor with N = 2*BN
  int b_resident;             // 1: the whole weight tile [taps*Cin x BN] stays in smem for the CTA's lifetime
  uint32_t bres_off;          // byte offset of the resident weight area
  int prec;                   // MP_PREC_SPLIT_BF16 | MP_PREC_F16X2 (single fp16 activation plane, fp16 hi / scaled-lo weights)
  uint32_t a_planes;          // activation planes per stage: 2 (hi, lo) or 1 (fp16)
  float lo_scale;             // weight of the [Bl] column half in the epilogue: 1 (bf16 split) or 2^-11 (fp16 scaled lo)
  int epi_tma;                // 1: fp16 output tile staged in swizzled shared memory and written by TMA bulk stores
  uint32_t tma_off;           // byte offset (1024-aligned) of the two TMA staging buffers
  uint32_t tma_buf_bytes;     // one staging buffer: (BN / 64) sub-tiles of [128 rows][64 fp16] = 16 KB each
  uint32_t tma_nbuf;          // 1 or 2 staging buffers
  uint32_t epi_bytes;         // size of the register-store staging area (0 when the TMA-store epilogue is used)
  int num_cchunks2, k2_off;   // fused 1x1 shortcut: channel chunks of the second source, K offset of its weights
  int stride2, in2_c_off;
  int out_fmt, res_fmt;       // effective MP_FMT_* of (out_hi, out_lo) and (res_hi, res_lo)
  int b_merged;               // 1: map_b_hi is a 3-D map (K, Cout_pad, plane) and one TMA load fetches [Bh | Bl]
  float acc_scale;            // fp16 weight planes were packed from w / acc_scale (a power of two)
  float out_q8_scale, res_q8_inv;   // F16_Q8 byte planes: scale written / 1 / (2048 * scale of the residual plane)
  unsigned int* sched;        // dynamic tile scheduler: {next-tile counter, finished-CTA counter}, or NULL = static walk
};

constexpr int SQ = 4;         // depth of the in-CTA tile-id queue (producer -> MMA issuer + epilogue warps)

// Global scheduler slots: both words are zero between launches (the last CTA of a launch resets them).  Two launches that
// may run at the same time never share a slot (sched_slot() below): eager launches rotate through a range private to
// their stream, launches recorded into a CUDA graph get a slot that nothing else will ever use.
constexpr unsigned SCHED_SLOTS = 65536, SCHED_EAGER = 32768, SCHED_RANGE = 2048;
__device__ unsigned int g_sched[SCHED_SLOTS][2];

struct __align__(16) bf16x8 {
  bf16 v[8];
};

constexpr int EPI_COLS = 32;                 // epilogue column chunk
constexpr int EPI_PITCH = EPI_COLS + 4;      // floats per staged row (+4: conflict-free 16-byte row writes)
constexpr uint32_t EPI_BYTES = TILE_M * EPI_PITCH * 4 + TILE_M * 8;   // staging tile + per-row output offsets

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

// ---- TMA bulk store (shared -> global) of a 5-D box, bulk-group completion
__device__ __forceinline__ void tma_store_5d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// weight tile(s) of one K chunk: [Bh | Bl] in one 3-D load when the two planes are one allocation, else two 2-D loads
__device__ __forceinline__ void load_b(const TcParams& p, uint32_t dst, const CUtensorMap* mh, const CUtensorMap* ml,
                                       uint32_t bar, int k0, int n0) {
  if (p.b_merged) {
    tma_load_3d(dst, mh, bar, k0, n0, 0);
  } else {
    tma_load_2d(dst, mh, bar, k0, n0);
    tma_load_2d(dst + p.b_bytes, ml, bar, k0, n0);
  }
}

// UMMA shared-memory descriptor (cute::UMMA::SmemDescriptor): start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1
// [46,48) | layout [61,64).  The upper word (SBO | version | layout) is constant per kernel, the lower word is
// (smem address >> 4) | LBO; stepping 16 bf16 along K inside the swizzle atom adds 2 to the lower word.
__device__ __forceinline__ uint32_t desc_hi_word(uint32_t sbo, uint32_t layout_type) {
  return ((sbo >> 4) & 0x3FFFu) | (1u << 14) | (layout_type << 29);
}
__device__ __forceinline__ uint32_t desc_lo_word(uint32_t addr) { return ((addr & 0x3FFFFu) >> 4) | (1u << 16); }
// Issue paths of the persistent kernels: called by ALL 32 lanes of the MMA warp in lock-step; `elect.sync` picks the one
// lane that issues.  (Issuing from inside an `if (lane == 0)` region instead makes the compiler wrap every UTCHMMA /
// UTCBAR in an active-lane loop -- ELECT, PLOP3, BRA.U.ANY -- which, on a single dependent instruction stream, capped
// the 64-cycle N = 128 MMAs at ~37 % tensor-pipe utilisation: ncu, profiles/r2_*.)
__device__ __forceinline__ void umma_lean_e(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}"
      ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_e(uint32_t bar) {
  asm volatile(
      "{\n\t.reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar)
      : "memory");
}
// FP8 (e4m3 x e4m3) MMA: K = 32 bytes per instruction, i.e. the same 32-byte descriptor step as a K = 16 fp16 MMA
__device__ __forceinline__ void umma_f8_e(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], da, db, %4, p;\n\t}"
      ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// cross terms of one 64-channel chunk: A bytes [x8 | xl8] (128 per row) against B bytes [wl8 | w8]: four K = 32 MMAs
template <int KS>
__device__ __forceinline__ void umma_chunk_q8(uint32_t tmem_d, uint32_t a_q, uint32_t b_q, uint32_t hi, uint32_t idesc,
                                              uint32_t acc_first) {
  const uint32_t aq = desc_lo_word(a_q), bq = desc_lo_word(b_q);
#pragma unroll
  for (int k = 0; k < KS; ++k) umma_f8_e(tmem_d, aq + 2 * k, bq + 2 * k, hi, idesc, k == 0 ? acc_first : 1u);
}
// one K-chunk (KS k-steps) of the 3-pass split product into accumulator tmem_d
template <int KS>
__device__ __forceinline__ void umma_chunk(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo,
                                           uint32_t hi, uint32_t idesc, uint32_t acc_first) {
  const uint32_t ah = desc_lo_word(a_hi), al = desc_lo_word(a_lo), bh = desc_lo_word(b_hi), bl = desc_lo_word(b_lo);
#pragma unroll
  for (int k = 0; k < KS; ++k) {
    umma_lean_e(tmem_d, al + 2 * k, bh + 2 * k, hi, idesc, k == 0 ? acc_first : 1u);
    umma_lean_e(tmem_d, ah + 2 * k, bl + 2 * k, hi, idesc, 1u);
    umma_lean_e(tmem_d, ah + 2 * k, bh + 2 * k, hi, idesc, 1u);
  }
}
// dual-B variant: two MMAs per K step -- Ah x [Bh|Bl] (N = 2*BN, Bl follows Bh in shared memory) and Al x Bh (N = BN)
template <int KS>
__device__ __forceinline__ void umma_chunk_dual(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi,
                                                uint32_t hi, uint32_t idesc, uint32_t idesc2, uint32_t acc_first) {
  const uint32_t ah = desc_lo_word(a_hi), al = desc_lo_word(a_lo), bh = desc_lo_word(b_hi);
#pragma unroll
  for (int k = 0; k < KS; ++k) {
    umma_lean_e(tmem_d, ah + 2 * k, bh + 2 * k, hi, idesc2, k == 0 ? acc_first : 1u);
    umma_lean_e(tmem_d, al + 2 * k, bh + 2 * k, hi, idesc, 1u);
  }
}
// fp16 two-pass variant: ONE MMA per K step, Ah x [Bh|Bl'] (N = 2*BN); the activation has no lo plane
template <int KS>
__device__ __forceinline__ void umma_chunk_h(uint32_t tmem_d, uint32_t a_hi, uint32_t b_hi, uint32_t hi, uint32_t idesc2,
                                             uint32_t acc_first) {
  const uint32_t ah = desc_lo_word(a_hi), bh = desc_lo_word(b_hi);
#pragma unroll
  for (int k = 0; k < KS; ++k) umma_lean_e(tmem_d, ah + 2 * k, bh + 2 * k, hi, idesc2, k == 0 ? acc_first : 1u);
}
__device__ __forceinline__ void umma_chunk_h_dyn(int ksteps, uint32_t tmem_d, uint32_t a_hi, uint32_t b_hi, uint32_t hi,
                                                 uint32_t idesc2, uint32_t acc_first) {
  if (ksteps == 4) umma_chunk_h<4>(tmem_d, a_hi, b_hi, hi, idesc2, acc_first);
  else if (ksteps == 2) umma_chunk_h<2>(tmem_d, a_hi, b_hi, hi, idesc2, acc_first);
  else umma_chunk_h<1>(tmem_d, a_hi, b_hi, hi, idesc2, acc_first);
}
__device__ __forceinline__ void umma_chunk_dual_dyn(int ksteps, uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo,
                                                    uint32_t b_hi, uint32_t hi, uint32_t idesc, uint32_t idesc2,
                                                    uint32_t acc_first) {
  if (ksteps == 4) umma_chunk_dual<4>(tmem_d, a_hi, a_lo, b_hi, hi, idesc, idesc2, acc_first);
  else if (ksteps == 2) umma_chunk_dual<2>(tmem_d, a_hi, a_lo, b_hi, hi, idesc, idesc2, acc_first);
  else umma_chunk_dual<1>(tmem_d, a_hi, a_lo, b_hi, hi, idesc, idesc2, acc_first);
}
__device__ __forceinline__ void umma_chunk_dyn(int ksteps, uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi,
                                               uint32_t b_lo, uint32_t hi, uint32_t idesc, uint32_t acc_first) {
  if (ksteps == 4) umma_chunk<4>(tmem_d, a_hi, a_lo, b_hi, b_lo, hi, idesc, acc_first);
  else if (ksteps == 2) umma_chunk<2>(tmem_d, a_hi, a_lo, b_hi, b_lo, hi, idesc, acc_first);
  else umma_chunk<1>(tmem_d, a_hi, a_lo, b_hi, b_lo, hi, idesc, acc_first);
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

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float v[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// Drain one 128 x BN accumulator (TMEM columns tmem_acc .. +BN of this warp's lane quarter) through the staging
// buffer: bias + residual + activation, warp-contiguous stores, per-column GroupNorm partial sums into s_stats.
// Called by all 256 epilogue threads; contains CTA-level named barriers.
__device__ __forceinline__ void epilogue_drain(const TcParams& p, float* epi, long long* row_off, double* s_stats,
                                               uint32_t tmem_acc, int n0, int64_t obase, int half, int r, int et,
                                               int lane, bool row_valid = true) {
  const bool vec4 = ((p.out_C | p.out_c_off | p.Cout) % 4) == 0, vec8 = ((p.out_C | p.out_c_off | p.Cout) % 8) == 0;
  const bool bias_vec = p.bias && ((reinterpret_cast<uintptr_t>(p.bias) & 15) == 0);
  if (half == 0) row_off[r] = row_valid ? obase : -1;   // rows of samples beyond the batch are never stored
  for (int c0 = 0; c0 < p.BN; c0 += EPI_COLS) {
    const int co0 = n0 + c0;
    if (co0 >= p.Cout) break;              // padded output channels (uniform across the CTA)
    const int ncol = min(min(EPI_COLS, p.BN - c0), p.Cout - co0);   // valid columns in this chunk (uniform)
    const int hc = half * 16;              // my first column inside the chunk
    const bool mine = hc < ncol;           // warp-uniform
    const bool fast = (ncol == EPI_COLS) && vec4;
    float v[16];
    if (mine) {
      tmem_ld16(tmem_acc + (uint32_t)(c0 + hc), v);
      if (p.dualb) {                           // Ah*Bl partial products live BN columns further
        float v2[16];
        tmem_ld16(tmem_acc + (uint32_t)(p.BN + c0 + hc), v2);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = fmaf(v2[i], p.lo_scale, v[i]) * p.acc_scale;
      }
      if (!row_valid) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = 0.f;
      } else if (fast) {
        if (bias_vec) {
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const float4 bb = __ldg(reinterpret_cast<const float4*>(p.bias + co0 + hc + i));
            v[i] += bb.x; v[i + 1] += bb.y; v[i + 2] += bb.z; v[i + 3] += bb.w;
          }
        } else if (p.bias) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] += __ldg(p.bias + co0 + hc + i);
        }
        if (p.res_f32) {
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const float4 rr = *reinterpret_cast<const float4*>(p.res_f32 + obase + co0 + hc + i);
            v[i] += rr.x; v[i + 1] += rr.y; v[i + 2] += rr.z; v[i + 3] += rr.w;
          }
        } else if (p.res_hi && p.res_fmt != MP_FMT_SPLIT_BF16) {
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const f16x4 rr = *reinterpret_cast<const f16x4*>(reinterpret_cast<const f16*>(p.res_hi) + obase + co0 + hc + i);
            v[i] += __half2float(rr.v[0]); v[i + 1] += __half2float(rr.v[1]);
            v[i + 2] += __half2float(rr.v[2]); v[i + 3] += __half2float(rr.v[3]);
          }
          if (p.res_fmt == MP_FMT_F16_Q8) {     // x = fp16 plane + e4m3 low part / 2048 (16 channels = 16 bytes)
            const int ch = p.out_c_off + co0 + hc;
            const uint4 raw = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(p.res_lo) +
                                                              2 * (obase - p.out_c_off) + mp_q8_off(ch) + 64);
            const uint8_t* b8 = reinterpret_cast<const uint8_t*>(&raw);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = fmaf(mp_e4m3_to_float(b8[i]), p.res_q8_inv, v[i]);
          }
        } else if (p.res_hi) {
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const float4 rr = mp_load_split4(p.res_hi, p.res_lo, obase + co0 + hc + i);
            v[i] += rr.x; v[i + 1] += rr.y; v[i + 2] += rr.z; v[i + 3] += rr.w;
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          if (hc + i < ncol) {
            const int64_t o = obase + co0 + hc + i;
            if (p.bias) v[i] += __ldg(p.bias + co0 + hc + i);
            if (p.res_f32) v[i] += p.res_f32[o];
            else if (p.res_hi && p.res_fmt == MP_FMT_F16) v[i] += __half2float(reinterpret_cast<const f16*>(p.res_hi)[o]);
            else if (p.res_hi) v[i] += mp_join(p.res_hi[o], p.res_lo[o]);
          }
        }
      }
      if (!row_valid) {
      } else if (p.act == MP_ACT_RELU) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
      } else if (p.act != MP_ACT_NONE) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = mp_apply_act(v[i], p.act);
      }
      // stage my half row (16-byte stores; row pitch 36 floats => conflict-free across the warp)
#pragma unroll
      for (int i = 0; i < 16; i += 4)
        *reinterpret_cast<float4*>(epi + r * EPI_PITCH + hc + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
    }
    epi_bar();
    // ---- write-out: warp-contiguous rows (256 threads)
    if (p.out_f32) {
      if (vec4) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int row = j * 32 + (et >> 3), cq = (et & 7) * 4;
          if (cq < ncol && row_off[row] >= 0)
            *reinterpret_cast<float4*>(p.out_f32 + row_off[row] + co0 + cq) =
                *reinterpret_cast<const float4*>(epi + row * EPI_PITCH + cq);
        }
      } else {
        for (int j = 0; j < 16; ++j) {
          const int row = j * 8 + (et >> 5);
          if (lane < ncol && row_off[row] >= 0) p.out_f32[row_off[row] + co0 + lane] = epi[row * EPI_PITCH + lane];
        }
      }
    }
    if (p.out_hi && p.out_fmt == MP_FMT_F16_Q8) {
      // fp16 plane + byte plane (vec8 is guaranteed: channel counts are multiples of 64)
      f16* oh = reinterpret_cast<f16*>(p.out_hi);
      uint8_t* oq = reinterpret_cast<uint8_t*>(p.out_lo);
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int row = j * 64 + (et >> 2), c8 = (et & 3) * 8;
        if (c8 < ncol && row_off[row] >= 0) {
          const float4 x0 = *reinterpret_cast<const float4*>(epi + row * EPI_PITCH + c8);
          const float4 x1 = *reinterpret_cast<const float4*>(epi + row * EPI_PITCH + c8 + 4);
          const float xs[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
          f16x8 h;
          uint2 a8, al8;
          mp_hq_pack8(xs, h, a8, al8, p.out_q8_scale);
          *reinterpret_cast<f16x8*>(oh + row_off[row] + co0 + c8) = h;
          uint8_t* q = oq + 2 * (row_off[row] - p.out_c_off) + mp_q8_off(p.out_c_off + co0 + c8);
          *reinterpret_cast<uint2*>(q) = a8;
          *reinterpret_cast<uint2*>(q + 64) = al8;
        }
      }
    } else if (p.out_hi && p.out_fmt == MP_FMT_F16) {
      f16* oh = reinterpret_cast<f16*>(p.out_hi);
      if (vec8) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int row = j * 64 + (et >> 2), c8 = (et & 3) * 8;
          if (c8 < ncol && row_off[row] >= 0) {
            const float4 x0 = *reinterpret_cast<const float4*>(epi + row * EPI_PITCH + c8);
            const float4 x1 = *reinterpret_cast<const float4*>(epi + row * EPI_PITCH + c8 + 4);
            f16x8 h;
            h.v[0] = mp_to_f16(x0.x); h.v[1] = mp_to_f16(x0.y); h.v[2] = mp_to_f16(x0.z); h.v[3] = mp_to_f16(x0.w);
            h.v[4] = mp_to_f16(x1.x); h.v[5] = mp_to_f16(x1.y); h.v[6] = mp_to_f16(x1.z); h.v[7] = mp_to_f16(x1.w);
            *reinterpret_cast<f16x8*>(oh + row_off[row] + co0 + c8) = h;
          }
        }
      } else {
        for (int j = 0; j < 16; ++j) {
          const int row = j * 8 + (et >> 5);
          if (lane < ncol && row_off[row] >= 0) oh[row_off[row] + co0 + lane] = mp_to_f16(epi[row * EPI_PITCH + lane]);
        }
      }
    } else if (p.out_hi) {
      if (vec8) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int row = j * 64 + (et >> 2), c8 = (et & 3) * 8;
          if (c8 < ncol && row_off[row] >= 0) {
            const float4 x0 = *reinterpret_cast<const float4*>(epi + row * EPI_PITCH + c8);
            const float4 x1 = *reinterpret_cast<const float4*>(epi + row * EPI_PITCH + c8 + 4);
            const int64_t o = row_off[row] + co0 + c8;
            bf16x8 h, l;
            mp_split2(x0.x, h.v[0], l.v[0]); mp_split2(x0.y, h.v[1], l.v[1]); mp_split2(x0.z, h.v[2], l.v[2]);
            mp_split2(x0.w, h.v[3], l.v[3]); mp_split2(x1.x, h.v[4], l.v[4]); mp_split2(x1.y, h.v[5], l.v[5]);
            mp_split2(x1.z, h.v[6], l.v[6]); mp_split2(x1.w, h.v[7], l.v[7]);
            *reinterpret_cast<bf16x8*>(p.out_hi + o) = h;
            *reinterpret_cast<bf16x8*>(p.out_lo + o) = l;
          }
        }
      } else {
        for (int j = 0; j < 16; ++j) {
          const int row = j * 8 + (et >> 5);
          if (lane < ncol && row_off[row] >= 0) {
            const int64_t o = row_off[row] + co0 + lane;
            mp_split2(epi[row * EPI_PITCH + lane], p.out_hi[o], p.out_lo[o]);
          }
        }
      }
    }
    if (p.stats) {
      // column sums over 16-row slabs: thread (slab = et / 32, col = lane) -> s_part[slab][col]; after the barrier one
      // warp adds the 8 slabs in a fixed order (no atomics => bit-reproducible statistics)
      float* s_part = reinterpret_cast<float*>(s_stats + 2 * 256);
      const int slab = et >> 5;
      float sm = 0.f, sq = 0.f;
      if (lane < ncol) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float x = epi[(slab * 16 + j) * EPI_PITCH + lane];
          sm += x; sq += x * x;
        }
      }
      s_part[(slab * 32 + lane) * 2] = sm;
      s_part[(slab * 32 + lane) * 2 + 1] = sq;
    }
    epi_bar();                              // staging buffer may be overwritten
    if (p.stats && et < 32 && lane < ncol) {
      const float* s_part = reinterpret_cast<const float*>(s_stats + 2 * 256);
      double sm = 0.0, sq = 0.0;
#pragma unroll
      for (int sl = 0; sl < 8; ++sl) {
        sm += (double)s_part[(sl * 32 + lane) * 2];
        sq += (double)s_part[(sl * 32 + lane) * 2 + 1];
      }
      s_stats[2 * (c0 + lane)] += sm;
      s_stats[2 * (c0 + lane) + 1] += sq;
    }
  }
}

// fp16 two-pass epilogue, TMA-store flavour: the warp owning TMEM lane quarter q and column-block parity `half`
// drains 32-column blocks of one 128 x BN accumulator -- [Bh] + 2^-11 [Bl] + bias (+ fp16 residual), activation,
// round to fp16 -- into a 128B-swizzled staging tile ((BN/64) sub-tiles of [128 rows][64 fp16]) that one thread then
// writes to HBM with TMA bulk stores.  No per-thread global stores, no barrier inside the column loop.
// The fp16 residual of this thread's row (32 columns per block, at most two blocks for BN <= 128) is fetched into
// registers BEFORE the thread waits for the accumulator, so its latency hides behind the MMAs of the tile.
__device__ __forceinline__ void residual_prefetch_h(const TcParams& p, int n0, int64_t obase, int half, bool row_valid,
                                                    uint4 rp[8]) {
  const f16* res = reinterpret_cast<const f16*>(p.res_hi);
  if (!res || !row_valid) return;
  int blk = 0;
  for (int c0 = half * 32; c0 < p.BN && blk < 2; c0 += 64, ++blk) {
#pragma unroll
    for (int i = 0; i < 4; ++i) rp[blk * 4 + i] = *reinterpret_cast<const uint4*>(res + obase + n0 + c0 + i * 8);
  }
}

__device__ __forceinline__ void epilogue_stage_h(const TcParams& p, uint32_t stage, uint32_t tmem_acc, int n0,
                                                 int64_t obase, int half, int r, bool row_valid, const uint4 rp[8]) {
  const bool bias_vec = p.bias && ((reinterpret_cast<uintptr_t>(p.bias) & 15) == 0);
  const f16* res = reinterpret_cast<const f16*>(p.res_hi);
  int blk = 0;
  for (int c0 = half * 32; c0 < p.BN; c0 += 64, ++blk) {
    float v[32], v2[32];
    tmem_ld32(tmem_acc + (uint32_t)c0, v);
    tmem_ld32(tmem_acc + (uint32_t)(p.BN + c0), v2);
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = fmaf(v2[i], p.lo_scale, v[i]) * p.acc_scale;
    const int co0 = n0 + c0;
    if (bias_vec) {
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        const float4 bb = __ldg(reinterpret_cast<const float4*>(p.bias + co0 + i));
        v[i] += bb.x; v[i + 1] += bb.y; v[i + 2] += bb.z; v[i + 3] += bb.w;
      }
    } else if (p.bias) {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] += __ldg(p.bias + co0 + i);
    }
    if (res && row_valid) {
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        // blocks 0 and 1 come from the prefetch registers (static indexing keeps them in registers)
        const uint4 raw = blk == 0 ? rp[i >> 3] : blk == 1 ? rp[4 + (i >> 3)]
                                                           : *reinterpret_cast<const uint4*>(res + obase + co0 + i);
        const __half2* h2 = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 f = __half22float2(h2[k]);
          v[i + 2 * k] += f.x; v[i + 2 * k + 1] += f.y;
        }
      }
    }
    if (p.act == MP_ACT_RELU) {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
    } else if (p.act != MP_ACT_NONE) {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = mp_apply_act(v[i], p.act);
    }
    // 32 columns = four 16-byte chunks of this row; chunk index XOR (row & 7) = the 128B TMA swizzle
    const uint32_t sub = stage + (uint32_t)(c0 >> 6) * 16384u + (uint32_t)r * 128u;
    const uint32_t ch0 = (uint32_t)(c0 & 63) >> 3;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint32_t w[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const __half2 h = __floats2half2_rn(fminf(fmaxf(v[8 * j + 2 * k], -65504.f), 65504.f),
                                            fminf(fmaxf(v[8 * j + 2 * k + 1], -65504.f), 65504.f));
        w[k] = *reinterpret_cast<const uint32_t*>(&h);
      }
      const uint32_t addr = sub + (((ch0 + (uint32_t)j) ^ ((uint32_t)r & 7u)) << 4);
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3])
                   : "memory");
    }
  }
}

__device__ __forceinline__ void epilogue_flush_stats(const TcParams& p, double* s_stats, int n, int n0, int et) {
  epi_bar();   // the last chunk's fixed-order slab reduction (threads 0..31) must have landed
  const int cpg = p.gn_groups > 0 ? p.Cout / p.gn_groups : 1;
  for (int c = et; c < p.BN; c += 256) {
    const int co = n0 + c;
    if (co < p.Cout) {
      double* st = p.stats + ((int64_t)n * p.gn_groups + co / cpg) * 2;
      atomicAdd(st, s_stats[2 * c]);
      atomicAdd(st + 1, s_stats[2 * c + 1]);
    }
    s_stats[2 * c] = 0.0;
    s_stats[2 * c + 1] = 0.0;
  }
  epi_bar();
}

// ---- tile scheduler.  The TMA producer thread owns the tile sequence of its CTA: the first tile is blockIdx.x, every
// further one is drawn from a global atomic counter (p.sched) -- so a CTA that starts late, because another stream's
// kernel still held its SM, simply takes fewer tiles instead of delaying the whole launch -- or, with p.sched == NULL,
// the static walk blockIdx.x + k*gridDim.x.  Tile ids reach the MMA issuer and the epilogue warps through a small
// shared-memory queue guarded by mbarriers; -1 ends the walk.
__device__ __forceinline__ int sched_next(const TcParams& p, int tile) {
  return p.sched ? (int)(gridDim.x + atomicAdd(p.sched, 1u)) : tile + (int)gridDim.x;
}
__device__ __forceinline__ void sched_publish(uint32_t sfull, uint32_t sempty, volatile int* s_tile, uint32_t k, int tile) {
  const uint32_t slot = k % SQ;
  mbar_wait(sempty + slot * 8, ((k / SQ) & 1) ^ 1);
  s_tile[slot] = tile;
  mbar_arrive(sfull + slot * 8);            // release: the store above is visible to whoever acquires the barrier
}
__device__ __forceinline__ int sched_read(uint32_t sfull, volatile int* s_tile, uint32_t k) {
  const uint32_t slot = k % SQ;
  mbar_wait(sfull + slot * 8, (k / SQ) & 1);
  return s_tile[slot];
}
__device__ __forceinline__ void sched_finish(const TcParams& p) {
  if (p.sched) {                            // every CTA draws exactly one out-of-range id, then checks out
    const unsigned prev = atomicAdd(p.sched + 1, 1u);
    if (prev == gridDim.x - 1) { p.sched[0] = 0u; p.sched[1] = 0u; }
  }
}

// ------------------------------------------------------------------------------------------------ kernel v2
// Persistent: grid = min(#tiles, #SMs); every role walks tiles blockIdx.x, blockIdx.x + gridDim.x, ...
// TMEM holds TWO accumulators so the epilogue of tile i overlaps the MMAs of tile i+1.  The epilogue stages 32-column
// chunks in shared memory (padded rows) and writes them out with warp-contiguous 16-byte stores.
__global__ void __launch_bounds__(NUM_THREADS3, 1)
k_conv_tc2(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
           const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
           const __grid_constant__ CUtensorMap map_out, const __grid_constant__ CUtensorMap map_a2_hi,
           const __grid_constant__ CUtensorMap map_a2_lo, const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen_base = smem_raw + (smem_base - smem_u32(smem_raw));
  float* epi = reinterpret_cast<float*>(gen_base + p.epi_off);                    // [128][EPI_PITCH]
  long long* row_off = reinterpret_cast<long long*>(gen_base + p.epi_off + TILE_M * EPI_PITCH * 4);   // [128]
  const uint32_t bars = smem_base + p.epi_off + p.epi_bytes;   // full[S], empty[S], tmem_full[2], tmem_empty[2], bres
  const uint32_t bres_bar = bars + (2 * p.STAGES + 4) * 8;
  const uint32_t sfull = bars + (2 * p.STAGES + 5) * 8, sempty = sfull + SQ * 8;       // tile-id queue barriers
  volatile int* s_tile = reinterpret_cast<volatile int*>(gen_base + (sempty + SQ * 8 - smem_base));
  const uint32_t tmem_slot = sempty + SQ * 8 + SQ * 4;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen_base + (tmem_slot - smem_base));
  double* s_stats = reinterpret_cast<double*>(gen_base + (tmem_slot + 8 - smem_base));   // [BN][2]

  // warp index through a shuffle: tells the compiler it is warp-uniform, so role code stays on the uniform datapath
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  auto full_bar = [&](int s) { return bars + s * 8; };
  auto empty_bar = [&](int s) { return bars + (p.STAGES + s) * 8; };
  auto tfull_bar = [&](int b) { return bars + (2 * p.STAGES + b) * 8; };
  auto tempty_bar = [&](int b) { return bars + (2 * p.STAGES + 2 + b) * 8; };
  const int taps = p.KD * p.KH * p.KW;
  // K blocks per tile: live taps (taps that are out of bounds everywhere are skipped) + the fused shortcut's chunks
  const int num_kb = __popcll(p.tap_mask) * p.num_cchunks + p.num_cchunks2;

  if (threadIdx.x == 0) {
    // MP_PREC_F16_Q8: two MMA-issuing warps (warp 1: fp16 main product, warp 10: FP8 cross terms), so every barrier
    // that tracks MMA completion collects one tcgen05.commit per issuer
    const uint32_t nissue = p.prec == MP_PREC_F16_Q8 ? 2u : 1u;
    for (int s = 0; s < p.STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), nissue);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull_bar(b), nissue);
      mbar_init(tempty_bar(b), 8);      // one arrival per epilogue warp
    }
    mbar_init(bres_bar, 1);
    for (int i = 0; i < SQ; ++i) {
      mbar_init(sfull + i * 8, 1);
      mbar_init(sempty + i * 8, 8 + nissue);     // the MMA warp(s) + one lane of each epilogue warp
    }
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
    for (int i = threadIdx.x; i < 2 * p.BN; i += NUM_THREADS2) s_stats[i] = 0.0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot_ptr, 0);   // uniform for the compiler

  auto tile_coords = [&](int tile, int& n, int& d0, int& h0, int& w0, int& n0) {
    n0 = (tile % p.tiles_n) * p.BN;
    int t = tile / p.tiles_n;
    w0 = (t % p.tiles_w) * p.BW; t /= p.tiles_w;
    h0 = (t % p.tiles_h) * p.BH; t /= p.tiles_h;
    d0 = (t % p.tiles_d) * p.BD;
    n = (t / p.tiles_d) * p.BNb;           // first sample of the tile box
  };

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (lane == 0) {
      const int pd = p.KD / 2, ph = p.KH / 2, pw = p.KW / 2;
      uint32_t s = 0, ph_bit = 1;      // ring position; the producer waits on the "empty" phase, which starts at parity 1
      if (p.b_resident && (int)blockIdx.x < p.total_tiles) {
        // tiles_n == 1: every tile of this CTA uses the same weights -> fetch them once
        mbar_expect_tx(bres_bar, (uint32_t)num_kb * 2u * p.b_bytes);
        for (int i = 0; i < num_kb; ++i) {
          const uint32_t sb = smem_base + p.bres_off + (uint32_t)i * 2u * p.b_bytes;
          load_b(p, sb, &map_b_hi, &map_b_lo, bres_bar, i * p.CCHUNK, 0);
        }
      }
      const uint32_t tx = p.a_planes * p.a_bytes + (p.b_resident ? 0u : 2 * p.b_bytes);
      int tile = blockIdx.x;
      for (uint32_t tk = 0;; ++tk) {
        sched_publish(sfull, sempty, s_tile, tk, tile < p.total_tiles ? tile : -1);
        if (tile >= p.total_tiles) break;
        const int next_tile = sched_next(p, tile);      // drawn early: the atomic's latency hides behind this tile's loads
        int n, d0, h0, w0, n0;
        tile_coords(tile, n, d0, h0, w0, n0);
        int kw = -1, kh = 0, kd = 0;                  // tap -> (kd, kh, kw) without divisions
        for (int tap = 0; tap < taps; ++tap) {
          if (++kw == p.KW) { kw = 0; if (++kh == p.KH) { kh = 0; ++kd; } }
          if (!((p.tap_mask >> tap) & 1ull)) continue;
          for (int cc = 0; cc < p.num_cchunks; ++cc) {
            mbar_wait(empty_bar(s), ph_bit);
            const uint32_t sa = smem_base + s * p.stage_bytes;
            mbar_expect_tx(full_bar(s), tx);
            const int c0 = cc * p.CCHUNK;
            const int ci = c0 + p.in_c_off, wi = w0 * p.stride + kw - pw, hi_ = h0 * p.stride + kh - ph;
            tma_load_5d(sa, &map_a_hi, full_bar(s), ci, wi, hi_, d0 + kd - pd, n);
            if (p.a_planes == 2) tma_load_5d(sa + p.a_bytes, &map_a_lo, full_bar(s), ci, wi, hi_, d0 + kd - pd, n);
            if (!p.b_resident) {
              load_b(p, sa + p.a_planes * p.a_bytes, &map_b_hi, &map_b_lo, full_bar(s), tap * p.Cin + c0, n0);
            }
            if (++s == (uint32_t)p.STAGES) { s = 0; ph_bit ^= 1u; }
          }
        }
        for (int cc = 0; cc < p.num_cchunks2; ++cc) {      // fused 1x1 shortcut: centre tap of the second source
          mbar_wait(empty_bar(s), ph_bit);
          const uint32_t sa = smem_base + s * p.stage_bytes;
          mbar_expect_tx(full_bar(s), tx);
          const int c0 = cc * p.CCHUNK;
          tma_load_5d(sa, &map_a2_hi, full_bar(s), c0 + p.in2_c_off, w0 * p.stride2, h0 * p.stride2, d0, n);
          if (p.a_planes == 2)
            tma_load_5d(sa + p.a_bytes, &map_a2_lo, full_bar(s), c0 + p.in2_c_off, w0 * p.stride2, h0 * p.stride2, d0, n);
          if (!p.b_resident) load_b(p, sa + p.a_planes * p.a_bytes, &map_b_hi, &map_b_lo, full_bar(s), p.k2_off + c0, n0);
          if (++s == (uint32_t)p.STAGES) { s = 0; ph_bit ^= 1u; }
        }
        tile = next_tile;
      }
      sched_finish(p);
    }
  } else if (warp == 1 || warp == 10) {
    // ===================================================================== MMA issuers (whole warp, elected lane issues)
    // warp 10 exists for MP_PREC_F16_Q8 only: it issues the FP8 cross-term MMAs into the accumulator's second column
    // half while warp 1 issues the fp16 main product into the first
    const bool q8_issuer = warp == 10;
    if (!q8_issuer || p.prec == MP_PREC_F16_Q8) {
      const int ksteps = p.CCHUNK / 16;
      const uint32_t dhi = desc_hi_word(p.sbo, p.layout_type);
      uint32_t it = 0, s = 0, ph_bit = 0;        // smem ring position / phase (no div/mod on the issue path)
      if (p.b_resident && (int)blockIdx.x < p.total_tiles) mbar_wait(bres_bar, 0);
      for (;; ++it) {
        const int tile = __shfl_sync(0xffffffffu, sched_read(sfull, s_tile, it), 0);   // uniform loop exit
        if (lane == 0) mbar_arrive(sempty + (it % SQ) * 8);
        __syncwarp();
        if (tile < 0) break;
        const uint32_t b = it & 1;
        mbar_wait(tempty_bar(b), ((it >> 1) & 1) ^ 1);     // epilogue has drained this accumulator
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d_tmem = tmem_base + b * (uint32_t)p.acc_w;
        for (int i = 0; i < num_kb; ++i) {
          mbar_wait(full_bar(s), ph_bit);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t sa = smem_base + s * p.stage_bytes;
          const uint32_t a_hi = sa, a_lo = sa + p.a_bytes;
          const uint32_t b_hi =
              p.b_resident ? smem_base + p.bres_off + (uint32_t)i * 2u * p.b_bytes : sa + p.a_planes * p.a_bytes;
          const uint32_t b_lo = b_hi + p.b_bytes;
          if (p.prec == MP_PREC_F16_Q8) {
            // 64-channel chunks: [fp16 x fp16] -> columns [0, BN), [x8|xl8] x [wl8|w8] -> columns [BN, 2 BN)
            if (q8_issuer) umma_chunk_q8<4>(d_tmem + (uint32_t)p.BN, a_lo, b_lo, dhi, p.idesc, i > 0 ? 1u : 0u);
            else umma_chunk_h<4>(d_tmem, a_hi, b_hi, dhi, p.idesc, i > 0 ? 1u : 0u);
          } else if (p.prec == MP_PREC_F16X2) umma_chunk_h_dyn(ksteps, d_tmem, a_hi, b_hi, dhi, p.idesc2, i > 0 ? 1u : 0u);
          else if (p.dualb) umma_chunk_dual_dyn(ksteps, d_tmem, a_hi, a_lo, b_hi, dhi, p.idesc, p.idesc2, i > 0 ? 1u : 0u);
          else umma_chunk_dyn(ksteps, d_tmem, a_hi, a_lo, b_hi, b_lo, dhi, p.idesc, i > 0 ? 1u : 0u);
          umma_commit_e(empty_bar(s));
          if (++s == (uint32_t)p.STAGES) { s = 0; ph_bit ^= 1u; }
        }
        umma_commit_e(tfull_bar(b));
      }
    }
  } else {
    // ===================================================================== epilogue (warps 2..9, 256 threads)
    // Warp w may touch TMEM lanes (w % 4) * 32 .. +31; the two warps that share a lane quarter split every 32-column
    // chunk into halves of 16 columns.  Two warps per scheduler keep the dependent-issue latency hidden.
    const int q = warp & 3;                    // TMEM lane quarter
    const int half = (warp - 2) >> 2;          // which 16 columns of a chunk
    const int r = q * 32 + lane;               // accumulator row == position inside the tile
    const int et = threadIdx.x - 64;           // 0..255
    uint32_t it = 0;
    for (;; ++it) {
      const int tile = sched_read(sfull, s_tile, it);
      __syncwarp();
      if (lane == 0) mbar_arrive(sempty + (it % SQ) * 8);
      if (tile < 0) break;
      int n, d0, h0, w0, n0;
      tile_coords(tile, n, d0, h0, w0, n0);
      const uint32_t b = it & 1;
      const int sbox = p.BW * p.BH * p.BD;
      const int ni = r / sbox, rr = r - ni * sbox;
      const int ww = rr % p.BW, hh = (rr / p.BW) % p.BH, dd = rr / (p.BW * p.BH);
      const bool row_valid = n + ni < p.N;
      const int64_t pos = (((int64_t)(n + ni) * p.D + d0 + dd) * p.H + h0 + hh) * p.W + w0 + ww;
      const int64_t obase = pos * p.out_C + p.out_c_off;
      uint4 rp[8];
      if (p.epi_tma) residual_prefetch_h(p, n0, obase, half, row_valid, rp);
      mbar_wait(tfull_bar(b), (it >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (p.epi_tma) {
        const uint32_t stage = smem_base + p.tma_off + (p.tma_nbuf == 2 ? b : 0u) * p.tma_buf_bytes;
        if (et == 0) {                         // the bulk store that last read this staging buffer has drained it
          if (p.tma_nbuf == 2) bulk_wait_read<1>();
          else bulk_wait_read<0>();
        }
        epi_bar();
        epilogue_stage_h(p, stage, tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(b * p.acc_w), n0, obase, half, r,
                         row_valid, rp);
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(b));
        fence_proxy_async();                   // generic-proxy staging writes -> visible to the TMA engine
        epi_bar();
        if (et == 0) {
          for (int sub = 0; sub < (p.BN >> 6); ++sub)
            tma_store_5d(&map_out, stage + (uint32_t)sub * 16384u, p.out_c_off + n0 + sub * 64, w0, h0, d0, n);
          bulk_commit();
        }
        continue;
      }
      epilogue_drain(p, epi, row_off, s_stats, tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(b * p.acc_w), n0,
                     obase, half, r, et, lane, row_valid);
      // accumulator drained: hand the TMEM buffer back to the MMA warp
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(b));
      if (p.stats) epilogue_flush_stats(p, s_stats, n, n0, et);
    }
    if (p.epi_tma && et == 0) bulk_wait_all();   // outstanding bulk stores must finish reading shared memory
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------ kernel v4 (CTA pair)
// MP_PREC_F16_Q8 convolutions on `cta_group::2`: a cluster of two CTAs (one SM pair) computes a 256-position x 128-channel
// tile with M = 256 MMAs.  Each CTA stages ITS 128 positions of A (fp16 plane + FP8 byte plane) and HALF of the weight
// tile (64 of the 128 output channels), the leader CTA's two issuing warps (fp16 main product, FP8 cross terms) drive the
// tensor cores of both SMs, and each CTA's TMEM ends up with its own 128 rows of both accumulators.  Per SM and K step the
// tensor core now fetches 4 KB of A + 2 KB of B instead of 4 + 4 KB -- the single-CTA kernel sits exactly on the 128 B/clk
// shared-memory limit there (ncu: tensor pipe 70 %) -- and the weight tile crosses L2 -> SM once per pair instead of once
// per CTA.
//   barriers (same offsets in both CTAs): full[s]   leader only; the leader's producer expects the bytes of BOTH CTAs, the
//                                                   peer's TMA loads signal it through the cluster address (.cta_group::2)
//                                         empty[s]  per CTA; tcgen05.commit multicasts the two issuers' arrivals to both
//                                         tfull[b]  per CTA (multicast commit): accumulator b is complete
//                                         tempty[b] leader only: 8 epilogue warps of each CTA arrive (the peer's remotely)
//   tiles: static walk, item = cluster id + k * clusters, item -> (pair of position tiles, channel tile), channel fastest.
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_id_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_count_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_5d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1,
                                                int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1,
                                                int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
template <bool F8>
__device__ __forceinline__ void umma_pair_e(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc,
                                            uint32_t accumulate) {
  if (F8)
    asm volatile(
        "{\n\t.reg .pred p, e;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "mov.b64 da, {%1, %3};\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "@e tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], da, db, %4, p;\n\t}"
        ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(accumulate)
        : "memory");
  else
    asm volatile(
        "{\n\t.reg .pred p, e;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "mov.b64 da, {%1, %3};\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "@e tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, p;\n\t}"
        ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
// completion of every MMA this thread issued so far -> one arrival on the barrier at this offset in BOTH CTAs
__device__ __forceinline__ void umma_commit_pair_e(uint32_t bar) {
  asm volatile(
      "{\n\t.reg .pred e;\n\t.reg .b16 m;\n\t"
      "mov.b16 m, 3;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n\t}" ::"r"(bar)
      : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS3, 1)
k_conv_q8_pair(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
               const __grid_constant__ CUtensorMap map_b, const __grid_constant__ CUtensorMap map_a2_hi,
               const __grid_constant__ CUtensorMap map_a2_lo, const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen_base = smem_raw + (smem_base - smem_u32(smem_raw));
  float* epi = reinterpret_cast<float*>(gen_base + p.epi_off);
  long long* row_off = reinterpret_cast<long long*>(gen_base + p.epi_off + TILE_M * EPI_PITCH * 4);
  const uint32_t bars = smem_base + p.epi_off + p.epi_bytes;   // full[S], empty[S], tfull[2], tempty[2]
  const uint32_t tmem_slot = bars + (2 * p.STAGES + 4) * 8;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen_base + (tmem_slot - smem_base));
  double* s_stats = reinterpret_cast<double*>(gen_base + (tmem_slot + 8 - smem_base));   // unused (no GroupNorm statistics here)

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  auto full_bar = [&](int s) { return bars + s * 8; };
  auto empty_bar = [&](int s) { return bars + (p.STAGES + s) * 8; };
  auto tfull_bar = [&](int b) { return bars + (2 * p.STAGES + b) * 8; };
  auto tempty_bar = [&](int b) { return bars + (2 * p.STAGES + 2 + b) * 8; };
  const int taps = p.KD * p.KH * p.KW;
  const int num_kb = taps * p.num_cchunks + p.num_cchunks2;      // + the K chunks of a fused 1x1 shortcut
  const int items = (p.total_tiles / p.tiles_n / 2) * p.tiles_n;     // (pairs of position tiles) x channel tiles

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 2);       // two issuing warps, multicast commits
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull_bar(b), 2);
      mbar_init(tempty_bar(b), 16);     // 8 epilogue warps of each CTA
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a_hi)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a_lo)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_b)) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();                    // both CTAs' barriers are initialised before anything signals them
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot_ptr, 0);

  // item -> my position tile (pair index * 2 + rank) and the channel tile
  auto item_coords = [&](int item, int& n, int& d0, int& h0, int& w0, int& n0) {
    n0 = (item % p.tiles_n) * p.BN;
    int t = (item / p.tiles_n) * 2 + (int)rank;
    w0 = (t % p.tiles_w) * p.BW; t /= p.tiles_w;
    h0 = (t % p.tiles_h) * p.BH; t /= p.tiles_h;
    d0 = (t % p.tiles_d) * p.BD;
    n = t / p.tiles_d;
  };
  const int first = (int)cluster_id_x(), step = (int)cluster_count_x();

  if (warp == 0) {
    // ===================================================================== TMA producer (both CTAs)
    if (lane == 0) {
      const int pd = p.KD / 2, ph = p.KH / 2, pw = p.KW / 2;
      uint32_t s = 0, ph_bit = 1;
      const uint32_t tx_both = 2u * (2u * p.a_bytes + 2u * p.b_bytes);      // A (2 planes) + B half (2 planes), both CTAs
      for (int item = first; item < items; item += step) {
        int n, d0, h0, w0, n0;
        item_coords(item, n, d0, h0, w0, n0);
        int kw = -1, kh = 0, kd = 0;
        for (int tap = 0; tap < taps; ++tap) {
          if (++kw == p.KW) { kw = 0; if (++kh == p.KH) { kh = 0; ++kd; } }
          for (int cc = 0; cc < p.num_cchunks; ++cc) {
            mbar_wait(empty_bar(s), ph_bit);
            const uint32_t sa = smem_base + s * p.stage_bytes;
            const uint32_t fb = mapa_shared(full_bar(s), 0);                 // the LEADER's barrier, cluster address
            if (leader) mbar_expect_tx(full_bar(s), tx_both);
            const int c0 = cc * p.CCHUNK;
            tma_load_5d_2sm(sa, &map_a_hi, fb, c0, w0 + kw - pw, h0 + kh - ph, d0 + kd - pd, n);
            tma_load_5d_2sm(sa + p.a_bytes, &map_a_lo, fb, c0, w0 + kw - pw, h0 + kh - ph, d0 + kd - pd, n);
            tma_load_3d_2sm(sa + 2 * p.a_bytes, &map_b, fb, tap * p.Cin + c0, n0 + (int)rank * (p.BN / 2), 0);
            if (++s == (uint32_t)p.STAGES) { s = 0; ph_bit ^= 1u; }
          }
        }
        for (int cc = 0; cc < p.num_cchunks2; ++cc) {        // fused 1x1 shortcut: centre tap of the second source
          mbar_wait(empty_bar(s), ph_bit);
          const uint32_t sa = smem_base + s * p.stage_bytes;
          const uint32_t fb = mapa_shared(full_bar(s), 0);
          if (leader) mbar_expect_tx(full_bar(s), tx_both);
          const int c0 = cc * p.CCHUNK;
          tma_load_5d_2sm(sa, &map_a2_hi, fb, c0 + p.in2_c_off, w0, h0, d0, n);
          tma_load_5d_2sm(sa + p.a_bytes, &map_a2_lo, fb, c0 + p.in2_c_off, w0, h0, d0, n);
          tma_load_3d_2sm(sa + 2 * p.a_bytes, &map_b, fb, p.k2_off + c0, n0 + (int)rank * (p.BN / 2), 0);
          if (++s == (uint32_t)p.STAGES) { s = 0; ph_bit ^= 1u; }
        }
      }
    }
  } else if (warp == 1 || warp == 10) {
    // ===================================================================== MMA issuers (leader CTA only)
    if (leader) {
      const bool q8_issuer = warp == 10;
      const uint32_t dhi = desc_hi_word(p.sbo, p.layout_type);
      uint32_t it = 0, s = 0, ph_bit = 0;
      for (int item = first; item < items; item += step, ++it) {
        const uint32_t b = it & 1;
        mbar_wait(tempty_bar(b), ((it >> 1) & 1) ^ 1);     // both CTAs' epilogues have drained this accumulator
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d_tmem = tmem_base + b * (uint32_t)p.acc_w;
        for (int i = 0; i < num_kb; ++i) {
          mbar_wait(full_bar(s), ph_bit);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t sa = smem_base + s * p.stage_bytes;
          const uint32_t a_hi = sa, a_lo = sa + p.a_bytes, b_hi = sa + 2 * p.a_bytes, b_lo = b_hi + p.b_bytes;
          if (q8_issuer) {
            const uint32_t aq = desc_lo_word(a_lo), bq = desc_lo_word(b_lo);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_pair_e<true>(d_tmem + (uint32_t)p.BN, aq + 2 * k, bq + 2 * k, dhi, p.idesc, (i > 0 || k > 0) ? 1u : 0u);
          } else {
            const uint32_t ah = desc_lo_word(a_hi), bh = desc_lo_word(b_hi);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_pair_e<false>(d_tmem, ah + 2 * k, bh + 2 * k, dhi, p.idesc, (i > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit_pair_e(empty_bar(s));
          if (++s == (uint32_t)p.STAGES) { s = 0; ph_bit ^= 1u; }
        }
        umma_commit_pair_e(tfull_bar(b));
      }
    }
  } else {
    // ===================================================================== epilogue (warps 2..9 of both CTAs)
    const int q = warp & 3, half = (warp - 2) >> 2, r = q * 32 + lane, et = threadIdx.x - 64;
    const uint32_t tempty_leader0 = mapa_shared(tempty_bar(0), 0), tempty_leader1 = mapa_shared(tempty_bar(1), 0);
    uint32_t it = 0;
    for (int item = first; item < items; item += step, ++it) {
      int n, d0, h0, w0, n0;
      item_coords(item, n, d0, h0, w0, n0);
      const uint32_t b = it & 1;
      const int ww = r % p.BW, hh = (r / p.BW) % p.BH, dd = r / (p.BW * p.BH);
      const int64_t pos = (((int64_t)n * p.D + d0 + dd) * p.H + h0 + hh) * p.W + w0 + ww;
      const int64_t obase = pos * p.out_C + p.out_c_off;
      mbar_wait(tfull_bar(b), (it >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      epilogue_drain(p, epi, row_off, s_stats, tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(b * p.acc_w), n0, obase,
                     half, r, et, lane, true);
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(b ? tempty_leader1 : tempty_leader0);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  cluster_sync_all();                    // the peer may still be reading my shared memory / signalling my barriers
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------ kernel v3 (slab)
// 3x3(x3) convolutions with few output channels are bound by the L2 -> shared-memory feed of the A operand when
// every filter tap re-fetches its own shifted 128-position box.  Here the output tile is an (MT*BH) x BW patch
// (BH*BW = 128, BW a multiple of 8) and ONE TMA box of (MT*BH + 2) x BW pixel rows -- the patch plus its vertical
// halo, shifted by the horizontal tap kw -- serves all three vertical taps of all MT accumulators: tap kh of
// accumulator t simply starts (t*BH + kh)*BW pixel rows into the slab, which keeps the UMMA descriptor start
// 1024-byte aligned (BW % 8 == 0), so the swizzle phase is unchanged.  A traffic drops 2.7x (MT = 2: 34 rows
// instead of 96) and the weight tile of a tap is shared by the MT accumulators.  A slabs and B tiles travel in two
// independent mbarrier rings.
struct SlabExtra {
  int MT, SA, SB;
  uint32_t a_plane_bytes;     // one plane of one slab
  uint32_t b_ring_off;        // byte offset of the B ring
};

__global__ void __launch_bounds__(NUM_THREADS3, 1)
k_conv_tc3(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
           const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
           const __grid_constant__ CUtensorMap map_out, const __grid_constant__ CUtensorMap map_a2_hi,
           const __grid_constant__ CUtensorMap map_a2_lo, const TcParams p, const SlabExtra x) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen_base = smem_raw + (smem_base - smem_u32(smem_raw));
  float* epi = reinterpret_cast<float*>(gen_base + p.epi_off);
  long long* row_off = reinterpret_cast<long long*>(gen_base + p.epi_off + TILE_M * EPI_PITCH * 4);
  // barriers: fullA[SA], emptyA[SA], fullB[SB], emptyB[SB], tmem_full[2], tmem_empty[2]
  const uint32_t bars = smem_base + p.epi_off + p.epi_bytes;
  auto fullA = [&](int i) { return bars + i * 8; };
  auto emptyA = [&](int i) { return bars + (x.SA + i) * 8; };
  auto fullB = [&](int i) { return bars + (2 * x.SA + i) * 8; };
  auto emptyB = [&](int i) { return bars + (2 * x.SA + x.SB + i) * 8; };
  auto tfull_bar = [&](int b) { return bars + (2 * x.SA + 2 * x.SB + b) * 8; };
  auto tempty_bar = [&](int b) { return bars + (2 * x.SA + 2 * x.SB + 2 + b) * 8; };
  const uint32_t sfull = bars + (2 * x.SA + 2 * x.SB + 4) * 8, sempty = sfull + SQ * 8;   // tile-id queue barriers
  volatile int* s_tile = reinterpret_cast<volatile int*>(gen_base + (sempty + SQ * 8 - smem_base));
  const uint32_t tmem_slot = sempty + SQ * 8 + SQ * 4;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen_base + (tmem_slot - smem_base));
  double* s_stats = reinterpret_cast<double*>(gen_base + (tmem_slot + 8 - smem_base));

  // warp index through a shuffle: tells the compiler it is warp-uniform, so role code stays on the uniform datapath
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const uint32_t a_stage = p.a_planes * x.a_plane_bytes, b_stage = 2 * p.b_bytes;
  const uint32_t row_bytes = (uint32_t)p.CCHUNK * 2u;

  if (threadIdx.x == 0) {
    // with MT = 2 accumulators per weight tile TWO warps issue MMAs (one accumulator each): every "consumed" barrier then
    // collects one tcgen05.commit per issuer
    const uint32_t nissue = x.MT == 2 ? 2u : 1u;
    for (int i = 0; i < x.SA; ++i) { mbar_init(fullA(i), 1); mbar_init(emptyA(i), nissue); }
    for (int i = 0; i < x.SB; ++i) { mbar_init(fullB(i), 1); mbar_init(emptyB(i), nissue); }
    for (int b = 0; b < 2; ++b) { mbar_init(tfull_bar(b), nissue); mbar_init(tempty_bar(b), 8); }
    for (int i = 0; i < SQ; ++i) { mbar_init(sfull + i * 8, 1); mbar_init(sempty + i * 8, 8 + nissue); }
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
    for (int i = threadIdx.x; i < 2 * p.BN; i += NUM_THREADS2) s_stats[i] = 0.0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot_ptr, 0);   // uniform for the compiler

  // super-tile id -> (n, d, h0, w0, n0); tiles_h counts super-tiles of MT*BH rows, tiles_d = D (BD = 1)
  auto tile_coords = [&](int tile, int& n, int& d0, int& h0, int& w0, int& n0) {
    n0 = (tile % p.tiles_n) * p.BN;
    int t = tile / p.tiles_n;
    w0 = (t % p.tiles_w) * p.BW; t /= p.tiles_w;
    h0 = (t % p.tiles_h) * (p.BH * x.MT); t /= p.tiles_h;
    d0 = t % p.tiles_d;
    n = t / p.tiles_d;
  };

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (lane == 0) {
      const int pd = p.KD / 2, pw = p.KW / 2;
      uint32_t sidx = 0, a_ph = 1, bidx = 0, b_ph = 1;   // producer waits on the "empty" phase: starts at parity 1
      int tile = blockIdx.x;
      for (uint32_t tk = 0;; ++tk) {
        sched_publish(sfull, sempty, s_tile, tk, tile < p.total_tiles ? tile : -1);
        if (tile >= p.total_tiles) break;
        const int next_tile = sched_next(p, tile);
        int n, d0, h0, w0, n0;
        tile_coords(tile, n, d0, h0, w0, n0);
        for (int kd = 0; kd < p.KD; ++kd)
          for (int kw = 0; kw < p.KW; ++kw)
            for (int cc = 0; cc < p.num_cchunks; ++cc) {
              const int c0 = cc * p.CCHUNK;
              {
                mbar_wait(emptyA(sidx), a_ph);
                const uint32_t sa = smem_base + sidx * a_stage;
                mbar_expect_tx(fullA(sidx), a_stage);
                tma_load_5d(sa, &map_a_hi, fullA(sidx), c0 + p.in_c_off, w0 + kw - pw, h0 - 1, d0 + kd - pd, n);
                if (p.a_planes == 2)
                  tma_load_5d(sa + x.a_plane_bytes, &map_a_lo, fullA(sidx), c0 + p.in_c_off, w0 + kw - pw, h0 - 1,
                              d0 + kd - pd, n);
                if (++sidx == (uint32_t)x.SA) { sidx = 0; a_ph ^= 1u; }
              }
              for (int kh = 0; kh < 3; ++kh) {
                mbar_wait(emptyB(bidx), b_ph);
                const uint32_t sb = smem_base + x.b_ring_off + bidx * b_stage;
                mbar_expect_tx(fullB(bidx), b_stage);
                const int tap = (kd * 3 + kh) * p.KW + kw;
                load_b(p, sb, &map_b_hi, &map_b_lo, fullB(bidx), tap * p.Cin + c0, n0);
                if (++bidx == (uint32_t)x.SB) { bidx = 0; b_ph ^= 1u; }
              }
            }
        for (int cc = 0; cc < p.num_cchunks2; ++cc) {      // fused 1x1 shortcut: one slab (centre column) + one weight tile
          const int c0 = cc * p.CCHUNK;
          mbar_wait(emptyA(sidx), a_ph);
          const uint32_t sa = smem_base + sidx * a_stage;
          mbar_expect_tx(fullA(sidx), a_stage);
          tma_load_5d(sa, &map_a2_hi, fullA(sidx), c0 + p.in2_c_off, w0 * p.stride2, (h0 - 1) * p.stride2, d0, n);
          if (p.a_planes == 2)
            tma_load_5d(sa + x.a_plane_bytes, &map_a2_lo, fullA(sidx), c0 + p.in2_c_off, w0 * p.stride2,
                        (h0 - 1) * p.stride2, d0, n);
          if (++sidx == (uint32_t)x.SA) { sidx = 0; a_ph ^= 1u; }
          mbar_wait(emptyB(bidx), b_ph);
          const uint32_t sb = smem_base + x.b_ring_off + bidx * b_stage;
          mbar_expect_tx(fullB(bidx), b_stage);
          load_b(p, sb, &map_b_hi, &map_b_lo, fullB(bidx), p.k2_off + c0, n0);
          if (++bidx == (uint32_t)x.SB) { bidx = 0; b_ph ^= 1u; }
        }
        tile = next_tile;
      }
      sched_finish(p);
    }
  } else if (warp == 1 || warp == 10) {
    // ===================================================================== MMA issuers (whole warp, elected lane issues)
    // warp 1 owns accumulator 0 (and all of them when MT = 1); warp 10 owns accumulator 1 when MT = 2: the single
    // dependent instruction stream of one issuing warp cannot keep 64-cycle (N <= 128) MMAs back to back
    const int t0 = warp == 1 ? 0 : 1, tstep = x.MT == 2 ? 2 : 1;
    if (t0 < x.MT && (warp == 1 || x.MT == 2)) {
      const int ksteps = p.CCHUNK / 16;
      const uint32_t dhi = desc_hi_word(p.sbo, p.layout_type);
      uint32_t it = 0, sidx = 0, a_ph = 0, bidx = 0, b_ph = 0;   // ring positions / phases (no div/mod on the issue path)
      const int groups = p.KD * p.KW * p.num_cchunks;
      const uint32_t kh_step = (uint32_t)p.BW * row_bytes, acc_step = (uint32_t)(p.BH * p.BW) * row_bytes;
      for (;; ++it) {
        const int tile = __shfl_sync(0xffffffffu, sched_read(sfull, s_tile, it), 0);   // uniform loop exit
        if (lane == 0) mbar_arrive(sempty + (it % SQ) * 8);
        __syncwarp();
        if (tile < 0) break;
        const uint32_t b = it & 1;
        mbar_wait(tempty_bar(b), ((it >> 1) & 1) ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d_tmem0 = tmem_base + b * (uint32_t)(x.MT * p.acc_w);
        bool first = true;
        for (int g = 0; g < groups; ++g) {
          mbar_wait(fullA(sidx), a_ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t sa = smem_base + sidx * a_stage;
          for (int kh = 0; kh < 3; ++kh) {
            mbar_wait(fullB(bidx), b_ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t b_hi = smem_base + x.b_ring_off + bidx * b_stage, b_lo = b_hi + p.b_bytes;
            for (int t = t0; t < x.MT; t += tstep) {
              const uint32_t a_hi = sa + (uint32_t)t * acc_step + (uint32_t)kh * kh_step;
              const uint32_t a_lo = a_hi + x.a_plane_bytes;
              const uint32_t d_tmem = d_tmem0 + (uint32_t)(t * p.acc_w);
              if (p.prec == MP_PREC_F16X2) umma_chunk_h_dyn(ksteps, d_tmem, a_hi, b_hi, dhi, p.idesc2, first ? 0u : 1u);
              else if (p.dualb) umma_chunk_dual_dyn(ksteps, d_tmem, a_hi, a_lo, b_hi, dhi, p.idesc, p.idesc2, first ? 0u : 1u);
              else umma_chunk_dyn(ksteps, d_tmem, a_hi, a_lo, b_hi, b_lo, dhi, p.idesc, first ? 0u : 1u);
            }
            first = false;
            umma_commit_e(emptyB(bidx));
            if (++bidx == (uint32_t)x.SB) { bidx = 0; b_ph ^= 1u; }
          }
          umma_commit_e(emptyA(sidx));
          if (++sidx == (uint32_t)x.SA) { sidx = 0; a_ph ^= 1u; }
        }
        for (int g = 0; g < p.num_cchunks2; ++g) {           // fused 1x1 shortcut: the slab's centre row offset (kh = 1)
          mbar_wait(fullA(sidx), a_ph);
          mbar_wait(fullB(bidx), b_ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t sa = smem_base + sidx * a_stage;
          const uint32_t b_hi = smem_base + x.b_ring_off + bidx * b_stage, b_lo = b_hi + p.b_bytes;
          for (int t = t0; t < x.MT; t += tstep) {
            const uint32_t a_hi = sa + (uint32_t)t * acc_step + kh_step;
            const uint32_t a_lo = a_hi + x.a_plane_bytes;
            const uint32_t d_tmem = d_tmem0 + (uint32_t)(t * p.acc_w);
            if (p.prec == MP_PREC_F16X2) umma_chunk_h_dyn(ksteps, d_tmem, a_hi, b_hi, dhi, p.idesc2, first ? 0u : 1u);
            else if (p.dualb) umma_chunk_dual_dyn(ksteps, d_tmem, a_hi, a_lo, b_hi, dhi, p.idesc, p.idesc2, first ? 0u : 1u);
            else umma_chunk_dyn(ksteps, d_tmem, a_hi, a_lo, b_hi, b_lo, dhi, p.idesc, first ? 0u : 1u);
          }
          first = false;
          umma_commit_e(emptyB(bidx));
          if (++bidx == (uint32_t)x.SB) { bidx = 0; b_ph ^= 1u; }
          umma_commit_e(emptyA(sidx));
          if (++sidx == (uint32_t)x.SA) { sidx = 0; a_ph ^= 1u; }
        }
        umma_commit_e(tfull_bar(b));
      }
    }
  } else {
    // ===================================================================== epilogue (warps 2..9)
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int r = q * 32 + lane;
    const int et = threadIdx.x - 64;
    uint32_t it = 0, se = 0;
    for (;; ++it) {
      const int tile = sched_read(sfull, s_tile, it);
      __syncwarp();
      if (lane == 0) mbar_arrive(sempty + (it % SQ) * 8);
      if (tile < 0) break;
      int n, d0, h0, w0, n0;
      tile_coords(tile, n, d0, h0, w0, n0);
      const uint32_t b = it & 1;
      const int ww = r % p.BW, hh = r / p.BW;
      // MT <= 2: both accumulators' residual rows are in flight (in two statically indexed register sets) while the MMAs
      // of the tile finish
      uint4 rp0[8], rp1[8];
      auto obase_of = [&](int t) {
        const int64_t pos = (((int64_t)n * p.D + d0) * p.H + h0 + t * p.BH + hh) * p.W + w0 + ww;
        return pos * p.out_C + p.out_c_off;
      };
      if (p.epi_tma) {
        residual_prefetch_h(p, n0, obase_of(0), half, true, rp0);
        if (x.MT == 2) residual_prefetch_h(p, n0, obase_of(1), half, true, rp1);
      }
      mbar_wait(tfull_bar(b), (it >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      auto drain = [&](int t, const uint4* rp) {
        const int64_t obase = obase_of(t);
        const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((b * x.MT + t) * p.acc_w);
        if (p.epi_tma) {
          const uint32_t stage = smem_base + p.tma_off + (p.tma_nbuf == 2 ? (se & 1u) : 0u) * p.tma_buf_bytes;
          ++se;
          if (et == 0) {
            if (p.tma_nbuf == 2) bulk_wait_read<1>();
            else bulk_wait_read<0>();
          }
          epi_bar();
          epilogue_stage_h(p, stage, tacc, n0, obase, half, r, true, rp);
          fence_proxy_async();
          epi_bar();
          if (et == 0) {
            for (int sub = 0; sub < (p.BN >> 6); ++sub)
              tma_store_5d(&map_out, stage + (uint32_t)sub * 16384u, p.out_c_off + n0 + sub * 64, w0, h0 + t * p.BH, d0, n);
            bulk_commit();
          }
        } else {
          epilogue_drain(p, epi, row_off, s_stats, tacc, n0, obase, half, r, et, lane);
        }
      };
      drain(0, rp0);
      if (x.MT == 2) drain(1, rp1);
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(b));
      if (p.stats) epilogue_flush_stats(p, s_stats, n, n0, et);
    }
    if (p.epi_tma && et == 0) bulk_wait_all();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------ weight gradient (tcgen05)
// dW[co][tap][ci] = sum_p dY[p][co] * X[p + shift(tap)][ci]: one GEMM per filter tap whose K axis is the POSITION axis.  Both
// operands are channels-last tensors, i.e. [K = position][MN = channel] with the channel contiguous: "MN-major" in UMMA terms.
// A TMA box of (64 channels, bw, bh, bd, bn positions) with SWIZZLE_128B lands in shared memory as [positions][128 B] rows --
// exactly the canonical MN-major SWIZZLE_128B layout ((8,n),(8,k)):((1,LBO),(8,SBO)) (16-byte units): 8-row groups 1 024 B apart
// (SBO), 64-channel sub-tiles LBO = positions * 128 B apart.  So the tensor cores read dY and X as they lie in HBM: no transpose,
// no im2col; the tap shift is the box origin of the X load and the zero border is TMA's out-of-bounds fill (which also pads
// channel counts up to the 128 x BN tile and ragged boxes at the volume edge).  Three bf16 passes (hi*hi + hi*lo + lo*hi) with fp32
// accumulation in TMEM, one accumulator per tap of the CTA's tap group (the un-shifted operand is staged once per group),
// split-K over position boxes with a vectorised fp32 RED epilogue.
struct WgParams {
  float* dw;
  int Cout, Cin, T, KD, KH, KW;
  int swap;                       // 0: M = output channels (dY), N = input channels (X);  1: M = input channels, N = output channels
  int BN, TG;                     // N tile (multiple of 16, <= 128); taps per CTA
  int tiles_m, tiles_n, tap_groups;
  int bw, bh, bd, bn, kblk;       // position box (K block)
  int nbw, nbh, nbd;              // boxes per dimension (the sample axis takes the rest)
  int total_boxes, boxes_per_slice;
  int STAGES;
  int s_sub, x_sub;               // 64-channel sub-tiles of the shared (dY) / shifted (X) operand per stage
  uint32_t sub_bytes, stage_bytes;
  uint32_t idesc, tmem_cols;
};
constexpr int WG_TC_THREADS = 192;        // warp 0: TMA producer, warp 1: TMEM + MMA issue, warps 2-5: epilogue

__global__ void __launch_bounds__(WG_TC_THREADS, 1)
k_wgrad_tc(const __grid_constant__ CUtensorMap map_s_hi, const __grid_constant__ CUtensorMap map_s_lo,
           const __grid_constant__ CUtensorMap map_x_hi, const __grid_constant__ CUtensorMap map_x_lo, const WgParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen_base = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bars = smem_base + (uint32_t)p.STAGES * p.stage_bytes;      // full[S], empty[S], done
  const uint32_t done_bar = bars + 2 * p.STAGES * 8;
  const uint32_t tmem_slot = done_bar + 8;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen_base + (tmem_slot - smem_base));
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;

  // tile: (m tile, n tile, tap group); slice of the K (position box) axis
  int tile = blockIdx.x;
  const int tg = tile % p.tap_groups; tile /= p.tap_groups;
  const int n0 = (tile % p.tiles_n) * p.BN;
  const int m0 = (tile / p.tiles_n) * 128;
  const int tap0 = tg * p.TG;
  const int ntap = min(p.TG, p.T - tap0);
  const int box0 = blockIdx.y * p.boxes_per_slice;
  const int nbox = min(p.boxes_per_slice, p.total_boxes - box0);

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.STAGES; ++s) {
      mbar_init(bars + s * 8, 1);
      mbar_init(bars + (p.STAGES + s) * 8, 1);
    }
    mbar_init(done_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_s_hi)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_s_lo)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_x_hi)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_x_lo)) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot_ptr, 0);

  // shared operand S = dY (channel origin: m0 or n0), shifted operand X (the other origin)
  const int s_c0 = p.swap ? n0 : m0, x_c0 = p.swap ? m0 : n0;
  const uint32_t s_plane = (uint32_t)p.s_sub * p.sub_bytes, x_plane = (uint32_t)p.x_sub * p.sub_bytes;
  const uint32_t x_off = 2u * s_plane;                 // stage layout: [S hi][S lo] then per tap [X hi][X lo]

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (lane == 0 && nbox > 0) {
      const uint32_t tx = 2u * s_plane + (uint32_t)ntap * 2u * x_plane;
      uint32_t s = 0, ph_bit = 1;
      for (int b = box0; b < box0 + nbox; ++b) {
        int t = b;
        const int w0 = (t % p.nbw) * p.bw; t /= p.nbw;
        const int h0 = (t % p.nbh) * p.bh; t /= p.nbh;
        const int d0 = (t % p.nbd) * p.bd;
        const int ns = (t / p.nbd) * p.bn;
        mbar_wait(bars + (p.STAGES + s) * 8, ph_bit);
        const uint32_t sa = smem_base + s * p.stage_bytes, fb = bars + s * 8;
        mbar_expect_tx(fb, tx);
        for (int j = 0; j < p.s_sub; ++j) {
          tma_load_5d(sa + j * p.sub_bytes, &map_s_hi, fb, s_c0 + 64 * j, w0, h0, d0, ns);
          tma_load_5d(sa + s_plane + j * p.sub_bytes, &map_s_lo, fb, s_c0 + 64 * j, w0, h0, d0, ns);
        }
        for (int i = 0; i < ntap; ++i) {
          const int tap = tap0 + i;
          const int kw = tap % p.KW, kh = (tap / p.KW) % p.KH, kd = tap / (p.KW * p.KH);
          const int ws = w0 + kw - p.KW / 2, hs = h0 + kh - p.KH / 2, ds = d0 + kd - p.KD / 2;
          const uint32_t xb = sa + x_off + (uint32_t)i * 2u * x_plane;
          for (int j = 0; j < p.x_sub; ++j) {
            tma_load_5d(xb + j * p.sub_bytes, &map_x_hi, fb, x_c0 + 64 * j, ws, hs, ds, ns);
            tma_load_5d(xb + x_plane + j * p.sub_bytes, &map_x_lo, fb, x_c0 + 64 * j, ws, hs, ds, ns);
          }
        }
        if (++s == (uint32_t)p.STAGES) { s = 0; ph_bit ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer (whole warp converged, one elected lane)
    if (nbox > 0) {
      // MN-major SWIZZLE_128B descriptors: SBO = 1 024 B between 8-position groups, LBO = one 64-channel sub-tile
      const uint32_t dhi = desc_hi_word(1024u, 2u);
      const uint32_t lbo = ((p.sub_bytes >> 4) & 0x3FFFu) << 16;
      auto dlo = [&](uint32_t addr) { return ((addr & 0x3FFFFu) >> 4) | lbo; };
      const int ksteps = p.kblk / 16;
      uint32_t s = 0, ph_bit = 0;
      for (int b = 0; b < nbox; ++b) {
        mbar_wait(bars + s * 8, ph_bit);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t sa = smem_base + s * p.stage_bytes;
        const uint32_t s_hi = dlo(sa), s_lo = dlo(sa + s_plane);
        for (int i = 0; i < ntap; ++i) {
          const uint32_t xb = sa + x_off + (uint32_t)i * 2u * x_plane;
          const uint32_t x_hi = dlo(xb), x_lo = dlo(xb + x_plane);
          const uint32_t a_hi = p.swap ? x_hi : s_hi, a_lo = p.swap ? x_lo : s_lo;
          const uint32_t b_hi = p.swap ? s_hi : x_hi, b_lo = p.swap ? s_lo : x_lo;
          const uint32_t acc = tmem_base + (uint32_t)(i * p.BN);
          for (int k = 0; k < ksteps; ++k) {
            const uint32_t ko = (uint32_t)k * 128u;          // 16 positions x 128 B, in 16-byte units
            umma_lean_e(acc, a_lo + ko, b_hi + ko, dhi, p.idesc, (b > 0 || k > 0) ? 1u : 0u);
            umma_lean_e(acc, a_hi + ko, b_lo + ko, dhi, p.idesc, 1u);
            umma_lean_e(acc, a_hi + ko, b_hi + ko, dhi, p.idesc, 1u);
          }
        }
        umma_commit_e(bars + (p.STAGES + s) * 8);            // frees the stage once its MMAs have read it
        if (++s == (uint32_t)p.STAGES) { s = 0; ph_bit ^= 1u; }
      }
      umma_commit_e(done_bar);
    }
  } else if (nbox > 0) {
    // ===================================================================== epilogue: TMEM -> fp32 RED into dW[co][tap][ci]
    const int q = warp & 3;                                   // TMEM lane quarter this warp may read
    const int m = m0 + q * 32 + lane;
    mbar_wait(done_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int CM = p.swap ? p.Cin : p.Cout, CN = p.swap ? p.Cout : p.Cin;
    for (int i = 0; i < ntap; ++i) {
      const int tap = tap0 + i;
      for (int c = 0; c < p.BN; c += 32) {
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(i * p.BN + c), v);
        if (m >= CM) continue;
        if (!p.swap) {          // thread = one output channel, 32 consecutive input channels: 16-byte vector REDs
          float* row = p.dw + ((int64_t)m * p.T + tap) * p.Cin;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const int ci = n0 + c + j;
            if (ci + 3 < CN) atomicAdd(reinterpret_cast<float4*>(row + ci), make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
            else
              for (int e = 0; e < 4; ++e)
                if (ci + e < CN) atomicAdd(row + ci + e, v[j + e]);
          }
        } else {                // thread = one input channel: lanes of a warp are consecutive addresses for every column
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int co = n0 + c + j;
            if (co < CN) atomicAdd(p.dw + ((int64_t)co * p.T + tap) * p.Cin + m, v[j]);
          }
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

// cuTensorMapEncodeTiled costs 1-2 us and a convolution launch needs up to 7 maps: at batch 1 that is a third of the
// eager launch cost.  Descriptors depend only on (pointer, extents, strides, box, type, swizzle), and PyTorch's caching
// allocator hands the same pointers out step after step, so the encoded maps are kept in a small per-process table
// (bounded; cleared wholesale when full; mp_release_caches() empties it).
struct MapKey {
  uint64_t w[24];
  bool operator==(const MapKey& o) const { return memcmp(w, o.w, sizeof(w)) == 0; }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    uint64_t h = 1469598103934665603ull;
    for (uint64_t x : k.w) { h ^= x; h *= 1099511628211ull; }
    return (size_t)h;
  }
};
std::unordered_map<MapKey, CUtensorMap, MapKeyHash>& map_cache() {
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> c;
  return c;
}
std::mutex g_map_mu;

CUresult encode_cached(CUtensorMap* m, CUtensorMapDataType dt, cuuint32_t rank, void* ptr, const cuuint64_t* dims,
                       const cuuint64_t* strides, const cuuint32_t* box, const cuuint32_t* estr, CUtensorMapInterleave il,
                       CUtensorMapSwizzle swz, CUtensorMapL2promotion l2, CUtensorMapFloatOOBfill oob) {
  static int enabled = [] { const char* e = getenv("MPB200_NO_MAP_CACHE"); return (e && atoi(e)) ? 0 : 1; }();
  if (!enabled) return get_encoder()(m, dt, rank, ptr, dims, strides, box, estr, il, swz, l2, oob);
  MapKey k;
  memset(&k, 0, sizeof(k));
  int dev = 0;
  cudaGetDevice(&dev);
  k.w[0] = reinterpret_cast<uint64_t>(ptr);
  k.w[1] = (uint64_t)dt | ((uint64_t)rank << 8) | ((uint64_t)il << 16) | ((uint64_t)swz << 24) | ((uint64_t)l2 << 32) |
           ((uint64_t)oob << 40) | ((uint64_t)dev << 48);
  for (cuuint32_t i = 0; i < rank; ++i) {
    k.w[2 + i] = dims[i];
    k.w[8 + i] = i + 1 < rank ? strides[i] : 0;
    k.w[14 + i] = ((uint64_t)box[i] << 32) | estr[i];
  }
  {
    std::lock_guard<std::mutex> lock(g_map_mu);
    auto it = map_cache().find(k);
    if (it != map_cache().end()) { *m = it->second; return CUDA_SUCCESS; }
  }
  CUresult r = get_encoder()(m, dt, rank, ptr, dims, strides, box, estr, il, swz, l2, oob);
  if (r == CUDA_SUCCESS) {
    std::lock_guard<std::mutex> lock(g_map_mu);
    if (map_cache().size() >= 16384) map_cache().clear();
    map_cache().emplace(k, *m);
  }
  return r;
}

struct Plan {
  TcParams p;
  int tiles_m, tiles_n;
  uint32_t smem_bytes;
  CUtensorMapSwizzle swz;
  bool slab;
  SlabExtra x;
};

bool allow_dual() {
  static int v = [] { const char* e = getenv("MPB200_TC_NO_DUALB"); return (e && atoi(e)) ? 0 : 1; }();
  return v != 0;
}

bool allow_tma_epi() {
  static int v = [] { const char* e = getenv("MPB200_TC_NO_TMA_EPI"); return (e && atoi(e)) ? 0 : 1; }();
  return v != 0;
}

// TMA-store epilogue: fp16 output plane only, whole 64-channel sub-tiles, 16-byte aligned channel windows
bool tma_epi_ok(const mp_conv_desc* d, int bn) {
  const int out_C = d->out_C > 0 ? d->out_C : d->Cout;
  // worth it where the epilogue, not the main loop, bounds a tile: K = taps*Cin up to 1152
  if ((int64_t)d->KD * d->KH * d->KW * d->Cin > 1152) return false;
  if (d->out_fmt > MP_FMT_F16 || d->res_fmt > MP_FMT_F16 || d->out_fmt == MP_FMT_SPLIT_BF16 || d->res_fmt == MP_FMT_SPLIT_BF16)
    return false;
  return allow_tma_epi() && d->prec == MP_PREC_F16X2 && d->out_hi && !d->out_f32 && !d->stats && bn % 64 == 0 &&
         d->Cout == d->Cout_pad && out_C % 8 == 0 && d->out_c_off % 8 == 0 &&
         (!d->res_hi || (reinterpret_cast<uintptr_t>(d->res_hi) & 15) == 0) && !d->res_f32;
}

int g_num_sms[64] = {0};

int next_pow2(int v) { int r = 32; while (r < v) r <<= 1; return r; }

// returns 0 and fills plan when the shape is supported
int make_plan(const mp_conv_desc* d, Plan& pl, bool report) {
  TcParams& p = pl.p;
  auto fail = [&](const char* why) { return report ? mp_set_error("mp_conv_tc: unsupported shape: %s", why) : 1; };
  if (d->Cin % 16 != 0) return fail("Cin % 16 != 0");
  if (d->Cout_pad % 16 != 0) return fail("Cout_pad % 16 != 0");
  if (d->prec != MP_PREC_SPLIT_BF16 && d->prec != MP_PREC_F16X2 && d->prec != MP_PREC_F16_Q8) return fail("unknown prec");
  const bool f16x2 = d->prec == MP_PREC_F16X2;
  const bool f16q8 = d->prec == MP_PREC_F16_Q8;
  if (f16q8 && (d->Cin % 64 || d->Cin2 % 64)) return fail("fp16 + fp8 mode needs input channels in multiples of 64");
  {
    const int native = f16q8 ? MP_FMT_F16_Q8 : f16x2 ? MP_FMT_F16 : MP_FMT_SPLIT_BF16;
    p.out_fmt = d->out_fmt ? d->out_fmt : native;
    p.res_fmt = d->res_fmt ? d->res_fmt : native;
    if (p.out_fmt < MP_FMT_SPLIT_BF16 || p.out_fmt > MP_FMT_F16_Q8 || p.res_fmt < MP_FMT_SPLIT_BF16 || p.res_fmt > MP_FMT_F16_Q8)
      return fail("unknown out_fmt / res_fmt");
    const int out_C = d->out_C > 0 ? d->out_C : d->Cout;
    if ((p.out_fmt == MP_FMT_F16_Q8 && d->out_hi) || (p.res_fmt == MP_FMT_F16_Q8 && d->res_hi))
      if (out_C % 64 || d->out_c_off % 64 || d->Cout % 64) return fail("the fp16 + fp8 plane format needs 64-channel groups");
  }
  p.prec = d->prec;
  p.a_planes = f16x2 ? 1u : 2u;
  p.lo_scale = f16x2 ? 1.0f / 2048.0f : f16q8 ? d->corr_scale : 1.0f;
  p.acc_scale = ((f16x2 || f16q8) && d->acc_scale > 0.f) ? d->acc_scale : 1.0f;
  p.out_q8_scale = d->out_q8_scale > 0.f ? d->out_q8_scale : 1.0f;
  p.res_q8_inv = 1.0f / (2048.0f * (d->res_q8_scale > 0.f ? d->res_q8_scale : 1.0f));
  const uint32_t ap = p.a_planes;
  // fused 1x1 shortcut (second source appended to K)
  if (d->Cin2 < 0 || (d->Cin2 > 0 && (!d->in2_hi || (!f16x2 && !d->in2_lo)))) return fail("bad second source");
  if (d->Cin2 % 16 != 0) return fail("Cin2 % 16 != 0");
  p.stride2 = d->stride2 > 0 ? d->stride2 : 1;
  p.in2_c_off = d->in2_c_off;
  if (p.stride2 != 1 && p.stride2 != 2) return fail("stride2 must be 1 or 2");
  if (d->Cin2 > 0 && (d->in2_c_off % 8 || (d->in2_C > 0 && d->in2_C < d->in2_c_off + d->Cin2))) return fail("bad second-source channel window");
  p.k2_off = d->KD * d->KH * d->KW * d->Cin;
  // kind::f16 instruction descriptor: D = f32 (bit 4); A/B format bf16 = 1 at bits 7 / 10, fp16 = 0
  const uint32_t idesc_fmt = (1u << 4) | ((f16x2 || f16q8) ? 0u : ((1u << 7) | (1u << 10)));
  {
    const int cboth = d->Cin | d->Cin2;          // the channel chunk must divide both sources
    p.CCHUNK = (cboth % 64 == 0) ? 64 : (cboth % 32 == 0) ? 32 : 16;
  }
  p.layout_type = p.CCHUNK == 64 ? 2u : p.CCHUNK == 32 ? 4u : 6u;
  pl.swz = p.CCHUNK == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : p.CCHUNK == 32 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                                         : CU_TENSOR_MAP_SWIZZLE_32B;
  p.sbo = 8u * p.CCHUNK * 2u;
  p.num_cchunks = d->Cin / p.CCHUNK;
  // ABI v2 fields
  p.stride = d->stride > 0 ? d->stride : 1;
  p.in_c_off = d->in_c_off;
  p.out_C = d->out_C > 0 ? d->out_C : d->Cout;
  p.out_c_off = d->out_c_off;
  if (p.stride != 1 && p.stride != 2) return fail("stride must be 1 or 2");
  if (d->H % p.stride || d->W % p.stride) return fail("H, W must be multiples of the stride");
  if (p.in_c_off % 8) return fail("in_c_off must be a multiple of 8 (16-byte TMA coordinate granularity)");
  const int Ho = d->H / p.stride, Wo = d->W / p.stride;
  // output tile box
  auto pow2_le = [](int v, int cap) { int r = 1; while (r * 2 <= v && r * 2 <= cap) r *= 2; return r; };
  p.BW = pow2_le(Wo, TILE_M);
  p.BH = pow2_le(Ho, TILE_M / p.BW);
  p.BD = pow2_le(d->D, TILE_M / (p.BW * p.BH));
  if (Wo % p.BW || Ho % p.BH || d->D % p.BD) return fail("grid not divisible by the tile box");
  p.tiles_w = Wo / p.BW; p.tiles_h = Ho / p.BH; p.tiles_d = d->D / p.BD;
  p.N = d->N;
  p.BNb = 1;
  if (p.BW * p.BH * p.BD != TILE_M) {
    // fewer than 128 positions per sample (FlowField tower, model.py:446-456): the tile box spans several samples
    const int sbox = p.BW * p.BH * p.BD;
    if (TILE_M % sbox || p.tiles_w * p.tiles_h * p.tiles_d != 1) return fail("sample grid is not a power-of-two box");
    if (d->stats) return fail("fused GroupNorm statistics need >= 128 positions per sample");
    p.BNb = TILE_M / sbox;
  }
  pl.tiles_m = ((d->N + p.BNb - 1) / p.BNb) * p.tiles_d * p.tiles_h * p.tiles_w;
  p.tap_mask = 0;
  {
    const int pd = d->KD / 2, ph = d->KH / 2, pw = d->KW / 2;
    for (int kd = 0; kd < d->KD; ++kd)
      for (int kh = 0; kh < d->KH; ++kh)
        for (int kw = 0; kw < d->KW; ++kw) {
          const bool live = abs(kd - pd) < d->D && abs(kh - ph) < d->H && abs(kw - pw) < d->W;
          if (live) p.tap_mask |= 1ull << ((kd * d->KH + kh) * d->KW + kw);
        }
  }
  if (d->KD * d->KH * d->KW > 64) return fail("more than 64 filter taps");
  // N tile: largest multiple-of-16 divisor of Cout_pad up to the cap, shrunk while the grid under-fills the GPU
  static int bn_cap = [] { const char* e = getenv("MPB200_TC_BN_MAX"); int v = e ? atoi(e) : 256; return v < 16 ? 16 : (v > 256 ? 256 : v); }();
  const int cands[] = {256, 192, 128, 96, 64, 48, 32, 16};
  p.BN = 0;
  for (int c : cands)
    if (c <= bn_cap && (!(f16x2 || f16q8) || c <= 128) && d->Cout_pad % c == 0) { p.BN = c; break; }   // two column halves
  if (!p.BN) return fail("no N tile");
  while ((int64_t)pl.tiles_m * (d->Cout_pad / p.BN) < 148 && p.BN % 32 == 0 && p.BN > 32) p.BN /= 2;
  pl.tiles_n = d->Cout_pad / p.BN;
  pl.slab = false;
  // ---- slab (vertical-halo reuse) kernel for 3x3(x3) convolutions with <= 128 output channels
  static int allow_slab = [] { const char* e = getenv("MPB200_TC_NO_SLAB"); return (e && atoi(e)) ? 0 : 1; }();
  if (allow_slab && !f16q8 && p.stride == 1 && d->KH == 3 && d->KW == 3 && (d->KD == 1 || d->KD == 3) &&
      d->Cout_pad <= 128 && d->W % 8 == 0 && d->H % 16 == 0) {
    const int bn = d->Cout_pad;
    const bool tma_s = tma_epi_ok(d, bn);
    static int nbuf_env = [] { const char* e = getenv("MPB200_TC_TMA_NBUF"); return e ? atoi(e) : 0; }();
    const uint32_t nbuf_s = nbuf_env ? (uint32_t)nbuf_env : (bn <= 64 ? 2u : 1u);      // 32 KB of staging either way
    const uint32_t tma_bytes_s = tma_s ? nbuf_s * (uint32_t)(bn / 64) * 16384u + 1024u /*alignment*/ : 0u;
    const uint32_t fixed_s = 1024 + 1024 + 2 * 256 * sizeof(double) + 2048 /*s_part*/ + (tma_s ? 0u : EPI_BYTES) + tma_bytes_s;
    static int mt_cap = [] { const char* e = getenv("MPB200_TC_SLAB_MT"); return e ? atoi(e) : 2; }();
    static int cc_force = [] { const char* e = getenv("MPB200_TC_SLAB_CC"); return e ? atoi(e) : 0; }();
    static int sb_want = [] { const char* e = getenv("MPB200_TC_SLAB_SB"); int v = e ? atoi(e) : 3; return v < 2 ? 2 : v; }();
    // The weight ring must be deep enough to cover the TMA latency (a B tile feeds only MT*CCHUNK/16*2 MMAs):
    // prefer the widest channel chunk that still leaves >= sb_want weight stages, else the deepest ring found.
    // Measured on B200 (profiles/): the widest channel chunk wins (64-byte TMA rows and twice the barrier traffic cost
    // more than a deeper weight ring buys), two accumulators per weight tile win over one; so: widest chunk first,
    // then MT = 2 before MT = 1, then the deepest A ring that leaves >= sb_want weight stages.
    auto choose = [&](int mt_first) -> bool {
      int best_cc = 0, best_sa = 0, best_sb = 0, best_mt = 0;
      for (int cc = (cc_force ? cc_force : p.CCHUNK); cc >= 16 && !best_cc; cc = (cc_force ? 0 : cc / 2)) {
        if (d->Cin % cc || d->Cin2 % cc) continue;
        for (int mt_try = mt_first; mt_try >= 1 && !best_cc; --mt_try) {
          const uint32_t b_stage = 2u * bn * cc * 2u;
          const uint32_t a_plane = (uint32_t)(mt_try * 16 + 2) * 8u * cc * 2u;
          for (int sa = 3; sa >= 2; --sa) {
            if (fixed_s + sa * ap * a_plane + sb_want * b_stage > SMEM_LIMIT) continue;
            int sb = (int)((SMEM_LIMIT - fixed_s - sa * ap * a_plane) / b_stage);
            if (sb > 12) sb = 12;
            best_sb = sb; best_cc = cc; best_sa = sa; best_mt = mt_try;
            break;
          }
        }
      }
      const int mt = best_mt;
      if (!best_cc) return false;
      const int cc = best_cc, sa = best_sa, sb = best_sb;
      const uint32_t b_stage = 2u * bn * cc * 2u;
      const uint32_t a_plane = (uint32_t)(mt * 16 + 2) * 8u * cc * 2u;
      pl.slab = true;
      pl.x.MT = mt; pl.x.SA = sa; pl.x.SB = sb;
      pl.x.a_plane_bytes = a_plane;
      pl.x.b_ring_off = sa * ap * a_plane;
      p.CCHUNK = cc;
      p.layout_type = cc == 64 ? 2u : cc == 32 ? 4u : 6u;
      pl.swz = cc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : cc == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
      p.sbo = 8u * cc * 2u;
      p.num_cchunks = d->Cin / cc;
      p.num_cchunks2 = d->Cin2 / cc;
      p.BN = bn;
      pl.tiles_n = 1;
      p.BW = 8; p.BH = 16; p.BD = 1; p.BNb = 1;
      p.tiles_w = d->W / 8; p.tiles_h = d->H / (16 * mt); p.tiles_d = d->D;
      pl.tiles_m = d->N * p.tiles_d * p.tiles_h * p.tiles_w;
      p.a_bytes = a_plane;
      p.b_bytes = (uint32_t)bn * cc * 2u;
      p.stage_bytes = 0;
      p.STAGES = sa;
      p.b_resident = 0;
      p.bres_off = 0;
      p.epi_tma = tma_s ? 1 : 0;
      p.tma_off = (pl.x.b_ring_off + sb * b_stage + 1023u) & ~1023u;
      p.tma_buf_bytes = tma_s ? (uint32_t)(bn / 64) * 16384u : 0u;
      p.tma_nbuf = nbuf_s;
      p.epi_bytes = tma_s ? 0u : EPI_BYTES;
      p.epi_off = tma_s ? p.tma_off + p.tma_nbuf * p.tma_buf_bytes : pl.x.b_ring_off + sb * b_stage;
      pl.smem_bytes = fixed_s + pl.x.b_ring_off + sb * b_stage;
      p.dualb = ((allow_dual() || f16x2) && 2 * mt * 2 * bn <= 512) ? 1 : 0;
      p.acc_w = p.dualb ? 2 * bn : bn;
      p.tmem_cols = next_pow2(2 * mt * p.acc_w);
      p.tiles_n = 1;
      p.total_tiles = pl.tiles_m;
      return true;
    };
    // fp16 two-pass mode always uses the [Bh|Bl] accumulator layout (2*bn TMEM columns per accumulator, two tiles in
    // flight), so two accumulators per weight tile fit only up to bn = 64
    const bool mt2_ok = d->H % 32 == 0 && mt_cap >= 2 && 2 * 2 * bn <= 512 && (!f16x2 || 2 * 2 * 2 * bn <= 512);
    choose(mt2_ok ? 2 : 1);
  }
  if (pl.slab) {
    if (f16x2 && !p.dualb) return fail("fp16 two-pass mode needs the [Bh|Bl] column layout (BN <= 128)");
    p.idesc = idesc_fmt | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
    p.idesc2 = idesc_fmt | ((uint32_t)(2 * p.BN >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
    p.D = d->D; p.H = Ho; p.W = Wo; p.Cin = d->Cin; p.Cout = d->Cout;
    p.KD = d->KD; p.KH = d->KH; p.KW = d->KW;
    p.bias = d->bias; p.res_f32 = d->res_f32; p.res_hi = (const bf16*)d->res_hi; p.res_lo = (const bf16*)d->res_lo;
    p.out_f32 = d->out_f32; p.out_hi = (bf16*)d->out_hi; p.out_lo = (bf16*)d->out_lo;
    p.stats = d->stats; p.gn_groups = d->gn_groups; p.act = d->act;
    p.S = (int64_t)d->D * Ho * Wo;
    return 0;
  }
  const bool tma_g = tma_epi_ok(d, p.BN);
  // two staging buffers where shared memory is plentiful (narrow tiles, tiny K); one otherwise
  const uint32_t nbuf_g = (p.BN <= 64 || (int64_t)d->KD * d->KH * d->KW * d->Cin <= 128) ? 2u : 1u;
  const uint32_t tma_bytes_g = tma_g ? nbuf_g * (uint32_t)(p.BN / 64) * 16384u + 1024u /*alignment*/ : 0u;
  const uint32_t fixed = 1024 /*align slack*/ + 1024 /*barriers, tmem slot*/ + 2 * 256 * sizeof(double) +
                         (tma_g ? 0u : EPI_BYTES) + 2048 /*s_part*/ + tma_bytes_g;
  // weight-resident mode: one N tile and the whole [taps*Cin x BN] weight tile (hi+lo) fits beside >= 3 A stages;
  // try the widest channel chunk first, then narrower ones (smaller A stages).
  static int allow_res = [] { const char* e = getenv("MPB200_TC_NO_BRES"); return (e && atoi(e)) ? 0 : 1; }();
  p.b_resident = 0;
  const uint32_t ktot = (uint32_t)d->KD * d->KH * d->KW * d->Cin + (uint32_t)d->Cin2;
  const uint32_t bres_bytes = ktot * (uint32_t)p.BN * 4u;
  const int ntaps = d->KD * d->KH * d->KW;
  const bool all_taps = p.tap_mask == (ntaps == 64 ? ~0ull : ((1ull << ntaps) - 1ull));
  if (allow_res && all_taps && pl.tiles_n == 1 && pl.tiles_m >= 4 * 148 && bres_bytes < SMEM_LIMIT) {
    for (int cc = p.CCHUNK; cc >= (f16q8 ? 64 : 16); cc /= 2) {       // (fp16 + fp8 mode: 64-channel chunks only)
      const uint32_t a_stage = ap * TILE_M * cc * 2u;
      if (fixed + bres_bytes + 3 * a_stage <= SMEM_LIMIT && bres_bytes < (1u << 20)) {
        p.b_resident = 1;
        p.CCHUNK = cc;
        p.layout_type = cc == 64 ? 2u : cc == 32 ? 4u : 6u;
        pl.swz = cc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : cc == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
        p.sbo = 8u * cc * 2u;
        p.num_cchunks = d->Cin / cc;
        break;
      }
    }
  }
  p.num_cchunks2 = d->Cin2 / p.CCHUNK;
  p.a_bytes = TILE_M * p.CCHUNK * 2;
  p.b_bytes = p.BN * p.CCHUNK * 2;
  p.stage_bytes = p.b_resident ? ap * p.a_bytes : ap * p.a_bytes + 2 * p.b_bytes;
  const uint32_t avail = SMEM_LIMIT - fixed - (p.b_resident ? bres_bytes : 0);
  int stages = (int)(avail / p.stage_bytes);
  if (stages > 8) stages = 8;
  if (stages < 2) return fail("tile does not fit shared memory");
  p.STAGES = stages;
  pl.smem_bytes = fixed + stages * p.stage_bytes + (p.b_resident ? bres_bytes : 0);
  p.bres_off = stages * p.stage_bytes;
  p.epi_tma = tma_g ? 1 : 0;
  p.tma_off = (p.bres_off + (p.b_resident ? bres_bytes : 0) + 1023u) & ~1023u;
  p.tma_buf_bytes = tma_g ? (uint32_t)(p.BN / 64) * 16384u : 0u;
  p.tma_nbuf = nbuf_g;
  p.epi_bytes = tma_g ? 0u : EPI_BYTES;
  p.epi_off = tma_g ? p.tma_off + p.tma_nbuf * p.tma_buf_bytes : p.bres_off + (p.b_resident ? bres_bytes : 0);
  p.dualb = ((allow_dual() || f16x2 || f16q8) && 2 * 2 * p.BN <= 512) ? 1 : 0;
  if ((f16x2 || f16q8) && !p.dualb) return fail("fp16 modes need the two-half accumulator layout (BN <= 128)");
  if (f16q8 && p.CCHUNK != 64) return fail("fp16 + fp8 mode needs 64-channel K chunks");
  p.acc_w = p.dualb ? 2 * p.BN : p.BN;
  p.tmem_cols = next_pow2(2 * p.acc_w);
  p.idesc2 = idesc_fmt | ((uint32_t)(2 * p.BN >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
  p.tiles_n = pl.tiles_n;
  p.total_tiles = pl.tiles_m * pl.tiles_n;
  p.idesc = idesc_fmt | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
  p.D = d->D; p.H = Ho; p.W = Wo; p.Cin = d->Cin; p.Cout = d->Cout;
  p.KD = d->KD; p.KH = d->KH; p.KW = d->KW;
  p.bias = d->bias; p.res_f32 = d->res_f32; p.res_hi = (const bf16*)d->res_hi; p.res_lo = (const bf16*)d->res_lo;
  p.out_f32 = d->out_f32; p.out_hi = (bf16*)d->out_hi; p.out_lo = (bf16*)d->out_lo;
  p.stats = d->stats; p.gn_groups = d->gn_groups; p.act = d->act;
  p.S = (int64_t)d->D * Ho * Wo;
  if ((int64_t)pl.tiles_m * pl.tiles_n > 0x7fffffff) return fail("too many tiles");
  return 0;
}

int encode_act_map(CUtensorMap* m, const void* ptr, const mp_conv_desc* d, const Plan& pl) {
  const cuuint64_t C = (cuuint64_t)(d->in_C > 0 ? d->in_C : d->Cin);
  const cuuint32_t st = (cuuint32_t)pl.p.stride;
  cuuint64_t dims[5] = {C, (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->D, (cuuint64_t)d->N};
  cuuint64_t strides[4] = {C * 2, C * 2 * d->W, C * 2 * d->W * d->H, C * 2 * d->W * d->H * d->D};
  // with elementStrides = s the box spans BW*s input elements and TMA keeps every s-th one (BW of them)
  cuuint32_t box[5] = {(cuuint32_t)pl.p.CCHUNK, (cuuint32_t)pl.p.BW * st,
                       (cuuint32_t)(pl.slab ? pl.x.MT * pl.p.BH + 2 : pl.p.BH * st), (cuuint32_t)pl.p.BD,
                       (cuuint32_t)(pl.slab ? 1 : pl.p.BNb)};
  cuuint32_t estr[5] = {1, st, st, 1, 1};
  const CUtensorMapDataType dt = pl.p.prec != MP_PREC_SPLIT_BF16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUresult r = encode_cached(m, dt, 5, const_cast<void*>(ptr), dims, strides, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, pl.swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : mp_set_error("mp_conv_tc: cuTensorMapEncodeTiled(activation) failed (%d)", (int)r);
}

// fp16 output plane as a 5-D tensor (C, W, H, D, N) of OUTPUT extents; box = one 64-channel sub-tile of the tile box
int encode_out_map(CUtensorMap* m, void* ptr, const mp_conv_desc* d, const Plan& pl) {
  const TcParams& p = pl.p;
  const cuuint64_t C = (cuuint64_t)p.out_C;
  cuuint64_t dims[5] = {C, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.D, (cuuint64_t)d->N};
  cuuint64_t strides[4] = {C * 2, C * 2 * p.W, C * 2 * p.W * p.H, C * 2 * p.W * p.H * p.D};
  cuuint32_t box[5] = {64, (cuuint32_t)p.BW, (cuuint32_t)p.BH, (cuuint32_t)p.BD, (cuuint32_t)(pl.slab ? 1 : p.BNb)};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = encode_cached(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, ptr, dims, strides, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : mp_set_error("mp_conv_tc: cuTensorMapEncodeTiled(output) failed (%d)", (int)r);
}

// second source of the fused shortcut: [N, D, Ho*stride2, Wo*stride2, in2_C]; same tile box as the main operand, centre tap
int encode_act2_map(CUtensorMap* m, const void* ptr, const mp_conv_desc* d, const Plan& pl) {
  const TcParams& p = pl.p;
  const cuuint64_t C = (cuuint64_t)(d->in2_C > 0 ? d->in2_C : d->Cin2);
  const cuuint32_t st = (cuuint32_t)p.stride2;
  const cuuint64_t W2 = (cuuint64_t)p.W * st, H2 = (cuuint64_t)p.H * st;       // p.W / p.H are OUTPUT extents
  cuuint64_t dims[5] = {C, W2, H2, (cuuint64_t)d->D, (cuuint64_t)d->N};
  cuuint64_t strides[4] = {C * 2, C * 2 * W2, C * 2 * W2 * H2, C * 2 * W2 * H2 * d->D};
  cuuint32_t box[5] = {(cuuint32_t)p.CCHUNK, (cuuint32_t)p.BW * st,
                       (cuuint32_t)((pl.slab ? pl.x.MT * p.BH + 2 : p.BH) * st), (cuuint32_t)p.BD,
                       (cuuint32_t)(pl.slab ? 1 : p.BNb)};
  cuuint32_t estr[5] = {1, st, st, 1, 1};
  const CUtensorMapDataType dt = p.prec != MP_PREC_SPLIT_BF16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUresult r = encode_cached(m, dt, 5, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             pl.swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : mp_set_error("mp_conv_tc: cuTensorMapEncodeTiled(second source) failed (%d)", (int)r);
}

int encode_w_map(CUtensorMap* m, const void* ptr, const mp_conv_desc* d, const Plan& pl, int box_rows = 0) {
  const cuuint64_t ktot = (cuuint64_t)d->KD * d->KH * d->KW * d->Cin + (cuuint64_t)d->Cin2;
  if (pl.p.b_merged) {     // both planes in one box: (K chunk, BN rows, 2 planes) -> [Bh tile | Bl tile] in shared memory
    const CUtensorMapDataType dt3 = pl.p.prec != MP_PREC_SPLIT_BF16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    cuuint64_t dims3[3] = {ktot, (cuuint64_t)d->Cout_pad, 2};
    cuuint64_t strides3[2] = {ktot * 2, ktot * 2 * (cuuint64_t)d->Cout_pad};
    cuuint32_t box3[3] = {(cuuint32_t)pl.p.CCHUNK, (cuuint32_t)(box_rows ? box_rows : pl.p.BN), 2};
    cuuint32_t estr3[3] = {1, 1, 1};
    CUresult r3 = encode_cached(m, dt3, 3, const_cast<void*>(ptr), dims3, strides3, box3, estr3, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                pl.swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r3 == CUDA_SUCCESS ? 0 : mp_set_error("mp_conv_tc: cuTensorMapEncodeTiled(weights, 3-D) failed (%d)", (int)r3);
  }
  cuuint64_t dims[2] = {ktot, (cuuint64_t)d->Cout_pad};
  cuuint64_t strides[1] = {ktot * 2};
  cuuint32_t box[2] = {(cuuint32_t)pl.p.CCHUNK, (cuuint32_t)pl.p.BN};
  cuuint32_t estr[2] = {1, 1};
  const CUtensorMapDataType dt = pl.p.prec != MP_PREC_SPLIT_BF16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUresult r = encode_cached(m, dt, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, pl.swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : mp_set_error("mp_conv_tc: cuTensorMapEncodeTiled(weights) failed (%d)", (int)r);
}

int launch_pair(const mp_conv_desc* d, Plan pl, void* stream) {
  TcParams& p = pl.p;
  p.b_bytes = (uint32_t)(p.BN / 2) * p.CCHUNK * 2u;            // each CTA stages half of the weight tile
  p.stage_bytes = 2u * p.a_bytes + 2u * p.b_bytes;
  const uint32_t fixed = 1024 + 1024 + 2 * 256 * sizeof(double) + EPI_BYTES + 2048;
  int stages = (int)((SMEM_LIMIT - fixed) / p.stage_bytes);
  if (stages > 8) stages = 8;
  MP_REQUIRE(stages >= 3, "mp_conv_tc: pair kernel does not fit shared memory");
  p.STAGES = stages;
  p.epi_tma = 0;
  p.epi_bytes = EPI_BYTES;
  p.epi_off = (uint32_t)stages * p.stage_bytes;
  p.tmem_cols = 512;                                           // two accumulator pairs of 2 x 128 columns
  p.acc_w = 2 * p.BN;
  p.dualb = 1;
  p.sched = nullptr;
  const uint32_t idesc_fmt = (1u << 4);                        // D = f32, A / B = fp16 (kind::f16) resp. e4m3 (kind::f8f6f4)
  p.idesc = idesc_fmt | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);     // M = 256 across the pair
  const uint32_t smem_bytes = fixed + (uint32_t)stages * p.stage_bytes;
  CUtensorMap ma_hi, ma_lo, mb;
  if (int e = encode_act_map(&ma_hi, d->in_hi, d, pl)) return e;
  if (int e = encode_act_map(&ma_lo, d->in_lo, d, pl)) return e;
  if (int e = encode_w_map(&mb, d->w_hi, d, pl, p.BN / 2)) return e;
  CUtensorMap ma2_hi = ma_hi, ma2_lo = ma_lo;
  if (d->Cin2 > 0) {
    MP_REQUIRE((((uintptr_t)d->in2_hi | (uintptr_t)d->in2_lo) & 15) == 0, "mp_conv_tc: second source must be 16-byte aligned");
    if (int e = encode_act2_map(&ma2_hi, d->in2_hi, d, pl)) return e;
    if (int e = encode_act2_map(&ma2_lo, d->in2_lo, d, pl)) return e;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  static bool attr_done[64] = {false};
  if (dev < 0 || dev >= 64 || !attr_done[dev]) {
    cudaError_t ae = cudaFuncSetAttribute(k_conv_q8_pair, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT);
    MP_REQUIRE(ae == cudaSuccess, "mp_conv_tc: cannot opt in to %u B shared memory: %s", SMEM_LIMIT, cudaGetErrorString(ae));
    if (dev >= 0 && dev < 64) attr_done[dev] = true;
  }
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int items = (pl.tiles_m / 2) * pl.tiles_n;
  int clusters = sms / 2;
  if (clusters > items) clusters = items;
  k_conv_q8_pair<<<2 * clusters, NUM_THREADS3, smem_bytes, mp_stream(stream)>>>(ma_hi, ma_lo, mb, ma2_hi, ma2_lo, p);
  MP_LAUNCH_CHECK("mp_conv_tc (pair)");
  return 0;
}

// One counter slot for a launch on `stream` (NULL = use the static walk).  Rules that keep concurrently running kernels
// apart: (1) eager launches rotate through SCHED_RANGE slots owned by their stream -- launches on one stream are
// serialised, so a slot is idle again long before the rotation returns to it; (2) a launch that is being captured into a
// CUDA graph is baked into that graph for good, so it takes a slot from a bump allocator that never hands it out twice
// (one graph cannot run concurrently with itself); (3) when either pool is exhausted (more than 16 streams, more than
// 32768 captured launches per device) the launch falls back to the static tile walk.
unsigned int* sched_slot(int dev, cudaStream_t stream) {
  struct PerDevice {
    unsigned int* base = nullptr;
    std::unordered_map<cudaStream_t, unsigned> range_of, rot;
    unsigned next_range = 0, next_captured = SCHED_EAGER;
  };
  static PerDevice state[64];
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  PerDevice& st = state[dev];
  if (!st.base) {
    void* sym = nullptr;
    if (cudaGetSymbolAddress(&sym, g_sched) != cudaSuccess) return nullptr;
    st.base = static_cast<unsigned int*>(sym);
  }
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(stream, &cs) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  if (cs != cudaStreamCaptureStatusNone) {
    if (st.next_captured >= SCHED_SLOTS) return nullptr;
    return st.base + 2 * (st.next_captured++);
  }
  auto it = st.range_of.find(stream);
  if (it == st.range_of.end()) {
    if ((st.next_range + 1) * SCHED_RANGE > SCHED_EAGER) return nullptr;
    it = st.range_of.emplace(stream, st.next_range++).first;
  }
  unsigned& r = st.rot[stream];
  const unsigned slot = it->second * SCHED_RANGE + (r++ % SCHED_RANGE);
  return st.base + 2 * slot;
}

}  // namespace

extern "C" int mp_release_caches(void) {
  std::lock_guard<std::mutex> lock(g_map_mu);
  map_cache().clear();
  return 0;
}

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
  {
    static int no_merge = [] { const char* e = getenv("MPB200_TC_NO_BMERGE"); return (e && atoi(e)) ? 1 : 0; }();
    const size_t plane = ((size_t)d->KD * d->KH * d->KW * d->Cin + d->Cin2) * d->Cout_pad * 2;
    pl.p.b_merged = (!no_merge &&
                     reinterpret_cast<const char*>(d->w_lo) == reinterpret_cast<const char*>(d->w_hi) + plane) ? 1 : 0;
  }
  // ---- CTA-pair kernel (cta_group::2) for the fp16 + FP8 convolutions whose tiles pair up (MPB200_TC_PAIR=0 disables it)
  {
    const char* pe = getenv("MPB200_TC_PAIR");          // read per launch: tests flip it inside one process
    const int allow_pair = (pe && !atoi(pe)) ? 0 : 1;
    const TcParams& q = pl.p;
    const int ntaps = d->KD * d->KH * d->KW;
    const bool all_taps = q.tap_mask == (ntaps == 64 ? ~0ull : ((1ull << ntaps) - 1ull));
    if (allow_pair && d->prec == MP_PREC_F16_Q8 && !pl.slab && q.BN == 128 && q.CCHUNK == 64 && q.b_merged &&
        (d->Cin2 == 0 || q.stride2 == 1) && q.stride == 1 && q.BNb == 1 && pl.tiles_m % 2 == 0 && q.in_c_off == 0 && !d->stats && !q.b_resident &&
        all_taps && (d->in_C == 0 || d->in_C == d->Cin))
      return launch_pair(d, pl, stream);
  }
  CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo;
  if (int e = encode_act_map(&ma_hi, d->in_hi, d, pl)) return e;
  if (int e = encode_act_map(&ma_lo, pl.p.a_planes == 2 ? d->in_lo : d->in_hi, d, pl)) return e;   // unused when 1 plane
  if (int e = encode_w_map(&mb_hi, d->w_hi, d, pl)) return e;
  if (pl.p.b_merged) mb_lo = mb_hi;
  else if (int e = encode_w_map(&mb_lo, d->w_lo, d, pl)) return e;
  CUtensorMap ma2_hi = ma_hi, ma2_lo = ma_lo;     // placeholders when there is no fused shortcut
  if (d->Cin2 > 0) {
    MP_REQUIRE((((uintptr_t)d->in2_hi | (uintptr_t)d->in2_lo) & 15) == 0, "mp_conv_tc: second source must be 16-byte aligned");
    if (int e = encode_act2_map(&ma2_hi, d->in2_hi, d, pl)) return e;
    if (int e = encode_act2_map(&ma2_lo, pl.p.a_planes == 2 ? d->in2_lo : d->in2_hi, d, pl)) return e;
  }
  CUtensorMap mo = mb_hi;          // placeholder when the kernel stores with ordinary instructions
  if (pl.p.epi_tma) {
    MP_REQUIRE((reinterpret_cast<uintptr_t>(d->out_hi) & 15) == 0, "mp_conv_tc: output must be 16-byte aligned");
    if (int e = encode_out_map(&mo, d->out_hi, d, pl)) return e;
  }
  {
    static bool attr_done[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !attr_done[dev]) {
      cudaError_t ae = cudaFuncSetAttribute(k_conv_tc2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT);
      if (ae == cudaSuccess)
        ae = cudaFuncSetAttribute(k_conv_tc3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT);
      MP_REQUIRE(ae == cudaSuccess, "mp_conv_tc: cannot opt in to %u B shared memory: %s", SMEM_LIMIT,
                 cudaGetErrorString(ae));
      if (dev >= 0 && dev < 64) {
        attr_done[dev] = true;
        cudaDeviceGetAttribute(&g_num_sms[dev], cudaDevAttrMultiProcessorCount, dev);
      }
    }
    // dynamic tile scheduling (default; MPB200_TC_STATIC=1 restores the static walk)
    static int dyn = [] { const char* e = getenv("MPB200_TC_STATIC"); return (e && atoi(e)) ? 0 : 1; }();
    pl.p.sched = nullptr;
    if (dyn) pl.p.sched = sched_slot((dev >= 0 && dev < 64) ? dev : 0, mp_stream(stream));
    int sms = (dev >= 0 && dev < 64 && g_num_sms[dev] > 0) ? g_num_sms[dev] : 148;
    int grid = pl.p.total_tiles < sms ? pl.p.total_tiles : sms;
    if (pl.slab)
      k_conv_tc3<<<grid, NUM_THREADS3, pl.smem_bytes, mp_stream(stream)>>>(ma_hi, ma_lo, mb_hi, mb_lo, mo, ma2_hi, ma2_lo, pl.p,
                                                                             pl.x);
    else
      k_conv_tc2<<<grid, NUM_THREADS3, pl.smem_bytes, mp_stream(stream)>>>(ma_hi, ma_lo, mb_hi, mb_lo, mo, ma2_hi, ma2_lo, pl.p);
  }
  MP_LAUNCH_CHECK("mp_conv_tc");
  return 0;
}

// ---- weight gradient on tcgen05 (row f-2): see k_wgrad_tc
namespace {
int wg_pow2_at_least(int v, int cap) {
  int r = 1;
  while (r < v && r < cap) r <<= 1;
  return r;
}
int encode_wg_map(CUtensorMap* m, const void* ptr, int C, int N, int D, int H, int W, const WgParams& p) {
  cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
  cuuint64_t strides[4] = {(cuuint64_t)C * 2, (cuuint64_t)C * 2 * W, (cuuint64_t)C * 2 * W * H, (cuuint64_t)C * 2 * W * H * D};
  cuuint32_t box[5] = {64u, (cuuint32_t)p.bw, (cuuint32_t)p.bh, (cuuint32_t)p.bd, (cuuint32_t)p.bn};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = encode_cached(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(ptr), dims, strides, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : mp_set_error("mp_conv_wgrad_tc: cuTensorMapEncodeTiled failed (%d)", (int)r);
}
}  // namespace

extern "C" int mp_conv_wgrad_tc_supported(int Cin, int Cout, int KD, int KH, int KW) {
  return (Cin > 0 && Cout > 0 && Cin % 8 == 0 && Cout % 8 == 0 && KD % 2 == 1 && KH % 2 == 1 && KW % 2 == 1 &&
          KD * KH * KW <= 4096 && get_encoder() != nullptr) ? 1 : 0;
}

extern "C" int mp_conv_wgrad_tc(const void* x_hi, const void* x_lo, const void* dy_hi, const void* dy_lo, float* dw, int N, int D,
                                int H, int W, int Cin, int Cout, int KD, int KH, int KW, void* stream) {
  MP_REQUIRE(x_hi && x_lo && dy_hi && dy_lo && dw, "mp_conv_wgrad_tc: null pointer");
  MP_REQUIRE(N > 0 && D > 0 && H > 0 && W > 0, "mp_conv_wgrad_tc: bad dims");
  MP_REQUIRE(mp_conv_wgrad_tc_supported(Cin, Cout, KD, KH, KW) == 1,
             "mp_conv_wgrad_tc: needs channel counts in multiples of 8 and odd kernel sizes, got %d -> %d, %dx%dx%d", Cin, Cout, KD,
             KH, KW);
  MP_REQUIRE((((uintptr_t)x_hi | (uintptr_t)x_lo | (uintptr_t)dy_hi | (uintptr_t)dy_lo | (uintptr_t)dw) & 15) == 0,
             "mp_conv_wgrad_tc: operands must be 16-byte aligned");
  WgParams p;
  memset(&p, 0, sizeof(p));
  p.dw = dw; p.Cout = Cout; p.Cin = Cin; p.KD = KD; p.KH = KH; p.KW = KW; p.T = KD * KH * KW;
  // M side in 128-channel tiles, N side in 64 / 128: put the roles so that the zero padding of the tile is smallest
  // (N tile: 128 channels, or the whole N side rounded up to the MMA's multiple of 16 when it is narrower)
  auto n_tile = [](int c) { return c >= 128 ? 128 : (c + 15) / 16 * 16; };
  auto padded = [&](int cm, int cn) { return (int64_t)((cm + 127) / 128 * 128) * ((cn + n_tile(cn) - 1) / n_tile(cn) * n_tile(cn)); };
  p.swap = padded(Cin, Cout) < padded(Cout, Cin) ? 1 : 0;
  const int CM = p.swap ? Cin : Cout, CN = p.swap ? Cout : Cin;
  p.BN = n_tile(CN);
  p.tiles_m = (CM + 127) / 128;
  p.tiles_n = (CN + p.BN - 1) / p.BN;
  // One tap per CTA with 64-position K blocks and a four-wave split-K measured best (tools/wgrad_tc_bench.py: 512 -> 512 3x3 at
  // 289 useful TFLOP/s against 201 with three taps sharing the dY tile over 32-position blocks)
  p.TG = 1;
  int waves = 4;
  {   // tuning overrides (tools/wgrad_tc_bench.py): MPB200_WG_TG = taps per CTA (1..3), MPB200_WG_WAVES = split-K target in waves
    const char* e = getenv("MPB200_WG_TG");
    if (e && atoi(e) >= 1 && atoi(e) <= 3 && atoi(e) <= p.T) p.TG = atoi(e);
    e = getenv("MPB200_WG_WAVES");
    if (e && atoi(e) >= 1 && atoi(e) <= 16) waves = atoi(e);
  }
  p.tap_groups = (p.T + p.TG - 1) / p.TG;
  p.kblk = p.TG == 1 ? 64 : 32;
  if (p.TG == 2) p.kblk = 32;
  // position box of kblk positions: W first, then H, D, and the sample axis takes what is left
  int rem = p.kblk;
  p.bw = wg_pow2_at_least(W, rem < 8 ? rem : 8); rem /= p.bw;
  p.bh = wg_pow2_at_least(H, rem); rem /= p.bh;
  p.bd = wg_pow2_at_least(D, rem); rem /= p.bd;
  p.bn = rem;
  p.nbw = (W + p.bw - 1) / p.bw; p.nbh = (H + p.bh - 1) / p.bh; p.nbd = (D + p.bd - 1) / p.bd;
  const int64_t boxes = (int64_t)p.nbw * p.nbh * p.nbd * ((N + p.bn - 1) / p.bn);
  MP_REQUIRE(boxes < 0x7fffffff, "mp_conv_wgrad_tc: too many position boxes");
  p.total_boxes = (int)boxes;
  p.s_sub = p.swap ? (p.BN + 63) / 64 : 2;
  p.x_sub = p.swap ? 2 : (p.BN + 63) / 64;
  p.sub_bytes = (uint32_t)p.kblk * 128u;
  p.stage_bytes = 2u * p.sub_bytes * (uint32_t)(p.s_sub + p.TG * p.x_sub);
  p.STAGES = (int)((SMEM_LIMIT - 2048u) / p.stage_bytes);
  if (p.STAGES > 6) p.STAGES = 6;
  MP_REQUIRE(p.STAGES >= 2, "mp_conv_wgrad_tc: stage of %u B does not fit twice", p.stage_bytes);
  p.tmem_cols = next_pow2((uint32_t)(p.TG * p.BN) < 32u ? 32u : (uint32_t)(p.TG * p.BN));
  // kind::f16, D = f32, A / B = bf16, both operands MN-major (bits 15 / 16), N = BN, M = 128
  p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const int64_t tiles = (int64_t)p.tiles_m * p.tiles_n * p.tap_groups;
  int dev = 0;
  cudaGetDevice(&dev);
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  // split K so that the grid is about two waves of one CTA per SM; every slice keeps at least 4 boxes
  int64_t slices = (waves * (int64_t)sms + tiles - 1) / tiles;
  if (slices > (p.total_boxes + 3) / 4) slices = (p.total_boxes + 3) / 4;
  if (slices < 1) slices = 1;
  if (slices > 65535) slices = 65535;
  p.boxes_per_slice = (int)((p.total_boxes + slices - 1) / slices);
  slices = (p.total_boxes + p.boxes_per_slice - 1) / p.boxes_per_slice;
  MP_REQUIRE(tiles < 0x7fffffff, "mp_conv_wgrad_tc: too many tiles");
  CUtensorMap ms_hi, ms_lo, mx_hi, mx_lo;
  if (int e = encode_wg_map(&ms_hi, dy_hi, Cout, N, D, H, W, p)) return e;
  if (int e = encode_wg_map(&ms_lo, dy_lo, Cout, N, D, H, W, p)) return e;
  if (int e = encode_wg_map(&mx_hi, x_hi, Cin, N, D, H, W, p)) return e;
  if (int e = encode_wg_map(&mx_lo, x_lo, Cin, N, D, H, W, p)) return e;
  {
    static bool attr_done[64] = {false};
    if (dev < 0 || dev >= 64 || !attr_done[dev]) {
      cudaError_t ae = cudaFuncSetAttribute(k_wgrad_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT);
      MP_REQUIRE(ae == cudaSuccess, "mp_conv_wgrad_tc: cannot opt in to %u B shared memory: %s", SMEM_LIMIT, cudaGetErrorString(ae));
      if (dev >= 0 && dev < 64) attr_done[dev] = true;
    }
  }
  cudaError_t e = cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)Cout * p.T * Cin, mp_stream(stream));
  MP_REQUIRE(e == cudaSuccess, "mp_conv_wgrad_tc: memset: %s", cudaGetErrorString(e));
  const uint32_t smem_bytes = (uint32_t)p.STAGES * p.stage_bytes + 2048u;
  dim3 grid((unsigned)tiles, (unsigned)slices);
  k_wgrad_tc<<<grid, WG_TC_THREADS, smem_bytes, mp_stream(stream)>>>(ms_hi, ms_lo, mx_hi, mx_lo, p);
  MP_LAUNCH_CHECK("mp_conv_wgrad_tc");
  return 0;
}
