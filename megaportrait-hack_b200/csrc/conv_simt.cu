// CUDA-core (fp32 FMA) implicit-GEMM convolution with the same contract as mp_conv_tc.
//
// Role: (a) shapes the tensor-core kernel does not take (Cin=3 stem, Cout=3 heads, the launch-bound FlowField
// tower whose M is 4..4096 positions), (b) the on-device cross-check of mp_conv_tc in the parity tests.
// Operands are the same split-bf16 planes; they are re-joined to fp32 (x = hi + lo) before the FMA, so the two
// kernels differ only by the dropped lo*lo term and the accumulation order.
//
// Tile: 64 output positions x 64 output channels per CTA of 256 threads (4x4 outputs per thread), K walked in
// chunks of 16 input channels per filter tap.  Bound: fp32 FMA pipe (not a roofline kernel).
#include "common.cuh"

namespace {

constexpr int BM = 64, BN = 64, BK = 16, THREADS = 256;

struct ConvArgs {
  const bf16 *in_hi, *in_lo, *w_hi, *w_lo;
  const float* bias;
  const float* res_f32;
  const bf16 *res_hi, *res_lo;
  float* out_f32;
  bf16 *out_hi, *out_lo;
  double* stats;
  int N, D, H, W, Cin, Cout, KD, KH, KW, gn_groups, act;   // H, W: INPUT extents
  int Ho, Wo, stride, in_c_off, in_C, out_c_off, out_C;
  int64_t M;      // N*D*Ho*Wo
  int64_t S;      // D*Ho*Wo
  int Ktot;       // taps*Cin
};

__global__ void __launch_bounds__(THREADS) k_conv_simt(ConvArgs a) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  __shared__ double s_stats[BN][2];

  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;          // 16x16 threads, each 4 rows x 4 cols
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;

  // loader mapping: A: row = tid/4 (64 rows), 4 channels at (tid%4)*4; B: cout = tid/4, 4 k at (tid%4)*4
  const int lrow = tid / 4, lq = (tid % 4) * 4;
  const int64_t am = m0 + lrow;
  const bool arow_ok = am < a.M;
  int an = 0, ad = 0, ah = 0, aw = 0;
  if (arow_ok) {
    int64_t r = am;
    aw = (int)(r % a.Wo); r /= a.Wo;
    ah = (int)(r % a.Ho); r /= a.Ho;
    ad = (int)(r % a.D);
    an = (int)(r / a.D);
  }
  const int bco = n0 + lrow;
  const bool brow_ok = bco < a.Cout;
  const bool vec_ok = ((a.Cin | a.in_C | a.in_c_off) % 4) == 0;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int pd = a.KD / 2, ph = a.KH / 2, pw = a.KW / 2;
  for (int kd = 0; kd < a.KD; ++kd)
    for (int kh = 0; kh < a.KH; ++kh)
      for (int kw = 0; kw < a.KW; ++kw) {
        const int tap = (kd * a.KH + kh) * a.KW + kw;
        // taps that are out of bounds for every position (extent-1 axes of the FlowField tower) carry only zeros
        if ((a.D == 1 && kd != pd) || (a.H == 1 && kh != ph) || (a.W == 1 && kw != pw)) continue;
        const int id = ad + kd - pd, ih = ah * a.stride + kh - ph, iw = aw * a.stride + kw - pw;
        const bool in_ok = arow_ok && id >= 0 && id < a.D && ih >= 0 && ih < a.H && iw >= 0 && iw < a.W;
        const int64_t in_base = ((((int64_t)an * a.D + id) * a.H + ih) * a.W + iw) * a.in_C + a.in_c_off;
        for (int c0 = 0; c0 < a.Cin; c0 += BK) {
          // ---- A chunk
          float av[4] = {0.f, 0.f, 0.f, 0.f};
          const int c = c0 + lq;
          if (in_ok) {
            if (vec_ok && c + 3 < a.Cin) {
              float4 q = mp_load_split4(a.in_hi, a.in_lo, in_base + c);
              av[0] = q.x; av[1] = q.y; av[2] = q.z; av[3] = q.w;
            } else {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                if (c + k < a.Cin) av[k] = mp_join(a.in_hi[in_base + c + k], a.in_lo[in_base + c + k]);
            }
          }
          // ---- B chunk
          float bv[4] = {0.f, 0.f, 0.f, 0.f};
          if (brow_ok) {
            const int64_t wb = (int64_t)bco * a.Ktot + (int64_t)tap * a.Cin;
            if (vec_ok && c + 3 < a.Cin) {
              float4 q = mp_load_split4(a.w_hi, a.w_lo, wb + c);
              bv[0] = q.x; bv[1] = q.y; bv[2] = q.z; bv[3] = q.w;
            } else {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                if (c + k < a.Cin) bv[k] = mp_join(a.w_hi[wb + c + k], a.w_lo[wb + c + k]);
            }
          }
          __syncthreads();
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            As[lq + k][lrow] = av[k];
            Bs[lq + k][lrow] = bv[k];
          }
          __syncthreads();
#pragma unroll
          for (int kk = 0; kk < BK; ++kk) {
            float4 ar = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            float4 br = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            float af[4] = {ar.x, ar.y, ar.z, ar.w}, bf[4] = {br.x, br.y, br.z, br.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
              for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(af[i], bf[j], acc[i][j]);
          }
        }
      }

  // ---- epilogue: bias + residual + activation, stores, GroupNorm statistics
  const bool tile_one_sample = a.stats && (a.S % BM == 0);
  if (a.stats) {
    for (int i = tid; i < BN * 2; i += THREADS) (&s_stats[0][0])[i] = 0.0;
    __syncthreads();
  }
  const int cpg = a.gn_groups > 0 ? a.Cout / a.gn_groups : 1;
  float csum[4] = {0.f, 0.f, 0.f, 0.f}, csq[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + ty * 4 + i;
    if (m >= a.M) continue;
    const int n = (int)(m / a.S);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = n0 + tx * 4 + j;
      if (co >= a.Cout) continue;
      const int64_t o = m * a.out_C + a.out_c_off + co;
      float v = acc[i][j];
      if (a.bias) v += a.bias[co];
      if (a.res_f32) v += a.res_f32[o];
      else if (a.res_hi) v += mp_join(a.res_hi[o], a.res_lo[o]);
      v = mp_apply_act(v, a.act);
      if (a.out_f32) a.out_f32[o] = v;
      if (a.out_hi) mp_split2(v, a.out_hi[o], a.out_lo[o]);
      if (a.stats) {
        if (tile_one_sample) {
          csum[j] += v;
          csq[j] += v * v;
        } else {
          double* st = a.stats + ((int64_t)n * a.gn_groups + co / cpg) * 2;
          atomicAdd(st, (double)v);
          atomicAdd(st + 1, (double)v * v);
        }
      }
    }
  }
  if (tile_one_sample) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      atomicAdd(&s_stats[tx * 4 + j][0], (double)csum[j]);
      atomicAdd(&s_stats[tx * 4 + j][1], (double)csq[j]);
    }
    __syncthreads();
    const int n = (int)(m0 / a.S);
    for (int i = tid; i < BN; i += THREADS) {
      const int co = n0 + i;
      if (co < a.Cout) {
        double* st = a.stats + ((int64_t)n * a.gn_groups + co / cpg) * 2;
        atomicAdd(st, s_stats[i][0]);
        atomicAdd(st + 1, s_stats[i][1]);
      }
    }
  }
}

}  // namespace

int mp_conv_validate(const mp_conv_desc* d, const char* who) {
  MP_REQUIRE(d, "%s: null descriptor", who);
  MP_REQUIRE(d->prec == MP_PREC_SPLIT_BF16 || d->prec == MP_PREC_F16X2 || d->prec == MP_PREC_F16_Q8, "%s: unknown prec", who);
  MP_REQUIRE(d->prec != MP_PREC_F16_Q8 || d->corr_scale > 0.f, "%s: MP_PREC_F16_Q8 needs corr_scale", who);
  const bool one_plane = d->prec == MP_PREC_F16X2;
  MP_REQUIRE(d->in_hi && (d->in_lo || one_plane) && d->w_hi && d->w_lo, "%s: null operand", who);
  MP_REQUIRE(d->out_f32 || (d->out_hi && (d->out_lo || (one_plane && d->out_fmt != MP_FMT_SPLIT_BF16 && d->out_fmt != MP_FMT_F16_Q8) || d->out_fmt == MP_FMT_F16)),
             "%s: no output", who);
  MP_REQUIRE(!one_plane || !d->stats, "%s: fused GroupNorm statistics are not available in fp16 two-pass mode", who);
  MP_REQUIRE(d->N > 0 && d->D > 0 && d->H > 0 && d->W > 0 && d->Cin > 0 && d->Cout > 0, "%s: bad dims", who);
  MP_REQUIRE(d->KD % 2 == 1 && d->KH % 2 == 1 && d->KW % 2 == 1, "%s: kernel extents must be odd", who);
  MP_REQUIRE(d->Cout_pad >= d->Cout, "%s: Cout_pad < Cout", who);
  MP_REQUIRE(!d->stats || (d->gn_groups > 0 && d->Cout % d->gn_groups == 0), "%s: bad gn_groups", who);
  MP_REQUIRE(d->stride >= 0 && d->stride <= 2, "%s: stride must be 1 or 2", who);
  MP_REQUIRE(d->in_c_off >= 0 && d->out_c_off >= 0, "%s: negative channel offset", who);
  MP_REQUIRE(d->in_C == 0 || d->in_C >= d->in_c_off + d->Cin, "%s: input channel window exceeds in_C", who);
  MP_REQUIRE(d->out_C == 0 || d->out_C >= d->out_c_off + d->Cout, "%s: output channel window exceeds out_C", who);
  MP_REQUIRE(d->in_C != 0 || d->in_c_off == 0, "%s: in_c_off needs in_C", who);
  MP_REQUIRE(d->out_C != 0 || d->out_c_off == 0, "%s: out_c_off needs out_C", who);
  return 0;
}

extern "C" int mp_conv_simt(const mp_conv_desc* d, void* stream) {
  if (int e = mp_conv_validate(d, "mp_conv_simt")) return e;
  MP_REQUIRE(d->prec == MP_PREC_SPLIT_BF16, "mp_conv_simt: only the split-bf16 operand format is implemented");
  MP_REQUIRE(d->Cin2 == 0, "mp_conv_simt: the fused 1x1 shortcut is implemented by mp_conv_tc only");
  MP_REQUIRE(d->out_fmt <= MP_FMT_SPLIT_BF16 && d->res_fmt <= MP_FMT_SPLIT_BF16, "mp_conv_simt: split-bf16 planes only");
  ConvArgs a;
  a.in_hi = (const bf16*)d->in_hi; a.in_lo = (const bf16*)d->in_lo;
  a.w_hi = (const bf16*)d->w_hi; a.w_lo = (const bf16*)d->w_lo;
  a.bias = d->bias; a.res_f32 = d->res_f32; a.res_hi = (const bf16*)d->res_hi; a.res_lo = (const bf16*)d->res_lo;
  a.out_f32 = d->out_f32; a.out_hi = (bf16*)d->out_hi; a.out_lo = (bf16*)d->out_lo; a.stats = d->stats;
  a.N = d->N; a.D = d->D; a.H = d->H; a.W = d->W; a.Cin = d->Cin; a.Cout = d->Cout;
  a.KD = d->KD; a.KH = d->KH; a.KW = d->KW; a.gn_groups = d->gn_groups; a.act = d->act;
  a.stride = d->stride > 0 ? d->stride : 1;
  MP_REQUIRE(d->H % a.stride == 0 && d->W % a.stride == 0, "mp_conv_simt: H, W must be multiples of the stride");
  a.Ho = d->H / a.stride; a.Wo = d->W / a.stride;
  a.in_c_off = d->in_c_off; a.in_C = d->in_C > 0 ? d->in_C : d->Cin;
  a.out_c_off = d->out_c_off; a.out_C = d->out_C > 0 ? d->out_C : d->Cout;
  a.S = (int64_t)d->D * a.Ho * a.Wo;
  a.M = a.S * d->N;
  a.Ktot = d->KD * d->KH * d->KW * d->Cin;
  int64_t gx = (a.M + BM - 1) / BM;
  MP_REQUIRE(gx <= 0x7fffffff, "mp_conv_simt: M too large");
  dim3 grid((unsigned)gx, (d->Cout + BN - 1) / BN);
  k_conv_simt<<<grid, THREADS, 0, mp_stream(stream)>>>(a);
  MP_LAUNCH_CHECK("mp_conv_simt");
  return 0;
}
