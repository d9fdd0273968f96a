/* mpb200.h -- C ABI of libmpb200.so: the B200 (sm_100a) kernels behind the Gbase volumetric forward path.
 *
 * The reference (johndpope/MegaPortrait-hack, model.py) has no native code and no FFI: every operator on its hot
 * path is a PyTorch ATen call.  The entry points below are therefore what a ctypes binding of that path binds
 * (INTEGRATION.md shows the stub); each one names the reference call site (model.py:line) it replaces.
 *
 * Conventions
 *   - plain pointers and sizes only; all pointers are DEVICE pointers unless the name ends in `_host`;
 *   - every call is asynchronous and stream-ordered on `stream` (a cudaStream_t passed as void*); no hidden
 *     synchronisation, no default-stream use, no allocation except the per-process TMA-descriptor encoder lookup;
 *     calls may come from any host thread and any number of streams / CUDA graphs per device (mp_conv_tc hands its
 *     tile-scheduler counters out per stream and per captured launch);
 *   - return value 0 = success; non-zero = error, text available from mp_last_error() (thread-local);
 *     never throws, never exits;
 *   - "NCDHW" = the reference's contiguous fp32 layout; "CL" = channels-last [N, D, H, W, C] (2-D tensors use D=1);
 *   - "split" = a pair of bf16 planes (hi, lo) with x ~= hi + lo (16 mantissa bits): the operand format of the
 *     3-pass bf16 tensor-core convolution;
 *   - "fp16 plane" = one fp16 value per element: the activation format of the two-pass MP_PREC_F16X2 convolutions
 *     that serve the pooled motion-encoder trunks.
 */
#ifndef MPB200_H
#define MPB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPB200_ABI_VERSION 4

/* activation codes shared by several entry points */
enum { MP_ACT_NONE = 0, MP_ACT_RELU = 1, MP_ACT_RELU_TANH = 2, MP_ACT_SIGMOID = 3, MP_ACT_TANH = 4 };

int mp_abi_version(void);
const char* mp_last_error(void);
/* 1 if the current device is sm_100 (tcgen05/TMA kernels usable), 0 otherwise, <0 on error. */
int mp_device_supported(void);

/* ---------------------------------------------------------------- layout ------------------------------------ */
/* NCDHW fp32 -> CL.  Any of out_f32 / (out_hi,out_lo) may be NULL.  S = D*H*W.  (boundary of every module) */
int mp_nchw_to_cl(const float* in, float* out_f32, void* out_hi, void* out_lo, int N, int C, int64_t S, void* stream);
/* CL -> NCDHW fp32; source is in_f32 if non-NULL else the split pair. */
int mp_cl_to_nchw(const float* in_f32, const void* in_hi, const void* in_lo, float* out, int N, int C, int64_t S,
                  void* stream);
/* NCHW fp32 with C <= 16 channels (RGB frames) -> split CL [N,S,16], channel axis zero-padded to 16: the input of the
 * tensor-core stems (nn.Conv2d(3, 64, ...), model.py:211; resnet.py:192; mysixdrepnet.py:1230). */
int mp_nchw_to_cl_pad16(const float* in, void* out_hi, void* out_lo, int N, int C, int64_t S, void* stream);
/* fp32 -> split planes, elementwise (n elements). */
int mp_split(const float* in, void* out_hi, void* out_lo, int64_t n, void* stream);

/* AvgPool2d(2,2) / AvgPool3d(2,2) on CL fp32 (model.py:231, 576-580).  pool_d = 1 for 2-D. */
int mp_avgpool2_cl(const float* in, float* out_f32, void* out_hi, void* out_lo, int N, int D, int H, int W, int C,
                   int pool_d, void* stream);
/* nn.Upsample(scale 2, 'trilinear'|'bilinear', align_corners=True) on CL (model.py:585-589, 733-743).
 * up_d = 1 keeps D (2-D bilinear).  Source: in_f32 if non-NULL else split.  Either output may be NULL. */
int mp_upsample2x_linear_cl(const float* in_f32, const void* in_hi, const void* in_lo, float* out_f32, void* out_hi,
                            void* out_lo, int N, int D, int H, int W, int C, int up_d, void* stream);
/* The same bilinear x2 (2-D, align_corners=True) from split planes to the F16_Q8 plane pair (fp16 plane + FP8 byte plane
 * written with the per-tensor power-of-two q8_scale): input of the fp16 + FP8 convolutions of G2d's up-blocks
 * (model.py:733-743).  C % 64 == 0. */
int mp_upsample2x_bilinear_hq(const void* in_hi, const void* in_lo, void* out_h16, void* out_q8, int N, int H, int W, int C,
                              float q8_scale, void* stream);
/* nn.Upsample(scale_factor=(sd,sh,sw)) nearest on CL fp32 (model.py:427-433). */
int mp_upsample_nearest_cl(const float* in, float* out_f32, void* out_hi, void* out_lo, int N, int D, int H, int W,
                           int C, int sd, int sh, int sw, void* stream);

/* ---------------------------------------------------------------- normalisation ----------------------------- */
/* Accumulate per-(sample, group) sum and sum-of-squares of a CL fp32 tensor into stats[N][G][2] (double, must be
 * zeroed by the caller; mp_conv can accumulate the same statistics in its epilogue).  (F.group_norm, model.py:116) */
int mp_gn_stats(const float* x, double* stats, int N, int64_t S, int C, int G, void* stream);
/* stats -> per-(sample, channel) scale/shift ab[N][C][2] such that GN(x)*g2+b2 == x*a + b.
 * gamma/beta (GroupNorm affine) and gamma2/beta2 (AdaptiveGroupNorm's second affine, model.py:314-316) may be NULL. */
int mp_gn_finalize(const double* stats, const float* gamma, const float* beta, const float* gamma2,
                   const float* beta2, float* ab, int N, int64_t S, int C, int G, float eps, void* stream);
/* out = act(x * a[n,c] + b[n,c] + residual) on CL.  ab may be NULL (identity); residual is res_f32, else the split
 * pair, else none.  Outputs: out_f32 and/or split, either may be NULL.  x may alias out_f32. */
int mp_affine_act_cl(const float* x, const float* ab, const float* res_f32, const void* res_hi, const void* res_lo,
                     float* out_f32, void* out_hi, void* out_lo, int N, int64_t S, int C, int act, void* stream);

/* ---------------------------------------------------------------- convolution ------------------------------- */
typedef struct mp_conv_desc {
  /* input activation, split CL [N, D, H, W, Cin] */
  const void* in_hi;
  const void* in_lo;
  /* packed weights, split, K-major: [Cout_pad][taps*Cin], tap index = (kd*KH + kh)*KW + kw, cin fastest */
  const void* w_hi;
  const void* w_lo;
  const float* bias;      /* [Cout] or NULL */
  /* optional residual added before the activation, CL [N, D, H, W, Cout] */
  const float* res_f32;
  const void* res_hi;
  const void* res_lo;
  /* outputs, CL [N, D, H, W, Cout]; any may be NULL */
  float* out_f32;
  void* out_hi;
  void* out_lo;
  double* stats;          /* [N][gn_groups][2] accumulated over the written values, or NULL */
  int N, D, H, W, Cin, Cout;
  int KD, KH, KW;         /* odd, "same" padding, stride 1 (nn.Conv2d/3d(k, padding=k//2), model.py:95-107,374-375) */
  int Cout_pad;           /* rows in the packed weight matrix (>= Cout, multiple of 16) */
  int gn_groups;
  int act;                /* MP_ACT_* applied after bias + residual */
  /* --- ABI v2 extensions (0 = default) --------------------------------------------------------------------- */
  int stride;             /* 1 (default) or 2 on H and W: output grid is (D, H/stride, W/stride); resnet/RepVGG
                             stride-2 3x3 (pad 1) and 1x1 convolutions (resnet.py:101-118, mysixdrepnet.py:1230-1246) */
  int in_c_off;           /* first input channel read (grouped convolutions run one launch per group) */
  int in_C;               /* channels per position of the input tensor (default Cin) */
  int out_c_off;          /* first output channel written */
  int out_C;              /* channels per position of the output / residual tensors (default Cout) */
  /* --- ABI v3 extension (0 = default) ---------------------------------------------------------------------- */
  int prec;               /* MP_PREC_SPLIT_BF16 (0): 3-pass split-bf16 as described above.
                             MP_PREC_F16X2 (1): two-pass fp16 for the pooled encoder trunks (Emtn, model.py:888-907):
                             in_hi / res_hi / out_hi are SINGLE fp16 planes (in_lo, res_lo, out_lo ignored; x rounded
                             to 11 significant bits), w_hi = fp16(w), w_lo = fp16((w - w_hi) * 2048); the kernel
                             accumulates x*w_hi and x*w_lo in fp32 and adds the second sum scaled by 2^-11.
                             mp_conv_tc only. */
  /* Fused 1x1 shortcut (mp_conv_tc only; 0 / NULL = none): out = act(conv(in) + conv1x1_stride2(in2) + bias + res).
   * The residual blocks whose shortcut is Conv2d(1x1) + BatchNorm folded in (ResBlock2D, model.py:616-640; the
   * down-sampling BasicBlock / Bottleneck, resnet.py:101-118) add it inside the accumulator instead of through HBM:
   * in2 is a second activation [N, D, H/stride*stride2, W/stride*stride2, in2_C] in the same operand format as `in`,
   * its Cin2 channels (window starting at in2_c_off) extend K: packed weight rows are [taps*Cin | Cin2] long. */
  int Cin2, in2_C, in2_c_off, stride2;
  const void* in2_hi;
  const void* in2_lo;
  /* MP_PREC_F16_Q8 (prec = 2): fp16 main product + both cross terms in FP8 (e4m3) -- two pass-units instead of three.
   * Operand format "F16_Q8": hi = fp16 plane; lo = a byte plane with 2 bytes per element, laid out per 64-channel
   * group as [64 x e4m3(x)] [64 x e4m3((x - fp16(x)) * 2048)] (weights: [e4m3(wl * 2048 * sw)] [e4m3(w * sw)], sw a
   * power of two); the kernel accumulates fp16(x)*fp16(w) with kind::f16 MMAs and x8*wl8 + xl8*w8 with kind::f8f6f4
   * MMAs (twice the rate) side by side in TMEM and adds the second sum scaled by corr_scale = 1 / (2048 * sw).
   * Channels must be multiples of 64.  out_fmt / res_fmt select the plane format of out_hi/out_lo and res_hi/res_lo
   * independently of `prec` (0 = the native format of `prec`), so format changes ride on a convolution's epilogue. */
  int out_fmt, res_fmt;
  float corr_scale;
  /* Per-tensor power-of-two scale of an F16_Q8 byte plane (0 = 1): the plane holds [e4m3(x * s)] [e4m3((x - fp16 x) *
   * 2048 * s)], so activations far from 1 in magnitude still use e4m3's 2^-6 .. 448 normal range (|x * s| > 448 would
   * saturate the cross term, |x * s| < 2^-9 would drop it).  out_q8_scale: the scale this launch writes out_lo with;
   * res_q8_scale: the scale res_lo was written with.  The scale of the INPUT plane is folded into corr_scale by the
   * caller: corr_scale = 1 / (2048 * sw * s_in). */
  float out_q8_scale, res_q8_scale;
  /* fp16 weight planes (MP_PREC_F16X2, MP_PREC_F16_Q8) may be packed from w * sw, sw a power of two that moves the
   * largest weight into [1, 2) -- so small weights do not sink into fp16's subnormal range; the kernel multiplies the
   * finished accumulator by acc_scale = 1 / sw before bias / residual / activation (0 = 1). */
  float acc_scale;
} mp_conv_desc;

enum { MP_PREC_SPLIT_BF16 = 0, MP_PREC_F16X2 = 1, MP_PREC_F16_Q8 = 2 };
enum { MP_FMT_NATIVE = 0, MP_FMT_SPLIT_BF16 = 1, MP_FMT_F16 = 2, MP_FMT_F16_Q8 = 3 };

/* ---------------------------------------------------------------- frame decode (row f-4) ------------------- */
/* JPEG frames -> uint8 RGB HWC on the device through nvJPEG (a CUDA-toolkit library, dlopen'ed at first use):
 * `Image.open(path).convert("RGB")` of inference.py:10-13 / the per-frame decode of EmoDataset.py:180-247 for JPEG
 * sources.  jpeg_host[i] / nbytes[i] are HOST pointers to n bitstreams of identical size H x W; out_u8 [n, H, W, 3] is a
 * DEVICE buffer in the layout mp_frames_u8_to_f32 reads.  Stream-ordered.  mp_jpeg_info parses the header only. */
int mp_jpeg_info(const unsigned char* jpeg_host, size_t nbytes, int* width, int* height);
int mp_decode_jpeg_frames(const unsigned char* const* jpeg_host, const size_t* nbytes, int n, unsigned char* out_u8,
                          int H, int W, void* stream);

/* Drops the library's host-side caches (encoded TMA descriptors keyed by pointer + geometry; they hold no device
 * memory and never dereference the pointer, so stale entries are harmless -- this only returns the host memory). */
int mp_release_caches(void);

/* Implicit-GEMM convolution on tcgen05 tensor cores fed by TMA (fp32 accumulate in TMEM; operand format by `prec`).
 * Requires Cin % 16 == 0 and a 128-position output tile that is a box of the (D,H,W) grid. */
int mp_conv_tc(const mp_conv_desc* desc, void* stream);
/* Same contract on CUDA cores (fp32 FMA): odd shapes (Cin=3 stem, Cout=3 heads, the FlowField tower) and the
 * on-device cross-check of mp_conv_tc. */
int mp_conv_simt(const mp_conv_desc* desc, void* stream);
/* 1 if mp_conv_tc accepts the shape. */
int mp_conv_tc_supported(const mp_conv_desc* desc);

/* ---------------------------------------------------------------- pooling (motion encoder trunks) ---------- */
/* nn.MaxPool2d(kernel_size=3, stride=2, padding=1) on a split CL tensor [N,1,H,W,C] -> [N,1,H/2,W/2,C] (resnet.py:196). */
int mp_maxpool3x3s2_cl(const void* in_hi, const void* in_lo, void* out_hi, void* out_lo, int N, int H, int W, int C,
                       void* stream);
/* nn.AdaptiveAvgPool2d(1) on a CL tensor (fp32 if in_f32 != NULL else split): [N,S,C] -> out [N,C] fp32. */
int mp_global_avgpool_cl(const float* in_f32, const void* in_hi, const void* in_lo, float* out, int N, int64_t S, int C,
                         void* stream);

/* ---------------------------------------------------------------- fp16 single-plane trunk operators -------- */
/* RGB stem as a GEMM: NCHW fp32 frames [N,C<=3,H,W] -> fp16 CL patches [N, H/stride, W/stride, 32], channel
 * (kh*3+kw)*C + c holds in[n, c, ho*stride+kh-1, wo*stride+kw-1] (zero outside the frame), channels >= 9*C are zero:
 * the 3x3 (pad 1) stems nn.Conv2d(3, 64, 3, stride) of the motion-encoder trunks (resnet.py:192,
 * mysixdrepnet.py:1230) then run as one 1x1 convolution with K = 32 on the tensor cores. */
int mp_im2col3x3_f16(const float* in, void* out_h, int N, int C, int H, int W, int stride, void* stream);
/* nn.MaxPool2d(3, 2, 1) on an fp16 CL tensor [N,1,H,W,C] -> [N,1,H/2,W/2,C] (resnet.py:196). */
int mp_maxpool3x3s2_cl_f16(const void* in_h, void* out_h, int N, int H, int W, int C, void* stream);
/* The whole stem of a CIFAR-style ResNet-18 trunk in one kernel: nn.Conv2d(3, C, 3, 1, 1) + folded BatchNorm + ReLU +
 * nn.MaxPool2d(3, 2, 1) (resnet.py:192-197, 271-273): NCHW fp32 frames [N,3,H,W] -> fp16 CL [N,1,H/2,W/2,C]; the
 * full-resolution C-channel tensor exists only in shared memory.  w_hi / w_lo: the [C, 32] fp16 planes of an MP_PREC_F16X2
 * weight pack over the patch channels (kh*3+kw)*3 + c (27 used), acc_scale as in mp_conv_desc.  C = 64 or 128 (two trunks
 * reading the same frame); H, W multiples of 16. */
int mp_stem3x3_relu_maxpool_f16(const float* x, const void* w_hi, const void* w_lo, const float* bias, float acc_scale,
                                void* out_h, int N, int H, int W, int C, void* stream);
/* nn.AdaptiveAvgPool2d(1) on an fp16 CL tensor [N,S,C] -> out [N,C] fp32 (fp32 accumulation). */
int mp_global_avgpool_cl_f16(const void* in_h, float* out, int N, int64_t S, int C, void* stream);

/* ---------------------------------------------------------------- warping ----------------------------------- */
/* F.grid_sample(v, grid, 'bilinear', 'border', align_corners=True) for 5-D input (model.py:1062).
 * v [N,C,D,H,W], grid [N,Do,Ho,Wo,3] (x,y,z in [-1,1]), out [N,C,Do,Ho,Wo], all fp32 NCDHW. */
int mp_grid_sample3d(const float* v, const float* grid, float* out, int N, int C, int D, int H, int W, int Do,
                     int Ho, int Wo, void* stream);
/* Same result through a channels-last copy of v in a caller-owned workspace (mp_gather_workspace_bytes): 2-5x
 * faster for jittery / random grids (8 channels per fetched sector instead of 1).  Requires C % 4 == 0. */
size_t mp_gather_workspace_bytes(int N, int C, int D, int H, int W);
int mp_grid_sample3d_ws(const float* v, const float* grid, float* out, void* workspace, size_t workspace_bytes, int N,
                        int C, int D, int H, int W, int Do, int Ho, int Wo, void* stream);
int mp_apply_warping_field_ws(const float* v, const float* warp_field, float* out, void* workspace,
                              size_t workspace_bytes, int N, int C, int D, int H, int W, int Df, int Hf, int Wf,
                              void* stream);
/* Brick-staged variants of the two NCDHW drop-ins (same results bit for bit): per output tile the cells that its
 * voxels sample are bounded, the covering (d,h,w) brick of every channel is fetched ONCE by a TMA box load into shared
 * memory (zero fill outside the volume = ATen's within_bounds test), the 8 taps are read from shared memory with the
 * tile's voxels bucketed by bank so that jittery grids stay conflict-free, and each result tile leaves through one TMA
 * box store.  Tiles whose sampling region does not fit the brick fall back to direct gathers inside the same kernel.
 * No whole-volume transpose.  Requires W % 4 == 0, Wo % 4 == 0, D/H/W <= 1023, 16-byte aligned v / out.
 * workspace (optional, caller-owned, mp_gs_brick_workspace_bytes: one byte per output tile): when given, the kernel
 * records which tiles it served and a second launch gathers the others at full occupancy (exits at once when there are
 * none); when NULL, unfit tiles are gathered inside the brick kernel (slow for grids that are mostly unfit).
 * flags: bit 0 = do not bucket by bank (A/B).  mp_gs_brick_tune: {tz, ty, tx, BD, BH, BW, threads, groups} overrides,
 * 0 = automatic (test / profiling hook; process-wide). */
size_t mp_gs_brick_workspace_bytes(int N, int Do, int Ho, int Wo);
int mp_grid_sample3d_brick(const float* v, const float* grid, float* out, void* workspace, size_t workspace_bytes, int N,
                           int C, int D, int H, int W, int Do, int Ho, int Wo, int flags, void* stream);
int mp_apply_warping_field_brick(const float* v, const float* warp_field, float* out, void* workspace,
                                 size_t workspace_bytes, int N, int C, int D, int H, int W, int Df, int Hf, int Wf,
                                 int flags, void* stream);
int mp_gs_brick_tune(const int* cfg, int n);
/* apply_warping_field(v, warp_field) (model.py:1028-1065) in the reference layout: v [N,C,D,H,W],
 * warp_field [N,3,Df,Hf,Wf] -> out [N,C,D,H,W]; the flow resample, identity grid, the reference's
 * re-normalisation and the trilinear border gather run in one kernel. */
int mp_apply_warping_field(const float* v, const float* warp_field, float* out, int N, int C, int D, int H, int W,
                           int Df, int Hf, int Wf, void* stream);
/* Backward of apply_warping_field (row f-2, first operator of the training step: train.py:188 back-propagates through
 * model.py:1062): grad_v [N,C,D,H,W] and / or grad_warp_field [N,3,Df,Hf,Wf] (either may be NULL; both must be
 * ZERO-INITIALISED by the caller, the kernel accumulates with fp32 atomics) from grad_out [N,C,D,H,W], following ATen's
 * grid_sampler_3d_backward for (bilinear, border, align_corners=True) and the trilinear align_corners=True flow
 * resample.  Accumulation order is not fixed: results are reproducible to fp32 rounding, not bit for bit. */
int mp_apply_warping_field_backward(const float* grad_out, const float* v, const float* warp_field, float* grad_v,
                                    float* grad_warp_field, int N, int C, int D, int H, int W, int Df, int Hf, int Wf,
                                    void* stream);
/* Weight gradient of a stride-1 "same" convolution (row f-2; nn.Conv2d / nn.Conv3d / Conv2d_WS / Conv3D_WS of the path,
 * model.py:61-86, 439-471, back-propagated by train.py:188): x [N,D,H,W,Cin] and dy [N,D,H,W,Cout] channels-last fp32,
 * dw [Cout, KD*KH*KW, Cin] fp32 (the K order of the packed forward weights; overwritten).  One GEMM per filter tap with
 * K = positions on the tensor cores, three bf16 passes (fp32-grade), split-K with fp32 RED: reproducible to fp32 rounding,
 * not bit for bit.  Cin, Cout multiples of 4; odd kernel sizes (padding k / 2). */
int mp_conv_wgrad(const float* x, const float* dy, float* dw, int N, int D, int H, int W, int Cin, int Cout, int KD, int KH,
                  int KW, void* stream);
/* The same weight gradient on tcgen05 (round 2, second session): x and dy as split-bf16 plane pairs (hi, lo; channels-last, the
 * operand format of mp_conv_tc).  K = positions: both operands are read as they lie in HBM ("MN-major" UMMA descriptors over TMA
 * boxes of 64 channels x 32 / 64 positions), tap shift = TMA box origin, zero border and channel padding = TMA out-of-bounds
 * fill, one TMEM accumulator per tap of a 3-tap group, split-K with 16-byte fp32 RED.  Cin, Cout multiples of 8.
 * mp_conv_wgrad_tc_supported() tells whether a shape can take this path (else mp_conv_wgrad). */
int mp_conv_wgrad_tc_supported(int Cin, int Cout, int KD, int KH, int KW);
int mp_conv_wgrad_tc(const void* x_hi, const void* x_lo, const void* dy_hi, const void* dy_lo, float* dw, int N, int D, int H,
                     int W, int Cin, int Cout, int KD, int KH, int KW, void* stream);
/* OIDHW fp32 weights [Cout, Cin, T] -> the split-bf16 K-major operand planes [rows_pad][K] of mp_conv_tc in one launch:
 * dgrad = 0: the forward operand (rows = Cout, K = tap * Cin + ci); dgrad = 1: the operand of the data gradient (rows = Cin,
 * K = tap * Cout + co, taps reversed).  Rows >= the real row count are zero.  Bit-identical to ops.pack_conv on fp32 weights. */
int mp_pack_conv_weights(const float* w, void* out_hi, void* out_lo, int Cout, int Cin, int T, int rows_pad, int dgrad, void* stream);
/* Backward of mp_upsample2x_linear_cl (nn.Upsample(x2, bilinear / trilinear, align_corners=True), model.py:585-589, 733-743,
 * back-propagated by train.py:318): grad_out [N, D*up_d, 2H, 2W, C] -> grad_in [N, D, H, W, C], channels-last fp32; the adjoint
 * evaluated as a gather with the forward's own index rule (deterministic, no atomics). */
int mp_upsample2x_linear_backward_cl(const float* grad_out, float* grad_in, int N, int D, int H, int W, int C, int up_d, void* stream);
/* Patches of an RGB frame for the weight gradient of a stem convolution (Eapp 7x7, model.py:211; the 3x3 / 7x7 stems of the
 * ResNets, resnet.py:192, torchvision): x NCHW fp32 [N, C <= 4, H, W] -> split-bf16 rows [N, H/stride, W/stride, Kpad], element
 * (kh*KW + kw)*C + c.  dW is then one mp_conv_wgrad_tc call with a 1x1 filter over Kpad "input channels". */
int mp_im2col_rgb_split(const float* x, void* out_hi, void* out_lo, int N, int C, int H, int W, int KH, int KW, int stride, int Kpad,
                        void* stream);
/* nn.MaxPool2d(3, stride 2, padding 1) of the ResNet stems for the training path (resnet.py:195, torchvision resnet50; row f-2),
 * channels-last fp32: forward writes the maxima and, per element, the window position (kh*3+kw) of the first maximum in ATen's
 * scan order (idx: one byte per output element); backward gathers grad_out through those positions (deterministic). */
int mp_maxpool3x3s2_forward_idx(const float* in, float* out, void* idx, int N, int H, int W, int C, void* stream);
int mp_maxpool3x3s2_backward(const float* grad_out, const void* idx, float* grad_in, int N, int H, int W, int C, void* stream);
/* dL/dbias = column sums of dy [P, C] -> db [C] (overwritten). */
int mp_bias_grad(const float* dy, float* db, int64_t P, int C, void* stream);
/* nn.GroupNorm backward (model.py:302-316, 439-471) on channels-last fp32 tensors [N, S, C]: dx (fp32), dgamma / dbeta [C]
 * (double, overwritten) from dy, the forward input x, the forward's `stats` [N,G,2] (mp_gn_stats or a convolution epilogue)
 * and gamma (NULL = 1).  workspace: N*(2*G + 2*C) doubles (group sums + the per-channel coefficient table). */
int mp_group_norm_backward(const float* x, const float* dy, const double* stats, const float* gamma, float* dx, double* dgamma,
                           double* dbeta, double* workspace, int N, int64_t S, int C, int G, float eps, void* stream);
/* WarpGenerator tail (model.py:965-973): 64^3 field = affine_grid(theta[N,3,4], align_corners=False) +
 * trilinear(em 16^3 -> 64^3, align_corners=False).  em is CL [N,E,E,E,3]; out is NCDHW [N,3,G,G,G]. */
int mp_warp_field(const float* em_cl, const float* theta, float* out, int N, int E, int G, void* stream);
/* Fused pipeline warp on CL volumes: grid built on the fly from em (CL [N,E,E,E,3]) + theta[N,3,4] exactly as
 * mp_warp_field + mp_apply_warping_field would, then the trilinear gather of v (CL fp32 [Nv,D,H,W,C], Nv = 1 or N:
 * one source volume may serve the whole driver batch).  sum_d = 0: out is CL [N,D,H,W,C];
 * sum_d = 1: torch.sum(dim=2) is fused (model.py:1171) and out is CL [N,1,H,W,C].  Outputs fp32 and/or split. */
int mp_warp_fused_cl(const float* v, const float* em_cl, const float* theta, float* out_f32, void* out_hi,
                     void* out_lo, int N, int Nv, int C, int D, int H, int W, int E, int G, int sum_d, void* stream);

/* G2d output head fused into one pass (model.py:748-751): GroupNorm(32,64) -> ReLU -> Conv2d(64,3,3,pad 1) -> act.
 * x [N,H,W,Cin] CL fp32 (the last decoder block's output), ab [N][Cin][2] from mp_gn_finalize, weight_host [Cout][Cin][3][3]
 * and bias_host [Cout] (may be NULL) are HOST pointers (the weights travel as kernel parameters), out [N,Cout,H,W] NCHW
 * fp32.  Cin = 64, Cout = 3.  fp32 FMA arithmetic. */
int mp_gn_relu_conv3x3_head(const float* x, const float* ab, const float* weight_host, const float* bias_host, float* out,
                            int N, int H, int W, int Cin, int Cout, int act, void* stream);

/* ---------------------------------------------------------------- frame I/O ------------- */
/* inference.py:16-20: transforms.ToTensor() + Normalize([mean],[std]) -- uint8 HWC [N,H,W,3] -> fp32 NCHW [N,3,H,W]. */
int mp_frames_u8_to_f32(const void* in_u8, float* out, int N, int H, int W, float mean, float std, void* stream);
/* inference.py:36-43: ((x + shift) * scale * 255).astype(uint8), CHW -> HWC, optional channel reversal
 * (cv2.cvtColor(.., COLOR_BGR2RGB)) -- fp32 NCHW [N,3,H,W] -> uint8 HWC [N,H,W,3]; values are clamped to [0, 255]. */
int mp_frames_f32_to_u8(const float* in, void* out_u8, int N, int H, int W, float shift, float scale, int reverse_channels,
                        void* stream);

/* ---------------------------------------------------------------- image pyramid ----------------------------- */
/* AntiAliasInterpolation2d (model.py:683-691): zero-pad, depthwise ks x ks filter, nearest subsample by `step`.
 * x [N,C,H,W] fp32 NCHW, kernel [ks*ks] (same for every channel), out [N,C,H/step,W/step]. */
int mp_blur_subsample(const float* x, const float* kernel, float* out, int N, int C, int H, int W, int ks, int step,
                      void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MPB200_H */
