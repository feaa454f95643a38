/* libsg_b200 — C ABI of the B200-native scene-graph -> image hot path.
 *
 * The reference (ashual/scene_generation) is pure PyTorch and has no FFI of its own;
 * each entry point below replaces the ATen/cuDNN dispatch of the cited reference call site and is
 * what a maintainer would bind (ctypes stub: INTEGRATION.md).  Conventions:
 *   - plain pointers + sizes, no torch types; all pointers are DEVICE pointers unless noted;
 *   - the caller owns every buffer (inputs, outputs, workspaces); nothing is retained after return;
 *   - every call only ENQUEUES work on `stream` (no device synchronisation, no default stream);
 *   - return 0 on success, non-zero on error; sg_last_error() gives the message (thread local);
 *   - there is no CPU fallback: without an sm_100a device the calls fail.
 * Tensor formats: "NCHW f32" is the reference's layout; "NHWC bf16" is channels-last bf16 with a
 * physical channel count Cp (multiple of 8, channels >= C are zero) — the operand format of the
 * tensor-core kernels.
 */
#ifndef SG_B200_H
#define SG_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* sg_stream_t; /* cudaStream_t */

/* ---- runtime ------------------------------------------------------------------------------- */
const char* sg_last_error(void);
const char* sg_version(void);
int sg_arch(void);                       /* 100 = built for sm_100a */
unsigned long long sg_launch_count(void);/* kernels launched by this library in this process */
void sg_reset_launch_count(void);
void sg_add_launch_count(unsigned long long n); /* account launches replayed from a captured CUDA graph */

/* ---- layout.py:64-184  masks_to_layout / _boxes_to_grid / _pool_samples ---------------------- */
/* mask_dtype: 0 f32, 1 i64, 2 u8.  img_ranges: int32 (N,2) object range [start,end) per image
 * (objects of an image are contiguous, layout.py:152-155).  out_format: 0 NCHW f32 (N,D,H,W),
 * 1 NHWC bf16 (N,H,W,Cp). */
int sg_masks_to_layout_fwd(const float* vecs, const float* boxes, const void* masks, int mask_dtype,
                           const int* img_ranges, int O, int D, int M, int N, int H, int W,
                           int align_corners, int out_format, int Cp, void* out, sg_stream_t stream);
/* d vecs (O,D) f32 always — columns outside [c_begin, c_end) (c_begin % 8 == 0; pass 0, D for all) are written as
 * zeros: model.py:165-168 only the appearance part of a layout vector carries a gradient; d masks (O,M,M) f32 when
 * dmasks != NULL.  ws: optional f32 scratch of 8*O*D floats (per-band partial sums, added in band order). */
int sg_masks_to_layout_bwd(const float* vecs, const float* boxes, const void* masks, int mask_dtype,
                           const int* img_ranges, int O, int D, int M, int N, int H, int W,
                           int align_corners, int grad_format, int Cp, const void* grad_out, int c_begin, int c_end,
                           float* dvecs, float* dmasks, float* ws, long long ws_floats, sg_stream_t stream);
/* test_mode=True compositing (layout.py:157-169); mass_ws: O floats of workspace. */
int sg_masks_to_layout_test(const float* vecs, const float* boxes, const void* masks, int mask_dtype,
                            const int* img_ranges, int O, int D, int M, int N, int H, int W,
                            int align_corners, int out_format, int Cp, float* mass_ws, void* out,
                            sg_stream_t stream);

/* ---- graph.py:74-116  GraphTripleConv gather / pooled scatter -------------------------------- */
/* edges: int64 (T,2) [subject, object] (graph.py:75-76).  CSR of incidences per object:
 * seg_ptr int32 (O+1), seg_src int32 (2T) with seg_src = 2*t + role (role 1 = object slot), in the
 * reference's accumulation order (all subject uses in triple order, then all object uses). */
int sg_gconv_gather_fwd(const float* obj_vecs, const float* pred_vecs, const long long* edges, int O, int T,
                        int Do, int Dp, int out_dtype, int ld_out, void* out, sg_stream_t stream);
int sg_gconv_pool_fwd(const float* new_t, int ldt, int col_o, const int* seg_ptr, const int* seg_src, int O,
                      int H, int avg, int out_dtype, int ld_out, void* out, sg_stream_t stream);
int sg_gconv_pool_bwd(const float* dpooled, const float* dnew_p, const long long* edges, const int* seg_ptr,
                      int T, int H, int Dout, int avg, int ld_out, float* dnew_t, sg_stream_t stream);
int sg_gconv_gather_bwd(const float* dcur, int ldc, const int* seg_ptr, const int* seg_src, int O, int T,
                        int Do, int Dp, float* dobj, float* dpred, sg_stream_t stream);

/* ---- bilinear.py:26-130,246-275  crop_bbox_batch ---------------------------------------------- */
int sg_crop_bbox_fwd(const float* feats, const float* boxes, const long long* box_to_feats, int N, int C,
                     int H, int W, int B, int HH, int WW, int align_corners, int out_format, int Cp,
                     void* out, sg_stream_t stream);
int sg_crop_bbox_bwd(const float* boxes, const long long* box_to_feats, int N, int C, int H, int W, int B,
                     int HH, int WW, int align_corners, int grad_format, int Cp, const void* grad_out,
                     float* dfeats, sg_stream_t stream);

/* ---- generators.py:62-91, layers.py:234-273, discriminators.py:87-245, graph.py:85,120 --------
 * Dense contractions (nn.Conv2d / ConvTranspose2d / Linear fprop + dgrad) as ONE tcgen05
 * implicit-GEMM kernel:   y[img,h,w,co] = epi( sum_{tap} sum_{ci} x[img, plane(tap), h+dh(tap)+in_h0,
 *                                                    w+dw(tap)+in_w0, ci] * wgt[co, wtap(tap), ci] )
 * x is NHWC bf16 viewed as [N][P][H][W][C] (P = parity planes for stride-2, else 1); out-of-range
 * coordinates read as zero (TMA fill) which implements zero padding.  Up to 4 output "phases"
 * (sub-pixel decomposition of transposed / strided-adjoint convs) each with its own tap list and
 * output offset; output address = img*os_img + (h*oh_mul+oh_off)*os_h + (w*ow_mul+ow_off)*os_w + co*os_c. */
#define SG_MAX_TAPS 64
#define SG_ACT_NONE 0
#define SG_ACT_RELU 1
#define SG_ACT_LEAKY 2
#define SG_ACT_TANH 3
#define SG_ACT_SIGMOID 4
typedef struct {
  int16_t dh, dw, plane, wtap;
} sg_tap_t;
typedef struct {
  int tap_begin, ntaps, oh_off, ow_off;
} sg_phase_t;
typedef struct {
  const void* x;          /* bf16 [x_N][x_P][x_H][x_W][x_C], x_C % 8 == 0 */
  int x_N, x_P, x_H, x_W, x_C;
  const void* w;          /* bf16 [w_Cout][w_taps][w_C], w_C % 8 == 0 */
  int w_Cout, w_taps, w_C;
  void* y;                /* f32 or bf16 */
  int y_dtype;            /* 0 f32, 1 bf16 */
  long long y_os_img, y_os_h, y_os_w, y_os_c; /* element strides; y_os_c = 1 for NHWC */
  int Hout, Wout;         /* per-phase logical output extent */
  int oh_mul, ow_mul;
  int in_h0, in_w0;
  int nphases;
  sg_phase_t phases[4];
  int ntaps;
  sg_tap_t taps[SG_MAX_TAPS];
  const float* bias;      /* (Cout) or NULL */
  int act;
  float slope;
  float* stats;           /* NULL, or f32 (x_N, stats_slots, Cout, 2): partial sum / sum-of-squares of the
                             pre-activation output per (image, slot, channel) for InstanceNorm / BatchNorm.  Every
                             entry is written exactly once, by a fixed-order reduction (no atomics, no zeroing needed):
                             bit-reproducible.  sg_norm_finalize adds the slots in slot order. */
  int stats_slots;        /* slots per (image, channel) this launch writes: sg_conv_stats_slots() */
  int w_img_rows;         /* 0: one weight tensor for all images.  > 0: per-image weights, w is
                             [x_N * w_img_rows][w_taps][w_C] and image i uses rows [i * w_img_rows, + w_Cout)
                             (channel-compacted layouts, see sg_pack_weight_cmap); needs Hout*Wout >= 128 */
  int w_row0;             /* per-image weights only: first row (output channel) within an image's block, so that
                             a restricted dgrad covers rows [w_row0, w_row0 + w_Cout); 0 otherwise */
  int w_mn;               /* 1: "transposed" use of an fprop weight tensor (the dgrad of a convolution reads the same
                             bf16 copy as its fprop): w is [w_rows][w_taps][w_C], the contraction runs over the ROWS
                             (and x's channels) and output channel n is column w_col0 + n:
                               y[.., n] = sum_tap sum_k x[.., k] * w[k, wtap, w_col0 + n],  n < w_Cout */
  int w_rows, w_col0;     /* w_mn only; w_col0 % 8 == 0, w_col0 + w_Cout <= w_C */
} sg_conv_desc_t;
int sg_conv_tc(const sg_conv_desc_t* desc, sg_stream_t stream);
/* number of partial-sum slots per (image, channel) sg_conv_tc writes to desc->stats for this geometry
 * (= phases x M tiles per image); host-only, needs no pointers in desc. */
int sg_conv_stats_slots(const sg_conv_desc_t* desc, int* slots);

/* Weight gradient (cuDNN wgrad in the reference):
 *   dw[co, wtap, ci] += sum_{img,h,w} dy[img, pa, h+dha, w+dwa, co] * x[img, pb, h+dhb, w+dwb, ci]
 * for every entry of the tap table (a = dy side, b = x side).  dw is f32 [Cout][w_taps][dw_C] and is
 * OVERWRITTEN.  When the reduction over pixels is split across CTAs, every split stores its partial sums to a slab
 * of the workspace `ws` and a second kernel adds the slabs in split order, so the result is bit-reproducible;
 * without a workspace the reduction is not split.
 * At least one side must have zero tap offsets and extents equal to (Hred, Wred), so that pixels
 * outside the reduction extent read as zero through the TMA fill. */
typedef struct {
  int16_t dha, dwa, pa, dhb, dwb, pb, wtap, pad;
} sg_wtap_t;
typedef struct {
  const void* dy;         /* bf16 [N][dy_P][dy_H][dy_W][dy_C] */
  int N, dy_P, dy_H, dy_W, dy_C;
  const void* x;          /* bf16 [N][x_P][x_H][x_W][x_C] */
  int x_P, x_H, x_W, x_C;
  int Hred, Wred;         /* reduction extent (h in [0,Hred), w in [0,Wred)) */
  float* dw;              /* f32 [Cout][w_taps][dw_C] */
  int Cout, Cin, w_taps, dw_C;
  int ntaps;
  sg_wtap_t taps[SG_MAX_TAPS];
  int ksplit;             /* 0 = auto */
  int per_image;          /* 1: dw is f32 [N][Cout][w_taps][dw_C], one slab per image (no reduction over
                             images; the per-image weight gradients of a channel-compacted operand) */
  float* ws;              /* f32 scratch for split partial sums: the reduction uses at most ws_floats / (Cout * w_taps *
                             dw_C) splits; stream-ordered reuse is fine, never share between streams; may be NULL */
  long long ws_floats;
} sg_wgrad_desc_t;
int sg_wgrad_tc(const sg_wgrad_desc_t* desc, sg_stream_t stream);

/* Hardware probe (not on any product path): clocks per tcgen05.mma (M = 128, K = 16, N = BN in {16, 64, 128, 256})
 * issued back to back by one thread over fixed shared-memory operands; mode 0 = the four K steps of a k-block into one
 * accumulator, 1 = the same K step, 2 = four accumulators round robin.  out: one float on the device. */
int sg_probe_mma_rate(int BN, int iters, int mode, float* out, sg_stream_t stream);
/* Hardware probe (groundwork for an smem-resident halo tile, not on any product path): a 3x3 convolution of one
 * 16x8-pixel tile, 64 -> 64 channels, whose nine taps read ONE halo tile through tap-shifted UMMA descriptors.
 * x: bf16 [1][1][18][10][64], w: bf16 [64][9][64], y: f32 [128][64]; mode 0 / 1 = descriptor base_offset 0 / derived
 * from the start address.  tests/halo_probe.py compares both against sg_conv_tc. */
int sg_probe_shifted_desc(const void* x, const void* w, float* y, int mode, sg_stream_t stream);

/* Input gradient of a stride-1 "valid" convolution with a tiny output-channel count (the generator's last
 * 7x7 conv 64 -> 3, generators.py:87) as a direct CUDA-core convolution:
 *   dx[n,u,v,ci] = sum_{kh,kw,co} dz[n,u-kh,v-kw,co] * w[co,kh,kw,ci],  u < H+k-1, v < W+k-1
 * dz: bf16 [N][H][W][dzC] (first Cout channels used), w: the f32 master [Cout][k*k][Cin], dx: bf16
 * [N][H+k-1][W+k-1][Cin].  Cout <= 4, k <= 7, Cin in {32, 64}. */
int sg_dgrad_small_cout(const void* dz, int dzC, const float* w, int Cout, int k, int Cin, int N, int H, int W,
                        void* dx, sg_stream_t stream);

/* Tap-unrolled copy of the output gradient of a tiny-Cout stride-1 "valid" convolution over the PADDED pixel grid:
 *   col[n,u,v, co*k*k + kh*k + kw] = dz[n, u-kh, v-kw, co]   (zero outside; bf16 [N][H+k-1][W+k-1][Kp], Kp % 8 == 0)
 * so that both adjoints of the layer (generators.py:87) are ordinary tensor-core GEMMs: dgrad = sg_conv_tc with one
 * tap over col (K = Cout*k*k), wgrad = sg_wgrad_tc with one tap (dy side = col, x side = the padded operand). */
int sg_im2col_dz(const void* dz, int dzC, int Cout, int k, int N, int H, int W, int Kp, void* col, sg_stream_t stream);

/* Weight gradient of the same layer (tiny Cout): dw[co,kh*k+kw,ci] = sum dz[n,h,w,co] * xop[n,h+kh,w+kw,ci]
 * with the pre-padded bf16 operand xop [N][H+k-1][W+k-1][64]; dw f32 [Cout][k*k][64] is overwritten.
 * Cout <= 3, Cin == 64, k in {3, 7}.  ws: f32 scratch of SG_WGRAD_SMALL_BLOCKS * Cout*k*k*64 floats (per-CTA partial
 * sums, added in CTA order). */
#define SG_WGRAD_SMALL_BLOCKS 296
int sg_wgrad_small_cout(const void* dz, int dzC, const void* xop, int Cout, int k, int Cin, int N, int H, int W,
                        float* dw, float* ws, long long ws_floats, sg_stream_t stream);

/* ---- operand preparation ------------------------------------------------------------------ */
/* f32 (rows, cols) with row pitch ld_src -> bf16 (rows, ld_dst); columns >= cols are zero.  With
 * mask_y != NULL the value is multiplied by relu'/leaky' derived from the layer OUTPUT mask_y
 * (same shape/pitch as src): 1 where mask_y > 0 else `slope` (adjoint of the fused Linear+ReLU,
 * layers.py:215-231). */
int sg_cast_pad_bf16(const float* src, long long rows, int cols, long long ld_src, int ld_dst,
                     const float* mask_y, float slope, void* dst, sg_stream_t stream);
/* master weights f32 [Cout][taps][Cin] -> bf16 operand [Cout][taps][Cin_p] (fprop / wgrad-free B
 * operand) and, when wt != NULL, the transposed bf16 [Cin][taps][Cout_p] used by dgrad. */
int sg_pack_weight(const float* w, int Cout, int taps, int Cin, int Cin_p, int Cout_p, void* wk, void* wt,
                   sg_stream_t stream);

/* ---- channel-compacted layouts ----------------------------------------------------------------
 * model.py:165-168 builds layout vectors cat(one_hot(class), appearance): per image only the classes that
 * occur in it give non-zero layout channels.  A compacted layout keeps Cc (= 64) channels per image and a
 * channel map cmap int32 (N, Cc): dense channel of compact channel j of image n, or -1.  The consumer's
 * first convolution (generators.py:69, discriminators.py:211) then runs with per-image weights
 *   wk[n][co][tap][j]  = w[co][tap][cmap[n][j]]          (bf16, fprop / sg_conv_desc_t.w_img_rows = Cout)
 *   wt[n][j][tap][co]  = w[co][tap][cmap[n][j]]          (bf16 [N][Cc][taps][Cout_p], dgrad, w_img_rows = Cc)
 * (zero where cmap is -1 or >= Cin); either output may be NULL. */
int sg_pack_weight_cmap(const float* w, int Cout, int taps, int Cin, const int* cmap, int N, int Cc, int Cout_p,
                        void* wk, void* wt, sg_stream_t stream);
/* adjoint: per-image weight gradients dwc f32 [N][Cout][taps][Cc] (sg_wgrad_desc_t.per_image) summed, in image
 * order, into the dense dw f32 [Cout][taps][Cin], which is overwritten.  A dense channel may occur at most once in
 * an image's row of cmap; Cc <= 127. */
int sg_wgrad_cmap_scatter(const float* dwc, const int* cmap, int N, int Cout, int taps, int Cc, int Cin, float* dw,
                          sg_stream_t stream);

/* ---- layers.py:292-301 InstanceNorm2d / BatchNorm2d, ReLU / LeakyReLU, ReflectionPad2d,
 *      Interpolate(nearest x2), fused into one operand-writer pass (and its adjoint) ----------- */
/* conv-epilogue partial sums (n_img, n_slots, C, 2) (sg_conv_desc_t.stats) -> scale/shift/mean/rstd (n_img*C each);
 * the slots are added in slot order.  mode 0: InstanceNorm2d (affine=False), mode 1: BatchNorm2d train mode
 * (statistics over all n_img*count elements, running stats updated in place with the unbiased variance when
 * running_mean != NULL). */
int sg_norm_finalize(const float* stats, int n_slots, int mode, int n_img, int C, float count, float eps, const float* gamma,
                     const float* beta, float* running_mean, float* running_var, float momentum, float* scale,
                     float* shift, float* save_mean, float* save_rstd, sg_stream_t stream);
typedef struct {
  const void* src;        /* bf16 NHWC [N][H][W][C], C % 8 == 0, C <= 2048 (one thread per 8 channels), 4 N < 65536,
                             one image's padded operand below 2^31 elements (raw conv output) */
  int N, H, W, C;
  const float* scale;     /* (N*C) or NULL (identity) */
  const float* shift;
  int act;                /* SG_ACT_NONE / RELU / LEAKY */
  float slope;
  const void* res;        /* bf16 residual addressed img*res_os_img + h*res_os_h + w*res_os_w + c, or NULL */
  long long res_os_img, res_os_h, res_os_w;
  int up;                 /* 1, or 2 = nearest-neighbour x2 upsampling before padding */
  int pad;                /* halo width */
  int pad_mode;           /* 0 zeros, 1 reflection */
  int planes;             /* 0: out [N][1][Hp][Wp][C]; 1: parity planes [N][4][ceil(Hp/2)][ceil(Wp/2)][C] */
} sg_nap_desc_t;
int sg_norm_act_pad_fwd(const sg_nap_desc_t* d, void* out, sg_stream_t stream);
/* adjoint: grad has the layout of the forward output.  With save_mean != NULL the norm backward
 * dsrc = scale * (g' - mean(g') - xhat * mean(g' xhat)) is applied (bn=1: statistics over all images);
 * sums is f32 scratch of (parts + 1)*N*C*2 floats, parts = sg_norm_act_pad_bwd_parts(N,H,W,C) <= SG_NAP_MAX_PARTS
 * (partial S1 = sum g', S2 = sum g' xhat per reduction CTA and image, then their sums in part order: no atomics, no
 * zeroing) plus, for bn=1, C*2 more floats behind them that return the batch totals (= d beta, d gamma for BatchNorm).
 * dsrc is plain bf16 NHWC or (out_planes=1) parity planes.
 * dres (optional, bf16, addressed with the residual's res_os_* strides) receives the folded
 * gradient of the residual input. */
#define SG_NAP_MAX_PARTS 32
int sg_norm_act_pad_bwd(const sg_nap_desc_t* d, const void* grad, const float* save_mean, const float* save_rstd,
                        int bn, float count, float* sums, int out_planes, void* dsrc, void* dres,
                        sg_stream_t stream);
int sg_norm_act_pad_bwd_parts(int N, int H, int W, int C);   /* host-only; 0 on bad arguments */
/* f32 NCHW grad * act'(y) (tanh / sigmoid heads, generators.py:87, model.py:107) -> bf16 NHWC (Cp). */
int sg_act_bwd_nchw(const float* dy, const float* y, int N, int C, int H, int W, int act, int Cp, void* out,
                    sg_stream_t stream);
/* NCHW (f32: dtype 0, i64: dtype 1) -> channels [c0, c0+C) of a bf16 NHWC (Cp) tensor, and back. */
int sg_nchw_to_nhwc(const void* src, int src_dtype, int N, int C, int H, int W, int Cp, int c0, void* out,
                    sg_stream_t stream);
int sg_nhwc_to_nchw(const void* src, int N, int C, int H, int W, int Cp, int c0, float* out, sg_stream_t stream);
/* discriminators.py:107-109: concat a broadcast one-hot class vector (cls int64 per image, n_cls wide)
 * behind the Cs feature channels -> [rows][Cd]; and the adjoint channel slice. */
int sg_concat_cond(const void* src, long long rows_per_img, int n_img, int Cs, int Cd, const long long* cls,
                   int n_cls, void* out, sg_stream_t stream);
int sg_slice_channels(const void* src, long long rows, int Cd, int Cs, void* out, sg_stream_t stream);
/* discriminators.py:99,184: AvgPool2d(3, stride 2, pad 1, count_include_pad=False), bf16 NHWC. */
int sg_avgpool3x3s2_fwd(const void* x, int N, int H, int W, int C, void* y, sg_stream_t stream);
int sg_avgpool3x3s2_bwd(const void* gy, int N, int H, int W, int C, void* gx, sg_stream_t stream);
/* MaxPool2d(2, 2) of torchvision's VGG19 (losses.py:187-196) on bf16 [N][H][W][C] (floor mode); the adjoint routes the
 * gradient to the first maximum of each window in scan order, like ATen. */
int sg_maxpool2x2_fwd(const void* x, int N, int H, int W, int C, void* y, sg_stream_t stream);
int sg_maxpool2x2_bwd(const void* gy, const void* x, int N, int H, int W, int C, void* gx, sg_stream_t stream);
/* layers.py:82-85 GlobalAvgPool: bf16 [N][HW][C] -> f32 [N][C], and the adjoint. */
int sg_gap_fwd(const void* x, int N, int HW, int C, float* y, sg_stream_t stream);
int sg_gap_bwd(const float* gy, int N, int HW, int C, void* gx, sg_stream_t stream);
/* bias gradient: column sums of bf16 [rows][ld] (first C columns) written to f32 out[C].  Fixed-order reduction:
 * ws = f32 scratch of SG_COLSUM_MAX_BLOCKS * C floats (one partial row per CTA, added in CTA order by a second
 * kernel; stream-ordered reuse is fine, never share between streams). */
#define SG_COLSUM_MAX_BLOCKS 296
int sg_colsum_bf16(const void* x, long long rows, int C, int ld, float* out, float* ws, long long ws_floats,
                   sg_stream_t stream);

/* ---- trainer.py:60,80,106,133 (torch.optim.Adam.step) + operand refresh ----------------------------------
 * Multi-tensor Adam (no weight decay, no amsgrad) over n_tensors parameter tensors, each given by its dense
 * physical storage (p, g, m, v: f32, same layout, numel[i] elements).  step[i] points at the parameter's f32
 * step counter ON THE DEVICE, already incremented for this step (bias corrections are evaluated in the kernel,
 * so the call can be captured in a CUDA graph).  wk[i] != NULL: the bf16 tensor-core operand of that weight,
 * rows of C[i] elements at pitch Cp[i] (see sg_pack_weight), is rewritten from the updated master in the same
 * pass; the pad columns [C, Cp) are left untouched. */
int sg_adam_pack(int n_tensors, void* const* p, void* const* g, void* const* m, void* const* v, void* const* wk,
                 void* const* step, const long long* numel, const int* C, const int* Cp, double lr, double beta1,
                 double beta2, double eps, sg_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SG_B200_H */
