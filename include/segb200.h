/* segb200 -- C ABI of the B200-native compute engine for SegmenTron's dense hot path.
 *
 * This is the drop-in boundary.  The reference has no FFI of its own for
 * this path other than the pybind module `segmentron._C` (segmentron/modules/csrc/vision.cpp:6-11)
 * and, for everything else, `torch.nn.functional` calls made from its nn.Module classes.  Each entry
 * point below names the reference call site(s) it replaces.  INTEGRATION.md shows the Python/ctypes
 * binding a maintainer of the reference would add.
 *
 * Conventions (segmentron/modules/csrc/criss_cross_attention/ca.h:25-72 is the model):
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless stated otherwise;
 *   - activations are NHWC ("channels_last"), element type `dtype` (SEGB200_BF16 / SEGB200_F16),
 *     channel pitch `*_ld` in elements so a tensor may be a channel slice of a wider buffer
 *     (this is how torch.cat on the hot path is eliminated); pitches and channel counts are
 *     multiples of 8 elements (16 bytes) unless stated otherwise;
 *   - every function enqueues work on `stream` (a cudaStream_t) and returns immediately:
 *     0 = ok, negative = argument error, positive = cudaError_t.  No function synchronises,
 *     allocates device memory or throws.  `segb200_last_error()` returns a thread-local message.
 *   - re-entrant; no global mutable state besides the lazily resolved driver entry point.
 */
#ifndef SEGB200_H_
#define SEGB200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SEGB200_VERSION 100

enum { SEGB200_BF16 = 0, SEGB200_F16 = 1, SEGB200_F32 = 2 };
enum { SEGB200_ACT_NONE = 0, SEGB200_ACT_RELU = 1, SEGB200_ACT_RELU6 = 2 };

int segb200_version(void);
const char* segb200_last_error(void);
/* Tuning knobs (process-global, not thread-safe; defaults are the measured best for one stream):
 *   "gemm_ring_kb"  : shared-memory ring of segb200_conv_gemm in KB (0 = 192 = whole SM).  A smaller ring leaves room for
 *                     a kernel of another stream to co-reside on the SM (dual-stream half-batch overlap).
 *   "dw_ring_slots" : cap on the row-ring depth of segb200_dwconv3x3 (0 = 12).
 *   "gemm_epi2_maxk": largest K (taps x padded cin) for which segb200_conv_gemm uses two epilogue warp-groups (default 512).
 *   "gemm_2cta"     : 0 (default) single-CTA tiles; 1 = CTA-pair kernel (tcgen05 cta_group::2, 256 x BN tiles, each CTA stages
 *                     half of the weight rows) for every eligible shape; 2 = only for K >= 1024.  Experimental.
 *   "gemm_bn128"    : 1 lets segb200_conv_gemm pick 128-wide N tiles when that saves >= 5 % of the persistent grid's rounds
 *                     (wave quantisation).  Default 0: measured slower (operand traffic per FLOP rises by a third). */
int segb200_set_option(const char* name, int value);

/* Diagnostics (only in a library built with -DSEGB200_DBG; otherwise returns -20): point subsequent
 * segb200_conv_gemm launches at 16 device uint64 counters that accumulate, over all CTAs, the clock cycles each role
 * spent waiting: [0] producer: ring slot free, [1] MMA: accumulator free, [2] MMA: operands landed, [3] epilogue:
 * accumulator ready.  NULL disables.  Not thread-safe; profiling only (tools/gemm_waits.py). */
int segb200_debug_set_counters(void* dev_ptr_16_u64);

/* ------------------------------------------------------------------------------------------
 * Dense convolution as an implicit GEMM on tcgen05 tensor cores, with the following
 * BatchNorm (folded to per-channel scale/shift), residual add and activation fused in the epilogue:
 *     y = act( conv(x, w) * scale[c] + shift[c] + residual )
 * Replaces: F.conv2d + F.batch_norm + F.relu sequences of `_ConvBNReLU` / `_ConvBN`
 * (modules/basic.py:65-77, :95-105), the pointwise half of `SeparableConv2d` (basic.py:42-43),
 * the 1x1 shortcut + add of `XceptionBlock` (backbones/xception.py:37-42), ResNet bottleneck
 * convs (backbones/resnet.py:50-58,78-79), ASPP / classifier 1x1 convs (modules/module.py:45-59,
 * models/deeplabv3_plus.py:62-64).
 *
 * wgt: packed [cout][kh*kw][cin_pad] (K-major), element type `dtype`, cin_pad = cin rounded up to a
 *      multiple of the K block (64, or 32/16 when cin < 64); build it with segb200_conv_kblock().
 * Output spatial size (ho, wo) and the top/left padding are explicit so that asymmetric cases
 * (space-to-depth stems) are expressible: tap (ky,kx) reads x[ho*stride + ky*dilation - pad_t, ...].
 */
typedef struct segb200_conv_args {
  const void* x;        /* [n][h][w][x_ld] */
  const void* wgt;      /* packed weights */
  const float* scale;   /* [cout] or NULL (1.0) */
  const float* shift;   /* [cout] or NULL (0.0) */
  const void* residual; /* [n][ho][wo][res_ld] or NULL */
  void* y;              /* [n][ho][wo][y_ld] */
  int32_t n, h, w, cin, x_ld;            /* cin, x_ld, y_ld, res_ld: multiples of 8 elements; cout: any >= 1 */
  int32_t ho, wo, cout, y_ld, res_ld;
  int32_t kh, kw, stride, dilation, pad_t, pad_l;
  int32_t act;          /* SEGB200_ACT_* */
  int32_t dtype;        /* SEGB200_BF16 | SEGB200_F16 */
  int32_t max_ctas;     /* 0 = number of SMs */
  int32_t y_f32;        /* 1: y is fp32 [n][ho][wo][y_ld floats] (y_ld % 4 == 0), residual must be NULL */
} segb200_conv_args;

int segb200_conv_kblock(int cin);                 /* K block (elements) the kernel will use for `cin` */
int segb200_conv_gemm(const segb200_conv_args* a, void* stream);

/* ------------------------------------------------------------------------------------------
 * Depthwise 3x3 convolution (groups = C, padding = dilation) with optional leading ReLU and the
 * following BatchNorm (+ ReLU/ReLU6) fused:  y = act( dw3x3(pre_relu ? relu(x) : x) + shift[c] )
 * The BN scale is folded into the fp32 weights by the caller.
 * Replaces: the depthwise half of `SeparableConv2d` (modules/basic.py:38-41,45-59) and the
 * depthwise `_ConvBNReLU` of `InvertedResidual` (basic.py:152-154).
 * wgt: fp32 [9][c] (tap-major: tap = ky*3+kx).
 */
typedef struct segb200_dwconv_args {
  const void* x;        /* [n][h][w][x_ld] */
  const float* wgt;     /* [9][c] */
  const float* shift;   /* [c] or NULL */
  void* y;              /* [n][ho][wo][y_ld] */
  int32_t n, h, w, c, x_ld, y_ld;
  int32_t ho, wo, stride, dilation;
  int32_t pre_relu, act;
  int32_t dtype;
} segb200_dwconv_args;

int segb200_dwconv3x3(const segb200_dwconv_args* a, void* stream);

/* ------------------------------------------------------------------------------------------
 * Input packing for stride-2 stems: NCHW image (fp32 / bf16 / fp16) -> space-to-depth NHWC
 *   out[n][i][j][(dy*2+dx)*c + ch] = x[n][ch][2i+dy][2j+dx]   (zero beyond the image; channels
 *   4c..ld-1 zero), so that a kxk stride-2 conv becomes a ceil((k+1)/2)^2 stride-1 conv on the
 *   tensor cores.  Replaces the first conv's input read (backbones/xception.py:131,
 *   mobilenet.py:79, resnet.py:116).   hs = ceil(h/2), ws = ceil(w/2).
 */
int segb200_pack_s2d(const void* x_nchw, int x_dtype, void* out, int out_dtype, int n, int c, int h, int w,
                     int out_ld, void* stream);

/* Global average pool over H*W: x [n][h][w][x_ld] -> out [n][c] (dtype), fp32 accumulation in a fixed
 * order (bit-reproducible, no atomics).  Replaces nn.AdaptiveAvgPool2d((1,1)) (module.py:52). */
int segb200_global_avgpool(const void* x, void* out, int n, int h, int w, int c, int x_ld, int dtype, void* stream);

/* Adaptive average pool to s x s bins (torch bin rule floor/ceil): out [n][s][s][out_ld].
 * Replaces nn.AdaptiveAvgPool2d(s) of PyramidPooling (module.py:89). */
int segb200_adaptive_avgpool(const void* x, void* out, int n, int h, int w, int c, int x_ld, int s, int out_ld,
                             int dtype, void* stream);

/* MaxPool2d(3, 2, 1) (-inf padding): x [n][h][w][x_ld] -> y [n][(h-1)/2+1][(w-1)/2+1][y_ld].
 * Replaces nn.MaxPool2d(3, 2, 1) of the ResNet stem (backbones/resnet.py:119). */
int segb200_maxpool3x3s2(const void* x, void* y, int n, int h, int w, int c, int x_ld, int y_ld, int dtype, void* stream);

/* HRNet fuse step: y = act(a + nearest_upsample(z, 2^k)), a/y [n][h][w][*], z [n][h>>k][w>>k][*]; k = 0 is a plain add.
 * Replaces nn.Upsample(scale_factor=2^k, 'nearest') + the running sum + ReLU of HighResolutionModule.forward
 * (backbones/hrnet.py:178-186, :215-232). */
int segb200_upsample_add(const void* a, const void* z, void* y, int n, int h, int w, int c, int a_ld, int z_ld, int y_ld,
                         int k, int act, int dtype, void* stream);

/* Bilinear resize NHWC -> NHWC channel slice.  align_corners as F.interpolate.  From a 1x1 source
 * this is the ASPP image-pooling broadcast (module.py:64).  Replaces F.interpolate at
 * module.py:64,96, deeplabv3_plus.py:71, hrnet_seg.py:57-59. */
int segb200_bilinear_nhwc(const void* x, void* y, int n, int hi, int wi, int c, int x_ld, int ho, int wo,
                          int y_ld, int align_corners, int dtype, void* stream);

/* Final logits upsample: NHWC [n][hi][wi][x_ld] (first c channels) -> NCHW [n][c][ho][wo] contiguous,
 * `out_dtype` may be the input dtype or SEGB200_F32.  Replaces F.interpolate at deeplabv3_plus.py:39,
 * danet.py:32-34.  If `argmax_out` is non-NULL also writes the per-pixel argmax class (uint8,
 * first-max tie-break like torch.argmax) -- the fused metric path of SURVEY.md 8(f1). */
int segb200_bilinear_nchw_out(const void* x, void* y, uint8_t* argmax_out, int n, int hi, int wi, int c,
                              int x_ld, int ho, int wo, int align_corners, int dtype, int out_dtype,
                              void* stream);

/* ------------------------------------------------------------------------------------------
 * Position attention (PAM_Module.forward, modules/module.py:112-131: two torch.bmm + nn.Softmax over an N x N matrix)
 * as a tiled softmax(Q K^T) V kernel on tcgen05 -- the N x N attention is never materialised (two exact passes:
 * row max / row sum, then P V).   y[b][i][:] = gamma * (sum_j softmax_j(q_i . k_j) v_j + bias_v) + x[b][i][:]
 *   q, k : [batch][n_tok][q_ld|k_ld], depth 64 (= in_dim/8 for DANet's 512 channels), outputs of the 1x1 convs incl. bias
 *   vt   : V transposed, [batch][dv][vt_ld] (vt_ld >= n_tok), WITHOUT the conv bias (pass it as bias_v; rows of the
 *          attention sum to 1);  dv % 256 == 0
 *   stat_m, stat_l : workspace, batch*n_tok floats each.   gamma: device pointer to one float. */
int segb200_pam_attention(const void* q, const void* k, const void* vt, const float* bias_v, const float* gamma,
                          const void* x, void* y, float* stat_m, float* stat_l, int batch, int n_tok, int dv, int q_ld,
                          int k_ld, int vt_ld, int x_ld, int y_ld, int dtype, void* stream);

/* The same kernel for the non-local / self-attention blocks of the other heads -- OCNet BaseAttentionBlock.forward
 * (models/ocnet.py:95-113: two torch.bmm around F.softmax, sim_map scaled by key_channels^-0.5):
 *   dk   : query/key depth, 64 or 256 (OCNet: key_channels = 256); q and k may be the same tensor (OCNet's f_query IS f_key) or q a
 *          pre-scaled copy -- the softmax scale is folded into q by the caller (256^-0.5 = 2^-4 is exact in 16-bit floats)
 *   gamma: device pointer to one float, or NULL (= 1);   x: residual, or NULL (none):  y = gamma * (softmax(q k^T) v + bias_v) + x */
int segb200_nonlocal_attention(const void* q, const void* k, const void* vt, const float* bias_v, const float* gamma,
                               const void* x, void* y, float* stat_m, float* stat_l, int batch, int n_tok, int dk, int dv,
                               int q_ld, int k_ld, int vt_ld, int x_ld, int y_ld, int dtype, void* stream);

/* ------------------------------------------------------------------------------------------
 * Channel attention (CAM_Module, modules/module.py:142-162).  The two matrix products run on segb200_conv_gemm
 * (E = X^T X with y_f32 = 1 on a [C][N] transposed copy; y = gamma*(A X) + x with scale = gamma, residual = x);
 * this entry point is the softmax in between:  att[r][j] = softmax_j(rowmax_r(E) - E[r][j])  (fp32 in, dtype out). */
int segb200_cam_softmax(const float* energy, void* att, int rows, int c, int e_ld, int att_ld, int dtype, void* stream);

/* Criss-cross attention (CrissCrossAttention, modules/cc_attention.py:62-72; replaces _C.ca_forward + F.softmax and
 * _C.ca_map_forward + gamma*out + x of csrc/criss_cross_attention/ca_cuda.cu:8-36,94-120).
 *   att [n][h][w][att_ld] fp32 (att_ld >= h+w-1): channels [0,w) = same-row keys (self included), [w,h+w-1) = same-column
 *   keys with the pixel's own row skipped (j = i<y ? i : i+1), softmax over those h+w-1 entries.
 *   y = gamma * sum_z att[p][z] v[key(p,z)] + x   (gamma: device pointer to one float, the module's nn.Parameter). */
int segb200_cca_weight_softmax(const void* q, const void* k, float* att, int n, int h, int w, int c, int q_ld, int k_ld,
                               int att_ld, int dtype, void* stream);
int segb200_cca_map(const float* att, const void* v, const void* x, void* y, const float* gamma, int n, int h, int w, int c,
                    int att_ld, int v_ld, int x_ld, int y_ld, int dtype, void* stream);

/* Layout converters at the module boundary: logical-NCHW contiguous <-> NHWC (channel pitch ld). */
int segb200_nchw_to_nhwc(const void* x, int x_dtype, void* y, int y_dtype, int n, int c, int h, int w, int y_ld,
                         void* stream);
int segb200_nhwc_to_nchw(const void* x, int x_dtype, void* y, int y_dtype, int n, int c, int h, int w, int x_ld,
                         void* stream);
/* Same transpose with a padded pixel pitch: y[n][c][pitch] (pitch >= h*w; the tail is NOT written).  Produces the
 * K-major [C][N] operand of CAM's Gram matrix. */
int segb200_nhwc_to_cn(const void* x, void* y, int n, int c, int hw, int x_ld, int pitch, int dtype, void* stream);

/* ==========================================================================================
 * TRAINING PATH: the kernels behind `loss.backward()` of tools/train.py:135-147 for the
 * conv / BatchNorm / pooling / interpolate / cross-entropy graph of DeepLabv3+.  Activations and activation gradients are
 * NHWC 16-bit; parameter gradients and BatchNorm statistics are fp32.
 * ========================================================================================== */

/* Convolution weight gradient on tcgen05 tensor cores:  dw[cout][kh*kw][cin] += sum_pixels dy[p][cout] * x[p + tap][cin]
 * (fp32, ACCUMULATED with red.global.add -- zero dw first for a plain gradient).  Same geometry arguments as
 * segb200_conv_gemm (x is the conv INPUT, dy the gradient of its OUTPUT [n][ho][wo][dy_ld]).  Replaces the
 * cudnnConvolutionBackwardFilter call autograd makes for every nn.Conv2d (groups = 1) of the model.
 * The [cout][tap][cin] layout is the `channels_last` physical layout of a torch OIHW weight gradient. */
typedef struct segb200_wgrad_args {
  const void* x;        /* [n][h][w][x_ld] */
  const void* dy;       /* [n][ho][wo][dy_ld] */
  float* dw;            /* [cout][kh*kw][cin] fp32, accumulated */
  int32_t n, h, w, cin, x_ld;            /* cin, x_ld, dy_ld multiples of 8; cout any >= 1 */
  int32_t ho, wo, cout, dy_ld;
  int32_t kh, kw, stride, dilation, pad_t, pad_l;
  int32_t dtype;        /* SEGB200_BF16 | SEGB200_F16 */
  int32_t max_ctas;     /* 0 = number of SMs */
  int32_t splits;       /* pixel splits per (cout tile, cin tile, tap); 0 = automatic (whole waves) */
} segb200_wgrad_args;
int segb200_conv_wgrad(const segb200_wgrad_args* a, void* stream);
int segb200_wgrad_debug_swap(int v);   /* diagnostics only: exchange the LBO/SBO descriptor fields */

/* The data gradient of a conv needs no entry point of its own: it is segb200_conv_gemm on dy with the weights packed
 * transposed and tap-flipped ([cin][kh*kw reversed][cout_pad]); stride-2 convs go through segb200_stride2_place. */

/* ---- two-level, fixed-order column reductions (bit-reproducible; no float atomics) ----
 * A reduction over `rows` pixels of `c` channels writes partial[(slab*K + k)*c + ch]; segb200_reduce_slabs() returns the
 * slab count the kernels will use (size the partial buffer with it). */
int segb200_reduce_slabs(long long rows, int c, int max_slabs);
int segb200_reduce_partials(const float* partial, int slabs, int k, int c, float* out, long long stride_k,
                            long long stride_c, int accumulate, float scale, void* stream);

/* Train-mode BatchNorm2d (F.batch_norm(training=True), modules/batch_norm.py / torch):
 *   bn_stats    : partial[(slab*2 + {0,1})*c + ch] = per-slab sum / sum of squares of x [rows][x_ld]
 *   bn_finalize : biased batch variance, mean, invstd; scale = gamma*invstd, shift = beta - mean*scale; running stats
 *                 updated with `momentum` (running_var with the unbiased variance).  `count` = rows (x world size when
 *                 the partial sums were all-reduced for SyncBatchNorm, tools/train.py:73-79).
 *   bn_apply    : z = act(y*scale + shift + residual) * nc_scale[n][c]   (residual: resnet.py:78-79; nc_scale: the
 *                 Dropout2d channel mask of module.py:60; all three optional) */
int segb200_bn_stats(const void* x, long long rows, int c, int x_ld, int dtype, float* partial, int max_slabs, void* stream);
int segb200_bn_finalize(const float* partial, int slabs, int c, double count, const float* gamma, const float* beta,
                        float* running_mean, float* running_var, float momentum, float eps, float* mean, float* invstd,
                        float* scale, float* shift, void* stream);
int segb200_bn_apply(const void* y, const float* scale, const float* shift, const void* residual, const float* nc_scale,
                     void* z, long long rows, long long rows_per_img, int c, int y_ld, int res_ld, int z_ld, int act,
                     int dtype, void* stream);
/* Backward of the same unit.  g = dz * nc_scale * act'(pre-activation); the activation mask is taken from the stored
 * output z when z != NULL (required when a residual entered the activation), else recomputed as y*scale + shift
 * (one activation read less per pass).
 *   bn_bwd_reduce   : per-slab sum(g), sum(g*y)
 *   bn_bwd_finalize : sums[2][c] = {sum(g), sum(g*xhat) = invstd*(sum(g*y) - mean*sum(g))}; dgamma += sums[1], dbeta += sums[0]
 *   bn_bwd_apply    : dy = scale*(g - sums[0]/count - xhat*sums[1]/count);  dres (+)= g   (sums == NULL: dy = g*scale) */
int segb200_bn_bwd_reduce(const void* dz, const void* z, const void* y, const float* scale, const float* shift,
                          const float* nc_scale, float* partial, long long rows, long long rows_per_img, int c, int dz_ld,
                          int z_ld, int y_ld, int act, int dtype, int max_slabs, void* stream);
int segb200_bn_bwd_finalize(const float* partial, int slabs, int c, const float* mean, const float* invstd, float* sums,
                            float* dgamma, float* dbeta, void* stream);
int segb200_bn_bwd_apply(const void* dz, const void* z, const void* y, const float* mean, const float* invstd,
                         const float* scale, const float* shift, const float* sums, double count, const float* nc_scale,
                         void* dy, void* dres, int dres_accumulate, long long rows, long long rows_per_img, int c, int dz_ld,
                         int z_ld, int y_ld, int dy_ld, int dres_ld, int act, int dtype, void* stream);

/* MaxPool2d(3,2,1) backward (first-maximum rule of torch), gather form: dx [n][h][w][dx_ld]. */
int segb200_maxpool3x3s2_bwd(const void* x, const void* dy, void* dx, int n, int h, int w, int c, int x_ld, int dy_ld,
                             int dx_ld, int dtype, void* stream);
/* The pair the training plan uses: the forward also records the argmax tap (uint8 [n][ho][wo][c], 0..8, first maximum), the
 * backward reads 4 x (index + gradient) per input vector instead of re-scanning the windows. */
int segb200_maxpool3x3s2_idx(const void* x, void* y, uint8_t* idx, int n, int h, int w, int c, int x_ld, int y_ld, int dtype,
                             void* stream);
int segb200_maxpool3x3s2_bwd_idx(const uint8_t* idx, const void* dy, void* dx, int n, int h, int w, int c, int dy_ld, int dx_ld,
                                 int dtype, void* stream);
/* Backward of segb200_bilinear_nhwc (gather form): dx [n][hi][wi] (+)= gscale * W^T dy [n][ho][wo]; gscale = optional
 * DEVICE pointer to one float (the 1/valid-pixel-count of the loss). */
int segb200_bilinear_nhwc_bwd(const void* dy, void* dx, int n, int hi, int wi, int c, int dx_ld, int ho, int wo, int dy_ld,
                              int align_corners, int accumulate, const float* gscale, int dtype, void* stream);
/* Fused final up-sampling + nn.CrossEntropyLoss(ignore_index) (models/deeplabv3_plus.py:39 + solver/loss.py:16-46): the
 * [n][nclass][ho][wo] logits tensor is never materialised.  target: int64 [n][ho][wo].  Writes
 *   dfull [n][ho][wo][d_ld] = softmax - onehot (0 at ignored pixels), out3 = {mean loss, 1/valid, valid};
 * partial: workspace of 2*segb200_upsample_ce_blocks() floats. */
int segb200_upsample_ce_blocks(int n, int ho, int wo);
int segb200_upsample_ce(const void* logits, const long long* target, void* dfull, float* partial, float* out3, int n,
                        int hi, int wi, int nclass, int x_ld, int ho, int wo, int d_ld, int align_corners,
                        int ignore_index, int dtype, void* stream);
/* Depthwise 3x3 weight gradient (stride 1, padding = dilation): partial[(slab*9 + tap)*c + ch]; finish with
 * segb200_reduce_partials(k = 9).  The depthwise DATA gradient is segb200_dwconv3x3 with tap-reversed weights. */
int segb200_dw_wgrad(const void* x, const void* dy, float* partial, int n, int h, int w, int c, int x_ld, int dy_ld,
                     int dilation, int pre_relu, int dtype, int max_slabs, void* stream);
/* y[b][p][:] (+)= v[b][:] * scale -- gradient of the global average pool / broadcast of a pooled feature. */
int segb200_nc_broadcast(const void* v, void* y, int n, long long hw, int c, int v_ld, int y_ld, float scale,
                         int accumulate, int dtype, void* stream);
/* Stride-2 data gradients.  t [n][ceil(h/2)][ceil(w/2)], z [n][h][w].  mode 0: z = zero-inserted t (then a stride-1
 * segb200_conv_gemm with flipped weights is the 3x3/2 data gradient); mode 1: z[2i][2j] += t[i][j] (1x1/2). */
int segb200_stride2_place(const void* t, void* z, int n, int h, int w, int c, int t_ld, int z_ld, int mode, int dtype,
                          void* stream);
/* Weight packing by index table: dst[i] = index[i] >= 0 ? (dst_dtype) src[index[i]] : 0; and its adjoint. */
int segb200_gather_cast(const float* src, const int* index, void* dst, long long n, int dst_dtype, void* stream);
int segb200_scatter_add(const float* src, const int* index, float* dst, long long n, void* stream);
/* torch.optim.SGD(momentum, weight_decay) step on flat fp32 buffers (solver/optimizer.py:50-51). */
int segb200_sgd_step(float* p, const float* g, float* m, long long n, float lr, float momentum, float weight_decay,
                     float grad_scale, void* stream);

/* Criss-cross attention BACKWARD (replaces _C.ca_backward / _C.ca_map_backward, csrc/criss_cross_attention/ca_cuda.cu:38-92,
 * :122-177, plus the softmax backward and the gamma gradient autograd runs around them, modules/cc_attention.py:62-72).
 *   cca_weight_bwd : D[p][z] = dy[p].v[key(p,z)] (== ca_map_backward's dw), dE = gamma * A * (D - sum_z A D)  -> de [n][h][w][att_ld];
 *                    dgamma_partial[block] = per-block sum of A.D (finish with segb200_reduce_partials; segb200_cca_weight_bwd_blocks())
 *   cca_gather     : out[p] (+)= scale * sum_z a[p][z] src[key(p,z)]        (ca_backward's dt with a = dE, src = k; == ca_map_forward)
 *   cca_scatter    : out[r] (+)= scale * sum_{(p,z): key(p,z)=r} a[p][z] src[p]   (ca_backward's df with a = dE, src = q;
 *                    ca_map_backward's dg with a = A, src = dy, scale_dev = gamma); scale_dev: optional DEVICE float multiplied in. */
int segb200_cca_weight_bwd_blocks(int n, int h, int w);
int segb200_cca_weight_bwd(const void* dy, const void* v, const float* att, float* de, float* dgamma_partial, const float* gamma,
                           int n, int h, int w, int c, int dy_ld, int v_ld, int att_ld, int dtype, void* stream);
int segb200_cca_gather(const float* a, const void* src, void* out, int n, int h, int w, int c, int a_ld, int src_ld, int out_ld,
                       float scale, int accumulate, int dtype, void* stream);
int segb200_cca_scatter(const float* a, const void* src, void* out, int n, int h, int w, int c, int a_ld, int src_ld, int out_ld,
                        float scale, const float* scale_dev, int accumulate, int dtype, void* stream);

/* ------------------------------------------------------------------------------------------
 * SyncBatchNorm statistics exchange fused into the finalize kernels, over NVLink peer memory (csrc/syncbn.cu).
 * Replaces the two per-layer collectives of `nn.SyncBatchNorm` (tools/train.py:73-79; NaiveSyncBatchNorm's all_reduce,
 * modules/batch_norm.py:161-183): the finalize kernel of every rank stores its [2][c] partial sums into every peer's symmetric
 * buffer (P2P stores), releases a per-(slot, source rank, CTA) epoch flag, waits for all peers' flags and sums the contributions
 * in rank order (bit-identical on all ranks).  `peers_dev` = device array [world] of the ranks' buffer base pointers as mapped in
 * THIS process (torch.distributed._symmetric_memory rendezvous); slot s of a buffer holds data at float offset `data_off` =
 * s * segb200_syncbn_slot_floats(world, cmax) and flags at u32 offset `flag_off` (host-assigned, identical on all ranks);
 * `epoch_dev` = device u32 bumped once per training step (segb200_counter_add) before the step's first exchange; flags must be
 * zero-initialised and epochs start at 1.  world == 1 degenerates to the plain finalize.
 *   bn_finalize_sync     : partial [slabs][2][c] (local) -> exchange -> mean / invstd / scale / shift (+ running stats), count_total
 *                          = rows of ALL ranks
 *   bn_bwd_finalize_sync : partial (local sum g, sum g*y) -> dgamma / dbeta (+)= local sums -> exchange -> sums [2][c] over all ranks */
int segb200_syncbn_slot_floats(int world, int cmax);
int segb200_syncbn_slot_flags(int world);
int segb200_counter_add(void* counter_u32, int value, void* stream);
int segb200_bn_finalize_sync(const float* partial, int slabs, int c, double count_total, const float* gamma, const float* beta,
                             float* running_mean, float* running_var, float momentum, float eps, float* mean, float* invstd,
                             float* scale, float* shift, const void* peers_dev, int world, int rank, int cmax, long long data_off,
                             long long flag_off, const void* epoch_dev, void* stream);
int segb200_bn_bwd_finalize_sync(const float* partial, int slabs, int c, const float* mean, const float* invstd, float* sums,
                                 float* dgamma, float* dbeta, const void* peers_dev, int world, int rank, int cmax,
                                 long long data_off, long long flag_off, const void* epoch_dev, void* stream);

/* ------------------------------------------------------------------------------------------
 * The reference's only native module, `segmentron._C` (segmentron/modules/csrc/vision.cpp:6-11), function by function, on the
 * reference's OWN layout: contiguous NCHW tensors of element type `dtype` (SEGB200_F32 / F16 / BF16), fp32 accumulation,
 * outputs fully written (no zero-initialisation needed), no atomics.  Signatures follow csrc/criss_cross_attention/ca.h:25-72:
 *   ca_forward(t, f) -> weight            t, f [n][c][h][w];  weight [n][h+w-1][h][w]           (ca_cuda.cu:8-36, :188-212)
 *   ca_backward(dw, t, f) -> dt, df       dw like weight;  dt, df like t                          (ca_cuda.cu:38-92, :214-248)
 *   ca_map_forward(weight, g) -> out      g, out [n][c][h][w]                                     (ca_cuda.cu:94-120, :250-276)
 *   ca_map_backward(dout, weight, g) -> dw, dg                                                   (ca_cuda.cu:122-177, :278-312)
 * segmentron_b200/c_shim.py wraps them as a module object with exactly the reference's four Python-visible functions so that
 * segmentron/modules/cc_attention.py:11-45 (_CAWeight / _CAMap) runs unchanged. */
int segb200_ca_forward(const void* t, const void* f, void* weight, int n, int c, int h, int w, int dtype, void* stream);
int segb200_ca_backward(const void* dw, const void* t, const void* f, void* dt, void* df, int n, int c, int h, int w, int dtype,
                        void* stream);
int segb200_ca_map_forward(const void* weight, const void* g, void* out, int n, int c, int h, int w, int dtype, void* stream);
int segb200_ca_map_backward(const void* dout, const void* weight, const void* g, void* dw, void* dg, int n, int c, int h, int w,
                            int dtype, void* stream);

/* Backward of segb200_upsample_add, y = act(a + nearest_up_{2^k}(z)) -- the HRNet fuse sum (backbones/hrnet.py:178-186,215-232):
 * g = dy * [y > 0] (act = relu, mask from the stored output); da (+)= g; dz (+)= sum of g over each 2^k x 2^k block.  da or dz may
 * be NULL. */
int segb200_upsample_add_bwd(const void* dy, const void* y, void* da, void* dz, int n, int h, int w, int c, int dy_ld, int y_ld,
                             int da_ld, int dz_ld, int k, int act, int accumulate_a, int accumulate_z, int dtype, void* stream);

/* CAM_Module backward glue (modules/module.py:134-162; autograd through bmm / softmax / bmm in the reference).  With
 * G[c1][c2] = sum_p dy[p,c1] x[p,c2] (a segb200_conv_gemm with fp32 output):
 *   cam_softmax_bwd : r[c1] = sum_c2 A G -> dgamma_partial[c1];  dE = -(*gamma) * A * (G - r)         (fp32 [c][de_ld])
 *   cam_bwd_pack    : w1[c2][c1] = (*gamma) A[c1][c2],  w2 = dE + dE^T   as 16-bit GEMM weights [c][w_ld] (zero K padding), so that
 *                     dx = dy + conv_gemm(dy, w1) + conv_gemm(x, w2). */
int segb200_cam_softmax_bwd(const void* att, const float* g, const float* gamma, float* de, float* dgamma_partial, int c, int att_ld,
                            int g_ld, int de_ld, int dtype, void* stream);
int segb200_cam_bwd_pack(const void* att, const float* de, const float* gamma, void* w1, void* w2, int c, int att_ld, int de_ld,
                         int w_ld, int dtype, void* stream);

/* Backward of segb200_adaptive_avgpool (nn.AdaptiveAvgPool2d(s) of PyramidPooling, modules/module.py:82-97): dx[p] (+)= sum over the
 * (possibly overlapping) bins containing p of dy[bin] / area(bin). */
int segb200_adaptive_avgpool_bwd(const void* dy, void* dx, int n, int h, int w, int c, int dy_ld, int dx_ld, int s, int accumulate,
                                 int dtype, void* stream);

/* TRAINING-mode PAM_Module (modules/module.py:100-131) on a materialised attention matrix, like the reference's bmm / softmax / bmm:
 *   row_softmax     : P[r][j] = softmax_j(S[r][j]) for j < n (16-bit), columns n .. p_ld-1 = 0   (S: fp32 GEMM output Q K^T)
 *   row_softmax_bwd : r = sum_j P D -> partial[r] (the gamma-gradient term);  dS = (*gamma) * P * (D - r), zero padded
 *                     (D = dy V^T, fp32 GEMM output).  dV, dQ, dK then are tcgen05 GEMMs over P^T, dS, dS^T (segb200_nhwc_to_cn). */
int segb200_row_softmax(const float* s, void* p, int rows, int n, int s_ld, int p_ld, int dtype, void* stream);
int segb200_row_softmax_bwd(const void* p, const float* d, const float* gamma, void* ds, float* partial, int rows, int n, int p_ld,
                            int d_ld, int ds_ld, int dtype, void* stream);

/* OPT-IN variant of segb200_dw_wgrad (same arguments, same partial[(slab*9 + tap)*c + ch] layout, slab count from
 * segb200_dw_wgrad_v2_slabs): a thread owns four channels and walks contiguous pixels of an image row with the 3x3 window held in
 * registers (3 new 8-byte loads per pixel instead of 9 16-byte ones) and packed fma.rn.f32x2.  Not the default until measured. */
int segb200_dw_wgrad_v2_slabs(long long rows, int c, int max_slabs);
int segb200_dw_wgrad_v2(const void* x, const void* dy, float* partial, int n, int h, int w, int c, int x_ld, int dy_ld,
                        int dilation, int pre_relu, int dtype, int max_slabs, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * EVALUATION METRIC -- replaces segmentron/utils/score.py:83-113 (batch_pix_accuracy,
 * batch_intersection_union: two argmax passes, three host transfers and three torch.histc calls behind a device synchronise,
 * score.py:49,108-110) and the accumulation of SegmentationMetric.update (score.py:50-55).
 *   counts: unsigned 64-bit [2 + 3*nclass], ACCUMULATED by seg_metric / seg_metric_lowres:
 *     [0] correct (label >= 0 and argmax of the logits TRUNCATED to integers == label, score.py:86-90), [1] labeled (label >= 0),
 *     [2+c] inter[c], [2+nclass+c] pred[c] (argmax of the float logits over labeled pixels), [2+2*nclass+c] lab[c] (label == c);
 *     argmax ties go to the lowest class (torch.argmax).  target: int64 [n][h][w], negative = ignored.
 *   seg_metric        : logits = full-resolution NCHW [n][nclass][h][w], dtype SEGB200_BF16 / F16 / F32.
 *   seg_metric_lowres : logits = the classifier's NHWC 16-bit output [n][hi][wi][x_ld]; the final bilinear up-sampling to (ho, wo)
 *                       (deeplabv3_plus.py:39) is fused with the arithmetic and out_dtype rounding of segb200_bilinear_nchw_out.
 *   seg_metric_accumulate : total_pixels[0..1] (int64) += correct, labeled; total_inter[c] (f32) += inter; total_union[c] (f32) +=
 *                       pred + lab - inter (score.py:112, float32 like the reference's totals); then counts := 0.  No host sync.
 * 1 <= nclass <= 64. */
int segb200_seg_metric(const void* logits, int dtype, const long long* target, int n, int nclass, int h, int w,
                       unsigned long long* counts, void* stream);
int segb200_seg_metric_lowres(const void* logits_nhwc, int dtype, int x_ld, int hi, int wi, int align_corners, int out_dtype,
                              const long long* target, int n, int nclass, int ho, int wo, unsigned long long* counts,
                              void* stream);
int segb200_seg_metric_accumulate(unsigned long long* counts, int nclass, long long* total_pixels, float* total_inter,
                                  float* total_union, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * MULTI-SCALE + FLIP EVALUATION -- the data movement of SegBaseModel.evaluate (segmentron/models/segbase.py:44-79).
 *   eval_prepare    : image fp32 NCHW [b][c][h][w] -> out fp32 [(1+flip) b][c][hp][wp]: images 0..b-1 = F.interpolate(image,
 *                     (height, width), bilinear, align_corners=True) (:63, _resize_image :82) zero padded at the bottom / right
 *                     to (hp, wp) (_pad_image :86-107); images b..2b-1 (flip != 0) = the same, horizontally flipped AFTER the
 *                     padding (_flip_image :114, applied to the padded image :71), so one model call serves both passes.
 *   eval_accumulate : logits [(1+flip) b][k][hp][wp] (dtype) -> scores [b][k][h][w] (same dtype):
 *                     outputs = logits[0:b][..., :height, :width] (+ flip(logits[b:2b])[..., :height, :width], rounded to dtype, :69-71);
 *                     scores (accumulate ? += : =) F.interpolate(outputs, (h, w), bilinear, align_corners=True) (:73-78). */
int segb200_eval_prepare(const float* image, float* out, int b, int c, int h, int w, int height, int width, int hp, int wp,
                         int flip, void* stream);
int segb200_eval_accumulate(const void* logits, void* scores, int dtype, int b, int k, int hp, int wp, int height, int width,
                            int h, int w, int flip, int accumulate, void* stream);

/* GPU-side input transform: transforms.ToTensor() + transforms.Normalize(mean, std) (tools/train.py:36-39) on a
 * batch of decoded uint8 HWC images on the device: out[n][c][y][x] = (img[n][y][x][c] / 255 - mean[c]) / std[c], fp32, bit-identical
 * to torchvision.  mean / stdv: device pointers to c floats; c <= 4. */
int segb200_image_normalize(const unsigned char* img, float* out, int n, int h, int w, int c, const float* mean, const float* stdv,
                            void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SEGB200_H_ */
