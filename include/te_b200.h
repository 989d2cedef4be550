/*
 * te_b200.h — C ABI of libte_b200.so, the B200 (sm_100a) kernels behind TransEditor's
 * generator / discriminator hot path.
 *
 * Conventions (all entry points):
 *   - plain pointers and sizes only; every data pointer is a DEVICE pointer unless noted;
 *   - nothing allocates, nothing synchronises, everything is enqueued on `stream`
 *     (a cudaStream_t passed as void*; NULL = legacy default stream);
 *   - returns 0 on success, a negative te_status otherwise; te_last_error() gives the
 *     message of the last failure on the calling thread;
 *   - stateless / re-entrant apart from a per-process cache of immutable TMA descriptors.
 *
 * Each entry point names the reference interface it replaces (paths relative to the
 * TransEditor tree).  INTEGRATION.md shows the binding a reference maintainer adds.
 */
#ifndef TE_B200_H
#define TE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum { TE_F32 = 0, TE_BF16 = 1, TE_F16 = 2, TE_F64 = 3 } te_dtype;
/* OR-ed into te_fused_bias_act's dtype: x/out are bf16 or f16 while `bias` points to FLOAT32 values
 * (the f32 master parameter is used directly, no per-call down-cast kernel). */
#define TE_BIAS_F32 0x100

typedef enum {
  TE_OK = 0,
  TE_ERR_INVALID = -1,     /* bad argument (shape, dtype, alignment, null pointer) */
  TE_ERR_UNSUPPORTED = -2, /* valid request this build has no kernel for */
  TE_ERR_CUDA = -3         /* a CUDA runtime / driver call failed; see te_last_error() */
} te_status;

/* Library version (major*1000 + minor) and last error text of the calling thread. */
int te_version(void);
const char* te_last_error(void);

/* ---------------------------------------------------------------------------------------------
 * fused bias + activation.
 * Replaces: fused.fused_bias_act(input, bias, refer, act, grad, alpha, scale)
 *           utils/op/fused_bias_act.cpp:11-17 -> fused_bias_act_kernel.cu:52-99 (kernel :18-49).
 *   x' = x + bias[(i / step_b) % size_b]            (bias == NULL: no bias)
 *   act*10+grad: 10,11 -> y = x'        12,32 -> y = 0
 *                30    -> y = x' > 0 ? x' : alpha*x'
 *                31    -> y = ref > 0 ? x' : alpha*x'   (ref == NULL reads as 0, as the reference)
 *   out = y * scale
 * `n` elements, contiguous.  NCHW: step_b = H*W, size_b = C.  Channels-last: step_b = 1.
 * dtypes: f32, bf16, f16, f64 (bias/ref same dtype as x).
 */
int te_fused_bias_act(void* out, const void* x, const void* bias, const void* ref, int act,
                      int grad, float alpha, float scale, int64_t n, int64_t step_b,
                      int64_t size_b, int dtype, void* stream);

/* Backward of the above in ONE pass (replaces FusedLeakyReLUFunctionBackward.forward,
 * utils/op/fused_act.py:20-38, which runs fused_bias_act(grad=1) and then a separate
 * `.sum()` reduction that re-reads the tensor):
 *   grad_in = (ref > 0 ? g : alpha*g) * scale ;  grad_bias[c] += sum of grad_in over channel c.
 * grad_bias is FLOAT32 [size_b] (FLOAT64 when dtype is TE_F64) and must be zeroed by the caller
 * (accumulated with atomics); pass NULL to skip the reduction. */
int te_fused_bias_act_bwd(void* grad_in, void* grad_bias, const void* g, const void* ref,
                          float alpha, float scale, int64_t n, int64_t step_b, int64_t size_b,
                          int dtype, void* stream);

/* ---------------------------------------------------------------------------------------------
 * upfirdn2d: zero-insert upsample -> pad/crop -> 2-D FIR (true convolution) -> decimate.
 * Replaces: upfirdn2d.upfirdn2d(input[major,in_h,in_w,minor], kernel[kh,kw], up_x, up_y,
 *           down_x, down_y, pad_x0, pad_x1, pad_y0, pad_y1)
 *           utils/op/upfirdn2d.cpp:12-19 -> upfirdn2d_kernel.cu:140-272 (kernel :52-137).
 * out[major,out_h,out_w,minor], out_h = (in_h*up_y + pad_y0 + pad_y1 - kh)/down_y + 1 (:167-168).
 * `fir` is FLOAT32 [kh,kw] on the device, kh,kw <= 16.  Unlike the reference, EVERY parameter
 * combination is computed (the reference launches nothing for unmatched modes, :172,216).
 * NCHW tensors: major = N*C, minor = 1.  Channels-last: major = N, minor = C.
 * dtypes: f32, bf16, f16, f64.
 */
int te_upfirdn2d(void* out, const void* in, const float* fir, int64_t major, int in_h, int in_w,
                 int minor, int kh, int kw, int up_x, int up_y, int down_x, int down_y, int pad_x0,
                 int pad_x1, int pad_y0, int pad_y1, int dtype, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Gather convolution (SIMT, f32/f64, NCHW) — the full-precision parity engine.
 * Replaces the F.conv2d / F.conv_transpose2d(groups=batch) calls of ModulatedConv2d
 * (model_spatial_query.py:318,327,333) and EqualConv2d (:177-183) WITHOUT materialising
 * per-sample weights: conv(x, W*s*d) == d * conv(x*s, W)  (SURVEY.md App. A.5).
 *
 *   y[b,o,oy,ox] = act( out_scale[b,o] * SUM_{i,ky,kx} W(o,i,ky,kx) * in_scale[b,i] * x[b,i,iy,ix]
 *                       + noise_w[0]*noise[b*noise_bstride + oy*wout + ox] + bias[o] )
 *   ty = oy*down + ky - pad_y ; tap valid iff ty >= 0, ty % up == 0, iy = ty/up < hin (same for x)
 *   W(o,i,ky,kx) = w[o*w_so + i*w_si + (flip ? (kh-1-ky)*kw + (kw-1-kx) : ky*kw + kx)]
 *   act: 0 = identity, 1 = leaky_relu(0.2)*sqrt(2) (FusedLeakyReLU, utils/op/fused_act.py:72-90)
 * in_scale, out_scale, bias, noise may be NULL.  Covers conv2d (up=1, down=stride), conv_transpose2d
 * (up=stride, flip=1, pad=k-1-p) and every data-gradient of those (swap up/down, swap w_so/w_si,
 * toggle flip, pad -> k-1-pad).
 */
typedef struct {
  int batch, cin, hin, win, cout, hout, wout, kh, kw;
  int up, down, pad_y, pad_x, flip;
  int64_t w_so, w_si; /* element strides of the weight's output / input channel index */
  int act;
  int64_t noise_bstride; /* 0: one noise map shared by the batch */
} te_conv_geom;

int te_conv2d_simt(void* y, const void* x, const void* w, const void* in_scale,
                   const void* out_scale, const void* bias, const void* noise, const void* noise_w,
                   const te_conv_geom* g, int dtype, void* stream);

/* Weight gradient of the gather convolution (same geometry struct; act/noise ignored):
 *   gw(o,i,ky,kx) += SUM_{b,oy,ox} out_scale[b,o]*gy[b,o,oy,ox] * in_scale[b,i]*x[b,i,iy,ix]
 * written through the same (w_so, w_si, flip) addressing; gw must be zeroed by the caller
 * (split-K accumulation with atomics).  Replaces autograd's conv weight-gradient of the calls above. */
int te_conv2d_wgrad_simt(void* gw, const void* x, const void* gy, const void* in_scale,
                         const void* out_scale, const te_conv_geom* g, int dtype, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Fused Adam + EMA over flat parameter buffers (f32).
 * Replaces: torch.optim.Adam.step (train_spatial_query.py:464-473 use) and the per-parameter
 * EMA loop `accumulate` (train_spatial_query.py:56-61), one launch instead of ~650.
 *   m = b1*m + (1-b1)*g ; v = b2*v + (1-b2)*g*g
 *   p -= lr * (m/(1-b1^t)) / (sqrt(v/(1-b2^t)) + eps)          (torch.optim.Adam semantics)
 *   ema = ema_decay*ema + (1-ema_decay)*p    (ema == NULL: skipped)
 *   grad_scale multiplies g first (1/world_size after a sum-all-reduce).
 */
int te_adam_ema(float* p, const float* g, float* m, float* v, float* ema, int64_t n, float lr,
                float beta1, float beta2, float eps, int step, float ema_decay, float grad_scale,
                void* stream);

/* Same update with the step count t read from DEVICE memory (*step_dev >= 1), so the launch can live
 * inside a captured CUDA graph and be replayed every iteration. */
int te_adam_ema_devstep(float* p, const float* g, float* m, float* v, float* ema, int64_t n, float lr,
                        float beta1, float beta2, float eps, const int* step_dev, float ema_decay,
                        float grad_scale, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Tensor-core (tcgen05 + TMA + TMEM) implicit-GEMM convolution, bf16 operands / f32 accumulate,
 * channels-last: the speed path of ModulatedConv2d / EqualConv2d (same reference call sites as
 * te_conv2d_simt: model_spatial_query.py:177-183,318,327,333) and of their data gradients.
 *
 * Geometry = a TAP TABLE over an ANCHOR grid (a = (ay, ax), 0 <= ay < grid_h, 0 <= ax < grid_w):
 *   y[b, ay*out_stride + out_off_y, ax*out_stride + out_off_x, :] =
 *       act( out_scale[b,:] * SUM_t  x[b, ay*in_stride + tap_dy[t], ax*in_stride + tap_dx[t], :] . W[tap_w[t]]^T
 *            + bias[:] )
 *   x [batch, hin, win, cin] bf16 ; W [w_slices, cout, cin] bf16 (K contiguous) ; y [batch, hout, wout, cout]
 *   bf16 (out_f32: float).  Reads outside x are zeros (that is the padding).
 *   conv2d stride 1:      in_stride 1, out_stride 1, taps (ky-p, kx-p)
 *   conv2d stride 2:      in_stride 2, out_stride 1, taps (ky, kx)
 *   conv_transpose2d s2:  in_stride 1, out_stride 2, one call per output parity class (out_off)
 * in_scale is not applied here: the caller folds the modulation into x or into per-sample weights
 * (w_bstride != 0: sample b uses W + b*w_bstride, requires >= 128 anchors per sample).
 * Requirements: cin % 8 == 0, cout % 8 == 0, 16-byte aligned pointers.
 */
typedef struct {
  int batch, hin, win, cin;
  int hout, wout, cout;
  int ntaps;                 /* 1..9 */
  int tap_dy[9], tap_dx[9];  /* input offset of tap t */
  int tap_w[9];              /* weight slice of tap t */
  int w_slices;              /* slices in W (per sample when w_bstride != 0) */
  int in_stride, out_stride, out_off_y, out_off_x;
  int grid_h, grid_w;
  int act;                   /* 0 identity, 1 leaky_relu(0.2) * act_gain, 2 leaky_relu(0.01) * act_gain (nn.LeakyReLU), 3 PReLU(slope[c]) */
  int out_f32;
  int64_t w_bstride;
  /* optional extras; an all-zero tail keeps the plain behaviour */
  float act_gain;            /* gain after the leaky ReLU; 0 selects sqrt(2) (FusedLeakyReLU's scale) for act 1, 1 for act 2 */
  float wgrad_alpha;         /* te_conv_wgrad_tc accumulates wgrad_alpha * gradient; 0 selects 1 */
  const void* residual;      /* [batch, hout, wout, cout] in the OUTPUT's dtype (bf16, or float when out_f32), added
                                AFTER bias/activation (ResBlock skip sum, model_spatial_query.py:795-797), or NULL */
  const void* slope;         /* act 3: FLOAT32 [cout] negative slopes (nn.PReLU of the pSp trunk, helpers.py:90,112) */
  int split;                 /* 0/1: x, W (and g of the weight gradient) are plain bf16 tensors.  2 or 3: the
                                SPLIT-OPERAND fp32 mode — each of them is `split` bf16 planes [split][...same layout...]
                                holding hi = bf16(v), mid = bf16(v - hi)(, lo = bf16(v - hi - mid)) of f32 values v
                                (te_split_bf16, te_pack_weights_tc); the kernel sums the plane-pair products
                                hi*hi + hi*mid + mid*hi (+ mid*mid + hi*lo + lo*hi) in the f32 accumulator: the
                                reference's fp32 F.conv2d arithmetic to ~2^-16 (2^-24) per product, on tensor cores */
  int reserved;
} te_tc_conv_desc;

int te_conv_tc(void* y, const void* x, const void* w, const float* out_scale, const float* bias,
               const te_tc_conv_desc* d, void* stream);

/* Weight gradient for the same geometry (act/out_f32/w_bstride ignored):
 *   gw[tap_w[t]][m][n] += SUM_{b, a} g[b, a*out_stride + out_off, m] * x[b, a*in_stride + tap_d[t], n]
 * g [batch, hout, wout, cout] bf16, x [batch, hin, win, cin] bf16, gw [w_slices, cout, cin] FLOAT32,
 * accumulated (caller zeroes it).  Replaces autograd's convolution weight gradient. */
int te_conv_wgrad_tc(float* gw, const void* g, const void* x, const te_tc_conv_desc* d, void* stream);

/* te_conv_wgrad_tc's tap-major accumulator -> the master weight's layout (what autograd hands to the optimiser):
 *   trans 0:  out[b][o][i][t] = ws[b][t][o*ld + i]        trans 1:  out[b][o][i][t] = ws[b][t][i*ld + o]
 * ws [batch][taps][rows][ld] f32 (rows x ld = the padded (cout, cin) of the convolution descriptor), out
 * [batch][o_dim][i_dim][taps] f32 = the [O, I, k, k] weight gradient (batch > 1: per-sample weights); taps = 1, 4
 * or 9.  clear != 0: every element read is written back as 0, so the accumulator can be reused by the next
 * te_conv_wgrad_tc launch without a memset.  Replaces the permute + contiguous copy after every weight-gradient
 * launch (and, with clear, the zero fill before it). */
int te_wgrad_unpack(float* out, float* ws, int batch, int o_dim, int i_dim, int taps, int rows, int ld, int trans,
                    int clear, void* stream);

/* Convenience form: stride 1, padding k/2, kh = kw in {1,3}, bf16 output. */
int te_conv2d_tc(void* y, const void* x, const void* w, const float* out_scale, const float* bias,
                 int batch, int hin, int win, int cin, int cout, int kh, int kw, int act,
                 int64_t w_bstride, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Channels-last per-(sample, channel) scaling and its gradient reduction: the modulation s[b,cin] and
 * demodulation d[b,cout] multiplies of ModulatedConv2d applied to ACTIVATIONS instead of being folded
 * into per-sample weights (model_spatial_query.py:299-304).
 *   te_scale_bc:  y[b,p,c] = x[b,p,c] * s[b,c]            x,y [batch, pixels, channels], s FLOAT32
 *   te_dot_bc:    out[b,c] += SUM_p a[b,p,c] * b[b,p,c]   out FLOAT32, zeroed by the caller
 * dtypes: bf16, f16, f32; channels % (16/sizeof(T)) == 0.
 */
int te_scale_bc(void* y, const void* x, const float* s, int64_t batch, int64_t pixels, int channels,
                int dtype, void* stream);
int te_dot_bc(float* out, const void* a, const void* b, int64_t batch, int64_t pixels, int channels,
              int dtype, void* stream);

/* Split-operand planes of an f32 channels-last activation tensor, optionally modulated on the way:
 *   v = x[b,p,c] * (s ? s[b,c] : 1) ;  dst[0] = hi = bf16(v), dst[1] = bf16(v - hi), dst[2] = bf16(v - hi - mid)
 * x [batch, pixels, channels] FLOAT32, dst [nseg][batch, pixels, channels] bf16 (nseg = 1, 2 or 3; nseg 1 is a plain
 * rounding conversion), s FLOAT32 [batch, channels] or NULL.  channels % 8 == 0, 16-byte aligned pointers.
 * Feeds te_conv_tc / te_conv_wgrad_tc with te_tc_conv_desc.split = nseg: the fp32 parity mode of the reference's
 * F.conv2d / F.conv_transpose2d calls (model_spatial_query.py:177-183,318,327,333) on tensor cores. */
int te_split_bf16(void* dst, const float* x, const float* s, int64_t batch, int64_t pixels, int channels, int nseg,
                  void* stream);

/* Repack a TABLE of f32 master weights [out_ch, in_ch, K, K] (K*K = taps = 1 or 9, contiguous) into te_conv_tc's bf16
 * operand layouts in ONE launch: dst_n = [taps, out_pad, in_pad] (in_ch contiguous; the forward convolution's weights),
 * dst_t = [taps, in_pad, out_pad] (the data gradient's), each value times `scale`; pads = channel counts rounded up to
 * 8, never written (allocate the buffers zeroed once).  Either destination may be NULL.  `tasks` is a HOST array.
 * Replaces the `weight * self.scale` elementwise kernel the reference runs on every forward
 * (model_spatial_query.py:178,299) plus the layout change cuDNN does internally: done once per optimiser step. */
typedef struct te_pack_task {
  const float* src;
  void* dst_n;
  void* dst_t;
  int out_ch, in_ch, taps;
  float scale;
  int split; /* 0/1: one bf16 plane; 2, 3: that many split-operand planes [split][taps][..][..] (te_tc_conv_desc.split) */
} te_pack_task;
int te_pack_weights_tc(const te_pack_task* tasks, int n_tasks, void* stream);

/* Self-test of the tcgen05 GEMM core: D[M,N] (f32) = A[M,K] (bf16, K-major) * B[N,K]^T (bf16). */
int te_gemm_tc_selftest(float* d, const void* a, const void* b, int m, int n, int k, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Cross-attention core of the dual-space transformer (model_spatial_query.py:888-894):
 *   q [B,M,128] (from P), k,v [B,L,128] (from Z), 4 heads x 32, scale = 128^-0.5, M = L = 16
 *   sim[b,h,m,l] = softmax_l(q.k * scale) ; o[b,h,c,m] = SUM_l sim*v ; out = o viewed [B,128,L]
 *   permuted to [B,L,128] exactly as `sv.reshape(N, planes, L).permute(0, 2, 1)` (:894).
 * f32.  sim_out (may be NULL) receives the [B,4,M,L] similarity the reference returns.
 */
int te_attn_core(float* out, float* sim_out, const float* q, const float* k, const float* v,
                 int batch, int tokens, void* stream);

/* ---------------------------------------------------------------------------------------------
 * The whole dual-space interaction network in ONE launch.
 * Replaces: the `self.interact[i](x, spatialcode)` loop of Generator.forward
 *           (model_spatial_query.py:668-679) over AttentionBlock.forward (:920-936) and
 *           Attention.forward (:883-901): per block
 *     xn = layer_norm(x over [16, in_dim])                       (no affine, eps 1e-5)
 *     q = EL_q(P)  k = EL_k(xn)  v = EL_v(xn)                    (EL = EqualLinear, :194-221)
 *     att = softmax_l(q.k * 128^-0.5) v   per head (4 x 32)      (te_attn_core's definition)
 *     x1 = (in_dim != 512 ? EL_proj(x) : x) + EL_o(att)
 *     x2 = x1 + EL_m2(gelu_erf(EL_m1(layer_norm(x1))))
 *   EL(x) = x (W * lr_mul / sqrt(in))^T + b * lr_mul, W stored [out, in] row-major as the reference's parameter.
 * 16 query tokens (P) x 16 key tokens (Z), width 512, 128 attention planes; f32 in and out.
 * precision 0: products in 3xTF32 (error ~1e-6 relative: the fp32 parity bar holds); 1: single-pass TF32 (what
 * cuBLAS does under allow_tf32, ~1e-3).  One thread-block CLUSTER of 8 (or 4, chosen per batch so that the
 * samples fit the device in the fewest waves) CTAs per sample: every CTA owns 1/8 of each layer's output columns
 * and streams only that slice of the weights; activations are exchanged through distributed shared memory.
 *
 * Block 0 reads x0 [B,16,blocks[0].in_dim] and p0 [B,16,blocks[0].param_dim]; blocks >= 1 read the previous
 * block's output and p [B,16,512] (their in_dim and param_dim must be 512).  in_dim / param_dim: multiples of
 * 16 in [64, 528].  w_proj / b_proj are required iff in_dim != 512, else NULL.  n_blocks in [1, 8].  Weight
 * matrices (and the buffers receiving their gradients) must be 16-byte aligned.
 */
typedef struct te_attn_block {
  const float* w_proj; const float* b_proj; /* [512, in_dim], [512]   (AttentionBlock.proj)        */
  const float* w_q;    const float* b_q;    /* [128, param_dim], [128] (Attention.q_transform)     */
  const float* w_k;    const float* b_k;    /* [128, in_dim], [128]                                */
  const float* w_v;    const float* b_v;    /* [128, in_dim], [128]                                */
  const float* w_o;    const float* b_o;    /* [512, 128], [512]       (Attention.proj)            */
  const float* w_m1;   const float* b_m1;   /* [512, 512], [512]       (AttentionBlock.mlp[0])     */
  const float* w_m2;   const float* b_m2;   /* [512, 512], [512]       (AttentionBlock.mlp[2])     */
  int in_dim;
  int param_dim;
} te_attn_block;

/* Sizes (in floats) of the two workspaces below for this stack and batch. */
int te_attn_stack_workspace(const te_attn_block* blocks, int n_blocks, int batch, int64_t* save_floats,
                            int64_t* gws_floats);

/* How many clusters (= samples) of the forward / backward kernel the device holds at once
 * (cudaOccupancyMaxActiveClusters): [0] with 8 CTAs per sample, [1] with 4.  Batches beyond that run in waves. */
int te_attn_stack_occupancy(int fwd_clusters[2], int bwd_clusters[2]);

/* y [B,16,512].  `save` (save_floats, may be NULL for inference) receives what the backward pass needs. */
int te_attn_stack_fwd(float* y, const float* x0, const float* p0, const float* p,
                      const te_attn_block* blocks, int n_blocks, int batch, float lr_mul, int precision,
                      float* save, void* stream);

/* First-order backward of te_attn_stack_fwd (two launches: the per-sample data-gradient chain, then one
 * grouped weight-gradient GEMM over all B*16 rows).  gy [B,16,512] is the gradient of y.
 * Writes (never accumulates): g_x0, g_p0 (shapes of x0, p0), g_p [B,16,512] (summed over blocks >= 1; may be
 * NULL when n_blocks == 1), and through `grads` — the same struct with every pointer naming the OUTPUT buffer
 * of that parameter's gradient (in_dim/param_dim ignored).  `gws` is scratch of gws_floats. */
int te_attn_stack_bwd(float* g_x0, float* g_p0, float* g_p, const te_attn_block* grads, const float* gy,
                      const float* x0, const float* p0, const float* p, const te_attn_block* blocks,
                      int n_blocks, int batch, float lr_mul, int precision, const float* save, float* gws,
                      void* stream);

/* ---------------------------------------------------------------------------------------------
 * Data formats either side of the path.
 *
 * te_image_prep — the device-side tail of the reference's input pipeline: `self.transform(img)` in
 * utils/dataset.py:38-41 with the transform built at train_spatial_query.py:511-517
 * (RandomHorizontalFlip, ToTensor, Normalize(0.5, 0.5, inplace)).  The host keeps the entropy decode
 * (PIL, as in the reference) and the coin flips; the H2D copy carries uint8 pixels (1/4 of the f32 bytes).
 *   src_hwc  uint8 [B, H, W, 3]   decoded RGB pixels (np.asarray(PIL image))
 *   flip     uint8 [B] or NULL    1 = mirror that image horizontally
 *   dst_nchw f32 [B, 3, H, W] or NULL   ((x / 255) - 0.5) / 0.5, bit-identical to the torch ops
 *   dst_nhwc8 [B, H, W, 8] of nhwc_dtype (TE_F32 / TE_BF16) or NULL: the same values channels-last with
 *            the three channels zero-padded to 8 (the tensor-core from-RGB convolution's operand)
 *
 * te_image_quantize — network output to pixels: torchvision.utils.save_image(normalize=True,
 * range=(low, high)) per element (test_spatial_query.py:82-88, train_spatial_query.py:345-351):
 * clamp(x, low, high); (x - low) / max(high - low, 1e-5); * 255; + 0.5; clamp(0, 255); truncate.
 *   src   [B, 3, H, W] elements of `dtype` (TE_F32 / TE_BF16) addressed through the four element strides
 *         (NCHW or channels-last storage)
 *   dst_hwc uint8 [B, H, W, 3]
 */
int te_image_prep(float* dst_nchw, void* dst_nhwc8, const uint8_t* src_hwc, const uint8_t* flip, int batch,
                  int h, int w, int nhwc_dtype, void* stream);
int te_image_quantize(uint8_t* dst_hwc, const void* src, int batch, int h, int w, int64_t stride_b,
                      int64_t stride_c, int64_t stride_y, int64_t stride_x, float low, float high, int dtype,
                      void* stream);

/* ---------------------------------------------------------------------------------------------
 * Grouped EqualLinear (model_spatial_query.py:194-221) for the small-M linears of the path, many layers per
 * launch: the 2 x 16 per-column mapping linears incl. their PixelNorm (:75-81, :626-646), the style modulations
 * (ModulatedConv2d.modulation, :283), adjust_style (:686), the discriminator's final linears (:838-841), and —
 * with w_trans — their data gradients.  f32 in and out; products on TF32 tensor cores: precision 0 = 3 x TF32
 * (~1e-6 relative, the fp32 parity mode), 1 = single-pass TF32.
 *   y[m, n] = act( alpha * r[m] * SUM_k x[m*x_rs + k*x_cs] * B(k, n) + bias[n] * bias_mul )
 *   B(k, n) = w[n*w_ld + k] (w_trans 0) or w[k*w_ld + n] (w_trans 1);  y element (m, n) at y[m*y_rs + n*y_cs]
 *   r[m] = rsqrt(mean_k x[m, k]^2 + 1e-8) when pixel_norm (K <= 512), else 1; written to rnorm_out[m] if not NULL
 *   act 0: none; 1: leaky_relu(0.2) * sqrt(2) (fused_leaky_relu)
 *   k_splits > 1: the reduction is cut into that many CTAs which atomically ADD into y (caller zeroes y; no act).
 * `tasks` is a HOST array; all pointers in it are device pointers.
 */
typedef struct te_linear_task {
  const float* x; int64_t x_rs, x_cs;
  const float* w; int64_t w_ld; int w_trans;
  const float* bias; float bias_mul;
  float* y; int64_t y_rs, y_cs;
  int m, n, k;
  float alpha;
  int act, pixel_norm, k_splits;
  float* rnorm_out;
} te_linear_task;
int te_linear_grouped(const te_linear_task* tasks, int n_tasks, int precision, void* stream);

/* Weight / bias gradients of the same layers, many per launch (exact f32 FMA; output-bound):
 *   gw[n*k_dim + k] = alpha * SUM_m g[m*g_rs + n*g_cs] * x_scale[m] * x[m*x_rs + k*x_cs]
 *   gbias[n] = bias_mul * SUM_m g[m, n]     (gbias may be NULL; x_scale may be NULL = 1)
 * Writes, never accumulates. */
typedef struct te_linear_wgrad_task {
  const float* g; int64_t g_rs, g_cs;
  const float* x; int64_t x_rs, x_cs;
  const float* x_scale;
  float* gw; float* gbias;
  int m, n, k;
  float alpha, bias_mul;
} te_linear_wgrad_task;
int te_linear_wgrad_grouped(const te_linear_wgrad_task* tasks, int n_tasks, void* stream);

/* ---------------------------------------------------------------------------------------------
 * The discriminator's from-RGB layer: ConvLayer(3, C, 1) = EqualConv2d(3, C, 1) + FusedLeakyReLU(C)
 * (model_spatial_query.py:806-808 with :731-777, :156-185) as one streaming kernel each way.
 *   fwd:  y[b, p, o] = leaky_relu(SUM_c x[b, c, p] * w[o, c] * wscale + bias[o], slope) * gain
 *         x f32 NCHW [B, 3, H, W] (the loader's / generator's image), y channels-last [B, H, W, C] of `dtype`
 *         (TE_F32 / TE_BF16), w f32 [C, 3] (the [C, 3, 1, 1] master weight), bias f32 [C] or NULL.
 *   bwd:  with gp = g * (out > 0 ? gain : gain * slope)  (the mask comes from the saved OUTPUT, utils/op/fused_act.py:27-29)
 *         gw[o, c] += wscale * SUM_{b,p} gp * x      gbias[o] += SUM_{b,p} gp      (ACCUMULATE: zero them first)
 *         gx[b, c, p] = wscale * SUM_o gp * w[o, c]  (written; NULL = not needed)
 * C must be 8, 16, ..., 256 (a power of two).
 */
int te_from_rgb_fwd(void* y, const float* x, const float* w, const float* bias, int batch, int h, int wd, int cout,
                    float wscale, float slope, float gain, int dtype, void* stream);
int te_from_rgb_bwd(float* gw, float* gbias, float* gx, const void* g, const void* out, const float* x,
                    const float* w, int batch, int h, int wd, int cout, float wscale, float slope, float gain,
                    int dtype, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Blur followed by bias + leaky ReLU in one pass: the tail of an UPSAMPLING StyledConv,
 *   out = self.blur(conv_transpose2d(...)) (model_spatial_query.py:318-322) ... self.activate(out) (:398-400),
 * i.e. upfirdn2d(x, fir, up=1, down=1, pad) then fused_leaky_relu(., bias, slope, gain):
 *   out[n, y, x, c] = leaky_relu(SUM_k fir-taps * x + bias[c], slope) * gain
 * Channels-last tensors only ([major, H, W, minor], minor % (128 / sizeof(T)) == 0), FIR up to 4 x 4; bias FLOAT32
 * [minor].  Returns TE_ERR_UNSUPPORTED when the geometry is not the TMA-staged kernel's (the caller then runs
 * te_upfirdn2d and te_fused_bias_act).
 */
int te_upfirdn2d_bias_act(void* out, const void* x, const float* fir, const float* bias, int64_t major, int in_h,
                          int in_w, int minor, int kh, int kw, int pad_x0, int pad_x1, int pad_y0, int pad_y1,
                          float slope, float gain, int dtype, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Weight-only factor of ModulatedConv2d's demodulation (model_spatial_query.py:301-303 with the style factored out:
 * demod[b,o] = rsqrt(SUM_i style[b,i]^2 * energy[o,i] + eps)) and its gradient, f32:
 *   te_weight_energy:      energy[r] = coef * SUM_t w[r*taps + t]^2     r over the O*I rows of the [O,I,k,k] weight
 *   te_weight_energy_bwd:  gw[r*taps + t] = w[r*taps + t] * g[r] * coef2
 */
int te_weight_energy(float* energy, const float* w, int64_t rows, int taps, float coef, void* stream);
int te_weight_energy_bwd(float* gw, const float* w, const float* g, int64_t rows, int taps, float coef2, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TE_B200_H */
