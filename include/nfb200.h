/*
 * nfb200.h -- C ABI of libnfb200.so: B200 (sm_100a) kernels for the normalizing-flow hot path
 * of tatsy/normalizing-flows-pytorch (flow-layer forward / inverse + per-sample log|det J|).
 *
 * Conventions
 *   - Every pointer is a DEVICE pointer to contiguous fp32 data unless stated otherwise; the caller
 *     owns all memory.  No allocation, no host synchronisation, no stdout: every entry point only
 *     enqueues kernels on `stream` (a cudaStream_t passed as void*), so a whole stack is capturable
 *     in a CUDA graph.
 *   - Return value: 0 = success; > 0 = the cudaError_t of the launch; < 0 = argument error
 *     (NFB_ERR_*).  nfb_error_string() maps any of them to text.  Nothing throws across the ABI.
 *   - "z" tensors are (B, C, H, W) contiguous (NCHW) or (B, C) with H = W = 1.  D = C*H*W.
 *   - Log-det: `ldj_out[b] = ldj_in[b] + delta_b`.  Pass the same pointer twice for the reference's
 *     in-place `log_df_dz += ...` (coupling.py:110, modules.py:249,480); pass a different buffer for
 *     the layers that return a new tensor (Logit, modules.py:150).
 *   - z_out may alias z_in (in-place update of the transformed half; the pass-through half is then
 *     not touched at all).
 *   - Split modes (AbstractCoupling.__init__, coupling.py:16-30; index formulas in DESIGN.md):
 *       NFB_SPLIT_1D       squeeze1d / unsqueeze1d      (squeeze.py:64-83)   z0 = even entries
 *       NFB_SPLIT_CHECKER  checker_split / checker_merge (squeeze.py:32-61)  space-to-depth, blocks (a,d)|(b,c)
 *       NFB_SPLIT_CHANNEL  channel_split / channel_merge (squeeze.py:5-17)   first half | second half
 *     `odd` != 0 swaps the roles of the two halves.
 *   - The conditioner output `params` is the (B, P*c0, h, w) tensor the reference's `self.net(z1)`
 *     returns (coupling.py:105,176), consumed in place -- no slicing copies.
 */
#ifndef NFB200_H_
#define NFB200_H_

#ifdef __cplusplus
extern "C" {
#endif

typedef void* nfb_stream_t; /* cudaStream_t */

#define NFB_SPLIT_1D 0
#define NFB_SPLIT_CHECKER 1
#define NFB_SPLIT_CHANNEL 2

#define NFB_OK 0
#define NFB_ERR_NULL (-1)        /* required pointer is NULL */
#define NFB_ERR_SHAPE (-2)       /* non-positive or inconsistent dimension */
#define NFB_ERR_SPLIT (-3)       /* shape not splittable in this mode (odd H/W or odd C): squeeze.py view/split would fail */
#define NFB_ERR_UNSUPPORTED (-4) /* valid request outside what the kernels implement (e.g. K > NFB_MAX_MIXTURES) */

#define NFB_MAX_MIXTURES 32
#define NFB_MAX_BINS 32

int nfb_version(void);
const char* nfb_error_string(int code);
/* number of kernel launches enqueued by this library since load (all entry points); for bench accounting */
unsigned long long nfb_launch_count(void);

/* ---------------- coupling bijections (coupling.py) ---------------- */

/* AffineCoupling._transform (coupling.py:104-112): s = tanh(s_raw)*a + b; z0' = z0*exp(s) + t; ldj += sum(s).
 * params = (B, 2*c0, h, w): channels [0,c0) = t, [c0,2c0) = s_raw.  s_log_scale / s_bias: device scalars. */
int nfb_affine_coupling_fwd(const float* z_in, float* z_out, const float* params, const float* ldj_in,
                            float* ldj_out, const float* s_log_scale, const float* s_bias, int B, int C, int H,
                            int W, int mode, int odd, nfb_stream_t stream);
/* AffineCoupling._inverse_transform (coupling.py:114-122): y0' = exp(-s)*(y0 - t); ldj -= sum(s). */
int nfb_affine_coupling_inv(const float* z_in, float* z_out, const float* params, const float* ldj_in,
                            float* ldj_out, const float* s_log_scale, const float* s_bias, int B, int C, int H,
                            int W, int mode, int odd, nfb_stream_t stream);

/* AdditiveCoupling (coupling.py:69-79): z0' = z0 + sign*t; params = (B, c0, h, w) = t.  No log-det. */
int nfb_additive_coupling(const float* z_in, float* z_out, const float* params, float sign, int B, int C, int H,
                          int W, int mode, int odd, nfb_stream_t stream);

/* MixLogAttnCoupling._transform (coupling.py:172-190) incl. MixLogCDF.forward (modules.py:190-194) and
 * Logit(1e-5).forward: params = (B, (2+3K)*c0, h, w), sections [a | b | logpi(K) | mu(K) | s(K)],
 * mixture channel = k*c0 + m (coupling.py:180-182). */
int nfb_mixlog_coupling_fwd(const float* z_in, float* z_out, const float* params, const float* ldj_in,
                            float* ldj_out, const float* a_log_scale, const float* a_bias, int B, int C, int H,
                            int W, int mode, int odd, int K, nfb_stream_t stream);
/* MixLogAttnCoupling._inverse_transform (coupling.py:192-210) incl. the bisection of MixLogCDF.backward
 * (modules.py:196-212).  The reference stops after 25 iterations when every element has converged and runs
 * all 100 when any element stalls (val == x); `stall_flag` (device int, zeroed by the call) reproduces that
 * global rule without a host sync: phase 1 runs 25 iterations and raises the flag on a stall, phase 2 (same
 * call) continues to 100 iterations only if the flag is set.  scratch: 2*B*n0 floats (lo/hi). */
int nfb_mixlog_coupling_inv(const float* z_in, float* z_out, const float* params, const float* ldj_in,
                            float* ldj_out, const float* a_log_scale, const float* a_bias, float* scratch,
                            int* stall_flag, int B, int C, int H, int W, int mode, int odd, int K,
                            nfb_stream_t stream);

/* MixLogCDF.forward (modules.py:190-194) as a standalone layer, with mix_logistic_logpdf / mix_logistic_logcdf and the
 * logistic_* helpers (modules.py:64-97) inside: x (B, n), log_pi / mu / s (B, K, n) as given (no normalisation);
 * y = exp(mix_logcdf(x)), ldj_out[b] = ldj_in[b] + sum_n mix_logpdf(x). */
int nfb_mixlogcdf_fwd(const float* x, float* y, const float* log_pi, const float* mu, const float* s,
                      const float* ldj_in, float* ldj_out, int B, int n, int K, nfb_stream_t stream);
/* MixLogCDF.backward (modules.py:196-212): bisection on [-1e3, 1e3], the reference's global 25 / 100 iteration rule as in
 * nfb_mixlog_coupling_inv (scratch: 2*B*n floats, stall_flag: device int); ldj_out[b] = ldj_in[b] - sum_n mix_logpdf(x). */
int nfb_mixlogcdf_inv(const float* y, float* x, const float* log_pi, const float* mu, const float* s, const float* ldj_in,
                      float* ldj_out, float* scratch, int* stall_flag, int B, int n, int K, nfb_stream_t stream);

/* Rational-quadratic spline coupling (Durkan et al. 2019; no counterpart in the reference -- parity unpinned):
 * params = (B, (3K-1)*c0, h, w), sections [widths(K) | heights(K) | derivatives(K-1)], bin channel = k*c0 + m;
 * identity outside [-bound, bound], boundary derivatives 1. */
int nfb_rqs_coupling_fwd(const float* z_in, float* z_out, const float* params, const float* ldj_in,
                         float* ldj_out, int B, int C, int H, int W, int mode, int odd, int K, float bound,
                         nfb_stream_t stream);
int nfb_rqs_coupling_inv(const float* z_in, float* z_out, const float* params, const float* ldj_in,
                         float* ldj_out, int B, int C, int H, int W, int mode, int odd, int K, float bound,
                         nfb_stream_t stream);

/* squeeze (split) / unsqueeze (merge) of AbstractCoupling (coupling.py:33,35): bit-exact permutations.
 * z0_out / z1_out are (B, c0, h, w) contiguous; either may be NULL to skip it (merge: a NULL half reads as zeros --
 * the gradient of a split of which only one half was used). */
int nfb_coupling_split(const float* z, float* z0_out, float* z1_out, int B, int C, int H, int W, int mode,
                       int odd, nfb_stream_t stream);
int nfb_coupling_merge(const float* z0, const float* z1, float* z_out, int B, int C, int H, int W, int mode,
                       int odd, nfb_stream_t stream);

/* ---------------- per-channel affine layers (modules.py) ---------------- */

/* ActNorm.forward (modules.py:246-250): z = (z - bias)/exp(log_scale); ldj -= sum(log_scale)*HW. */
int nfb_actnorm_fwd(const float* z_in, float* z_out, const float* ldj_in, float* ldj_out, const float* log_scale,
                    const float* bias, int B, int C, int HW, nfb_stream_t stream);
/* ActNorm.backward (modules.py:252-256): y = y*exp(log_scale) + bias; ldj += sum(log_scale)*HW. */
int nfb_actnorm_inv(const float* z_in, float* z_out, const float* ldj_in, float* ldj_out, const float* log_scale,
                    const float* bias, int B, int C, int HW, nfb_stream_t stream);
/* ActNorm data-dependent init (modules.py:238-244): log_scale = log(std_unbiased + eps), bias = mean. */
int nfb_actnorm_init(const float* z, float* log_scale_out, float* bias_out, int B, int C, int HW, float eps,
                     nfb_stream_t stream);
/* flow BatchNorm.forward (modules.py:300-305): x = (x-mean)/sqrt(var)*exp(log_gamma)+beta;
 * ldj += sum(log_gamma - 0.5 log var)*HW. */
int nfb_bnflow_fwd(const float* z_in, float* z_out, const float* ldj_in, float* ldj_out, const float* mean,
                   const float* var, const float* log_gamma, const float* beta, int B, int C, int HW,
                   nfb_stream_t stream);
/* flow BatchNorm.backward (modules.py:315-320). */
int nfb_bnflow_inv(const float* z_in, float* z_out, const float* ldj_in, float* ldj_out, const float* mean,
                   const float* var, const float* log_gamma, const float* beta, int B, int C, int HW,
                   nfb_stream_t stream);
/* flow BatchNorm train-mode statistics (modules.py:285-287): mean, biased variance + eps. */
int nfb_bnflow_batch_stats(const float* z, float* mean_out, float* var_out, int B, int C, int HW, float eps,
                           nfb_stream_t stream);

/* Per-channel raw moments in fp64 over (batch, pixels): moments_out[c] = sum x, moments_out[C + c] = sum x^2 (device
 * double[2C]).  Additive across ranks: the sharded form of the statistics behind ActNorm's init (modules.py:238-244) and
 * the train-mode BatchNorm (modules.py:285-287); see nfb200.parallel. */
int nfb_channel_moments(const float* z, double* moments_out, int B, int C, int HW, nfb_stream_t stream);

/* Logit.forward (modules.py:146-150): x = clamp(x, lo, hi); y = logit(x); ldj += sum(-(y - 2 softplus(y))). */
int nfb_logit_fwd(const float* z_in, float* z_out, const float* ldj_in, float* ldj_out, float lo, float hi, int B,
                  int D, nfb_stream_t stream);
/* Logit.backward (modules.py:152-155): sigmoid; ldj += sum(x - 2 softplus(x)). */
int nfb_logit_inv(const float* z_in, float* z_out, const float* ldj_in, float* ldj_out, int B, int D,
                  nfb_stream_t stream);

/* ---------------- invertible 1x1 convolution (modules.py:441-497) ---------------- */

/* W = P (L o tril + I)(U o triu + diag(sign_s exp(log_s)))  (modules.py:471-473) -> W_out (C*C, row-major);
 * if Winv_out != NULL also W^-1 = U'^-1 L'^-1 P^T (fp64 substitution on the device, rounded once) which
 * replaces the per-call lu_solve of modules.py:490. */
int nfb_invconv1x1_weight(const float* P, const float* L, const float* U, const float* log_s,
                          const float* sign_s, float* W_out, float* Winv_out, int C, nfb_stream_t stream);
/* out[b,:,p] = M @ z[b,:,p]; ldj += sign * sum(log_s) * HW.  forward: M = W, sign = +1 (modules.py:477-480);
 * inverse: M = W^-1, sign = -1 (modules.py:490-495).  z_out must NOT alias z_in.  ldj_out == NULL: matrix apply only. */
int nfb_invconv1x1_apply(const float* z_in, float* z_out, const float* ldj_in, float* ldj_out, const float* M,
                         const float* log_s, float sign, int B, int C, int HW, nfb_stream_t stream);

/* Fused flow-step front end: ActNorm.forward (modules.py:246-250) followed by InvertibleConv1x1.forward
 * (modules.py:470-482) in one pass over z -- bit-identical to the two calls back to back.  Returns
 * NFB_ERR_UNSUPPORTED for shapes the tiled kernel does not take (HW % 4 != 0): run the layers separately. */
int nfb_actnorm_invconv_fwd(const float* z_in, float* z_out, const float* ldj_in, float* ldj_out,
                            const float* log_scale, const float* bias, const float* W, const float* log_s, int B, int C,
                            int HW, nfb_stream_t stream);

/* Inverse of that pair in one pass: InvertibleConv1x1.backward (modules.py:484-497, Winv from nfb_invconv1x1_weight)
 * followed by ActNorm.backward (modules.py:252-256) -- bit-identical to the two calls back to back. */
int nfb_invconv_actnorm_inv(const float* y_in, float* y_out, const float* ldj_in, float* ldj_out, const float* Winv,
                            const float* log_s, const float* log_scale, const float* bias, int B, int C, int HW,
                            nfb_stream_t stream);

/* ---------------- Squeeze2d / Unsqueeze2d (squeeze.py:153-189) ---------------- */

/* (B,C,H,W) -> (B,4C,H/2,W/2): out[b,4c+2dy+dx,i,j] = in[b,c,2i+dy,2j+dx]; odd swaps the two channel halves. */
int nfb_squeeze2d(const float* z_in, float* z_out, int B, int C, int H, int W, int odd, nfb_stream_t stream);
/* exact inverse; C, H, W are those of the UNSQUEEZED (output) tensor. */
int nfb_unsqueeze2d(const float* z_in, float* z_out, int B, int C, int H, int W, int odd, nfb_stream_t stream);

/* ---------------- likelihood (main.py:83-85) ---------------- */

/* nll_rows[b] = 0.5*||z_b||^2 + 0.5*D*log(2 pi) - ldj[b]  (= -(log N(z_b;0,I) + ldj_b)), ||z||^2 accumulated in
 * fp64 and rounded once.  nll_rows (device float[B]) is required.  sum_out (device double[2], may be NULL) =
 * { sum_b nll_rows[b] in fp64 (fixed order: deterministic), B } -- the 2-element payload that is all-reduced
 * across GPUs. */
int nfb_gauss_nll(const float* z, const float* ldj, float* nll_rows, double* sum_out, int B, int D,
                  nfb_stream_t stream);

/* ---------------- conditioner networks (modules.py:342-438, weight_norm.py) ---------------- */

/* Fold WeightNorm (weight_norm.py:40: w = v*g/(||v||_dim0 + eps)) of one conv/linear weight:
 * v (O, I*kk), g (I*kk) -> w_out (O, I*kk). */
int nfb_weight_norm(const float* v, const float* g, float* w_out, int O, int Ikk, float eps, nfb_stream_t stream);

/* ConvNet (modules.py:416-438) / MLP (modules.py:391-413) in eval mode as ONE kernel.
 * nfb_resnet_pack folds WeightNorm, and every BatchNorm that directly follows a conv/linear, into one packed weight
 * buffer (nfb_resnet_pack_size floats; call again whenever a parameter changes).  `tensors` is a HOST array of 38
 * device pointers: (weight_v, weight_g, bias) of in_block.0, mid_block.0.net.2, mid_block.0.net.5, mid_block.1.net.2,
 * mid_block.1.net.5, out_block.2, then (weight, bias, running_mean, running_var) of mid_block.0.net.0,
 * mid_block.0.net.3, mid_block.1.net.0, mid_block.1.net.3, out_block.0.  conv != 0: ConvNet (3x3 / 1x1), else MLP. */
int nfb_resnet_pack_size(int in_ch, int out_ch, int conv);
int nfb_resnet_pack(const float* const* tensors, float* packed, int in_ch, int out_ch, int conv, float wn_eps,
                    float bn_eps, nfb_stream_t stream);
/* params_out (B, out_ch, h, w) = ConvNet(z1).  mode = NFB_SPLIT_CHECKER / NFB_SPLIT_CHANNEL: `src` is the coupling's
 * z (B, C, H, W) and the conditioner input z1 = the pass-through half is gathered on the fly (coupling.py:33,105);
 * mode < 0: `src` is the (B, in_ch, H, W) conditioner input itself (C ignored).  Supported spatial sizes of the
 * conditioner input: 32x32 (one 2-CTA cluster per sample), 16x16, 8x8, 4x4 (else NFB_ERR_UNSUPPORTED). */
int nfb_convnet_fwd(const float* src, float* params_out, const float* packed, int B, int C, int H, int W, int mode,
                    int odd, int in_ch, int out_ch, nfb_stream_t stream);
/* Same with an explicit kernel selection (the library keeps no mutable state).  flags = 0 is nfb_convnet_fwd: the
 * tensor-core kernel (tcgen05, error-compensated half-precision split: x = hi + 2^-10 lo, both fp16 -- the error of an
 * fp32 convolution) where it applies (16x16 / 8x8 / 4x4), else the FP32-FFMA kernels.  Domain of the default operand format:
 * |conditioner input|, |activation| and |folded weight| < 65504; beyond it the affected samples come back as NaN (never
 * a finite wrong number) -- use NFB_CONV_TF32 or NFB_CONV_FFMA for such data.
 *   NFB_CONV_FFMA            use the FP32-FFMA kernels
 *   NFB_CONV_TF32            tensor-core kernel with 3xTF32 operands (full fp32 range, ~1.25x the time at 16x16)
 *   NFB_CONV_SINGLE          16x16 maps: one sample per CTA at a time (default: two in flight per CTA when the batch gives
 *                            every SM more than one sample, or with NFB_CONV_PAIR)
 *   NFB_CONV_VARIANT_MASK    thread-tile variant of the FP32-FFMA kernel for this spatial size (0 = default)
 *   NFB_CONV_GROUPS(g)       accumulator groups per layer of the tensor-core kernel, 1..4 (0 = default 3; at most the
 *                            k-steps of one tap: 2 with the default operands, 4 with NFB_CONV_TF32)
 *   NFB_CONV_PAIR            8x8 / 4x4 maps: two independent 128-position tiles per CTA taking turns on the tensor core
 *                            (more work per SM-second, half the CTAs): for several batches in flight; without the flag
 *                            it is chosen only when the batch alone gives every SM two tiles
 *   NFB_CONV_DEBUG(bits)     profiling knobs of the tensor-core kernel (skip MMAs / TMEM loads / ...): WRONG results; honoured
 *                            only by a library built with -DNFB_TC_DEBUG_KNOBS, ignored otherwise */
#define NFB_CONV_VARIANT_MASK 0x7
#define NFB_CONV_FFMA 0x8
#define NFB_CONV_TF32 0x10000
#define NFB_CONV_ITERS_SHIFT 18 /* with NFB_CONV_PAIR: (1 + value) loop iterations per CTA (fewer, longer CTAs), value 0..3;
                                  measured on Glow-32 with 5 and 10 batches in flight: slower (75.0 k -> 62.9 k / 67.9 k samples/s) */
#define NFB_CONV_ITERS(n) ((n) << NFB_CONV_ITERS_SHIFT)
#define NFB_CONV_SINGLE 0x20000
#define NFB_CONV_PAIR 0x80
#define NFB_CONV_GROUPS_SHIFT 4
#define NFB_CONV_GROUPS(g) ((g) << NFB_CONV_GROUPS_SHIFT)
#define NFB_CONV_DEBUG_SHIFT 8
#define NFB_CONV_DEBUG(bits) ((bits) << NFB_CONV_DEBUG_SHIFT)
int nfb_convnet_fwd_ex(const float* src, float* params_out, const float* packed, int B, int C, int H, int W, int mode,
                       int odd, int in_ch, int out_ch, int flags, nfb_stream_t stream);
/* AffineCoupling.forward with its ConvNet conditioner as ONE kernel, in place on z (coupling.py:32-36,104-112 +
 * modules.py:416-438): z1 = pass-through half of z is gathered by the kernel, params = ConvNet(z1) is computed on the
 * tensor cores (tcgen05, error-compensated FP16 split; NFB_CONV_TF32: 3xTF32) and consumed from tensor memory -- it
 * never reaches shared or global memory -- z0 <- z0*exp(tanh(s_raw)*a + b) + t is written over z0 inside z, ldj[b] += sum(s).  z1 stays untouched, so
 * the result equals the reference's merged output.  `packed` from nfb_resnet_pack(in_ch = c0, out_ch = 2*c0).
 * Conditioner input sizes 16x16 / 8x8 / 4x4; otherwise (or with NFB_CONV_FFMA in flags) NFB_ERR_UNSUPPORTED: run
 * nfb_convnet_fwd + nfb_affine_coupling_fwd.  flags as for nfb_convnet_fwd_ex. */
int nfb_convnet_affine_fwd(float* z, float* ldj, const float* packed, const float* s_log_scale, const float* s_bias,
                           int B, int C, int H, int W, int mode, int odd, int flags, nfb_stream_t stream);
/* nfb_convnet_affine_fwd followed, inside the same kernel, by the NEXT flow step's ActNorm.forward (modules.py:246-250) and
 * InvertibleConv1x1.forward (modules.py:470-480) on the samples the CTA has just finished: z <- W ((z' - bias) / exp(log_scale)),
 * ldj += (sum log_s - sum log_scale) * H*W, all in place -- a Glow flow step is then ONE launch (glow.py:27-29: this
 * step's coupling + the next step's ActNorm and 1x1 conv).  next_W: the matrix from nfb_invconv1x1_weight.  Implemented
 * where it beats a separate launch -- 16x16 conditioner maps with C = 3 or 12, 8x8 maps with C = 12 (small per-pixel
 * matrices); otherwise NFB_ERR_UNSUPPORTED: run nfb_convnet_affine_fwd and nfb_actnorm_invconv_fwd. */
int nfb_convnet_affine_step_fwd(float* z, float* ldj, const float* packed, const float* s_log_scale, const float* s_bias,
                                const float* next_log_scale, const float* next_bias, const float* next_W,
                                const float* next_log_s, int B, int C, int H, int W, int mode, int odd, int flags,
                                nfb_stream_t stream);

/* Flow++ conditioner (coupling.py:160-167: Conv2d(in,32,3) -> GatedConv2d -> LayerNorm -> GatedAttn(4 heads) ->
 * LayerNorm -> Conv2d(32,out,3); modules.py:519-578) as ONE kernel.  `tensors`: HOST array of 15 device pointers:
 * net.0 weight packed by nfb_pack_conv3x3, net.0.bias, net.1.op weight packed, net.1.op.bias, net.2.weight, net.2.bias,
 * net.3.pos_emb, net.3.conv1.weight (96,32), net.3.conv1.bias, net.3.conv2.weight (64,32), net.3.conv2.bias,
 * net.4.weight, net.4.bias, net.5 weight packed, net.5.bias.  src / mode / spatial sizes as for nfb_convnet_fwd. */
int nfb_pack_conv3x3(const float* w, float* out, int O, int I, nfb_stream_t stream); /* (O,I,3,3) -> [O/32][I][9][32] */
int nfb_flowpp_cond_fwd(const float* const* tensors, const float* src, float* params_out, int B, int C, int H, int W,
                        int mode, int odd, int in_ch, int out_ch, nfb_stream_t stream);

/* Flow++ conditioner of the 1-D couplings (coupling.py:142-158: Linear -> GatedLinear -> LayerNorm -> GatedAttn -> LayerNorm
 * -> Linear; a 1-D input is a single attention token, so A = Q exactly) as ONE kernel.  `tensors`: HOST array of the same
 * 15 device pointers as nfb_flowpp_cond_fwd, all UNPACKED (net.0.weight (32,in), net.1.op.weight (32,64), net.5.weight
 * (out,32)).  mode = NFB_SPLIT_1D: src = z (B, D), z1 gathered; mode < 0: src = (B, in_ch).  NFB_ERR_UNSUPPORTED when the
 * network does not fit in shared memory. */
int nfb_flowpp_mlp_fwd(const float* const* tensors, const float* src, float* params_out, int B, int D, int mode, int odd,
                       int in_ch, int out_ch, nfb_stream_t stream);

/* params_out (B, out_ch) = MLP(z1).  mode = NFB_SPLIT_1D: src = z (B, C), z1 gathered (squeeze.py:64-72);
 * mode < 0: src = (B, in_ch). */
int nfb_mlp_fwd(const float* src, float* params_out, const float* packed, int B, int C, int mode, int odd, int in_ch,
                int out_ch, nfb_stream_t stream);

/* ---------------- gradients of the bijective layers (the reference gets them from autograd) -------
 * Convention: gy = dLoss/d(z_out) (layout of z), gldj = dLoss/d(ldj_out) (device float[B], may be NULL = zeros).  Every
 * layer passes gldj through unchanged to its ldj input, so only gz and the parameter gradients are produced.  Each
 * backward recomputes tanh/exp from the layer's saved INPUT.  `scratch`: device doubles (sizes below), zeroed by the
 * call; batch-wide sums are accumulated in fp64 and rounded to fp32 once. */

/* AffineCoupling.forward (coupling.py:104-112): gz (z0 slots: gy*exp(s), z1 slots: gy -- the conditioner's own
 * gradient wrt z1 is added by the caller), gparams (B, 2*c0, h, w) = [gy0 | gs*a*(1-tanh^2)], gs = gy0*z0*exp(s)+gldj,
 * g_s_log_scale = sum gs*tanh(s_raw), g_s_bias = sum gs (device scalars, may be NULL).  scratch: 2 doubles. */
int nfb_affine_coupling_bwd(const float* z_in, const float* params, const float* gy, const float* gldj, float* gz,
                            float* gparams, float* g_s_log_scale, float* g_s_bias, double* scratch,
                            const float* s_log_scale, const float* s_bias, int B, int C, int H, int W, int mode, int odd,
                            nfb_stream_t stream);
/* MixLogAttnCoupling.forward (coupling.py:172-190): gparams (B, (2+3K)*c0, h, w) in the layout of params; the mixture
 * responsibilities are evaluated in the log domain (no division by an underflowed density).  scratch: 2 doubles. */
int nfb_mixlog_coupling_bwd(const float* z_in, const float* params, const float* gy, const float* gldj, float* gz,
                            float* gparams, float* g_a_log_scale, float* g_a_bias, double* scratch,
                            const float* a_log_scale, const float* a_bias, int B, int C, int H, int W, int mode, int odd,
                            int K, nfb_stream_t stream);
/* RQ-spline coupling, forward direction (nfb_rqs_coupling_fwd): gparams (B, (3K-1)*c0, h, w). */
int nfb_rqs_coupling_bwd(const float* z_in, const float* params, const float* gy, const float* gldj, float* gz,
                         float* gparams, int B, int C, int H, int W, int mode, int odd, int K, float bound,
                         nfb_stream_t stream);
/* ActNorm.forward (modules.py:246-250): gz = gy/exp(log_scale); g_bias[c] = -sum gz; g_log_scale[c] = -sum gy*y -
 * HW*sum_b gldj[b].  scratch: 2C doubles. */
int nfb_actnorm_bwd(const float* gy, const float* z_in, const float* gldj, const float* log_scale, const float* bias,
                    float* gz, float* g_log_scale, float* g_bias, double* scratch, int B, int C, int HW,
                    nfb_stream_t stream);
/* flow BatchNorm.forward (modules.py:300-305) with mean/var treated as constants -- the reference copies the batch
 * statistics into buffers through .data (modules.py:288-289), so autograd never sees them.  gx = gy*exp(log_gamma)/
 * sqrt(var); g_log_gamma[c] = sum gy*xhat*exp(log_gamma) + HW*sum_b gldj[b]; g_beta[c] = sum gy (either may be NULL:
 * affine=False).  scratch: 2C doubles. */
int nfb_bnflow_bwd(const float* gy, const float* x_in, const float* gldj, const float* mean, const float* var,
                   const float* log_gamma, float* gx, float* g_log_gamma, float* g_beta, double* scratch, int B, int C,
                   int HW, nfb_stream_t stream);
/* InvertibleConv1x1.forward (modules.py:470-482).  gz = W^T gy is nfb_invconv1x1_apply(gy, gz, NULL, NULL, W^T, ...);
 * gW[i,j] = sum_{b,p} gy[b,i,p] z[b,j,p] (C*C floats).  Position chunks store partial tiles in `scratch`
 * (nfb_invconv1x1_wgrad_scratch floats, need not be zeroed) and a second kernel adds them: no atomics, deterministic. */
long long nfb_invconv1x1_wgrad_scratch(int B, int C, int HW);
int nfb_invconv1x1_wgrad(const float* gy, const float* z_in, float* gW, float* scratch, int B, int C, int HW,
                         nfb_stream_t stream);
/* chain rule through W = P (L o tril + I)(U o triu + diag(sign_s exp(log_s))) (modules.py:471-473) plus the log-det
 * term: gL, gU (C*C, zero outside the strict triangles), g_log_s (C). */
int nfb_invconv1x1_weight_bwd(const float* gW, const float* P, const float* L, const float* U, const float* log_s,
                              const float* sign_s, const float* gldj, float* gL, float* gU, float* g_log_s, int B, int C,
                              int HW, nfb_stream_t stream);
/* Logit.forward (modules.py:146-150): gx = gy/(x(1-x)) + gldj[b]*(1/(1-x) - 1/x) inside [lo, hi], 0 outside (clamp). */
int nfb_logit_bwd(const float* x_in, const float* gy, const float* gldj, float* gx, float lo, float hi, int B, int D,
                  nfb_stream_t stream);
/* nfb_gauss_nll rows: gz = z * g_rows[b], gldj = -g_rows[b] (gldj may be NULL). */
int nfb_gauss_nll_bwd(const float* z, const float* g_rows, float* gz, float* gldj, int B, int D, nfb_stream_t stream);

/* ---------------- train-mode ConvNet conditioner (modules.py:416-438 under net.train()) and its backward ----------------
 * BatchNorm on batch statistics makes every layer depend on the whole batch, so the network runs layer by layer over
 * (B, 32, h, w) activations (h x w in {32x32, 16x16, 8x8, 4x4}; else NFB_ERR_UNSUPPORTED and the caller uses the library path).
 * nfb200/flows/conditioner_train.py strings these together as one autograd function. */

/* WeightNorm (weight_norm.py:40) of v (O, I, KK) / g (I, KK), KK in {1, 9}: w_nat (O, I, KK) plus the two packed layouts
 * the convolution kernel reads -- w_fwd [O/32][I/32][32 ci][KK][32 o] and, for the data gradient, w_bwd
 * [I/32][O/32][32 o][KK flipped][32 i].  Both packed buffers hold ceil(O/32)*ceil(I/32)*32*KK*32 floats; the padding
 * is written (as zeros) by the call. */
int nfb_wn_pack_train(const float* v, const float* g, float* w_nat, float* w_fwd, float* w_bwd, int O, int I, int KK,
                      float eps, nfb_stream_t stream);
/* gradient of that map: gw (O, I*KK) -> gv (O, I*KK), gg (I*KK). */
int nfb_wn_bwd(const float* v, const float* g, const float* gw, float* gv, float* gg, int O, int Ikk, float eps,
               nfb_stream_t stream);
/* The same two maps for all n (<= 8) WeightNorm layers of one conditioner in ONE launch.  ptrs: HOST array of 5 device
 * pointers per layer -- pack: (v, g, w_nat, w_fwd, w_bwd); backward: (v, g, gw, gv, gg); dims: HOST array of (O, I, KK)
 * per layer. */
int nfb_wn_pack_train_multi(const void* const* ptrs, const int* dims, int n, float eps, nfb_stream_t stream);
int nfb_wn_bwd_multi(const void* const* ptrs, const int* dims, int n, float eps, nfb_stream_t stream);
/* out (B, Cout, h, w) = conv_ks(in (B, Cin, h, w); packed w) + bias (+ skip); ks in {1, 3}, padding ks/2.  stats (device
 * double[2*Cout], may be NULL): per-channel sum and sum of squares of `out` for the BatchNorm that follows, accumulated
 * with fp64 atomics; zeroed by the call unless stats_zeroed != 0 (the caller carved it from one zero-filled arena).  The
 * data gradient of a layer is the same call with w_bwd and Cin / Cout exchanged. */
int nfb_conv_train(const float* in, const float* w_packed, const float* bias, const float* skip, float* out, double* stats,
                   int stats_zeroed, int B, int Cin, int Cout, int h, int w, int ks, nfb_stream_t stream);
/* Data gradient of a layer fused with the BatchNorm+ReLU backward reduction of the layer below it:
 * U (B, Cout, h, w) = conv_ks(gy (B, Cin, h, w); w_bwd) * [a > 0], sums = (sum U | sum U * xhat) -- what
 * nfb_conv_train + nfb_bn_relu_bwd_reduce compute in two passes.  a / x: that BatchNorm's output (after ReLU) / input. */
int nfb_conv_train_dgrad_bnrelu(const float* gy, const float* w_bwd, const float* a, const float* x, const float* mean_rstd,
                                float* U, double* sums, int sums_zeroed, int B, int Cin, int Cout, int h, int w, int ks,
                                nfb_stream_t stream);
/* gw (Cout, Cin, ks, ks) = sum over batch and pixels of gy (B, Cout, h, w) x a (B, Cin, h, w); gb (Cout, may be NULL) =
 * sum of gy.  Sample groups write partial sums into `scratch` (nfb_conv_train_wgrad_scratch floats, need not be zeroed)
 * and a second kernel adds them up: no atomics, deterministic. */
long long nfb_conv_train_wgrad_scratch(int B, int Cin, int Cout, int h, int w, int ks);
int nfb_conv_train_wgrad(const float* gy, const float* a, float* gw, float* gb, float* scratch, int B, int Cin, int Cout,
                         int h, int w, int ks, nfb_stream_t stream);
/* a = relu(BatchNorm_train(x)) from the moments `stats` (sum | sum of squares over batch and pixels); stores
 * mean_rstd (float[2C]) for the backward pass and updates running_mean / running_var (may be NULL) with `momentum`
 * (running_var takes the unbiased variance, like nn.BatchNorm2d). */
int nfb_bn_relu_fwd(const float* x, const double* stats, const float* gamma, const float* beta, float* running_mean,
                    float* running_var, float momentum, float eps, float* a_out, float* mean_rstd, int B, int C, int HW,
                    nfb_stream_t stream);
/* U = ga * [a > 0]; sums (device double[2C], zeroed by the call unless sums_zeroed != 0) = (sum U | sum U * xhat). */
int nfb_bn_relu_bwd_reduce(const float* ga, const float* a, const float* x, const float* mean_rstd, float* U, double* sums,
                           int sums_zeroed, int B, int C, int HW, nfb_stream_t stream);
/* gx = gamma * rstd * (U - mean(U) - xhat * mean(U xhat)) (+ add, may be NULL); g_gamma = sum U xhat, g_beta = sum U. */
int nfb_bn_bwd_apply(const float* U, const float* x, const float* mean_rstd, const float* gamma, const double* sums,
                     const float* add, float* gx, float* g_gamma, float* g_beta, int B, int C, int HW, nfb_stream_t stream);

/* (B, C) rows <-> (B/HW, C, HW) channel planes (B % HW == 0): the train-mode MLP conditioner (modules.py:391-413) runs on
 * the convolution-layer kernels above as six 1x1 convolutions over planes of HW rows each. */
int nfb_rows_to_planes(const float* rows, float* planes, int B, int C, int HW, nfb_stream_t stream);
int nfb_planes_to_rows(const float* planes, float* rows, int B, int C, int HW, nfb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* NFB200_H_ */
