/*
 * e3b.h -- C ABI of libe3b.so: the B200-native (sm_100a) kernels behind the elektronn3 UNet /
 * Predictor hot path.
 *
 * The reference has no FFI of its own: its seam is the torch.nn.Module protocol, and the arithmetic
 * is delegated to torch ATen/cuDNN at the call sites cited on each entry point below (paths relative
 * to the reference root).  This library replaces those library calls.  It is bound from Python with
 * ctypes (elektronn3_b200/_lib.py); INTEGRATION.md shows the binding a reference maintainer would add.
 *
 * Conventions
 *  - plain C, no torch types: raw device pointers, int sizes, a cudaStream_t passed as void*.
 *  - no allocation, no ownership transfer: the caller (PyTorch's allocator) owns every buffer.
 *  - every function returns 0 on success, non-zero on error; e3b_last_error() returns the message of
 *    the calling thread's last error.  Nothing throws, nothing calls exit().
 *  - all functions only ENQUEUE work on `stream` (asynchronous w.r.t. the host) and are re-entrant.
 *  - internal fp32 layout "QP" (quad-planar): float32 (N, Cq, D, H, W, 4), Cq = ceil8(C)/4,
 *    channel c -> plane c/4, lane c%4, padding channels exactly 0.  2D data uses D = 1.
 *    Carries convolution outputs and gradients w.r.t. activations (what the elementwise kernels read).
 *  - operand layout "QH": float16 (N, Ch, D, H, W, 8), Ch = ceil16(C)/8, channel c -> plane c/8, lane c%8,
 *    padding channels exactly 0.  Carries everything a forward / dgrad MMA reads: the network input,
 *    activations, pooled activations and conv-output gradients.  fp16 has TF32's 10 explicit mantissa
 *    bits (the reference's GPU arithmetic); gradients are stored times a per-tensor power of two.
 */
#ifndef E3B_H
#define E3B_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define E3B_VERSION 230

int e3b_version(void);
const char* e3b_last_error(void);
/* number of kernels launched by this library since load (all threads); bench.py's gpu_launches */
int64_t e3b_launch_count(void);

/* ---- layout conversion at the module boundary ------------------------------------------------
 * NCDHW float32 -> QH (e3b_pack_ncdhw, e3b_gather_tiles) and QP -> NCDHW float32 (e3b_unpack_qp).  Used for the network input (reference: trainer.py:515 `inp.to(device)`)
 * and by tests.  src may be a sub-box of a larger volume (Predictor tiles, inference.py:179-189):
 * (Dv,Hv,Wv) are the extents of the allocation, (z0,y0,x0) the origin of the box inside it (may be
 * negative / overhanging: out-of-volume voxels read as 0, which is tiled_apply's zero padding). */
int e3b_pack_ncdhw(const float* src, void* dst_qh, int N, int C, int D, int H, int W,
                   int Dv, int Hv, int Wv, int z0, int y0, int x0, void* stream);
int e3b_unpack_qp(const float* src_qp, float* dst, int N, int C, int D, int H, int W, void* stream);

/* Predictor tile gather (inference.py:179-189): tile b of the batch is the box of a single-sample
 * volume (Cv, Dv, Hv, Wv) whose origin is origins[3*b..3*b+2] (device int32, may overhang -> 0).
 * flip (bit 0: D, bit 1: H, bit 2: W) mirrors the tile in those dims while gathering: the test-time
 * augmentation FlipAugment.forward (inference.py:215-223) as an index transform instead of a torch.flip copy. */
int e3b_gather_tiles(const float* vol, const int32_t* origins, void* dst_qh, int B, int C,
                     int D, int H, int W, int Dv, int Hv, int Wv, int flip, void* stream);

/* ---- weight packing ---------------------------------------------------------------------------
 * torch parameter layouts -> the fp16 K-major no-swizzle shared-memory image the conv kernel streams.
 *   mode 0: conv forward.  w (Co, C0+C1, kd,kh,kw); K space = [pad16(C0) | pad16(C1)], N space = pad16(Co)
 *   mode 1: conv dgrad.    K space = pad16(Co); N space = [pad8(C0) | pad8(C1)] padded to 16; taps flipped
 *   mode 2: transposed conv k=s forward. w (Ci, Co, sd,sh,sw); K = pad16(Ci), N = taps * pad16(Co)
 *   mode 3: transposed conv dgrad on the space-to-depth gradient: K = pad16(taps * pad8(Co)), N = pad16(Ci)
 *   mode 4 / 5: the contents of mode 0 / 1 as the image of the z-stacked kernel (e3b_conv_args.variant = 1):
 *           [chunk16][(dy,dx) tap][k half][3 z taps x N][8], 3x3x3 taps and N <= 80 only
 * `scale` (optional, [Co]) multiplies output channel co (eval-mode BatchNorm folding, mode 0 only).
 * `wscale` (optional, device scalar): a power of two multiplied into every weight before the fp16 rounding, chosen by
 * the caller so that max|w| lands in [1, 2): tiny (or huge) weights then keep TF32's 10 mantissa bits instead of falling
 * into fp16's subnormal range; the convolution undoes it through e3b_conv_args.w_unscale.
 * e3b_packed_weight_floats() returns the size of `dst` in units of 4 bytes. */
int64_t e3b_packed_weight_floats(int mode, int C0, int C1, int Co, int kd, int kh, int kw);
int e3b_pack_weights(int mode, const float* w, const float* scale, const float* wscale, void* dst, int C0, int C1, int Co,
                     int kd, int kh, int kw, void* stream);

/* All weight images of a network in ONE launch (a training step re-packs every image).  The caller describes the images
 * once (pointers are those of the parameters and of persistent destination buffers), e3b_pack_jobs_fill() turns the
 * descriptions into the library's device-table format in a HOST buffer of e3b_pack_job_table_bytes(njobs) bytes, the caller
 * copies that buffer to the device (once) and calls e3b_pack_weights_batched() with the device copy every step. */
typedef struct e3b_pack_job {
    const float* w; const float* scale; const float* wscale; void* dst;    /* as e3b_pack_weights */
    int32_t mode, C0, C1, Co, kd, kh, kw;
} e3b_pack_job;
int64_t e3b_pack_job_table_bytes(int njobs);
int e3b_pack_jobs_fill(const e3b_pack_job* jobs, int njobs, void* host_table, int64_t* total_blocks);
int e3b_pack_weights_batched(const void* device_table, int njobs, int64_t total_blocks, void* stream);

/* The per-tensor power-of-two weight scales (`wscale` above) of all weight tensors of a network in one launch:
 * table[2*i] = 2^k, table[2*i+1] = 2^-k with k = -floor(log2(max |w_i|)), so that the scaled maximum lands in [1, 2)
 * (k = 0 for all-zero or non-finite tensors, |k| <= 100).  device_jobs: njobs descriptors in DEVICE memory; scratch:
 * 2 * njobs uint32 of device memory, zero before the first call (every call leaves it zero again). */
typedef struct e3b_ws_job { const float* w; int64_t n; } e3b_ws_job;
int e3b_weight_scales(const e3b_ws_job* device_jobs, int njobs, float* table, uint32_t* scratch, void* stream);

/* ---- convolution ------------------------------------------------------------------------------
 * Implicit-GEMM convolution on tcgen05 tensor cores (kind::f16: fp16 operands, fp32 accumulate).
 * Replaces nn.Conv3d/Conv2d behind conv3 (models/unet.py:131-149) incl. its dgrad, the virtual
 * torch.cat((updec, enc), 1) in front of UpConv.conv1 (unet.py:399), and nn.ConvTranspose3d/2d
 * behind upconv2 (unet.py:152-165, scatter=1). */
typedef struct e3b_conv_args {
    const void* src0; int32_t C0;              /* QH source 0 (N, C0, D, H, W) */
    const void* src1; int32_t C1;              /* optional QH source 1 = channels after source 0 */
    int32_t N, D, H, W;                        /* extents of source 0 */
    int32_t D1, H1, W1;                        /* extents of source 1 (>= source 0's) */
    int32_t off1_d, off1_h, off1_w;            /* autocrop centre-crop offset into source 1 (unet.py:303-324) */
    int32_t kd, kh, kw;                        /* taps per dim: 1 or 3 */
    int32_t pd, ph, pw;                        /* zero padding per dim (0..2) */
    const void* wpk;                           /* e3b_pack_weights image */
    const float* bias; int32_t n_bias;         /* bias[n_bias] per output channel (columns >= n_bias get 0) or NULL */
    int32_t n_total;                           /* padded N space (multiple of 16) */
    void* dst0; int32_t Cd0;                   /* QP (fp32) output, channels [0, Cd0); QH with half_out */
    float* dst1; int32_t Cd1;                  /* optional 2nd QP output: channels after pad8(Cd0) */
    int32_t relu;                              /* fuse ReLU (eval-mode BN folded into weights) */
    int32_t half_out;                          /* write dst0 as a QH operand tensor: the output feeds the next MMA */
    const float* out_scale;                    /* optional device scalar multiplied into the accumulators before the
                                                  bias (e3b_norm_bwd_args.dy_scale + 2: undoes the gradient scale) */
    const float* w_unscale;                    /* optional second device scalar, multiplied likewise: 2^-k of a weight image
                                                  packed with the power-of-two scale 2^k (e3b_pack_weights wscale) */
    double* stats; int32_t stats_channels;     /* optional [N][stats_channels][2] sum/sumsq of the output (fp64;
                                                  zeroed by this call) for the following Group/BatchNorm */
    int32_t scatter, sd, sh, sw;               /* transposed conv: column = tap*pad16(Cd0)+co -> fine voxel */
    int32_t Ds, Hs, Ws;                        /* scatter: (cropped) fine output extents (autocrop :294-301) */
    int32_t force_tz;                          /* 0 = auto; tests only */
    int32_t variant;                           /* 0: halo-tile kernel (wpk packed with mode 0..3);
                                                  1: z-stacked kernel (wpk packed with mode 4/5), see e3b_conv_variant */
} e3b_conv_args;
int e3b_conv(const e3b_conv_args* args, void* stream);
/* Which kernel serves a convolution best: 1 = the z-stacked kernel (3x3x3 taps, n_total <= 80, weight image
 * resident in shared memory: the three z taps are stacked in the MMA's N so that narrow layers are not
 * operand-fetch bound), else 0.  The caller packs the weights accordingly (mode 4/5 vs 0/1) and passes the same
 * value in e3b_conv_args.variant.  Both kernels compute the same function. */
int e3b_conv_variant(int C0, int C1, int n_total, int kd, int kh, int kw, int scatter);
/* Developer aid: with the environment variable E3B_CONV_DEBUG set, e3b_conv accumulates per-role cycle
 * counters (producer / MMA issuer / epilogue barrier waits); this reads (and optionally resets) them. */
int e3b_debug_conv_counters(unsigned long long* out16, int reset);
/* Developer aid (environment variable E3B_ZS_DEBUG): progress words of the z-stacked kernel's roles, 16 per CTA,
 * kept in host-mapped memory so that they can be read after a trapped launch. */
int e3b_debug_zs_read(uint32_t* out, int n);
/* Developer aid (E3B_ZS_PROF): per-role cycle counters of the z-stacked kernel summed over CTAs (producer total /
 * wait, issuer total / waits / issue, epilogue total / wait / load / store / statistics / release). */
int e3b_debug_zs_prof(unsigned long long* out16, int reset);
/* phase timeline of the last e3b_norm_bwd_fused launch made with E3B_FUSED_PROF set: [cta 2][round 4][stamp 8] (ns) */
int e3b_debug_fused_prof(unsigned long long* out64);

/* Weight gradient: dW[tap][ci][co] = sum_voxels x[v + tap - pad][ci] * dy[v][co]  (conv backward-filter
 * of nn.Conv3d at unet.py:131-149; with taps=1 on (x, space-to-depth dy) also ConvTranspose's).
 * fp16 operands (kind::f16), fp32 accumulate.  The operands are the QH tensors themselves, read as MN-major MMA operands
 * (K = 16 consecutive x-voxels per MMA; stencil shifts are 16-byte start-address offsets, y taps are stacked in the
 * MMA's M, z taps in its N):  src0 (N, C0, D, H, W);  src1 (N, C1, D1, H1, W1) read at the voxel offset off1 (the
 * centre-cropped skip tensor of autocrop, unet.py:303-324: any offset);  dy (N, Co, Do, Ho, Wo), usually carrying the
 * power-of-two scale of e3b_norm_bwd_args.dy_scale.  Result is written in torch layout:
 *   layout 0: dw (Co, C0+C1, kd, kh, kw)        (Conv)
 *   layout 1: dw (C0, Co/ntap_up, sd, sh, sw) with dy channels = tap*pad8(Co_up)+co  (ConvTranspose)
 * `workspace` holds split-K partials: e3b_wgrad_workspace_floats() floats. */
typedef struct e3b_wgrad_args {
    const void* src0; int32_t C0;
    const void* src1; int32_t C1;
    int32_t N, D, H, W;
    int32_t D1, H1, W1, off1_d, off1_h, off1_w;
    const void* dy; int32_t Co;
    int32_t kd, kh, kw, pd, ph, pw;
    float* dw; int32_t layout; int32_t up_taps; int32_t up_co;
    float* workspace;
    const float* dy_unscale;                   /* optional device scalar multiplied into dw: 2^-k of the scaled gradient
                                                  tensor (e3b_norm_bwd_args.dy_scale + 2) */
    int32_t defer_reduce;                      /* 1: leave the split-K partials in `workspace`; dw is written later by
                                                  e3b_wgrad_reduce_batched called with the same arguments */
} e3b_wgrad_args;
int64_t e3b_wgrad_workspace_floats(const e3b_wgrad_args* args);
int e3b_wgrad(const e3b_wgrad_args* args, void* stream);
/* The split-K reductions of n deferred e3b_wgrad calls (args[i] as passed to them; workspaces still alive) in one launch per
 * 16 layers: the weight gradients of a backward pass are only needed by the optimizer at its end. */
int e3b_wgrad_reduce_batched(const e3b_wgrad_args* args, int n, void* stream);

/* ---- normalisation + activation (+ pooling) -----------------------------------------------------
 * get_normalization (unet.py:77-111) + get_activation (:183-199) + MaxPool(ceil_mode) (:225-229).
 * Activation code `relu` (here and in e3b_norm_bwd_args): 0 identity ('lin'), 1 the leaky-ReLU family with negative slope
 * `act_slope` (0 = nn.ReLU, 0.1 = 'leaky', (lower+upper)/2 = eval-mode nn.RReLU), 2 nn.SiLU.  nn.PReLU(num_parameters=1) is code 1
 * with the learned slope read from device memory (`act_slope_dev`); its gradient comes out of the three-kernel backward.
 * mode: 0 none, 1 group/instance (G groups), 2 batch (training: batch stats + running update),
 *       3 batch eval (running stats).
 * finalize: stats [N][C][2] (from e3b_conv) -> per-(n,c) scale/shift and mean/rstd ([N][pad8(C)]). */
int e3b_norm_finalize(const double* stats, int mode, int G, int N, int C, int64_t S,
                      const float* gamma, const float* beta, float eps,
                      float* running_mean, float* running_var, float momentum,
                      float* scale, float* shift, float* mean, float* rstd, void* stream);
/* a = act(y*scale+shift): y QP (fp32), a QH; if pooled != NULL also pooled = maxpool_{(pk_d,pk_h,pk_w), ceil}(a) (QH).
 * scale/shift NULL = identity; a NULL = only the pooled tensor is written.  y_is_half: y is itself a QH
 * activation (eval path: the conv epilogue already activated it) and is only pooled.
 * pool_idx (optional, uint8 (N, pad8(C)/4, Dp, Hp, Wp, 4)): per pooled voxel and channel the window slot
 * ((dz*pk_h+dy)*pk_w+dx) of the first maximum -- what nn.MaxPool3d(return_indices) would give; consumed by
 * e3b_norm_bwd_*. */
int e3b_norm_act(const void* y, const float* scale, const float* shift, void* a, void* pooled, uint8_t* pool_idx,
                 int N, int C, int D, int H, int W, int pk_d, int pk_h, int pk_w, int relu, float act_slope,
                 const float* act_slope_dev, int y_is_half, void* stream);

/* backward of conv -> norm -> relu [-> pool] as autograd derives it:
 *   dr  = (g0 + g1 + unpool(gp)) * act'(y*scale+shift)   ([a > 0] for ReLU)   g0,g1: same extents as y (either may be NULL)
 *   reduce:   sums[n][c] = (sum dr, sum dr*xhat)   (fp64 atomics, [N][pad8(C)][2]);  amax = (max|dr|, max|xhat|)
 *   finalize: m1,m2 per (n,c); dgamma, dbeta, dbias (conv bias grad); a bound on |dy| -> dy_scale[0]
 *   apply:    dy = rstd * (gamma*dr - m1 - xhat*m2), written as the QH operand 2^k * dy (k from the bound so
 *             that |2^k dy| <= 2^14; 2^k -> dy_scale[1], 2^-k -> dy_scale[2]), or space-to-depth (s2d=1:
 *             channel = tap*pad8(C)+c on the grid (ceil(D/sd),..)) for the transposed conv backward. */
typedef struct e3b_norm_bwd_args {
    const float* y;                            /* conv output (pre-norm); with scale == NULL: the activation itself */
    const float* scale; const float* shift;    /* forward affine [N][pad8(C)] from e3b_norm_finalize: the activation
                                                  a = tf32(relu(y*scale+shift)) is recomputed, not re-read */
    const float* g0; const float* g1; const float* gp;
    const uint8_t* pool_idx;                   /* forward arg-max slots of the pooling (required with gp) */
    int32_t N, C, D, H, W;
    int32_t pk_d, pk_h, pk_w;                  /* pooling kernel of gp (if gp) */
    int32_t mode, G; float eps;
    const float* gamma; const float* mean; const float* rstd;   /* mean/rstd [N][pad8(C)] */
    const double* fwd_stats;                   /* [N][C][2] forward sums (for dbias) */
    double* sums;                              /* [N][pad8(C)][2] */
    uint32_t* amax;                            /* [N][pad8(C)][2] workspace (float bits) */
    float* dy_scale;                           /* [4]: bound bits, 2^k, 2^-k, unused */
    float* m1; float* m2;                      /* [N][pad8(C)] */
    float* dgamma; float* dbeta; float* dbias; /* [C] each; may be NULL */
    void* dy; int32_t s2d, sd, sh, sw;         /* QH output (scaled) */
    int32_t relu;
    int32_t g1_crop;                           /* 1: g1 is the gradient of a CENTRE-CROPPED view of this tensor (autocrop,
                                                  unet.py:303-324; the backward of the slice is a zero pad): it has extents
                                                  (g1_D,g1_H,g1_W) and is added inside the box starting at (g1_od,g1_oh,g1_ow) */
    int32_t g1_od, g1_oh, g1_ow, g1_D, g1_H, g1_W;
    float act_slope;                           /* negative slope of activation code relu == 1 (0 = ReLU) */
    const float* act_slope_dev;                /* nn.PReLU: the slope in device memory (overrides act_slope; three-kernel path) */
    double* slope_sums;                        /* nn.PReLU: workspace [N][pad8(C)] (zeroed by e3b_norm_bwd_reduce) */
    float* dslope;                             /* nn.PReLU: gradient of the slope, written by e3b_norm_bwd_finalize */
} e3b_norm_bwd_args;
/* The three passes as ONE persistent kernel (reduce, grid barrier, finalize, apply), one sample at a time for per-sample
 * statistics (group / instance / none): the apply pass re-reads the sample out of L2, so y and the incoming gradient cross
 * HBM once.  Needs sums, amax and dy_scale as one contiguous workspace (in this order), at most 512 channels and, for
 * s2d, extents divisible by the stride, and no PReLU slope gradient (dslope); otherwise use the three calls below. */
int e3b_norm_bwd_fused(const e3b_norm_bwd_args* args, void* stream);
int e3b_norm_bwd_reduce(const e3b_norm_bwd_args* args, void* stream);
int e3b_norm_bwd_finalize(const e3b_norm_bwd_args* args, void* stream);
int e3b_norm_bwd_apply(const e3b_norm_bwd_args* args, void* stream);

/* ---- options off the default UNet configuration ---------------------------------------------------
 * merge_mode='add' (unet.py:399-401): dst = a + b[centre crop at (off_d, off_h, off_w)], QH operand tensors;
 * a, dst (N, C, D, H, W), b (N, C, D1, H1, W1) -- the skip tensor after autocrop (unet.py:303-324). */
int e3b_add_qh(const void* a, const void* b, void* dst, int N, int C, int D, int H, int W, int D1, int H1, int W1,
               int off_d, int off_h, int off_w, void* stream);
/* up_mode='resizeconv_*' (ResizeConv, unet.py:411-449): nn.Upsample(scale_factor = (sd, sh, sw) in {1, 2}, mode 'nearest'
 * (linear = 0) or 'trilinear' / 'bilinear' with align_corners=False (linear = 1)) of a QH tensor (N, C, d, h, w), written
 * into a QH tensor of extents (Dp, Hp, Wp) in which fine voxel f sits at f + off per axis, only fine voxels [0, R) per axis
 * are stored and everything else is zero: the caller makes the zero padding of the conv that follows explicit (off = 1 on
 * 3-tap axes) and folds autocrop's crop of that conv's OUTPUT (unet.py:294-301) into R, so that the conv itself is a
 * plain VALID one.  e3b_upsample_bwd_qp is the transpose: the fp32 QP gradient of that padded tensor (the dgrad output)
 * gathered onto the coarse grid. */
int e3b_upsample_qh(const void* src, void* dst, int N, int C, int d, int h, int w, int Dp, int Hp, int Wp, int sd, int sh, int sw,
                    int off_d, int off_h, int off_w, int Rd, int Rh, int Rw, int linear, void* stream);
int e3b_upsample_bwd_qp(const float* gfine, float* gcoarse, int N, int C, int d, int h, int w, int Dp, int Hp, int Wp, int sd, int sh,
                        int sw, int off_d, int off_h, int off_w, int Rd, int Rh, int Rw, int linear, void* stream);

/* The residual shortcut of resunet's ConvBlock (models/resunet.py:252-261: `y = conv2(..); y += proj(inp); y = norm2(y)`):
 * y (QP fp32, N x C x S voxels, in place) += r, r the QP fp32 output of the 1x1x1 projection or -- identity shortcut -- the
 * QH activation itself (r_is_half).  stats (optional, [N][C][2] fp64, zeroed here): sum / sum of squares of the SUM, what
 * e3b_norm_finalize needs for the norm that follows.
 * e3b_qp_axpy: dst (QP fp32, in place) += alpha * src (QP fp32 or QH fp16; alpha a device scalar, NULL = 1): the
 * shortcut's gradient added to the gradient of the block input (with src = the scaled fp16 gradient dy and
 * alpha = e3b_norm_bwd_args.dy_scale + 2 for the identity shortcut). */
int e3b_residual_add(float* y_qp, const void* r, int r_is_half, double* stats, int N, int C, int64_t S, void* stream);
int e3b_qp_axpy(float* dst_qp, const void* src, int src_is_half, const float* alpha, int N, int C, int64_t S, void* stream);

/* ---- 1x1x1 head --------------------------------------------------------------------------------
 * conv_final (unet.py:881,912) fused with Predictor's Softmax(1) / Argmax (inference.py:443-456,202-212).
 * out_mode 0: logits float NCDHW; 1: softmax float NCDHW; 2: argmax uint8 (N,1,D,H,W).
 * Writes only the box [c0, c0+cn) per dim of each sample into a destination of extents (Dd,Hd,Wd) at
 * origin dst_origin[3*n..] (device int32; NULL = 0): the Predictor's crop-and-place
 * (inference.py:147-151,188-197).  Destination batch index = dst_n[n] (NULL = n). */
typedef struct e3b_head_args {
    const void* a; int32_t N, C, D, H, W;      /* QH input */
    const float* w; const float* b; int32_t Co;/* torch (Co, C, 1,1,1) */
    int32_t out_mode;
    void* dst; int32_t Dd, Hd, Wd;
    int32_t c0_d, c0_h, c0_w, cn_d, cn_h, cn_w;
    const int32_t* dst_origin;
    int32_t dst_single;                        /* 1: all tiles write into sample 0 of dst */
    int32_t flip;                              /* bits D/H/W: `a` is the network output on a mirrored tile; voxel v of the box
                                                  is read at its mirrored position (FlipAugment.backward, inference.py:225-226) */
    int32_t accumulate; float acc_scale;       /* out_mode 0/1: dst = (accumulate ? dst : 0) + acc_scale * value (acc_scale 0
                                                  is read as 1): the mean over test-time augmentations (inference.py:507-517) */
    int32_t use_threshold; float threshold;    /* out_mode 2: probabilities <= threshold are zeroed before the argmax
                                                  (nn.Threshold(t, 0) + Argmax, inference.py:448-453); implies softmax */
    int32_t round_half;                        /* out_mode 0/1: values rounded through fp16 (Predictor(float16=True) returns what a
                                                  .half() model computed, inference.py:402-408,445-446) */
} e3b_head_args;
int e3b_head(const e3b_head_args* args, void* stream);
/* argmax over the channels of a probability volume (N, C, S) float -> uint8 (N, 1, S), with the optional threshold of
 * e3b_head_args: the deferred argmax after the test-time-augmentation mean (inference.py:519-523). */
int e3b_prob_argmax(const float* prob, uint8_t* dst, int N, int C, int64_t S, int use_threshold, float threshold, void* stream);
/* backward: dl NCDHW (N,Co,D,H,W), a QH -> da QP; dw (Co,C), db (Co) via workspace double[Co*(C+1)] (zeroed here) */
int e3b_head_bwd(const float* dl, const void* a, const float* w, float* da, float* dw, float* db,
                 double* workspace, int N, int C, int Co, int D, int H, int W, void* stream);

/* ---- Dice loss ----------------------------------------------------------------------------------
 * The generalised Dice loss of modules/loss.py:165-233 (`dice_loss` / `DiceLoss.forward`) on the network's NCDHW logits
 * (N, C, S voxels):  loss = mean_c w_c * (1 - (2 * sum p_c t_c + smooth) / (sum p_c + sum t_c + smooth + eps)),
 * p = softmax(logits) if apply_softmax else logits, t = one-hot target.  One read of the logits forward (the reference
 * materialises probabilities, the one-hot tensor, their product and their sum), one read + one write backward.
 * target: dense class indices int64 (N, S), or target_onehot float (N, C, S) (exactly one of the two).
 * weight: [weight_n] with weight_n = 1 or C (NULL = 1).  sums: workspace double[3*C] (zeroed here);
 * loss: float[1]; coef: float[2*C] handed to e3b_dice_bwd together with the upstream gradient gout (device float[1]). */
int e3b_dice_fwd(const float* logits, const int64_t* target, const float* target_onehot, const float* weight, int weight_n,
                 int N, int C, int64_t S, int apply_softmax, float smooth, float eps, double* sums, float* loss, float* coef,
                 void* stream);
int e3b_dice_bwd(const float* logits, const int64_t* target, const float* target_onehot, const float* coef, const float* gout,
                 float* dlogits, int N, int C, int64_t S, int apply_softmax, void* stream);

#ifdef __cplusplus
}
#endif
#endif
