/*
 * vqb200.h — C ABI of libvqb200.so, the B200 (sm_100a) native layer under the
 * vqgan-training hot path (Encoder -> reg -> Decoder fwd/bwd + LPIPS/VGG + PatchD + losses).
 *
 * The reference (cloneofsimo/vqgan-training) has NO native layer: its boundary is the Python
 * surface (ae.py / utils.py / vae_trainer.py) and every device op is a PyTorch library call.
 * Each entry point below therefore cites the reference *call site(s)* whose ATen/cuDNN library
 * call it replaces (file:line into the reference tree).
 *
 * Conventions
 *   - plain pointers + sizes only; no torch / C++ types. All pointers are DEVICE pointers unless
 *     the name ends in _host. `stream` is a cudaStream_t passed as void*.
 *   - activations are NHWC bf16 with C a multiple of 8 ("internal layout"); master weights,
 *     gradients and module-boundary tensors are fp32 NCHW / OIHW (the reference's layout).
 *   - every function returns 0 on success or a negative VQB_E* code; vqb_last_error() gives a
 *     message. There is no CPU fallback: on a machine without an sm_100 device the compute entry
 *     points fail with VQB_ENODEVICE.
 */
#ifndef VQB200_H_
#define VQB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VQB_OK 0
#define VQB_EINVAL (-1)    /* bad shape / alignment / flag combination */
#define VQB_ENODEVICE (-2) /* no sm_100 device or driver entry point missing */
#define VQB_ECUDA (-3)     /* a CUDA runtime / driver call failed */

#define VQB_MAX_VIEWS 16
#define VQB_MAX_TAPS 16

/* epilogue flags of vqb_conv_gemm */
#define VQB_EPI_BIAS 1   /* += bias[c]                                                    */
#define VQB_EPI_RES 2    /* += res[pixel][c]   (bf16, same addressing as out)             */
#define VQB_EPI_RELU 4   /* max(.,0)                                                      */
#define VQB_EPI_MASK 8   /* *= (mask[pixel][c] > 0)  (bf16, same addressing as out)       */
#define VQB_EPI_STATS 16 /* accumulate per-(n,channel) sum / sum-of-squares of the bf16-rounded
                            output into stats[n][Cout][2] (fp32 atomics) for GroupNorm      */

/* A strided 4-D view [Nv][Hv][Wv][C] (channel stride 1) of an NHWC bf16 tensor. */
typedef struct VqbView {
    int64_t offset;     /* element offset from the tensor base pointer */
    int32_t Wv, Hv, Nv; /* extents */
    int32_t _pad;
    int64_t sw, sh, sn; /* strides in elements */
} VqbView;

/* One filter tap: reads view `view` at (w + dw, h + dh); out-of-range reads are zero. */
typedef struct VqbTap {
    int32_t view, dw, dh, _pad;
} VqbTap;

/*
 * Implicit-GEMM convolution  out[n,h,w,co] = epi( sum_t sum_c A_view(t)[n, h+dh_t, w+dw_t, c] * Wp[co][t][c] ).
 * Covers: 3x3 s1 p1 and 1x1 convs, their dgrad (packed transposed/rotated weights), the
 * stride-2 (0,1,0,1)-padded Downsample conv (4 parity views), its transposed dgrad (per output
 * parity class), and the non-overlapping k4s4/k2s2 PatchDiscriminator heads (one view per tap).
 * Replaces: nn.Conv2d forward + autograd dgrad at ae.py:105-117,143-154,160-167,197-199,230-232,
 * 282-284,307-309; torchvision VGG convs reached from utils.py:95-111,150-154; heads utils.py:156-185.
 */
typedef struct VqbConvDesc {
    int32_t C;       /* channels of the A tensor = K per tap (multiple of 8)     */
    int32_t Cout;    /* GEMM N                                                   */
    int32_t N, H, W; /* output pixel grid, GEMM M = N*H*W                        */
    int32_t nviews, ntaps;
    int32_t flags;   /* VQB_EPI_*                                                */
    int32_t out_f32; /* 0: out is bf16, 1: out is fp32                           */
    int32_t _pad;
    int64_t on, oh, ow, oc; /* out element address = out + n*on + h*oh + w*ow + c*oc */
    VqbView views[VQB_MAX_VIEWS];
    VqbTap taps[VQB_MAX_TAPS];
} VqbConvDesc;

int vqb_conv_gemm(const VqbConvDesc* d, const void* a, const void* w_packed /* bf16 [Cout][ntaps*C] */,
                  const float* bias, const void* res, const void* mask, void* out, float* stats, void* stream);

/*
 * Weight gradient  dWp[co][t][c] = sum_{n,h,w} dy[n,h,w,co] * X_view(t)[n, h+dh_t, w+dw_t, c]
 * as a split-K tcgen05 GEMM with MN-major operands. `partial` is fp32 [ksplit][Cout][ntaps*C];
 * vqb_wgrad_reduce sums the splits and writes the OIHW fp32 gradient.
 * Replaces: the wgrad half of convolution_backward for every trainable conv (autograd of the
 * call sites listed at vqb_conv_gemm).
 */
typedef struct VqbWgradDesc {
    int32_t C;       /* channels of x                       */
    int32_t Cout;    /* channels of dy                      */
    int32_t N, H, W; /* dy pixel grid                       */
    int32_t nviews, ntaps;
    int32_t ksplit;
    int64_t ld_override; /* row pitch (floats) of partial; 0 = ntaps*roundup(C,64)                     */
    int64_t col_offset;  /* first column of partial this launch writes (several launches, one buffer) */
    VqbView dy_view;     /* normally the dense view of dy                                              */
    VqbView views[VQB_MAX_VIEWS];
    VqbTap taps[VQB_MAX_TAPS];
} VqbWgradDesc;

/*
 * Data-gradient launch fused with the statistics pass of a GroupNorm(+swish) backward. When the conv being
 * differentiated consumed y = swish(GroupNorm(x)) (ae.py:124-131: norm1 -> conv1, norm2 -> conv2), its data gradient IS
 * the dy of that GroupNorm; the epilogue reads the matching x tile (TMA-prefetched, same addressing as `out`) and
 * accumulates cs[n][c] = (sum_p du, sum_p du * xhat), du = dy * swish'(gamma*xhat + beta), into the pre-zeroed
 * cs[N][Cout][2] with fp32 atomics. vqb_gn_silu_bwd_pre then only finalises and applies (x, dy are not read a second
 * time for the reduction: 10 -> 6 bytes per element for the GroupNorm backward).
 */
typedef struct VqbGnBwdFuse {
    const void* x;      /* bf16 NHWC input of the GroupNorm, same geometry as `out` */
    const float* mr;    /* [N][groups][2] mean, rstd saved by the forward */
    const float* gamma; /* [Cout] */
    const float* beta;  /* [Cout] */
    float* cs;          /* [N][Cout][2], pre-zeroed */
    int32_t groups;
    int32_t _pad;
} VqbGnBwdFuse;
int vqb_conv_gemm_gnbwd(const VqbConvDesc* d, const void* a, const void* w_packed, const float* bias, void* out,
                        const VqbGnBwdFuse* gn, void* stream);
int vqb_conv_gnbwd_ok(const VqbConvDesc* d, int groups);

/* 1 if vqb_conv_gemm supports VQB_EPI_STATS for this descriptor (staged epilogue, whole sub-tiles inside one image) */
int vqb_conv_stats_ok(const VqbConvDesc* d);

int vqb_wgrad_gemm(const VqbWgradDesc* d, const void* dy, const void* x, float* partial, void* stream);

/* number of fp32 columns per Cout row of the wgrad partial buffer: ntaps * roundup(C, 64) */
int vqb_wgrad_cols(int ntaps, int C);

/*
 * grad[co][ci][tap] (OIHW fp32) (+)= sum_s partial[s][co][slot*C64 + ci], slot -> tap through tapmap_dev (int32[nslots],
 * device memory). Deterministic (no atomics). Replaces the tail of aten::convolution_backward (weight gradient layout).
 */
int vqb_wgrad_reduce(const float* partial, float* grad, int ksplit, int Cout, int CoutPad, int Cin, int T, int nslots,
                     int C64, const int* tapmap_dev, int accumulate, void* stream);

/*
 * OIHW fp32 master weights -> bf16 [R][nslots][Kpad] GEMM operand (R = Cin if transpose else Cout; transpose = dgrad
 * layout; tapmap_dev selects / reorders filter taps, e.g. the 180-degree rotation of the data gradient).
 * Replaces: the per-step fp32->bf16 weight casts of torch.autocast (vae_trainer.py:453,623) and cuDNN's internal
 * filter transforms.
 */
int vqb_pack_weights(const float* w_oihw, void* out, int Cout, int Cin, int T, int nslots, const int* tapmap_dev,
                     int transpose, int Kpad, void* stream);

/*
 * Module-boundary layout conversion. y[n,h,w,c] = (x[n,c,h,w] - shift[c]) * inv_scale[c] as bf16 NHWC with Cpad
 * channels (pad = 0); shift/inv_scale may be NULL. The scaled form is LPIPS/PatchD ScalingLayer (utils.py:70-71).
 * vqb_nhwc_to_nchw is the inverse / the backward of it (gx = g * inv_scale).
 */
/* folded variants (slot -> SET of taps as a bit mask): nearest-2x upsample fused into 4 phase convs with 2x2 taps */
int vqb_pack_weights_fold(const float* w_oihw, void* out, int Cout, int Cin, int T, int nslots, const int* tapmask_dev,
                          int transpose, int Kpad, void* stream);
int vqb_wgrad_reduce_fold(const float* partial, float* grad, int ksplit, int Cout, int CoutPad, int Cin, int T,
                          int nslots, int C64, const int* tapmask_dev, void* stream);

int vqb_nchw_to_nhwc(const float* x, void* y, int N, int C, int H, int W, int Cpad, const float* shift,
                     const float* inv_scale, void* stream);
int vqb_nhwc_to_nchw(const void* g, float* gx, int N, int C, int H, int W, int Cpad, const float* inv_scale,
                     void* stream);
/* variants that write / read the interior of a zero-framed [N][H+2p][W+2p][Cpad] buffer (first-layer "fat pixel" conv:
 * three horizontally adjacent 8-channel pixels are one 24-channel K run, 3 taps instead of 9) */
int vqb_nchw_to_nhwc_pad(const float* x, void* y, int N, int C, int H, int W, int Cpad, int pad, const float* shift,
                         const float* inv_scale, void* stream);
int vqb_nhwc_to_nchw_pad(const void* g, float* gx, int N, int C, int H, int W, int Cpad, int pad,
                         const float* inv_scale, void* stream);

/*
 * FP32GroupNorm (+ swish) forward / backward on bf16 NHWC: 32 groups, biased variance, eps inside the sqrt, fp32
 * statistics (ae.py:41-53 + ae.py:13-14). mr = [N][G][2] (mean, rstd) kept for the backward.
 * fwd workspace ws: N*C*2 doubles; bwd workspace ws: N*C*2 + N*G*2 floats. `add` (optional) is summed into dx.
 */
int vqb_gn_silu_fwd(const void* x, void* y, const float* gamma, const float* beta, float* mr, double* ws, int N,
                    int HW, int C, int G, float eps, int silu, void* stream);
/* forward when the producing conv already accumulated chsums[N][C][2] (VQB_EPI_STATS): finalise + apply only */
int vqb_gn_silu_fwd_pre(const void* x, void* y, const float* gamma, const float* beta, float* mr, const float* chsums,
                        int N, int HW, int C, int G, float eps, int silu, void* stream);
/* dx_colsum (optional [C] fp32): per-channel sums of dx = bias gradient of the conv that produced x, same pass */
int vqb_gn_silu_bwd(const void* x, const void* dy, const void* add, void* dx, const float* gamma, const float* beta,
                    const float* mr, float* dgamma, float* dbeta, float* ws, int N, int HW, int C, int G, int silu,
                    float* dx_colsum, void* stream);
/* same, when cs[N][C][2] = (sum du, sum du*xhat) was already accumulated by vqb_conv_gemm_gnbwd: finalise + apply only
 * (ws: N*G*2 floats) */
int vqb_gn_silu_bwd_pre(const void* x, const void* dy, const void* add, void* dx, const float* gamma, const float* beta,
                        const float* mr, const float* cs, float* dgamma, float* dbeta, float* ws, int N, int HW, int C,
                        int G, int silu, float* dx_colsum, void* stream);

#ifdef VQB_DEBUG
/* bring-up experiment only (csrc/dbg_shift.cu, libvqb200_dbg.so): one M=128,N=64,K=64 MMA whose A descriptor starts
 * shift_rows rows into a TMA-written 128B-swizzled tile with 8-row groups sbo_bytes apart */
int vqb_dbg_shift_mma(const void* X, int R, const void* B, float* out, int shift_rows, int sbo_bytes, int base_offset,
                      void* stream);
#endif

/*
 * Wavelet front-end of the encoder (--use_wavelet; utils.py:229-247): F.pad(x, 2) + grouped 6x6 stride-2 conv with the
 * four fixed analysis filters filt[4][6][6] (device, fp32), fused with the NCHW fp32 -> NHWC bf16 conversion:
 * y[n][ho][wo][c*4 + band], Cpad channels (pad = 0). Input-side op, no gradient (the input is data).
 */
int vqb_wavelet_fwd(const float* x, void* y, const float* filt, int N, int C, int H, int W, int Cpad, void* stream);

/* nearest-neighbour x2 up-sampling (ae.py:165) and its backward (2x2 sum), bf16 NHWC */
int vqb_upsample2x_fwd(const void* x, void* y, int N, int H, int W, int C, void* stream);
int vqb_upsample2x_bwd(const void* dy, void* dx, int N, int H, int W, int C, void* stream);

/* out[c] = sum over P pixels of x[p][c] : Conv2d bias gradient */
int vqb_colsum(const void* x, float* out, int64_t P, int C, void* stream);

/*
 * 2x2/2 max-pool of the VGG16 trunk (torchvision features[4,9,16,23]) and its backward: first-maximum tie rule of
 * ATen; relu_mask additionally gates by x > 0 (x is a post-ReLU activation); `add` (optional) is summed into dx.
 */
int vqb_maxpool2_fwd(const void* x, void* y, int N, int Ho, int Wo, int C, void* stream);
int vqb_maxpool2_bwd(const void* x, const void* dy, const void* add, void* dx, int N, int Ho, int Wo, int C,
                     int relu_mask, void* stream);

/*
 * One LPIPS layer (utils.py:44-53,134-140): out[n] += mean_p sum_c w[c] (f0/(|f0|+1e-10) - f1/(|f1|+1e-10))^2,
 * and its backward w.r.t. f0 only (frozen trunk, target branch carries no gradient), gated by f0 > 0.
 */
int vqb_lpips_tail_fwd(const void* f0, const void* f1, const float* w, float* out, int N, int HW, int C, void* stream);
int vqb_lpips_tail_bwd(const void* f0, const void* f1, const float* w, const float* g, void* df0, int N, int HW, int C,
                       void* stream);
/*
 * Train-mode variants: the nn.Dropout(0.5) in front of every lin layer (utils.py:79-89) is live in the reference's
 * training loop (LPIPS is never put in eval mode, vae_trainer.py:477). Element (n, p, c) of the squared-difference tensor
 * is kept (and scaled by 2) iff bit ((n*HW + p)*C + c) of a counter-based hash stream of `seed` is set; forward and
 * backward regenerate the same bits, and vqb_lpips_dropout_mask writes them out ([N][HW][C] bytes, 1 = keep) so that a
 * test can feed the identical mask to the reference arithmetic.
 */
int vqb_lpips_tail_fwd_dropout(const void* f0, const void* f1, const float* w, float* out, int N, int HW, int C,
                               uint64_t seed, void* stream);
int vqb_lpips_tail_bwd_dropout(const void* f0, const void* f1, const float* w, const float* g, void* df0, int N, int HW,
                               int C, uint64_t seed, void* stream);
int vqb_lpips_dropout_mask(uint64_t seed, int N, int HW, int C, uint8_t* mask, void* stream);

/*
 * Multi-head self-attention core of AttnBlock (ae.py:74-93): qkv [N][T][3C] bf16 (q | k | v channel blocks, heads of 64
 * channels) -> out [N][T][C] = softmax(q k^T / 8) v, flash-style (no T x T matrix in HBM). lse [N][C/64][T] is saved for
 * the backward; dvec is a workspace of the same shape. Replaces F.scaled_dot_product_attention + einops rearranges.
 */
int vqb_attn_fwd(const void* qkv, void* out, float* lse, int N, int T, int C, void* stream);
int vqb_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse, float* dvec, void* dqkv, int N,
                 int T, int C, void* stream);

/*
 * VQ codebook nearest neighbour (BASELINE.json config 4; the reference has no VQ — semantics pinned by
 * oracle/vq_oracle.py): idx[i] = argmin_j sum_c (z[i][c]-e[j][c])^2 in canonical fp32 order (bit-exact vs the oracle,
 * first index on ties), zq = e[idx], *sqerr += sum (zq - z)^2 when sqerr != NULL.
 */
int vqb_vq_argmin(const float* z, const float* e, long long* idx, float* zq, float* sqerr, int M, int K, int D,
                  void* stream);

/*
 * Optimizer step over ONE flat fp32 buffer holding every tensor of a model (each tensor padded to a multiple of 1024
 * elements): torch.optim.AdamW semantics (decoupled weight decay, bias correction) with up to VQB_ADAMW_MAX_GROUPS
 * hyper-parameter groups. chunk_group[i] (device, uint8) = group of 1024-element chunk i, 255 = skip (the owning tensor
 * received no gradient). grads are multiplied by grad_scale first. groups_host is HOST memory (read during the call).
 * Replaces: optimizer_G.step()/optimizer_D.step() = AdamW(lr groups, wd 1e-3, betas (0.9, 0.95)) at
 * vae_trainer.py:455-475 with the cosine-with-warmup learning rate of :486-490 passed in as groups_host[].lr.
 */
#define VQB_ADAMW_MAX_GROUPS 4
typedef struct VqbAdamwGroup {
    float lr, beta1, beta2, eps, weight_decay;
    int32_t step; /* 1-based step count of this group (bias correction) */
} VqbAdamwGroup;
int vqb_adamw_flat(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, const uint8_t* chunk_group,
                   int64_t nchunks, int ngroups, const VqbAdamwGroup* groups_host, float grad_scale, void* stream);
/* CUDA-graph friendly form: the per-group hyper-parameters (incl. bias corrections) are read from a 28-float DEVICE
 * record at kernel run time. vqb_adamw_fill_record (host function) builds that record from groups_host into host memory;
 * the caller copies it to record_dev on the launch stream before every launch / graph replay. */
int vqb_adamw_fill_record(int ngroups, const VqbAdamwGroup* groups_host, float* record_host);
int vqb_adamw_flat_dev(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, const uint8_t* chunk_group,
                       int64_t nchunks, const float* record_dev, float grad_scale, void* stream);

/*
 * Re-pack every cached bf16 GEMM operand of the fp32 OIHW master weights in one launch (after an optimizer step).
 * jobs_dev: DEVICE array of njobs descriptors; total_blocks = sum over jobs of ceil(R/8)*ceil(Kpad/64)
 * (R = Cin if transpose else Cout; one block packs an 8-row x 64-k tile for every slot; T <= 16).
 *   out[r*ld_r + (slot/sg)*ld_g + (slot%sg)*Kpad + k] = bf16(transpose ? w[k][r][.] : w[r][k][.]); fold: tapmap entries are
 *   bit masks of taps summed in fp32 (vqb_pack_weights_fold), else tap indices (vqb_pack_weights).
 * Replaces: the per-step fp32->bf16 weight casts of torch.autocast (vae_trainer.py:453,623).
 */
typedef struct VqbPackJob {
    const float* w;
    void* out;
    const int* tapmap;
    int32_t Cout, Cin, T, nslots, transpose, Kpad, fold, sg, ld_g, ld_r;
    int32_t first_block; /* prefix sum of the tile-block counts of the preceding jobs */
    int32_t _pad;
} VqbPackJob;
int vqb_pack_weights_multi(const VqbPackJob* jobs_dev, int njobs, int total_blocks, void* stream);

/* library / device info */
const char* vqb_last_error(void);
int vqb_version(void);
int vqb_device_ok(void); /* 1 if the current device is sm_100 and the TMA driver entry point resolved */
int vqb_kernel_launch_count(void); /* number of kernels this library launched in this process */
int vqb_set_debug_mode(int mode);   /* perf-experiment switches: only the -DVQB_DEBUG build accepts mode != 0 */

#ifdef __cplusplus
}
#endif
#endif /* VQB200_H_ */
