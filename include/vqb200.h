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
    VqbView dy_view; /* normally the dense view of dy       */
    VqbView views[VQB_MAX_VIEWS];
    VqbTap taps[VQB_MAX_TAPS];
} VqbWgradDesc;

int vqb_wgrad_gemm(const VqbWgradDesc* d, const void* dy, const void* x, float* partial, void* stream);

/* library / device info */
const char* vqb_last_error(void);
int vqb_version(void);
int vqb_device_ok(void); /* 1 if the current device is sm_100 and the TMA driver entry point resolved */
int vqb_kernel_launch_count(void); /* number of kernels this library launched in this process */

#ifdef __cplusplus
}
#endif
#endif /* VQB200_H_ */
