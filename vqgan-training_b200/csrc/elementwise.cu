// HBM-bound kernels of the VAE path: layout conversion at the module boundary, weight packing,
// fused GroupNorm(+SiLU) forward / backward, nearest-2x up-sampling, bias gradients, wgrad split
// reduction. All activations are NHWC bf16 with C % 8 == 0; every thread moves 16-byte vectors and
// owns a FIXED 8-channel slot (its channel vector index never changes while it strides over pixels),
// so per-channel affine terms / reductions stay in registers.
//
// Reference semantics: FP32GroupNorm ae.py:41-53 (32 groups, biased variance, eps inside sqrt,
// fp32 math), swish ae.py:13-14, Upsample ae.py:157-167 (nearest), Conv2d bias gradients.
#ifndef VQB_EXACT_SIGMOID
#define VQB_EXACT_SIGMOID 0
#endif
#include "common.cuh"
#include "ptx.cuh"

#include <cstdlib>

namespace vqb {

__device__ __forceinline__ void load8(const __nv_bfloat16* p, float (&f)[8]) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ uint4 ldg16(const __nv_bfloat16* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ void cvt8(const uint4& u, float (&f)[8]) {
    float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float (&f)[8]) {
    uint4 u;
    u.x = pack_bf16x2(f[0], f[1]);
    u.y = pack_bf16x2(f[2], f[3]);
    u.z = pack_bf16x2(f[4], f[5]);
    u.w = pack_bf16x2(f[6], f[7]);
    *reinterpret_cast<uint4*>(p) = u;
}
// sigmoid(x) = 0.5 + 0.5 tanh(x/2) with the single-instruction MUFU.TANH (abs error of the sigmoid <= ~2.5e-4, an order
// of magnitude below the bf16 rounding of the activations it multiplies): one SFU op instead of ex2 + rcp. The GroupNorm
// kernels sit at the SFU / FP32-issue / HBM triple point (ncu: XU 42 %, issue 52 %, DRAM 57 %), so this is time.
__device__ __forceinline__ float sigmoidf_(float x) {
#if VQB_EXACT_SIGMOID
    return 1.f / (1.f + __expf(-x));
#else
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * x));
    return fmaf(0.5f, t, 0.5f);
#endif
}

// tanh for the swish derivative: MUFU.TANH by default; with -DVQB_EXACT_SIGMOID the exact form through exp
__device__ __forceinline__ float tanh_fast(float x) {
#if VQB_EXACT_SIGMOID
    return 2.f / (1.f + __expf(-2.f * x)) - 1.f;
#else
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(x));
    return t;
#endif
}

// ------------------------------------------------------------------ weight packing
// out[r][slot][k] (bf16), r < R, k < Kpad:  transpose ? w[k][r][tap] : w[r][k][tap]   (w is OIHW fp32,
// tap = tapmap[slot] indexes KH*KW), zero for k >= K.
__global__ void pack_weights_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int Cout, int Cin,
                                    int T, int nslots, const int* __restrict__ tapmap, int transpose, int Kpad,
                                    int fold /* 0 none */) {
    const int R = transpose ? Cin : Cout;
    const int K = transpose ? Cout : Cin;
    const int64_t total = static_cast<int64_t>(R) * nslots * Kpad;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int k = static_cast<int>(i % Kpad);
        const int slot = static_cast<int>((i / Kpad) % nslots);
        const int r = static_cast<int>(i / (static_cast<int64_t>(Kpad) * nslots));
        float v = 0.f;
        if (k < K) {
            const int tap = tapmap[slot];
            const int co = transpose ? k : r, ci = transpose ? r : k;
            v = w[(static_cast<int64_t>(co) * Cin + ci) * T + tap];
        }
        out[i] = __float2bfloat16(v);
    }
}

// Folded packing: out[r][slot][k] = sum over the taps in tapmask[slot] (bit t = tap t) of w[..][tap]; the fp32 sum is
// rounded to bf16 once. Used by the nearest-2x-upsample + conv3x3 fusion (4 phase convs with 2x2 folded taps).
__global__ void pack_weights_fold_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int Cout,
                                         int Cin, int T, int nslots, const int* __restrict__ tapmask, int transpose,
                                         int Kpad) {
    const int R = transpose ? Cin : Cout;
    const int K = transpose ? Cout : Cin;
    const int64_t total = static_cast<int64_t>(R) * nslots * Kpad;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int k = static_cast<int>(i % Kpad);
        const int slot = static_cast<int>((i / Kpad) % nslots);
        const int r = static_cast<int>(i / (static_cast<int64_t>(Kpad) * nslots));
        float v = 0.f;
        if (k < K) {
            const int mask = tapmask[slot];
            const int co = transpose ? k : r, ci = transpose ? r : k;
            const float* wp = w + (static_cast<int64_t>(co) * Cin + ci) * T;
            for (int t = 0; t < T; ++t)
                if ((mask >> t) & 1) v += wp[t];
        }
        out[i] = __float2bfloat16(v);
    }
}

// ------------------------------------------------------------------ layout conversion
// y[n,h,w,c] = (x[n,c,h,w] - shift[c]) * inv_scale[c]   (bf16 NHWC, channels >= C zero)
// pad > 0: y is [N][H+2pad][W+2pad][Cpad] (pre-zeroed) and only its interior is written (W needed then)
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, int N, int C, int HW,
                                    int Cpad, const float* __restrict__ shift, const float* __restrict__ inv_scale,
                                    int W, int pad) {
    const int64_t total = static_cast<int64_t>(N) * HW;
    const int H = HW / W;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t n = i / HW, p = i % HW;
        const float* xp = x + n * C * HW + p;
        int64_t opix = i;
        if (pad) {
            const int h = static_cast<int>(p / W), w = static_cast<int>(p % W);
            opix = (n * (H + 2 * pad) + h + pad) * (W + 2 * pad) + w + pad;
        }
        __nv_bfloat16* yp = y + opix * Cpad;
        for (int c0 = 0; c0 < Cpad; c0 += 8) {
            float f[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int c = c0 + j;
                float v = 0.f;
                if (c < C) {
                    v = xp[static_cast<int64_t>(c) * HW];
                    if (shift) v = (v - shift[c]) * inv_scale[c];
                }
                f[j] = v;
            }
            store8(yp + c0, f);
        }
    }
}

// gx[n,c,h,w] = g[n,h,w,c] * inv_scale[c]   (fp32 NCHW out)
__global__ void nhwc_to_nchw_kernel(const __nv_bfloat16* __restrict__ g, float* __restrict__ gx, int N, int C, int HW,
                                    int Cpad, const float* __restrict__ inv_scale, int W, int pad) {
    const int64_t total = static_cast<int64_t>(N) * HW;
    const int H = HW / W;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t n = i / HW, p = i % HW;
        int64_t ipix = i;
        if (pad) {
            const int h = static_cast<int>(p / W), w = static_cast<int>(p % W);
            ipix = (n * (H + 2 * pad) + h + pad) * (W + 2 * pad) + w + pad;
        }
        const __nv_bfloat16* gp = g + ipix * Cpad;
        float* xp = gx + n * C * HW + p;
        for (int c = 0; c < C; ++c) {
            float v = __bfloat162float(gp[c]);
            if (inv_scale) v *= inv_scale[c];
            xp[static_cast<int64_t>(c) * HW] = v;
        }
    }
}

// ------------------------------------------------------------------ GroupNorm forward
// grid (chunks, N); thread t owns channel vector cv = t % V (V = C/8) and pixel rows t / V + k*R.
__global__ void gn_stats_kernel(const __nv_bfloat16* __restrict__ x, double* __restrict__ sums /* [N][C][2] */, int HW,
                                int C, int pix_per_chunk) {
    extern __shared__ float sm[];  // [C][2]
    const int V = C >> 3, R = blockDim.x / V;
    const int cv = threadIdx.x % V, pr = threadIdx.x / V;
    const int n = blockIdx.y;
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sm[i] = 0.f;
    __syncthreads();
    float s[8], q[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
    const int p0 = blockIdx.x * pix_per_chunk;
    const int p1 = min(HW, p0 + pix_per_chunk);
    if (pr < R) {
        const __nv_bfloat16* xb = x + (static_cast<int64_t>(n) * HW) * C + cv * 8;
        int p = p0 + pr;
        for (; p + 3 * R < p1; p += 4 * R) {  // 4 independent 16-byte loads in flight per thread
            uint4 u[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) u[k] = ldg16(xb + static_cast<int64_t>(p + k * R) * C);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float f[8];
                cvt8(u[k], f);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    s[j] += f[j];
                    q[j] += f[j] * f[j];
                }
            }
        }
        for (; p < p1; p += R) {
            float f[8];
            load8(xb + static_cast<int64_t>(p) * C, f);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                s[j] += f[j];
                q[j] += f[j] * f[j];
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            atomicAdd(&sm[(cv * 8 + j) * 2], s[j]);
            atomicAdd(&sm[(cv * 8 + j) * 2 + 1], q[j]);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x)
        atomicAdd(&sums[static_cast<int64_t>(n) * 2 * C + i], static_cast<double>(sm[i]));
}

// mean / rstd per (n, group) from the per-channel double sums.
__global__ void gn_finalize_kernel(const double* __restrict__ sums, float* __restrict__ mr /* [N][G][2] */, int N,
                                   int C, int G, int HW, float eps) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N * G) return;
    const int n = i / G, g = i % G, cpg = C / G;
    double s = 0, q = 0;
    for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
        s += sums[(static_cast<int64_t>(n) * C + c) * 2];
        q += sums[(static_cast<int64_t>(n) * C + c) * 2 + 1];
    }
    const double m = static_cast<double>(cpg) * HW;
    const double mean = s / m;
    double var = q / m - mean * mean;
    if (var < 0) var = 0;
    mr[i * 2] = static_cast<float>(mean);
    mr[i * 2 + 1] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
}

// same, from per-channel fp32 sums accumulated by the producing convolution's epilogue (VQB_EPI_STATS)
__global__ void gn_finalize_f32_kernel(const float* __restrict__ sums, float* __restrict__ mr, int N, int C, int G,
                                       int HW, float eps) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N * G) return;
    const int n = i / G, g = i % G, cpg = C / G;
    double s = 0, q = 0;
    for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
        s += static_cast<double>(sums[(static_cast<int64_t>(n) * C + c) * 2]);
        q += static_cast<double>(sums[(static_cast<int64_t>(n) * C + c) * 2 + 1]);
    }
    const double m = static_cast<double>(cpg) * HW;
    const double mean = s / m;
    double var = q / m - mean * mean;
    if (var < 0) var = 0;
    mr[i * 2] = static_cast<float>(mean);
    mr[i * 2 + 1] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
}

__global__ void gn_apply_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                const float* __restrict__ mr, const float* __restrict__ gamma,
                                const float* __restrict__ beta, int HW, int C, int G, int pix_per_chunk, int silu) {
    const int V = C >> 3, R = blockDim.x / V;
    const int cv = threadIdx.x % V, pr = threadIdx.x / V;
    if (pr >= R) return;
    const int n = blockIdx.y, cpg = C / G;
    float a[8], b[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = cv * 8 + j, g = c / cpg;
        const float mean = mr[(n * G + g) * 2], rstd = mr[(n * G + g) * 2 + 1];
        a[j] = rstd * gamma[c];
        b[j] = beta[c] - mean * a[j];
    }
    const int p0 = blockIdx.x * pix_per_chunk;
    const int p1 = min(HW, p0 + pix_per_chunk);
    const int64_t base = (static_cast<int64_t>(n) * HW) * C + cv * 8;
    int p = p0 + pr;
    for (; p + 3 * R < p1; p += 4 * R) {
        uint4 u4[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) u4[k] = ldg16(x + base + static_cast<int64_t>(p + k * R) * C);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float f[8];
            cvt8(u4[k], f);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float u = fmaf(a[j], f[j], b[j]);
                f[j] = silu ? u * sigmoidf_(u) : u;
            }
            store8(y + base + static_cast<int64_t>(p + k * R) * C, f);
        }
    }
    for (; p < p1; p += R) {
        float f[8];
        load8(x + base + static_cast<int64_t>(p) * C, f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float u = fmaf(a[j], f[j], b[j]);
            f[j] = silu ? u * sigmoidf_(u) : u;
        }
        store8(y + base + static_cast<int64_t>(p) * C, f);
    }
}

// ------------------------------------------------------------------ GroupNorm backward
// per-(n,channel) sums of du and du*xhat, du = dy * silu'(u), u = xhat*gamma + beta
template <int U>
__global__ void gn_bwd_reduce_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
                                     const float* __restrict__ mr, const float* __restrict__ gamma,
                                     const float* __restrict__ beta, float* __restrict__ cs /* [N][C][2] */, int HW,
                                     int C, int G, int pix_per_chunk, int silu) {
    extern __shared__ float sm[];  // [C][2]
    const int V = C >> 3, R = blockDim.x / V;
    const int cv = threadIdx.x % V, pr = threadIdx.x / V;
    const int n = blockIdx.y, cpg = C / G;
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sm[i] = 0.f;
    __syncthreads();
    if (pr < R) {
        // trimmed form (this kernel sat on the FP32-issue / SFU limit): h = x*a2 + b2 (= u/2), t = tanh(h),
        // 2*silu'(u) = (1 + t)(1 + h - h t); s1 = sum 2du, s2 = sum 2du (x - mean); scaled by 1/2 and rstd/2 at the end
        float mean[8], rstd[8], a2[8], b2[8], s1[8], s2[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = cv * 8 + j, g = c / cpg;
            mean[j] = mr[(n * G + g) * 2];
            rstd[j] = mr[(n * G + g) * 2 + 1];
            const float a = gamma[c] * rstd[j];
            a2[j] = 0.5f * a;
            b2[j] = 0.5f * (beta[c] - mean[j] * a);
            s1[j] = s2[j] = 0.f;
        }
        const int p0 = blockIdx.x * pix_per_chunk;
        const int p1 = min(HW, p0 + pix_per_chunk);
        const int64_t base = (static_cast<int64_t>(n) * HW) * C + cv * 8;
        // 4 pixel rows (8 x 16-byte loads) in flight per thread: with 2 rows the kernel sat at 57 % of HBM bandwidth on
        // load latency (ncu: 16 warps/SM, long-scoreboard stalls)
        for (int p = p0 + pr; p < p1; p += U * R) {
            uint4 ux[U], ud[U];
#pragma unroll
            for (int k = 0; k < U; ++k) {
                const bool in = (p + k * R) < p1;
                ux[k] = in ? ldg16(x + base + static_cast<int64_t>(p + k * R) * C) : make_uint4(0, 0, 0, 0);
                ud[k] = in ? ldg16(dy + base + static_cast<int64_t>(p + k * R) * C) : make_uint4(0, 0, 0, 0);
            }
#pragma unroll
            for (int k = 0; k < U; ++k) {
                if ((p + k * R) >= p1) break;
                float f[8], d[8];
                cvt8(ux[k], f);
                cvt8(ud[k], d);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float du2 = 2.f * d[j];
                    if (silu) {
                        const float h = fmaf(f[j], a2[j], b2[j]);
                        const float t = tanh_fast(h);
                        const float r = fmaf(-h, t, h + 1.f);
                        du2 = d[j] * fmaf(t, r, r);
                    }
                    s1[j] += du2;
                    s2[j] = fmaf(du2, f[j] - mean[j], s2[j]);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            atomicAdd(&sm[(cv * 8 + j) * 2], 0.5f * s1[j]);
            atomicAdd(&sm[(cv * 8 + j) * 2 + 1], 0.5f * rstd[j] * s2[j]);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) atomicAdd(&cs[static_cast<int64_t>(n) * 2 * C + i], sm[i]);
}

// gs[n][g] = (sum_c gamma_c*s1, sum_c gamma_c*s2) / m ; dgamma[c] = sum_n s2 ; dbeta[c] = sum_n s1
__global__ void gn_bwd_finalize_kernel(const float* __restrict__ cs, const float* __restrict__ gamma,
                                       float* __restrict__ gs /* [N][G][2] */, float* __restrict__ dgamma,
                                       float* __restrict__ dbeta, int N, int C, int G, int HW) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int cpg = C / G;
    if (i < N * G) {
        const int n = i / G, g = i % G;
        float a = 0.f, b = 0.f;
        for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
            a += gamma[c] * cs[(static_cast<int64_t>(n) * C + c) * 2];
            b += gamma[c] * cs[(static_cast<int64_t>(n) * C + c) * 2 + 1];
        }
        const float m = static_cast<float>(cpg) * HW;
        gs[i * 2] = a / m;
        gs[i * 2 + 1] = b / m;
    }
    if (i < C) {
        float a = 0.f, b = 0.f;
        for (int n = 0; n < N; ++n) {
            a += cs[(static_cast<int64_t>(n) * C + i) * 2];
            b += cs[(static_cast<int64_t>(n) * C + i) * 2 + 1];
        }
        dbeta[i] = a;
        dgamma[i] = b;
    }
}

// dx = rstd * (du*gamma - S1 - xhat*S2) (+ add)
template <bool ADD, int U>
__global__ void __launch_bounds__(256, 2)
gn_bwd_apply_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
                    const __nv_bfloat16* __restrict__ add, __nv_bfloat16* __restrict__ dx,
                    const float* __restrict__ mr, const float* __restrict__ gs, const float* __restrict__ gamma,
                    const float* __restrict__ beta, int HW, int C, int G, int pix_per_chunk, int silu,
                    float* __restrict__ colsum /* [C] or null */) {
    extern __shared__ float sm[];  // [C] (only when colsum != null)
    const int V = C >> 3, R = blockDim.x / V;
    const int cv = threadIdx.x % V, pr = threadIdx.x / V;
    if (colsum) {
        for (int i = threadIdx.x; i < C; i += blockDim.x) sm[i] = 0.f;
        __syncthreads();
    }
    if (pr < R) {
        const int n = blockIdx.y, cpg = C / G;
        // trimmed form: dx = 2du * (gamma rstd / 2) - rstd S1 - (x - mean) rstd^2 S2, 2du as in the reduce kernel
        float mean[8], a2[8], b2[8], k0[8], k2[8], cs8[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = cv * 8 + j, g = c / cpg;
            mean[j] = mr[(n * G + g) * 2];
            const float rstd = mr[(n * G + g) * 2 + 1];
            const float a = gamma[c] * rstd;
            a2[j] = 0.5f * a;
            b2[j] = 0.5f * (beta[c] - mean[j] * a);
            k0[j] = -rstd * gs[(n * G + g) * 2];
            k2[j] = -rstd * rstd * gs[(n * G + g) * 2 + 1];
            cs8[j] = 0.f;
        }
        const int p0 = blockIdx.x * pix_per_chunk;
        const int p1 = min(HW, p0 + pix_per_chunk);
        const int64_t base = (static_cast<int64_t>(n) * HW) * C + cv * 8;
        // U pixel rows in flight per thread (2 or 3 16-byte loads each): U = 3 without the skip-gradient operand, 2 with it
        for (int pp = p0 + pr; pp < p1; pp += U * R) {
            uint4 ux[U], ud[U], ua[U];
#pragma unroll
            for (int k = 0; k < U; ++k) {
                const bool in = (pp + k * R) < p1;
                const int64_t off = base + static_cast<int64_t>(pp + k * R) * C;
                ux[k] = in ? ldg16(x + off) : make_uint4(0, 0, 0, 0);
                ud[k] = in ? ldg16(dy + off) : make_uint4(0, 0, 0, 0);
                if (ADD) ua[k] = in ? ldg16(add + off) : make_uint4(0, 0, 0, 0);
            }
#pragma unroll
            for (int k = 0; k < U; ++k) {
                if ((pp + k * R) >= p1) break;
                float f[8], d[8], r[8];
                cvt8(ux[k], f);
                cvt8(ud[k], d);
                if (ADD) cvt8(ua[k], r);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float du2 = 2.f * d[j];
                    if (silu) {
                        const float h = fmaf(f[j], a2[j], b2[j]);
                        const float t = tanh_fast(h);
                        const float rr = fmaf(-h, t, h + 1.f);
                        du2 = d[j] * fmaf(t, rr, rr);
                    }
                    float v = fmaf(du2, a2[j], fmaf(f[j] - mean[j], k2[j], k0[j]));
                    if (ADD) v += r[j];
                    f[j] = v;
                    // column sums of the bf16 values actually written (= bias gradient of the conv that produced x)
                    cs8[j] += __bfloat162float(__float2bfloat16(v));
                }
                store8(dx + base + static_cast<int64_t>(pp + k * R) * C, f);
            }
        }
        if (colsum) {
#pragma unroll
            for (int j = 0; j < 8; ++j) atomicAdd(&sm[cv * 8 + j], cs8[j]);
        }
    }
    if (colsum) {
        __syncthreads();
        for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(&colsum[i], sm[i]);
    }
}


// ------------------------------------------------------------------ GroupNorm backward, persistent L2-pipelined form
// One persistent launch does reduce AND apply. The two-kernel form reads x and dy twice from HBM (10 B/element, 12 with
// the skip gradient). Here the work is ordered  R(0) R(1) A(0) R(2) A(1) ... A(N-1)  (R(n) = statistics of sample n,
// A(n) = dx of sample n): when A(n) re-reads x, dy of sample n they were streamed at most one sample ago and (for every
// layer of the FLUX config: x + dy of one sample <= 34 MB of the 126 MB L2) are still L2 resident — HBM traffic drops to
// read-once + write-once = 6 (8) B/element. R loads carry an L2 evict_last policy, A loads / dx stores evict_first.
// Sync: R units add their per-channel partials to cs[n] (fp32 atomics) and bump done[n]; A units spin (acquire) until
// done[n] == units. Every CTA walks the same global order and only ever waits on work that precedes its own position in
// every CTA's list, and the grid is sized to be fully co-resident, so the wait cannot deadlock.
//
// Per element (trimmed: the reduce kernel sat on the FP32-issue / SFU limit):  h = x*a2 + b2 (= u/2), t = tanh(h),
// 2*silu'(u) = (1 + t) * (1 + h - h*t);  du2 = dy * that;  S_du = sum du2 / 2,  S_duxh = rstd/2 * sum du2*(x - mean);
// dx = du2*(gamma*rstd/2) - rstd*gS1 - (x - mean)*rstd^2*gS2 (+ add).
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint4 ldg16_hint(const __nv_bfloat16* ptr, uint64_t pol) {
    uint4 v;
    asm volatile("ld.global.nc.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(ptr), "l"(pol));
    return v;
}
__device__ __forceinline__ void stg16_hint(__nv_bfloat16* ptr, const uint4& v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.v4.u32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(ptr), "r"(v.x), "r"(v.y), "r"(v.z),
                 "r"(v.w), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ float tanh_approx(float x) {
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(x));
    return t;
}

template <bool ADD, bool SILU>
__global__ void __launch_bounds__(256, 2)
gn_bwd_persistent_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
                         const __nv_bfloat16* __restrict__ add, __nv_bfloat16* __restrict__ dx,
                         const float* __restrict__ mr, const float* __restrict__ gamma, const float* __restrict__ beta,
                         float* cs /* [N][C][2], zeroed */, int* done /* [groups], zeroed */,
                         float* __restrict__ colsum, int N, int HW, int C, int G, int S /* samples per group */,
                         int ups /* units per sample */, int pix_per_unit, int depth, int hints) {
    extern __shared__ float sm[];  // [2C] partial sums | [C] dx column sums ; then [2G] group sums
    float* sm_gs = sm + 2 * C;
    const int V = C >> 3, R = blockDim.x / V;
    const int cv = threadIdx.x % V, pr = threadIdx.x / V;
    const int cpg = C / G;
    const uint64_t pol_keep = l2_policy_evict_last(), pol_drop = l2_policy_evict_first();
    const int ngroups = (N + S - 1) / S;
    float csum[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) csum[j] = 0.f;

    for (int step = 0; step < ngroups + depth; ++step) {
        // ---------------------------------------------------------------- R(step): statistics of sample group `step`
        if (step < ngroups) {
            const int n0 = step * S, ns = min(S, N - n0), units = ns * ups;
            for (int u = blockIdx.x; u < units; u += gridDim.x) {
                const int n = n0 + u / ups, ch = u % ups;
                float mean[8], a2[8], b2[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int g = (cv * 8 + j) / cpg;
                    mean[j] = mr[(n * G + g) * 2];
                    const float a = __ldg(gamma + cv * 8 + j) * mr[(n * G + g) * 2 + 1];
                    a2[j] = 0.5f * a;
                    b2[j] = 0.5f * (__ldg(beta + cv * 8 + j) - mean[j] * a);
                }
                const int64_t base = (static_cast<int64_t>(n) * HW) * C + cv * 8;
                for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sm[i] = 0.f;
                __syncthreads();
                float s1[8], s2[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) s1[j] = s2[j] = 0.f;
                const int p0 = ch * pix_per_unit, p1 = min(HW, p0 + pix_per_unit);
                for (int p = p0 + pr; p < p1; p += 3 * R) {
                    uint4 ux[3], ud[3];
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const bool in = (p + k * R) < p1;
                        const __nv_bfloat16* xp = x + base + static_cast<int64_t>(p + k * R) * C;
                        const __nv_bfloat16* dp = dy + base + static_cast<int64_t>(p + k * R) * C;
                        ux[k] = in ? (hints ? ldg16_hint(xp, pol_keep) : ldg16(xp)) : make_uint4(0, 0, 0, 0);
                        ud[k] = in ? (hints ? ldg16_hint(dp, pol_keep) : ldg16(dp)) : make_uint4(0, 0, 0, 0);
                    }
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        if ((p + k * R) >= p1) break;
                        float f[8], d[8];
                        cvt8(ux[k], f);
                        cvt8(ud[k], d);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            float du2 = 2.f * d[j];
                            if (SILU) {
                                const float h = fmaf(f[j], a2[j], b2[j]);
                                const float t = tanh_approx(h);
                                const float r = fmaf(-h, t, h + 1.f);
                                du2 = d[j] * fmaf(t, r, r);
                            }
                            s1[j] += du2;
                            s2[j] = fmaf(du2, f[j] - mean[j], s2[j]);
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    atomicAdd(&sm[(cv * 8 + j) * 2], s1[j]);
                    atomicAdd(&sm[(cv * 8 + j) * 2 + 1], s2[j]);
                }
                __syncthreads();
                for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
                    const int g = (i >> 1) / cpg;
                    const float rstd = mr[(n * G + g) * 2 + 1];
                    // cs[n][c] = (sum du, sum du*xhat)
                    atomicAdd(&cs[static_cast<int64_t>(n) * 2 * C + i], sm[i] * ((i & 1) ? 0.5f * rstd : 0.5f));
                }
                __threadfence();
                __syncthreads();
                if (threadIdx.x == 0) atomicAdd(&done[step], 1);
            }
        }
        // ---------------------------------------------------------------- A(step - depth): dx of that sample group
        if (step >= depth) {
            const int gi = step - depth;
            const int n0 = gi * S, ns = min(S, N - n0), units = ns * ups;
            if (static_cast<int>(blockIdx.x) < units) {
                if (threadIdx.x == 0) {
                    int v;
                    do {
                        asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(done + gi) : "memory");
                    } while (v < units);
                }
                __syncthreads();
                int cur_n = -1;
                float mean[8], a2[8], b2[8], k0[8], k2[8];  // (du2 * a2 = du * gamma * rstd)
                for (int u = blockIdx.x; u < units; u += gridDim.x) {
                    const int n = n0 + u / ups, ch = u % ups;
                    if (n != cur_n) {
                        cur_n = n;
                        __syncthreads();
                        // group sums gs[g] = (sum_c gamma_c cs0, sum_c gamma_c cs1) / m from the complete cs[n]
                        for (int g = threadIdx.x; g < G; g += blockDim.x) {
                            float a = 0.f, b = 0.f;
                            for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
                                a = fmaf(gamma[c], __ldcg(&cs[(static_cast<int64_t>(n) * C + c) * 2]), a);
                                b = fmaf(gamma[c], __ldcg(&cs[(static_cast<int64_t>(n) * C + c) * 2 + 1]), b);
                            }
                            const float m = static_cast<float>(cpg) * HW;
                            sm_gs[g * 2] = a / m;
                            sm_gs[g * 2 + 1] = b / m;
                        }
                        __syncthreads();
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const int g = (cv * 8 + j) / cpg;
                            mean[j] = mr[(n * G + g) * 2];
                            const float rstd = mr[(n * G + g) * 2 + 1];
                            const float a = __ldg(gamma + cv * 8 + j) * rstd;
                            a2[j] = 0.5f * a;
                            b2[j] = 0.5f * (__ldg(beta + cv * 8 + j) - mean[j] * a);
                            k0[j] = -rstd * sm_gs[g * 2];
                            k2[j] = -rstd * rstd * sm_gs[g * 2 + 1];
                        }
                    }
                    const int64_t base = (static_cast<int64_t>(n) * HW) * C + cv * 8;
                    const int p0 = ch * pix_per_unit, p1 = min(HW, p0 + pix_per_unit);
                    constexpr int U = 2;
                    for (int pp = p0 + pr; pp < p1; pp += U * R) {
                        uint4 ux[U], ud[U], ua[U];
#pragma unroll
                        for (int k = 0; k < U; ++k) {
                            const bool in = (pp + k * R) < p1;
                            const int64_t off = base + static_cast<int64_t>(pp + k * R) * C;
                            ux[k] = in ? (hints ? ldg16_hint(x + off, pol_drop) : ldg16(x + off)) : make_uint4(0, 0, 0, 0);
                            ud[k] = in ? (hints ? ldg16_hint(dy + off, pol_drop) : ldg16(dy + off)) : make_uint4(0, 0, 0, 0);
                            if (ADD) ua[k] = in ? ldg16(add + off) : make_uint4(0, 0, 0, 0);
                        }
#pragma unroll
                        for (int k = 0; k < U; ++k) {
                            if ((pp + k * R) >= p1) break;
                            float f[8], d[8], r8[8];
                            cvt8(ux[k], f);
                            cvt8(ud[k], d);
                            if (ADD) cvt8(ua[k], r8);
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                float du2 = 2.f * d[j];
                                if (SILU) {
                                    const float h = fmaf(f[j], a2[j], b2[j]);
                                    const float t = tanh_approx(h);
                                    const float r = fmaf(-h, t, h + 1.f);
                                    du2 = d[j] * fmaf(t, r, r);
                                }
                                float v = fmaf(du2, a2[j], fmaf(f[j] - mean[j], k2[j], k0[j]));
                                if (ADD) v += r8[j];
                                f[j] = v;
                                csum[j] += __bfloat162float(__float2bfloat16(v));
                            }
                            uint4 o;
                            o.x = pack_bf16x2(f[0], f[1]);
                            o.y = pack_bf16x2(f[2], f[3]);
                            o.z = pack_bf16x2(f[4], f[5]);
                            o.w = pack_bf16x2(f[6], f[7]);
                            __nv_bfloat16* op = dx + base + static_cast<int64_t>(pp + k * R) * C;
                            if (hints) stg16_hint(op, o, pol_drop); else *reinterpret_cast<uint4*>(op) = o;
                        }
                    }
                }
                __syncthreads();
            }
        }
    }
    // dx column sums (= bias gradient of the conv that produced x): once per CTA
    if (colsum) {
        for (int i = threadIdx.x; i < C; i += blockDim.x) sm[i] = 0.f;
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 8; ++j) atomicAdd(&sm[cv * 8 + j], csum[j]);
        __syncthreads();
        for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(&colsum[i], sm[i]);
    }
}

// dgamma[c] = sum_n cs[n][c][1], dbeta[c] = sum_n cs[n][c][0]
__global__ void gn_bwd_param_grads_kernel(const float* __restrict__ cs, float* __restrict__ dgamma,
                                          float* __restrict__ dbeta, int N, int C) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= C) return;
    float a = 0.f, b = 0.f;
    for (int n = 0; n < N; ++n) {
        a += cs[(static_cast<int64_t>(n) * C + i) * 2];
        b += cs[(static_cast<int64_t>(n) * C + i) * 2 + 1];
    }
    dbeta[i] = a;
    dgamma[i] = b;
}

// ------------------------------------------------------------------ nearest 2x up-sampling
__global__ void upsample2x_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int N, int H,
                                  int W, int C) {
    const int V = C >> 3;
    const int64_t total = static_cast<int64_t>(N) * H * W * V;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int cv = static_cast<int>(i % V);
        const int64_t pix = i / V;
        const int w = static_cast<int>(pix % W);
        const int h = static_cast<int>((pix / W) % H);
        const int64_t n = pix / (static_cast<int64_t>(W) * H);
        const uint4 u = *reinterpret_cast<const uint4*>(x + pix * C + cv * 8);
        __nv_bfloat16* o = y + ((n * 2 * H + 2 * h) * 2 * W + 2 * w) * C + cv * 8;
        *reinterpret_cast<uint4*>(o) = u;
        *reinterpret_cast<uint4*>(o + C) = u;
        *reinterpret_cast<uint4*>(o + static_cast<int64_t>(2) * W * C) = u;
        *reinterpret_cast<uint4*>(o + static_cast<int64_t>(2) * W * C + C) = u;
    }
}

// dx[n,h,w,c] = sum of the 2x2 block of dy (fp32 add, one rounding)
__global__ void upsample2x_bwd_kernel(const __nv_bfloat16* __restrict__ dy, __nv_bfloat16* __restrict__ dx, int N,
                                      int H, int W, int C) {
    const int V = C >> 3;
    const int64_t total = static_cast<int64_t>(N) * H * W * V;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int cv = static_cast<int>(i % V);
        const int64_t pix = i / V;
        const int w = static_cast<int>(pix % W);
        const int h = static_cast<int>((pix / W) % H);
        const int64_t n = pix / (static_cast<int64_t>(W) * H);
        const __nv_bfloat16* s = dy + ((n * 2 * H + 2 * h) * 2 * W + 2 * w) * C + cv * 8;
        float a[8], b[8], c[8], d[8];
        load8(s, a);
        load8(s + C, b);
        load8(s + static_cast<int64_t>(2) * W * C, c);
        load8(s + static_cast<int64_t>(2) * W * C + C, d);
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = (a[j] + b[j]) + (c[j] + d[j]);
        store8(dx + pix * C + cv * 8, a);
    }
}

// ------------------------------------------------------------------ column sums (bias gradient)
__global__ void colsum_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ out /* [C] zeroed */,
                              int64_t P, int C, int pix_per_chunk) {
    extern __shared__ float sm[];  // [C]
    const int V = C >> 3, R = blockDim.x / V;
    const int cv = threadIdx.x % V, pr = threadIdx.x / V;
    for (int i = threadIdx.x; i < C; i += blockDim.x) sm[i] = 0.f;
    __syncthreads();
    if (pr < R) {
        float s[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) s[j] = 0.f;
        const int64_t p0 = static_cast<int64_t>(blockIdx.x) * pix_per_chunk;
        const int64_t p1 = min(P, p0 + pix_per_chunk);
        int64_t p = p0 + pr;
        for (; p + 3 * R < p1; p += 4 * R) {
            uint4 u[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) u[k] = ldg16(x + (p + k * R) * C + cv * 8);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float f[8];
                cvt8(u[k], f);
#pragma unroll
                for (int j = 0; j < 8; ++j) s[j] += f[j];
            }
        }
        for (; p < p1; p += R) {
            float f[8];
            load8(x + p * C + cv * 8, f);
#pragma unroll
            for (int j = 0; j < 8; ++j) s[j] += f[j];
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) atomicAdd(&sm[cv * 8 + j], s[j]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(&out[i], sm[i]);
}

// ------------------------------------------------------------------ wgrad split reduction
// grad[co][ci][tap_src] (+)= sum_s partial[s][co][slot*C64 + ci]   (OIHW fp32; slot -> tap via tapmap;
// several slots may map to the same tap when weights were folded)
__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ grad, int ksplit, int Cout,
                                    int CoutPad, int Cin, int T, int nslots, int C64,
                                    const int* __restrict__ tapmap, int accumulate) {
    const int64_t total = static_cast<int64_t>(Cout) * Cin * T;
    const int64_t ld = static_cast<int64_t>(nslots) * C64;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        // ci fastest: coalesced reads of the (ksplit x larger) partial buffer, strided 4-byte writes
        const int ci = static_cast<int>(i % Cin);
        const int tap = static_cast<int>((i / Cin) % T);
        const int co = static_cast<int>(i / (static_cast<int64_t>(T) * Cin));
        float acc = 0.f;
        bool any = false;
        for (int slot = 0; slot < nslots; ++slot) {
            if (tapmap[slot] != tap) continue;
            any = true;
            for (int s = 0; s < ksplit; ++s)
                acc += partial[(static_cast<int64_t>(s) * CoutPad + co) * ld + static_cast<int64_t>(slot) * C64 + ci];
        }
        const int64_t o = (static_cast<int64_t>(co) * Cin + ci) * T + tap;
        if (any || !accumulate) grad[o] = accumulate ? grad[o] + acc : acc;
    }
}

// Folded variant: slot s contributes to every tap in tapmask[s] (transpose of pack_weights_fold).
__global__ void wgrad_reduce_fold_kernel(const float* __restrict__ partial, float* __restrict__ grad, int ksplit,
                                         int Cout, int CoutPad, int Cin, int T, int nslots, int C64,
                                         const int* __restrict__ tapmask) {
    const int64_t total = static_cast<int64_t>(Cout) * Cin * T;
    const int64_t ld = static_cast<int64_t>(nslots) * C64;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int ci = static_cast<int>(i % Cin);
        const int tap = static_cast<int>((i / Cin) % T);
        const int co = static_cast<int>(i / (static_cast<int64_t>(T) * Cin));
        float acc = 0.f;
        for (int slot = 0; slot < nslots; ++slot) {
            if (!((tapmask[slot] >> tap) & 1)) continue;
            for (int s = 0; s < ksplit; ++s)
                acc += partial[(static_cast<int64_t>(s) * CoutPad + co) * ld + static_cast<int64_t>(slot) * C64 + ci];
        }
        grad[(static_cast<int64_t>(co) * Cin + ci) * T + tap] = acc;
    }
}


// ------------------------------------------------------------------ wavelet front-end (utils.py:229-247)
// y[n, ho, wo, c*4 + band] = sum_{i,j < 6} x[n, c, 2ho + i - 2, 2wo + j - 2] * filt[band][i][j]   (zero outside):
// the reference's F.pad(2) + grouped 6x6 stride-2 conv, fused with the NCHW fp32 -> NHWC bf16 layout conversion the
// encoder's first conv needs. One thread per output pixel; the 6x6 windows of neighbouring pixels overlap 9x, which the
// L1/L2 absorbs (the whole input batch is a few tens of MB).
__global__ void wavelet_fwd_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                   const float* __restrict__ filt, int N, int C, int H, int W, int Cpad) {
    __shared__ float f[4 * 36];
    for (int i = threadIdx.x; i < 144; i += blockDim.x) f[i] = filt[i];
    __syncthreads();
    const int Ho = H / 2, Wo = W / 2;
    const int64_t total = static_cast<int64_t>(N) * Ho * Wo;
    for (int64_t p = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; p < total;
         p += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int wo = static_cast<int>(p % Wo), ho = static_cast<int>((p / Wo) % Ho);
        const int64_t n = p / (static_cast<int64_t>(Wo) * Ho);
        __nv_bfloat16* yp = y + p * Cpad;
        for (int c = 0; c < C; ++c) {
            const float* xp = x + (n * C + c) * static_cast<int64_t>(H) * W;
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                const int h = 2 * ho + i - 2;
                if (h < 0 || h >= H) continue;
#pragma unroll
                for (int j = 0; j < 6; ++j) {
                    const int w = 2 * wo + j - 2;
                    if (w < 0 || w >= W) continue;
                    const float v = __ldg(xp + static_cast<int64_t>(h) * W + w);
#pragma unroll
                    for (int b = 0; b < 4; ++b) acc[b] = fmaf(v, f[b * 36 + i * 6 + j], acc[b]);
                }
            }
#pragma unroll
            for (int b = 0; b < 4; ++b) yp[c * 4 + b] = __float2bfloat16(acc[b]);
        }
        for (int c = 4 * C; c < Cpad; ++c) yp[c] = __float2bfloat16(0.f);
    }
}

static inline int gs_blocks(int64_t total, int threads) {
    int64_t b = (total + threads - 1) / threads;
    const int64_t cap = static_cast<int64_t>(num_sms() > 0 ? num_sms() : 148) * 16;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return static_cast<int>(b);
}

// thread-block shape for the "fixed channel vector" kernels
static inline int cv_threads(int C) {
    const int V = C / 8;
    int t = (256 / V) * V;
    if (t < V) t = V;
    return t;
}
// Grid for the fixed-channel-vector kernels: (chunks, N) blocks of cv_threads(C) threads. The block count is made a
// whole number of waves for the kernel's actual occupancy (these kernels are HBM bound; a 60 %-full second wave was
// costing ~20 %), with at least 4 row-iterations of work per thread.
template <typename Kernel>
static inline void cv_grid(int HW, int C, int N, Kernel kernel, size_t smem, int& chunks, int& pix_per_chunk) {
    const int V = C / 8;
    const int T = cv_threads(C);
    const int R = T / V > 0 ? T / V : 1;
    int bpsm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bpsm, kernel, T, smem) != cudaSuccess || bpsm < 1) bpsm = 4;
    const int concurrent = (num_sms() > 0 ? num_sms() : 148) * bpsm;
    const int min_ppc = R * 4;
    int best_chunks = 1;
    for (int waves = 1; waves <= 8; ++waves) {
        int c = (concurrent * waves) / N;  // chunks per image so that N*c <= waves*concurrent
        if (c < 1) c = 1;
        const int ppc = (HW + c - 1) / c;
        best_chunks = c;
        if (ppc <= 2048 || ppc <= min_ppc) break;  // small enough pieces: stop adding waves
    }
    int ppc = (HW + best_chunks - 1) / best_chunks;
    if (ppc < min_ppc) ppc = min_ppc;
    ppc = ((ppc + R - 1) / R) * R;
    pix_per_chunk = ppc;
    chunks = (HW + ppc - 1) / ppc;
}

}  // namespace vqb

using namespace vqb;

extern "C" {

int vqb_pack_weights(const float* w, void* out, int Cout, int Cin, int T, int nslots, const int* tapmap_dev,
                     int transpose, int Kpad, void* stream) {
    VQB_CHECK(w && out && tapmap_dev, "vqb_pack_weights: null pointer");
    VQB_CHECK(Kpad % 8 == 0 && Kpad >= (transpose ? Cout : Cin), "vqb_pack_weights: bad Kpad=%d", Kpad);
    const int R = transpose ? Cin : Cout;
    const int64_t total = static_cast<int64_t>(R) * nslots * Kpad;
    pack_weights_kernel<<<gs_blocks(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        w, static_cast<__nv_bfloat16*>(out), Cout, Cin, T, nslots, tapmap_dev, transpose, Kpad, 0);
    VQB_CUDA(cudaGetLastError());
    count_launch();
    return VQB_OK;
}

int vqb_pack_weights_fold(const float* w, void* out, int Cout, int Cin, int T, int nslots, const int* tapmask_dev,
                          int transpose, int Kpad, void* stream) {
    VQB_CHECK(w && out && tapmask_dev, "vqb_pack_weights_fold: null pointer");
    VQB_CHECK(Kpad % 8 == 0 && Kpad >= (transpose ? Cout : Cin) && T <= 31, "vqb_pack_weights_fold: bad Kpad/T");
    const int R = transpose ? Cin : Cout;
    const int64_t total = static_cast<int64_t>(R) * nslots * Kpad;
    pack_weights_fold_kernel<<<gs_blocks(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        w, static_cast<__nv_bfloat16*>(out), Cout, Cin, T, nslots, tapmask_dev, transpose, Kpad);
    VQB_CUDA(cudaGetLastError());
    count_launch();
    return VQB_OK;
}

int vqb_wgrad_reduce_fold(const float* partial, float* grad, int ksplit, int Cout, int CoutPad, int Cin, int T,
                          int nslots, int C64, const int* tapmask_dev, void* stream) {
    VQB_CHECK(partial && grad && tapmask_dev && T <= 31, "vqb_wgrad_reduce_fold: bad arguments");
    const int64_t total = static_cast<int64_t>(Cout) * Cin * T;
    wgrad_reduce_fold_kernel<<<gs_blocks(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        partial, grad, ksplit, Cout, CoutPad, Cin, T, nslots, C64, tapmask_dev);
    VQB_CUDA(cudaGetLastError());
    count_launch();
    return VQB_OK;
}

int vqb_nchw_to_nhwc(const float* x, void* y, int N, int C, int H, int W, int Cpad, const float* shift,
                     const float* inv_scale, void* stream) {
    VQB_CHECK(x && y && Cpad % 8 == 0 && Cpad >= C, "vqb_nchw_to_nhwc: bad arguments");
    const int64_t total = static_cast<int64_t>(N) * H * W;
    nchw_to_nhwc_kernel<<<gs_blocks(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        x, static_cast<__nv_bfloat16*>(y), N, C, H * W, Cpad, shift, inv_scale, W, 0);
    VQB_CUDA(cudaGetLastError());
    count_launch();
    return VQB_OK;
}

// Same, into the interior of a PRE-ZEROED [N][H+2pad][W+2pad][Cpad] buffer (zero frame for the "fat pixel" first-layer conv)
int vqb_nchw_to_nhwc_pad(const float* x, void* y, int N, int C, int H, int W, int Cpad, int pad, const float* shift,
                         const float* inv_scale, void* stream) {
    VQB_CHECK(x && y && Cpad % 8 == 0 && Cpad >= C && pad >= 0, "vqb_nchw_to_nhwc_pad: bad arguments");
    const int64_t total = static_cast<int64_t>(N) * H * W;
    nchw_to_nhwc_kernel<<<gs_blocks(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        x, static_cast<__nv_bfloat16*>(y), N, C, H * W, Cpad, shift, inv_scale, W, pad);
    VQB_CUDA(cudaGetLastError());
    count_launch();
    return VQB_OK;
}

// inverse: reads the interior of a padded NHWC buffer
int vqb_nhwc_to_nchw_pad(const void* g, float* gx, int N, int C, int H, int W, int Cpad, int pad,
                         const float* inv_scale, void* stream) {
    VQB_CHECK(g && gx && Cpad % 8 == 0 && Cpad >= C && pad >= 0, "vqb_nhwc_to_nchw_pad: bad arguments");
    const int64_t total = static_cast<int64_t>(N) * H * W;
    nhwc_to_nchw_kernel<<<gs_blocks(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(g), gx, N, C, H * W, Cpad, inv_scale, W, pad);
    VQB_CUDA(cudaGetLastError());
    count_launch();
    return VQB_OK;
}

int vqb_nhwc_to_nchw(const void* g, float* gx, int N, int C, int H, int W, int Cpad, const float* inv_scale,
                     void* stream) {
    VQB_CHECK(g && gx && Cpad % 8 == 0 && Cpad >= C, "vqb_nhwc_to_nchw: bad arguments");
    const int64_t total = static_cast<int64_t>(N) * H * W;
    nhwc_to_nchw_kernel<<<gs_blocks(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(g), gx, N, C, H * W, Cpad, inv_scale, W, 0);
    VQB_CUDA(cudaGetLastError());
    count_launch();
    return VQB_OK;
}

// GroupNorm(+SiLU) forward. ws: >= N*C*2 doubles (zeroed here); mr: [N][G][2] floats (mean, rstd) kept for backward.
int vqb_gn_silu_fwd(const void* x, void* y, const float* gamma, const float* beta, float* mr, double* ws, int N,
                    int HW, int C, int G, float eps, int silu, void* stream) {
    VQB_CHECK(x && y && gamma && beta && mr && ws, "vqb_gn_silu_fwd: null pointer");
    VQB_CHECK(C % 8 == 0 && C % G == 0 && C <= 2048, "vqb_gn_silu_fwd: C=%d G=%d unsupported", C, G);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    VQB_CUDA(cudaMemsetAsync(ws, 0, sizeof(double) * 2 * N * C, st));
    int chunks, ppc;
    const int T = cv_threads(C);
    cv_grid(HW, C, N, gn_stats_kernel, 2 * C * sizeof(float), chunks, ppc);
    gn_stats_kernel<<<dim3(chunks, N), T, 2 * C * sizeof(float), st>>>(static_cast<const __nv_bfloat16*>(x), ws, HW, C,
                                                                        ppc);
    gn_finalize_kernel<<<(N * G + 127) / 128, 128, 0, st>>>(ws, mr, N, C, G, HW, eps);
    cv_grid(HW, C, N, gn_apply_kernel, 0, chunks, ppc);
    gn_apply_kernel<<<dim3(chunks, N), T, 0, st>>>(static_cast<const __nv_bfloat16*>(x),
                                                   static_cast<__nv_bfloat16*>(y), mr, gamma, beta, HW, C, G, ppc,
                                                   silu);
    VQB_CUDA(cudaGetLastError());
    count_launch(3);
    return VQB_OK;
}

// GroupNorm(+SiLU) forward when the per-(n, channel) sums [N][C][2] (sum, sum of squares; fp32) were already produced by
// the convolution that wrote x (vqb_conv_gemm with VQB_EPI_STATS): finalise + one apply pass, no statistics pass.
int vqb_gn_silu_fwd_pre(const void* x, void* y, const float* gamma, const float* beta, float* mr, const float* chsums,
                        int N, int HW, int C, int G, float eps, int silu, void* stream) {
    VQB_CHECK(x && y && gamma && beta && mr && chsums, "vqb_gn_silu_fwd_pre: null pointer");
    VQB_CHECK(C % 8 == 0 && C % G == 0 && C <= 2048, "vqb_gn_silu_fwd_pre: C=%d G=%d unsupported", C, G);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    gn_finalize_f32_kernel<<<(N * G + 127) / 128, 128, 0, st>>>(chsums, mr, N, C, G, HW, eps);
    int chunks, ppc;
    const int T = cv_threads(C);
    cv_grid(HW, C, N, gn_apply_kernel, 0, chunks, ppc);
    gn_apply_kernel<<<dim3(chunks, N), T, 0, st>>>(static_cast<const __nv_bfloat16*>(x),
                                                   static_cast<__nv_bfloat16*>(y), mr, gamma, beta, HW, C, G, ppc,
                                                   silu);
    VQB_CUDA(cudaGetLastError());
    count_launch(2);
    return VQB_OK;
}

// GroupNorm(+SiLU) backward. ws: >= N*C*2 + N*G*2 floats. dx may alias dy. add (optional) is summed into dx.
// dx_colsum (optional, [C] fp32, overwritten): per-channel sums of dx over all N*HW pixels, i.e. the bias gradient of the
// convolution whose output this GroupNorm normalised, produced in the same pass instead of by vqb_colsum.
static int gn_silu_bwd_impl(const void* x, const void* dy, const void* add, void* dx, const float* gamma,
                           const float* beta, const float* mr, const float* cs_pre, float* dgamma, float* dbeta, float* ws,
                           int N, int HW, int C, int G, int silu, float* dx_colsum, void* stream) {
    VQB_CHECK(x && dy && dx && gamma && beta && mr && dgamma && dbeta && ws, "vqb_gn_silu_bwd: null pointer");
    VQB_CHECK(C % 8 == 0 && C % G == 0 && C <= 2048, "vqb_gn_silu_bwd: C=%d G=%d unsupported", C, G);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int chunks, ppc;
    const int T = cv_threads(C);
    const float* cs = cs_pre;
    float* gsum = ws;
    // Persistent L2-pipelined form (reduce + apply in one launch, x / dy read from HBM once): used when one sample's
    // x + dy (+ add) comfortably fits the L2 next to the following sample's, and there is enough work to pipeline.
    // VQB_GN_BWD_PERSISTENT=1 opts into the persistent form. Measured (tools/gn_bwd_bench.py, N=32, profiles/
    // r02_gn_bwd_variants.txt): 1.5-2x SLOWER than the two-kernel form on every shape (e.g. 128 ch @ 256^2: 1010 us vs
    // 682 us) — two 128-register CTAs per SM keep too few loads in flight and every unit pays barrier + fence + atomics —
    // so the default stays the two-kernel form.
    static const int gnp_mode = [] {
        const char* e = getenv("VQB_GN_BWD_PERSISTENT");
        return e ? atoi(e) : 0;
    }();
    static const int gnp_depth = [] { const char* e = getenv("VQB_GNP_DEPTH"); return e ? atoi(e) : 1; }();
    static const int gnp_hints = [] { const char* e = getenv("VQB_GNP_HINTS"); return e ? atoi(e) : 1; }();
    static const int gnp_mb = [] { const char* e = getenv("VQB_GNP_MB"); return e ? atoi(e) : 36; }();
    const int64_t sample_bytes = static_cast<int64_t>(HW) * C * 2 * (add ? 3 : 2);
    if (!cs_pre && gnp_mode && T <= 256 && T % (C / 8) == 0 && sample_bytes <= (static_cast<int64_t>(gnp_mb) << 20) &&
        static_cast<int64_t>(N) * HW * C >= (1 << 18)) {
        const int V = C / 8, R = T / V;
        float* csw = ws;                                                   // [N][C][2]
        int* done = reinterpret_cast<int*>(ws + static_cast<int64_t>(N) * C * 2);  // [<= N] (ws has N*G*2 floats there)
        VQB_CUDA(cudaMemsetAsync(csw, 0, sizeof(float) * (2 * N * C + N), st));
        if (dx_colsum) VQB_CUDA(cudaMemsetAsync(dx_colsum, 0, sizeof(float) * C, st));
        const size_t smem = (2 * C + 2 * G) * sizeof(float);
        auto launch = [&](auto kern) -> int {
            int bpsm = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bpsm, kern, T, smem) != cudaSuccess || bpsm < 1)
                return set_error(VQB_ECUDA, "vqb_gn_silu_bwd: occupancy query failed");
            if (bpsm > 2) bpsm = 2;
            const int grid = (num_sms() > 0 ? num_sms() : 148) * bpsm;
            // sample groups of S samples whose x + dy (+ add) fit the L2 budget; each group is cut into ~grid units
            int S = static_cast<int>((static_cast<int64_t>(gnp_mb) << 20) / sample_bytes);
            if (S < 1) S = 1;
            if (S > N) S = N;
            const int min_ppu = R * 4;
            int ups = (grid + S - 1) / S;  // units per sample
            if (ups < 1) ups = 1;
            if (HW / ups < min_ppu) ups = HW / min_ppu > 0 ? HW / min_ppu : 1;
            int ppu = (HW + ups - 1) / ups;
            ppu = ((ppu + R - 1) / R) * R;
            ups = (HW + ppu - 1) / ppu;
            kern<<<grid, T, smem, st>>>(static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(dy),
                                        static_cast<const __nv_bfloat16*>(add), static_cast<__nv_bfloat16*>(dx), mr,
                                        gamma, beta, csw, done, dx_colsum, N, HW, C, G, S, ups, ppu, gnp_depth,
                                        gnp_hints);
            return VQB_OK;
        };
        int rc;
        if (add)
            rc = silu ? launch(gn_bwd_persistent_kernel<true, true>) : launch(gn_bwd_persistent_kernel<true, false>);
        else
            rc = silu ? launch(gn_bwd_persistent_kernel<false, true>) : launch(gn_bwd_persistent_kernel<false, false>);
        if (rc != VQB_OK) return rc;
        gn_bwd_param_grads_kernel<<<(C + 127) / 128, 128, 0, st>>>(csw, dgamma, dbeta, N, C);
        VQB_CUDA(cudaGetLastError());
        count_launch(2);
        return VQB_OK;
    }
    if (!cs_pre) {  // statistics pass (skipped when the consumer conv's data-gradient epilogue produced them)
        float* csw = ws;
        gsum = ws + static_cast<int64_t>(N) * C * 2;
        VQB_CUDA(cudaMemsetAsync(csw, 0, sizeof(float) * 2 * N * C, st));
        static const int red_u = [] { const char* e = getenv("VQB_GN_RED_U"); return e ? atoi(e) : 4; }();
        auto launch_red = [&](auto kern) {
            cv_grid(HW, C, N, kern, 2 * C * sizeof(float), chunks, ppc);
            kern<<<dim3(chunks, N), T, 2 * C * sizeof(float), st>>>(
                static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(dy), mr, gamma, beta, csw, HW,
                C, G, ppc, silu);
        };
        if (red_u == 6) launch_red(gn_bwd_reduce_kernel<6>);
        else if (red_u == 8) launch_red(gn_bwd_reduce_kernel<8>);
        else if (red_u == 2) launch_red(gn_bwd_reduce_kernel<2>);
        else launch_red(gn_bwd_reduce_kernel<4>);
        cs = csw;
        count_launch();
    }
    const int fin = (N * G > C ? N * G : C);
    gn_bwd_finalize_kernel<<<(fin + 127) / 128, 128, 0, st>>>(cs, gamma, gsum, dgamma, dbeta, N, C, G, HW);
    const size_t cs_smem = dx_colsum ? C * sizeof(float) : 0;
    if (dx_colsum) VQB_CUDA(cudaMemsetAsync(dx_colsum, 0, sizeof(float) * C, st));
    if (add) {
        cv_grid(HW, C, N, gn_bwd_apply_kernel<true, 2>, cs_smem, chunks, ppc);
        gn_bwd_apply_kernel<true, 2><<<dim3(chunks, N), T, cs_smem, st>>>(
            static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(dy),
            static_cast<const __nv_bfloat16*>(add), static_cast<__nv_bfloat16*>(dx), mr, gsum, gamma, beta, HW, C, G,
            ppc, silu, dx_colsum);
    } else {
        cv_grid(HW, C, N, gn_bwd_apply_kernel<false, 3>, cs_smem, chunks, ppc);
        gn_bwd_apply_kernel<false, 3><<<dim3(chunks, N), T, cs_smem, st>>>(
            static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(dy), nullptr,
            static_cast<__nv_bfloat16*>(dx), mr, gsum, gamma, beta, HW, C, G, ppc, silu, dx_colsum);
    }
    VQB_CUDA(cudaGetLastError());
    count_launch(2);
    return VQB_OK;
}

int vqb_gn_silu_bwd(const void* x, const void* dy, const void* add, void* dx, const float* gamma, const float* beta,
                    const float* mr, float* dgamma, float* dbeta, float* ws, int N, int HW, int C, int G, int silu,
                    float* dx_colsum, void* stream) {
    return gn_silu_bwd_impl(x, dy, add, dx, gamma, beta, mr, nullptr, dgamma, dbeta, ws, N, HW, C, G, silu, dx_colsum,
                            stream);
}

int vqb_gn_silu_bwd_pre(const void* x, const void* dy, const void* add, void* dx, const float* gamma, const float* beta,
                        const float* mr, const float* cs, float* dgamma, float* dbeta, float* ws, int N, int HW, int C,
                        int G, int silu, float* dx_colsum, void* stream) {
    VQB_CHECK(cs, "vqb_gn_silu_bwd_pre: null cs");
    return gn_silu_bwd_impl(x, dy, add, dx, gamma, beta, mr, cs, dgamma, dbeta, ws, N, HW, C, G, silu, dx_colsum, stream);
}

int vqb_wavelet_fwd(const float* x, void* y, const float* filt, int N, int C, int H, int W, int Cpad, void* stream) {
    VQB_CHECK(x && y && filt && H % 2 == 0 && W % 2 == 0 && Cpad % 8 == 0 && Cpad >= 4 * C,
              "vqb_wavelet_fwd: bad arguments (H, W even; Cpad >= 4C, multiple of 8)");
    const int64_t total = static_cast<int64_t>(N) * (H / 2) * (W / 2);
    wavelet_fwd_kernel<<<gs_blocks(total, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        x, static_cast<__nv_bfloat16*>(y), filt, N, C, H, W, Cpad);
    VQB_CUDA(cudaGetLastError());
    count_launch();
    return VQB_OK;
}

int vqb_upsample2x_fwd(const void* x, void* y, int N, int H, int W, int C, void* stream) {
    VQB_CHECK(x && y && C % 8 == 0, "vqb_upsample2x_fwd: bad arguments");
    const int64_t total = static_cast<int64_t>(N) * H * W * (C / 8);
    upsample2x_kernel<<<gs_blocks(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(y), N, H, W, C);
    VQB_CUDA(cudaGetLastError());
    count_launch();
    return VQB_OK;
}

int vqb_upsample2x_bwd(const void* dy, void* dx, int N, int H, int W, int C, void* stream) {
    VQB_CHECK(dy && dx && C % 8 == 0, "vqb_upsample2x_bwd: bad arguments");
    const int64_t total = static_cast<int64_t>(N) * H * W * (C / 8);
    upsample2x_bwd_kernel<<<gs_blocks(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(dy), static_cast<__nv_bfloat16*>(dx), N, H, W, C);
    VQB_CUDA(cudaGetLastError());
    count_launch();
    return VQB_OK;
}

// out[c] = sum over P pixels of x[p][c]  (bias gradient). out is overwritten.
int vqb_colsum(const void* x, float* out, int64_t P, int C, void* stream) {
    VQB_CHECK(x && out && C % 8 == 0 && C <= 2048, "vqb_colsum: bad arguments");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    VQB_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * C, st));
    const int V = C / 8, T = cv_threads(C), R = T / V;
    int bpsm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bpsm, colsum_kernel, T, C * sizeof(float)) != cudaSuccess || bpsm < 1)
        bpsm = 4;
    const int64_t concurrent = static_cast<int64_t>(num_sms() > 0 ? num_sms() : 148) * bpsm;
    int64_t nblk = concurrent;
    while (nblk < concurrent * 8 && (P + nblk - 1) / nblk > 2048) nblk += concurrent;
    int64_t ppc = (P + nblk - 1) / nblk;
    if (ppc < R * 4) ppc = R * 4;
    ppc = ((ppc + R - 1) / R) * R;
    const int chunks = static_cast<int>((P + ppc - 1) / ppc);
    colsum_kernel<<<chunks, T, C * sizeof(float), st>>>(static_cast<const __nv_bfloat16*>(x), out, P, C,
                                                        static_cast<int>(ppc));
    VQB_CUDA(cudaGetLastError());
    count_launch();
    return VQB_OK;
}

int vqb_wgrad_reduce(const float* partial, float* grad, int ksplit, int Cout, int CoutPad, int Cin, int T, int nslots,
                     int C64, const int* tapmap_dev, int accumulate, void* stream) {
    VQB_CHECK(partial && grad && tapmap_dev, "vqb_wgrad_reduce: null pointer");
    const int64_t total = static_cast<int64_t>(Cout) * Cin * T;
    wgrad_reduce_kernel<<<gs_blocks(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        partial, grad, ksplit, Cout, CoutPad, Cin, T, nslots, C64, tapmap_dev, accumulate);
    VQB_CUDA(cudaGetLastError());
    count_launch();
    return VQB_OK;
}

}  // extern "C"
