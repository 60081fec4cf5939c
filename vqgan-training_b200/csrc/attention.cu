// Multi-head self-attention core of AttnBlock (ae.py:74-93): softmax(q k^T / sqrt(64)) v per head of 64 channels,
// flash-attention style (online softmax, no T x T matrix in HBM), warp-level tensor-core MMA
// (mma.sync.m16n8k16 bf16 -> fp32). The 1x1 qkv / proj_out convolutions and the GroupNorm around it run on the
// tcgen05 conv / GN kernels; this file is only the [T x T] part: T = (H/8)(W/8) = 1024 tokens at 256^2, 8 heads at
// C = 512, 4.3 GFLOP per image and block — latency/occupancy bound, not worth a TMEM pipeline.
//
// Layout: qkv [N][T][3C] bf16 (channel blocks q | k | v; head h owns channels h*64..h*64+63 of each block, the
// "b (h d) x y -> b h (x y) d" rearrange of ae.py:79-89 is pure addressing), out [N][T][C] bf16, lse [N][heads][T] fp32.
//
// Backward: D = rowsum(dO * O); one kernel owns key tiles and produces dK, dV; one owns query tiles and produces dQ.
// Both recompute P from q, k and the saved log-sum-exp.
#include "common.cuh"
#include "ptx.cuh"

namespace vqb {

constexpr int kHD = 64;   // head dim
constexpr int kTQ = 64;   // rows per block (4 warps x 16)
constexpr int kLD = 72;   // smem row pitch in bf16 (144 B: conflict-free 32-bit fragment loads)

__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// 64 x 64 bf16 tile: rows row0.. of a [T][ld] matrix (column offset already applied to src) -> dst[64][kLD]; rows >= T zero.
__device__ __forceinline__ void load_tile(__nv_bfloat16* dst, const __nv_bfloat16* src, int64_t ld, int row0, int T) {
    for (int i = threadIdx.x; i < 64 * 8; i += blockDim.x) {
        const int r = i >> 3, v = i & 7;
        uint4 u = make_uint4(0, 0, 0, 0);
        if (row0 + r < T) u = __ldg(reinterpret_cast<const uint4*>(src + static_cast<int64_t>(row0 + r) * ld + v * 8));
        *reinterpret_cast<uint4*>(dst + r * kLD + v * 8) = u;
    }
}
// same tile stored transposed: dst[col][row]
__device__ __forceinline__ void load_tile_t(__nv_bfloat16* dst, const __nv_bfloat16* src, int64_t ld, int row0, int T) {
    for (int i = threadIdx.x; i < 64 * 8; i += blockDim.x) {
        const int r = i >> 3, v = i & 7;
        uint4 u = make_uint4(0, 0, 0, 0);
        if (row0 + r < T) u = __ldg(reinterpret_cast<const uint4*>(src + static_cast<int64_t>(row0 + r) * ld + v * 8));
        const __nv_bfloat16* e = reinterpret_cast<const __nv_bfloat16*>(&u);
#pragma unroll
        for (int j = 0; j < 8; ++j) dst[(v * 8 + j) * kLD + r] = e[j];
    }
}
// A fragments (16 rows of this warp x 64 cols) of a [64][kLD] smem tile
__device__ __forceinline__ void load_a_frags(const __nv_bfloat16* s, int warp, int lane, uint32_t (&a)[4][4]) {
    const int r = warp * 16 + (lane >> 2), c = (lane & 3) * 2;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
        a[ks][0] = *reinterpret_cast<const uint32_t*>(s + r * kLD + ks * 16 + c);
        a[ks][1] = *reinterpret_cast<const uint32_t*>(s + (r + 8) * kLD + ks * 16 + c);
        a[ks][2] = *reinterpret_cast<const uint32_t*>(s + r * kLD + ks * 16 + c + 8);
        a[ks][3] = *reinterpret_cast<const uint32_t*>(s + (r + 8) * kLD + ks * 16 + c + 8);
    }
}
// acc[nt] += A(16 x 64) * B where B[k][n] = s[n][k] (s is a [64 n][kLD] smem tile, k contiguous)
__device__ __forceinline__ void mma_a_bT(float (&acc)[8][4], const uint32_t (&a)[4][4], const __nv_bfloat16* s, int lane) {
    const int n = lane >> 2, c = (lane & 3) * 2;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks)
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const uint32_t b0 = *reinterpret_cast<const uint32_t*>(s + (nt * 8 + n) * kLD + ks * 16 + c);
            const uint32_t b1 = *reinterpret_cast<const uint32_t*>(s + (nt * 8 + n) * kLD + ks * 16 + c + 8);
            mma16816(acc[nt], a[ks], b0, b1);
        }
}
// accumulator (16 x 64 fp32, 8 n-tiles) -> bf16 A fragments over k = the 64 columns
__device__ __forceinline__ void acc_to_a(const float (&p)[8][4], uint32_t (&a)[4][4]) {
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
        a[kk][0] = pack_bf16x2(p[2 * kk][0], p[2 * kk][1]);
        a[kk][1] = pack_bf16x2(p[2 * kk][2], p[2 * kk][3]);
        a[kk][2] = pack_bf16x2(p[2 * kk + 1][0], p[2 * kk + 1][1]);
        a[kk][3] = pack_bf16x2(p[2 * kk + 1][2], p[2 * kk + 1][3]);
    }
}

// ------------------------------------------------------------------------------------------------ forward
__global__ void __launch_bounds__(128) attn_fwd_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                       __nv_bfloat16* __restrict__ out, float* __restrict__ lse, int T,
                                                       int C, float scale) {
    __shared__ __align__(16) __nv_bfloat16 sQ[64 * kLD];
    __shared__ __align__(16) __nv_bfloat16 sK[64 * kLD];
    __shared__ __align__(16) __nv_bfloat16 sVt[64 * kLD];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * kTQ, h = blockIdx.y, n = blockIdx.z, heads = gridDim.y;
    const int64_t ld = 3 * static_cast<int64_t>(C);
    const __nv_bfloat16* base = qkv + static_cast<int64_t>(n) * T * ld;
    load_tile(sQ, base + h * kHD, ld, q0, T);
    __syncthreads();
    uint32_t qa[4][4];
    load_a_frags(sQ, warp, lane, qa);
    float m_i[2] = {-INFINITY, -INFINITY}, l_i[2] = {0.f, 0.f};
    float o[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
    const int nkt = (T + 63) / 64;
    for (int kt = 0; kt < nkt; ++kt) {
        __syncthreads();
        load_tile(sK, base + C + h * kHD, ld, kt * 64, T);
        load_tile_t(sVt, base + 2 * C + h * kHD, ld, kt * 64, T);
        __syncthreads();
        float s[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
        mma_a_bT(s, qa, sK, lane);
        float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int key = kt * 64 + nt * 8 + (lane & 3) * 2 + (j & 1);
                float v = s[nt][j] * scale;
                if (key >= T) v = -INFINITY;
                s[nt][j] = v;
                mx[j >> 1] = fmaxf(mx[j >> 1], v);
            }
        float alpha[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
            const float mnew = fmaxf(m_i[r], mx[r]);
            alpha[r] = (m_i[r] == -INFINITY) ? 0.f : __expf(m_i[r] - mnew);
            m_i[r] = mnew;
        }
        float rs[2] = {0.f, 0.f};
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float pv = (s[nt][j] == -INFINITY) ? 0.f : __expf(s[nt][j] - m_i[j >> 1]);
                s[nt][j] = pv;
                rs[j >> 1] += pv;
            }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            rs[r] += __shfl_xor_sync(0xffffffffu, rs[r], 1);
            rs[r] += __shfl_xor_sync(0xffffffffu, rs[r], 2);
            l_i[r] = l_i[r] * alpha[r] + rs[r];
        }
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
            for (int j = 0; j < 4; ++j) o[nt][j] *= alpha[j >> 1];
        uint32_t pa[4][4];
        acc_to_a(s, pa);
        mma_a_bT(o, pa, sVt, lane);  // B[k=key][n=d] = Vt[d][key]
    }
    const int r0 = q0 + warp * 16 + (lane >> 2);
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int q = r0 + r * 8;
        if (q < T) {
            const float inv = 1.f / l_i[r];
            __nv_bfloat16* op = out + (static_cast<int64_t>(n) * T + q) * C + h * kHD + (lane & 3) * 2;
#pragma unroll
            for (int nt = 0; nt < 8; ++nt)
                *reinterpret_cast<uint32_t*>(op + nt * 8) = pack_bf16x2(o[nt][2 * r] * inv, o[nt][2 * r + 1] * inv);
            if ((lane & 3) == 0) lse[(static_cast<int64_t>(n) * heads + h) * T + q] = m_i[r] + __logf(l_i[r]);
        }
    }
}

// D[n][h][q] = sum_d dO * O
__global__ void attn_bwd_prep_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ dout,
                                     float* __restrict__ dvec, int T, int C, int heads, int64_t total) {
    const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x / 32) + (threadIdx.x >> 5);  // one warp per (n,q,h)
    if (i >= total) return;
    const int lane = threadIdx.x & 31;
    const int h = static_cast<int>(i % heads);
    const int64_t nq = i / heads;
    const int64_t off = nq * C + h * kHD + lane * 2;
    const float2 a = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(o + off));
    const float2 b = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(dout + off));
    float s = a.x * b.x + a.y * b.y;
#pragma unroll
    for (int k = 16; k > 0; k >>= 1) s += __shfl_xor_sync(0xffffffffu, s, k);
    if (lane == 0) {
        const int64_t n = nq / T, q = nq % T;
        dvec[(n * heads + h) * T + q] = s;
    }
}

// ------------------------------------------------------------------------------------------------ backward: dK, dV
__global__ void __launch_bounds__(128) attn_bwd_dkdv_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                            const __nv_bfloat16* __restrict__ dout,
                                                            const float* __restrict__ lse,
                                                            const float* __restrict__ dvec,
                                                            __nv_bfloat16* __restrict__ dqkv, int T, int C, float scale) {
    extern __shared__ __align__(16) uint8_t smem_dyn[];
    __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem_dyn);
    __nv_bfloat16* sQt = sQ + 64 * kLD;
    __nv_bfloat16* sdO = sQt + 64 * kLD;
    __nv_bfloat16* sdOt = sdO + 64 * kLD;
    float* sLse = reinterpret_cast<float*>(sdOt + 64 * kLD);
    float* sD = sLse + 64;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int k0 = blockIdx.x * 64, h = blockIdx.y, n = blockIdx.z, heads = gridDim.y;
    const int64_t ld = 3 * static_cast<int64_t>(C);
    const __nv_bfloat16* base = qkv + static_cast<int64_t>(n) * T * ld;
    const __nv_bfloat16* dob = dout + static_cast<int64_t>(n) * T * C + h * kHD;
    // K and V rows of this warp as A fragments (staged through sQ / sdO once)
    load_tile(sQ, base + C + h * kHD, ld, k0, T);
    load_tile(sdO, base + 2 * C + h * kHD, ld, k0, T);
    __syncthreads();
    uint32_t ka[4][4], va[4][4];
    load_a_frags(sQ, warp, lane, ka);
    load_a_frags(sdO, warp, lane, va);
    float dk[8][4], dv[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dk[i][j] = dv[i][j] = 0.f;
    const int key_r0 = k0 + warp * 16 + (lane >> 2);
    const int nqt = (T + 63) / 64;
    for (int qt = 0; qt < nqt; ++qt) {
        __syncthreads();
        load_tile(sQ, base + h * kHD, ld, qt * 64, T);
        load_tile_t(sQt, base + h * kHD, ld, qt * 64, T);
        load_tile(sdO, dob, C, qt * 64, T);
        load_tile_t(sdOt, dob, C, qt * 64, T);
        if (threadIdx.x < 64) {
            const int q = qt * 64 + threadIdx.x;
            sLse[threadIdx.x] = q < T ? lse[(static_cast<int64_t>(n) * heads + h) * T + q] : 0.f;
            sD[threadIdx.x] = q < T ? dvec[(static_cast<int64_t>(n) * heads + h) * T + q] : 0.f;
        }
        __syncthreads();
        float st[8][4], dp[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) st[i][j] = dp[i][j] = 0.f;
        mma_a_bT(st, ka, sQ, lane);   // S^T[key][q] = sum_d K[key][d] Q[q][d]
        mma_a_bT(dp, va, sdO, lane);  // dP^T[key][q] = sum_d V[key][d] dO[q][d]
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int ql = nt * 8 + (lane & 3) * 2 + (j & 1);
                const int key = key_r0 + (j >> 1) * 8;
                const bool ok = (qt * 64 + ql < T) && (key < T);
                const float pt = ok ? __expf(st[nt][j] * scale - sLse[ql]) : 0.f;
                st[nt][j] = pt;
                dp[nt][j] = pt * (dp[nt][j] - sD[ql]) * scale;
            }
        uint32_t pa[4][4], dsa[4][4];
        acc_to_a(st, pa);
        acc_to_a(dp, dsa);
        mma_a_bT(dv, pa, sdOt, lane);  // dV[key][d] += sum_q P^T[key][q] dO[q][d]
        mma_a_bT(dk, dsa, sQt, lane);  // dK[key][d] += sum_q dS^T[key][q] Q[q][d]
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int key = key_r0 + r * 8;
        if (key < T) {
            __nv_bfloat16* kp = dqkv + (static_cast<int64_t>(n) * T + key) * ld + C + h * kHD + (lane & 3) * 2;
            __nv_bfloat16* vp = kp + C;
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                *reinterpret_cast<uint32_t*>(kp + nt * 8) = pack_bf16x2(dk[nt][2 * r], dk[nt][2 * r + 1]);
                *reinterpret_cast<uint32_t*>(vp + nt * 8) = pack_bf16x2(dv[nt][2 * r], dv[nt][2 * r + 1]);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ backward: dQ
__global__ void __launch_bounds__(128) attn_bwd_dq_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                          const __nv_bfloat16* __restrict__ dout,
                                                          const float* __restrict__ lse, const float* __restrict__ dvec,
                                                          __nv_bfloat16* __restrict__ dqkv, int T, int C, float scale) {
    __shared__ __align__(16) __nv_bfloat16 sK[64 * kLD];
    __shared__ __align__(16) __nv_bfloat16 sKt[64 * kLD];
    __shared__ __align__(16) __nv_bfloat16 sV[64 * kLD];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * 64, h = blockIdx.y, n = blockIdx.z, heads = gridDim.y;
    const int64_t ld = 3 * static_cast<int64_t>(C);
    const __nv_bfloat16* base = qkv + static_cast<int64_t>(n) * T * ld;
    const __nv_bfloat16* dob = dout + static_cast<int64_t>(n) * T * C + h * kHD;
    load_tile(sK, base + h * kHD, ld, q0, T);
    load_tile(sV, dob, C, q0, T);
    __syncthreads();
    uint32_t qa[4][4], doa[4][4];
    load_a_frags(sK, warp, lane, qa);
    load_a_frags(sV, warp, lane, doa);
    const int qr0 = q0 + warp * 16 + (lane >> 2);
    float lse_r[2], d_r[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int q = qr0 + r * 8;
        lse_r[r] = q < T ? lse[(static_cast<int64_t>(n) * heads + h) * T + q] : 0.f;
        d_r[r] = q < T ? dvec[(static_cast<int64_t>(n) * heads + h) * T + q] : 0.f;
    }
    float dq[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dq[i][j] = 0.f;
    const int nkt = (T + 63) / 64;
    for (int kt = 0; kt < nkt; ++kt) {
        __syncthreads();
        load_tile(sK, base + C + h * kHD, ld, kt * 64, T);
        load_tile_t(sKt, base + C + h * kHD, ld, kt * 64, T);
        load_tile(sV, base + 2 * C + h * kHD, ld, kt * 64, T);
        __syncthreads();
        float s[8][4], dp[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) s[i][j] = dp[i][j] = 0.f;
        mma_a_bT(s, qa, sK, lane);    // S[q][key]
        mma_a_bT(dp, doa, sV, lane);  // dP[q][key] = sum_d dO[q][d] V[key][d]
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int key = kt * 64 + nt * 8 + (lane & 3) * 2 + (j & 1);
                const int q = qr0 + (j >> 1) * 8;
                const bool ok = (key < T) && (q < T);
                const float pv = ok ? __expf(s[nt][j] * scale - lse_r[j >> 1]) : 0.f;
                dp[nt][j] = pv * (dp[nt][j] - d_r[j >> 1]) * scale;
            }
        uint32_t dsa[4][4];
        acc_to_a(dp, dsa);
        mma_a_bT(dq, dsa, sKt, lane);  // dQ[q][d] += sum_key dS[q][key] K[key][d]
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int q = qr0 + r * 8;
        if (q < T) {
            __nv_bfloat16* qp = dqkv + (static_cast<int64_t>(n) * T + q) * ld + h * kHD + (lane & 3) * 2;
#pragma unroll
            for (int nt = 0; nt < 8; ++nt)
                *reinterpret_cast<uint32_t*>(qp + nt * 8) = pack_bf16x2(dq[nt][2 * r], dq[nt][2 * r + 1]);
        }
    }
}

}  // namespace vqb

using namespace vqb;

extern "C" {

// out[n][t][h*64+d] = softmax_t'(q.k/8) v ; lse [N][C/64][T] saved for the backward. Replaces
// F.scaled_dot_product_attention + the einops rearranges at ae.py:79-89.
int vqb_attn_fwd(const void* qkv, void* out, float* lse, int N, int T, int C, void* stream) {
    VQB_CHECK(qkv && out && lse, "vqb_attn_fwd: null pointer");
    VQB_CHECK(C % 64 == 0 && T > 0 && N > 0, "vqb_attn_fwd: C=%d must be a multiple of the head dim 64", C);
    dim3 grid((T + 63) / 64, C / 64, N);
    attn_fwd_kernel<<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(qkv), static_cast<__nv_bfloat16*>(out), lse, T, C, 0.125f);
    VQB_CUDA(cudaGetLastError());
    count_launch();
    return VQB_OK;
}

// dqkv [N][T][3C] <- gradients of q, k, v. dvec: workspace [N][C/64][T] floats.
int vqb_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse, float* dvec, void* dqkv, int N,
                 int T, int C, void* stream) {
    VQB_CHECK(qkv && out && dout && lse && dvec && dqkv, "vqb_attn_bwd: null pointer");
    VQB_CHECK(C % 64 == 0 && T > 0 && N > 0, "vqb_attn_bwd: C=%d must be a multiple of the head dim 64", C);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int heads = C / 64;
    const int64_t total = static_cast<int64_t>(N) * T * heads;
    attn_bwd_prep_kernel<<<static_cast<int>((total + 7) / 8), 256, 0, st>>>(
        static_cast<const __nv_bfloat16*>(out), static_cast<const __nv_bfloat16*>(dout), dvec, T, C, heads, total);
    dim3 grid((T + 63) / 64, heads, N);
    const size_t smem = 4 * 64 * kLD * sizeof(__nv_bfloat16) + 2 * 64 * sizeof(float);
    static bool attr = false;
    if (!attr) {
        VQB_CUDA(cudaFuncSetAttribute(attn_bwd_dkdv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        attr = true;
    }
    attn_bwd_dkdv_kernel<<<grid, 128, smem, st>>>(static_cast<const __nv_bfloat16*>(qkv),
                                                   static_cast<const __nv_bfloat16*>(dout), lse, dvec,
                                                   static_cast<__nv_bfloat16*>(dqkv), T, C, 0.125f);
    attn_bwd_dq_kernel<<<grid, 128, 0, st>>>(static_cast<const __nv_bfloat16*>(qkv),
                                             static_cast<const __nv_bfloat16*>(dout), lse, dvec,
                                             static_cast<__nv_bfloat16*>(dqkv), T, C, 0.125f);
    VQB_CUDA(cudaGetLastError());
    count_launch(3);
    return VQB_OK;
}

}  // extern "C"
