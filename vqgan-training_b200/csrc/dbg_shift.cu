#ifdef VQB_DEBUG  // bring-up experiment: only part of libvqb200_dbg.so (build_native.py --debug)
// Bring-up experiment (not on the product path): does tcgen05.mma accept a 128B-swizzled K-major A operand whose
// descriptor start is shifted by whole 128-byte rows inside a TMA-written tile, and whose 8-row groups are SBO bytes
// apart for SBO not a multiple of 1024? This decides whether one activation halo tile in shared memory can serve all
// nine taps of a 3x3 convolution (tap shift = descriptor start offset) instead of nine separate TMA boxes.
//   out[m][n] = sum_k X[shift + (m/8)*(sbo/128) + m%8][k] * B[n][k]      m < 128, n < 64, k < 64
#include "common.cuh"
#include "ptx.cuh"

namespace vqb {

struct DbgShiftParams {
    CUtensorMap xmap, bmap;
    int32_t R, shift, sbo, base_offset;
    float* out;
};

__global__ void __launch_bounds__(128, 1) dbg_shift_kernel(const __grid_constant__ DbgShiftParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sA = base;                 // R rows x 128 B
    uint8_t* sB = base + 64 * 1024;     // 64 rows x 128 B
    uint64_t* bar = reinterpret_cast<uint64_t*>(sB + 8192);
    uint64_t* done = bar + 1;
    uint32_t* slot = reinterpret_cast<uint32_t*>(done + 1);
    const uint32_t warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_init(done, 1);
        fence_mbar_init();
    }
    if (warp == 0) {
        tmem_alloc(slot, 64);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *slot;
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(bar, static_cast<uint32_t>(p.R) * 128u + 8192u);
        for (int r0 = 0; r0 < p.R; r0 += 128) tma_load_2d(&p.xmap, bar, sA + r0 * 128, 0, r0);
        tma_load_2d(&p.bmap, bar, sB, 0, 0);
        mbar_wait(bar, 0);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(sA) + static_cast<uint32_t>(p.shift) * 128u;
        uint64_t da = make_smem_desc(a_addr, 0, static_cast<uint32_t>(p.sbo), 2);
        da |= static_cast<uint64_t>(p.base_offset & 7) << 49;
        const uint64_t db = make_smem_desc(smem_u32(sB), 0, 1024, 2);
        const uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
        for (int k = 0; k < 4; ++k) umma_bf16(tmem, da + 2 * k, db + 2 * k, idesc, k ? 1u : 0u);
        umma_commit(done);
    }
    mbar_wait(done, 0);
    tc_fence_after();
    uint32_t v0[32], v1[32];
    const uint32_t taddr = tmem + ((warp * 32u) << 16);
    tmem_ld32(taddr, v0);
    tmem_ld32(taddr + 32, v1);
    tmem_ld_wait();
    float* o = p.out + static_cast<size_t>(threadIdx.x) * 64;
    for (int j = 0; j < 32; ++j) {
        o[j] = __uint_as_float(v0[j]);
        o[32 + j] = __uint_as_float(v1[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem, 64);
    }
}

}  // namespace vqb

using namespace vqb;

// X [R][64] bf16 (R % 128 == 0, R <= 512), B [64][64] bf16 (row n = output column, K contiguous), out [128][64] fp32.
extern "C" int vqb_dbg_shift_mma(const void* X, int R, const void* B, float* out, int shift_rows, int sbo_bytes,
                                 int base_offset, void* stream) {
    VQB_CHECK(X && B && out && R > 0 && R % 128 == 0 && R <= 512, "vqb_dbg_shift_mma: bad arguments");
    DbgShiftParams p;
    p.R = R;
    p.shift = shift_rows;
    p.sbo = sbo_bytes;
    p.base_offset = base_offset;
    p.out = out;
    {
        uint64_t dims[2] = {64, static_cast<uint64_t>(R)};
        uint64_t str[1] = {128};
        uint32_t box[2] = {64, 128};
        int rc = encode_tmap_bf16(&p.xmap, X, 2, dims, str, box, 128);
        if (rc != VQB_OK) return rc;
        uint64_t bd[2] = {64, 64};
        uint32_t bb[2] = {64, 64};
        rc = encode_tmap_bf16(&p.bmap, B, 2, bd, str, bb, 128);
        if (rc != VQB_OK) return rc;
    }
    const size_t smem = 1024 + 64 * 1024 + 8192 + 64;
    VQB_CUDA(cudaFuncSetAttribute(dbg_shift_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    dbg_shift_kernel<<<1, 128, smem, static_cast<cudaStream_t>(stream)>>>(p);
    VQB_CUDA(cudaGetLastError());
    return VQB_OK;
}

#endif  // VQB_DEBUG
