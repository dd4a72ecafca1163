// Shared pieces of the tcgen05 kernels: tile constants, the mbarrier set, the single-thread MMA issuer, the tf32 hi/lo
// split + swizzled store helpers and the weight-image packer.  Included by lstm_tc.cu and proj_tc.cu.
#pragma once
#include "tc_common.cuh"

namespace stmgcn {
namespace tc {

constexpr int kTileM = 128;
constexpr int kKB = 32;              // k-block: one 128-byte swizzle row of fp32
constexpr int kHid = 64;
constexpr int kGateCols = 256;       // 4H
constexpr int kMaxStages = 4;
constexpr int kAccs = 2;
constexpr int kABytes = kTileM * kKB * 4;            // 16 KB per hi or lo A tile

struct Barriers {
    uint64_t full[kMaxStages];
    uint64_t empty[kMaxStages];
    uint64_t tmem_full[kAccs];
    uint64_t tmem_empty[kAccs];
    uint32_t tmem_base;
};

__device__ __forceinline__ void init_barriers(Barriers* b, int stages, int n_epi_threads, int n_loaders) {
    for (int s = 0; s < stages; ++s) {
        mbar_init(&b->full[s], n_loaders + 1);
        mbar_init(&b->empty[s], 1);
    }
    for (int a = 0; a < kAccs; ++a) {
        mbar_init(&b->tmem_full[a], 1);
        mbar_init(&b->tmem_empty[a], n_epi_threads);
    }
    fence_barrier_init();
}

// The MMA warp: for every tile, for every k-block: wait operands, issue 3 x 4 MMAs, release the stage.
template <int N, int STAGES, int PROF_KERNEL>
__device__ __forceinline__ void mma_issuer(Barriers* bar, uint8_t* smem, int stage_bytes, int b_bytes, int nkb,
                                           int n_tiles, uint32_t tmem_base, int lane) {
    constexpr uint32_t idesc = idesc_tf32(kTileM, N);
    TC_PROF_DECL
    (void)lane;
    const bool leader = elect_one_sync();      // (not `leader`: see elect_one_sync in tc_common.cuh)
    uint32_t it = 0, tcount = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tcount) {
        const int a = tcount & 1;
        const uint32_t aph = (tcount >> 1) & 1;
        mbar_wait(&bar->tmem_empty[a], aph ^ 1, 2);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)a * N;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
            const int s = it % STAGES;
            const uint32_t ph = (it / STAGES) & 1;
            mbar_wait(&bar->full[s], ph, 1);
            tc_fence_after();
            if (leader) {
                const uint32_t st = smem_u32(smem + (size_t)s * stage_bytes);
                const uint64_t a_hi = smem_desc_k_sw128(st);
                const uint64_t a_lo = smem_desc_k_sw128(st + kABytes);
                const uint64_t b_hi = smem_desc_k_sw128(st + 2 * kABytes);
                const uint64_t b_lo = smem_desc_k_sw128(st + 2 * kABytes + b_bytes);
#pragma unroll
                for (int pass = 0; pass < 3; ++pass) {
                    const uint64_t da = (pass == 1) ? a_lo : a_hi;
                    const uint64_t db = (pass == 2) ? b_lo : b_hi;
#pragma unroll
                    for (int k = 0; k < kKB / 8; ++k) {
                        const uint32_t acc = (kb > 0 || pass > 0 || k > 0) ? 1u : 0u;
                        mma_tf32(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, acc);
                    }
                }
                mma_commit(&bar->empty[s]);
            }
            __syncwarp();
        }
        if (leader) mma_commit(&bar->tmem_full[a]);
        __syncwarp();
    }
    TC_PROF_FLUSH(PROF_KERNEL * 3 + 1, leader)
}

__device__ __forceinline__ void split_store(uint8_t* st, uint32_t off, const float4& v) {
    float4 hi, lo;
    hi.x = tf32_hi(v.x); hi.y = tf32_hi(v.y); hi.z = tf32_hi(v.z); hi.w = tf32_hi(v.w);
    lo.x = tf32_lo(v.x, hi.x); lo.y = tf32_lo(v.y, hi.y); lo.z = tf32_lo(v.z, hi.z); lo.w = tf32_lo(v.w, hi.w);
    *reinterpret_cast<float4*>(st + off) = hi;
    *reinterpret_cast<float4*>(st + kABytes + off) = lo;
}


// transpose-store one float4 (4 consecutive M/N indices mn..mn+3 of row k) into a K-major swizzled tile pair
__device__ __forceinline__ void split_store_t(uint8_t* hi_tile, uint8_t* lo_tile, int mn, int k, const float4& v) {
    const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const uint32_t off = sw128_offset((uint32_t)(mn + j), (uint32_t)k);
        const float hi = tf32_hi(vv[j]);
        *reinterpret_cast<float*>(hi_tile + off) = hi;
        *reinterpret_cast<float*>(lo_tile + off) = tf32_lo(vv[j], hi);
    }
}

// Generic K-major hi/lo image of a logical B[n][k] = src[n*rs + k*cs]: per 32-wide k-block [hi | lo], each an
// [n_rows][32] fp32 tile with the 128-byte swizzle.
static __global__ void pack_image_kernel(const float* __restrict__ src, int n_rows, int k_cols, int64_t rs, int64_t cs,
                                         float* __restrict__ img, int tile_rows) {
    const int total = n_rows * k_cols;
    const int tile_floats = tile_rows * kKB;      // tile_rows >= n_rows: extra rows keep what the caller put there (zeros)
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int n = e / k_cols, k = e % k_cols;
        const float v = src[(int64_t)n * rs + (int64_t)k * cs];
        const float hi = tf32_hi(v);
        const float lo = tf32_lo(v, hi);
        const int kb = k / kKB, kk = k % kKB;
        const uint32_t off = sw128_offset((uint32_t)n, (uint32_t)kk) / 4;
        float* base = img + (size_t)kb * (2 * tile_floats);
        base[off] = hi;
        base[tile_floats + off] = lo;
    }
}

}  // namespace tc
}  // namespace stmgcn
