// K1: one Chebyshev recurrence step on node-major features (replaces the dense einsum GCN.py:35 and the
// matrix-matrix recurrence GCN.py:125-135):   Y = alpha * op(A) X + beta * Z + gamma * U.
//
// One warp owns one output row x one 128-float column tile; lanes hold a float4 each, so every non-zero drives one
// fully coalesced 512 B gather of the neighbour's feature run.  The grid is column-tile-major (blockIdx.y = column tile)
// so the CTAs resident at any moment share one column tile of X (N x 512 B = 2 MB at N=4096): gathers are served by
// L2/L1, HBM sees X, Z, U once and Y once.
// Why there is no shared-memory-staged variant: the kernel's time is its gather volume nnz*F*4 B (2.87 GB per launch at
// BASELINE configs[2]) divided by what an on-chip level can serve for RANDOM row gathers.  tools/gather_probe.cu measures
// exactly that on B200 (profiles/r2_gather_probe.log): 512-byte segment gathers through L2/L1 17.9 TB/s (165 us for this
// launch's volume), an unstructured shared-memory strip 7.4-8.7 TB/s (bank conflicts), a persistent L1-resident 128-byte
// column tile 13.6-15.1 TB/s.  A round-1 strip kernel with a bank-aware sliced ELL reached 228 us.  This kernel runs at
// 170-187 us = 88-97 % of the best measured gather rate; the 70 %-of-HBM target would need 65 TB/s of gather bandwidth.
#include "common.cuh"
#include <stdlib.h>

namespace stmgcn {
void graph_view(const stmgcn_graph* g, bool transpose, int64_t* n, int64_t* nnz, const int32_t** rowptr,
                const int32_t** colidx, const float** vals, bool* ok);
}
using namespace stmgcn;

namespace {

constexpr int kRowsPerCta = 32;
constexpr int kWarpsPerCta = 8;

template <int VEC>
struct VecT;
template <>
struct VecT<4> {
    using type = float4;
};
template <>
struct VecT<1> {
    using type = float;
};

__device__ __forceinline__ void fma_vec(float4& a, float s, const float4& x) {
    a.x = fmaf(s, x.x, a.x);
    a.y = fmaf(s, x.y, a.y);
    a.z = fmaf(s, x.z, a.z);
    a.w = fmaf(s, x.w, a.w);
}
__device__ __forceinline__ void fma_vec(float& a, float s, const float& x) { a = fmaf(s, x, a); }
__device__ __forceinline__ float4 zero_vec(float4*) { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ float zero_vec(float*) { return 0.f; }
__device__ __forceinline__ float4 axpbypcz(float a, const float4& x, float b, const float4& y, float c,
                                           const float4& z) {
    return make_float4(a * x.x + b * y.x + c * z.x, a * x.y + b * y.y + c * z.y, a * x.z + b * y.z + c * z.z,
                       a * x.w + b * y.w + c * z.w);
}
__device__ __forceinline__ float axpbypcz(float a, float x, float b, float y, float c, float z) {
    return a * x + b * y + c * z;
}

template <int VEC>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
spmm_row_gather_kernel(int64_t n, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                       const float* __restrict__ vals, float alpha, const float* __restrict__ x, float beta,
                       const float* __restrict__ z, float gamma, const float* __restrict__ u,
                       float* __restrict__ y, int64_t f_total) {
    using V = typename VecT<VEC>::type;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int64_t f = ((int64_t)blockIdx.y * 32 + lane) * VEC;      // first feature of this lane
    const bool live = f < f_total;
    const int64_t row0 = (int64_t)blockIdx.x * kRowsPerCta;
    for (int r = warp; r < kRowsPerCta; r += kWarpsPerCta) {
        const int64_t row = row0 + r;
        if (row >= n) break;
        const int32_t beg = rowptr[row], end = rowptr[row + 1];
        V acc0 = zero_vec((V*)nullptr), acc1 = zero_vec((V*)nullptr);
        int32_t i = beg;
        if (live) {
            for (; i + 4 <= end; i += 4) {
                const int32_t c0 = __ldg(colidx + i), c1 = __ldg(colidx + i + 1);
                const int32_t c2 = __ldg(colidx + i + 2), c3 = __ldg(colidx + i + 3);
                const float v0 = __ldg(vals + i), v1 = __ldg(vals + i + 1);
                const float v2 = __ldg(vals + i + 2), v3 = __ldg(vals + i + 3);
                const V x0 = *reinterpret_cast<const V*>(x + (int64_t)c0 * f_total + f);
                const V x1 = *reinterpret_cast<const V*>(x + (int64_t)c1 * f_total + f);
                const V x2 = *reinterpret_cast<const V*>(x + (int64_t)c2 * f_total + f);
                const V x3 = *reinterpret_cast<const V*>(x + (int64_t)c3 * f_total + f);
                fma_vec(acc0, v0, x0);
                fma_vec(acc1, v1, x1);
                fma_vec(acc0, v2, x2);
                fma_vec(acc1, v3, x3);
            }
            for (; i < end; ++i) {
                const int32_t c0 = __ldg(colidx + i);
                const float v0 = __ldg(vals + i);
                const V x0 = *reinterpret_cast<const V*>(x + (int64_t)c0 * f_total + f);
                fma_vec(acc0, v0, x0);
            }
            V acc = axpbypcz(1.f, acc0, 1.f, acc1, 0.f, acc0);
            const int64_t off = row * f_total + f;
            V zz = zero_vec((V*)nullptr), uu = zero_vec((V*)nullptr);
            if (z != nullptr) zz = *reinterpret_cast<const V*>(z + off);
            if (u != nullptr) uu = *reinterpret_cast<const V*>(u + off);
            *reinterpret_cast<V*>(y + off) = axpbypcz(alpha, acc, beta, zz, gamma, uu);
        }
    }
}

}  // namespace

extern "C" int32_t stmgcn_cheb_spmm_step(const stmgcn_graph_t* g, int32_t transpose, float alpha,
                                         const float* x, float beta, const float* z, float gamma,
                                         const float* u, float* y, int64_t f_total, void* stream) {
    STMGCN_REQUIRE(g && x && y, STMGCN_ERR_ARG, "cheb_spmm_step: null pointer");
    STMGCN_REQUIRE(x != y, STMGCN_ERR_ARG, "cheb_spmm_step: y must not alias x");
    STMGCN_REQUIRE(f_total > 0, STMGCN_ERR_SHAPE, "cheb_spmm_step: f_total=%lld", (long long)f_total);
    int64_t n, nnz;
    const int32_t *rp, *ci;
    const float* va;
    bool ok;
    graph_view(g, transpose != 0, &n, &nnz, &rp, &ci, &va, &ok);
    STMGCN_REQUIRE(ok, STMGCN_ERR_STATE, "cheb_spmm_step: transpose requested but handle has none");
    cudaStream_t st = (cudaStream_t)stream;
    const bool vec4 = (f_total % 4 == 0) && aligned16(x) && aligned16(y) && (!z || aligned16(z)) &&
                      (!u || aligned16(u));
    const int vec = vec4 ? 4 : 1;
    const int64_t col_tiles = ceil_div(f_total, 32 * vec);
    STMGCN_REQUIRE(col_tiles <= 65535, STMGCN_ERR_SHAPE, "cheb_spmm_step: f_total=%lld too wide", (long long)f_total);
    dim3 grid((unsigned)ceil_div(n, kRowsPerCta), (unsigned)col_tiles);
    if (vec4)
        spmm_row_gather_kernel<4><<<grid, kWarpsPerCta * 32, 0, st>>>(n, rp, ci, va, alpha, x, beta, z, gamma, u, y, f_total);
    else
        spmm_row_gather_kernel<1><<<grid, kWarpsPerCta * 32, 0, st>>>(n, rp, ci, va, alpha, x, beta, z, gamma, u, y, f_total);
    count_launch();
    return check_launch("cheb_spmm_step");
}
