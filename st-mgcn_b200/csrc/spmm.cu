// K1: one Chebyshev recurrence step on node-major features (replaces the dense einsum GCN.py:35 and the
// matrix-matrix recurrence GCN.py:125-135):   Y = alpha * op(A) X + beta * Z + gamma * U.
//
// v1 kernel ("row gather"): one warp owns one output row x one 128-float column tile; lanes hold a float4
// each, so every non-zero drives one fully coalesced 512 B gather of the neighbour's feature run.  The
// grid is column-tile-major (blockIdx.y = column tile) so the CTAs resident at any moment share one column
// tile of X (N x 512 B = 2 MB at N=4096): gathers are served by L2/L1, HBM sees X, Z, U once and Y once.
#include "common.cuh"
#include <stdlib.h>

namespace stmgcn {
void graph_view(const stmgcn_graph* g, bool transpose, int64_t* n, int64_t* nnz, const int32_t** rowptr,
                const int32_t** colidx, const float** vals, bool* ok);
bool graph_ell_view(const stmgcn_graph* g, bool transpose, int32_t* n_slices, const int32_t** perm, const int32_t** off,
                    const uint2** ent);
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

// ---------------------------------------------------------------------------------------------------------
// v2 "strip" kernel: the CTA stages an (N x 8 fp32) column strip of X in shared memory (row j at byte 32*j) and every
// thread owns one output row of the strip, walking the graph's bank-aware sliced ELL (graph.cu): coalesced 8-byte
// {col, val} reads, two conflict-free 16-byte shared loads and 8 FMAs per non-zero.  HBM sees X, Z, U once and Y
// once; the gather volume nnz*32 B per strip is served by shared memory instead of L2.
// ---------------------------------------------------------------------------------------------------------
constexpr int kStripW = 8;
constexpr int kStripThreads = 512;

__device__ __forceinline__ void cp_async16_(void* smem, const void* gmem) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem) : "memory");
}

__global__ void __launch_bounds__(kStripThreads, 1)
spmm_strip_kernel(int64_t n, int32_t n_slices, const int32_t* __restrict__ perm, const int32_t* __restrict__ off,
                  const uint2* __restrict__ ent, float alpha, const float* __restrict__ x, float beta,
                  const float* __restrict__ z, float gamma, const float* __restrict__ u, float* __restrict__ y,
                  int64_t f_total) {
    extern __shared__ __align__(16) float xs[];                 // [n][8]
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int64_t f0 = (int64_t)blockIdx.x * kStripW;
    for (int64_t i = tid; i < n * 2; i += kStripThreads)        // two 16-byte pieces per row
        cp_async16_(xs + (i >> 1) * kStripW + (i & 1) * 4, x + (i >> 1) * f_total + f0 + (i & 1) * 4);
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();

    const float4* xs4 = reinterpret_cast<const float4*>(xs);
    const int par = lane & 1;
    // slices of this CTA: blockIdx.y-th share, interleaved over the warps
    for (int s = blockIdx.y * (kStripThreads / 32) + warp; s < n_slices; s += gridDim.y * (kStripThreads / 32)) {
        const int row = perm[s * 32 + lane];
        const int32_t beg = off[s];
        const int width = (off[s + 1] - beg) >> 5;
        const uint2* e = ent + beg + lane;
        // accA accumulates the half read first (half `par`), accB the other one: no per-slot selects
        float4 accA = make_float4(0.f, 0.f, 0.f, 0.f), accB = accA;
        int k = 0;
        if (width >= 4) {                 // software pipeline: the {col,val} entries of the next 4 slots are in flight
            uint2 n0 = e[0], n1 = e[32], n2 = e[64], n3 = e[96];
            for (; k + 4 <= width; k += 4) {
                const uint2 e0 = n0, e1 = n1, e2 = n2, e3 = n3;
                if (k + 8 <= width) {
                    n0 = e[(k + 4) * 32]; n1 = e[(k + 5) * 32]; n2 = e[(k + 6) * 32]; n3 = e[(k + 7) * 32];
                }
                const float4 a0 = xs4[e0.x * 2 + par], b0 = xs4[e0.x * 2 + (par ^ 1)];
                const float4 a1 = xs4[e1.x * 2 + par], b1 = xs4[e1.x * 2 + (par ^ 1)];
                const float4 a2 = xs4[e2.x * 2 + par], b2 = xs4[e2.x * 2 + (par ^ 1)];
                const float4 a3 = xs4[e3.x * 2 + par], b3 = xs4[e3.x * 2 + (par ^ 1)];
                fma_vec(accA, __uint_as_float(e0.y), a0); fma_vec(accB, __uint_as_float(e0.y), b0);
                fma_vec(accA, __uint_as_float(e1.y), a1); fma_vec(accB, __uint_as_float(e1.y), b1);
                fma_vec(accA, __uint_as_float(e2.y), a2); fma_vec(accB, __uint_as_float(e2.y), b2);
                fma_vec(accA, __uint_as_float(e3.y), a3); fma_vec(accB, __uint_as_float(e3.y), b3);
            }
        }
        for (; k < width; ++k) {
            const uint2 e0 = e[k * 32];
            const float4 a0 = xs4[e0.x * 2 + par], b0 = xs4[e0.x * 2 + (par ^ 1)];
            fma_vec(accA, __uint_as_float(e0.y), a0); fma_vec(accB, __uint_as_float(e0.y), b0);
        }
        if (row >= 0) {
            const float4 lo = par ? accB : accA, hi = par ? accA : accB;
            const int64_t o = (int64_t)row * f_total + f0;
            float4 zl = make_float4(0.f, 0.f, 0.f, 0.f), zh = zl, ul = zl, uh = zl;
            if (z != nullptr) { zl = *reinterpret_cast<const float4*>(z + o); zh = *reinterpret_cast<const float4*>(z + o + 4); }
            if (u != nullptr) { ul = *reinterpret_cast<const float4*>(u + o); uh = *reinterpret_cast<const float4*>(u + o + 4); }
            *reinterpret_cast<float4*>(y + o) = axpbypcz(alpha, lo, beta, zl, gamma, ul);
            *reinterpret_cast<float4*>(y + o + 4) = axpbypcz(alpha, hi, beta, zh, gamma, uh);
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
    {   // v2: shared-memory strip kernel when the strip of all N rows fits and the graph carries its ELL form
        int32_t n_slices;
        const int32_t *perm, *off;
        const uint2* ent;
        const size_t smem = (size_t)n * kStripW * sizeof(float);
        // Opt-in (STMGCN_SPMM_STRIP=1): measured on B200 at cfg3 the strip kernel takes 228 us per spatial step against
        // 171 us for the L2 row-gather kernel (its shared-memory wavefront count, 1.3x the conflict-free ideal, is the
        // limit), so the row-gather kernel stays the default.
        static int use_strip = -1;
        if (use_strip < 0) use_strip = getenv("STMGCN_SPMM_STRIP") ? 1 : 0;
        if (use_strip && f_total % kStripW == 0 && smem <= 200 * 1024 && n >= 8 && aligned16(x) && aligned16(y) &&
            (!z || aligned16(z)) && (!u || aligned16(u)) &&
            graph_ell_view(g, transpose != 0, &n_slices, &perm, &off, &ent)) {
            static bool attr_done = false;
            if (!attr_done) {
                STMGCN_CUDA(cudaFuncSetAttribute(spmm_strip_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
                attr_done = true;
            }
            const int64_t strips = f_total / kStripW;
            STMGCN_REQUIRE(strips <= 2147483647, STMGCN_ERR_SHAPE, "cheb_spmm_step: f_total too wide");
            // few strips (temporal GCN): split the rows over several CTAs per strip so every SM has work
            int64_t splits = 1;
            const int64_t max_splits = ceil_div(n_slices, kStripThreads / 32);
            while (strips * splits < 6 * (int64_t)sm_count() && splits < max_splits) ++splits;   // >= 6 waves: small tail
            dim3 grid((unsigned)strips, (unsigned)splits);
            spmm_strip_kernel<<<grid, kStripThreads, smem, st>>>(n, n_slices, perm, off, ent, alpha, x, beta, z, gamma, u, y,
                                                                 f_total);
            count_launch();
            return check_launch("cheb_spmm_step(strip)");
        }
    }
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
