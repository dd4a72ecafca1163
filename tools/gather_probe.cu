// Gather-bandwidth microbenchmark behind the Chebyshev SpMM's ceiling argument.
// The SpMM at BASELINE configs[2] (N = 4096 regions, F = B*H = 4096 feature columns, ~42 non-zeros per row) moves
// nnz * F * 4 B = 2.87 GB of gathered feature data per launch against 0.2 GB of algorithmic HBM traffic, so its time is
// set by how fast an on-chip level can serve random row gathers.  This program measures exactly that, nothing else:
//   l2  : every warp gathers 512-byte row segments (one float4 per lane) of random rows of X[N][F] -- the access
//         pattern of spmm_row_gather_kernel -- column-tile-major so a wave shares a 2 MB column tile (L1 + L2 serve it);
//         variants: default caching (L1 allowed) and ld.global.cg (L2 only)
//   smem: a CTA stages an [N][W] fp32 strip in shared memory and every thread gathers W floats of random rows with
//         LDS.128 (the access pattern of a shared-memory strip kernel), W = 8 and 12
// Output: GB/s of gathered bytes.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/gather_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>

__device__ __forceinline__ uint32_t lcg(uint32_t& s) { s = s * 1664525u + 1013904223u; return s >> 8; }

template <bool CG>
__global__ void __launch_bounds__(256) l2_gather(const float* __restrict__ x, float* __restrict__ y, int n, int64_t f,
                                                 int deg, const int* __restrict__ cols) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t fo = ((int64_t)blockIdx.y * 32 + lane) * 4;
    const int row0 = blockIdx.x * 32;
    for (int r = warp; r < 32; r += 8) {
        const int row = row0 + r;
        if (row >= n) break;
        float4 a0 = make_float4(0, 0, 0, 0), a1 = a0;
        const int* c = cols + (int64_t)row * deg;
        for (int i = 0; i + 4 <= deg; i += 4) {
            const int c0 = __ldg(c + i), c1 = __ldg(c + i + 1), c2 = __ldg(c + i + 2), c3 = __ldg(c + i + 3);
            float4 v0, v1, v2, v3;
            const float* p0 = x + (int64_t)c0 * f + fo; const float* p1 = x + (int64_t)c1 * f + fo;
            const float* p2 = x + (int64_t)c2 * f + fo; const float* p3 = x + (int64_t)c3 * f + fo;
            if (CG) {
                v0 = __ldcg(reinterpret_cast<const float4*>(p0)); v1 = __ldcg(reinterpret_cast<const float4*>(p1));
                v2 = __ldcg(reinterpret_cast<const float4*>(p2)); v3 = __ldcg(reinterpret_cast<const float4*>(p3));
            } else {
                v0 = *reinterpret_cast<const float4*>(p0); v1 = *reinterpret_cast<const float4*>(p1);
                v2 = *reinterpret_cast<const float4*>(p2); v3 = *reinterpret_cast<const float4*>(p3);
            }
            a0.x += v0.x + v2.x; a0.y += v0.y + v2.y; a0.z += v0.z + v2.z; a0.w += v0.w + v2.w;
            a1.x += v1.x + v3.x; a1.y += v1.y + v3.y; a1.z += v1.z + v3.z; a1.w += v1.w + v3.w;
        }
        *reinterpret_cast<float4*>(y + (int64_t)row * f + fo) = make_float4(a0.x + a1.x, a0.y + a1.y, a0.z + a1.z, a0.w + a1.w);
    }
}

// one thread = one output row of the strip; W floats per gather
template <int W>
__global__ void __launch_bounds__(512, 1) smem_gather(const float* __restrict__ x, float* __restrict__ y, int n, int64_t f,
                                                      int deg, const int* __restrict__ cols) {
    extern __shared__ __align__(16) float xs[];
    const int64_t f0 = (int64_t)blockIdx.x * W;
    for (int i = threadIdx.x; i < n * (W / 4); i += 512)
        reinterpret_cast<float4*>(xs)[i] = *reinterpret_cast<const float4*>(x + (int64_t)(i / (W / 4)) * f + f0 + (i % (W / 4)) * 4);
    __syncthreads();
    const float4* xs4 = reinterpret_cast<const float4*>(xs);
    for (int row = threadIdx.x; row < n; row += 512) {
        float4 acc[W / 4];
#pragma unroll
        for (int j = 0; j < W / 4; ++j) acc[j] = make_float4(0, 0, 0, 0);
        // column indices transposed per 32-row slice so a warp's index loads are coalesced (ELL order)
        const int* c = cols + (int64_t)(row & ~31) * deg + (row & 31);
        for (int i = 0; i < deg; ++i) {
            const int cc = __ldg(c + i * 32);
#pragma unroll
            for (int j = 0; j < W / 4; ++j) {
                const float4 v = xs4[cc * (W / 4) + j];
                acc[j].x += v.x; acc[j].y += v.y; acc[j].z += v.z; acc[j].w += v.w;
            }
        }
#pragma unroll
        for (int j = 0; j < W / 4; ++j) *reinterpret_cast<float4*>(y + (int64_t)row * f + f0 + 4 * j) = acc[j];
    }
}

// L1-resident variant: ONE persistent CTA per 128-byte column tile (32 floats per row): the SM's whole gather stream hits
// one N x 128 B = 512 KB slice of X, of which the ~200 KB L1 keeps a good part.  8 lanes fetch one row segment, so a
// warp-wide 128-bit load covers 4 different rows (4 L1 wavefronts).
__global__ void __launch_bounds__(1024, 1) l1_tile_gather(const float* __restrict__ x, float* __restrict__ y, int n, int64_t f,
                                                          int deg, const int* __restrict__ cols, int n_tiles) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane >> 3, l8 = lane & 7;                 // 4 rows per warp step, 8 lanes x float4 per row
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t fo = (int64_t)tile * 32 + l8 * 4;
        for (int row = warp * 4 + sub; row < n; row += 32 * 4) {
            float4 a0 = make_float4(0, 0, 0, 0), a1 = a0;
            const int* c = cols + (int64_t)row * deg;
            for (int i = 0; i + 4 <= deg; i += 4) {
                const int c0 = __ldg(c + i), c1 = __ldg(c + i + 1), c2 = __ldg(c + i + 2), c3 = __ldg(c + i + 3);
                const float4 v0 = *reinterpret_cast<const float4*>(x + (int64_t)c0 * f + fo);
                const float4 v1 = *reinterpret_cast<const float4*>(x + (int64_t)c1 * f + fo);
                const float4 v2 = *reinterpret_cast<const float4*>(x + (int64_t)c2 * f + fo);
                const float4 v3 = *reinterpret_cast<const float4*>(x + (int64_t)c3 * f + fo);
                a0.x += v0.x + v2.x; a0.y += v0.y + v2.y; a0.z += v0.z + v2.z; a0.w += v0.w + v2.w;
                a1.x += v1.x + v3.x; a1.y += v1.y + v3.y; a1.z += v1.z + v3.z; a1.w += v1.w + v3.w;
            }
            *reinterpret_cast<float4*>(y + (int64_t)row * f + fo) = make_float4(a0.x + a1.x, a0.y + a1.y, a0.z + a1.z, a0.w + a1.w);
        }
    }
}

int main() {
    const int n = 4096, deg = 44;
    const int64_t f = 4096;
    float *x, *y;
    int* cols;
    cudaMalloc(&x, n * f * 4); cudaMalloc(&y, n * f * 4); cudaMalloc(&cols, (size_t)n * deg * 4);
    cudaMemset(x, 0, n * f * 4);
    std::vector<int> h((size_t)n * deg);
    uint32_t s = 12345;
    for (auto& v : h) { s = s * 1664525u + 1013904223u; v = (s >> 8) % n; }
    cudaMemcpy(cols, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const double gathered = (double)n * deg * f * 4;
    auto report = [&](const char* name, float ms, int reps) {
        printf("{\"probe\": \"%s\", \"us_per_launch\": %.1f, \"gathered_GB\": %.3f, \"gather_GBps\": %.0f}\n", name,
               ms * 1e3 / reps, gathered / 1e9, gathered / (ms / reps * 1e-3) / 1e9);
    };
    const int reps = 10;
    float ms;
    dim3 grid(n / 32, (unsigned)(f / 128));
    for (int v = 0; v < 2; ++v) {
        for (int i = 0; i < reps + 2; ++i) {
            if (i == 2) cudaEventRecord(e0);
            if (v == 0) l2_gather<false><<<grid, 256>>>(x, y, n, f, deg, cols);
            else l2_gather<true><<<grid, 256>>>(x, y, n, f, deg, cols);
        }
        cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        report(v == 0 ? "l2_gather_512B_segments_L1_allowed" : "l2_gather_512B_segments_ldcg_L2_only", ms, reps);
    }
    {
        cudaFuncSetAttribute(smem_gather<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, n * 8 * 4);
        for (int i = 0; i < reps + 2; ++i) {
            if (i == 2) cudaEventRecord(e0);
            smem_gather<8><<<(unsigned)(f / 8), 512, n * 8 * 4>>>(x, y, n, f, deg, cols);
        }
        cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        report("smem_gather_strip_W8_lds128_random_rows", ms, reps);
    }
    {
        const int64_t f12 = 4092;        // 341 strips of 12 columns
        cudaFuncSetAttribute(smem_gather<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, n * 12 * 4);
        for (int i = 0; i < reps + 2; ++i) {
            if (i == 2) cudaEventRecord(e0);
            smem_gather<12><<<(unsigned)(f12 / 12), 512, n * 12 * 4>>>(x, y, n, f, deg, cols);
        }
        cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        printf("{\"probe\": \"smem_gather_strip_W12_lds128_random_rows\", \"us_per_launch\": %.1f, \"gather_GBps\": %.0f}\n",
               ms * 1e3 / reps, (double)n * deg * f12 * 4 / (ms / reps * 1e-3) / 1e9);
    }
    for (int carve = 0; carve < 2; ++carve) {
        // 3 graphs worth of column tiles (384 work items over 148 persistent CTAs), L1 carve-out maximised
        cudaFuncSetAttribute(l1_tile_gather, cudaFuncAttributePreferredSharedMemoryCarveout, carve == 0 ? 0 : 50);
        const int n_tiles = (int)(f / 32) * 3;
        float* x3; float* y3;
        cudaMalloc(&x3, 3 * n * f * 4); cudaMalloc(&y3, 3 * n * f * 4);
        cudaMemset(x3, 0, 3 * n * f * 4);
        for (int i = 0; i < reps + 2; ++i) {
            if (i == 2) cudaEventRecord(e0);
            l1_tile_gather<<<148, 1024>>>(x3, y3, n, 3 * f, deg, cols, n_tiles);
        }
        cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        printf("{\"probe\": \"l1_tile_gather_128B_tiles_persistent_%s\", \"us_per_launch_3graphs\": %.1f, \"us_per_graph\": %.1f, \"gather_GBps\": %.0f}\n",
               carve == 0 ? "maxL1" : "carve50", ms * 1e3 / reps, ms * 1e3 / reps / 3, 3 * gathered / (ms / reps * 1e-3) / 1e9);
        cudaFree(x3); cudaFree(y3);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("cuda: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
