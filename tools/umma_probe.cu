// Stand-alone probe of the bf16 tcgen05 operand layouts the LSTM kernels rely on (st-mgcn_b200/csrc/tc16.cuh):
//   test 0: A K-major [128 x 64]  . B K-major  [64 x 64]^T           (gate GEMM form)
//   test 1: A K-major [128 x 64]  . B MN-major (tile [64 k][64 n])    (data-gradient form: B is the weight tile transposed)
//   test 2: A MN-major (2 tiles [128 k][64 m], LBO apart) . B MN-major (tile [128 k][64 n])   (weight-gradient form)
//   test 3: test 0 with A in fp16 and B in bf16 (mixed kinds -- information only)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I st-mgcn_b200/csrc tools/umma_probe.cu -o gpurun_out/umma_probe
#include "tc16.cuh"
#include <cuda_fp16.h>
#include <vector>
#include <cstdio>
#include <cmath>
#include <cstdlib>

using namespace stmgcn::tc;

struct Args {
    const float* a;   // logical A[m][k] fp32 (m < 128)
    const float* b;   // logical B[n][k]
    float* d;         // [128][n]
    int n, k, test;
};

__global__ void __launch_bounds__(128, 1) probe_kernel(Args p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    uint8_t* at = smem;                 // A tiles: up to 2 x 16 KB
    uint8_t* bt = smem + 2 * kTile16Bytes;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(&tmem_slot, 64);
    auto to16 = [&](float v, bool half) -> uint16_t {
        if (half) { __half h = __float2half_rn(v); return *reinterpret_cast<uint16_t*>(&h); }
        __nv_bfloat16 h = __float2bfloat16_rn(v); return *reinterpret_cast<uint16_t*>(&h);
    };
    const bool a_half = p.test == 3;
    if (p.test == 0 || p.test == 3 || p.test == 1) {
        // A K-major: tile row = m, col = k (k < 64)
        for (int e = tid; e < 128 * 64; e += 128) {
            const int m = e / 64, k = e % 64;
            *reinterpret_cast<uint16_t*>(at + sw128_off16(m, k)) = to16(p.a[m * p.k + k], a_half);
        }
    } else {
        // A MN-major: tile j (m in [64j, 64j+64)): row = k (k < 128), col = m - 64j
        for (int e = tid; e < 128 * 128; e += 128) {
            const int m = e / 128, k = e % 128;
            *reinterpret_cast<uint16_t*>(at + (m / 64) * kTile16Bytes + sw128_off16(k, m % 64)) = to16(p.a[m * p.k + k], false);
        }
    }
    if (p.test == 0 || p.test == 3) {
        for (int e = tid; e < p.n * 64; e += 128) {        // B K-major: row = n, col = k
            const int n = e / 64, k = e % 64;
            *reinterpret_cast<uint16_t*>(bt + sw128_off16(n, k)) = to16(p.b[n * p.k + k], false);
        }
    } else {
        for (int e = tid; e < p.n * p.k; e += 128) {       // B MN-major: row = k, col = n (n < 64)
            const int n = e / p.k, k = e % p.k;
            *reinterpret_cast<uint16_t*>(bt + sw128_off16(k, n)) = to16(p.b[n * p.k + k], false);
        }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (tid == 0) {
        const uint32_t a0 = smem_u32(at), b0 = smem_u32(bt);
        const int a_mn = (p.test == 2), b_mn = (p.test == 1 || p.test == 2);
        uint32_t idesc = idesc_bf16(128, p.n, a_mn, b_mn);
        if (p.test == 3) idesc &= ~(1u << 7);              // a_format = F16
        for (int ks = 0; ks < p.k / 16; ++ks) {
            const uint64_t da = a_mn ? desc16_mn(a0 + ks * 2048, kTile16Bytes) : desc16_k(a0) + (uint64_t)(2 * ks);
            const uint64_t db = b_mn ? desc16_mn(b0 + ks * 2048, kTile16Bytes) : desc16_k(b0) + (uint64_t)(2 * ks);
            mma_bf16(tmem, da, db, idesc, ks > 0);
        }
        mma_commit(&bar);
    }
    mbar_wait_raw(&bar, 0);
    tc_fence_after();
    for (int c0 = 0; c0 < p.n; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
        tmem_ld_wait();
        for (int j = 0; j < 32; ++j) p.d[(warp * 32 + lane) * p.n + c0 + j] = __uint_as_float(v[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 64);
}

static float bf(float v) { __nv_bfloat16 h = __float2bfloat16_rn(v); return __bfloat162float(h); }
static float hf(float v) { __half h = __float2half_rn(v); return __half2float(h); }

int main() {
    int fails = 0;
    for (int test = 0; test < 4; ++test) {
        const int n = 64, k = (test == 2) ? 128 : 64;
        std::vector<float> a(128 * k), b(n * k), d(128 * n), ref(128 * n);
        srand(1 + test);
        for (auto& v : a) v = (rand() % 2001 - 1000) / 1000.f;
        for (auto& v : b) v = (rand() % 2001 - 1000) / 1000.f;
        for (int m = 0; m < 128; ++m)
            for (int j = 0; j < n; ++j) {
                double s = 0;
                for (int kk = 0; kk < k; ++kk) s += (double)(test == 3 ? hf(a[m * k + kk]) : bf(a[m * k + kk])) * bf(b[j * k + kk]);
                ref[m * n + j] = (float)s;
            }
        float *da, *db, *dd;
        cudaMalloc(&da, a.size() * 4); cudaMalloc(&db, b.size() * 4); cudaMalloc(&dd, d.size() * 4);
        cudaMemcpy(da, a.data(), a.size() * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(db, b.data(), b.size() * 4, cudaMemcpyHostToDevice);
        cudaMemset(dd, 0, d.size() * 4);
        Args p{da, db, dd, n, k, test};
        const size_t smem = 1024 + 3 * kTile16Bytes;
        cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        probe_kernel<<<1, 128, smem>>>(p);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(d.data(), dd, d.size() * 4, cudaMemcpyDeviceToHost);
        double maxerr = 0, maxref = 0;
        for (size_t i = 0; i < d.size(); ++i) { maxerr = fmax(maxerr, fabs(d[i] - ref[i])); maxref = fmax(maxref, fabs(ref[i])); }
        const bool ok = (e == cudaSuccess) && maxerr <= 1e-4 * maxref;
        printf("umma_probe test %d: %s  (cuda: %s, max err %.3e, max |ref| %.3e)\n", test, ok ? "PASS" : "FAIL",
               cudaGetErrorString(e), maxerr, maxref);
        if (!ok && test != 3) ++fails;
        if (e != cudaSuccess) { cudaDeviceReset(); }
        cudaFree(da); cudaFree(db); cudaFree(dd);
    }
    return fails ? 1 : 0;
}
