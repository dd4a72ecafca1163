// Small memory-bound kernels around the three hot kernels: layout change of the observations, the context
// gate's tiny FC (STMGCN.py:42-43) and the fusion over graphs + output FC (STMGCN.py:116-118).
#include "common.cuh"

using namespace stmgcn;

namespace {

// obs (B,T,N,C) -> xo (N,B,T,C), xt (N,B,T) = sum_c      (STMGCN.py:36, :39, :47)
__global__ void obs_to_node_major_kernel(const float* __restrict__ obs, float* __restrict__ xo,
                                         float* __restrict__ xt, int64_t b_sz, int64_t t_len, int64_t n,
                                         int64_t c_in) {
    // one thread per (n, b, t); consecutive threads walk t then b (coalesced writes, strided reads;
    // the whole tensor is ~13 MB at 4096 regions x 64 windows x 12 steps).
    const int64_t total = n * b_sz * t_len;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t t = i % t_len;
        const int64_t b = (i / t_len) % b_sz;
        const int64_t nn = i / (t_len * b_sz);
        const float* src = obs + ((b * t_len + t) * n + nn) * c_in;
        float sum = 0.f;
        for (int64_t c = 0; c < c_in; ++c) {
            const float v = src[c];
            sum += v;
            if (xo != nullptr) xo[i * c_in + c] = v;
        }
        xt[i] = sum;
    }
}

// ---- context gate ----------------------------------------------------------------------------------
// one CTA per window b; T threads-worth of work looped over blockDim
__global__ void gate_fwd_kernel(const float* __restrict__ pool, int t_len, float inv_n,
                                const float* __restrict__ fcw, const float* __restrict__ fcb,
                                float* __restrict__ z, float* __restrict__ a1, float* __restrict__ s) {
    extern __shared__ float sm[];            // z[T], r1[T]
    float* zs = sm;
    float* rs = sm + t_len;
    const int64_t b = blockIdx.x;
    for (int j = threadIdx.x; j < t_len; j += blockDim.x) {
        const float v = pool[b * t_len + j] * inv_n;
        zs[j] = v;
        z[b * t_len + j] = v;
    }
    __syncthreads();
    for (int j = threadIdx.x; j < t_len; j += blockDim.x) {
        float acc = fcb[j];
        for (int i = 0; i < t_len; ++i) acc = fmaf(fcw[j * t_len + i], zs[i], acc);
        a1[b * t_len + j] = acc;
        rs[j] = fmaxf(acc, 0.f);
    }
    __syncthreads();
    for (int j = threadIdx.x; j < t_len; j += blockDim.x) {
        float acc = fcb[j];
        for (int i = 0; i < t_len; ++i) acc = fmaf(fcw[j * t_len + i], rs[i], acc);
        s[b * t_len + j] = sigmoidf_(acc);
    }
}

__global__ void gate_bwd_kernel(const float* __restrict__ d_s, const float* __restrict__ z,
                                const float* __restrict__ a1, const float* __restrict__ s, int t_len,
                                const float* __restrict__ fcw, float* __restrict__ d_fcw,
                                float* __restrict__ d_fcb, float* __restrict__ d_z) {
    extern __shared__ float sm[];            // da2[T], da1[T], r1[T], z[T]
    float* da2 = sm;
    float* da1 = sm + t_len;
    float* r1 = sm + 2 * t_len;
    float* zs = sm + 3 * t_len;
    const int64_t b = blockIdx.x;
    for (int j = threadIdx.x; j < t_len; j += blockDim.x) {
        const float sv = s[b * t_len + j];
        da2[j] = d_s[b * t_len + j] * sv * (1.f - sv);
        r1[j] = fmaxf(a1[b * t_len + j], 0.f);
        zs[j] = z[b * t_len + j];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < t_len; i += blockDim.x) {
        float acc = 0.f;                                     // d r1[i] = sum_j da2[j] fcw[j,i]
        for (int j = 0; j < t_len; ++j) acc = fmaf(da2[j], fcw[j * t_len + i], acc);
        da1[i] = (a1[b * t_len + i] > 0.f) ? acc : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < t_len; i += blockDim.x) {
        float acc = 0.f;                                     // d z[i] = sum_j da1[j] fcw[j,i]
        for (int j = 0; j < t_len; ++j) acc = fmaf(da1[j], fcw[j * t_len + i], acc);
        d_z[b * t_len + i] = acc;
        atomicAdd(&d_fcb[i], da2[i] + da1[i]);
    }
    for (int e = threadIdx.x; e < t_len * t_len; e += blockDim.x) {
        const int j = e / t_len, i = e % t_len;              // fc used twice: both uses accumulate
        atomicAdd(&d_fcw[e], da2[j] * r1[i] + da1[j] * zs[i]);
    }
}

// ---- fusion over graphs + output FC -------------------------------------------------------------------
constexpr int kMaxGraphs = 8;
struct GraphPtrs {
    const float* g[kMaxGraphs];
};

// one warp per node-major row r = n*B + b
__global__ void fuse_out_fwd_kernel(GraphPtrs gp, int m, int64_t n, int64_t b_sz, int gdim, int c_out,
                                    const float* __restrict__ fcw, const float* __restrict__ fcb,
                                    float* __restrict__ feat, float* __restrict__ y) {
    const int lane = threadIdx.x & 31;
    const int64_t rows = n * b_sz;
    for (int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < rows;
         r += (int64_t)gridDim.x * (blockDim.x >> 5)) {
        const int64_t nn = r / b_sz, b = r % b_sz;
        for (int c = 0; c < c_out; ++c) {
            float dot = 0.f;
            for (int g = lane; g < gdim; g += 32) {
                float v = 0.f;
                for (int k = 0; k < m; ++k) v += gp.g[k][r * gdim + g];
                if (c == 0) feat[r * gdim + g] = v;
                dot = fmaf(v, fcw[c * gdim + g], dot);
            }
            dot = warp_sum(dot);
            if (lane == 0) y[(b * n + nn) * c_out + c] = dot + fcb[c];
        }
    }
}

__global__ void fuse_out_bwd_kernel(const float* __restrict__ d_y, const float* __restrict__ feat, int64_t n,
                                    int64_t b_sz, int gdim, int c_out, const float* __restrict__ fcw,
                                    float* __restrict__ d_feat, float* __restrict__ d_fcw,
                                    float* __restrict__ d_fcb) {
    extern __shared__ float sacc[];          // c_out*gdim + c_out
    const int lane = threadIdx.x & 31;
    const int n_acc = c_out * gdim + c_out;
    for (int e = threadIdx.x; e < n_acc; e += blockDim.x) sacc[e] = 0.f;
    __syncthreads();
    const int64_t rows = n * b_sz;
    for (int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < rows;
         r += (int64_t)gridDim.x * (blockDim.x >> 5)) {
        const int64_t nn = r / b_sz, b = r % b_sz;
        const float* dyr = d_y + (b * n + nn) * c_out;
        for (int g = lane; g < gdim; g += 32) {
            const float fv = feat[r * gdim + g];
            float acc = 0.f;
            for (int c = 0; c < c_out; ++c) {
                const float dv = dyr[c];
                acc = fmaf(dv, fcw[c * gdim + g], acc);
                atomicAdd(&sacc[c * gdim + g], dv * fv);
            }
            d_feat[r * gdim + g] = acc;
        }
        if (lane < c_out) atomicAdd(&sacc[c_out * gdim + lane], dyr[lane]);
        for (int c = 32 + lane; c < c_out; c += 32) atomicAdd(&sacc[c_out * gdim + c], dyr[c]);
    }
    __syncthreads();
    for (int e = threadIdx.x; e < c_out * gdim; e += blockDim.x) atomicAdd(&d_fcw[e], sacc[e]);
    for (int e = threadIdx.x; e < c_out; e += blockDim.x) atomicAdd(&d_fcb[e], sacc[c_out * gdim + e]);
}

}  // namespace

extern "C" {

int32_t stmgcn_obs_to_node_major(const float* obs, float* xo, float* xt, int64_t b, int64_t t, int64_t n,
                                 int64_t c, void* stream) {
    STMGCN_REQUIRE(obs && xt, STMGCN_ERR_ARG, "obs_to_node_major: null pointer");
    STMGCN_REQUIRE(b > 0 && t > 0 && n > 0 && c > 0, STMGCN_ERR_SHAPE, "obs_to_node_major: bad shape");
    STMGCN_REQUIRE(xo != nullptr || c == 1, STMGCN_ERR_ARG, "obs_to_node_major: xo required when C > 1");
    const int64_t total = n * b * t;
    const int64_t blocks = ceil_div(total, 256);
    const int64_t cap = (int64_t)sm_count() * 16;
    obs_to_node_major_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream>>>(
        obs, xo, xt, b, t, n, c);
    count_launch();
    return check_launch("obs_to_node_major");
}

int32_t stmgcn_gate_fwd(const float* pool, int64_t b, int32_t t, int64_t n_regions, const float* fcw,
                        const float* fcb, float* z, float* a1, float* s, void* stream) {
    STMGCN_REQUIRE(pool && fcw && fcb && z && a1 && s, STMGCN_ERR_ARG, "gate_fwd: null pointer");
    STMGCN_REQUIRE(b > 0 && t > 0 && t <= 4096 && n_regions > 0, STMGCN_ERR_SHAPE, "gate_fwd: bad shape");
    const int threads = t <= 32 ? 32 : (t <= 128 ? 128 : 256);
    gate_fwd_kernel<<<(unsigned)b, threads, 2 * t * sizeof(float), (cudaStream_t)stream>>>(
        pool, t, 1.0f / (float)n_regions, fcw, fcb, z, a1, s);
    count_launch();
    return check_launch("gate_fwd");
}

int32_t stmgcn_gate_bwd(const float* d_s, const float* z, const float* a1, const float* s, int64_t b,
                        int32_t t, const float* fcw, float* d_fcw, float* d_fcb, float* d_z, void* stream) {
    STMGCN_REQUIRE(d_s && z && a1 && s && fcw && d_fcw && d_fcb && d_z, STMGCN_ERR_ARG, "gate_bwd: null pointer");
    STMGCN_REQUIRE(b > 0 && t > 0 && t <= 2048, STMGCN_ERR_SHAPE, "gate_bwd: bad shape");
    const int threads = t <= 32 ? 32 : (t <= 128 ? 128 : 256);
    gate_bwd_kernel<<<(unsigned)b, threads, 4 * t * sizeof(float), (cudaStream_t)stream>>>(
        d_s, z, a1, s, t, fcw, d_fcw, d_fcb, d_z);
    count_launch();
    return check_launch("gate_bwd");
}

int32_t stmgcn_fuse_out_fwd(const float* const* g, int32_t m, int64_t n, int64_t b, int32_t gdim, int32_t c,
                            const float* fcw, const float* fcb, float* feat, float* y, void* stream) {
    STMGCN_REQUIRE(g && fcw && fcb && feat && y, STMGCN_ERR_ARG, "fuse_out_fwd: null pointer");
    STMGCN_REQUIRE(m >= 1 && m <= kMaxGraphs, STMGCN_ERR_SHAPE, "fuse_out_fwd: M=%d (max %d)", m, kMaxGraphs);
    STMGCN_REQUIRE(n > 0 && b > 0 && gdim > 0 && c > 0, STMGCN_ERR_SHAPE, "fuse_out_fwd: bad shape");
    GraphPtrs gp;
    for (int k = 0; k < kMaxGraphs; ++k) gp.g[k] = k < m ? g[k] : nullptr;
    for (int k = 0; k < m; ++k) STMGCN_REQUIRE(gp.g[k], STMGCN_ERR_ARG, "fuse_out_fwd: g[%d] null", k);
    const int64_t rows = n * b;
    const int64_t blocks = ceil_div(rows, 8);
    const int64_t cap = (int64_t)sm_count() * 8;
    fuse_out_fwd_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream>>>(
        gp, m, n, b, gdim, c, fcw, fcb, feat, y);
    count_launch();
    return check_launch("fuse_out_fwd");
}

int32_t stmgcn_fuse_out_bwd(const float* d_y, const float* feat, int64_t n, int64_t b, int32_t gdim,
                            int32_t c, const float* fcw, float* d_feat, float* d_fcw, float* d_fcb,
                            void* stream) {
    STMGCN_REQUIRE(d_y && feat && fcw && d_feat && d_fcw && d_fcb, STMGCN_ERR_ARG, "fuse_out_bwd: null pointer");
    STMGCN_REQUIRE(n > 0 && b > 0 && gdim > 0 && c > 0, STMGCN_ERR_SHAPE, "fuse_out_bwd: bad shape");
    const size_t smem = ((size_t)c * gdim + c) * sizeof(float);
    STMGCN_REQUIRE(smem <= 48 * 1024, STMGCN_ERR_SHAPE, "fuse_out_bwd: C*G=%d too large", c * gdim);
    const int64_t rows = n * b;
    const int64_t blocks = ceil_div(rows, 8);
    const int64_t cap = (int64_t)sm_count() * 4;
    fuse_out_bwd_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, smem, (cudaStream_t)stream>>>(
        d_y, feat, n, b, gdim, c, fcw, d_feat, d_fcw, d_fcb);
    count_launch();
    return check_launch("fuse_out_bwd");
}

}  // extern "C"
