// K2 on the 5th-gen tensor cores (p = q = 64, the reference's lstm_hidden_dim / gcn_hidden_dim, Main.py:62-63):
//   forward : out[128 x 64]  = act( [T_0X | T_1X | ... | T_KX][128 x Ks*64] . W[Ks*64 x 64] + b )   (GCN.py:37-42)
//             -- the A operand is read segment by segment straight from the Chebyshev stack the SpMM steps wrote
//             (no torch.cat), split to tf32 hi/lo in the loader, accumulated in TMEM (3xTF32, tc_common.cuh)
//   backward: dZ = dOut (.) [out > 0] is formed in the loader (and written out for the weight-gradient kernel, with
//             the bias gradient as a by-product);  U[128 x Ks*64] = dZ[128 x 64] . W^T  -> U_k segments
// Same CTA anatomy as lstm_tc.cu: 8 loader warps (register ping-pong), 1 MMA warp, 4 or 8 epilogue warps.
#include "tc_pipeline.cuh"

using namespace stmgcn;
using namespace stmgcn::tc;

namespace {

constexpr int kPLoaderWarps = 8;
constexpr int kPLoaders = kPLoaderWarps * 32;
constexpr int kMaxSeg = 8;

template <int N>
struct PCfg {
    static constexpr int kEpiWarps = (N == 64) ? 4 : 8;
    static constexpr int kThreads = (kEpiWarps + kPLoaderWarps + 1) * 32;
    static constexpr int kBBytes = N * kKB * 4;
    static constexpr int kStageBytes = 2 * kABytes + 2 * kBBytes;
    static constexpr int kStages = (N == 64) ? 4 : 2;
    static constexpr int kTmemCols = (2 * N < 32) ? 32 : 2 * N;
};

struct PTail {
    float bias[64];
    float s_db[kPLoaderWarps][64];
    Barriers bar;
};
template <int N>
constexpr size_t psmem() { return 1024 + (size_t)PCfg<N>::kStages * PCfg<N>::kStageBytes + sizeof(PTail); }

struct PParams {
    const float* seg[kMaxSeg];   // forward: A segments (rows x 64)
    int nkb;                     // k-blocks (2 per 64-wide segment)
    // backward (dz mode): A = d_out (.) [out > 0]
    const float* d_out;          // (rows, 64) or nullptr
    const float* out_act;        // (rows, 64) forward output (mask source)
    int act;
    float* dz_out;               // (rows, 64) or nullptr (second pass of a > 4-support backward: dZ is already stored)
    float* dbias;                // (64) += or nullptr
    const float* wimg;
    const float* bias;           // forward epilogue
    float* out;                  // forward: (rows, 64)
    float* u;                    // backward: U_k = u + k*stride_u, (rows, 64) each
    int64_t stride_u;
    int ks_out;                  // backward: number of valid U segments (<= 4)
    int64_t rows;
    int n_tiles;
};

template <int N, bool DZ>
__global__ void __launch_bounds__(PCfg<N>::kThreads, 1) proj_rows_tc_kernel(const __grid_constant__ PParams p) {
    using Cfg = PCfg<N>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the __shared__ address space (LDS/STS, not generic LD/ST)
    PTail* tail = (PTail*)(smem + (size_t)Cfg::kStages * Cfg::kStageBytes);
    Barriers* bar = &tail->bar;
    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;
    constexpr int kMmaWarp = Cfg::kEpiWarps + kPLoaderWarps;

    if (tid == 0) init_barriers(bar, Cfg::kStages, Cfg::kEpiWarps * 32, kPLoaders / 2);   // one loader group per k-block
    if (warp == kMmaWarp) tmem_alloc(&bar->tmem_base, Cfg::kTmemCols);
    for (int i = tid; i < 64; i += Cfg::kThreads) tail->bias[i] = (!DZ && p.bias) ? p.bias[i] : 0.f;
    for (int i = tid; i < kPLoaderWarps * 64; i += Cfg::kThreads) (&tail->s_db[0][0])[i] = 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bar->tmem_base;

    if (warp >= Cfg::kEpiWarps && warp < kMmaWarp) {
        // ===================== loaders: two groups alternate k-blocks (see lstm_tc.cu on the proxy fence) ==========
        TC_PROF_DECL
        constexpr int kGroups = 2, kGT = kPLoaders / kGroups, kPer = 1024 / kGT;
        const int ltid = tid - Cfg::kEpiWarps * 32;
        const int grp = ltid / kGT, gtid = ltid % kGT;
        const int c = gtid & 7, rsub = gtid >> 3, lwarp = ltid >> 5;
        const int my_tiles = (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
        const int total = my_tiles * p.nkb;
        for (int j = grp; j < total; j += kGroups) {
            const int tile = blockIdx.x + (j / p.nkb) * gridDim.x, kb = j % p.nkb;
            const int koff = (kb & 1) * kKB + c * 4;
            float4 v[kPer];
            if (DZ) {
                float4 sb = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int i = 0; i < kPer; ++i) {
                    const int64_t r = (int64_t)tile * kTileM + rsub + (kGT / 8) * i;
                    float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (r < p.rows) {
                        d = *reinterpret_cast<const float4*>(p.d_out + r * 64 + koff);
                        if (p.act == STMGCN_ACT_RELU) {
                            const float4 m = *reinterpret_cast<const float4*>(p.out_act + r * 64 + koff);
                            if (!(m.x > 0.f)) d.x = 0.f;
                            if (!(m.y > 0.f)) d.y = 0.f;
                            if (!(m.z > 0.f)) d.z = 0.f;
                            if (!(m.w > 0.f)) d.w = 0.f;
                        }
                    }
                    v[i] = d;
                    sb.x += d.x; sb.y += d.y; sb.z += d.z; sb.w += d.w;
                }
#pragma unroll
                for (int o = 8; o <= 16; o <<= 1) {
                    sb.x += __shfl_xor_sync(0xffffffffu, sb.x, o); sb.y += __shfl_xor_sync(0xffffffffu, sb.y, o);
                    sb.z += __shfl_xor_sync(0xffffffffu, sb.z, o); sb.w += __shfl_xor_sync(0xffffffffu, sb.w, o);
                }
                if (lane < 8) {
                    float4* acc = reinterpret_cast<float4*>(&tail->s_db[lwarp][koff]);
                    float4 t = *acc;
                    t.x += sb.x; t.y += sb.y; t.z += sb.z; t.w += sb.w;
                    *acc = t;
                }
            } else {
                const float* seg = p.seg[kb >> 1];
#pragma unroll
                for (int i = 0; i < kPer; ++i) {
                    const int64_t r = (int64_t)tile * kTileM + rsub + (kGT / 8) * i;
                    v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (seg != nullptr && r < p.rows) v[i] = *reinterpret_cast<const float4*>(seg + r * 64 + koff);
                }
            }
            const int s = j % Cfg::kStages;
            const uint32_t ph = (j / Cfg::kStages) & 1;
            mbar_wait(&bar->empty[s], ph ^ 1, 0);
            uint8_t* st = smem + (size_t)s * Cfg::kStageBytes;
            if (gtid == 0) {
                mbar_arrive_expect_tx(&bar->full[s], 2 * Cfg::kBBytes);
                const float* src = p.wimg + (size_t)kb * (2 * Cfg::kBBytes / 4);
                bulk_g2s(st + 2 * kABytes, src, Cfg::kBBytes, &bar->full[s]);
                bulk_g2s(st + 2 * kABytes + Cfg::kBBytes, src + Cfg::kBBytes / 4, Cfg::kBBytes, &bar->full[s]);
            }
#pragma unroll
            for (int i = 0; i < kPer; ++i) {
                const int row = rsub + (kGT / 8) * i;
                split_store(st, (uint32_t)row * 128u + (uint32_t)((c ^ (row & 7)) << 4), v[i]);
            }
            fence_proxy_async_smem();
            mbar_arrive(&bar->full[s]);
            if (DZ) {          // dZ tape for the weight-gradient kernel, stored after the hand-off (see lstm_tc.cu)
#pragma unroll
                for (int i = 0; i < kPer; ++i) {
                    const int64_t r = (int64_t)tile * kTileM + rsub + (kGT / 8) * i;
                    if (r < p.rows && p.dz_out != nullptr) *reinterpret_cast<float4*>(p.dz_out + r * 64 + koff) = v[i];
                }
            }
        }
        TC_PROF_FLUSH(11, ltid == 0)
    } else if (warp == kMmaWarp) {
        mma_issuer<N, Cfg::kStages, 4>(bar, smem, Cfg::kStageBytes, Cfg::kBBytes, p.nkb, p.n_tiles, tmem_base, lane);
    } else {
        // ===================== epilogue =====================
        TC_PROF_DECL
        const int q = warp & 3, part = warp >> 2;          // N = 256: two column halves
        uint32_t tcount = 0;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++tcount) {
            const int a = tcount & 1;
            const uint32_t aph = (tcount >> 1) & 1;
            const int64_t r = (int64_t)tile * kTileM + q * 32 + lane;
            const bool valid = r < p.rows;
            mbar_wait(&bar->tmem_full[a], aph, 3);
            tc_fence_after();
            const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)a * N;
            constexpr int kChunksPerWarp = (N == 64) ? 2 : 4;
#pragma unroll 1
            for (int ci = 0; ci < kChunksPerWarp; ++ci) {
                const int chunk = part * kChunksPerWarp + ci;
                uint32_t v[32];
                tmem_ld32(t_row + chunk * 32, v);
                tmem_ld_wait();
                if (!valid) continue;
                if (N == 64) {
                    float* dst = p.out + r * 64 + chunk * 32;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float4 o;
                        o.x = __uint_as_float(v[4 * j + 0]) + tail->bias[chunk * 32 + 4 * j + 0];
                        o.y = __uint_as_float(v[4 * j + 1]) + tail->bias[chunk * 32 + 4 * j + 1];
                        o.z = __uint_as_float(v[4 * j + 2]) + tail->bias[chunk * 32 + 4 * j + 2];
                        o.w = __uint_as_float(v[4 * j + 3]) + tail->bias[chunk * 32 + 4 * j + 3];
                        if (p.act == STMGCN_ACT_RELU) {
                            o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
                        }
                        *reinterpret_cast<float4*>(dst + 4 * j) = o;
                    }
                } else {
                    const int col = chunk * 32;                    // U_k, k = col / 64
                    if ((col >> 6) >= p.ks_out) continue;
                    float* dst = p.u + (int64_t)(col >> 6) * p.stride_u + r * 64 + (col & 63);
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        *reinterpret_cast<uint4*>(dst + 4 * j) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                }
            }
            tc_fence_before();
            mbar_arrive(&bar->tmem_empty[a]);
        }
        TC_PROF_FLUSH(12, tid == 0)
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == kMmaWarp) tmem_dealloc(tmem_base, Cfg::kTmemCols);
    if (DZ && p.dbias != nullptr) {
        for (int i = tid; i < 64; i += Cfg::kThreads) {
            float v = 0.f;
#pragma unroll
            for (int w = 0; w < kPLoaderWarps; ++w) v += tail->s_db[w][i];
            atomicAdd(&p.dbias[i], v);
        }
    }
}

}  // namespace

namespace stmgcn {

bool proj_tc_applicable(int ks, int p, int q, const void* a, const void* b, const void* c) {
    return p == 64 && q == 64 && ks >= 1 && ks <= 8 && aligned16(a) && aligned16(b) && (!c || aligned16(c));
}

// forward: out = act(sum_k S_k W_k + bias); wimg = image of B[n][k] = W[k][n] (2*ks k-blocks of [hi|lo] [64][32])
int32_t launch_proj_fwd_tc(const float* s, int64_t stride_k, int ks, int64_t rows, const float* wimg, const float* bias,
                           int act, float* out, cudaStream_t st) {
    auto kern = proj_rows_tc_kernel<64, false>;
    if (int32_t rc = ensure_dyn_smem((const void*)kern, psmem<64>())) return rc;
    PParams p{};
    for (int k = 0; k < ks; ++k) p.seg[k] = s + (int64_t)k * stride_k;
    p.nkb = 2 * ks;
    p.wimg = wimg;
    p.bias = bias;
    p.act = act;
    p.out = out;
    p.rows = rows;
    p.n_tiles = (int)ceil_div(rows, kTileM);
    const int grid = p.n_tiles < sm_count() ? p.n_tiles : sm_count();
    kern<<<grid, PCfg<64>::kThreads, psmem<64>(), st>>>(p);
    count_launch();
    return check_launch("proj_fwd_tc");
}

// backward data: dZ = d_out (.) mask (written to dz_out, bias gradient accumulated), U_k = dZ W_k^T (if u != nullptr);
// wimg_t = image of B[n = k*64+i][k' = j] = W[n][j]  (2 k-blocks of [hi|lo] [ks*64][32])
int32_t launch_proj_bwd_tc(const float* d_out, const float* out_act, int act, int64_t rows, int ks, const float* wimg_t,
                           float* dz_out, float* dbias, float* u, int64_t stride_u, cudaStream_t st) {
    auto kern = proj_rows_tc_kernel<256, true>;
    if (int32_t rc = ensure_dyn_smem((const void*)kern, psmem<256>())) return rc;
    PParams p{};
    p.nkb = 2;
    p.ks_out = ks;
    p.d_out = d_out;
    p.out_act = out_act;
    p.act = act;
    p.dz_out = dz_out;
    p.dbias = dbias;
    p.wimg = wimg_t;
    p.u = u;
    p.stride_u = stride_u;
    p.rows = rows;
    p.n_tiles = (int)ceil_div(rows, kTileM);
    const int grid = p.n_tiles < sm_count() ? p.n_tiles : sm_count();
    kern<<<grid, PCfg<256>::kThreads, psmem<256>(), st>>>(p);
    count_launch();
    return check_launch("proj_bwd_tc");
}

int32_t launch_pack_image(const float* src, int n_rows, int k_cols, int64_t rs, int64_t cs, float* img, int tile_rows,
                          cudaStream_t st) {
    pack_image_kernel<<<(n_rows * k_cols + 255) / 256, 256, 0, st>>>(src, n_rows, k_cols, rs, cs, img, tile_rows);
    count_launch();
    return check_launch("pack_image");
}

}  // namespace stmgcn
