// K3b on the 5th-gen tensor cores: one LSTM layer-step for a 128-row tile as
//     gates[128 x 256] = [h_below | h_prev][128 x Kd] . Wp[Kd x 256]        (reference: nn.LSTM, STMGCN.py:48)
// with tcgen05.mma kind::tf32 in the 3xTF32 scheme (tc_common.cuh), fp32 accumulators in TMEM, and the cell
// nonlinearity + state update fused into the TMEM->register epilogue.  H = 64 only (the reference's value,
// Main.py:62); other sizes take the exact-FFMA path in lstm.cu.
//
// CTA = 9 warps, persistent over row tiles, 1 CTA / SM (192 KB of operand stages, all 512 TMEM columns):
//   warps 0-3  epilogue : tcgen05.ld its TMEM lane quadrant, bias + layer-0 input term, sigmoid/tanh, c/h update,
//                         stores h, c (and the gate tape for the backward)
//   warps 4-7  loaders  : read the A rows from HBM (coalesced 128-bit), split into tf32 hi/lo and write both
//                         K-major 128B-swizzled operand tiles; one thread also starts the bulk copies (TMA unit)
//                         of the pre-swizzled hi/lo weight images for the same k-block
//   warp  8    MMA      : single-thread tcgen05.mma issue, 12 MMAs (3 passes x 4 k-slices) per 32-wide k-block;
//                         tcgen05.commit releases operand stages and publishes accumulators
// Pipelines: 2 operand stages (full/empty mbarriers), 2 TMEM accumulators (tmem_full/tmem_empty).
#include "tc_common.cuh"

using namespace stmgcn;
using namespace stmgcn::tc;

namespace {

constexpr int kTileM = 128;
constexpr int kTileN = 256;          // 4H, H = 64
constexpr int kKB = 32;              // k-block: one 128-byte swizzle row of fp32
constexpr int kHid = 64;
constexpr int kMaxC = 4;
constexpr int kStages = 2;
constexpr int kAccs = 2;
constexpr int kABytes = kTileM * kKB * 4;            // 16 KB per hi or lo tile
constexpr int kBBytes = kTileN * kKB * 4;            // 32 KB per hi or lo tile
constexpr int kStageBytes = 2 * kABytes + 2 * kBBytes;   // 96 KB
constexpr int kThreads = 288;
constexpr int kNumLoaders = 128;
constexpr int kNumEpi = 128;

struct SmemTail {
    float bias[kTileN];
    float wx[kMaxC * kTileN];
    uint64_t full[kStages];
    uint64_t empty[kStages];
    uint64_t tmem_full[kAccs];
    uint64_t tmem_empty[kAccs];
    uint32_t tmem_base;
};
constexpr size_t kSmemBytes = 1024 /*align slack*/ + (size_t)kStages * kStageBytes + sizeof(SmemTail);

struct CellParams {
    const float* seg0;       // (rows, 64) first K segment  (h_below for l>0, h_prev for l==0) or nullptr = zeros
    const float* seg1;       // (rows, 64) second K segment (h_prev for l>0) or nullptr
    int nkb;                 // k-blocks: 2 per segment
    const float* wimg;       // nkb x [hi 32 KB | lo 32 KB] pre-swizzled weight images
    const float* bias;       // (256)
    const float* wx;         // (C,256) layer 0 only, else nullptr
    const float* xo;         // (rows, T, C)
    const float* sg;         // (B, T)
    int c_in, t, t_len;
    int64_t b_inner;
    const float* c_prev;     // (rows,64) or nullptr
    float* h_out;
    float* c_out;
    float* gates_out;        // (rows,256) or nullptr
    int64_t rows;
    int n_tiles;
};

__global__ void __launch_bounds__(kThreads, 1) lstm_cell_tc_kernel(const __grid_constant__ CellParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    SmemTail* tail = (SmemTail*)(smem + (size_t)kStages * kStageBytes);
    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;

    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&tail->full[s], kNumLoaders + 1);
            mbar_init(&tail->empty[s], 1);
        }
        for (int a = 0; a < kAccs; ++a) {
            mbar_init(&tail->tmem_full[a], 1);
            mbar_init(&tail->tmem_empty[a], kNumEpi);
        }
        fence_barrier_init();
    }
    if (warp == 8) tmem_alloc(&tail->tmem_base, 512);
    for (int i = tid; i < kTileN; i += kThreads) tail->bias[i] = p.bias[i];
    if (p.wx != nullptr)
        for (int i = tid; i < p.c_in * kTileN; i += kThreads) tail->wx[i] = p.wx[i];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tail->tmem_base;

    if (warp >= 4 && warp < 8) {
        // ===================== loaders / tf32 splitters =====================
        const int ltid = tid - 128;
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
            const int64_t row_base = (int64_t)tile * kTileM;
            for (int kb = 0; kb < p.nkb; ++kb, ++it) {
                const int s = it & 1;
                const uint32_t ph = (it >> 1) & 1;
                mbar_wait(&tail->empty[s], ph ^ 1);
                uint8_t* st = smem + (size_t)s * kStageBytes;
                if (ltid == 0) {
                    mbar_arrive_expect_tx(&tail->full[s], 2 * kBBytes);
                    const float* src = p.wimg + (size_t)kb * (2 * kBBytes / 4);
                    bulk_g2s(st + 2 * kABytes, src, kBBytes, &tail->full[s]);
                    bulk_g2s(st + 2 * kABytes + kBBytes, src + kBBytes / 4, kBBytes, &tail->full[s]);
                }
                const float* seg = (kb >> 1) ? p.seg1 : p.seg0;
                const int koff = (kb & 1) * kKB;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int idx = ltid + i * kNumLoaders;
                    const int row = idx >> 3, c = idx & 7;
                    const int64_t r = row_base + row;
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (seg != nullptr && r < p.rows)
                        v = *reinterpret_cast<const float4*>(seg + r * kHid + koff + c * 4);
                    float4 hi, lo;
                    hi.x = tf32_hi(v.x); hi.y = tf32_hi(v.y); hi.z = tf32_hi(v.z); hi.w = tf32_hi(v.w);
                    lo.x = tf32_lo(v.x, hi.x); lo.y = tf32_lo(v.y, hi.y); lo.z = tf32_lo(v.z, hi.z); lo.w = tf32_lo(v.w, hi.w);
                    const uint32_t off = (uint32_t)row * 128u + (uint32_t)((c ^ (row & 7)) << 4);
                    *reinterpret_cast<float4*>(st + off) = hi;
                    *reinterpret_cast<float4*>(st + kABytes + off) = lo;
                }
                fence_proxy_async_smem();
                mbar_arrive(&tail->full[s]);
            }
        }
    } else if (warp == 8) {
        // ===================== MMA issuer =====================
        constexpr uint32_t idesc = idesc_tf32(kTileM, kTileN);
        uint32_t it = 0, tcount = 0;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++tcount) {
            const int a = tcount & 1;
            const uint32_t aph = (tcount >> 1) & 1;
            mbar_wait(&tail->tmem_empty[a], aph ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)a * kTileN;
            for (int kb = 0; kb < p.nkb; ++kb, ++it) {
                const int s = it & 1;
                const uint32_t ph = (it >> 1) & 1;
                mbar_wait(&tail->full[s], ph);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t st = smem_u32(smem + (size_t)s * kStageBytes);
                    const uint64_t a_hi = smem_desc_k_sw128(st);
                    const uint64_t a_lo = smem_desc_k_sw128(st + kABytes);
                    const uint64_t b_hi = smem_desc_k_sw128(st + 2 * kABytes);
                    const uint64_t b_lo = smem_desc_k_sw128(st + 2 * kABytes + kBBytes);
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
                    mma_commit(&tail->empty[s]);
                }
                __syncwarp();
            }
            if (lane == 0) mma_commit(&tail->tmem_full[a]);
            __syncwarp();
        }
    } else {
        // ===================== epilogue: LSTM cell =====================
        const int etid = tid;            // 0..127 == TMEM lane == row in tile
        uint32_t tcount = 0;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++tcount) {
            const int a = tcount & 1;
            const uint32_t aph = (tcount >> 1) & 1;
            const int64_t r = (int64_t)tile * kTileM + etid;
            const bool valid = r < p.rows;
            float xs[kMaxC];
#pragma unroll
            for (int c = 0; c < kMaxC; ++c) xs[c] = 0.f;
            if (p.wx != nullptr && valid) {
                const float sv = p.sg[(r % p.b_inner) * p.t_len + p.t];
#pragma unroll
                for (int c = 0; c < kMaxC; ++c)
                    if (c < p.c_in) xs[c] = p.xo[(r * p.t_len + p.t) * p.c_in + c] * sv;
            }
            mbar_wait(&tail->tmem_full[a], aph);
            tc_fence_after();
            const uint32_t t_row = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)a * kTileN;
#pragma unroll 1
            for (int chunk = 0; chunk < kTileN / 32; ++chunk) {
                uint32_t v[32];
                tmem_ld32(t_row + chunk * 32, v);
                tmem_ld_wait();
                if (valid) {
                    float cp[8];
                    if (p.c_prev != nullptr) {
                        const float4 c0 = *reinterpret_cast<const float4*>(p.c_prev + r * kHid + chunk * 8);
                        const float4 c1 = *reinterpret_cast<const float4*>(p.c_prev + r * kHid + chunk * 8 + 4);
                        cp[0] = c0.x; cp[1] = c0.y; cp[2] = c0.z; cp[3] = c0.w;
                        cp[4] = c1.x; cp[5] = c1.y; cp[6] = c1.z; cp[7] = c1.w;
                    } else {
#pragma unroll
                        for (int u = 0; u < 8; ++u) cp[u] = 0.f;
                    }
                    float hn[8], cn[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int col = chunk * 32 + 4 * u;
                        float pi = __uint_as_float(v[4 * u + 0]) + tail->bias[col + 0];
                        float pf = __uint_as_float(v[4 * u + 1]) + tail->bias[col + 1];
                        float pg = __uint_as_float(v[4 * u + 2]) + tail->bias[col + 2];
                        float po = __uint_as_float(v[4 * u + 3]) + tail->bias[col + 3];
                        if (p.wx != nullptr) {
#pragma unroll
                            for (int c = 0; c < kMaxC; ++c) {
                                if (c < p.c_in) {
                                    pi = fmaf(xs[c], tail->wx[c * kTileN + col + 0], pi);
                                    pf = fmaf(xs[c], tail->wx[c * kTileN + col + 1], pf);
                                    pg = fmaf(xs[c], tail->wx[c * kTileN + col + 2], pg);
                                    po = fmaf(xs[c], tail->wx[c * kTileN + col + 3], po);
                                }
                            }
                        }
                        const float gi = sigmoidf_(pi), gf = sigmoidf_(pf), gg = tanhf_(pg), go = sigmoidf_(po);
                        cn[u] = fmaf(gf, cp[u], gi * gg);
                        hn[u] = go * tanhf_(cn[u]);
                        if (p.gates_out != nullptr)
                            *reinterpret_cast<float4*>(p.gates_out + r * kTileN + col) = make_float4(gi, gf, gg, go);
                    }
                    float* hd = p.h_out + r * kHid + chunk * 8;
                    float* cd = p.c_out + r * kHid + chunk * 8;
                    *reinterpret_cast<float4*>(hd) = make_float4(hn[0], hn[1], hn[2], hn[3]);
                    *reinterpret_cast<float4*>(hd + 4) = make_float4(hn[4], hn[5], hn[6], hn[7]);
                    *reinterpret_cast<float4*>(cd) = make_float4(cn[0], cn[1], cn[2], cn[3]);
                    *reinterpret_cast<float4*>(cd + 4) = make_float4(cn[4], cn[5], cn[6], cn[7]);
                }
            }
            tc_fence_before();
            mbar_arrive(&tail->tmem_empty[a]);
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 8) tmem_dealloc(tmem_base, 512);
}

// wp (kd, 256) [k][n] -> per k-block: hi image [256][32] then lo image, both K-major with the 128-byte swizzle
__global__ void pack_weights_tc_kernel(const float* __restrict__ wp, int kd, float* __restrict__ img) {
    const int total = kd * kTileN;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int k = e / kTileN, n = e % kTileN;
        const float v = wp[e];
        const float hi = tf32_hi(v);
        const float lo = tf32_lo(v, hi);
        const int kb = k / kKB, kk = k % kKB;
        const uint32_t off = sw128_offset((uint32_t)n, (uint32_t)kk) / 4;
        float* base = img + (size_t)kb * (2 * kBBytes / 4);
        base[off] = hi;
        base[kBBytes / 4 + off] = lo;
    }
}

}  // namespace

namespace stmgcn {

// Called from stmgcn_lstm_step_fwd (lstm.cu) when the tensor-core path applies.
int32_t launch_lstm_cell_tc(const float* seg0, const float* seg1, int nseg, const float* wimg, const float* bias,
                            const float* wx, const float* xo, const float* sg, int c_in, int t, int t_len,
                            int64_t b_inner, const float* c_prev, float* h_out, float* c_out, float* gates_out,
                            int64_t rows, cudaStream_t st) {
    static bool attr_done = false;
    if (!attr_done) {
        STMGCN_CUDA(cudaFuncSetAttribute(lstm_cell_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)kSmemBytes));
        attr_done = true;
    }
    CellParams p;
    p.seg0 = seg0;
    p.seg1 = seg1;
    p.nkb = 2 * nseg;
    p.wimg = wimg;
    p.bias = bias;
    p.wx = wx;
    p.xo = xo;
    p.sg = sg;
    p.c_in = c_in;
    p.t = t;
    p.t_len = t_len;
    p.b_inner = b_inner;
    p.c_prev = c_prev;
    p.h_out = h_out;
    p.c_out = c_out;
    p.gates_out = gates_out;
    p.rows = rows;
    p.n_tiles = (int)ceil_div(rows, kTileM);
    const int grid = p.n_tiles < sm_count() ? p.n_tiles : sm_count();
    lstm_cell_tc_kernel<<<grid, kThreads, kSmemBytes, st>>>(p);
    count_launch();
    return check_launch("lstm_cell_tc");
}

}  // namespace stmgcn

extern "C" int32_t stmgcn_lstm_pack_tc(const float* wp, int32_t kd, int32_t hid, float* img, void* stream) {
    STMGCN_REQUIRE(wp && img, STMGCN_ERR_ARG, "lstm_pack_tc: null pointer");
    STMGCN_REQUIRE(hid == kHid && kd > 0 && kd % kKB == 0, STMGCN_ERR_SHAPE,
                   "lstm_pack_tc: tensor-core path needs hid == 64 and kd %% 32 == 0 (got hid=%d kd=%d)", hid, kd);
    const int total = kd * kTileN;
    pack_weights_tc_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(wp, kd, img);
    count_launch();
    return check_launch("lstm_pack_tc");
}
