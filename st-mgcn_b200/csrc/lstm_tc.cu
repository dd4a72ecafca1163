// K3b on the 5th-gen tensor cores (tcgen05, TMEM accumulators, 3xTF32 -- see tc_common.cuh), H = 64 only
// (the reference's value, Main.py:62); other sizes take the exact-FFMA kernels in lstm.cu.
//
//  forward  (one launch per layer-step):  gates[128 x 256] = [h_below | h_prev][128 x Kd] . Wp[Kd x 256]
//           + fused LSTM cell epilogue (reference: nn.LSTM inside CG_LSTM, STMGCN.py:48).
//  backward (one launch per layer-step):  the loader warps turn the saved gates into the pre-activation
//           gradients dA (BPTT pointwise math), write dA back over the tape (consumed by stmgcn_lstm_wgrad) and
//           straight into the swizzled operand tiles of  [dx_below | dh_prev][128 x Kd] = dA[128 x 256] . Wp^T.
//
// CTA anatomy (persistent over 128-row tiles, 1 CTA / SM):
//   producer warp  : one thread issues every global->shared transfer through the TMA unit: tensor loads of raw A
//                    tiles / gate slices (CUtensorMap built per launch), 4 KB bulk copies of tile-blocked workspace
//                    slices, bulk copies (UBLKCP) of the pre-swizzled hi/lo weight images
//   loader warps   : shared memory -> registers -> (backward: BPTT pointwise math) -> tf32 hi/lo split -> K-major
//                    128B-swizzled operand tiles.  (Register-load variants, where these warps read HBM themselves,
//                    remain as fallbacks: STMGCN_FWD_TMA=0 / STMGCN_BWD_TMA=0, row-major workspaces.)
//   MMA warp       : one thread issues tcgen05.mma kind::tf32, 12 per 32-wide k-block (3 passes x 4 k-slices);
//                    tcgen05.commit releases operand stages / publishes accumulators
//   epilogue warps : tcgen05.ld their TMEM lane quadrant -> registers -> math -> forward: staging tile -> TMA tensor
//                    stores (gate tape, h, c); backward: direct stores of [dx_below | dh_prev]
// Pipelines: operand stages (full/empty mbarriers), raw slots (raw_full/raw_empty), two TMEM accumulators.
#include "tc_pipeline.cuh"
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

using namespace stmgcn;
using namespace stmgcn::tc;

namespace {

constexpr int kNumLoaders = 512;     // 16 loader warps (backward): the loaders are latency-bound, TLP is what helps
constexpr int kLoaderWarps = kNumLoaders / 32;

// element (row r, unit u) of a (rows x 64) workspace: row-major or tile-blocked (each 8-unit k-block slice of a
// 128-row tile is one contiguous 4 KB run, so the loader's reads and the epilogue's writes are full lines)
__device__ __forceinline__ int64_t ws_off(int blocked, int64_t r, int unit) {
    return blocked ? ((((r >> 7) * 8 + (unit >> 3)) * kTileM + (r & 127)) * 8 + (unit & 7)) : r * kHid + unit;
}


// =====================================================================================================
// forward cell
// =====================================================================================================
constexpr int kFwdN = kGateCols;
constexpr int kFwdStages = 2;
constexpr int kFwdBBytes = kFwdN * kKB * 4;                       // 32 KB
constexpr int kFwdStageBytes = 2 * kABytes + 2 * kFwdBBytes;      // 96 KB
constexpr int kFwdEpiWarps = 16;                                  // 4 per TMEM lane quadrant: the epilogue is
// 8 loader warps.  Register loads from HBM: two groups alternate k-blocks (their load latencies overlap).  TMA-fed: one
// group of all 8 warps per k-block (nothing to overlap, the split latency of a k-block halves); 4 warps were measured
// slower (6.55 vs 5.42 ms per branch forward).
template <bool A_TMA> struct FwdCfg {
    static constexpr int kLoaderWarps = 8;
    static constexpr int kGroups = A_TMA ? 1 : 2;
    static constexpr int kLoaders = kLoaderWarps * 32;
    static constexpr int kThreads = (kFwdEpiWarps + kLoaderWarps + 2) * 32;   // + MMA warp + TMA producer warp
};
constexpr int kStagingBytes = 32 * 16 * 4;                        // per epilogue warp: [32 rows][16 cols] fp32

struct FwdTail {
    float bias[kGateCols];
    Barriers bar;
    uint64_t a_full[kFwdStages];      // TMA-fed variant: the raw A tile of the stage has landed
};
constexpr size_t kFwdSmem = 1024 + (size_t)kFwdStages * kFwdStageBytes + (size_t)kFwdEpiWarps * kStagingBytes +
                            sizeof(FwdTail);

static_assert(kFwdSmem <= 232448, "forward kernel exceeds the 227 KB shared-memory limit");

struct CellParams {
    const float* seg0;       // (rows, 64) first K segment  (h_below for l>0, h_prev for l==0) or nullptr = zeros
    const float* seg1;       // (rows, 64) second K segment (h_prev for l>0) or nullptr
    int nkb;                 // k-blocks: 2 per segment (+1 auxiliary block for layer 0)
    int aux;                 // layer 0: k-block 2 holds [x*s (C cols) | 1 | 0...] so that x.W_ih + b runs on the MMA
    const float* wimg;       // nkb x [hi 32 KB | lo 32 KB] pre-swizzled weight images
    const float* bias;       // (256) or nullptr when folded into the auxiliary block
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
    int blocked_cs;          // c_prev / c_out use the tile-blocked layout (see ws_off)
    int prefetch;            // bulk L2 prefetch of the next tile's inputs (STMGCN_TC_PREFETCH=1; default off)
    int gates_tma;           // gate tape leaves through TMA tensor stores (gates_map) instead of per-thread stores
    int dbg_skip_hc;         // timing experiments only (STMGCN_DBG_SKIP_HC=1): drop the h / c stores (wrong results)
    alignas(64) CUtensorMap gates_map;   // (rows, 256) fp32 slice of the tape, box 32 rows x 16 columns, 64-byte swizzle
    int hc_tma;                          // h / c leave through TMA tensor stores too (h_map, c_map; needs gates_tma)
    alignas(64) CUtensorMap h_map;       // (rows, 64) slice of hs, box 32 rows x 4 units
    alignas(64) CUtensorMap c_map;       // cs slice: (rows, 64) row-major, or (rows_pad * 8, 8) when tile-blocked
    alignas(64) CUtensorMap seg0_map;    // A_TMA: (rows, 64) slices of the hidden-state tape, box 128 rows x 32 columns,
    alignas(64) CUtensorMap seg1_map;    //        128-byte swizzle == the K-major operand tile layout
};

// A_TMA: the raw fp32 A tile of every k-block is written by ONE TMA tensor load straight into the stage's hi tile (the
// 128-byte TMA swizzle is the operand layout); the loader warps then read it back from shared memory, and write the
// tf32 hi part in place and the lo part next to it -- no per-thread global loads (whose latency the proxy fence's MEMBAR
// would expose), the load is issued the moment the MMA warp frees the stage.
template <bool GATES_TMA, bool A_TMA>
__global__ void __launch_bounds__((FwdCfg<A_TMA>::kThreads), 1) lstm_cell_tc_kernel(const __grid_constant__ CellParams p) {
    constexpr int kFwdLoaderWarps = FwdCfg<A_TMA>::kLoaderWarps;
    constexpr int kFwdLoaders = FwdCfg<A_TMA>::kLoaders;
    constexpr int kFwdThreads = FwdCfg<A_TMA>::kThreads;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the __shared__ address space (LDS/STS, not generic LD/ST)
    uint8_t* staging = smem + (size_t)kFwdStages * kFwdStageBytes;
    FwdTail* tail = (FwdTail*)(staging + (size_t)kFwdEpiWarps * kStagingBytes);
    Barriers* bar = &tail->bar;
    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;
    constexpr int kMmaWarp = kFwdEpiWarps + kFwdLoaderWarps;

    pdl_launch_dependents();
    if (tid == 0) {
        for (int s = 0; s < kFwdStages; ++s) mbar_init(&tail->a_full[s], 1);
        init_barriers(bar, kFwdStages, kFwdEpiWarps * 32, kFwdLoaders / FwdCfg<A_TMA>::kGroups);   // one loader group per k-block
    }
    if (warp == kMmaWarp) tmem_alloc(&bar->tmem_base, 512);
    for (int i = tid; i < kGateCols; i += kFwdThreads) tail->bias[i] = p.bias ? p.bias[i] : 0.f;
    pdl_wait();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bar->tmem_base;

    if (warp >= kFwdEpiWarps && warp < kMmaWarp) {
        // ===================== loaders / tf32 splitters =====================
        // NOTE fence.proxy.async compiles to MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC: the MEMBAR waits for every outstanding
        // memory operation of the thread, so loads prefetched by the same thread for a later k-block would be waited
        // for at the current k-block's fence.  Memory-level parallelism therefore comes from two independent loader
        // GROUPS that alternate k-blocks (group g owns k-blocks g, g+2, ... of this CTA's sequence): while one group
        // waits for its loads, the other converts / stores / fences.
        TC_PROF_DECL
        constexpr int kGroups = FwdCfg<A_TMA>::kGroups, kGT = kFwdLoaders / kGroups, kPer = 1024 / kGT;   // float4 per thread per k-block
        const int ltid = tid - kFwdEpiWarps * 32;
        const int grp = ltid / kGT, gtid = ltid % kGT;
        const int c = gtid & 7, rsub = gtid >> 3;          // rows rsub + (kGT/8)*i, 16-byte chunk c of the 128-byte row
        const int my_tiles = (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
        const int total = my_tiles * p.nkb;
        for (int j = grp; j < total; j += kGroups) {
            const int tile = blockIdx.x + (j / p.nkb) * gridDim.x, kb = j % p.nkb;
            constexpr int kChunk = kPer < 8 ? kPer : 8;    // float4 per thread in flight
            static_assert(kPer % kChunk == 0 && (A_TMA || kPer == kChunk), "loader mapping");
            const float* seg = (kb >> 1) ? p.seg1 : p.seg0;
            const bool aux_blk = p.aux && kb == 2;
            const bool from_tma = A_TMA && !aux_blk && seg != nullptr;
            const int s = j % kFwdStages;
            const uint32_t ph = (j / kFwdStages) & 1;
            uint8_t* st = smem + (size_t)s * kFwdStageBytes;
            if (A_TMA) mbar_wait(&tail->a_full[s], ph, 0);     // stage is free and (for TMA blocks) the raw tile is in place
#pragma unroll 1
            for (int h0 = 0; h0 < kPer; h0 += kChunk) {
                float4 buf[kChunk];
                if (aux_blk) {                             // auxiliary block: modulated input columns and the constant 1
#pragma unroll
                    for (int i = 0; i < kChunk; ++i) {
                        const int64_t r = (int64_t)tile * kTileM + rsub + (kGT / 8) * (h0 + i);
                        float v[4] = {0.f, 0.f, 0.f, 0.f};
                        if (c < 2 && r < p.rows) {
                            const float sv = p.sg[(r % p.b_inner) * p.t_len + p.t];
#pragma unroll
                            for (int jj = 0; jj < 4; ++jj) {
                                const int col = 4 * c + jj;
                                if (col < p.c_in) v[jj] = p.xo[(r * p.t_len + p.t) * p.c_in + col] * sv;
                                else if (col == p.c_in) v[jj] = 1.0f;
                            }
                        }
                        buf[i] = make_float4(v[0], v[1], v[2], v[3]);
                    }
                } else if (from_tma) {
#pragma unroll
                    for (int i = 0; i < kChunk; ++i) {
                        const int row = rsub + (kGT / 8) * (h0 + i);
                        buf[i] = *reinterpret_cast<const float4*>(st + (uint32_t)row * 128u + (uint32_t)((c ^ (row & 7)) << 4));
                    }
                } else {
                    const int koff = (kb & 1) * kKB + c * 4;
#pragma unroll
                    for (int i = 0; i < kChunk; ++i) {
                        const int64_t r = (int64_t)tile * kTileM + rsub + (kGT / 8) * (h0 + i);
                        buf[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (seg != nullptr && r < p.rows) buf[i] = *reinterpret_cast<const float4*>(seg + r * kHid + koff);
                    }
                }
                if (!A_TMA) {
                    mbar_wait(&bar->empty[s], ph ^ 1, 0);
                    if (gtid == 0) {
                        mbar_arrive_expect_tx(&bar->full[s], 2 * kFwdBBytes);
                        const float* src = p.wimg + (size_t)kb * (2 * kFwdBBytes / 4);
                        bulk_g2s(st + 2 * kABytes, src, kFwdBBytes, &bar->full[s]);
                        bulk_g2s(st + 2 * kABytes + kFwdBBytes, src + kFwdBBytes / 4, kFwdBBytes, &bar->full[s]);
                    }
                }
#pragma unroll
                for (int i = 0; i < kChunk; ++i) {
                    const int row = rsub + (kGT / 8) * (h0 + i);
                    split_store(st, (uint32_t)row * 128u + (uint32_t)((c ^ (row & 7)) << 4), buf[i]);
                }
            }
            fence_proxy_async_smem();
            mbar_arrive(&bar->full[s]);
            if (p.prefetch && gtid == 0 && kb == 0) {          // L2 prefetch of this CTA's next tile (contiguous rows)
                const int nt = tile + gridDim.x;
                if (nt < p.n_tiles) {
                    const int64_t r0 = (int64_t)nt * kTileM;
                    const int64_t nr = (p.rows - r0) < kTileM ? (p.rows - r0) : kTileM;
                    const uint32_t bh = (uint32_t)(nr * kHid * 4);
                    if (p.seg0) prefetch_l2(p.seg0 + r0 * kHid, bh);
                    if (p.seg1) prefetch_l2(p.seg1 + r0 * kHid, bh);
                    if (p.c_prev) prefetch_l2(p.c_prev + r0 * kHid, bh);     // (a 128-row tile is contiguous in both layouts)
                }
            }
        }
        TC_PROF_FLUSH(0, ltid == 0)
    } else if (warp == kMmaWarp) {
        mma_issuer<kFwdN, kFwdStages, 0>(bar, smem, kFwdStageBytes, kFwdBBytes, p.nkb, p.n_tiles, tmem_base, lane);
    } else if (warp == kMmaWarp + 1) {
        // ===================== producer (A_TMA): raw A tile + weight images for every freed stage =====================
        if (A_TMA && lane == 0) {
            const int my_tiles = (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
            const int total = my_tiles * p.nkb;
            for (int j = 0; j < total; ++j) {
                const int tile = blockIdx.x + (j / p.nkb) * gridDim.x, kb = j % p.nkb;
                const int s = j % kFwdStages;
                const uint32_t ph = (j / kFwdStages) & 1;
                mbar_wait_raw(&bar->empty[s], ph ^ 1);
                uint8_t* st = smem + (size_t)s * kFwdStageBytes;
                const float* seg = (kb >> 1) ? p.seg1 : p.seg0;
                if (!(p.aux && kb == 2) && seg != nullptr) {
                    mbar_arrive_expect_tx(&tail->a_full[s], kABytes);
                    tma_load_2d(st, (kb >> 1) ? &p.seg1_map : &p.seg0_map, (kb & 1) * kKB, tile * kTileM, &tail->a_full[s]);
                } else {
                    mbar_arrive(&tail->a_full[s]);         // the loaders build this block themselves
                }
                mbar_arrive_expect_tx(&bar->full[s], 2 * kFwdBBytes);
                const float* src = p.wimg + (size_t)kb * (2 * kFwdBBytes / 4);
                bulk_g2s(st + 2 * kABytes, src, kFwdBBytes, &bar->full[s]);
                bulk_g2s(st + 2 * kABytes + kFwdBBytes, src + kFwdBBytes / 4, kFwdBBytes, &bar->full[s]);
            }
        }
    } else {
        // ===================== epilogue: LSTM cell =====================
        // 16 warps: TMEM lane quadrant q = warp & 3 (rows 32q..32q+31 of the tile), column quarter part = warp >> 2
        // (units 16*part .. 16*part+15), processed as four 16-column pieces of 4 units each.
        TC_PROF_DECL
        const int q = warp & 3, part = warp >> 2;
        float* stg = reinterpret_cast<float*>(staging + (size_t)warp * kStagingBytes);
        uint32_t tcount = 0;
        // c_{t-1} of this warp's 16 units for a whole tile is fetched one TILE ahead, right after the last proxy fence of
        // the previous tile: the MEMBAR inside fence.proxy.async waits for every outstanding load of the thread, so a
        // load issued inside the piece loop would cost each piece a full DRAM latency
        float4 cpv[4];
        auto fetch_cprev = [&](int tile_n) {
            const int64_t rn = (int64_t)tile_n * kTileM + q * 32 + lane;
            const bool okn = p.c_prev != nullptr && tile_n < p.n_tiles && rn < p.rows;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                cpv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (okn) cpv[j] = *reinterpret_cast<const float4*>(p.c_prev + ws_off(p.blocked_cs, rn, part * 16 + 4 * j));
            }
        };
        fetch_cprev((int)blockIdx.x);
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++tcount) {
            const int a = tcount & 1;
            const uint32_t aph = (tcount >> 1) & 1;
            const int64_t r0 = (int64_t)tile * kTileM + q * 32;     // first row of this warp
            const int64_t r = r0 + lane;
            const bool valid = r < p.rows;
            mbar_wait(&bar->tmem_full[a], aph, 3);
            tc_fence_after();
            const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)a * kFwdN + (uint32_t)part * 64;
#pragma unroll
            for (int pc = 0; pc < 4; ++pc) {                        // piece: 16 gate columns = 4 units
                uint32_t v[16];
                tmem_ld16(t_row + pc * 16, v);
                const int unit0 = part * 16 + pc * 4;
                const float cp[4] = {cpv[pc].x, cpv[pc].y, cpv[pc].z, cpv[pc].w};
                tmem_ld_wait();
                float hn[4], cn[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int col = 4 * (unit0 + u);
                    const float4 bv = *reinterpret_cast<const float4*>(&tail->bias[col]);
                    const float pi = __uint_as_float(v[4 * u + 0]) + bv.x;
                    const float pf = __uint_as_float(v[4 * u + 1]) + bv.y;
                    const float pg = __uint_as_float(v[4 * u + 2]) + bv.z;
                    const float po = __uint_as_float(v[4 * u + 3]) + bv.w;
                    const float gi = sigmoidf_(pi), gf = sigmoidf_(pf), gg = tanhf_(pg), go = sigmoidf_(po);
                    cn[u] = fmaf(gf, cp[u], gi * gg);
                    hn[u] = go * tanhf_(cn[u]);
                    v[4 * u + 0] = __float_as_uint(gi);
                    v[4 * u + 1] = __float_as_uint(gf);
                    v[4 * u + 2] = __float_as_uint(gg);
                    v[4 * u + 3] = __float_as_uint(go);
                }
                if (p.gates_out != nullptr) {
                    // gate tape: this warp's [32 rows][16 columns] piece goes through its staging tile (16-byte chunks
                    // XOR-swizzled with (row >> 1) & 3 == the TMA 64-byte swizzle, conflict-free for the lanes)
                    if (GATES_TMA) {                   // previous piece's tensor store must have read the tile
                        if (lane == 0 && p.dbg_skip_hc != 2) bulk_wait_read0();
                    }
                    __syncwarp();
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        *reinterpret_cast<uint4*>(stg + lane * 16 + ((u ^ ((lane >> 1) & 3)) << 2)) =
                            make_uint4(v[4 * u], v[4 * u + 1], v[4 * u + 2], v[4 * u + 3]);
                    if (GATES_TMA) {
                        // one TMA tensor store per piece: no per-thread global stores, rows past the end are clipped
                        fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0) {
                            tma_store_2d(&p.gates_map, smem_u32(stg), 4 * unit0, (int)r0);
                            bulk_commit_group();
                        }
                    } else {
                        // a store instruction writes 8 row segments of 64 contiguous bytes
                        __syncwarp();
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int row = i * 8 + (lane >> 2), qq = lane & 3;
                            const uint4 val = *reinterpret_cast<const uint4*>(stg + row * 16 + ((qq ^ ((row >> 1) & 3)) << 2));
                            if (r0 + row < p.rows)
                                *reinterpret_cast<uint4*>(p.gates_out + (r0 + row) * kGateCols + 4 * unit0 + qq * 4) = val;
                        }
                    }
                }
                if (GATES_TMA && p.hc_tma) {
                    // h / c through the same staging tile once the tape store has read it: the epilogue threads then
                    // issue no global stores at all, so the proxy fences (MEMBAR) have nothing outstanding to wait for
                    if (lane == 0 && p.dbg_skip_hc != 2) bulk_wait_read0();
                    __syncwarp();
                    *reinterpret_cast<float4*>(stg + lane * 4) = make_float4(hn[0], hn[1], hn[2], hn[3]);
                    *reinterpret_cast<float4*>(stg + 128 + lane * 4) = make_float4(cn[0], cn[1], cn[2], cn[3]);
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_2d(&p.h_map, smem_u32(stg), unit0, (int)r0);
                        if (p.blocked_cs)
                            tma_store_2d(&p.c_map, smem_u32(stg + 128), unit0 & 7,
                                         (tile * 8 + (unit0 >> 3)) * kTileM + q * 32);
                        else
                            tma_store_2d(&p.c_map, smem_u32(stg + 128), unit0, (int)r0);
                        bulk_commit_group();
                    }
                } else if (valid && p.dbg_skip_hc != 1) {
                    // h / c after the tape hand-off: the proxy fence above does not have to wait for these stores
                    *reinterpret_cast<float4*>(p.h_out + r * kHid + unit0) = make_float4(hn[0], hn[1], hn[2], hn[3]);
                    *reinterpret_cast<float4*>(p.c_out + ws_off(p.blocked_cs, r, unit0)) = make_float4(cn[0], cn[1], cn[2], cn[3]);
                }
            }
            fetch_cprev(tile + (int)gridDim.x);
            tc_fence_before();
            mbar_arrive(&bar->tmem_empty[a]);
        }
        if (GATES_TMA && lane == 0) bulk_wait0();      // shared memory stays valid until the last tensor store is done
        TC_PROF_FLUSH(2, tid == 0)
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == kMmaWarp) tmem_dealloc(tmem_base, 512);
}

// =====================================================================================================
// forward cell, N-split variant with RESIDENT weight halves  --  opt-in (STMGCN_FWD_NSPLIT=1), kept for study
// Written at the end of round 1 from the profile of lstm_cell_tc_kernel.  It passes the
// golden and TC-vs-FFMA parity tests but is 18 % SLOWER than lstm_cell_tc_kernel (6.39 vs 5.4 ms per branch forward at
// cfg3): every A tile is split to tf32 hi/lo by two CTAs, and that costs more than the resident weights gain.
//
// The gate columns are unit-interleaved (col = 4*unit + gate), so columns [128h, 128h+128) are a self-contained half
// (units 32h .. 32h+31).  CTA c owns half h = c & 1 for the whole launch: the hi/lo weight images of its half for all
// k-blocks (4 x 32 KB) are loaded ONCE and stay in shared memory; only the 16 KB raw A tiles stream through two stages
// (TMA tensor load into the hi tile, split in place); four 128-column TMEM accumulators decouple the MMA warp from the
// epilogue.  CTAs 2i and 2i+1 walk the same tile sequence, so the second read of an A tile hits L2.
// =====================================================================================================
constexpr int kNsN = 128;                                         // gate columns per CTA
constexpr int kNsStages = 2;
constexpr int kNsAStage = 2 * kABytes;                            // 32 KB: A hi | A lo
constexpr int kNsBHalf = kNsN * kKB * 4;                          // 16 KB: [128 n][32 k] hi or lo of one k-block
constexpr int kNsBkb = 2 * kNsBHalf;                              // 32 KB per k-block: hi | lo
constexpr int kNsMaxKb = 4;
constexpr int kNsAccs = 4;
constexpr int kNsEpiWarps = 16;
constexpr int kNsSplitWarps = 8;
constexpr int kNsThreads = (kNsEpiWarps + kNsSplitWarps + 2) * 32;
struct NsTail {
    float bias[kGateCols];
    uint64_t a_full[kNsStages];
    uint64_t full[kNsStages];
    uint64_t empty[kNsStages];
    uint64_t b_full;
    uint64_t tmem_full[kNsAccs];
    uint64_t tmem_empty[kNsAccs];
    uint32_t tmem_base;
};
constexpr size_t kNsSmem = 1024 + (size_t)kNsMaxKb * kNsBkb + (size_t)kNsStages * kNsAStage +
                           (size_t)kNsEpiWarps * kStagingBytes + sizeof(NsTail);
static_assert(kNsSmem <= 232448, "N-split forward kernel exceeds the 227 KB shared-memory limit");

__global__ void __launch_bounds__(kNsThreads, 1) lstm_cell_nsplit_kernel(const __grid_constant__ CellParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the __shared__ address space (LDS/STS, not generic LD/ST)
    uint8_t* bres = smem;                                          // resident weight half: nkb x [hi 16 KB | lo 16 KB]
    uint8_t* stages = bres + (size_t)kNsMaxKb * kNsBkb;
    uint8_t* staging = stages + (size_t)kNsStages * kNsAStage;
    NsTail* tail = (NsTail*)(staging + (size_t)kNsEpiWarps * kStagingBytes);
    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;
    constexpr int kMmaWarp = kNsEpiWarps + kNsSplitWarps;
    constexpr int kSplitters = kNsSplitWarps * 32;
    const int half = blockIdx.x & 1;
    const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;     // host launches an even grid
    const int my_tiles = pair < p.n_tiles ? (p.n_tiles - pair + npairs - 1) / npairs : 0;

    pdl_launch_dependents();
    if (tid == 0) {
        for (int s = 0; s < kNsStages; ++s) {
            mbar_init(&tail->a_full[s], 1);
            mbar_init(&tail->full[s], kSplitters);
            mbar_init(&tail->empty[s], 1);
        }
        mbar_init(&tail->b_full, 1);
        for (int a = 0; a < kNsAccs; ++a) {
            mbar_init(&tail->tmem_full[a], 1);
            mbar_init(&tail->tmem_empty[a], kNsEpiWarps * 32);
        }
        fence_barrier_init();
    }
    if (warp == kMmaWarp) tmem_alloc(&tail->tmem_base, 512);
    for (int i = tid; i < kGateCols; i += kNsThreads) tail->bias[i] = p.bias ? p.bias[i] : 0.f;
    pdl_wait();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tail->tmem_base;

    if (warp >= kNsEpiWarps && warp < kMmaWarp) {
        // ===================== splitters: raw A tile (in the stage's hi tile) -> tf32 hi in place + lo =====================
        const int gtid = tid - kNsEpiWarps * 32;
        constexpr int kPer = 1024 / kSplitters;                    // 4 float4 per thread per k-block
        const int c = gtid & 7, rsub = gtid >> 3;                  // rows rsub + 32*i, 16-byte chunk c of the 128-byte row
        const int total = my_tiles * p.nkb;
        for (int j = 0; j < total; ++j) {
            const int tile = pair + (j / p.nkb) * npairs, kb = j % p.nkb;
            const float* seg = (kb >> 1) ? p.seg1 : p.seg0;
            const bool aux_blk = p.aux && kb == 2;
            const bool from_tma = !aux_blk && seg != nullptr;
            const int s = j % kNsStages;
            const uint32_t ph = (j / kNsStages) & 1;
            uint8_t* st = stages + (size_t)s * kNsAStage;
            mbar_wait_raw(&tail->a_full[s], ph);
            float4 buf[kPer];
#pragma unroll
            for (int i = 0; i < kPer; ++i) {
                const int row = rsub + (kSplitters / 8) * i;
                const int64_t r = (int64_t)tile * kTileM + row;
                if (from_tma) {
                    buf[i] = *reinterpret_cast<const float4*>(st + (uint32_t)row * 128u + (uint32_t)((c ^ (row & 7)) << 4));
                } else if (aux_blk) {                              // [x*s (C cols) | 1 | 0 ...]
                    float v[4] = {0.f, 0.f, 0.f, 0.f};
                    if (c < 2 && r < p.rows) {
                        const float sv = p.sg[(r % p.b_inner) * p.t_len + p.t];
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) {
                            const int col = 4 * c + jj;
                            if (col < p.c_in) v[jj] = p.xo[(r * p.t_len + p.t) * p.c_in + col] * sv;
                            else if (col == p.c_in) v[jj] = 1.0f;
                        }
                    }
                    buf[i] = make_float4(v[0], v[1], v[2], v[3]);
                } else {
                    buf[i] = make_float4(0.f, 0.f, 0.f, 0.f);     // absent segment (h_{-1} = 0)
                }
            }
#pragma unroll
            for (int i = 0; i < kPer; ++i) {
                const int row = rsub + (kSplitters / 8) * i;
                split_store(st, (uint32_t)row * 128u + (uint32_t)((c ^ (row & 7)) << 4), buf[i]);   // hi at st, lo at st + kABytes
            }
            fence_proxy_async_smem();
            mbar_arrive(&tail->full[s]);
        }
    } else if (warp == kMmaWarp) {
        // ===================== MMA issuer: A from the stage, B from the resident half =====================
        constexpr uint32_t idesc = idesc_tf32(kTileM, kNsN);
        mbar_wait_raw(&tail->b_full, 0);
        tc_fence_after();
        uint32_t it = 0;
        for (int item = 0; item < my_tiles; ++item) {
            const int a = item % kNsAccs;
            const uint32_t aph = (uint32_t)(item / kNsAccs) & 1;
            mbar_wait_raw(&tail->tmem_empty[a], aph ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)a * kNsN;
            for (int kb = 0; kb < p.nkb; ++kb, ++it) {
                const int s = it % kNsStages;
                const uint32_t ph = (it / kNsStages) & 1;
                mbar_wait_raw(&tail->full[s], ph);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t st = smem_u32(stages + (size_t)s * kNsAStage);
                    const uint32_t bb = smem_u32(bres + (size_t)kb * kNsBkb);
                    const uint64_t a_hi = smem_desc_k_sw128(st);
                    const uint64_t a_lo = smem_desc_k_sw128(st + kABytes);
                    const uint64_t b_hi = smem_desc_k_sw128(bb);
                    const uint64_t b_lo = smem_desc_k_sw128(bb + kNsBHalf);
#pragma unroll
                    for (int pass = 0; pass < 3; ++pass) {
                        const uint64_t da = (pass == 1) ? a_lo : a_hi;
                        const uint64_t db = (pass == 2) ? b_lo : b_hi;
#pragma unroll
                        for (int k = 0; k < kKB / 8; ++k)
                            mma_tf32(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc,
                                     (kb > 0 || pass > 0 || k > 0) ? 1u : 0u);
                    }
                    mma_commit(&tail->empty[s]);
                }
                __syncwarp();
            }
            if (lane == 0) mma_commit(&tail->tmem_full[a]);
            __syncwarp();
        }
    } else if (warp == kMmaWarp + 1) {
        // ===================== producer: resident weight half once, then one raw A tile per freed stage =====================
        if (lane == 0) {
            mbar_arrive_expect_tx(&tail->b_full, (uint32_t)p.nkb * kNsBkb);
            for (int kb = 0; kb < p.nkb; ++kb) {
                const float* img = p.wimg + (size_t)kb * (2 * kFwdBBytes / 4);              // [hi 256 x 32 | lo 256 x 32]
                bulk_g2s(bres + (size_t)kb * kNsBkb, img + (size_t)half * (kNsBHalf / 4), kNsBHalf, &tail->b_full);
                bulk_g2s(bres + (size_t)kb * kNsBkb + kNsBHalf, img + kFwdBBytes / 4 + (size_t)half * (kNsBHalf / 4), kNsBHalf,
                         &tail->b_full);
            }
            const int total = my_tiles * p.nkb;
            for (int j = 0; j < total; ++j) {
                const int tile = pair + (j / p.nkb) * npairs, kb = j % p.nkb;
                const int s = j % kNsStages;
                const uint32_t ph = (j / kNsStages) & 1;
                mbar_wait_raw(&tail->empty[s], ph ^ 1);
                const float* seg = (kb >> 1) ? p.seg1 : p.seg0;
                if (!(p.aux && kb == 2) && seg != nullptr) {
                    mbar_arrive_expect_tx(&tail->a_full[s], kABytes);
                    tma_load_2d(stages + (size_t)s * kNsAStage, (kb >> 1) ? &p.seg1_map : &p.seg0_map, (kb & 1) * kKB,
                                tile * kTileM, &tail->a_full[s]);
                } else {
                    mbar_arrive(&tail->a_full[s]);
                }
            }
        }
    } else {
        // ===================== epilogue: LSTM cell on this CTA's 32 units =====================
        // 16 warps: TMEM lane quadrant q = warp & 3, column quarter part = warp >> 2 (32 of the 128 columns = 8 units),
        // processed as two 16-column pieces of 4 units each
        const int q = warp & 3, part = warp >> 2;
        float* stg = reinterpret_cast<float*>(staging + (size_t)warp * kStagingBytes);
        const int ubase = 32 * half + 8 * part;                    // first unit of this warp
        float4 cpv[2];
        auto fetch_cprev = [&](int tile_n) {
            const int64_t rn = (int64_t)tile_n * kTileM + q * 32 + lane;
            const bool okn = p.c_prev != nullptr && tile_n < p.n_tiles && rn < p.rows;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                cpv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (okn) cpv[j] = *reinterpret_cast<const float4*>(p.c_prev + ws_off(p.blocked_cs, rn, ubase + 4 * j));
            }
        };
        fetch_cprev(pair);
        for (int item = 0; item < my_tiles; ++item) {
            const int tile = pair + item * npairs;
            const int a = item % kNsAccs;
            const uint32_t aph = (uint32_t)(item / kNsAccs) & 1;
            const int64_t r0 = (int64_t)tile * kTileM + q * 32;
            mbar_wait_raw(&tail->tmem_full[a], aph);
            tc_fence_after();
            const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)a * kNsN + (uint32_t)part * 32;
#pragma unroll
            for (int pc = 0; pc < 2; ++pc) {
                uint32_t v[16];
                tmem_ld16(t_row + pc * 16, v);
                const int unit0 = ubase + pc * 4;
                const float cp[4] = {cpv[pc].x, cpv[pc].y, cpv[pc].z, cpv[pc].w};
                tmem_ld_wait();
                float hn[4], cn[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float4 bv = *reinterpret_cast<const float4*>(&tail->bias[4 * (unit0 + u)]);
                    const float gi = sigmoidf_(__uint_as_float(v[4 * u + 0]) + bv.x);
                    const float gf = sigmoidf_(__uint_as_float(v[4 * u + 1]) + bv.y);
                    const float gg = tanhf_(__uint_as_float(v[4 * u + 2]) + bv.z);
                    const float go = sigmoidf_(__uint_as_float(v[4 * u + 3]) + bv.w);
                    cn[u] = fmaf(gf, cp[u], gi * gg);
                    hn[u] = go * tanhf_(cn[u]);
                    v[4 * u + 0] = __float_as_uint(gi);
                    v[4 * u + 1] = __float_as_uint(gf);
                    v[4 * u + 2] = __float_as_uint(gg);
                    v[4 * u + 3] = __float_as_uint(go);
                }
                if (p.gates_out != nullptr) {
                    if (lane == 0) bulk_wait_read0();
                    __syncwarp();
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        *reinterpret_cast<uint4*>(stg + lane * 16 + ((u ^ ((lane >> 1) & 3)) << 2)) =
                            make_uint4(v[4 * u], v[4 * u + 1], v[4 * u + 2], v[4 * u + 3]);
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_2d(&p.gates_map, smem_u32(stg), 4 * unit0, (int)r0);
                        bulk_commit_group();
                    }
                }
                if (lane == 0) bulk_wait_read0();
                __syncwarp();
                *reinterpret_cast<float4*>(stg + lane * 4) = make_float4(hn[0], hn[1], hn[2], hn[3]);
                *reinterpret_cast<float4*>(stg + 128 + lane * 4) = make_float4(cn[0], cn[1], cn[2], cn[3]);
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    tma_store_2d(&p.h_map, smem_u32(stg), unit0, (int)r0);
                    if (p.blocked_cs)
                        tma_store_2d(&p.c_map, smem_u32(stg + 128), unit0 & 7, (tile * 8 + (unit0 >> 3)) * kTileM + q * 32);
                    else
                        tma_store_2d(&p.c_map, smem_u32(stg + 128), unit0, (int)r0);
                    bulk_commit_group();
                }
            }
            fetch_cprev(tile + npairs);
            tc_fence_before();
            mbar_arrive(&tail->tmem_empty[a]);
        }
        if (lane == 0) bulk_wait0();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == kMmaWarp) tmem_dealloc(tmem_base, 512);
}

// =====================================================================================================
// backward: pointwise dA in the loader + data GEMM  [dx_below | dh_prev] = dA . Wp^T
// =====================================================================================================
constexpr int kBwdEpiWarps = 4;
constexpr int kBwdNkb = kGateCols / kKB;                          // 8
// TMA-fed variant: a producer thread streams every (tile, k-block) item -- the gate slice through a tensor map, the
// five tile-blocked (128 x 8) vectors as 4 KB bulk copies -- into one raw slot per loader group, one item ahead of the
// group's arithmetic.  Bulk copies are tracked by mbarriers, not by the thread's scoreboard, so (unlike register
// prefetches) they are not waited for by the MEMBAR that fence.proxy.async implies.
constexpr int kRawGates = kTileM * kKB * 4;                       // 16 KB  [128 rows][32 gate columns]
constexpr int kRawVec = kTileM * 8 * 4;                           // 4 KB   [128 rows][8 units]
constexpr int kRawX = kTileM * 12 * 4;                            // 6 KB   layer 0: the tile's rows of xo (T <= 12, C = 1)
constexpr int kRawSlot = kRawGates + 5 * kRawVec;                 // 36 KB: gates | dh_rec | dh_in | c_t | c_prev | dc (| x)
constexpr int kRawSlots = 2;                                      // one per loader group
constexpr int kBwdSgMax = 1024;                                   // batch entries of the gate / its adjoint kept in smem
constexpr int kBwdMaxC = 1;                                       // layer-0 TC backward handles input_dim 1

template <int N>
struct BwdTailT {
    float s_db[kLoaderWarps][kGateCols];                        // per-loader-warp private partial sums
    float s_dwx[N == 64 ? kLoaderWarps : 1][kGateCols];         // layer 0 only (input_dim == 1)
    float s_ds[N == 64 ? kBwdSgMax : 1];                        // layer 0 only: d s[b, t] partial sums
    float s_sg[N == 64 ? kBwdSgMax : 1];                        // layer 0 only: s[b, t]
    float s_wx[N == 64 ? kGateCols : 1];                        // layer 0 only: W_ih row (input_dim == 1)
    Barriers bar;
    uint64_t raw_full[kRawSlots];
    uint64_t raw_empty[kRawSlots];
};
template <int N, bool TMA>
struct BwdCfg {
    using BwdTail = BwdTailT<N>;
    static constexpr int kStages = TMA ? 2 : 3;
    static constexpr int kWarps = kBwdEpiWarps + kLoaderWarps + 1 + (TMA ? 1 : 0);     // + the producer warp
    static constexpr int kThreads = kWarps * 32;
    static constexpr int kBBytes = N * kKB * 4;
    static constexpr int kStageBytes = 2 * kABytes + 2 * kBBytes;
    static constexpr int kSlotBytes = kRawSlot + (N == 64 ? kRawX : 0);
    static constexpr size_t kRawBytes = TMA ? (size_t)kRawSlots * kSlotBytes : 0;
    static constexpr size_t kSmem = 1024 + (size_t)kStages * kStageBytes + kRawBytes + sizeof(BwdTail);
    static constexpr int kTmemCols = 2 * N;      // 256 or 128 (power of two >= 32)
    static_assert(kSmem <= 232448, "backward kernel exceeds the 227 KB shared-memory limit");
};

struct BwdParams {
    float* gates;            // (rows,256) in: post-activation i,f,g,o (interleaved); out: dA
    const float* c_t;        // (rows,64)
    const float* c_prev;     // (rows,64) or nullptr
    const float* dh_in;      // (rows,64) or nullptr   gradient from the layer above (or d_top)
    float* dh_rec;           // (rows,64) in: recurrent gradient for this step; out: for step t-1
    float* dc;               // (rows,64) in/out
    float* dx_out;           // (rows,64) or nullptr (layer 0): gradient for the layer below
    const float* wimg_t;     // 8 k-blocks x [hi | lo] images of Wp^T, [N][32] each
    float* dbp;              // (256) +=
    const float* wx;         // layer 0: (C,256), else nullptr
    float* dwx;              // (C,256) +=
    const float* xo;
    const float* sg;
    float* d_s;              // (B,T) +=
    int c_in, t, t_len;
    int64_t b_inner;
    int64_t rows;
    int n_tiles;
    int prefetch;
    int blocked;             // dh_in / dh_rec / dc / dx_out use the tile-blocked layout [tile][unit/8][128 rows][8 units]
    int first;               // t == T-1: the incoming dh_rec / dc are zero by definition and are not read
    alignas(64) CUtensorMap gates_map;   // TMA-fed variant: (rows, 256) slice, box 128 rows x 32 columns, no swizzle
};

template <int N, bool TMA>
__global__ void __launch_bounds__((BwdCfg<N, TMA>::kThreads), 1) lstm_bwd_tc_kernel(const __grid_constant__ BwdParams p) {
    using Cfg = BwdCfg<N, TMA>;
    constexpr int kBwdStages = Cfg::kStages;
    constexpr int kBwdThreads = Cfg::kThreads;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the __shared__ address space (LDS/STS, not generic LD/ST)
    using BwdTail = typename Cfg::BwdTail;
    uint8_t* raw = smem + (size_t)kBwdStages * Cfg::kStageBytes;
    BwdTail* tail = (BwdTail*)(raw + Cfg::kRawBytes);
    Barriers* bar = &tail->bar;
    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;
    constexpr int kMmaWarp = kBwdEpiWarps + kLoaderWarps;
    constexpr bool l0 = (N == 64);               // layer 0 <=> kd = 64 (no layer below; carries the gate adjoint)
    const bool ds_smem = l0 && p.b_inner <= kBwdSgMax;
    // layer 0, TMA-fed: the tile's xo rows arrive by bulk copy and s[., t] / W_ih sit in shared memory, so the loader
    // warps issue no global loads at all (any outstanding load would be waited for at the next proxy fence)
    const bool x_tma = TMA && l0 && ds_smem && p.c_in == 1 && p.t_len * 4 * kTileM <= kRawX && (p.t_len % 4) == 0;

    pdl_launch_dependents();
    if (tid == 0) {
        if (TMA)
            for (int s = 0; s < kRawSlots; ++s) {
                mbar_init(&tail->raw_full[s], 1);                   // the producer's arrive.expect_tx
                mbar_init(&tail->raw_empty[s], kNumLoaders / 2);    // every thread of the slot's loader group
            }
        init_barriers(bar, kBwdStages, kBwdEpiWarps * 32, kNumLoaders / 2);     // one loader group per k-block
    }
    if (warp == kMmaWarp) tmem_alloc(&bar->tmem_base, Cfg::kTmemCols);
    for (int i = tid; i < kLoaderWarps * kGateCols; i += kBwdThreads) {
        (&tail->s_db[0][0])[i] = 0.f;
        if (l0) (&tail->s_dwx[0][0])[i] = 0.f;
    }
    if (ds_smem)
        for (int i = tid; i < (int)p.b_inner; i += kBwdThreads) tail->s_ds[i] = 0.f;
    if (x_tma) {
        for (int i = tid; i < (int)p.b_inner; i += kBwdThreads) tail->s_sg[i] = p.sg[(int64_t)i * p.t_len + p.t];
        for (int i = tid; i < kGateCols; i += kBwdThreads) tail->s_wx[i] = p.wx[i];
    }
    pdl_wait();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bar->tmem_base;

    if (warp >= kBwdEpiWarps && warp < kMmaWarp) {
        // ===================== loaders: BPTT pointwise -> dA -> operand tiles =====================
        // two independent loader groups alternate k-blocks (see the note in lstm_cell_tc_kernel);
        // thread <-> (unit c of the k-block, rows rsub + kRowStep*i): one LSTM cell per (i, k-block)
        TC_PROF_DECL
        constexpr int kGroups = 2, kGT = kNumLoaders / kGroups;
        constexpr int kCells = kTileM * 8 / kGT;              // LSTM cells per thread per k-block
        constexpr int kRowStep = kGT / 8;
        static_assert(kBwdNkb % kGroups == 0, "k-blocks of a tile must split evenly over the loader groups");
        const int ltid = tid - kBwdEpiWarps * 32;
        const int grp = ltid / kGT, gtid = ltid % kGT;
        const int c = gtid & 7, rsub = gtid >> 3, lwarp = ltid >> 5;
        struct CellIn { float4 g; float dh, dh2, ct, cp, dc; };
        float xs[kCells], dxs[kCells], xraw[kCells];
        const int my_tiles = (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
        const int total = my_tiles * kBwdNkb;
        for (int j = grp; j < total; j += kGroups) {
            const int tile = blockIdx.x + (j / kBwdNkb) * gridDim.x, kb = j % kBwdNkb;
            CellIn buf[kCells];
            float4 wv = make_float4(0.f, 0.f, 0.f, 0.f);
            {
                const int unit = kb * 8 + c;
                if (l0) {
                    wv = x_tma ? *reinterpret_cast<const float4*>(&tail->s_wx[4 * unit])
                               : __ldg(reinterpret_cast<const float4*>(p.wx + 4 * unit));
                    if (kb == grp && !x_tma) {     // this group's first k-block of the tile: modulated inputs of its rows
#pragma unroll
                        for (int i = 0; i < kCells; ++i) {
                            const int64_t r = (int64_t)tile * kTileM + rsub + kRowStep * i;
                            xs[i] = (r < p.rows) ? p.xo[(r * p.t_len + p.t) * p.c_in] * p.sg[(r % p.b_inner) * p.t_len + p.t] : 0.f;
                            dxs[i] = 0.f;
                        }
                    }
                }
                if (TMA) {
                    // this group's raw slot was filled by the producer warp while the group worked on its previous item
                    const uint8_t* rs = raw + (size_t)grp * Cfg::kSlotBytes;
                    mbar_wait(&tail->raw_full[grp], (uint32_t)((j / kGroups) & 1), 1);
                    if (x_tma && (kb == grp || kb == kBwdNkb - kGroups + grp)) {
                        // first item of the tile: modulated inputs x*s; last item: raw x for the gate adjoint
                        const float* rx = reinterpret_cast<const float*>(rs + kRawSlot);
                        const uint32_t b0 = (uint32_t)(((int64_t)tile * kTileM) % p.b_inner);
#pragma unroll
                        for (int i = 0; i < kCells; ++i) {
                            const int row = rsub + kRowStep * i;
                            const bool ok = (int64_t)tile * kTileM + row < p.rows;
                            const float xv = ok ? rx[row * p.t_len + p.t] : 0.f;
                            if (kb == grp) {
                                xs[i] = xv * tail->s_sg[(b0 + (uint32_t)row) % (uint32_t)p.b_inner];
                                dxs[i] = 0.f;
                            } else {
                                xraw[i] = xv;
                            }
                        }
                    }
#pragma unroll
                    for (int i = 0; i < kCells; ++i) {
                        const int row = rsub + kRowStep * i;
                        const bool ok = (int64_t)tile * kTileM + row < p.rows;       // padding rows hold garbage
                        const float* vec = reinterpret_cast<const float*>(rs + kRawGates) + row * 8 + c;
                        const float4 g = *reinterpret_cast<const float4*>(rs + row * 128 + c * 16);
                        const float dh = p.first ? 0.f : vec[0];
                        const float dh2 = p.dh_in ? vec[kRawVec / 4] : 0.f;
                        const float ct = vec[2 * (kRawVec / 4)];
                        const float cp = p.c_prev ? vec[3 * (kRawVec / 4)] : 0.f;
                        const float dcv = p.first ? 0.f : vec[4 * (kRawVec / 4)];
                        buf[i].g = ok ? g : make_float4(0.f, 0.f, 0.f, 0.f);
                        buf[i].dh = ok ? dh : 0.f;
                        buf[i].dh2 = ok ? dh2 : 0.f;
                        buf[i].ct = ok ? ct : 0.f;
                        buf[i].cp = ok ? cp : 0.f;
                        buf[i].dc = ok ? dcv : 0.f;
                    }
                    mbar_arrive(&tail->raw_empty[grp]);        // slot may be refilled (item j + 2)
                } else {
#pragma unroll
                for (int i = 0; i < kCells; ++i) {
                    const int64_t r = (int64_t)tile * kTileM + rsub + kRowStep * i;
                    buf[i].g = make_float4(0.f, 0.f, 0.f, 0.f);
                    buf[i].dh = buf[i].dh2 = buf[i].ct = buf[i].cp = buf[i].dc = 0.f;
                    if (r < p.rows) {
                        buf[i].g = *reinterpret_cast<const float4*>(p.gates + r * kGateCols + 4 * unit);
                        const int64_t eb = ws_off(p.blocked, r, unit);
                        if (!p.first) buf[i].dh = p.dh_rec[eb];
                        if (p.dh_in) buf[i].dh2 = p.dh_in[eb];
                        buf[i].ct = p.c_t[eb];                      // the cell-state tape shares the workspace layout
                        buf[i].cp = p.c_prev ? p.c_prev[eb] : 0.f;
                        if (!p.first) buf[i].dc = p.dc[eb];
                    }
                }
                }
            }
            const int64_t row_base = (int64_t)tile * kTileM;
            const int unit = kb * 8 + c;
            float4 da[kCells];
            float dcn[kCells];
            float4 sb = make_float4(0.f, 0.f, 0.f, 0.f), sx = sb;
#pragma unroll
            for (int i = 0; i < kCells; ++i) {
                const float4 g = buf[i].g;
                const float dh = buf[i].dh + buf[i].dh2;
                const float tc_ = tanhf_(buf[i].ct);
                const float dcv = buf[i].dc + dh * g.w * (1.f - tc_ * tc_);
                da[i].x = dcv * g.z * g.x * (1.f - g.x);
                da[i].y = dcv * buf[i].cp * g.y * (1.f - g.y);
                da[i].z = dcv * g.x * (1.f - g.z * g.z);
                da[i].w = dh * tc_ * g.w * (1.f - g.w);
                dcn[i] = dcv * g.y;
                sb.x += da[i].x; sb.y += da[i].y; sb.z += da[i].z; sb.w += da[i].w;
                if (l0) {
                    sx.x = fmaf(xs[i], da[i].x, sx.x); sx.y = fmaf(xs[i], da[i].y, sx.y);
                    sx.z = fmaf(xs[i], da[i].z, sx.z); sx.w = fmaf(xs[i], da[i].w, sx.w);
                    dxs[i] += da[i].x * wv.x + da[i].y * wv.y + da[i].z * wv.z + da[i].w * wv.w;
                }
            }
            // bias (and layer-0 input-weight) gradient partials: reduce over the 4 lanes that share this unit,
            // then one shared-memory atomic per value
#pragma unroll
            for (int o = 8; o <= 16; o <<= 1) {
                sb.x += __shfl_xor_sync(0xffffffffu, sb.x, o); sb.y += __shfl_xor_sync(0xffffffffu, sb.y, o);
                sb.z += __shfl_xor_sync(0xffffffffu, sb.z, o); sb.w += __shfl_xor_sync(0xffffffffu, sb.w, o);
                if (l0) {
                    sx.x += __shfl_xor_sync(0xffffffffu, sx.x, o); sx.y += __shfl_xor_sync(0xffffffffu, sx.y, o);
                    sx.z += __shfl_xor_sync(0xffffffffu, sx.z, o); sx.w += __shfl_xor_sync(0xffffffffu, sx.w, o);
                }
            }
            if (lane < 8) {        // lanes 0-7 own units c = 0..7 of this warp's private accumulator row
                float4* acc = reinterpret_cast<float4*>(&tail->s_db[lwarp][4 * unit]);
                float4 t = *acc;
                t.x += sb.x; t.y += sb.y; t.z += sb.z; t.w += sb.w;
                *acc = t;
                if (l0) {
                    float4* accx = reinterpret_cast<float4*>(&tail->s_dwx[l0 ? lwarp : 0][4 * unit]);
                    float4 tx = *accx;
                    tx.x += sx.x; tx.y += sx.y; tx.z += sx.z; tx.w += sx.w;
                    *accx = tx;
                }
            }
            const int s = j % kBwdStages;
            const uint32_t ph = (j / kBwdStages) & 1;
            mbar_wait(&bar->empty[s], ph ^ 1, 0);
            uint8_t* st = smem + (size_t)s * Cfg::kStageBytes;
            if (gtid == 0) {
                mbar_arrive_expect_tx(&bar->full[s], 2 * Cfg::kBBytes);
                const float* src = p.wimg_t + (size_t)kb * (2 * Cfg::kBBytes / 4);
                bulk_g2s(st + 2 * kABytes, src, Cfg::kBBytes, &bar->full[s]);
                bulk_g2s(st + 2 * kABytes + Cfg::kBBytes, src + Cfg::kBBytes / 4, Cfg::kBBytes, &bar->full[s]);
            }
#pragma unroll
            for (int i = 0; i < kCells; ++i) {
                const int row = rsub + kRowStep * i;
                split_store(st, (uint32_t)row * 128u + (uint32_t)((c ^ (row & 7)) << 4), da[i]);
            }
            fence_proxy_async_smem();
            mbar_arrive(&bar->full[s]);
            // tape updates AFTER the hand-off: the proxy fence (MEMBAR.ALL.CTA) of the NEXT k-block is the first point
            // that waits for these stores, a whole iteration later
#pragma unroll
            for (int i = 0; i < kCells; ++i) {
                const int64_t r = row_base + rsub + kRowStep * i;
                if (r < p.rows) {
                    p.dc[ws_off(p.blocked, r, unit)] = dcn[i];
                    *reinterpret_cast<float4*>(p.gates + r * kGateCols + 4 * unit) = da[i];
                }
            }
            if (p.prefetch && gtid == 0 && kb == grp && grp == 0) {   // L2 prefetch of this CTA's next tile
                const int nt = tile + gridDim.x;
                if (nt < p.n_tiles) {
                    const int64_t r0 = (int64_t)nt * kTileM;
                    const int64_t nr = (p.rows - r0) < kTileM ? (p.rows - r0) : kTileM;
                    const uint32_t bh = (uint32_t)(nr * kHid * 4);
                    prefetch_l2(p.gates + r0 * kGateCols, (uint32_t)(nr * kGateCols * 4));
                    prefetch_l2(p.c_t + r0 * kHid, bh);
                    prefetch_l2(p.dh_rec + r0 * kHid, bh);
                    prefetch_l2(p.dc + r0 * kHid, bh);
                    if (p.c_prev) prefetch_l2(p.c_prev + r0 * kHid, bh);
                    if (p.dh_in) prefetch_l2(p.dh_in + r0 * kHid, bh);
                }
            }
            if (l0 && kb == kBwdNkb - kGroups + grp) {     // this group's last k-block of the tile: gate adjoint partial
                // d s[b,t] += dxmod[r] * xo[r,t]   (STMGCN.py:44, C = 1); b = r mod B with one 64-bit modulo per tile
                const uint32_t bt0 = (uint32_t)(row_base % p.b_inner);
#pragma unroll
                for (int i = 0; i < kCells; ++i) {
                    const int64_t r = row_base + rsub + kRowStep * i;
                    float d = dxs[i];
                    d += __shfl_xor_sync(0xffffffffu, d, 1);
                    d += __shfl_xor_sync(0xffffffffu, d, 2);
                    d += __shfl_xor_sync(0xffffffffu, d, 4);
                    if (c == 0 && r < p.rows) {
                        const float contrib = d * (x_tma ? xraw[i] : p.xo[(r * p.t_len + p.t) * p.c_in]);
                        const int64_t b = p.b_inner <= 0x7fffffff
                                              ? (int64_t)((bt0 + (uint32_t)(rsub + kRowStep * i)) % (uint32_t)p.b_inner)
                                              : r % p.b_inner;
                        if (ds_smem) atomicAdd(&tail->s_ds[b], contrib);
                        else atomicAdd(&p.d_s[b * p.t_len + p.t], contrib);
                    }
                }
            }
        }
        TC_PROF_FLUSH((N == 128 ? 3 : 6), ltid == 0)
    } else if (warp == kMmaWarp) {
        mma_issuer<N, kBwdStages, (N == 128 ? 1 : 2)>(bar, smem, Cfg::kStageBytes, Cfg::kBBytes, kBwdNkb, p.n_tiles, tmem_base, lane);
    } else if (TMA && warp == kMmaWarp + 1) {
        // ===================== producer: TMA loads of the raw inputs, one item ahead per loader group =====================
        if (lane == 0) {
            const int my_tiles = (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
            const int total = my_tiles * kBwdNkb;
            const uint32_t bytes = kRawGates + kRawVec * (1 + (p.c_prev ? 1 : 0) + (p.dh_in ? 1 : 0) + (p.first ? 0 : 2));
            for (int j = 0; j < total; ++j) {
                const int slot = j & 1, n = j >> 1;
                const int tile = blockIdx.x + (j / kBwdNkb) * gridDim.x, kb = j % kBwdNkb;
                if (n > 0) mbar_wait_raw(&tail->raw_empty[slot], (uint32_t)((n - 1) & 1));
                uint8_t* rs = raw + (size_t)slot * Cfg::kSlotBytes;
                uint64_t* fb = &tail->raw_full[slot];
                const bool want_x = x_tma && (kb < 2 || kb >= kBwdNkb - 2);
                const int64_t r0 = (int64_t)tile * kTileM;
                const uint32_t xbytes = (uint32_t)(((p.rows - r0) < kTileM ? (p.rows - r0) : kTileM) * p.t_len * 4);
                mbar_arrive_expect_tx(fb, bytes + (want_x ? xbytes : 0u));
                if (want_x) bulk_g2s(rs + kRawSlot, p.xo + r0 * p.t_len, xbytes, fb);
                tma_load_2d(rs, &p.gates_map, kb * kKB, tile * kTileM, fb);
                const int64_t off = ((int64_t)tile * 8 + kb) * (kRawVec / 4);      // tile-blocked (128 x 8) slice
                uint8_t* v = rs + kRawGates;
                if (!p.first) bulk_g2s(v, p.dh_rec + off, kRawVec, fb);
                if (p.dh_in) bulk_g2s(v + kRawVec, p.dh_in + off, kRawVec, fb);
                bulk_g2s(v + 2 * kRawVec, p.c_t + off, kRawVec, fb);
                if (p.c_prev) bulk_g2s(v + 3 * kRawVec, p.c_prev + off, kRawVec, fb);
                if (!p.first) bulk_g2s(v + 4 * kRawVec, p.dc + off, kRawVec, fb);
            }
        }
    } else {
        // ===================== epilogue: store [dx_below | dh_prev] =====================
        TC_PROF_DECL
        uint32_t tcount = 0;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++tcount) {
            const int a = tcount & 1;
            const uint32_t aph = (tcount >> 1) & 1;
            const int64_t r = (int64_t)tile * kTileM + warp * 32 + lane;
            const bool valid = r < p.rows;
            mbar_wait(&bar->tmem_full[a], aph, 3);
            tc_fence_after();
            const uint32_t t_row = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)a * N;
#pragma unroll 1
            for (int chunk = 0; chunk < N / 32; ++chunk) {
                uint32_t v[32];
                tmem_ld32(t_row + chunk * 32, v);
                tmem_ld_wait();
                if (valid) {
                    // columns [0,64) of a 128-wide result are dx_below, the last 64 are dh_prev
                    const int col = chunk * 32;
                    float* base = (N == 128 && col < 64) ? p.dx_out : p.dh_rec;
                    const int unit0 = col - ((N == 128 && col >= 64) ? 64 : 0);
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        *reinterpret_cast<uint4*>(base + ws_off(p.blocked, r, unit0 + 4 * j)) =
                            make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                }
            }
            tc_fence_before();
            mbar_arrive(&bar->tmem_empty[a]);
        }
        TC_PROF_FLUSH((N == 128 ? 5 : 8), tid == 0)
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == kMmaWarp) tmem_dealloc(tmem_base, Cfg::kTmemCols);
    for (int i = tid; i < kGateCols; i += kBwdThreads) {
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < kLoaderWarps; ++w) v += tail->s_db[w][i];
        atomicAdd(&p.dbp[i], v);
    }
    if (l0) {
        for (int i = tid; i < kGateCols; i += kBwdThreads) {
            float v = 0.f;
#pragma unroll
            for (int w = 0; w < kLoaderWarps; ++w) v += tail->s_dwx[l0 ? w : 0][i];
            atomicAdd(&p.dwx[i], v);
        }
        if (ds_smem)
            for (int i = tid; i < (int)p.b_inner; i += kBwdThreads) atomicAdd(&p.d_s[(int64_t)i * p.t_len + p.t], tail->s_ds[i]);
    }
}

// =====================================================================================================
// weight gradients:  dWp[kd x 256] += sum over (t, r) of [h_below_t | h_{t-1}][r, :]^T . dA_t[r, :]
// M = kd index (padded to 128), N = 256 gate columns, K = rows.  Both operands are row-major in HBM, i.e. K is the slow
// dimension: they are MN-major operands.  The loaders copy rows with coalesced float4 loads and store them as MN-major
// atoms with the 32-byte-base 128B swizzle (layout type SWIZZLE_128B_BASE32B = 1, see mn32_offset in tc_common.cuh;
// with the plain SWIZZLE_128B type and the MN-major descriptor bits the tf32 MMA returns zeros).  One TMEM
// accumulator lives for the whole kernel and is flushed with red.add.
// =====================================================================================================
constexpr int kWgStages = 2;
constexpr int kWgRows = 32;                                        // K per stage
constexpr int kWgABytes = 128 * kWgRows * 4;                       // 16 KB  [128 m][32 k] K-major
template <int N> struct WgCfg {
    static constexpr int kBBytes = N * kWgRows * 4;                // [N][32 k] K-major
    static constexpr int kStageBytes = 2 * kWgABytes + 2 * kBBytes;
    static constexpr size_t kSmem = 1024 + (size_t)kWgStages * kStageBytes + 64;
    static constexpr int kTmemCols = N < 32 ? 32 : N;
};
constexpr int kWgLoaderWarps = 16;
constexpr int kWgThreads = (kWgLoaderWarps + 1) * 32;              // 544

struct WgTail {
    uint64_t full[kWgStages];
    uint64_t empty[kWgStages];
    uint64_t done;
    uint32_t tmem_base;
};
static_assert(sizeof(WgTail) <= 64, "WgTail");

struct WgParams {
    const float* seg0;       // h_below tape base for this layer: (T, rows, 64) or nullptr (layer 0)
    const float* seg1;       // this layer's h tape base (T, rows, 64): read shifted by one step
    const float* h0;         // (rows, 64) value of h_{-1} or nullptr (zeros)
    const float* da;         // (T, rows, N)
    float* dwp;              // (kd, N) +=
    int shift1;              // 1: seg1 is read one step back (LSTM h_{t-1}); 0: same step (projection)
    int kd;                  // 128 or 64 (layer 0: only seg1)
    int t_len;
    int64_t rows;
    int64_t chunks_per_t;    // ceil(rows / 32)
    int64_t total_chunks;
};

template <int N>
__global__ void __launch_bounds__(kWgThreads, 1) lstm_wgrad_tc_kernel(const __grid_constant__ WgParams p) {
    using Cfg = WgCfg<N>;
    constexpr int kWgBBytes = Cfg::kBBytes;
    constexpr int kWgStageBytes = Cfg::kStageBytes;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the __shared__ address space (LDS/STS, not generic LD/ST)
    WgTail* tail = (WgTail*)(smem + (size_t)kWgStages * kWgStageBytes);
    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;
    constexpr int kMmaWarp = kWgLoaderWarps;
    constexpr int kLoaders = kWgLoaderWarps * 32;

    pdl_launch_dependents();
    if (tid == 0) {
        for (int s = 0; s < kWgStages; ++s) {
            mbar_init(&tail->full[s], kLoaders / 2);       // one loader group per chunk
            mbar_init(&tail->empty[s], 1);
        }
        mbar_init(&tail->done, 1);
        fence_barrier_init();
    }
    if (warp == kMmaWarp) tmem_alloc(&tail->tmem_base, Cfg::kTmemCols);
    pdl_wait();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tail->tmem_base;
    const bool has_work = (int64_t)blockIdx.x < p.total_chunks;

    if (warp < kMmaWarp) {
        // ===================== loaders: HBM rows -> tf32 hi/lo -> MN-major swizzled atoms =====================
        TC_PROF_DECL
        // two loader groups alternate row chunks (see the note in lstm_cell_tc_kernel about the proxy fence)
        constexpr int kGroups = 2, kGT = kLoaders / kGroups;
        constexpr int kNA = 1024 / kGT, kNB = (32 * N / 4) / kGT;
        static_assert(kNB >= 1, "loader mapping");
        const int ltid = tid;
        const int grp = ltid / kGT, gtid = ltid % kGT;
        const int64_t my_chunks = (p.total_chunks - (int64_t)blockIdx.x + gridDim.x - 1) / gridDim.x;
        for (int64_t j = grp; j < my_chunks; j += kGroups) {
            const int64_t chunk = blockIdx.x + j * gridDim.x;
            const int t = (int)(chunk / p.chunks_per_t);
            const int64_t r0 = (chunk % p.chunks_per_t) * kWgRows;
            const float* s0 = p.seg0 ? p.seg0 + (int64_t)t * p.rows * kHid : nullptr;
            const float* s1 = p.shift1 ? ((t > 0) ? p.seg1 + (int64_t)(t - 1) * p.rows * kHid : p.h0)
                                        : p.seg1 + (int64_t)t * p.rows * kHid;
            const float* dt = p.da + (int64_t)t * p.rows * N;
            float4 va[kNA], vb[kNB];
#pragma unroll
            for (int i = 0; i < kNA; ++i) {                   // A': 32 rows x 32 float4 (128 kd values), coalesced
                const int idx = gtid + i * kGT;
                const int row = idx >> 5, q = idx & 31;
                const int64_t r = r0 + row;
                // kd = 128: m 0..63 from seg0 (h_below), 64..127 from seg1 (h_prev); kd = 64: m 0..63 from seg1
                const float* src = (p.kd == 128) ? (q < 16 ? s0 : s1) : (q < 16 ? s1 : nullptr);
                va[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (src != nullptr && r < p.rows) va[i] = *reinterpret_cast<const float4*>(src + r * kHid + (q & 15) * 4);
            }
#pragma unroll
            for (int i = 0; i < kNB; ++i) {                   // B': 32 rows x N/4 float4, coalesced
                const int idx = gtid + i * kGT;
                const int row = idx / (N / 4), q = idx % (N / 4);
                const int64_t r = r0 + row;
                vb[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (r < p.rows) vb[i] = *reinterpret_cast<const float4*>(dt + r * N + q * 4);
            }
            const int s = (int)(j % kWgStages);
            const uint32_t ph = (uint32_t)(j / kWgStages) & 1;
            mbar_wait(&tail->empty[s], ph ^ 1, 0);
            uint8_t* st = smem + (size_t)s * kWgStageBytes;
#pragma unroll
            for (int i = 0; i < kNA; ++i) {
                const int idx = gtid + i * kGT;
                split_store(st, mn32_offset(idx & 31, idx >> 5, kWgRows), va[i]);      // hi at st, lo at st + kWgABytes
            }
#pragma unroll
            for (int i = 0; i < kNB; ++i) {
                const int idx = gtid + i * kGT;
                const uint32_t off = mn32_offset(idx % (N / 4), idx / (N / 4), kWgRows);
                float4 hi, lo;
                const float4 v = vb[i];
                hi.x = tf32_hi(v.x); hi.y = tf32_hi(v.y); hi.z = tf32_hi(v.z); hi.w = tf32_hi(v.w);
                lo.x = tf32_lo(v.x, hi.x); lo.y = tf32_lo(v.y, hi.y); lo.z = tf32_lo(v.z, hi.z); lo.w = tf32_lo(v.w, hi.w);
                *reinterpret_cast<float4*>(st + 2 * kWgABytes + off) = hi;
                *reinterpret_cast<float4*>(st + 2 * kWgABytes + kWgBBytes + off) = lo;
            }
            fence_proxy_async_smem();
            mbar_arrive(&tail->full[s]);
        }
        TC_PROF_FLUSH(9, ltid == 0)
        // ===================== epilogue (warps 0-3): accumulator rows = kd index -> red.add into dWp =====================
        if (warp < 4 && has_work) {
            mbar_wait(&tail->done, 0, 3);
            tc_fence_after();
            const int m = warp * 32 + lane;
            const uint32_t t_row = tmem_base + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
            for (int chunk32 = 0; chunk32 < N / 32; ++chunk32) {
                uint32_t v[32];
                tmem_ld32(t_row + chunk32 * 32, v);
                tmem_ld_wait();
                if (m < p.kd) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) atomicAdd(p.dwp + (int64_t)m * N + chunk32 * 32 + j, __uint_as_float(v[j]));
                }
            }
        }
    } else {
        // ===================== MMA issuer =====================
        constexpr uint32_t idesc = idesc_tf32(128, N, 1);          // both operands MN-major
        constexpr uint32_t kLbo = (kWgRows / 4) * 512, kSbo = 512;
        TC_PROF_DECL
        uint32_t it = 0;
        for (int64_t chunk = blockIdx.x; chunk < p.total_chunks; chunk += gridDim.x, ++it) {
            const int s = it % kWgStages;
            const uint32_t ph = (it / kWgStages) & 1;
            mbar_wait(&tail->full[s], ph, 1);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t st = smem_u32(smem + (size_t)s * kWgStageBytes);
#pragma unroll
                for (int pass = 0; pass < 3; ++pass) {
                    const uint32_t a_base = st + ((pass == 1) ? kWgABytes : 0);
                    const uint32_t b_base = st + 2 * kWgABytes + ((pass == 2) ? kWgBBytes : 0);
#pragma unroll
                    for (int ks = 0; ks < kWgRows / 8; ++ks) {      // one MMA consumes K = 8 rows = two 4-row atoms
                        const uint64_t da = smem_desc_mn_sw128(a_base + ks * 2 * kSbo, kLbo, kSbo, 1);
                        const uint64_t db = smem_desc_mn_sw128(b_base + ks * 2 * kSbo, kLbo, kSbo, 1);
                        mma_tf32(tmem_base, da, db, idesc, (it > 0 || pass > 0 || ks > 0) ? 1u : 0u);
                    }
                }
                mma_commit(&tail->empty[s]);
            }
            __syncwarp();
        }
        if (lane == 0 && has_work) mma_commit(&tail->done);
        __syncwarp();
        TC_PROF_FLUSH(10, lane == 0)
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == kMmaWarp) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}


// ---- TMA-fed variant of the LSTM weight-gradient reduction (N = 256); opt-in, measured slower (see launch_wgrad_tc) -----
// Same MMA formulation, K = 16 rows per operand stage.  A producer thread streams every 16-row item -- the rows of
// h_below / h_{t-1} (2 x 4 KB) and of dA (16 KB) are contiguous in HBM, so plain bulk copies do -- into a ring of four
// 24 KB raw slots, up to four items ahead of the arithmetic; the 16 loader warps (one group) only read shared memory,
// split to tf32 hi/lo and write the MN-major operand atoms.
constexpr int kWtRows = 16;
constexpr int kWtStages = 2;
constexpr int kWtSlots = 4;
constexpr int kWtN = kGateCols;
constexpr int kWtABytes = 128 * kWtRows * 4;                      // 8 KB
constexpr int kWtBBytes = kWtN * kWtRows * 4;                     // 16 KB
constexpr int kWtStageBytes = 2 * kWtABytes + 2 * kWtBBytes;      // 48 KB
constexpr int kWtRawSeg = kWtRows * kHid * 4;                     // 4 KB
constexpr int kWtRawDa = kWtRows * kWtN * 4;                      // 16 KB
constexpr int kWtRawBytes = 2 * kWtRawSeg + kWtRawDa;             // 24 KB
constexpr int kWtThreads = (kWgLoaderWarps + 2) * 32;             // 16 loader warps + MMA warp + producer warp
struct WtTail {
    uint64_t full[kWtStages];
    uint64_t empty[kWtStages];
    uint64_t raw_full[kWtSlots];
    uint64_t raw_empty[kWtSlots];
    uint64_t done;
    uint32_t tmem_base;
};
constexpr size_t kWtSmem = 1024 + (size_t)kWtStages * kWtStageBytes + (size_t)kWtSlots * kWtRawBytes + sizeof(WtTail);
static_assert(kWtSmem <= 232448, "wgrad TMA kernel exceeds the 227 KB shared-memory limit");

__global__ void __launch_bounds__(kWtThreads, 1) lstm_wgrad_tma_kernel(const __grid_constant__ WgParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the __shared__ address space (LDS/STS, not generic LD/ST)
    uint8_t* raw = smem + (size_t)kWtStages * kWtStageBytes;
    WtTail* tail = (WtTail*)(raw + (size_t)kWtSlots * kWtRawBytes);
    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;
    constexpr int kMmaWarp = kWgLoaderWarps;
    constexpr int kLoaders = kWgLoaderWarps * 32;
    constexpr int N = kWtN;

    pdl_launch_dependents();
    if (tid == 0) {
        for (int s = 0; s < kWtStages; ++s) {
            mbar_init(&tail->full[s], kLoaders);
            mbar_init(&tail->empty[s], 1);
        }
        for (int s = 0; s < kWtSlots; ++s) {
            mbar_init(&tail->raw_full[s], 1);
            mbar_init(&tail->raw_empty[s], kLoaders);
        }
        mbar_init(&tail->done, 1);
        fence_barrier_init();
    }
    if (warp == kMmaWarp) tmem_alloc(&tail->tmem_base, N);
    pdl_wait();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tail->tmem_base;
    const int64_t chunks_per_t = (p.rows + kWtRows - 1) / kWtRows;
    const int64_t total_chunks = chunks_per_t * p.t_len;
    const bool has_work = (int64_t)blockIdx.x < total_chunks;
    const int64_t my_chunks = has_work ? (total_chunks - (int64_t)blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    if (warp < kMmaWarp) {
        // ===================== loaders: raw slot -> tf32 hi/lo -> MN-major swizzled atoms =====================
        TC_PROF_DECL
        for (int64_t j = 0; j < my_chunks; ++j) {
            const int64_t chunk = blockIdx.x + j * gridDim.x;
            const int t = (int)(chunk / chunks_per_t);
            const int64_t r0 = (chunk % chunks_per_t) * kWtRows;
            const int nrow = (int)((p.rows - r0) < kWtRows ? (p.rows - r0) : kWtRows);
            const bool have0 = (p.kd == 128) && p.seg0 != nullptr;
            const bool have1 = p.shift1 ? ((t > 0) || p.h0 != nullptr) : true;
            const int slot = (int)(j % kWtSlots);
            const uint8_t* rs = raw + (size_t)slot * kWtRawBytes;
            mbar_wait(&tail->raw_full[slot], (uint32_t)((j / kWtSlots) & 1), 1);
            float4 va, vb[2];
            {
                const int row = tid >> 5, q = tid & 31;        // A': 16 rows x 32 float4 (128 kd values)
                // kd = 128: m 0..63 from seg0 (h_below), 64..127 from seg1 (h_prev); kd = 64: m 0..63 from seg1
                const bool from1 = (p.kd == 128) ? (q >= 16) : (q < 16);
                const bool ok = row < nrow && ((p.kd == 128) ? (from1 ? have1 : have0) : (from1 && have1));
                va = make_float4(0.f, 0.f, 0.f, 0.f);
                if (ok) va = *reinterpret_cast<const float4*>(rs + (from1 ? kWtRawSeg : 0) + row * (kHid * 4) + (q & 15) * 16);
            }
#pragma unroll
            for (int i = 0; i < 2; ++i) {                      // B': 16 rows x 64 float4
                const int idx = tid + i * kLoaders;
                const int row = idx >> 6, q = idx & 63;
                vb[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (row < nrow) vb[i] = *reinterpret_cast<const float4*>(rs + 2 * kWtRawSeg + row * (N * 4) + q * 16);
            }
            mbar_arrive(&tail->raw_empty[slot]);
            const int s = (int)(j % kWtStages);
            const uint32_t ph = (uint32_t)(j / kWtStages) & 1;
            mbar_wait(&tail->empty[s], ph ^ 1, 0);
            uint8_t* st = smem + (size_t)s * kWtStageBytes;
            {
                float4 hi, lo;
                hi.x = tf32_hi(va.x); hi.y = tf32_hi(va.y); hi.z = tf32_hi(va.z); hi.w = tf32_hi(va.w);
                lo.x = tf32_lo(va.x, hi.x); lo.y = tf32_lo(va.y, hi.y); lo.z = tf32_lo(va.z, hi.z); lo.w = tf32_lo(va.w, hi.w);
                const uint32_t off = mn32_offset(tid & 31, tid >> 5, kWtRows);
                *reinterpret_cast<float4*>(st + off) = hi;
                *reinterpret_cast<float4*>(st + kWtABytes + off) = lo;
            }
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int idx = tid + i * kLoaders;
                const uint32_t off = mn32_offset(idx & 63, idx >> 6, kWtRows);
                float4 hi, lo;
                const float4 v = vb[i];
                hi.x = tf32_hi(v.x); hi.y = tf32_hi(v.y); hi.z = tf32_hi(v.z); hi.w = tf32_hi(v.w);
                lo.x = tf32_lo(v.x, hi.x); lo.y = tf32_lo(v.y, hi.y); lo.z = tf32_lo(v.z, hi.z); lo.w = tf32_lo(v.w, hi.w);
                *reinterpret_cast<float4*>(st + 2 * kWtABytes + off) = hi;
                *reinterpret_cast<float4*>(st + 2 * kWtABytes + kWtBBytes + off) = lo;
            }
            fence_proxy_async_smem();
            mbar_arrive(&tail->full[s]);
        }
        TC_PROF_FLUSH(9, tid == 0)
        // ===================== epilogue (warps 0-3): accumulator rows = kd index -> red.add into dWp =====================
        if (warp < 4 && has_work) {
            mbar_wait(&tail->done, 0, 3);
            tc_fence_after();
            const int m = warp * 32 + lane;
            const uint32_t t_row = tmem_base + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
            for (int chunk32 = 0; chunk32 < N / 32; ++chunk32) {
                uint32_t v[32];
                tmem_ld32(t_row + chunk32 * 32, v);
                tmem_ld_wait();
                if (m < p.kd) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) atomicAdd(p.dwp + (int64_t)m * N + chunk32 * 32 + j, __uint_as_float(v[j]));
                }
            }
        }
    } else if (warp == kMmaWarp) {
        // ===================== MMA issuer =====================
        constexpr uint32_t idesc = idesc_tf32(128, N, 1);          // both operands MN-major
        constexpr uint32_t kLbo = (kWtRows / 4) * 512, kSbo = 512;
        TC_PROF_DECL
        for (int64_t it = 0; it < my_chunks; ++it) {
            const int s = (int)(it % kWtStages);
            const uint32_t ph = (uint32_t)(it / kWtStages) & 1;
            mbar_wait(&tail->full[s], ph, 1);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t st = smem_u32(smem + (size_t)s * kWtStageBytes);
#pragma unroll
                for (int pass = 0; pass < 3; ++pass) {
                    const uint32_t a_base = st + ((pass == 1) ? kWtABytes : 0);
                    const uint32_t b_base = st + 2 * kWtABytes + ((pass == 2) ? kWtBBytes : 0);
#pragma unroll
                    for (int ks = 0; ks < kWtRows / 8; ++ks) {      // one MMA consumes K = 8 rows = two 4-row atoms
                        const uint64_t da = smem_desc_mn_sw128(a_base + ks * 2 * kSbo, kLbo, kSbo, 1);
                        const uint64_t db = smem_desc_mn_sw128(b_base + ks * 2 * kSbo, kLbo, kSbo, 1);
                        mma_tf32(tmem_base, da, db, idesc, (it > 0 || pass > 0 || ks > 0) ? 1u : 0u);
                    }
                }
                mma_commit(&tail->empty[s]);
            }
            __syncwarp();
        }
        if (lane == 0 && has_work) mma_commit(&tail->done);
        __syncwarp();
        TC_PROF_FLUSH(10, lane == 0)
    } else {
        // ===================== producer: bulk copies of the raw rows, up to kWtSlots items ahead =====================
        if (lane == 0) {
            for (int64_t j = 0; j < my_chunks; ++j) {
                const int64_t chunk = blockIdx.x + j * gridDim.x;
                const int t = (int)(chunk / chunks_per_t);
                const int64_t r0 = (chunk % chunks_per_t) * kWtRows;
                const uint32_t nrow = (uint32_t)((p.rows - r0) < kWtRows ? (p.rows - r0) : kWtRows);
                const float* s0 = (p.kd == 128 && p.seg0) ? p.seg0 + (int64_t)t * p.rows * kHid : nullptr;
                const float* s1 = p.shift1 ? ((t > 0) ? p.seg1 + (int64_t)(t - 1) * p.rows * kHid : p.h0)
                                            : p.seg1 + (int64_t)t * p.rows * kHid;
                const float* dt = p.da + (int64_t)t * p.rows * N;
                const int slot = (int)(j % kWtSlots);
                const int64_t n = j / kWtSlots;
                if (n > 0) mbar_wait_raw(&tail->raw_empty[slot], (uint32_t)((n - 1) & 1));
                uint8_t* rs = raw + (size_t)slot * kWtRawBytes;
                uint64_t* fb = &tail->raw_full[slot];
                const uint32_t seg_b = nrow * kHid * 4, da_b = nrow * N * 4;
                mbar_arrive_expect_tx(fb, (s0 ? seg_b : 0u) + (s1 ? seg_b : 0u) + da_b);
                if (s0) bulk_g2s(rs, s0 + r0 * kHid, seg_b, fb);
                if (s1) bulk_g2s(rs + kWtRawSeg, s1 + r0 * kHid, seg_b, fb);
                bulk_g2s(rs + 2 * kWtRawSeg, dt + r0 * N, da_b, fb);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == kMmaWarp) tmem_dealloc(tmem_base, N);
}

}  // namespace

namespace stmgcn {

// Optional launch with the programmatic-stream-serialization attribute (STMGCN_PDL=1): consecutive layer-step kernels
// overlap the next kernel's prologue with the previous kernel's tail; every kernel calls pdl_wait() before it touches
// global memory, so the data dependencies of the stream order are kept.  Measured on B200 (cfg3, same box, A/B):
// 63.7 ms with the attribute vs 63.3 ms without -- the persistent one-CTA-per-SM grids leave nothing to overlap -- so
// it is off by default.
static int pdl_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("STMGCN_PDL");
        v = (e != nullptr && e[0] == '1') ? 1 : 0;
    }
    return v;
}
template <typename P>
static cudaError_t launch_pdl(void (*kernel)(P), int grid, int threads, size_t smem, cudaStream_t st, const P& p) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, p);
}

static int tc_prefetch_enabled() {
    static int v = -1;
    if (v < 0) {
        // measured on B200 at cfg3 (same box, alternating runs): 72.27 ms/step with the prefetch, 71.03 ms without --
        // the loads already stream well from DRAM and the prefetches only add L2 traffic.  Off unless asked for.
        const char* e = getenv("STMGCN_TC_PREFETCH");
        v = (e && e[0] == '1') ? 1 : 0;
    }
    return v;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)sym;
    }
    return fn;
}
// (rows, cols) fp32 row-major matrix, box = box_rows x box_cols
static bool make_tile_map(CUtensorMap* map, const float* base, int64_t rows, int cols, int box_cols, int box_rows,
                          CUtensorMapSwizzle swz) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (fn == nullptr) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)cols * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_NONE,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
static bool make_gates_map(CUtensorMap* map, float* base, int64_t rows, int box_cols, int box_rows, CUtensorMapSwizzle swz) {
    return make_tile_map(map, base, rows, kGateCols, box_cols, box_rows, swz);
}

static int env_flag(const char* name, int dflt) {
    const char* e = getenv(name);
    return e == nullptr ? dflt : (e[0] != '0');
}
static int fwd_tma_enabled() {        // forward: A operand tiles through TMA tensor loads (STMGCN_FWD_TMA=0 disables)
    static int v = -1;
    if (v < 0) v = env_flag("STMGCN_FWD_TMA", 1);
    return v;
}
static int gates_tma_enabled() {      // forward: gate tape through TMA tensor stores (STMGCN_GATES_TMA=0 disables)
    static int v = -1;
    if (v < 0) v = env_flag("STMGCN_GATES_TMA", 1);
    return v;
}
static int bwd_tma_enabled() {        // backward: TMA-fed loaders (STMGCN_BWD_TMA=0 selects the register-load variant)
    static int v = -1;
    if (v < 0) v = env_flag("STMGCN_BWD_TMA", 1);
    return v;
}

// Called from stmgcn_lstm_step_fwd (lstm.cu) when the tensor-core path applies.  aux != 0 (layer 0): the weight image
// carries a third k-block [W_ih^T ; b ; 0] and the loader feeds [x*s | 1 | 0] so x.W_ih + b comes out of the MMA.
int32_t launch_lstm_cell_tc(const float* seg0, const float* seg1, int nseg, int aux, const float* wimg, const float* bias,
                            const float* xo, const float* sg, int c_in, int t, int t_len, int64_t b_inner,
                            const float* c_prev, float* h_out, float* c_out, float* gates_out, int64_t rows,
                            int blocked_cs, cudaStream_t st) {
    static bool attr_done = false;
    if (!attr_done) {
        STMGCN_CUDA(cudaFuncSetAttribute(lstm_cell_tc_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFwdSmem));
        STMGCN_CUDA(cudaFuncSetAttribute(lstm_cell_tc_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFwdSmem));
        STMGCN_CUDA(cudaFuncSetAttribute(lstm_cell_tc_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFwdSmem));
        STMGCN_CUDA(cudaFuncSetAttribute(lstm_cell_tc_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFwdSmem));
        attr_done = true;
    }
    CellParams p;
    p.seg0 = seg0;
    p.seg1 = seg1;
    p.nkb = 2 * nseg + (aux ? 1 : 0);
    p.aux = aux;
    p.wimg = wimg;
    p.bias = aux ? nullptr : bias;
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
    p.blocked_cs = blocked_cs;
    p.prefetch = tc_prefetch_enabled();
    p.gates_tma = 0;
    {
        static int skip = -1;
        if (skip < 0) {          // 1: drop the h/c stores; 2: skip the staging-tile reuse waits (racy) -- timing experiments only
            const char* e = getenv("STMGCN_DBG_SKIP_HC");
            skip = e ? atoi(e) : 0;
        }
        p.dbg_skip_hc = skip;
    }
    memset(&p.gates_map, 0, sizeof(p.gates_map));
    if (gates_out != nullptr && gates_tma_enabled() && make_gates_map(&p.gates_map, gates_out, rows, 16, 32, CU_TENSOR_MAP_SWIZZLE_64B))
        p.gates_tma = 1;
    const int grid = p.n_tiles < sm_count() ? p.n_tiles : sm_count();
    // h / c through TMA stores (STMGCN_HC_TMA=0 disables): box = 32 rows x 4 units
    p.hc_tma = 0;
    memset(&p.h_map, 0, sizeof(p.h_map));
    memset(&p.c_map, 0, sizeof(p.c_map));
    {
        static int hc = -1;
        if (hc < 0) hc = env_flag("STMGCN_HC_TMA", 1);
        if (hc && p.gates_tma && p.dbg_skip_hc != 1 &&
            make_tile_map(&p.h_map, h_out, rows, kHid, 4, 32, CU_TENSOR_MAP_SWIZZLE_NONE) &&
            (blocked_cs ? make_tile_map(&p.c_map, c_out, (int64_t)p.n_tiles * 8 * kTileM, 8, 4, 32, CU_TENSOR_MAP_SWIZZLE_NONE)
                        : make_tile_map(&p.c_map, c_out, rows, kHid, 4, 32, CU_TENSOR_MAP_SWIZZLE_NONE)))
            p.hc_tma = 1;
    }
    // A operand through TMA: one tensor map per (rows, 64) hidden-state slice (absent segments are built as zeros)
    memset(&p.seg0_map, 0, sizeof(p.seg0_map));
    memset(&p.seg1_map, 0, sizeof(p.seg1_map));
    bool a_tma = fwd_tma_enabled() != 0;
    if (a_tma && seg0) a_tma = make_tile_map(&p.seg0_map, seg0, rows, kHid, kKB, kTileM, CU_TENSOR_MAP_SWIZZLE_128B);
    if (a_tma && seg1) a_tma = make_tile_map(&p.seg1_map, seg1, rows, kHid, kKB, kTileM, CU_TENSOR_MAP_SWIZZLE_128B);
    static int nsplit = -1;               // resident-weight N-split kernel: correct but measured slower, opt-in
    if (nsplit < 0) nsplit = env_flag("STMGCN_FWD_NSPLIT", 0);
    if (nsplit && p.gates_tma && p.hc_tma && a_tma && sm_count() >= 2) {
        static bool ns_attr = false;
        if (!ns_attr) {
            STMGCN_CUDA(cudaFuncSetAttribute(lstm_cell_nsplit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kNsSmem));
            ns_attr = true;
        }
        int pairs = sm_count() / 2;
        if (pairs > p.n_tiles) pairs = p.n_tiles;
        STMGCN_CUDA(launch_pdl(lstm_cell_nsplit_kernel, 2 * pairs, kNsThreads, kNsSmem, st, p));
        count_launch();
        return check_launch("lstm_cell_nsplit");
    }
    if (p.gates_tma) {
        if (a_tma) STMGCN_CUDA(launch_pdl(lstm_cell_tc_kernel<true, true>, grid, FwdCfg<true>::kThreads, kFwdSmem, st, p));
        else STMGCN_CUDA(launch_pdl(lstm_cell_tc_kernel<true, false>, grid, FwdCfg<false>::kThreads, kFwdSmem, st, p));
    } else {
        if (a_tma) STMGCN_CUDA(launch_pdl(lstm_cell_tc_kernel<false, true>, grid, FwdCfg<true>::kThreads, kFwdSmem, st, p));
        else STMGCN_CUDA(launch_pdl(lstm_cell_tc_kernel<false, false>, grid, FwdCfg<false>::kThreads, kFwdSmem, st, p));
    }
    count_launch();
    return check_launch("lstm_cell_tc");
}

// Called from stmgcn_lstm_step_bwd (lstm.cu).  kd = 128 (layers > 0) or 64 (layer 0).
int32_t launch_lstm_bwd_tc(int kd, float* gates, const float* c_t, const float* c_prev, const float* dh_in,
                           float* dh_rec, float* dc, float* dx_out, const float* wimg_t, float* dbp, const float* wx,
                           float* dwx, const float* xo, const float* sg, float* d_s, int c_in, int t, int t_len,
                           int64_t b_inner, int64_t rows, int blocked, cudaStream_t st) {
    static bool attr_done = false;
    if (!attr_done) {
        STMGCN_CUDA(cudaFuncSetAttribute(lstm_bwd_tc_kernel<128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)BwdCfg<128, false>::kSmem));
        STMGCN_CUDA(cudaFuncSetAttribute(lstm_bwd_tc_kernel<64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)BwdCfg<64, false>::kSmem));
        STMGCN_CUDA(cudaFuncSetAttribute(lstm_bwd_tc_kernel<128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)BwdCfg<128, true>::kSmem));
        STMGCN_CUDA(cudaFuncSetAttribute(lstm_bwd_tc_kernel<64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)BwdCfg<64, true>::kSmem));
        attr_done = true;
    }
    BwdParams p;
    p.gates = gates;
    p.c_t = c_t;
    p.c_prev = c_prev;
    p.dh_in = dh_in;
    p.dh_rec = dh_rec;
    p.dc = dc;
    p.dx_out = dx_out;
    p.wimg_t = wimg_t;
    p.dbp = dbp;
    p.wx = wx;
    p.dwx = dwx;
    p.xo = xo;
    p.sg = sg;
    p.d_s = d_s;
    p.c_in = c_in;
    p.t = t;
    p.t_len = t_len;
    p.b_inner = b_inner;
    p.rows = rows;
    p.n_tiles = (int)ceil_div(rows, kTileM);
    p.prefetch = tc_prefetch_enabled();
    p.blocked = blocked;
    p.first = (t == t_len - 1);
    const int grid = p.n_tiles < sm_count() ? p.n_tiles : sm_count();
    // TMA-fed loaders need the tile-blocked workspaces (contiguous 4 KB slices) and a tensor map of the gate slice
    memset(&p.gates_map, 0, sizeof(p.gates_map));
    const bool tma = blocked && bwd_tma_enabled() && make_gates_map(&p.gates_map, gates, rows, kKB, kTileM, CU_TENSOR_MAP_SWIZZLE_NONE);
    if (kd == 128) {
        if (tma) STMGCN_CUDA(launch_pdl(lstm_bwd_tc_kernel<128, true>, grid, BwdCfg<128, true>::kThreads, BwdCfg<128, true>::kSmem, st, p));
        else STMGCN_CUDA(launch_pdl(lstm_bwd_tc_kernel<128, false>, grid, BwdCfg<128, false>::kThreads, BwdCfg<128, false>::kSmem, st, p));
    } else {
        if (tma) STMGCN_CUDA(launch_pdl(lstm_bwd_tc_kernel<64, true>, grid, BwdCfg<64, true>::kThreads, BwdCfg<64, true>::kSmem, st, p));
        else STMGCN_CUDA(launch_pdl(lstm_bwd_tc_kernel<64, false>, grid, BwdCfg<64, false>::kThreads, BwdCfg<64, false>::kSmem, st, p));
    }
    count_launch();
    return check_launch("lstm_bwd_tc");
}

int lstm_tc_max_c_bwd() { return kBwdMaxC; }

// Weight-gradient reduction on the tensor cores.  LSTM (n = 256, shift1 = 1): called from stmgcn_lstm_wgrad;
// projection (n = 64, shift1 = 0, t_len = 1): called from stmgcn_proj_bwd, once per 128-row block of dW.
int32_t launch_wgrad_tc(const float* seg0, const float* seg1, const float* h0, int shift1, const float* da, int n,
                        float* dwp, int kd, int t_len, int64_t rows, cudaStream_t st) {
    static bool attr_done = false;
    if (!attr_done) {
        STMGCN_CUDA(cudaFuncSetAttribute(lstm_wgrad_tc_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)WgCfg<256>::kSmem));
        STMGCN_CUDA(cudaFuncSetAttribute(lstm_wgrad_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)WgCfg<64>::kSmem));
        STMGCN_CUDA(cudaFuncSetAttribute(lstm_wgrad_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWtSmem));
        attr_done = true;
    }
    WgParams p;
    p.seg0 = seg0;
    p.seg1 = seg1;
    p.h0 = h0;
    p.da = da;
    p.dwp = dwp;
    p.shift1 = shift1;
    p.kd = kd;
    p.t_len = t_len;
    p.rows = rows;
    p.chunks_per_t = ceil_div(rows, kWgRows);
    p.total_chunks = p.chunks_per_t * t_len;
    const int64_t grid = p.total_chunks < sm_count() ? p.total_chunks : sm_count();
    // TMA-fed variant for the LSTM shape: opt-in (STMGCN_WGRAD_TMA=1).  Measured on B200 at cfg3, same box: backward of one
    // branch 10.4 ms with it vs 9.9 ms with the register-load kernel -- its 16-row items double the per-byte cost of
    // fences and barriers, and the register-load loaders were 81 % busy, not starved.
    static int wg_tma = -1;
    if (wg_tma < 0) wg_tma = env_flag("STMGCN_WGRAD_TMA", 0);
    if (n == 256 && wg_tma && shift1 == 1) {
        const int64_t total16 = ceil_div(rows, kWtRows) * t_len;
        const int64_t grid16 = total16 < sm_count() ? total16 : sm_count();
        STMGCN_CUDA(launch_pdl(lstm_wgrad_tma_kernel, (int)grid16, kWtThreads, kWtSmem, st, p));
    } else if (n == 256)
        STMGCN_CUDA(launch_pdl(lstm_wgrad_tc_kernel<256>, (int)grid, kWgThreads, WgCfg<256>::kSmem, st, p));
    else
        STMGCN_CUDA(launch_pdl(lstm_wgrad_tc_kernel<64>, (int)grid, kWgThreads, WgCfg<64>::kSmem, st, p));
    count_launch();
    return check_launch("wgrad_tc");
}
int32_t launch_lstm_wgrad_tc(const float* seg0, const float* seg1, const float* h0, const float* da, float* dwp, int kd,
                             int t_len, int64_t rows, cudaStream_t st) {
    return launch_wgrad_tc(seg0, seg1, h0, 1, da, 256, dwp, kd, t_len, rows, st);
}

}  // namespace stmgcn

extern "C" int32_t stmgcn_lstm_pack_tc(const float* wp_fwd, int32_t kd_fwd, const float* wp_bwd, int32_t kd_bwd,
                                       int32_t hid, float* img_fwd, float* img_bwd, void* stream) {
    STMGCN_REQUIRE(wp_fwd && img_fwd, STMGCN_ERR_ARG, "lstm_pack_tc: null pointer");
    STMGCN_REQUIRE(hid == kHid && kd_fwd > 0 && kd_fwd % kKB == 0, STMGCN_ERR_SHAPE,
                   "lstm_pack_tc: tensor-core path needs hid == 64 and kd %% 32 == 0 (got hid=%d kd=%d)", hid, kd_fwd);
    cudaStream_t st = (cudaStream_t)stream;
    // forward operand B[n = gate col][k = kd index] = wp_fwd[k][n]
    pack_image_kernel<<<(kd_fwd * kGateCols + 255) / 256, 256, 0, st>>>(wp_fwd, kGateCols, kd_fwd, 1, kGateCols, img_fwd, kGateCols);
    count_launch();
    if (img_bwd != nullptr) {   // backward operand B[n = kd index][k = gate col] = wp_bwd[n][k]
        STMGCN_REQUIRE(wp_bwd && (kd_bwd == 64 || kd_bwd == 128), STMGCN_ERR_SHAPE, "lstm_pack_tc: kd_bwd=%d", kd_bwd);
        pack_image_kernel<<<(kd_bwd * kGateCols + 255) / 256, 256, 0, st>>>(wp_bwd, kd_bwd, kGateCols, kGateCols, 1, img_bwd, kd_bwd);
        count_launch();
    }
    return check_launch("lstm_pack_tc");
}

#ifdef STMGCN_TC_PROFILE
extern "C" int32_t stmgcn_dbg_tc_prof(unsigned long long* host_out, int32_t reset) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(host_out, stmgcn::tc::g_tc_prof, sizeof(unsigned long long) * 64);
    if (reset) {
        unsigned long long z[64] = {0};
        cudaMemcpyToSymbol(stmgcn::tc::g_tc_prof, z, sizeof(z));
    }
    return 0;
}
#endif
