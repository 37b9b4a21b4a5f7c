// K1 for narrow units (<= 16 mixtures per state), second generation: wide accumulators.
//
// Same arithmetic as score_tc.cu / score_simt.cu (LHMM.cal_observation_pro -> GMM.point ->
// gaussian_function, LHMM.py:163-187, Clustering.py:740-767, util.py:20-36,54-77):
//     c[t, g] = <[x_t (39), 1 | x_t^2 (39), 1], W_g>,     b[t, s] = logsumexp_{g in s} c[t, g]
// as the 3-product fp16 (hi, lo) contraction with fp32 accumulation in tensor memory.
//
// What changed against score_tc.cu (round 1: 36 % tensor-pipe activity at 16 mixtures): a tcgen05.mma fed
// from shared memory costs about 26 clk + operand bytes / 128 (the 128 B/clk shared-memory read port), so an
// N = 96 accumulator (two label positions) ran at 81 clk against a 48 clk tensor floor, and every utterance
// pulled 120 KB of pre-split fp16 frame-tile images plus its unit images through L2 -> SM.  Here
//   * PG label positions share one accumulator, N = PG * NPAD <= 256 (five positions = 240 columns at 16
//     mixtures): the 4 KB A read of an MMA is amortised over 240 columns (118 clk against a 120 clk floor);
//   * the unit images of the whole utterance stay resident (two slots of PG images; an utterance of <= 2*PG
//     positions loads them once), the frame tiles stream through ONE operand buffer;
//   * frame tiles arrive as fp32 (20 KB instead of 40 KB, quad-major so that shared-memory accesses are
//     conflict-free) and the 20 epilogue warps form the fp16 (hi, lo) [x | x^2] operand image between two
//     accumulators: 640 threads = 128 rows x 5 eight-feature chunks, ~60 instructions each.  (Four dedicated
//     converter warps, one row per thread, needed ~3 000 clk per tile next to the MUFU-bound epilogue warps
//     of their sub-partitions - as long as the two MMA batches of the tile; profiles/README.md.)
//
//   warp EW       TMA producer : per item the unit images (one cp.async.bulk per position), per tile one
//                                cp.async.bulk of its fp32 image
//   warp EW+1     MMA issuer   : 15 tcgen05.mma (M=128, N=PG*NPAD, K=16) per (tile, position group)
//   warps 0..EW-1 epilogue     : EPQ warps per TMEM lane quarter; a warp takes the positions p = w, w+EPQ, ..
//                                of the group: tcgen05.ld of a state's columns (one frame per thread),
//                                log-sum-exp, store of b.  When the LAST accumulator of a tile is complete
//                                (its commit covers every MMA that read the operand buffer) they first
//                                convert the next tile, then take the accumulator.
#include "tc_common.cuh"

#ifndef PC_K1W_POLY
#define PC_K1W_POLY 1  // exponentials on the FMA pipe: 0 none, 1 every fourth, 2 every second (measured at cfg 2: 125 / 119 / 125 us)
#endif

__device__ long long g_k1w_dbg[8192];

namespace {

using tc::T_KCH;
using tc::T_PIECE;
using tc::T_ROWS;
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;

template <int MIX>
struct WCfg {
    static constexpr int N_REAL = PC_EMIT * MIX;
    static constexpr int NPAD = (N_REAL + 15) & ~15;
    static constexpr int B_STAGE = 2 * T_KCH * NPAD * 16;  // one unit image (hi, lo)
    static constexpr int PG = MIX >= 16 ? 5 : (MIX >= 8 ? 7 : 11);  // positions per accumulator (odd: see NCOL)
    static constexpr int N_ACC = PG * NPAD;
    static constexpr int TM_STRIDE = 256, TM_BUFS = 2, TM_COLS = 512;
    static constexpr int NB = 2;  // B slots of PG unit images
    static constexpr int B_SLOT = PG * B_STAGE;
    static constexpr int RAW_BYTES = T_ROWS * PC_XS * 4;  // fp32 rows of one tile
    static constexpr int EPQ = PG < 5 ? PG : 5;           // epilogue warps per lane quarter
    static constexpr int EW = 4 * EPQ;
    static constexpr int W_PROD = EW, W_MMA = EW + 1;
    static constexpr int NTHREADS = (W_MMA + 1) * 32;
    static_assert(EW * 32 == T_ROWS * 5, "one (row, 8-feature chunk) per epilogue thread");
    // staging of an accumulator's emissions [128 rows][NCOL] for coalesced stores; an odd row stride keeps the
    // per-row writes of a warp (32 rows, one column) on 32 different banks
    static constexpr int NCOL = PC_EMIT * PG;
    static constexpr int STAGE_BYTES = T_ROWS * NCOL * 4;
    static constexpr int SMEM = 1024 + 2 * T_PIECE + RAW_BYTES + NB * B_SLOT + 2 * STAGE_BYTES;
    static_assert(NCOL % 2 == 1, "odd staging stride");
    static_assert(N_ACC <= 256 && N_ACC % 16 == 0, "accumulator width");
    static_assert(SMEM <= 227 * 1024, "shared memory budget");
};

struct WBars {
    uint64_t raw_full, raw_empty, a_full;
    uint64_t b_full[2], b_empty[2];
    uint64_t tm_full[2], tm_empty[2];
    uint32_t tmem_base;
};

template <int MIX, bool SCALED>
__device__ __forceinline__ float state_lse_w(const float (&v_in)[MIX], const float *__restrict__ scale) {
    float v[MIX];
#pragma unroll
    for (int e = 0; e < MIX; ++e) v[e] = SCALED ? v_in[e] * __ldg(scale + e) : v_in[e];
    return tc::lse_packed<MIX, PC_K1W_POLY>(v);  // every fourth exponential on the FMA pipe
}

template <int MIX>
__global__ void __launch_bounds__(WCfg<MIX>::NTHREADS, 1)
score_tc_wide_kernel(CorpusView v, const float *__restrict__ X, const float *__restrict__ W, int n_gauss,
                     float *__restrict__ b, int item_lo, int item_hi, int dbg) {
    using C = WCfg<MIX>;
    extern __shared__ __align__(1024) uint8_t smem[];
    WBars *bars = reinterpret_cast<WBars *>(smem);
    uint8_t *a_s = smem + 1024;               // operand image of the current tile: hi piece, lo piece
    uint8_t *raw_s = a_s + 2 * T_PIECE;       // fp32 rows of the next tile
    uint8_t *b_s = raw_s + C::RAW_BYTES;      // NB slots of PG unit images
    float *stage_s = reinterpret_cast<float *>(b_s + C::NB * C::B_SLOT);  // two emission staging buffers

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        tc::mbar_init(&bars->raw_full, 1); tc::mbar_init(&bars->raw_empty, C::EW);
        tc::mbar_init(&bars->a_full, C::EW);
        for (int i = 0; i < 2; ++i) {
            tc::mbar_init(&bars->b_full[i], 1); tc::mbar_init(&bars->b_empty[i], 1);
            tc::mbar_init(&bars->tm_full[i], 1); tc::mbar_init(&bars->tm_empty[i], C::EW);
        }
        tc::mbar_fence_init();
    }
    if (warp == C::W_MMA) tc::tmem_alloc(&bars->tmem_base, C::TM_COLS);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base_ = bars->tmem_base;

    const uint8_t *x32 = reinterpret_cast<const uint8_t *>(X) + pc_x32_offset(v.total_frames, v.n_xtiles);
    const uint8_t *w16 = reinterpret_cast<const uint8_t *>(W) + pc_w16_offset(n_gauss);
    const float *wscale = W + (size_t)n_gauss * PC_KA;
    const bool scaled_rows = reinterpret_cast<const int *>(wscale + n_gauss)[0] != 0;

    // running counters: tiles (raw / operand buffer phases), B slot fills, accumulators
    uint32_t n_tile = 0, n_bfill = 0, n_acc = 0;
    for (int item = item_lo + blockIdx.x; item < item_hi; item += gridDim.x) {
        const int u = v.sitem_utt[item];
        const int64_t f0 = v.frame_off[u];
        const int T = (int)(v.frame_off[u + 1] - f0);
        const int64_t p0 = v.pair_off[u];
        const int L_ = (int)(v.pair_off[u + 1] - p0);
        const int t_item = v.sitem_t0[item];
        const int nt_ = v.sitem_nt[item];
        const int NG_ = (L_ + C::PG - 1) / C::PG;        // position groups
        // resident: the item's unit images fit the B slots and are loaded once; otherwise they stream
        // through the slots once per tile
        const bool resident_ = NG_ <= C::NB;
        if (warp == C::W_PROD) {
            // ------------------------------------------------------------ TMA producer
            const int L = L_, nt = nt_, NG = NG_;
            auto load_b = [&](int g) {
                const int slot = n_bfill % C::NB;
                const int n_img = min(C::PG, L - g * C::PG);
                tc::mbar_wait(&bars->b_empty[slot], ((n_bfill / C::NB) & 1) ^ 1);
                if (lane == 0) {
                    tc::mbar_expect_tx(&bars->b_full[slot], n_img * C::B_STAGE);
                    for (int i = 0; i < n_img; ++i)
                        tc::tma_load_1d(b_s + slot * C::B_SLOT + i * C::B_STAGE,
                                        w16 + (size_t)v.labels[p0 + g * C::PG + i] * C::B_STAGE, C::B_STAGE,
                                        &bars->b_full[slot]);
                }
                ++n_bfill;
                __syncwarp();
            };
            auto load_raw = [&](int j) {  // the tile's fp32 image (quad-major, padding rows zero)
                const int64_t xt = v.xtile_off[u] + (t_item + j * T_ROWS) / T_ROWS;
                tc::mbar_wait(&bars->raw_empty, (n_tile & 1) ^ 1);
                if (lane == 0) {
                    tc::mbar_expect_tx(&bars->raw_full, PC_X32TILE_BYTES);
                    tc::tma_load_1d(raw_s, x32 + (size_t)xt * PC_X32TILE_BYTES, PC_X32TILE_BYTES, &bars->raw_full);
                }
                ++n_tile;
                __syncwarp();
            };
            // order = consumption order: first tile's rows, then the unit images, then the other tiles
            load_raw(0);
            if (resident_) {
                for (int g = 0; g < NG; ++g) load_b(g);
                for (int j = 1; j < nt; ++j) load_raw(j);
            } else {
                for (int j = 0; j < nt; ++j) {
                    if (j > 0) load_raw(j);
                    for (int g = 0; g < NG; ++g) load_b(g);
                }
            }
        } else if (warp == C::W_MMA) {
            // ------------------------------------------------------------ MMA issuer
            const int L = __reduce_max_sync(0xffffffffu, L_);
            const int nt = __reduce_max_sync(0xffffffffu, nt_);
            const int NG = (L + C::PG - 1) / C::PG;
            const bool resident = NG <= C::NB;
            const uint32_t tmem_base = __reduce_max_sync(0xffffffffu, tmem_base_);
            n_tile = __reduce_max_sync(0xffffffffu, n_tile);
            n_bfill = __reduce_max_sync(0xffffffffu, n_bfill);
            n_acc = __reduce_max_sync(0xffffffffu, n_acc);
            const uint32_t a_base = tc::smem_u32(a_s), b_base = tc::smem_u32(b_s);
            const uint32_t bfill0 = n_bfill;
            for (int j = 0; j < nt; ++j, ++n_tile) {
                if ((dbg & 32) && blockIdx.x == 0 && lane == 0 && n_tile < 500) g_k1w_dbg[6144 + n_tile] = clock64();
                tc::mbar_wait(&bars->a_full, n_tile & 1);
                for (int g = 0; g < NG; ++g, ++n_acc) {
                    const uint32_t fill = resident ? bfill0 + g : n_bfill;
                    const int slot = fill % C::NB, tb = n_acc % C::TM_BUFS;
                    const int n_img = min(C::PG, L - g * C::PG);
                    const uint32_t idesc = tc::umma_idesc_f16(T_ROWS, n_img * C::NPAD, 0, 0);
                    const bool rec = (dbg & 32) && blockIdx.x == 0 && lane == 0 && n_acc < 1000;
                    if (rec) g_k1w_dbg[n_acc * 4 + 0] = clock64();  // (includes the wait for the operand image when g == 0)
                    tc::mbar_wait(&bars->b_full[slot], (fill / C::NB) & 1);
                    if (rec) g_k1w_dbg[n_acc * 4 + 1] = clock64();
                    tc::mbar_wait(&bars->tm_empty[tb], ((n_acc / C::TM_BUFS) & 1) ^ 1);
                    if (rec) g_k1w_dbg[n_acc * 4 + 2] = clock64();
                    tc::tc_fence_after();
                    if (tc::elect_one()) {
                        const uint32_t d = tmem_base + tb * C::TM_STRIDE;
                        const uint32_t ah = a_base, al = a_base + T_PIECE;
                        const uint32_t bh = b_base + slot * C::B_SLOT, bl = bh + PC_WGROUP_BYTES / 2;
                        uint32_t accum = 0;
#pragma unroll
                        for (int q = 0; q < 3; ++q) {
                            const uint32_t ap = (q == 2) ? al : ah;
                            const uint32_t bp = (q == 1) ? bl : bh;
                            if (dbg & 2) break;
#pragma unroll
                            for (int k = 0; k < T_KCH / 2; ++k) {
                                const uint64_t ad = tc::umma_desc(ap + 2 * k * T_ROWS * 16, T_ROWS * 16, 128);
                                const uint64_t bd = tc::umma_desc(bp + 2 * k * 128, 128, PC_WGROUP_BYTES);
                                tc::mma_f16_ss(d, ad, bd, idesc, accum);
                                accum = 1;
                            }
                        }
                        tc::tc_commit(&bars->tm_full[tb]);
                        if (!resident || j == nt - 1) tc::tc_commit(&bars->b_empty[slot]);
                    }
                    if (rec) g_k1w_dbg[n_acc * 4 + 3] = clock64();
                    __syncwarp();
                    if (!resident) ++n_bfill;
                }
            }
            if (resident) n_bfill = bfill0 + NG;
        } else {
            // ------------------------------------------------------------ epilogue
            const int quarter = warp & 3, ew = warp >> 2;  // TMEM lane quarter == warp % 4
            const int r = quarter * 32 + lane;             // row of the tile == TMEM lane
            const int L = L_, NG = NG_;
            const int sp = pc_spad(L);
            const uint32_t tmem_base = tmem_base_;
            float *bu = b + v.emis_off[u];
            // the next tile in the FIFO of fp32 images -> operand buffer (thread = (row, 8-feature chunk))
            auto convert_next = [&]() {
                const bool rec = (dbg & 32) && blockIdx.x == 0 && threadIdx.x == 0 && n_tile < 500;
                if (rec) g_k1w_dbg[4096 + n_tile * 4 + 0] = clock64();
                tc::mbar_wait(&bars->raw_full, n_tile & 1);
                if (rec) g_k1w_dbg[4096 + n_tile * 4 + 1] = clock64();
                if (!(dbg & 8)) tc::convert_tile_chunk_qm(raw_s, threadIdx.x & (T_ROWS - 1), threadIdx.x / T_ROWS, a_s, a_s + T_PIECE);
                tc::fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    tc::mbar_arrive(&bars->a_full);
                    tc::mbar_arrive(&bars->raw_empty);
                }
                if (rec) g_k1w_dbg[4096 + n_tile * 4 + 3] = clock64();
                ++n_tile;
            };
            if (item == item_lo + (int)blockIdx.x) convert_next();  // the kernel's first tile: nothing to wait for
            for (int j = 0; j < nt_; ++j) {
                const int t0 = t_item + j * T_ROWS;
                const int rows = min(T_ROWS, T - t0);
                const bool live = quarter * 32 < rows;  // a quarter without frames skips the math
                for (int g = 0; g < NG; ++g, ++n_acc) {
                    const int tb = n_acc % C::TM_BUFS;
                    const int n_img = min(C::PG, L - g * C::PG);
                    const bool rece = (dbg & 32) && blockIdx.x == 0 && threadIdx.x == 0 && n_acc < 300;
                    if (rece) g_k1w_dbg[6656 + n_acc * 5 + 0] = clock64();
                    tc::mbar_wait(&bars->tm_full[tb], (n_acc / C::TM_BUFS) & 1);
                    tc::tc_fence_after();
                    if (rece) g_k1w_dbg[6656 + n_acc * 5 + 1] = clock64();
                    // every MMA that read the operand buffer has completed: refill it before the epilogue math
                    if (g == NG - 1 && (j + 1 < nt_ || item + (int)gridDim.x < item_hi)) convert_next();
                    if (rece) g_k1w_dbg[6656 + n_acc * 5 + 2] = clock64();
                    float *stage = stage_s + (n_acc & 1) * (T_ROWS * C::NCOL);
                    if (live && !(dbg & 16)) {
                        for (int i = ew; i < n_img; i += C::EPQ) {
                            const int p = g * C::PG + i;
                            const float *scale_p = wscale + (size_t)v.labels[p0 + p] * C::N_REAL;
#pragma unroll
                            for (int st = 0; st < PC_EMIT; ++st) {
                                const uint32_t taddr = tmem_base + tb * C::TM_STRIDE + i * C::NPAD + st * MIX +
                                                       ((uint32_t)(quarter * 32) << 16);
                                float vv[MIX];
                                if constexpr (MIX == 16) {
                                    tc::tmem_ld16(taddr, vv);
                                } else {
                                    float t8[8];
                                    tc::tmem_ld8(taddr, t8);
#pragma unroll
                                    for (int e = 0; e < MIX; ++e) vv[e] = t8[e];
                                }
                                tc::tmem_ld_wait();
                                stage[r * C::NCOL + PC_EMIT * i + st] =
                                    scaled_rows ? state_lse_w<MIX, true>(vv, scale_p + st * MIX)
                                                : state_lse_w<MIX, false>(vv, scale_p + st * MIX);
                            }
                        }
                    }
                    tc::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) tc::mbar_arrive(&bars->tm_empty[tb]);
                    if (rece) g_k1w_dbg[6656 + n_acc * 5 + 3] = clock64();
                    // the accumulator's emissions leave as runs of 3 * n_img consecutive floats per frame; the 32 frames
                    // of a lane quarter come from the EPQ warps of that quarter: one named barrier per quarter
                    asm volatile("bar.sync %0, %1;" ::"r"(1 + quarter), "n"(C::EPQ * 32) : "memory");
                    if (rece) g_k1w_dbg[6656 + n_acc * 5 + 4] = clock64();
                    if (!(dbg & (16 | 64))) {
                        // CPR threads per frame (a power of two >= NCOL: no division), EW*32 / CPR frames per pass
                        constexpr int CPR = C::NCOL <= 16 ? 16 : (C::NCOL <= 32 ? 32 : 64);
                        constexpr int RPP = C::EPQ * 32 / CPR;  // frames per pass of the quarter's threads
                        const int ncol = PC_EMIT * n_img;
                        const int tq = ew * 32 + lane;          // thread inside the quarter's group
                        const int c = tq & (CPR - 1);
                        float *o = bu + (size_t)t0 * sp + PC_EMIT * g * C::PG + c;
                        if (c < ncol) {
#pragma unroll
                            for (int rr = tq / CPR; rr < 32; rr += RPP) {
                                const int row = quarter * 32 + rr;
                                if (row < rows) o[(size_t)row * sp] = stage[row * C::NCOL + c];
                            }
                        }
                    }
                }
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == C::W_MMA) tc::tmem_dealloc(tmem_base_, C::TM_COLS);
}

template <int MIX>
int launch_wide(pc_handle h, const CorpusView &v, const float *X, const float *W, float *b, int item_lo,
                int item_hi, cudaStream_t st) {
    auto kern = score_tc_wide_kernel<MIX>;
    PC_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, WCfg<MIX>::SMEM));
    const int n = item_hi - item_lo;
    const int grid = n < h->sm_count ? n : h->sm_count;
    kern<<<grid, WCfg<MIX>::NTHREADS, WCfg<MIX>::SMEM, st>>>(v, X, W, v.n_units * PC_EMIT * MIX, b, item_lo, item_hi,
                                                            h->debug_flags);
    PC_LAUNCH_CHECK();
    h->launches++;
    return PC_OK;
}

}  // namespace

// block 0's clocks (option debug_flags & 32): per accumulator [4n..]: start, unit images landed, TMEM buffer free,
// MMAs issued; per tile [4096 + 4n..]: converter start, rows landed, operand buffer free, converted;
// [6144 + n]: MMA warp starts waiting for the tile's operand image
extern "C" int pc_debug_read_k1w(long long *host_out, int n) {
    return cudaMemcpyFromSymbol(host_out, g_k1w_dbg, sizeof(long long) * n) == cudaSuccess ? 0 : -2;
}

bool score_tc_wide_supported(int mix) { return mix == 4 || mix == 8 || mix == 16; }

int launch_score_tc_wide(pc_handle h, const CorpusView &v, const float *X, const float *W, int mix, float *b,
                         int item_lo, int item_hi, cudaStream_t st) {
    if (item_hi <= item_lo) return PC_OK;
    switch (mix) {
        case 4: return launch_wide<4>(h, v, X, W, b, item_lo, item_hi, st);
        case 8: return launch_wide<8>(h, v, X, W, b, item_lo, item_hi, st);
        case 16: return launch_wide<16>(h, v, X, W, b, item_lo, item_hi, st);
    }
    pc_set_error("launch_score_tc_wide: mix=%d not covered", mix);
    return PC_ERR_UNSUPPORTED;
}
