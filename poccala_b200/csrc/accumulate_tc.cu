// K3 (tensor-core generation): Baum-Welch sufficient statistics as two chained tcgen05
// contractions per 128-frame tile (FlashAttention-shaped, posteriors never leave the SM):
//
//   S[t, g]   = <[x_t | x_t^2], W_g>                      MMA1: frames x Gaussians, K = 80
//   P[t, g]   = exp(S[t, g] - b_t(state(g)) + lgam_t(state(g)))      (gamma_t(j, m), A.3)
//   D2[f, g] += sum_t [x_t | x_t^2][f] * P[t, g]          MMA2: features x Gaussians, K = frames
//
// Restates LHMM.update_acc -> Clustering.GMM.update_acc (LHMM.py:473-507, Clustering.py:653-680)
// in the linear-equivalent form of SURVEY A.4 (see accumulate_simt.cu).  All operands are fp16
// (hi, lo) pairs and every contraction is the 3-product error-compensated sum, fp32 accumulation
// in TMEM.  The frame tile that TMA drops into shared memory serves BOTH contractions: as the
// K-major A operand of MMA1 (K = feature) and, through an MN-major descriptor over the same bytes,
// as the A operand of MMA2 (M = feature, K = frame).  P (scaled by 2^15 to sit in the fp16 range)
// is written by the softmax warps as the MN-major B operand of MMA2.  Work is UNIT-major so that
// D2 stays in TMEM for a whole work item (<= 64 tiles of one unit) before one fp64-atomic flush.
//
//   warp 19     TMA producer : per item the unit's Gaussian rows (B), per tile 20 x 2 KiB of frames
//   warp 20     MMA issuer   : MMA1(i), then MMA2(i-1) (software pipelined), commits
//   warps 0-15  softmax      : two tile groups x (2 column halves x 4 lane quarters): tcgen05.ld S,
//                              exp2, hi/lo split, P tile
//   warps 16-18 flush        : D2 (lane = feature) -> atomicAdd(double) into acc, once per item
#include "tc_common.cuh"

__device__ long long g_acc_dbg[8192];

namespace {

using tc::T_KCH;
using tc::T_PIECE;
using tc::T_ROWS;
// warps 0-15 softmax, 16-18 flush, then the TMA producer and one MMA issuer per contraction: with a single
// issuer the barrier waits in front of each contraction (~400 clk per tile) left the tensor pipe idle
constexpr int W_FLUSH = 16, W_PROD = 19, W_MMA = 20, W_MMA2 = 21;
constexpr int NTHREADS = 22 * 32;
constexpr float LOG2E = 1.4426950408889634f;
// posteriors are stored as P * 2^15 so that the fp16 window [6e-8, 65504] covers [1.8e-12, 2]
constexpr float P_SHIFT = 15.f;
constexpr float P_UNSHIFT = 1.f / 32768.f;

// A (tile, unit) pair whose posteriors all vanish contributes exactly nothing: P = gamma_t(j,m) <=
// gamma_t(j), and the P tile stores P * 2^15 as fp16 (hi, lo), which is exactly (0, 0) below 2^-25.
// With every log gamma of the tile below log(2^-41) (one bit of margin for the rounding of S - b)
// the kernel would multiply by an all-zero P tile; such tiles are dropped from the work list.  On
// forced-alignment-like posteriors (a label position is occupied during a small part of its
// utterance) that is more than half of the tiles.
constexpr float ACTIVE_MIN_LGAM = PC_ACTIVE_MIN_LGAM;

// Four warps per 128-frame tile of an utterance (32 frames each): coalesced rows of log gamma (lane =
// state column), eight rows in flight per lane, running maximum per column, then
// active[(tile, position)] = 1 where any of the position's states exceeds ACTIVE_MIN_LGAM (the flags
// are zeroed by the launcher; NaN columns stay inactive).
constexpr int ACT_SPLIT = 4;
constexpr int ACT_THREADS = 128;
// Work items differ a lot in how many of their tiles are active; with a static round-robin over the
// SMs the kernel ends with its slowest SM.  The block that finishes last (ticket counter behind the
// item counts) orders the items by active tiles (counting sort, heaviest first); the main kernel deals
// them out in snake order.  Measured at the bench shape: 176 us against 192 us for a static round-robin
// over the run-major list and 202 us when only the items of one run of utterances are sorted - balance
// is worth more than the L2 hits of the per-position tile re-reads (DRAM reads 485 / 320 / 245 MB).
__global__ void __launch_bounds__(ACT_THREADS)
tile_active_kernel(CorpusView v, const float *__restrict__ lgam, int32_t *__restrict__ active, int dbg) {
    __shared__ int bin[64];
    __shared__ int is_last;
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int64_t xt = w / ACT_SPLIT;
    const int part = (int)(w - xt * ACT_SPLIT);
    if (xt < v.n_xtiles) {
        const int u = v.xtile_utt[xt], t0 = v.xtile_t0[xt];
        const int T = (int)(v.frame_off[u + 1] - v.frame_off[u]);
        const int64_t p0 = v.pair_off[u];
        const int L = (int)(v.pair_off[u + 1] - p0);
        const int sp = pc_spad(L), rows = min(PC_TILE_ROWS, T - t0);
        const int r_lo = part * (PC_TILE_ROWS / ACT_SPLIT), r_hi = min(rows, r_lo + PC_TILE_ROWS / ACT_SPLIT);
        const float *base = lgam + v.emis_off[u] + (size_t)t0 * sp;
        for (int c = lane; c < PC_EMIT * L && r_lo < r_hi; c += 32) {
            float m[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) m[j] = PC_NEG_INF;
            int r = r_lo;
            for (; r + 7 < r_hi; r += 8) {
#pragma unroll
                for (int j = 0; j < 8; ++j) m[j] = fmaxf(m[j], __ldg(base + (size_t)(r + j) * sp + c));
            }
            for (; r < r_hi; ++r) m[0] = fmaxf(m[0], __ldg(base + (size_t)r * sp + c));
#pragma unroll
            for (int j = 1; j < 8; ++j) m[0] = fmaxf(m[0], m[j]);
            if (m[0] > ACTIVE_MIN_LGAM) {
                const int64_t tile = v.pair_tile0[p0 + c / PC_EMIT] + t0 / PC_TILE_ROWS;
                if (dbg & 512) active[tile] = 1;
                else if (atomicExch(active + tile, 1) == 0) atomicAdd(v.item_act + v.tile_item[tile], 1);  // once per tile
            }
        }
    }
    if (dbg & 1024) return;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = atomicAdd(v.item_act + v.n_items, 1) == (int)gridDim.x - 1;
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    if (threadIdx.x == 0) v.item_act[v.n_items] = 0;  // the ticket counter is left clean for the next launch
    for (int i = threadIdx.x; i < 64; i += ACT_THREADS) bin[i] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < v.n_items; i += ACT_THREADS) atomicAdd(&bin[63 - min(__ldcg(v.item_act + i), 63)], 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        for (int k = 0; k < 64; ++k) { const int c = bin[k]; bin[k] = run; run += c; }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < v.n_items; i += ACT_THREADS)
        v.item_order[atomicAdd(&bin[63 - min(__ldcg(v.item_act + i), 63)], 1)] = i;
}

// For flags that K2 set (launch_forward_backward): one warp per item counts its active tiles, the block that
// finishes last orders the items.
__global__ void __launch_bounds__(1024)
item_order_kernel(CorpusView v) {
    __shared__ int bin[64];
    __shared__ int is_last;
    const int lane = threadIdx.x & 31;
    const int i = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (i < v.n_items) {
        const int64_t lo = v.item_tile_lo[i], hi = v.item_tile_lo[i + 1];  // <= 64 tiles
        const unsigned a0 = __ballot_sync(0xffffffffu, lo + lane < hi && __ldcg(v.tile_active + lo + lane) != 0);
        const unsigned a1 = __ballot_sync(0xffffffffu, lo + 32 + lane < hi && __ldcg(v.tile_active + lo + 32 + lane) != 0);
        if (lane == 0) v.item_act[i] = __popc(a0) + __popc(a1);
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = atomicAdd(v.item_act + v.n_items, 1) == (int)gridDim.x - 1;
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    if (threadIdx.x == 0) v.item_act[v.n_items] = 0;  // the ticket counter is left clean for the next launch
    for (int k = threadIdx.x; k < 64; k += blockDim.x) bin[k] = 0;
    __syncthreads();
    for (int k = threadIdx.x; k < v.n_items; k += blockDim.x) atomicAdd(&bin[63 - min(__ldcg(v.item_act + k), 63)], 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        for (int k = 0; k < 64; ++k) { const int c = bin[k]; bin[k] = run; run += c; }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < v.n_items; k += blockDim.x)
        v.item_order[atomicAdd(&bin[63 - min(__ldcg(v.item_act + k), 63)], 1)] = k;
}

// NC = Gaussians handled per work item (a unit's 3*MIX Gaussians, or a slice of them)
template <int MIX>
struct Cfg {
    static constexpr int N_UNIT = PC_EMIT * MIX;
    static constexpr int NC = N_UNIT <= 96 ? N_UNIT : 96;
    static constexpr int N_SLICES = (N_UNIT + NC - 1) / NC;
    static constexpr int NPAD = (NC + 15) & ~15;
    static constexpr int B_PIECE = T_KCH * NPAD * 16;
    static constexpr int P_PIECE = (NPAD / 8) * T_ROWS * 16;
    static constexpr int NA = NPAD <= 48 ? 3 : 2;
    // STACK: the (hi, lo) halves of the B operands are stacked along N, so one MMA of N = 2*NPAD
    // columns forms A*B_hi and A*B_lo at once and the A operand - whose 4 KB shared-memory read paces
    // an N = 48 MMA at ~65 clk - is fetched twice per K-step instead of three times:
    //   MMA1: S = (A_hi + A_lo) * [W_hi | W_lo]  (the unit image interleaves hi / lo row groups: SBO 1280)
    //   MMA2: D2 = (A_hi + A_lo)^T * [P_hi | P_lo]  (the lo piece of the P tile follows the hi piece)
    // The extra lo*lo products only add accuracy; the consumer adds the two column halves.
    static constexpr bool STACK = NPAD <= 64;
    static constexpr int S_COLS = STACK ? 2 * NPAD : NPAD;
    static constexpr int S_STRIDE = S_COLS <= 32 ? 32 : (S_COLS <= 64 ? 64 : 128);
    static constexpr int D2_COL = S_STRIDE * 2;
    static constexpr int D2_COLS = STACK ? 2 * NPAD : NPAD;
    static constexpr int TM_COLS = 512;
    static constexpr int SMEM = 1024 + NA * 2 * T_PIECE + 2 * B_PIECE + 2 * 2 * P_PIECE;
    static_assert(N_UNIT % NC == 0, "slices must tile the unit");
    static_assert(D2_COL + D2_COLS <= 512, "TMEM budget");
    static_assert(SMEM <= 227 * 1024, "shared memory budget");
};

struct Bars {
    uint64_t a_full[4], a_empty[4];
    uint64_t b_full, b_empty;
    uint64_t s_full[2], s_empty[2];
    uint64_t p_full[2], p_empty[2];
    uint64_t d2_full, d2_empty;
    uint32_t tmem_base;
    // per P buffer and lane quarter: which of the quarter's two 16-frame K-steps carry posterior mass
    uint8_t kact[2][4];
};

// S (TMEM, one frame per thread) -> P = 2^15 * exp(S*scale + lgam - b) -> fp16 hi / lo rows of the
// MN-major P tile, for the 8-Gaussian blocks [blk0, blk1) of the slice.  G0 = first Gaussian of the
// slice inside the unit (fixes the state of a column at compile time).
template <int MIX, int G0, bool SCALED, int HALF>
__device__ __forceinline__ void softmax_blocks(uint32_t taddr, const float (&dl)[PC_EMIT],
                                               const float *__restrict__ scale_g, uint8_t *ph,
                                               uint8_t *pl, int r) {
    using C = Cfg<MIX>;
    constexpr int NBLK = C::NPAD / 8;
    constexpr int PER = (NBLK + 1) / 2;
#pragma unroll
    for (int bb = 0; bb < PER; ++bb) {
        constexpr int blk0 = HALF * PER;
        const int blk = blk0 + bb;
        if (blk >= NBLK) break;
        float t8[8];
        if constexpr (C::STACK) {  // columns 16*blk + [0,8): A*W_hi, + [8,16): A*W_lo
            float t16[16];
            tc::tmem_ld16(taddr + blk * 16, t16);
            tc::tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 8; ++e) t8[e] = t16[e] + t16[8 + e];
        } else {
            tc::tmem_ld8(taddr + blk * 8, t8);
            tc::tmem_ld_wait();
        }
        uint32_t h[4], l[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float p[2];
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                // the state of a column is a compile-time constant after unrolling
                const int col_in_blk = 2 * e + k;
                const int col = blk * 8 + col_in_blk;
                const int st = (G0 + col) / MIX;
                const float d = dl[st < PC_EMIT ? st : PC_EMIT - 1];
                const float mul = SCALED ? __ldg(scale_g + col) * LOG2E : LOG2E;
                p[k] = (col < C::NC) ? tc::ex2(fmaf(t8[col_in_blk], mul, d)) : 0.f;
            }
            tc::split2(p[0], p[1], h[e], l[e]);
        }
        *reinterpret_cast<uint4 *>(ph + blk * T_ROWS * 16 + r * 16) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4 *>(pl + blk * T_ROWS * 16 + r * 16) = make_uint4(l[0], l[1], l[2], l[3]);
    }
}

template <int MIX>
__global__ void __launch_bounds__(NTHREADS, 1)
accumulate_tc_kernel(CorpusView v, const float *__restrict__ X, const float *__restrict__ W, int n_gauss,
                     const float *__restrict__ b, const float *__restrict__ lgam,
                     const int32_t *__restrict__ active, double *__restrict__ acc, int dbg) {
    using C = Cfg<MIX>;
    extern __shared__ __align__(1024) uint8_t smem[];
    Bars *bars = reinterpret_cast<Bars *>(smem);
    uint8_t *a_s = smem + 1024;
    uint8_t *b_s = a_s + C::NA * 2 * T_PIECE;
    uint8_t *p_s = b_s + 2 * C::B_PIECE;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 4; ++i) { tc::mbar_init(&bars->a_full[i], 1); tc::mbar_init(&bars->a_empty[i], 1); }
        tc::mbar_init(&bars->b_full, 1);
        tc::mbar_init(&bars->b_empty, 1);
        for (int i = 0; i < 2; ++i) {
            tc::mbar_init(&bars->s_full[i], 1); tc::mbar_init(&bars->s_empty[i], 8);
            tc::mbar_init(&bars->p_full[i], 8); tc::mbar_init(&bars->p_empty[i], 1);
        }
        tc::mbar_init(&bars->d2_full, 1);
        tc::mbar_init(&bars->d2_empty, 3);
        tc::mbar_fence_init();
    }
    // zero-fill the P tiles: Gaussian columns past the slice are never written
    for (int i = threadIdx.x; i < (2 * 2 * C::P_PIECE) / 16; i += NTHREADS)
        reinterpret_cast<uint4 *>(p_s)[i] = make_uint4(0u, 0u, 0u, 0u);
    tc::fence_proxy_async();
    if (warp == W_MMA) tc::tmem_alloc(&bars->tmem_base, C::TM_COLS);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base_ = bars->tmem_base;
    const uint32_t tmem_base = tmem_base_;

    const uint8_t *x16 = reinterpret_cast<const uint8_t *>(X) + pc_x16_offset(v.total_frames);
    const uint8_t *w16 = reinterpret_cast<const uint8_t *>(W) + pc_w16_offset(n_gauss);
    const float *wscale = W + (size_t)n_gauss * PC_KA;
    const bool scaled_rows = reinterpret_cast<const int *>(wscale + n_gauss)[0] != 0;
    constexpr int UNIT_IMG = ((C::N_UNIT + 15) & ~15) / 8 * PC_WGROUP_BYTES;  // bytes of a unit image
    constexpr uint32_t B_BYTES = C::NPAD / 8 * PC_WGROUP_BYTES;               // bytes of a slice

    uint32_t n_tile = 0, n_item = 0;  // running counters (tiles, items) of this CTA
    long long t_begin = 0;
    if ((dbg & 32) && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_begin));
    const int n_work = v.n_items * C::N_SLICES;
    for (int round = 0;; ++round, ++n_item) {
        // snake order over the sorted items: the SM that got the heaviest item of this round gets the
        // lightest of the next
        const int work = round * (int)gridDim.x + ((round & 1) ? (int)gridDim.x - 1 - (int)blockIdx.x : (int)blockIdx.x);
        if (round * (int)gridDim.x >= n_work) break;
        if (work >= n_work) {
            --n_item;  // nothing for this SM in the last, partial round
            continue;
        }
        const int item = v.item_order[work / C::N_SLICES], slice = work % C::N_SLICES;
        const int unit = v.item_unit[item];
        const int g0 = slice * C::NC;  // first Gaussian of the slice inside the unit
        const size_t gfirst = (size_t)unit * C::N_UNIT + g0;
        const int64_t lo = v.item_tile_lo[item], hi = v.item_tile_lo[item + 1];
        // the item's active tiles as a bit mask (an item has at most 64 tiles: two 32-bit masks); every warp forms the
        // same mask, so all roles skip the same tiles - and an item without any - consistently
        const uint32_t amask0 = __ballot_sync(0xffffffffu, lane < (int)(hi - lo) && __ldg(active + lo + lane) != 0);
        const uint32_t amask1 = __ballot_sync(0xffffffffu, 32 + lane < (int)(hi - lo) && __ldg(active + lo + 32 + lane) != 0);
        if ((amask0 | amask1) == 0u) {
            --n_item;  // compensates the loop increment: barrier parities count non-empty items only
            continue;
        }
        const int n_lo = __popc(amask0);
        const int n_tiles_ = n_lo + __popc(amask1);
        const int n_tiles = n_tiles_;
        // i-th active tile of the item -> tile index
        auto nth_tile = [&](int i) {
            return lo + (int64_t)(i < n_lo ? __fns(amask0, 0, i + 1) : 32 + __fns(amask1, 0, i - n_lo + 1));
        };

        if (warp == W_PROD) {
            // ------------------------------------------------------------ TMA producer
            tc::mbar_wait(&bars->b_empty, (n_item & 1) ^ 1);
            if (lane == 0) {
                tc::mbar_expect_tx(&bars->b_full, B_BYTES);
                tc::tma_load_1d(b_s, w16 + (size_t)unit * UNIT_IMG + (size_t)(g0 / 8) * PC_WGROUP_BYTES, B_BYTES,
                                &bars->b_full);
            }
            __syncwarp();
            for (int i = 0; i < n_tiles; ++i) {
                const uint32_t n = n_tile + i;
                const int slot = n % C::NA;
                const int64_t tile = nth_tile(i);
                tc::mbar_wait(&bars->a_empty[slot], ((n / C::NA) & 1) ^ 1);
                if (lane == 0) {
                    tc::mbar_expect_tx(&bars->a_full[slot], PC_XTILE_BYTES);
                    tc::tma_load_1d(a_s + slot * 2 * T_PIECE, x16 + (size_t)v.tile_xblk[tile] * PC_XTILE_BYTES,
                                    PC_XTILE_BYTES, &bars->a_full[slot]);
                }
                __syncwarp();
            }
        } else if (warp == W_MMA) {
            // ------------------------------------------------------------ MMA1 issuer: S = X * W^T
            constexpr uint32_t idesc1 = tc::umma_idesc_f16(T_ROWS, C::NPAD, 0, 0);
            constexpr uint32_t idesc1s = tc::umma_idesc_f16(T_ROWS, 2 * C::NPAD, 0, 0);
            const uint32_t a_base = tc::smem_u32(a_s), b_base = tc::smem_u32(b_s);
            // loop bounds / ring counters through redux.sync: uniform registers for the descriptors
            const int n_tiles = __reduce_max_sync(0xffffffffu, n_tiles_);
            const uint32_t tmem_base = __reduce_max_sync(0xffffffffu, tmem_base_);
            const uint32_t n_tile_u = __reduce_max_sync(0xffffffffu, n_tile);
            tc::mbar_wait(&bars->b_full, n_item & 1);
            for (int i = 0; i < n_tiles; ++i) {
                const uint32_t n = n_tile_u + i;
                const int slot = n % C::NA, sb = n & 1;
                const bool rec = (dbg & 32) && blockIdx.x == 0 && lane == 0 && n < 1000;
                if (rec) g_acc_dbg[n * 8 + 0] = clock64();
                tc::mbar_wait(&bars->a_full[slot], (n / C::NA) & 1);
                if (rec) g_acc_dbg[n * 8 + 1] = clock64();
                tc::mbar_wait(&bars->s_empty[sb], ((n >> 1) & 1) ^ 1);
                if (rec) g_acc_dbg[n * 8 + 2] = clock64();
                tc::tc_fence_after();
                if (tc::elect_one()) {
                    const uint32_t d = tmem_base + sb * C::S_STRIDE;
                    const uint32_t ah = a_base + slot * 2 * T_PIECE, al = ah + T_PIECE;
                    const uint32_t bh = b_base, bl = b_base + PC_WGROUP_BYTES / 2;
                    uint32_t accum = 0;
                    if constexpr (C::STACK) {
#pragma unroll
                        for (int q = 0; q < 2; ++q) {
                            const uint32_t ap = q ? al : ah;
#pragma unroll
                            for (int k = 0; k < T_KCH / 2; ++k) {
                                const uint64_t ad = tc::umma_desc(ap + 2 * k * T_ROWS * 16, T_ROWS * 16, 128);
                                const uint64_t bd = tc::umma_desc(bh + 2 * k * 128, 128, PC_WGROUP_BYTES / 2);
                                tc::mma_f16_ss(d, ad, bd, idesc1s, accum);
                                accum = 1;
                            }
                        }
                    } else {
#pragma unroll
                        for (int q = 0; q < 3; ++q) {
                            const uint32_t ap = (q == 2) ? al : ah;
                            const uint32_t bp = (q == 1) ? bl : bh;
#pragma unroll
                            for (int k = 0; k < T_KCH / 2; ++k) {
                                const uint64_t ad = tc::umma_desc(ap + 2 * k * T_ROWS * 16, T_ROWS * 16, 128);
                                const uint64_t bd = tc::umma_desc(bp + 2 * k * 128, 128, PC_WGROUP_BYTES);
                                tc::mma_f16_ss(d, ad, bd, idesc1, accum);
                                accum = 1;
                            }
                        }
                    }
                    tc::tc_commit(&bars->s_full[sb]);
                    if (i == n_tiles - 1) tc::tc_commit(&bars->b_empty);  // last use of this item's B
                }
                __syncwarp();
            }
        } else if (warp == W_MMA2) {
            // ------------------------------------------------------------ MMA2 issuer: D2[f, g] += sum_t A[t, f] * P[t, g]
            constexpr uint32_t idesc2 = tc::umma_idesc_f16(128, C::NPAD, 1, 1);
            constexpr uint32_t idesc2s = tc::umma_idesc_f16(128, 2 * C::NPAD, 1, 1);
            const uint32_t a_base = tc::smem_u32(a_s), p_base = tc::smem_u32(p_s);
            const int n_tiles = __reduce_max_sync(0xffffffffu, n_tiles_);
            const uint32_t tmem_base = __reduce_max_sync(0xffffffffu, tmem_base_);
            const uint32_t n_tile_u = __reduce_max_sync(0xffffffffu, n_tile);
            const uint32_t d2 = tmem_base + C::D2_COL;
            for (int i = 0; i < n_tiles; ++i) {
                const uint32_t n = n_tile_u + i;
                const int slot = n % C::NA, ps = n & 1;
                const bool rec = (dbg & 32) && blockIdx.x == 0 && lane == 0 && n < 1000;
                if (rec) g_acc_dbg[n * 8 + 3] = clock64();
                tc::mbar_wait(&bars->p_full[ps], (n >> 1) & 1);
                tc::mbar_wait(&bars->a_full[slot], (n / C::NA) & 1);  // long complete; this thread's own view of the tile
                if (rec) g_acc_dbg[n * 8 + 4] = clock64();
                if (i == 0) tc::mbar_wait(&bars->d2_empty, (n_item & 1) ^ 1);  // previous flush done
                tc::tc_fence_after();
                // active 16-frame K-steps of this tile (2 bits per lane quarter)
                uint32_t kmask;
                {
                    const uint32_t w = *reinterpret_cast<const volatile uint32_t *>(&bars->kact[ps][0]);
                    kmask = (w & 3u) | (((w >> 8) & 3u) << 2) | (((w >> 16) & 3u) << 4) | (((w >> 24) & 3u) << 6);
                    kmask = __reduce_or_sync(0xffffffffu, kmask);
                }
                if (tc::elect_one()) {
                    const uint32_t ah = a_base + slot * 2 * T_PIECE, al = ah + T_PIECE;
                    const uint32_t ph = p_base + ps * 2 * C::P_PIECE, pl = ph + C::P_PIECE;
                    uint32_t accum = (i == 0) ? 0u : 1u;  // first tile of the item resets D2
                    if constexpr (C::STACK) {
#pragma unroll
                        for (int q = 0; q < 2; ++q) {
                            const uint32_t ap = q ? al : ah;
#pragma unroll
                            for (int k = 0; k < T_ROWS / 16; ++k) {
                                if (!((kmask >> k) & 1u)) continue;
                                const uint64_t ad = tc::umma_desc(ap + k * 256, 128, T_ROWS * 16);
                                const uint64_t bd = tc::umma_desc(ph + k * 256, 128, T_ROWS * 16);  // hi then lo blocks
                                tc::mma_f16_ss(d2, ad, bd, idesc2s, accum);
                                accum = 1;
                            }
                        }
                    } else {
#pragma unroll
                        for (int q = 0; q < 3; ++q) {
                            const uint32_t ap = (q == 1) ? al : ah;
                            const uint32_t pp = (q == 2) ? pl : ph;
#pragma unroll
                            for (int k = 0; k < T_ROWS / 16; ++k) {
                                if (!((kmask >> k) & 1u)) continue;
                                // MN-major views: 8 frames x 16 B core matrices; LBO = 128 B between
                                // 8-frame groups, SBO = 2048 B between 8-feature / 8-Gaussian blocks
                                const uint64_t ad = tc::umma_desc(ap + k * 256, 128, T_ROWS * 16);
                                const uint64_t bd = tc::umma_desc(pp + k * 256, 128, T_ROWS * 16);
                                tc::mma_f16_ss(d2, ad, bd, idesc2, accum);
                                accum = 1;
                            }
                        }
                    }
                    tc::tc_commit(&bars->a_empty[slot]);
                    tc::tc_commit(&bars->p_empty[ps]);
                    if (i == n_tiles - 1) tc::tc_commit(&bars->d2_full);
                }
                __syncwarp();
            }
        } else if (warp < W_FLUSH) {
            // ------------------------------------------------------------ softmax: 2 tile groups x
            // (2 column halves x 4 lane quarters)
            const int grp = warp >> 3, half = (warp >> 2) & 1;
            const int r = (warp & 3) * 32 + lane;
            const float *scale_g = wscale + gfirst;
            for (int i = 0; i < n_tiles; ++i) {
                const uint32_t n = n_tile + i;
                if ((int)(n & 1) != grp) continue;
                const int sb = n & 1, ps = n & 1;
                const int64_t tile = nth_tile(i);
                const int rows = v.tile_rows[tile];
                const int tp = v.tile_tp[tile];
                // d = (lgam - b) * log2(e) + 15 for the unit's three states; -inf kills the row
                float dl[PC_EMIT];
                bool row_act = false;
                {
                    const size_t o = (size_t)v.tile_boff[tile] + (size_t)r * tp;
#pragma unroll
                    for (int s = 0; s < PC_EMIT; ++s) {
                        float d = PC_NEG_INF;
                        if (r < rows) {
                            const float lg = __ldg(lgam + o + s), bb = __ldg(b + o + s);
                            d = (lg == PC_NEG_INF) ? PC_NEG_INF : fmaf(lg - bb, LOG2E, P_SHIFT);
                            row_act = row_act || lg > ACTIVE_MIN_LGAM;
                        }
                        dl[s] = d;
                    }
                }
                // 16-frame K-steps of MMA2 whose frames all lack posterior mass have an exactly zero P
                // block (same argument as for whole tiles): the MMA warp skips them, and a warp whose
                // 32 frames are all dead does not form its part of P at all
                const uint32_t ract = __ballot_sync(0xffffffffu, row_act);
                const bool rec = (dbg & 32) && blockIdx.x == 0 && lane == 0 && (warp & 7) == 0 && n < 1000;
                if (rec) g_acc_dbg[n * 8 + 5] = clock64();
                tc::mbar_wait(&bars->s_full[sb], (n >> 1) & 1);
                tc::tc_fence_after();
                tc::mbar_wait(&bars->p_empty[ps], ((n >> 1) & 1) ^ 1);
                if (rec) g_acc_dbg[n * 8 + 6] = clock64();
                const uint32_t taddr = tmem_base + sb * C::S_STRIDE + ((uint32_t)((warp & 3) * 32) << 16);
                uint8_t *ph = p_s + ps * 2 * C::P_PIECE, *pl = ph + C::P_PIECE;
#define PC_SOFTMAX(G0_, SC_)                                                              \
    do {                                                                                  \
        if (half == 0) softmax_blocks<MIX, G0_, SC_, 0>(taddr, dl, scale_g, ph, pl, r);   \
        else softmax_blocks<MIX, G0_, SC_, 1>(taddr, dl, scale_g, ph, pl, r);             \
    } while (0)
                if (ract == 0u) {
                } else if (C::N_SLICES == 1 || slice == 0) {
                    if (scaled_rows) PC_SOFTMAX(0, true);
                    else PC_SOFTMAX(0, false);
                } else {
                    if (scaled_rows) PC_SOFTMAX(C::NC, true);
                    else PC_SOFTMAX(C::NC, false);
                }
#undef PC_SOFTMAX
                if (half == 0 && lane == 0)
                    bars->kact[ps][warp & 3] = (uint8_t)(((ract & 0xffffu) ? 1u : 0u) | ((ract >> 16) ? 2u : 0u));
                tc::tc_fence_before();
                tc::fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    tc::mbar_arrive(&bars->s_empty[sb]);
                    tc::mbar_arrive(&bars->p_full[ps]);
                }
                if (rec) g_acc_dbg[n * 8 + 7] = clock64();
            }
        } else if (warp < W_PROD) {
            // ------------------------------------------------------------ flush D2 (lane = feature)
            const int q = warp - W_FLUSH;  // TMEM lane quarter == warp % 4
            const int f = q * 32 + lane;
            tc::mbar_wait(&bars->d2_full, n_item & 1);
            tc::tc_fence_after();
            const uint32_t taddr = tmem_base + C::D2_COL + ((uint32_t)(q * 32) << 16);
            double *dst = acc + gfirst * PC_KA + f;
#pragma unroll
            for (int j = 0; j < C::NPAD / 16; ++j) {
                float t16[16];
                tc::tmem_ld16(taddr + j * 16, t16);
                if constexpr (C::STACK) {  // + the columns fed by P_lo
                    float u16[16];
                    tc::tmem_ld16(taddr + C::NPAD + j * 16, u16);
                    tc::tmem_ld_wait();
#pragma unroll
                    for (int e = 0; e < 16; ++e) t16[e] += u16[e];
                } else {
                    tc::tmem_ld_wait();
                }
                if (f < PC_KA) {
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        const int g = j * 16 + e;
                        if (g < C::NC)
                            atomicAdd(dst + (size_t)g * PC_KA, (double)(t16[e] * P_UNSHIFT));
                    }
                }
            }
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&bars->d2_empty);
        }
        n_tile += n_tiles;
    }
    tc::tc_fence_before();
    __syncthreads();
    if ((dbg & 32) && threadIdx.x == 0) {  // per block: wall nanoseconds, tiles and items processed
        long long t_end;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end));
        g_acc_dbg[4096 + blockIdx.x * 4 + 0] = t_begin;
        g_acc_dbg[4096 + blockIdx.x * 4 + 1] = t_end;
        g_acc_dbg[4096 + blockIdx.x * 4 + 2] = n_tile;
        g_acc_dbg[4096 + blockIdx.x * 4 + 3] = n_item;
    }
    if (warp == W_MMA) tc::tmem_dealloc(tmem_base, C::TM_COLS);
}

template <int MIX>
int launch_mix(pc_handle h, const CorpusView &v, const float *X, const float *W, const float *b,
               const float *lgam, double *acc, bool flags_fresh, cudaStream_t st) {
    auto kern = accumulate_tc_kernel<MIX>;
    PC_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<MIX>::SMEM));
    const int n_work = v.n_items * Cfg<MIX>::N_SLICES;
    const int grid = n_work < h->sm_count ? n_work : h->sm_count;
    if (flags_fresh) {
        item_order_kernel<<<(v.n_items + 31) / 32, 1024, 0, st>>>(v);
        PC_LAUNCH_CHECK();
        h->launches += 1;
    } else {
        const int64_t warps = v.n_xtiles * ACT_SPLIT;
        const int64_t blocks = (warps * 32 + ACT_THREADS - 1) / ACT_THREADS;
        // tile flags and the item counts behind them are one scratch range (the ticket counter cleans itself)
        PC_CUDA_TRY(cudaMemsetAsync(v.tile_active, 0, (size_t)((char *)(v.item_act + v.n_items) - (char *)v.tile_active), st));
        tile_active_kernel<<<(unsigned)blocks, ACT_THREADS, 0, st>>>(v, lgam, v.tile_active, h->debug_flags);
        PC_LAUNCH_CHECK();
        h->launches += 1;
    }
    if (h->debug_flags & 256) return PC_OK;  // timing experiment: pre-pass only
    kern<<<grid, NTHREADS, Cfg<MIX>::SMEM, st>>>(v, X, W, v.n_units * PC_EMIT * MIX, b, lgam, v.tile_active, acc,
                                                 h->debug_flags);
    PC_LAUNCH_CHECK();
    h->launches++;
    return PC_OK;
}

}  // namespace

// block 0's per-tile clocks (option debug_flags & 32): MMA warp [0] start, [1] tile landed, [2] S buffer
// free, [3] MMA1 issued / waiting for P, [4] P ready; softmax group [5] start, [6] S ready, [7] P written
extern "C" int pc_debug_read_acc(long long *host_out, int n) {
    return cudaMemcpyFromSymbol(host_out, g_acc_dbg, sizeof(long long) * n) == cudaSuccess ? 0 : -2;
}

bool accumulate_tc_supported(int mix) { return mix == 4 || mix == 8 || mix == 16 || mix == 32 || mix == 64; }

int launch_accumulate_tc(pc_handle h, const CorpusView &v, const float *X, const float *W, int mix,
                         const float *b, const float *lgam, double *acc, bool flags_fresh, cudaStream_t st) {
    if (v.n_items == 0) return PC_OK;
    switch (mix) {
        case 4: return launch_mix<4>(h, v, X, W, b, lgam, acc, flags_fresh, st);
        case 8: return launch_mix<8>(h, v, X, W, b, lgam, acc, flags_fresh, st);
        case 16: return launch_mix<16>(h, v, X, W, b, lgam, acc, flags_fresh, st);
        case 32: return launch_mix<32>(h, v, X, W, b, lgam, acc, flags_fresh, st);
        case 64: return launch_mix<64>(h, v, X, W, b, lgam, acc, flags_fresh, st);
    }
    pc_set_error("launch_accumulate_tc: mix=%d not covered", mix);
    return PC_ERR_UNSUPPORTED;
}
