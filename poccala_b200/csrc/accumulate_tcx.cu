// K3, second generation: Baum-Welch sufficient statistics with the Gaussians on the accumulator lanes.
//
// Same mathematics as accumulate_tc.cu / accumulate_simt.cu (LHMM.update_acc -> Clustering.GMM.update_acc,
// LHMM.py:473-507, Clustering.py:653-680, in the linear-equivalent form of SURVEY A.4):
//     acc[g] += sum_t gamma_t(j, m) * [x_t, 1 | x_t^2, 1],   gamma_t(j, m) = exp(c_t,g - b_t(j) + log gamma_t(j))
// but TRANSPOSED against the first generation, which put 128 frames of one (tile, unit) pair on the lanes:
//
//   S^T[g, t]   = <W_g, [x_t | x_t^2]>                MMA1: M = 128 Gaussians of the unit, N = 128 frames, K = 80
//   P[g, t]     = exp(S^T[g, t] - b_t + lgam_t)       one Gaussian per thread, written back to tensor memory
//   D2^T[g, f] += sum_t P[g, t] * [x_t | x_t^2][f]    MMA2: A = P from TENSOR MEMORY, N = 80 features, K = frames
//
// What that buys (profiles/README.md, round 2):
//   * the frame axis is an N / K dimension, so a batch is 128 frames GATHERED from the 32-frame blocks that
//     carry posterior mass (K2 leaves a 4-bit block mask per (tile, position)); the first generation
//     contracted whole 128-frame tiles of which about half the rows are dead, and re-fetched a 40 KB tile
//     image per (tile, unit) pair;
//   * the XU pipe (ex2 and the float -> half conversions of the hi / lo split) is the busiest unit of the softmax
//     warps (~3 000 clk per batch against 2 400 clk of MMAs).  Two ways around it were measured and dropped:
//     every second exponential on the FMA pipe (tc::ex2_fma) pushes the softmax warps, which hold 64
//     accumulator values each, into spills (183 us against 124 us); a bf16 lo part - a byte permute instead
//     of a conversion - is an illegal instruction (kind::f16 rejects A in bf16 against B in fp16).
//   * P never touches shared memory (tcgen05.st by the thread that owns the Gaussian's lane), the second
//     contraction reads its A operand from tensor memory and only 2.5 KB of B per MMA: no MN-major A operand
//     (1.25 clk per accumulator column in round 1);
//   * frames arrive as fp32 (tile images whose 32-row blocks are contiguous 5 KB pieces) and are split into fp16
//     (hi, lo) [x | x^2] in shared memory, as in score_tc_wide.cu.
// All contractions are the 3-product error-compensated sum with fp32 accumulation.  P is scaled by 2^15 so
// that the fp16 window covers [1.8e-12, 2].
//
//   warp 16     TMA producer : per work item the unit slice's Gaussian rows; per batch one 5 KB piece per block
//   warp 17     MMA issuer   : MMA1(n), then MMA2(n-1) (software pipelined), commits
//   warps 0-7   softmax      : quarter = warp % 4 (TMEM lanes), half = warp / 4 (64 of the 128 frames)
//   warps 8-11  prepare      : one batch ahead: fp32 staging -> fp16 (hi, lo) image, the batch's lgam - b rows
//   warps 12-15 flush        : D2^T (lane = Gaussian) -> registers -> atomicAdd(double), once per item
// (18 warps: with 22 the register file leaves 80 registers per thread and the softmax warps, which hold 64
// accumulator values each, spill - measured 160 us against 138 us at the bench shape)
#include "tc_common.cuh"

__device__ long long g_k3x_dbg[8192];

namespace {

using tc::T_KCH;
constexpr int NF = 128;                         // frames per batch
constexpr int NBLK = NF / PC_BLOCK_ROWS;        // blocks per batch
constexpr int X_PIECE = T_KCH * NF * 16;        // one fp16 piece (hi or lo) of a batch image
constexpr int STG_BYTES = NF * PC_XS * 4;       // fp32 staging of a batch: the tile image layout, [4 blocks][10 quads][32 rows][4]
constexpr int W_BYTES = 16 * PC_WGROUP_BYTES;   // 128 Gaussian rows of a unit image
constexpr int N_SOFT = 8, W_PREP = 8, N_PREP = 4, W_FLUSH = 12, W_PROD = 16, W_MMA = 17;
constexpr int NTHREADS = 18 * 32;
constexpr int TM_D1 = 0, TM_P = 128, TM_D2 = 384, TM_COLS = 512;  // D1: 128 cols, P: 2 x (64 hi + 64 lo), D2: 80
constexpr float LOG2E = 1.4426950408889634f;
constexpr float P_SHIFT = 15.f;
constexpr float P_UNSHIFT = 1.f / 32768.f;
constexpr int LIST_BYTES = 256 * (1 + N_PREP);  // block lists of the producer warp and the prepare warps
// Batch images are ring-buffered NXB deep.  With two, image n could only be formed after MMA2(n - 2), which
// the issue order puts behind MMA1(n - 1): prepare and the two contractions ran back to back (~5 000 clk per
// batch); with three they overlap.  Units of <= 96 Gaussians leave room for the third image: their two row
// buffers shrink to 96 rows each (the M = 128 contraction then reads 32 rows into whatever follows - lanes
// nobody looks at).
template <int MIX>
struct XCfg {
    static constexpr int N_UNIT = PC_EMIT * MIX;
    static constexpr int W_ROWS = ((N_UNIT + 15) & ~15) < 128 ? ((N_UNIT + 15) & ~15) : 128;  // rows a slice really holds
    static constexpr int W_BUF = W_ROWS / 8 * PC_WGROUP_BYTES;
    static constexpr int NXB = N_UNIT <= 96 ? 3 : 2;
    // fp32 stagings: the gathered 512-byte pieces come from HBM (2 000 - 3 000 clk from request to landing),
    // a third buffer lets the producer run two batches ahead of the conversion
    static constexpr int NSTG = (N_UNIT <= 48 || N_UNIT > 96) ? 3 : 2;
    static constexpr int SMEM = 1024 + 2 * W_BUF + NXB * 2 * X_PIECE + NSTG * STG_BYTES + NXB * PC_EMIT * NF * 4 + LIST_BYTES;
    static_assert(SMEM <= 227 * 1024, "shared memory budget");
};

struct XBars {
    uint64_t w_full[2], w_empty[2];
    uint64_t stg_full[3], stg_empty[3];
    uint64_t x_full[3], x_empty[3];
    uint64_t d1_full, d1_empty;
    uint64_t p_full[2], p_empty[2];
    uint64_t d2_full, d2_empty;
    uint32_t tmem_base;
};

// 32 lanes x 8 consecutive columns of tensor memory <- 8 registers per lane
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]),
                 "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// The activity masks of an item's tiles (<= 64 tiles, 4 bits each) and the walk over its active blocks.
// Every warp role builds the same structure, so all of them see the same sequence of batches.  A role that
// needs to know WHICH block is the n-th one passes 256 bytes of private shared memory for the list.
struct BlockWalk {
    int total;
    const uint8_t *list;  // n-th active block of the item: (tile inside the item << 2) | block
    __device__ __forceinline__ void init(const int32_t *__restrict__ active, int64_t lo, int64_t hi, int lane,
                                         uint8_t *list_smem) {
        const uint32_t m0 = (lo + lane < hi) ? (uint32_t)__ldg(active + lo + lane) & 0xFu : 0u;
        const uint32_t m1 = (lo + 32 + lane < hi) ? (uint32_t)__ldg(active + lo + 32 + lane) & 0xFu : 0u;
        const int c0 = __popc(m0), c1 = __popc(m1);
        int s0 = c0, s1 = c1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int a = __shfl_up_sync(0xffffffffu, s0, o), b2 = __shfl_up_sync(0xffffffffu, s1, o);
            if (lane >= o) { s0 += a; s1 += b2; }
        }
        const int t0 = __shfl_sync(0xffffffffu, s0, 31);
        total = t0 + __shfl_sync(0xffffffffu, s1, 31);
        list = list_smem;
        if (list_smem != nullptr) {
            __syncwarp();  // the previous item's list is no longer read
            int p = s0 - c0;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (m0 & (1u << k)) list_smem[p++] = (uint8_t)((lane << 2) | k);
            p = t0 + s1 - c1;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (m1 & (1u << k)) list_smem[p++] = (uint8_t)(((32 + lane) << 2) | k);
            __syncwarp();
        }
    }
    __device__ __forceinline__ void find(int n, int &tile, int &blk) const {
        const int e = list[n];
        tile = e >> 2;
        blk = e & 3;
    }
    // a per-tile quantity held per lane (a0 = tile lo + lane, a1 = tile lo + 32 + lane) -> its value for `tile`
    template <typename T>
    __device__ __forceinline__ T pick(T a0, T a1, int tile) const {
        const T x0 = __shfl_sync(0xffffffffu, a0, tile & 31), x1 = __shfl_sync(0xffffffffu, a1, tile & 31);
        return tile < 32 ? x0 : x1;
    }
};

// log gamma that did not come from K2: one warp per (tile of an utterance, 32-frame block), lane = state column
__global__ void __launch_bounds__(128)
block_active_kernel(CorpusView v, const float *__restrict__ lgam, int32_t *__restrict__ active) {
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int64_t xt = w / 4;
    const int part = (int)(w - xt * 4);
    if (xt >= v.n_xtiles) return;
    const int u = v.xtile_utt[xt], t0 = v.xtile_t0[xt];
    const int T = (int)(v.frame_off[u + 1] - v.frame_off[u]);
    const int64_t p0 = v.pair_off[u];
    const int L = (int)(v.pair_off[u + 1] - p0);
    const int sp = pc_spad(L), rows = min(PC_TILE_ROWS, T - t0);
    const int r_lo = part * PC_BLOCK_ROWS, r_hi = min(rows, r_lo + PC_BLOCK_ROWS);
    const float *base = lgam + v.emis_off[u] + (size_t)t0 * sp;
    for (int c = lane; c < PC_EMIT * L && r_lo < r_hi; c += 32) {
        float m = PC_NEG_INF;
        for (int r = r_lo; r < r_hi; ++r) m = fmaxf(m, __ldg(base + (size_t)r * sp + c));
        if (m > PC_ACTIVE_MIN_LGAM) atomicOr(active + v.pair_tile0[p0 + c / PC_EMIT] + t0 / PC_TILE_ROWS, 1 << part);
    }
}

// One warp per item counts its active blocks; the block that finishes last orders the items (heaviest first).
__global__ void __launch_bounds__(1024)
item_order_blocks_kernel(CorpusView v) {
    __shared__ int bin[257];
    __shared__ int is_last;
    const int lane = threadIdx.x & 31;
    const int i = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (i < v.n_items) {
        const int64_t lo = v.item_tile_lo[i], hi = v.item_tile_lo[i + 1];
        int c = 0;
        if (lo + lane < hi) c += __popc((uint32_t)__ldcg(v.tile_active + lo + lane) & 0xFu);
        if (lo + 32 + lane < hi) c += __popc((uint32_t)__ldcg(v.tile_active + lo + 32 + lane) & 0xFu);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        if (lane == 0) v.item_act[i] = c;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = atomicAdd(v.item_act + v.n_items, 1) == (int)gridDim.x - 1;
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    if (threadIdx.x == 0) v.item_act[v.n_items] = 0;  // the ticket counter is left clean for the next launch
    for (int k = threadIdx.x; k < 257; k += blockDim.x) bin[k] = 0;
    __syncthreads();
    for (int k = threadIdx.x; k < v.n_items; k += blockDim.x) atomicAdd(&bin[256 - min(__ldcg(v.item_act + k), 256)], 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        for (int k = 0; k < 257; ++k) { const int c = bin[k]; bin[k] = run; run += c; }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < v.n_items; k += blockDim.x)
        v.item_order[atomicAdd(&bin[256 - min(__ldcg(v.item_act + k), 256)], 1)] = k;
}

template <int MIX>
__global__ void __launch_bounds__(NTHREADS, 1)
accumulate_tcx_kernel(CorpusView v, const float *__restrict__ X, const float *__restrict__ W, int n_gauss,
                      const float *__restrict__ b, const float *__restrict__ lgam,
                      const int32_t *__restrict__ active, double *__restrict__ acc, int dbg) {
    constexpr int N_UNIT = PC_EMIT * MIX;                       // Gaussians of a unit
    constexpr int UNIT_ROWS = (N_UNIT + 15) & ~15;              // rows of a unit image
    constexpr int UNIT_IMG = UNIT_ROWS / 8 * PC_WGROUP_BYTES;
    constexpr int N_SLICES = (N_UNIT + 127) / 128;              // 128 Gaussians per work item slice
    constexpr int W_BUF = XCfg<MIX>::W_BUF, NXB = XCfg<MIX>::NXB, NSTG = XCfg<MIX>::NSTG;
    extern __shared__ __align__(1024) uint8_t smem[];
    XBars *bars = reinterpret_cast<XBars *>(smem);
    uint8_t *w_s = smem + 1024;                 // 2 slices of Gaussian rows (hi / lo interleaved per row group)
    uint8_t *x_s = w_s + 2 * W_BUF;             // NXB batch images: hi piece, lo piece
    uint8_t *stg_s = x_s + NXB * 2 * X_PIECE;   // NSTG fp32 stagings
    float *off_s = reinterpret_cast<float *>(stg_s + NSTG * STG_BYTES);  // NXB x [3][NF]: (lgam - b) * log2e + 15 per state
    uint8_t *list_s = reinterpret_cast<uint8_t *>(off_s + NXB * PC_EMIT * NF);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) {
            tc::mbar_init(&bars->w_full[i], 1); tc::mbar_init(&bars->w_empty[i], 1);
            tc::mbar_init(&bars->p_full[i], N_SOFT); tc::mbar_init(&bars->p_empty[i], 1);
        }
        for (int i = 0; i < 3; ++i) {
            tc::mbar_init(&bars->x_full[i], N_PREP); tc::mbar_init(&bars->x_empty[i], 1);
            tc::mbar_init(&bars->stg_full[i], 1); tc::mbar_init(&bars->stg_empty[i], N_PREP);
        }
        tc::mbar_init(&bars->d1_full, 1); tc::mbar_init(&bars->d1_empty, N_SOFT);
        tc::mbar_init(&bars->d2_full, 1); tc::mbar_init(&bars->d2_empty, 4);
        tc::mbar_fence_init();
    }
    // the Gaussian-row buffers are read 128 rows deep whatever the unit holds: no stale NaN patterns
    for (int i = threadIdx.x; i < 2 * W_BUF / 16; i += NTHREADS) reinterpret_cast<uint4 *>(w_s)[i] = make_uint4(0u, 0u, 0u, 0u);
    tc::fence_proxy_async();
    if (warp == W_MMA) tc::tmem_alloc(&bars->tmem_base, TM_COLS);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base_ = bars->tmem_base;

    const uint8_t *x32 = reinterpret_cast<const uint8_t *>(X) + pc_x32_offset(v.total_frames, v.n_xtiles);
    const uint8_t *w16 = reinterpret_cast<const uint8_t *>(W) + pc_w16_offset(n_gauss);
    const float *wscale = W + (size_t)n_gauss * PC_KA;
    const bool scaled_rows = reinterpret_cast<const int *>(wscale + n_gauss)[0] != 0;

    uint32_t n_batch = 0, n_item = 0;  // running counters of this CTA: batches, non-empty work items
    const int n_work = v.n_items * N_SLICES;
    // snake order over the sorted items: the SM that got the heaviest item of a round gets the lightest of the next.
    // The descriptor of the NEXT round's item is fetched while this round's item is processed (three dependent
    // look-ups, ~2 000 clk when done on demand).
    auto work_of = [&](int round) {
        return round * (int)gridDim.x + ((round & 1) ? (int)gridDim.x - 1 - (int)blockIdx.x : (int)blockIdx.x);
    };
    struct ItemDesc { int unit; int64_t lo, hi; };
    auto describe = [&](int round) {
        ItemDesc d = {0, 0, 0};
        const int work = work_of(round);
        if (round * (int)gridDim.x < n_work && work < n_work) {
            const int item = __ldg(v.item_order + work / N_SLICES);
            d.unit = __ldg(v.item_unit + item);
            d.lo = __ldg(v.item_tile_lo + item);
            d.hi = __ldg(v.item_tile_lo + item + 1);
        }
        return d;
    };
    ItemDesc next = describe(0);
    for (int round = 0;; ++round) {
        if (round * (int)gridDim.x >= n_work) break;
        const int work = work_of(round);
        const ItemDesc cur = next;
        next = describe(round + 1);
        if (work >= n_work) continue;
        const int slice = work % N_SLICES;
        const int unit = cur.unit;
        const int g0 = slice * 128;                       // first Gaussian of the slice inside the unit
        const int n_lanes = min(128, N_UNIT - g0);        // Gaussians of the slice
        const int64_t lo = cur.lo, hi = cur.hi;
        BlockWalk bw;
        bw.init(active, lo, hi, lane,
                warp == W_PROD ? list_s : (warp >= W_PREP && warp < W_FLUSH ? list_s + 256 * (1 + warp - W_PREP) : nullptr));
        if (bw.total == 0) continue;  // (every role skips the item: barrier parities count non-empty items)
        const int n_blocks = bw.total;
        const int n_bat = (n_blocks + NBLK - 1) / NBLK;
        const int wbuf = n_item & 1;

        if (warp == W_PROD) {
            // ------------------------------------------------------------ TMA producer
            const int w_rows = min(128, UNIT_ROWS - g0);
            // the image index of this lane's two tiles (no table look-ups inside the batch loop)
            const int64_t xb0 = (lo + lane < hi) ? __ldg(v.tile_xblk + lo + lane) : 0;
            const int64_t xb1 = (lo + 32 + lane < hi) ? __ldg(v.tile_xblk + lo + 32 + lane) : 0;
            tc::mbar_wait(&bars->w_empty[wbuf], ((n_item >> 1) & 1) ^ 1);
            if (lane == 0) {
                tc::mbar_expect_tx(&bars->w_full[wbuf], w_rows / 8 * PC_WGROUP_BYTES);
                tc::tma_load_1d(w_s + wbuf * W_BUF, w16 + (size_t)unit * UNIT_IMG + (size_t)(g0 / 8) * PC_WGROUP_BYTES,
                                w_rows / 8 * PC_WGROUP_BYTES, &bars->w_full[wbuf]);
            }
            __syncwarp();
            for (int k = 0; k < n_bat; ++k) {
                const uint32_t n = n_batch + k;
                const int sb = n % NSTG;
                const int nb = min(NBLK, n_blocks - k * NBLK);
                tc::mbar_wait(&bars->stg_empty[sb], ((n / NSTG) & 1) ^ 1);
                if (lane == 0) tc::mbar_expect_tx(&bars->stg_full[sb], nb * PC_BLOCK_ROWS * PC_XS * 4);
                __syncwarp();
                for (int s = 0; s < nb; ++s) {
                    int ti, blk;
                    bw.find(k * NBLK + s, ti, blk);
                    // a block is one contiguous 5 KB piece of its tile image, and of the staging (same layout)
                    const uint8_t *src = x32 + (size_t)bw.pick(xb0, xb1, ti) * PC_X32TILE_BYTES + blk * (PC_BLOCK_ROWS * PC_XS * 4);
                    uint8_t *dst = stg_s + sb * STG_BYTES + s * (PC_BLOCK_ROWS * PC_XS * 4);
                    if (lane == 0) tc::tma_load_1d(dst, src, PC_BLOCK_ROWS * PC_XS * 4, &bars->stg_full[sb]);
                }
                __syncwarp();
            }
        } else if (warp == W_MMA) {
            // ------------------------------------------------------------ MMA issuer
            constexpr uint32_t idesc1 = tc::umma_idesc_f16(128, NF, 0, 0);
            constexpr uint32_t idesc2 = tc::umma_idesc_f16(128, PC_KA, 0, 1);  // B = frames x features, features contiguous
            const uint32_t tmem_base = __reduce_max_sync(0xffffffffu, tmem_base_);
            const uint32_t nb0 = __reduce_max_sync(0xffffffffu, n_batch);
            const int nbat = __reduce_max_sync(0xffffffffu, n_bat);
            const int nblocks = __reduce_max_sync(0xffffffffu, n_blocks);
            const uint32_t w_base = tc::smem_u32(w_s) + wbuf * W_BUF, x_base = tc::smem_u32(x_s);
            tc::mbar_wait(&bars->w_full[wbuf], (n_item >> 1) & 1);
            // MMA2 of batch n (its P is complete): D2^T += P * X
            auto mma2 = [&](uint32_t n, int k_in_item) {
                const int pb = n & 1, xb = n % NXB;
                const int nb = min(NBLK, nblocks - k_in_item * NBLK);
                tc::mbar_wait(&bars->p_full[pb], (n >> 1) & 1);
                if (k_in_item == 0) tc::mbar_wait(&bars->d2_empty, (n_item & 1) ^ 1);  // the previous item's sums are in registers
                tc::tc_fence_after();
                if (tc::elect_one()) {
                    const uint32_t xh = x_base + xb * 2 * X_PIECE, xl = xh + X_PIECE;
                    const uint32_t ph = tmem_base + TM_P + pb * 128, pl = ph + 64;
                    uint32_t accum = k_in_item == 0 ? 0u : 1u;
                    if (!(dbg & 2)) {
                        for (int ks = 0; ks < 2 * nb; ++ks) {  // 16 frames per K-step, real blocks only
                            const uint64_t bh = tc::umma_desc(xh + ks * 256, 128, NF * 16);
                            const uint64_t bl = tc::umma_desc(xl + ks * 256, 128, NF * 16);
                            tc::mma_f16_ts(tmem_base + TM_D2, ph + 8 * ks, bh, idesc2, accum);
                            tc::mma_f16_ts(tmem_base + TM_D2, pl + 8 * ks, bh, idesc2, 1u);
                            tc::mma_f16_ts(tmem_base + TM_D2, ph + 8 * ks, bl, idesc2, 1u);
                            accum = 1u;
                        }
                    }
                    tc::tc_commit(&bars->p_empty[pb]);
                    tc::tc_commit(&bars->x_empty[xb]);
                    if (k_in_item == nbat - 1) tc::tc_commit(&bars->d2_full);
                }
                __syncwarp();
            };
            for (int k = 0; k < nbat; ++k) {
                const uint32_t n = nb0 + k;
                const int xb = n % NXB;
                const bool rec = (dbg & 32) && blockIdx.x == 0 && lane == 0 && n < 300;
                if (rec) g_k3x_dbg[n * 8 + 0] = clock64();
                tc::mbar_wait(&bars->x_full[xb], (n / NXB) & 1);
                if (rec) g_k3x_dbg[n * 8 + 1] = clock64();
                tc::mbar_wait(&bars->d1_empty, (n & 1) ^ 1);  // the softmax warps have read S of the previous batch
                if (rec) g_k3x_dbg[n * 8 + 2] = clock64();
                tc::tc_fence_after();
                if (tc::elect_one()) {
                    const uint32_t xh = x_base + xb * 2 * X_PIECE, xl = xh + X_PIECE;
                    const uint32_t wh = w_base, wl = w_base + PC_WGROUP_BYTES / 2;
                    uint32_t accum = 0;
                    if (!(dbg & 2)) {
#pragma unroll
                        for (int q = 0; q < 3; ++q) {
                            const uint32_t wp = (q == 1) ? wl : wh;
                            const uint32_t xp = (q == 2) ? xl : xh;
#pragma unroll
                            for (int kk = 0; kk < T_KCH / 2; ++kk) {
                                const uint64_t ad = tc::umma_desc(wp + 2 * kk * 128, 128, PC_WGROUP_BYTES);
                                const uint64_t bd = tc::umma_desc(xp + 2 * kk * NF * 16, NF * 16, 128);
                                tc::mma_f16_ss(tmem_base + TM_D1, ad, bd, idesc1, accum);
                                accum = 1;
                            }
                        }
                    }
                    tc::tc_commit(&bars->d1_full);
                    if (k == nbat - 1) tc::tc_commit(&bars->w_empty[wbuf]);  // last use of this item's Gaussian rows
                }
                if (rec) g_k3x_dbg[n * 8 + 3] = clock64();
                __syncwarp();
                if (k > 0) mma2(n - 1, k - 1);
                if (rec) g_k3x_dbg[n * 8 + 4] = clock64();
            }
            mma2(nb0 + nbat - 1, nbat - 1);
        } else if (warp < N_SOFT) {
            // ------------------------------------------------------------ softmax
            const int quarter = warp & 3, half = warp >> 2;
            const int g = quarter * 32 + lane;                    // Gaussian of the slice == TMEM lane
            const int st = min((g0 + g) / MIX, PC_EMIT - 1);      // its state (lanes past the slice: anything)
            const float mul = (scaled_rows && g < n_lanes) ? __ldg(wscale + (size_t)unit * N_UNIT + g0 + g) * LOG2E : LOG2E;
            const uint32_t tmem_base = tmem_base_;
            const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
            for (int k = 0; k < n_bat; ++k) {
                const uint32_t n = n_batch + k;
                const int sb = n & 1, xb = n % NXB;
                const float *off = off_s + xb * PC_EMIT * NF + st * NF + half * 64;
                const bool rec = (dbg & 32) && blockIdx.x == 0 && threadIdx.x == 0 && n < 300;
                if (rec) g_k3x_dbg[2400 + n * 8 + 0] = clock64();
                tc::mbar_wait(&bars->x_full[xb], (n / NXB) & 1);  // the batch's lgam - b rows are in off_s
                tc::mbar_wait(&bars->d1_full, n & 1);
                if (rec) g_k3x_dbg[2400 + n * 8 + 1] = clock64();
                tc::tc_fence_after();
                const uint32_t d1 = tmem_base + TM_D1 + half * 64 + lane_addr;
                const uint32_t ph = tmem_base + TM_P + sb * 128 + half * 32 + lane_addr, pl = ph + 64;
                float sv[4][16];
#pragma unroll
                for (int c = 0; c < 4; ++c) tc::tmem_ld16(d1 + 16 * c, sv[c]);
                tc::tmem_ld_wait();
                tc::tc_fence_before();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&bars->d1_empty);  // S is in registers: MMA1 of the next batch may run
                if (rec) g_k3x_dbg[2400 + n * 8 + 2] = clock64();
                tc::mbar_wait(&bars->p_empty[sb], ((n >> 1) & 1) ^ 1);  // MMA2 of batch n - 2 has read this P buffer
                if (rec) g_k3x_dbg[2400 + n * 8 + 3] = clock64();
                tc::tc_fence_after();
#pragma unroll
                for (int c = 0; c < 4; ++c) {  // 16 frames = one K-step of MMA2 per pass
                    uint32_t h[8], l[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const float d0 = off[16 * c + 2 * e], d1v = off[16 * c + 2 * e + 1];
                        float p0 = tc::ex2(fmaf(sv[c][2 * e], mul, d0)), p1 = tc::ex2(fmaf(sv[c][2 * e + 1], mul, d1v));
                        p0 = (d0 == PC_NEG_INF) ? 0.f : p0;  // frames without mass (and stale rows of a short batch)
                        p1 = (d1v == PC_NEG_INF) ? 0.f : p1;
                        if (dbg & 16) p0 = p1 = 0.f;
                        tc::split2(p0, p1, h[e], l[e]);
                    }
                    tmem_st8(ph + 8 * c, h);
                    tmem_st8(pl + 8 * c, l);
                }
                tmem_st_wait();
                tc::tc_fence_before();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&bars->p_full[sb]);
                if (rec) g_k3x_dbg[2400 + n * 8 + 4] = clock64();
            }
        } else if (warp < W_FLUSH) {
            // ------------------------------------------------------------ prepare: fp32 staging -> fp16 image, lgam - b rows
            const int tid = threadIdx.x - W_PREP * 32;  // frame of the batch
            constexpr int part = 0;
            // per-tile emission offsets, strides and frame counts of this lane's two tiles
            const bool h0 = lo + lane < hi, h1 = lo + 32 + lane < hi;
            const int64_t bo0 = h0 ? __ldg(v.tile_boff + lo + lane) : 0, bo1 = h1 ? __ldg(v.tile_boff + lo + 32 + lane) : 0;
            const int tp0 = h0 ? __ldg(v.tile_tp + lo + lane) : 0, tp1 = h1 ? __ldg(v.tile_tp + lo + 32 + lane) : 0;
            const int rw0 = h0 ? __ldg(v.tile_rows + lo + lane) : 0, rw1 = h1 ? __ldg(v.tile_rows + lo + 32 + lane) : 0;
            // lgam and b of frame `tid` of batch k, the unit's three states (issued one batch ahead of their use)
            float lgv[PC_EMIT], bbv[PC_EMIT];
            bool rowv = false;
            auto fetch = [&](int k) {
                const int nb = min(NBLK, n_blocks - k * NBLK);
                const int slot = tid / PC_BLOCK_ROWS, r = tid % PC_BLOCK_ROWS;
                int64_t boff = 0;
                int tp = 0, rows = 0;
                for (int s = 0; s < NBLK; ++s) {
                    int ti = 0, blk = 0;
                    if (s < nb) bw.find(k * NBLK + s, ti, blk);
                    const int64_t tb = bw.pick(bo0, bo1, ti);
                    const int tt = bw.pick(tp0, tp1, ti), tr = bw.pick(rw0, rw1, ti);
                    if (s == slot && s < nb) {
                        tp = tt;
                        boff = tb + (int64_t)blk * PC_BLOCK_ROWS * tt;
                        rows = min(PC_BLOCK_ROWS, tr - blk * PC_BLOCK_ROWS);
                    }
                }
                rowv = r < rows && part == 0;
#pragma unroll
                for (int s = 0; s < PC_EMIT; ++s) {
                    lgv[s] = rowv ? __ldg(lgam + boff + (int64_t)r * tp + s) : PC_NEG_INF;
                    bbv[s] = rowv ? __ldg(b + boff + (int64_t)r * tp + s) : 0.f;
                }
            };
            fetch(0);
            for (int k = 0; k < n_bat; ++k) {
                const uint32_t n = n_batch + k;
                const int sb = n % NSTG, xb = n % NXB;
                float *off = off_s + xb * PC_EMIT * NF;
                const bool rec = (dbg & 32) && blockIdx.x == 0 && tid == 0 && n < 300;
                if (rec) g_k3x_dbg[4800 + n * 8 + 0] = clock64();
                float dv[PC_EMIT];
#pragma unroll
                for (int s = 0; s < PC_EMIT; ++s)
                    dv[s] = (rowv && lgv[s] > PC_ACTIVE_MIN_LGAM - 1.f) ? fmaf(lgv[s] - bbv[s], LOG2E, P_SHIFT) : PC_NEG_INF;
                if (rec) g_k3x_dbg[4800 + n * 8 + 1] = clock64() + (long long)(dv[0] == 123.f);
                tc::mbar_wait(&bars->x_empty[xb], ((n / NXB) & 1) ^ 1);  // MMA2 of batch n - NXB has read the image (and off_s)
                if (rec) g_k3x_dbg[4800 + n * 8 + 2] = clock64();
                if (part == 0) {
#pragma unroll
                    for (int s = 0; s < PC_EMIT; ++s) off[s * NF + tid] = dv[s];
                }
                if (k + 1 < n_bat) fetch(k + 1);  // in flight while this batch is converted
                tc::mbar_wait(&bars->stg_full[sb], (n / NSTG) & 1);
                if (rec) g_k3x_dbg[4800 + n * 8 + 3] = clock64();
                if (!(dbg & 8)) {
#pragma unroll
                    for (int c = 0; c < 5; ++c)
                        tc::convert_tile_chunk_qm(stg_s + sb * STG_BYTES, tid, c, x_s + xb * 2 * X_PIECE,
                                                  x_s + xb * 2 * X_PIECE + X_PIECE);
                }
                tc::fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    tc::mbar_arrive(&bars->x_full[xb]);
                    tc::mbar_arrive(&bars->stg_empty[sb]);
                }
                if (rec) g_k3x_dbg[4800 + n * 8 + 4] = clock64();
            }
        } else if (warp < W_PROD) {
            // ------------------------------------------------------------ flush D2^T (lane = Gaussian)
            const int q = warp - W_FLUSH;  // TMEM lane quarter == warp % 4
            const int g = q * 32 + lane;
            tc::mbar_wait(&bars->d2_full, n_item & 1);
            tc::tc_fence_after();
            const uint32_t taddr = tmem_base_ + TM_D2 + ((uint32_t)(q * 32) << 16);
            float d[PC_KA];
#pragma unroll
            for (int j = 0; j < PC_KA / 16; ++j) {
                float t16[16];
                tc::tmem_ld16(taddr + j * 16, t16);
#pragma unroll
                for (int e = 0; e < 16; ++e) d[j * 16 + e] = t16[e];
            }
            tc::tmem_ld_wait();
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&bars->d2_empty);  // the sums are in registers: the next item may start
            if (g < n_lanes) {
                double *dst = acc + ((size_t)unit * N_UNIT + g0 + g) * PC_KA;
#pragma unroll
                for (int f = 0; f < PC_KA; ++f) atomicAdd(dst + f, (double)(d[f] * P_UNSHIFT));
            }
        }
        n_batch += n_bat;
        ++n_item;
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == W_MMA) tc::tmem_dealloc(tmem_base_, TM_COLS);
}

template <int MIX>
int launch_x(pc_handle h, const CorpusView &v, const float *X, const float *W, const float *b, const float *lgam,
             double *acc, bool flags_fresh, cudaStream_t st) {
    auto kern = accumulate_tcx_kernel<MIX>;
    constexpr int SMEM = XCfg<MIX>::SMEM;
    PC_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    constexpr int N_SLICES = (PC_EMIT * MIX + 127) / 128;
    const int n_work = v.n_items * N_SLICES;
    const int grid = n_work < h->sm_count ? n_work : h->sm_count;
    if (!flags_fresh) {
        const int64_t warps = v.n_xtiles * 4;
        const int64_t blocks = (warps * 32 + 127) / 128;
        // tile masks and the item counts behind them are one scratch range (the ticket counter cleans itself)
        PC_CUDA_TRY(cudaMemsetAsync(v.tile_active, 0, (size_t)((char *)(v.item_act + v.n_items) - (char *)v.tile_active), st));
        block_active_kernel<<<(unsigned)blocks, 128, 0, st>>>(v, lgam, v.tile_active);
        PC_LAUNCH_CHECK();
        h->launches += 1;
    }
    item_order_blocks_kernel<<<(v.n_items + 31) / 32, 1024, 0, st>>>(v);
    PC_LAUNCH_CHECK();
    h->launches += 1;
    kern<<<grid, NTHREADS, SMEM, st>>>(v, X, W, v.n_units * PC_EMIT * MIX, b, lgam, v.tile_active, acc, h->debug_flags);
    PC_LAUNCH_CHECK();
    h->launches++;
    return PC_OK;
}

}  // namespace

// block 0's clocks per batch n (debug_flags & 32): MMA warp [8n..]: start, image ready, S buffer free, MMA1 issued,
// MMA2(n-1) issued; softmax warp 0 [2400 + 8n..]: start, S ready, S read, P buffer free, P written; prepare warp
// [4800 + 8n..]: start, lgam / b rows loaded, image buffer free, staging landed, converted
extern "C" int pc_debug_read_k3x(long long *host_out, int n) {
    return cudaMemcpyFromSymbol(host_out, g_k3x_dbg, sizeof(long long) * n) == cudaSuccess ? 0 : -2;
}

bool accumulate_tcx_supported(int mix) { return mix == 4 || mix == 8 || mix == 16 || mix == 32 || mix == 64; }

int launch_accumulate_tcx(pc_handle h, const CorpusView &v, const float *X, const float *W, int mix, const float *b,
                          const float *lgam, double *acc, bool flags_fresh, cudaStream_t st) {
    if (v.n_items == 0) return PC_OK;
    switch (mix) {
        case 4: return launch_x<4>(h, v, X, W, b, lgam, acc, flags_fresh, st);
        case 8: return launch_x<8>(h, v, X, W, b, lgam, acc, flags_fresh, st);
        case 16: return launch_x<16>(h, v, X, W, b, lgam, acc, flags_fresh, st);
        case 32: return launch_x<32>(h, v, X, W, b, lgam, acc, flags_fresh, st);
        case 64: return launch_x<64>(h, v, X, W, b, lgam, acc, flags_fresh, st);
    }
    pc_set_error("launch_accumulate_tcx: mix=%d not covered", mix);
    return PC_ERR_UNSUPPORTED;
}
