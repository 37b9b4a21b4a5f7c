// K1 for WIDE units (64 mixtures: 192 Gaussians, one N = 192 accumulator per (frame tile, label position)).
//
// Same contraction and log-sum-exp as score_tc.cu; what changes is what stays in shared memory.  A unit
// image is 60 KiB (hi + lo), so the generic kernel holds two of them and two frame tiles, and streams
// every unit image of the utterance once per group of TWO tiles: measured at 100k utterances x 300 frames
// it is bound by the L2 -> SM fill (614 KB per 20-30 contractions), not by the tensor pipe.  Here all three
// tiles of a 300-frame utterance stay resident (120 KiB) and the unit images travel as hi and lo PIECES
// (30 KiB each, the de-interleaved image set of pack.cu): two slots for hi pieces, one for the lo piece.
// The three products of an accumulator are ordered so that the single lo slot turns around early:
//     first tile of a position :  Ah.Bh, Al.Bh, Ah.Bl      (the lo piece may still be in flight)
//     last tile of a position  :  Ah.Bl, Ah.Bh, Al.Bh      (the lo slot is released 10 MMAs before the end)
// so the next position's lo piece lands under ~20 MMAs, and the hi piece of the position after next under a
// whole position.  Per utterance the fill drops from 2 x 614 KB to 614 KB + 120 KB.
//
//   warp 12     TMA producer : frame tiles (40 KiB) and unit-image pieces in the MMA warp's order of use
//   warp 13     MMA issuer   : 15 tcgen05.mma (M = 128, N = 192, K = 16) per (tile, position)
//   warps 0-11  epilogue     : warp = (state, TMEM lane quarter); every warp takes every accumulator: tcgen05.ld of
//                              its state's 64 columns (one frame per thread), log-sum-exp, store
#include "tc_common.cuh"

#ifndef PC_K1B_POLY
#define PC_K1B_POLY 0  // exponentials on the FMA pipe: 0 none, 1 every fourth, 2 every second (measured per 12.5k utterances: 3.19 / 3.25 / 3.39 ms)
#endif

__device__ long long g_k1b_dbg[8192];

namespace {

using tc::T_KCH;
using tc::T_PIECE;
using tc::T_ROWS;
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;

constexpr int MIX = 64;
constexpr int N_REAL = PC_EMIT * MIX;          // 192 (already a multiple of 16)
constexpr int B_PIECE = T_KCH * N_REAL * 16;   // 30 720 B: hi or lo of one unit image
constexpr int SBO_PIECE = T_KCH * 128;         // 1 280 B between 8-row groups inside a piece
constexpr int G = 3;                           // resident frame tiles
constexpr int TM_STRIDE = 256, TM_BUFS = 2, TM_COLS = 512;
constexpr int EPI_WARPS = 4 * PC_EMIT;
constexpr int W_PROD = EPI_WARPS, W_MMA = W_PROD + 1;
constexpr int NTHREADS = (W_MMA + 1) * 32;
constexpr int SMEM = 1024 + G * 2 * T_PIECE + 3 * B_PIECE;
static_assert(SMEM <= 227 * 1024, "shared memory budget");

struct Bars {
    uint64_t a_full[G], a_empty[G];
    uint64_t h_full[2], h_empty[2];
    uint64_t l_full, l_empty;
    uint64_t tm_full[TM_BUFS], tm_empty[TM_BUFS];
    uint32_t tmem_base;
};

// log-sum-exp of the 64 component scores of one state; every fourth exponential runs on the FMA pipe
// (tc::ex2_fma): the MUFU unit is the epilogue's bottleneck (128 x 192 exponentials per accumulator)
template <bool SCALED>
__device__ __forceinline__ float state_lse64(uint32_t taddr, const float *__restrict__ scale) {
    float v[MIX];
#pragma unroll
    for (int jj = 0; jj < MIX / 16; ++jj) {
        float t16[16];
        tc::tmem_ld16(taddr + jj * 16, t16);
#pragma unroll
        for (int e = 0; e < 16; ++e) v[jj * 16 + e] = t16[e];
    }
    tc::tmem_ld_wait();
    if (SCALED) {
#pragma unroll
        for (int e = 0; e < MIX; ++e) v[e] *= __ldg(scale + e);
    }
    return tc::lse_packed<MIX, PC_K1B_POLY>(v);
}

__global__ void __launch_bounds__(NTHREADS, 1)
score_tc_big_kernel(CorpusView v, const float *__restrict__ X, const float *__restrict__ W, int n_gauss,
                    float *__restrict__ b, int item_lo, int item_hi, int dbg) {
    extern __shared__ __align__(1024) uint8_t smem[];
    Bars *bars = reinterpret_cast<Bars *>(smem);
    uint8_t *a_s = smem + 1024;
    uint8_t *h_s = a_s + G * 2 * T_PIECE;  // two hi pieces
    uint8_t *l_s = h_s + 2 * B_PIECE;      // one lo piece

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < G; ++i) { tc::mbar_init(&bars->a_full[i], 1); tc::mbar_init(&bars->a_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { tc::mbar_init(&bars->h_full[i], 1); tc::mbar_init(&bars->h_empty[i], 1); }
        tc::mbar_init(&bars->l_full, 1);
        tc::mbar_init(&bars->l_empty, 1);
        for (int i = 0; i < TM_BUFS; ++i) { tc::mbar_init(&bars->tm_full[i], 1); tc::mbar_init(&bars->tm_empty[i], EPI_WARPS); }
        tc::mbar_fence_init();
    }
    if (warp == W_MMA) tc::tmem_alloc(&bars->tmem_base, TM_COLS);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base_ = bars->tmem_base;

    const uint8_t *x16 = reinterpret_cast<const uint8_t *>(X) + pc_x16_offset(v.total_frames);
    const uint8_t *w16s = reinterpret_cast<const uint8_t *>(W) + pc_w16s_offset(n_gauss, N_REAL);
    const float *wscale = W + (size_t)n_gauss * PC_KA;
    const bool scaled_rows = reinterpret_cast<const int *>(wscale + n_gauss)[0] != 0;

    uint32_t n_a = 0, n_pos = 0, n_pair = 0;  // running counters: frame tiles, label positions, accumulators
    for (int item = item_lo + blockIdx.x; item < item_hi; item += gridDim.x) {
        const int u = v.sitem_utt[item];
        const int64_t f0 = v.frame_off[u];
        const int T = (int)(v.frame_off[u + 1] - f0);
        const int64_t p0 = v.pair_off[u];
        const int L_ = (int)(v.pair_off[u + 1] - p0);
        const int t_item = v.sitem_t0[item];
        const int nt_item = v.sitem_nt[item];
        for (int jg = 0; jg < nt_item; jg += G) {  // groups of G tiles
            const int nt_ = min(G, nt_item - jg);
            const int t_first = t_item + jg * T_ROWS;
            if (warp == W_PROD) {
                // -------------------------------------------------------- TMA producer
                const int L = L_, nt = nt_;
                auto load_a = [&](int t0) {
                    const int slot = n_a % G;
                    tc::mbar_wait(&bars->a_empty[slot], ((n_a / G) & 1) ^ 1);
                    if (lane == 0) {
                        tc::mbar_expect_tx(&bars->a_full[slot], PC_XTILE_BYTES);
                        tc::tma_load_1d(a_s + slot * 2 * T_PIECE,
                                        x16 + (size_t)(v.xtile_off[u] + t0 / T_ROWS) * PC_XTILE_BYTES, PC_XTILE_BYTES,
                                        &bars->a_full[slot]);
                    }
                    ++n_a;
                    __syncwarp();
                };
                auto load_b = [&](int p) {  // hi then lo piece of label position p
                    const uint32_t np = n_pos + p;
                    const uint8_t *img = w16s + (size_t)v.labels[p0 + p] * 2 * B_PIECE;
                    const int hs = np & 1;
                    tc::mbar_wait(&bars->h_empty[hs], ((np >> 1) & 1) ^ 1);
                    if (lane == 0) {
                        tc::mbar_expect_tx(&bars->h_full[hs], B_PIECE);
                        tc::tma_load_1d(h_s + hs * B_PIECE, img, B_PIECE, &bars->h_full[hs]);
                    }
                    __syncwarp();
                    tc::mbar_wait(&bars->l_empty, (np & 1) ^ 1);
                    if (lane == 0) {
                        tc::mbar_expect_tx(&bars->l_full, B_PIECE);
                        tc::tma_load_1d(l_s, img + B_PIECE, B_PIECE, &bars->l_full);
                    }
                    __syncwarp();
                };
                // the MMA warp's order of first use: A_0, B_0 (hi, lo), A_1 .. A_{nt-1}, B_1 .. B_{L-1}
                load_a(t_first);
                load_b(0);
                for (int j = 1; j < nt; ++j) load_a(t_first + j * T_ROWS);
                for (int p = 1; p < L; ++p) load_b(p);
                n_pos += L;
            } else if (warp == W_MMA) {
                // -------------------------------------------------------- MMA issuer
                constexpr uint32_t idesc = tc::umma_idesc_f16(T_ROWS, N_REAL, 0, 0);
                const uint32_t a_base = tc::smem_u32(a_s), h_base = tc::smem_u32(h_s), l_base = tc::smem_u32(l_s);
                const int L = __reduce_max_sync(0xffffffffu, L_);
                const int nt = __reduce_max_sync(0xffffffffu, nt_);
                const uint32_t tmem_base = __reduce_max_sync(0xffffffffu, tmem_base_);
                n_a = __reduce_max_sync(0xffffffffu, n_a);
                n_pos = __reduce_max_sync(0xffffffffu, n_pos);
                n_pair = __reduce_max_sync(0xffffffffu, n_pair);
                const uint32_t na0 = n_a;
                for (int p = 0; p < L; ++p, ++n_pos) {
                    const int hs = n_pos & 1;
                    if ((dbg & 32) && blockIdx.x == 0 && lane == 0 && n_pair < 600) g_k1b_dbg[n_pair * 8 + 7] = clock64();
                    tc::mbar_wait(&bars->h_full[hs], (n_pos >> 1) & 1);
                    for (int j = 0; j < nt; ++j, ++n_pair) {
                        const uint32_t na = na0 + j;
                        const int slot = na % G, tb = n_pair % TM_BUFS;
                        const bool rec = (dbg & 32) && blockIdx.x == 0 && lane == 0 && n_pair < 600;
                        if (rec) g_k1b_dbg[n_pair * 8 + 0] = clock64();
                        if (p == 0) tc::mbar_wait(&bars->a_full[slot], (na / G) & 1);
                        if (rec) g_k1b_dbg[n_pair * 8 + 1] = clock64();
                        tc::mbar_wait(&bars->tm_empty[tb], ((n_pair / TM_BUFS) & 1) ^ 1);
                        if (rec) g_k1b_dbg[n_pair * 8 + 2] = clock64();
                        const bool lo_first = j == nt - 1 && nt > 1;
                        // the lo piece: the position's first tile waits for it below, before its third product; the
                        // later tiles of the position come after that wait
                        if (rec) g_k1b_dbg[n_pair * 8 + 3] = clock64();
                        tc::tc_fence_after();
                        const uint32_t d = tmem_base + tb * TM_STRIDE;
                        const uint32_t ah = a_base + slot * 2 * T_PIECE, al = ah + T_PIECE;
                        const uint32_t bh = h_base + hs * B_PIECE;
                        auto product = [&](uint32_t ap, uint32_t bp, uint32_t first) {
                            uint32_t accum = first ? 0u : 1u;
#pragma unroll
                            for (int k = 0; k < T_KCH / 2; ++k) {
                                const uint64_t ad = tc::umma_desc(ap + 2 * k * T_ROWS * 16, T_ROWS * 16, 128);
                                const uint64_t bd = tc::umma_desc(bp + 2 * k * 128, 128, SBO_PIECE);
                                tc::mma_f16_ss(d, ad, bd, idesc, accum);
                                accum = 1;
                            }
                        };
                        if (lo_first) {
                            if (tc::elect_one()) {
                                product(ah, l_base, 1);
                                tc::tc_commit(&bars->l_empty);  // fires once every MMA that read the lo piece has run
                                product(ah, bh, 0);
                                product(al, bh, 0);
                            }
                        } else {
                            if (tc::elect_one()) {
                                product(ah, bh, 1);
                                product(al, bh, 0);
                            }
                            __syncwarp();
                            if (j == 0) {
                                if (rec) g_k1b_dbg[n_pair * 8 + 4] = clock64();
                                tc::mbar_wait(&bars->l_full, n_pos & 1);
                                if (rec) g_k1b_dbg[n_pair * 8 + 5] = clock64();
                                tc::tc_fence_after();
                            }
                            if (tc::elect_one()) {
                                product(ah, l_base, 0);
                                if (nt == 1) tc::tc_commit(&bars->l_empty);
                            }
                        }
                        if (tc::elect_one()) {
                            tc::tc_commit(&bars->tm_full[tb]);
                            if (p == L - 1) tc::tc_commit(&bars->a_empty[slot]);
                            if (j == nt - 1) tc::tc_commit(&bars->h_empty[hs]);
                        }
                        if (rec) g_k1b_dbg[n_pair * 8 + 6] = clock64();
                        __syncwarp();
                    }
                }
                n_a = na0 + nt;
            } else {
                // -------------------------------------------------------- epilogue: warp = (state, lane quarter)
                const int L = L_, nt = nt_;
                const int quarter = warp & 3, st = warp >> 2;
                const int r = quarter * 32 + lane;  // row of the tile == TMEM lane
                const int sp = pc_spad(L);
                float *out_u = b + v.emis_off[u] + st;
                for (int p = 0; p < L; ++p) {
                    const float *scale_g = wscale + (size_t)v.labels[p0 + p] * N_REAL + st * MIX;
                    for (int j = 0; j < nt; ++j, ++n_pair) {
                        const int tb = n_pair % TM_BUFS;
                        const int t0 = t_first + j * T_ROWS;
                        const int rows = min(T_ROWS, T - t0);
                        const bool rec = (dbg & 32) && blockIdx.x == 0 && threadIdx.x == 0 && n_pair < 600;
                        if (rec) g_k1b_dbg[4800 + n_pair * 4 + 0] = clock64();
                        tc::mbar_wait(&bars->tm_full[tb], (n_pair / TM_BUFS) & 1);
                        if (rec) g_k1b_dbg[4800 + n_pair * 4 + 1] = clock64();
                        tc::tc_fence_after();
                        const uint32_t taddr = tmem_base_ + tb * TM_STRIDE + st * MIX + ((uint32_t)(quarter * 32) << 16);
                        const float res = scaled_rows ? state_lse64<true>(taddr, scale_g) : state_lse64<false>(taddr, scale_g);
                        tc::tc_fence_before();
                        __syncwarp();
                        if (lane == 0) tc::mbar_arrive(&bars->tm_empty[tb]);
                        if (rec) g_k1b_dbg[4800 + n_pair * 4 + 2] = clock64();
                        if (r < rows) out_u[(size_t)(t0 + r) * sp + PC_EMIT * p] = res;
                    }
                }
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == W_MMA) tc::tmem_dealloc(tmem_base_, TM_COLS);
}

}  // namespace

// block 0's clocks (debug_flags & 32).  MMA warp, per accumulator n [8n..]: start, tile landed, accumulator free, lo piece
// (when needed first), before / after the late lo wait of a position's first tile, issued; [8n+7] on a position's first
// accumulator: before the hi-piece wait.  Epilogue warp 0 [4800 + 4n..]: start, accumulator ready, done.
extern "C" int pc_debug_read_k1b(long long *host_out, int n) {
    return cudaMemcpyFromSymbol(host_out, g_k1b_dbg, sizeof(long long) * n) == cudaSuccess ? 0 : -2;
}

bool score_tc_big_supported(int mix) { return mix == MIX; }

int launch_score_tc_big(pc_handle h, const CorpusView &v, const float *X, const float *W, int mix, float *b,
                        int item_lo, int item_hi, cudaStream_t st) {
    if (item_hi <= item_lo) return PC_OK;
    if (mix != MIX) {
        pc_set_error("launch_score_tc_big: mix=%d not covered", mix);
        return PC_ERR_UNSUPPORTED;
    }
    PC_CUDA_TRY(cudaFuncSetAttribute(score_tc_big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    const int n = item_hi - item_lo;
    const int grid = n < h->sm_count ? n : h->sm_count;
    score_tc_big_kernel<<<grid, NTHREADS, SMEM, st>>>(v, X, W, v.n_units * PC_EMIT * MIX, b, item_lo, item_hi,
                                                      h->debug_flags);
    PC_LAUNCH_CHECK();
    h->launches++;
    return PC_OK;
}
