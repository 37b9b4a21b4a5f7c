// K1 (tensor-core generation): GMM scoring as a TMA-fed tcgen05 contraction with a fused
// log-sum-exp.
//
// Same arithmetic as score_simt.cu (LHMM.cal_observation_pro -> GMM.point -> gaussian_function,
// LHMM.py:163-187, Clustering.py:740-767, util.py:20-36,54-77):
//     c[t, g] = <[x_t (39), 1 | x_t^2 (39), 1], W_g>,     b[t, s] = logsumexp_{g in s} c[t, g]
// Both operands live in HBM as fp16 (hi, lo) pairs (pack.cu) and the contraction is the
// error-compensated sum  A_hi*B_hi + A_hi*B_lo + A_lo*B_hi  with fp32 accumulation in TMEM, which
// reproduces the fp32 result (one fp16/bf16/tf32 pass does not meet the 1e-4 parity bound).
//
// Work decomposition is UTTERANCE-major: a work item is a group of <= G consecutive 128-frame
// tiles of one utterance; the tiles stay in shared memory while the Gaussians of every label
// position of the utterance stream through a ring of B stages.  Per SM this needs
// (G*40 KB + L*|B|) of shared-memory fill per G*L tile contractions, which keeps the kernel on the
// tensor pipe instead of the L2->SM fill rate.
//
//   warp 16     TMA producer : one cp.async.bulk per frame tile (40 KiB) and one per position (the
//                              unit's operand image); operands land directly in the UMMA layout
//   warp 17     MMA issuer   : 15 tcgen05.mma (M=128, N=NPAD, K=16) per (tile, position), commits
//   warps 0-15  epilogue     : one warpgroup per TMEM buffer, alternating (tile, position) pairs: tcgen05.ld
//                              (one frame per thread), log-sum-exp per state, coalesced store of b
// Shared-memory operand layouts (no swizzle, 8 rows x 16 B core matrices): frame tiles keep 16-byte
// chunk c of row r at c*2048 + r*16 (SBO = 128 B between 8-row groups, LBO = 2048 B between K chunks);
// Gaussian images are row-group major (LBO = 128 B, SBO = 2560 B), see pack.cu.
#include "tc_common.cuh"

__device__ long long g_pc_dbg[8192];

namespace {

using tc::T_KCH;
using tc::T_PIECE;
using tc::T_ROWS;
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;

template <int MIX>
struct Cfg {
    static constexpr int N_REAL = PC_EMIT * MIX;
    static constexpr int NPAD = (N_REAL + 15) & ~15;
    static constexpr int B_PIECE = T_KCH * NPAD * 16;
    static constexpr int B_STAGE = 2 * B_PIECE;  // hi, lo
    static constexpr int G = MIX <= 32 ? 3 : 2;                                // tiles per group
    // frame-tile slots: the group's tiles stay in shared memory for all label positions; one more
    // slot lets the producer start on the next group
    static constexpr int NA = MIX <= 32 ? 4 : 2;
    // PG label positions share one accumulator (one MMA of N = PG * NPAD columns, one hand-off to the
    // epilogue): with the frame tile in tensor memory an N = 48 MMA still costs ~32 clk of a 24 clk
    // floor and every accumulator costs a commit / wait round trip, so narrow units go in pairs
    static constexpr int PG = NPAD <= 64 ? 2 : 1;
    static constexpr int NB = 2;  // B slots (PG unit images each)
    static constexpr int N_ACC = PG * NPAD;
    static constexpr int TM_STRIDE = N_ACC <= 32 ? 32 : (N_ACC <= 64 ? 64 : (N_ACC <= 128 ? 128 : 256));
    static constexpr int TM_BUFS = TM_STRIDE >= 256 ? 2 : 4;
    // A_TMEM: frame tiles staged in tensor memory (tcgen05.cp once per tile, 40 columns hi + 40 lo)
    // as the A operand of every label position.  Measured (profiles/README.md): an N = 48 MMA drops
    // from ~70 to ~32 clk, but every tcgen05.cp of 128 rows x 32 B takes ~190 clk, i.e. ~5 600 clk
    // per group of three tiles; with two positions per MMA (N = 96) the shared-memory A read is
    // amortised just as well (~55 clk per MMA) without the staging, so the path is switched off.
    static constexpr bool A_TMEM = false;  // see profiles/README.md: staging cost ~190 clk per tcgen05.cp
    static constexpr int A_COL0 = TM_STRIDE * TM_BUFS;
    static constexpr int A_TILE_COLS = T_KCH * 8;  // 2 pieces x 5 K-steps x 8 columns
    static constexpr int TM_NEED = A_TMEM ? A_COL0 + G * A_TILE_COLS : TM_STRIDE * TM_BUFS;
    static_assert(TM_NEED <= 512, "tensor memory budget");
    static constexpr int TM_COLS = TM_NEED <= 32 ? 32 : (TM_NEED <= 64 ? 64 : (TM_NEED <= 128 ? 128 : (TM_NEED <= 256 ? 256 : 512)));
    // epilogue warps.  PG == 2: one warp per (lane quarter, position, state) - 24 warps that take every
    // accumulator, each thread reducing ONE state's mixtures (a thread that reduced a whole 48-column
    // block needed ~1 300 clk per block, more than the 800 clk the tensor pipe needs for the pair).
    // PG == 1 (wide units): one warpgroup per TMEM buffer, a thread reduces its frame's three states.
    static constexpr int EPI_WARPS = PG == 2 ? 4 * PG * PC_EMIT : 4 * TM_BUFS;
    static constexpr int W_PROD = EPI_WARPS, W_MMA = W_PROD + 1;
    static constexpr int NTHREADS = (W_MMA + 1) * 32;
    static constexpr int B_SLOT = PG * B_STAGE;
    static constexpr int SMEM = 1024 + NA * 2 * T_PIECE + NB * B_SLOT;
    static_assert(SMEM <= 227 * 1024, "shared memory budget");
};

struct Bars {
    uint64_t a_full[4], a_empty[4];
    uint64_t b_full[8], b_empty[8];
    uint64_t tm_full[4], tm_empty[4];
    uint32_t tmem_base;
};

// log-sum-exp of one state's MIX component scores held in TMEM columns [col0, col0 + MIX) of the
// calling thread's lane.  SCALED: multiply by the per-Gaussian power-of-two scale first.
template <int MIX, bool SCALED>
__device__ __forceinline__ float state_lse(const float (&v_in)[MIX], const float *__restrict__ scale) {
    float v[MIX];
#pragma unroll
    for (int e = 0; e < MIX; ++e) v[e] = SCALED ? v_in[e] * __ldg(scale + e) : v_in[e];
    return tc::lse_packed<MIX, 0>(v);
}

template <int MIX, bool SCALED>
__device__ __forceinline__ void epilogue_pair(uint32_t taddr, const float *__restrict__ scale,
                                              float (&res)[PC_EMIT]) {
    using C = Cfg<MIX>;
    if constexpr (MIX >= 16) {
#pragma unroll
        for (int s = 0; s < PC_EMIT; ++s) {
            float v[MIX];
#pragma unroll
            for (int jj = 0; jj < MIX / 16; ++jj) {
                float t16[16];
                tc::tmem_ld16(taddr + s * MIX + jj * 16, t16);
#pragma unroll
                for (int e = 0; e < 16; ++e) v[jj * 16 + e] = t16[e];
            }
            tc::tmem_ld_wait();
            res[s] = state_lse<MIX, SCALED>(v, scale + s * MIX);
        }
    } else {
        float all[C::NPAD];
#pragma unroll
        for (int jj = 0; jj < C::NPAD / 16; ++jj) {
            float t16[16];
            tc::tmem_ld16(taddr + jj * 16, t16);
#pragma unroll
            for (int e = 0; e < 16; ++e) all[jj * 16 + e] = t16[e];
        }
        tc::tmem_ld_wait();
#pragma unroll
        for (int s = 0; s < PC_EMIT; ++s) {
            float v[MIX];
#pragma unroll
            for (int e = 0; e < MIX; ++e) v[e] = all[s * MIX + e];
            res[s] = state_lse<MIX, SCALED>(v, scale + s * MIX);
        }
    }
}

template <int MIX>
__global__ void __launch_bounds__(Cfg<MIX>::NTHREADS, 1)
score_tc_kernel(CorpusView v, const float *__restrict__ X, const float *__restrict__ W, int n_gauss,
                float *__restrict__ b, int item_lo, int item_hi, int dbg) {
    using C = Cfg<MIX>;
    extern __shared__ __align__(1024) uint8_t smem[];
    Bars *bars = reinterpret_cast<Bars *>(smem);
    uint8_t *a_s = smem + 1024;
    uint8_t *b_s = a_s + C::NA * 2 * T_PIECE;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 4; ++i) {
            tc::mbar_init(&bars->a_full[i], 1); tc::mbar_init(&bars->a_empty[i], 1);
            tc::mbar_init(&bars->tm_full[i], 1); tc::mbar_init(&bars->tm_empty[i], C::PG == 2 ? C::EPI_WARPS : 4);
        }
        for (int i = 0; i < 8; ++i) { tc::mbar_init(&bars->b_full[i], 1); tc::mbar_init(&bars->b_empty[i], 1); }
        tc::mbar_fence_init();
    }
    if (warp == C::W_MMA) tc::tmem_alloc(&bars->tmem_base, C::TM_COLS);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base_ = bars->tmem_base;
    const uint32_t tmem_base = tmem_base_;

    const uint8_t *x16 = reinterpret_cast<const uint8_t *>(X) + pc_x16_offset(v.total_frames);
    const uint8_t *w16 = reinterpret_cast<const uint8_t *>(W) + pc_w16_offset(n_gauss);
    const float *wscale = W + (size_t)n_gauss * PC_KA;
    const bool scaled_rows = reinterpret_cast<const int *>(wscale + n_gauss)[0] != 0;

    uint32_t n_a = 0, n_b = 0, n_pair = 0;  // running counters: frame tiles, B stages, (tile, position) pairs
    for (int item = item_lo + blockIdx.x; item < item_hi; item += gridDim.x) {
        const int u = v.sitem_utt[item];
        const int64_t f0 = v.frame_off[u];
        const int T = (int)(v.frame_off[u + 1] - f0);
        const int64_t p0 = v.pair_off[u];
        const int L_ = (int)(v.pair_off[u + 1] - p0);
        const int L = L_;
        const int t_item = v.sitem_t0[item];
        const int nt_item = v.sitem_nt[item];
        for (int jg = 0; jg < nt_item; jg += C::G) {  // sub-groups of G tiles
            const int nt_ = min(C::G, nt_item - jg);
            const int nt = nt_;
            const int t_first = t_item + jg * T_ROWS;
            if (warp == C::W_PROD) {
                // -------------------------------------------------------- TMA producer
                auto load_a = [&](int uu, int t0) {   // one frame tile of utterance uu
                    const int slot = n_a % C::NA;
                    tc::mbar_wait(&bars->a_empty[slot], ((n_a / C::NA) & 1) ^ 1);
                    if (lane == 0 && (dbg & 8)) {
                        tc::mbar_arrive(&bars->a_full[slot]);
                    } else if (lane == 0) {
                        tc::mbar_expect_tx(&bars->a_full[slot], PC_XTILE_BYTES);
                        tc::tma_load_1d(a_s + slot * 2 * T_PIECE,
                                        x16 + (size_t)(v.xtile_off[uu] + t0 / T_ROWS) * PC_XTILE_BYTES,
                                        PC_XTILE_BYTES, &bars->a_full[slot]);
                    }
                    ++n_a;
                    __syncwarp();
                };
                auto load_b = [&](int pp) {           // the unit images of label positions pp*PG ..
                    const int slot = n_b % C::NB;
                    const int n_img = min(C::PG, L - pp * C::PG);
                    uint8_t *dst = b_s + slot * C::B_SLOT;
                    tc::mbar_wait(&bars->b_empty[slot], ((n_b / C::NB) & 1) ^ 1);
                    if (lane == 0 && (dbg & 4)) {
                        tc::mbar_arrive(&bars->b_full[slot]);
                    } else if (lane == 0) {
                        tc::mbar_expect_tx(&bars->b_full[slot], n_img * C::B_STAGE);
                        for (int i = 0; i < n_img; ++i)
                            tc::tma_load_1d(dst + i * C::B_STAGE, w16 + (size_t)v.labels[p0 + pp * C::PG + i] * C::B_STAGE,
                                            C::B_STAGE, &bars->b_full[slot]);
                    }
                    ++n_b;
                    __syncwarp();
                };
                const int NPG = (L + C::PG - 1) / C::PG;  // position groups of this utterance
                if constexpr (C::A_TMEM) {
                    // The MMA warp moves a group's tiles to tensor memory at once, which frees their
                    // slots: the NEXT group's tiles are requested right after this group's first unit
                    // images, a whole group of contractions before they are needed (HBM latency and
                    // the tcgen05.cp staging stay off the critical path).
                    if (item == item_lo + (int)blockIdx.x && jg == 0)
                        for (int j = 0; j < nt; ++j) load_a(u, t_first + j * T_ROWS);
                    const int lead = min(NPG, 2);  // B slots issued ahead of the next group's tiles
                    for (int p = 0; p < lead; ++p) load_b(p);
                    {
                        int nu = u, nt0 = t_first + C::G * T_ROWS, nn = nt_item - jg - C::G;  // next group
                        if (nn <= 0) {
                            const int nitem = item + (int)gridDim.x;
                            nn = 0;
                            if (nitem < item_hi) {
                                nu = v.sitem_utt[nitem];
                                nt0 = v.sitem_t0[nitem];
                                nn = v.sitem_nt[nitem];
                            }
                        }
                        nn = min(nn, C::G);
                        for (int j = 0; j < nn; ++j) load_a(nu, nt0 + j * T_ROWS);
                    }
                    for (int p = lead; p < NPG; ++p) load_b(p);
                } else {
                    // order = consumption order of the MMA warp: A_0, B_0, A_1 .. A_{nt-1}, B_1 .. B_{L-1}
                    load_a(u, t_first);
                    load_b(0);
                    for (int j = 1; j < nt; ++j) load_a(u, t_first + j * T_ROWS);
                    for (int p = 1; p < NPG; ++p) load_b(p);
                }
            } else if (warp == C::W_MMA) {
                // -------------------------------------------------------- MMA issuer
                // loop bounds and ring counters go through redux.sync so that the compiler can keep
                // them (and the UMMA descriptors derived from them) in uniform registers
                constexpr uint32_t idesc_full = tc::umma_idesc_f16(T_ROWS, C::N_ACC, 0, 0);
                constexpr uint32_t idesc_tail = tc::umma_idesc_f16(T_ROWS, C::NPAD, 0, 0);  // odd position left over
                const uint32_t a_base = tc::smem_u32(a_s), b_base = tc::smem_u32(b_s);
                const int L = __reduce_max_sync(0xffffffffu, L_);
                const int nt = __reduce_max_sync(0xffffffffu, nt_);
                const uint32_t tmem_base = __reduce_max_sync(0xffffffffu, tmem_base_);
                n_a = __reduce_max_sync(0xffffffffu, n_a);
                n_b = __reduce_max_sync(0xffffffffu, n_b);
                n_pair = __reduce_max_sync(0xffffffffu, n_pair);
                const uint32_t na0 = n_a;
                if constexpr (C::A_TMEM) {
                    // stage the group's frame tiles in tensor memory: 10 copies of 128 rows x 32 bytes
                    // per tile; the shared-memory slot is free again as soon as they have run, so the
                    // producer streams the next group's tiles under this group's contractions
                    for (int j = 0; j < nt; ++j) {
                        const uint32_t na = na0 + j;
                        const int slot = na % C::NA;
                        tc::mbar_wait(&bars->a_full[slot], (na / C::NA) & 1);
                        tc::tc_fence_after();
                        if (tc::elect_one()) {
                            const uint32_t ah = a_base + slot * 2 * T_PIECE, al = ah + T_PIECE;
                            const uint32_t ta = tmem_base + C::A_COL0 + j * C::A_TILE_COLS;
#pragma unroll
                            for (int k = 0; k < T_KCH / 2; ++k) {
                                tc::tmem_cp_128x256b(ta + 8 * k, tc::umma_desc(ah + 2 * k * T_ROWS * 16, T_ROWS * 16, 128));
                                tc::tmem_cp_128x256b(ta + T_KCH * 4 + 8 * k, tc::umma_desc(al + 2 * k * T_ROWS * 16, T_ROWS * 16, 128));
                            }
                            tc::tc_commit(&bars->a_empty[slot]);
                        }
                        __syncwarp();
                    }
                }
                const int NPG = (L + C::PG - 1) / C::PG;
                for (int p = 0; p < NPG; ++p, ++n_b) {
                    const int stage = n_b % C::NB;
                    const uint32_t idesc = (L - p * C::PG >= C::PG) ? idesc_full : idesc_tail;
                    tc::mbar_wait(&bars->b_full[stage], (n_b / C::NB) & 1);
                    for (int j = 0; j < nt; ++j, ++n_pair) {
                        const uint32_t na = na0 + j;
                        const int slot = na % C::NA, tb = n_pair % C::TM_BUFS;
                        const bool rec = (dbg & 32) && blockIdx.x == 0 && lane == 0 && n_pair < 1000;
                        if (rec) g_pc_dbg[n_pair * 8 + 0] = clock64();
                        if (!C::A_TMEM && p == 0) tc::mbar_wait(&bars->a_full[slot], (na / C::NA) & 1);
                        if (rec) g_pc_dbg[n_pair * 8 + 1] = clock64();
                        tc::mbar_wait(&bars->tm_empty[tb], ((n_pair / C::TM_BUFS) & 1) ^ 1);
                        if (rec) g_pc_dbg[n_pair * 8 + 2] = clock64();
                        tc::tc_fence_after();
                        if (tc::elect_one()) {
                            const uint32_t d = tmem_base + tb * C::TM_STRIDE;
                            const uint32_t ah = a_base + slot * 2 * T_PIECE, al = ah + T_PIECE;
                            const uint32_t bh = b_base + stage * C::B_SLOT, bl = bh + PC_WGROUP_BYTES / 2;
                            const uint32_t tah = tmem_base + C::A_COL0 + j * C::A_TILE_COLS, tal = tah + T_KCH * 4;
                            uint32_t accum = 0;
#pragma unroll
                            for (int q = 0; q < 3; ++q) {
                                const uint32_t ap = (q == 2) ? al : ah;
                                const uint32_t tap = (q == 2) ? tal : tah;
                                const uint32_t bp = (q == 1) ? bl : bh;
                                if (dbg & 2) break;
#pragma unroll
                                for (int k = 0; k < T_KCH / 2; ++k) {
                                    const uint64_t bd = tc::umma_desc(bp + 2 * k * 128, 128, PC_WGROUP_BYTES);
                                    if constexpr (C::A_TMEM) {
                                        tc::mma_f16_ts(d, tap + 8 * k, bd, idesc, accum);
                                    } else {
                                        const uint64_t ad = tc::umma_desc(ap + 2 * k * T_ROWS * 16, T_ROWS * 16, 128);
                                        tc::mma_f16_ss(d, ad, bd, idesc, accum);
                                    }
                                    accum = 1;
                                }
                            }
                            tc::tc_commit(&bars->tm_full[tb]);
                            if (!C::A_TMEM && p == NPG - 1) tc::tc_commit(&bars->a_empty[slot]);
                            if (j == nt - 1) tc::tc_commit(&bars->b_empty[stage]);
                        }
                        if (rec) g_pc_dbg[n_pair * 8 + 3] = clock64();
                        __syncwarp();
                    }
                }
                n_a = na0 + nt;
            } else if constexpr (C::PG == 2) {
                // -------------------------------------------------------- epilogue, narrow units
                // warp = (lane quarter, position inside the pair, state); every warp takes every
                // accumulator: tcgen05.ld of the state's MIX columns (one frame per thread), LSE, store
                const int quarter = warp & 3, role = warp >> 2;
                const int sub = role / PC_EMIT, st = role - sub * PC_EMIT;
                const int r = quarter * 32 + lane;  // row of the tile == TMEM lane
                const int sp = pc_spad(L);
                const int NPG = (L + C::PG - 1) / C::PG;
                for (int pp = 0; pp < NPG; ++pp) {
                    const int p = pp * C::PG + sub;
                    const bool have = p < L;  // an odd tail leaves the second position's warps idle
                    const float *scale_g = wscale + (size_t)v.labels[p0 + (have ? p : 0)] * C::N_REAL + st * MIX;
                    for (int j = 0; j < nt; ++j, ++n_pair) {
                        const int tb = n_pair % C::TM_BUFS;
                        const int t0 = t_first + j * T_ROWS;
                        const int rows = min(T_ROWS, T - t0);
                        tc::mbar_wait(&bars->tm_full[tb], (n_pair / C::TM_BUFS) & 1);
                        tc::tc_fence_after();
                        float res = 0.f;
                        if (have && !(dbg & 17)) {
                            const uint32_t taddr = tmem_base + tb * C::TM_STRIDE + sub * C::NPAD + st * MIX +
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
                            res = scaled_rows ? state_lse<MIX, true>(vv, scale_g) : state_lse<MIX, false>(vv, scale_g);
                        }
                        tc::tc_fence_before();
                        __syncwarp();
                        if (lane == 0) tc::mbar_arrive(&bars->tm_empty[tb]);
                        if (have && r < rows) b[v.emis_off[u] + (size_t)(t0 + r) * sp + PC_EMIT * p + st] = res;
                    }
                }
            } else {
                // -------------------------------------------------------- epilogue warpgroups, wide units
                const int grp = warp >> 2;
                const int r = (warp & 3) * 32 + lane;  // row of the tile == TMEM lane
                const int sp = pc_spad(L);
                for (int p = 0; p < L; ++p) {
                    const float *scale_g = wscale + (size_t)v.labels[p0 + p] * C::N_REAL;
                    for (int j = 0; j < nt; ++j, ++n_pair) {
                        if ((int)(n_pair % C::TM_BUFS) != grp) continue;
                        const int tb = n_pair % C::TM_BUFS;
                        const int t0 = t_first + j * T_ROWS;
                        const int rows = min(T_ROWS, T - t0);
                        float *out = b + v.emis_off[u] + (size_t)(t0 + r) * sp + PC_EMIT * p;
                        tc::mbar_wait(&bars->tm_full[tb], (n_pair / C::TM_BUFS) & 1);
                        tc::tc_fence_after();
                        const uint32_t taddr = tmem_base + tb * C::TM_STRIDE + ((uint32_t)((warp & 3) * 32) << 16);
                        float res[PC_EMIT] = {0.f, 0.f, 0.f};
                        if (!(dbg & 17)) {
                            if (scaled_rows) epilogue_pair<MIX, true>(taddr, scale_g, res);
                            else epilogue_pair<MIX, false>(taddr, scale_g, res);
                        }
                        tc::tc_fence_before();
                        __syncwarp();
                        if (lane == 0) tc::mbar_arrive(&bars->tm_empty[tb]);
                        if (r < rows) {
#pragma unroll
                            for (int s = 0; s < PC_EMIT; ++s) out[s] = res[s];
                        }
                    }
                }
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == C::W_MMA) tc::tmem_dealloc(tmem_base, C::TM_COLS);
}

template <int MIX>
int launch_mix(pc_handle h, const CorpusView &v, const float *X, const float *W, float *b, int item_lo,
               int item_hi, cudaStream_t st) {
    auto kern = score_tc_kernel<MIX>;
    PC_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<MIX>::SMEM));
    const int n = item_hi - item_lo;
    int grid = n < h->sm_count ? n : h->sm_count;
    kern<<<grid, Cfg<MIX>::NTHREADS, Cfg<MIX>::SMEM, st>>>(v, X, W, v.n_units * PC_EMIT * MIX, b, item_lo, item_hi,
                                                          h->debug_flags);
    PC_LAUNCH_CHECK();
    h->launches++;
    return PC_OK;
}

}  // namespace

extern "C" int pc_debug_read(long long *host_out, int n) {
    return cudaMemcpyFromSymbol(host_out, g_pc_dbg, sizeof(long long) * n) == cudaSuccess ? 0 : -2;
}

bool score_tc_supported(int mix) { return mix == 4 || mix == 8 || mix == 16 || mix == 32 || mix == 64; }

int launch_score_tc(pc_handle h, const CorpusView &v, const float *X, const float *W, int mix,
                    float *b, int item_lo, int item_hi, cudaStream_t st) {
    if (item_hi <= item_lo) return PC_OK;
    if (h->k1_kernel && score_tc_wide_supported(mix))  // narrow units: wide accumulators (score_tc_wide.cu)
        return launch_score_tc_wide(h, v, X, W, mix, b, item_lo, item_hi, st);
    if (h->k1_kernel && score_tc_big_supported(mix))  // 64 mixtures: resident tiles, unit images in pieces (score_tc_big.cu)
        return launch_score_tc_big(h, v, X, W, mix, b, item_lo, item_hi, st);
    switch (mix) {
        case 4: return launch_mix<4>(h, v, X, W, b, item_lo, item_hi, st);
        case 8: return launch_mix<8>(h, v, X, W, b, item_lo, item_hi, st);
        case 16: return launch_mix<16>(h, v, X, W, b, item_lo, item_hi, st);
        case 32: return launch_mix<32>(h, v, X, W, b, item_lo, item_hi, st);
        case 64: return launch_mix<64>(h, v, X, W, b, item_lo, item_hi, st);
    }
    pc_set_error("launch_score_tc: mix=%d not covered", mix);
    return PC_ERR_UNSUPPORTED;
}
