// K4: Viterbi forced alignment on the banded sentence HMM (one warp per utterance).
//
// Restates LHMM.viterbi (LHMM.py:546-609) as SURVEY A.6 verified bit-for-bit against the
// reference: fp64 scores p_t(j) = max(p_{t-1}(j-1) + log A[j-1,j], p_{t-1}(j) + log A[j,j]) + B[j,t]
// with the tie going to the LOWER predecessor index (np.where(tmp == max)[0][0]); end state = first
// argmax over all states; traceback through 1-bit backpointers kept in shared memory.  Only fp64
// adds and compares are used, so scores and paths are bit-identical to numpy given the same
// emissions, log-transitions and log-pi (those logs are computed on the host with numpy).
// A cell whose two candidates are both -inf gets backpointer 0 in the reference; such cells can
// never lie on the traced path of a finite end state, and for an all -inf column both
// implementations trace state 0, so one bit per cell is enough.
#include "common.cuh"

__device__ long long g_vit_dbg[8];

template <int SPL, typename E>
__global__ void viterbi_kernel(CorpusView v, const E *__restrict__ b,
                               const double *__restrict__ log_self,
                               const double *__restrict__ log_next,
                               const double *__restrict__ utt_logpi,
                               const double *__restrict__ state_logpi, int32_t *__restrict__ path,
                               int32_t *__restrict__ unit_path, double *__restrict__ score,
                               int words_per_warp) {
    extern __shared__ uint32_t bp_s[];
    const int lane = threadIdx.x & 31;
    const int warp_in_block = threadIdx.x >> 5;
    const int wid = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (wid >= v.n_utt) return;
    const int u = v.fb_order[wid];
    uint32_t *bp = bp_s + (size_t)warp_in_block * words_per_warp;
    const int64_t f0 = v.frame_off[u];
    const int T = (int)(v.frame_off[u + 1] - f0);
    const int64_t p0 = v.pair_off[u];
    const int L = (int)(v.pair_off[u + 1] - p0);
    const int NE = PC_EMIT * L;
    const int sp = pc_spad(L);
    const E *bu = b + v.emis_off[u];
    const double NINF = -INFINITY;

    double ls[SPL], ln[SPL], p[SPL];
    int kind[SPL];
    int64_t row[SPL];
#pragma unroll
    for (int q = 0; q < SPL; ++q) {
        const int s = lane * SPL + q;
        row[q] = 0;
        if (s == 0) {
            const int unit = v.labels[p0];
            kind[q] = 0;
            ls[q] = log_self[unit * PC_STATES];
            ln[q] = log_next[unit * PC_STATES];
        } else if (s <= NE) {
            const int pp = (s - 1) / PC_EMIT, r = (s - 1) - pp * PC_EMIT;
            const int unit = v.labels[p0 + pp];
            kind[q] = 1;
            row[q] = s - 1;
            ls[q] = log_self[unit * PC_STATES + 1 + r];
            ln[q] = log_next[unit * PC_STATES + 1 + r];  // into the exit state for s == NE: unused
        } else {
            kind[q] = 2;  // exit state (emission log 0) and padding: score stays -inf
            ls[q] = NINF;
            ln[q] = NINF;
        }
        double lpi = state_logpi ? ((s <= NE + 1) ? state_logpi[v.state_off[u] + s] : NINF)
                                 : utt_logpi[u];
        double e0 = kind[q] == 1 ? (double)bu[row[q]] : (kind[q] == 0 ? 0.0 : NINF);  // frame 0
        p[q] = (kind[q] == 2) ? NINF : lpi + e0;
    }
    const bool trace = blockIdx.x == 0 && threadIdx.x == 0;
    if (trace) g_vit_dbg[0] = clock64();
    uint32_t bits[SPL];
    double econst[SPL];  // emission of the states that do not read b: log 1 (entry), log 0 (exit, padding)
#pragma unroll
    for (int q = 0; q < SPL; ++q) {
        bits[q] = 0u;
        econst[q] = (kind[q] == 0) ? 0.0 : NINF;
    }
    // Frames go in chunks of VT_CH = 16 (8 with more than two states per lane: registers), aligned inside
    // the 32-bit backpointer words.  The
    // emission rows of a chunk are fetched one chunk ahead of the recurrence (they do not depend on it:
    // without the prefetch every step waited for its own global load, ~1 350 clk per frame).  Whole
    // chunks run unrolled without per-frame bounds checks or branches (the loop is issue-bound once the
    // SM holds its 28 utterances: ~100 instructions per frame before, ~35 now); the first chunk (frame 0
    // is the initialisation) and the last, partial one take the checked path.
    constexpr int VT_CH = (SPL <= 2) ? 16 : 8;
    auto load_rows = [&](int t0, E (&e)[VT_CH][SPL]) {
#pragma unroll
        for (int k = 0; k < VT_CH; ++k) {
            const int t = min(t0 + k, T - 1);
#pragma unroll
            for (int q = 0; q < SPL; ++q) e[k][q] = (kind[q] == 1) ? bu[(int64_t)t * sp + row[q]] : (E)0;
        }
    };
    // one frame: p_t(j) = max(p_{t-1}(j-1) + log A_{j-1,j}, p_{t-1}(j) + log A_jj) + B[j,t], tie -> j-1
    auto frame_step = [&](const E (&erow)[SPL], uint32_t bit) {
        double left = __shfl_up_sync(0xffffffffu, p[SPL - 1] + ln[SPL - 1], 1);
        if (lane == 0) left = NINF;
        double np_[SPL];
#pragma unroll
        for (int q = 0; q < SPL; ++q) {
            const double move = (q > 0) ? p[(q + SPL - 1) % SPL] + ln[(q + SPL - 1) % SPL] : left;
            const double stay = p[q] + ls[q];
            const bool take = move >= stay;  // tie -> lower index (j-1)
            const double best = take ? move : stay;
            const double e = kind[q] == 1 ? (double)erow[q] : econst[q];
            np_[q] = best + e;  // -inf for the exit state and the padding (their emission is log 0)
            bits[q] |= take ? bit : 0u;
        }
#pragma unroll
        for (int q = 0; q < SPL; ++q) p[q] = np_[q];
    };
    E e_nxt[VT_CH][SPL];
    load_rows(0, e_nxt);
    for (int t0 = 0; t0 < T; t0 += VT_CH) {
        E e_cur[VT_CH][SPL];
#pragma unroll
        for (int k = 0; k < VT_CH; ++k) {
#pragma unroll
            for (int q = 0; q < SPL; ++q) e_cur[k][q] = e_nxt[k][q];
        }
        load_rows(t0 + VT_CH, e_nxt);
        const uint32_t half = (uint32_t)(t0 & 31);  // bit position of the chunk's first frame in its word
        if (t0 > 0 && t0 + VT_CH <= T) {
#pragma unroll
            for (int k = 0; k < VT_CH; ++k) frame_step(e_cur[k], (1u << k) << half);
        } else {
#pragma unroll
            for (int k = 0; k < VT_CH; ++k) {
                const int t = t0 + k;
                if (t >= 1 && t < T) frame_step(e_cur[k], (1u << k) << half);
            }
        }
        if (((t0 + VT_CH) & 31) == 0 || t0 + VT_CH >= T) {  // the word of frames [32 * (t0 / 32), +32) is complete
#pragma unroll
            for (int q = 0; q < SPL; ++q) {
                bp[((t0 >> 5) * SPL + q) * 32 + lane] = bits[q];
                bits[q] = 0u;
            }
        }
    }
    if (trace) g_vit_dbg[1] = clock64();
    if (T == 1) {
#pragma unroll
        for (int q = 0; q < SPL; ++q) bp[q * 32 + lane] = 0u;
    }
    __syncwarp();
    // end state: first argmax over all states (LHMM.py:591)
    double best = NINF;
    int best_s = 0x7fffffff;
#pragma unroll
    for (int q = 0; q < SPL; ++q) {
        const int s = lane * SPL + q;
        if (p[q] > best) { best = p[q]; best_s = s; }
    }
    if (best == NINF) best_s = 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int os = __shfl_xor_sync(0xffffffffu, best_s, o);
        if (ob > best || (ob == best && os < best_s)) { best = ob; best_s = os; }
    }
    if (best_s == 0x7fffffff) best_s = 0;  // every score -inf: np.where(p == max)[0][0] == 0
    if (lane == 0) score[u] = best;
    // traceback, 32 frames at a time: every lane follows the same chain; the block's backpointer words
    // sit in registers (lane l holds the words of its own states) and the one for the current state
    // comes by shuffle.  The walk only records WHERE the path steps down (one bit per frame); lane k
    // then gets the state of frame 32*blk + k as the block's end state minus the steps taken after
    // frame k, so the stores are coalesced.  Whole blocks run without per-frame bounds checks.
    int cur = best_s;
    for (int blk = (T - 1) >> 5; blk >= 0; --blk) {
        uint32_t w[SPL];
#pragma unroll
        for (int q = 0; q < SPL; ++q) w[q] = bp[(blk * SPL + q) * 32 + lane];
        const int t_lo = blk * 32, t_hi = min(T - 1, t_lo + 31);
        const int end_state = cur;  // state of frame t_hi
        uint32_t steps = 0u;        // bit k: the path steps down between frame t_lo + k - 1 and t_lo + k
        auto back = [&](int k) {
            uint32_t wsel = w[0];
#pragma unroll
            for (int q = 1; q < SPL; ++q) wsel = ((cur % SPL) == q) ? w[q] : wsel;
            // (state 0 has no predecessor: its bit is set when every score of the column is -inf, move = stay = -inf)
            const uint32_t bit = cur > 0 ? (__shfl_sync(0xffffffffu, wsel, cur / SPL) >> k) & 1u : 0u;
            cur -= (int)bit;
            steps |= bit << k;
        };
        if (blk > 0 && t_hi == t_lo + 31) {
#pragma unroll
            for (int k = 31; k >= 0; --k) back(k);
        } else {
#pragma unroll
            for (int k = 31; k >= 0; --k) {
                if (t_lo + k <= t_hi && t_lo + k > 0) back(k);  // warp-uniform
            }
        }
        const int tt = t_lo + lane;
        if (tt <= t_hi) {
            // steps taken after frame tt inside this block: bits lane+1 .. 31
            const int keep = end_state - __popc(lane < 31 ? (steps >> (lane + 1)) : 0u);
            path[f0 + tt] = keep;
            if (unit_path) {
                const int pos = keep == 0 ? 0 : (keep > NE ? L - 1 : (keep - 1) / PC_EMIT);
                unit_path[f0 + tt] = v.labels[p0 + pos];
            }
        }
    }
    if (trace) g_vit_dbg[2] = clock64();
}

// block 0 clocks: [0] start, [1] recurrence done, [2] traceback done (tuning aid)
extern "C" int pc_debug_read_vit(long long *host_out) {
    return cudaMemcpyFromSymbol(host_out, g_vit_dbg, sizeof(long long) * 8) == cudaSuccess ? 0 : -2;
}

template <int SPL, typename E>
static int launch_vit(pc_handle h, const CorpusView &v, const E *b, const double *log_self,
                      const double *log_next, const double *utt_logpi, const double *state_logpi,
                      int32_t *path, int32_t *unit_path, double *score, cudaStream_t st) {
    const int words = ((v.max_frames + 31) / 32) * 32 * SPL;
    const size_t per_warp = (size_t)words * sizeof(uint32_t);
    int warps = (int)((200 * 1024) / per_warp);
    if (warps < 1) {
        pc_set_error("pc_viterbi: %d frames x %d labels needs %zu B of backpointers per utterance "
                     "(limit 200 KiB)", v.max_frames, v.max_labels, per_warp);
        return PC_ERR_UNSUPPORTED;
    }
    if (warps > 4) warps = 4;
    const size_t smem = per_warp * warps;
    auto kern = viterbi_kernel<SPL, E>;
    PC_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int blocks = (v.n_utt + warps - 1) / warps;
    kern<<<blocks, warps * 32, smem, st>>>(v, b, log_self, log_next, utt_logpi, state_logpi, path,
                                           unit_path, score, words);
    PC_LAUNCH_CHECK();
    h->launches++;
    return PC_OK;
}

template <typename E>
static int dispatch_vit(pc_handle h, const CorpusView &v, const E *b, const double *log_self,
                        const double *log_next, const double *utt_logpi, const double *state_logpi,
                        int32_t *path, int32_t *unit_path, double *score, cudaStream_t st) {
    const int states = PC_EMIT * v.max_labels + 2;
    if (states <= 32) return launch_vit<1, E>(h, v, b, log_self, log_next, utt_logpi, state_logpi, path, unit_path, score, st);
    if (states <= 64) return launch_vit<2, E>(h, v, b, log_self, log_next, utt_logpi, state_logpi, path, unit_path, score, st);
    if (states <= 128) return launch_vit<4, E>(h, v, b, log_self, log_next, utt_logpi, state_logpi, path, unit_path, score, st);
    if (states <= 256) return launch_vit<8, E>(h, v, b, log_self, log_next, utt_logpi, state_logpi, path, unit_path, score, st);
    pc_set_error("pc_viterbi: %d labels per utterance exceeds the limit of 84", v.max_labels);
    return PC_ERR_UNSUPPORTED;
}

int launch_viterbi(pc_handle h, const CorpusView &v, const float *b, const double *b64,
                   const double *log_self, const double *log_next, const double *utt_logpi,
                   const double *state_logpi, int32_t *path, int32_t *unit_path, double *score,
                   cudaStream_t st) {
    if (v.n_utt == 0) return PC_OK;
    if (b64) return dispatch_vit<double>(h, v, b64, log_self, log_next, utt_logpi, state_logpi, path, unit_path, score, st);
    return dispatch_vit<float>(h, v, b, log_self, log_next, utt_logpi, state_logpi, path, unit_path, score, st);
}
