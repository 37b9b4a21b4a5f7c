// Shared declarations for the sm_100a kernels behind include/poccala_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/poccala_b200.h"

#define PC_TILE_ROWS 128  // frames per work tile (one tcgen05 M=128 accumulator block)
#define PC_BLOCK_ROWS 32  // frames per activity block: K3 gathers the blocks of a tile that carry posterior mass
#define PC_XTILE_BYTES (2 * (PC_KA / 8) * PC_TILE_ROWS * 16)  // one frame-tile operand image (40 KiB)
#define PC_X32TILE_BYTES (PC_TILE_ROWS * PC_XS * 4)           // one frame tile as fp32, [4 blocks][10 quads][32 rows][4] (20 KiB)
#define PC_WGROUP_BYTES (2 * (PC_KA / 8) * 128)               // 8 Gaussian rows of a unit image (2560 B)

// byte offsets of the fp16 operand images inside the W / X buffers (pack.cu)
__host__ __device__ inline size_t pc_w16_offset(int64_t n_gauss) { return ((size_t)n_gauss * 328 + 127) & ~(size_t)127; }
// second image set behind the first: the same units with the hi and the lo halves as two contiguous pieces
// (row groups of 1 280 B each), for kernels that move them separately (score_tc_big.cu)
__host__ __device__ inline size_t pc_w16s_offset(int64_t n_gauss, int n_unit) {
    return pc_w16_offset(n_gauss) + (size_t)(n_gauss / n_unit) * 2 * (PC_KA / 8) * ((n_unit + 15) & ~15) * 16;
}
__host__ __device__ inline size_t pc_x16_offset(int64_t n_frames) { return ((size_t)n_frames * PC_XS * 4 + 127) & ~(size_t)127; }
__host__ __device__ inline size_t pc_x32_offset(int64_t n_frames, int64_t n_xtiles) {
    return pc_x16_offset(n_frames) + (size_t)n_xtiles * PC_XTILE_BYTES;
}
#define PC_NEG_INF (-INFINITY)
#define PC_L2_RUN_BYTES (32ll << 20)  // frame-tile images one (run, unit) block of K3 work items may span
// A (tile, unit) pair whose log gamma all lie below log(2^-41) contributes exactly nothing to K3 (see
// accumulate_tc.cu); K2 sets the activity flags while it writes the rows, K3's own pre-pass does it
// for log gamma that came from elsewhere.
#define PC_ACTIVE_MIN_LGAM (-41.f * 0.6931471805599453f)
#define PC_TR_CHUNKS 16  // blocks per unit in the transition reductions (fixed: the summation order depends on the corpus only)
#define PC_MAX_CHUNKS 8  // host-buffer entry point: transfer / prepare / score pipeline depth

// Device-side view of a corpus (all pointers device memory owned by pc_corpus_s).
struct CorpusView {
    int32_t n_utt;
    int32_t n_units;
    int64_t n_pairs;
    int64_t n_tiles;
    int32_t n_items;
    int32_t max_frames;
    int32_t max_labels;
    const int64_t *frame_off;   // [n_utt+1] first row of the utterance in X
    const int64_t *emis_off;    // [n_utt+1] first float of the utterance's [T][SP] emission block
    const int64_t *pair_off;    // [n_utt+1] first (utt, position) pair = first label
    const int64_t *state_off;   // [n_utt+1] first composite state (3L+2 per utterance)
    const int32_t *labels;      // [n_pairs] unit id, utterance-major
    const int32_t *pair_utt;    // [n_pairs] utterance of each utterance-major pair
    const int32_t *fb_order;    // [n_utt] utterances sorted by descending T*(3L+1)
    // unit-major work decomposition for K1 / K3
    const int64_t *sorted_pair;  // [n_pairs] utterance-major pair index, sorted by unit (stable)
    const int64_t *unit_pair_off;  // [n_units+1] range of sorted pairs per unit
    const int64_t *tile_pair;    // [n_tiles] utterance-major pair index of the tile
    const int32_t *tile_t0;      // [n_tiles] first frame (within the utterance) of the tile
    const int32_t *tile_rows;    // [n_tiles] frames in the tile (<= 128)
    const int32_t *tile_tp;      // [n_tiles] frame stride SP (floats) of the utterance's b / lgam block
    const int64_t *tile_xrow;    // [n_tiles] first row of the tile in X
    const int64_t *tile_boff;    // [n_tiles] float offset of (frame t0, state 0 of the pair) in b / lgam
    const int64_t *pair_tile0;   // [n_pairs] first unit-major tile of the pair (its tiles are consecutive)
    const int64_t *item_tile_lo;  // [n_items+1] tile range of each work item (one unit per item)
    const int32_t *item_unit;     // [n_items]
    float *scratch0;              // [total_frames] float4 per frame (K2 scratch: beta_hat of the entry state)
    float *scratch1;              // [emission floats] K2 scratch: beta_hat rows, same layout as b / lgam
    int32_t *tile_active;         // [n_tiles] K3 scratch: bit k = the tile's 32-frame block k carries posterior mass
    const int32_t *tile_item;     // [n_tiles] work item of each unit-major tile
    int32_t *item_act;            // [n_items + 1] K3 scratch, directly behind tile_active: active tiles of the item; [n_items] = block ticket counter
    int32_t *item_order;          // [n_items] K3 scratch: items sorted by active tiles, heaviest first
    double *trans_tmp;            // [n_units][PC_TR_CHUNKS][9] partial maxima / sums of the transition reductions
    int32_t *trans_cnt;           // [n_units] blocks of a unit that have delivered their partial (self-resetting)
    // utterance-major work decomposition for K1: groups of <= 3 consecutive tiles of one utterance
    int64_t total_frames;
    int32_t n_sitems;
    const int32_t *sitem_utt;     // [n_sitems]
    const int32_t *sitem_t0;      // [n_sitems] first frame of the group inside the utterance
    const int32_t *sitem_nt;      // [n_sitems] tiles in the group (1..3)
    // frame-tile operand images (one per 128-frame tile of every utterance, utterance-major)
    int64_t n_xtiles;
    const int64_t *xtile_off;     // [n_utt+1] first image of the utterance
    const int32_t *xtile_utt;     // [n_xtiles]
    const int32_t *xtile_t0;      // [n_xtiles] first frame of the image inside the utterance
    const int64_t *tile_xblk;     // [n_tiles] image index of each unit-major work tile
};

// emissions are time-major: b[t][s], s = 3*position + state, frame stride SP = 3L rounded up to 8
__host__ __device__ inline int pc_spad(int n_labels) { return (PC_EMIT * n_labels + 7) & ~7; }

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// log(exp(a)+exp(b)) in fp32 with -inf handling (util.py:54-77 semantics: an infinite maximum is
// returned unchanged).
__device__ __forceinline__ float logadd_f(float a, float b) {
    float m = fmaxf(a, b);
    float d = -fabsf(a - b);              // NaN only when a == b == -inf
    d = (m == PC_NEG_INF) ? 0.f : d;
    return m + __logf(1.f + __expf(d));
}

// Host-side launch helpers (api.cu)
const char *pc_set_error(const char *fmt, ...);
#define PC_CUDA_TRY(expr)                                                                 \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess) {                                                          \
            pc_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                         __LINE__);                                                       \
            return PC_ERR_CUDA;                                                           \
        }                                                                                 \
    } while (0)
#define PC_REQUIRE(cond, ...)          \
    do {                               \
        if (!(cond)) {                 \
            pc_set_error(__VA_ARGS__); \
            return PC_ERR_INVALID;     \
        }                              \
    } while (0)
#define PC_LAUNCH_CHECK()                                                                  \
    do {                                                                                   \
        cudaError_t _e = cudaGetLastError();                                               \
        if (_e != cudaSuccess) {                                                           \
            pc_set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, \
                         __LINE__);                                                        \
            return PC_ERR_CUDA;                                                            \
        }                                                                                  \
    } while (0)

struct pc_handle_s {
    int device;
    int sm_count;
    int use_tc;          // option "tensor_core"
    // option "debug_flags" (tuning aid for the tcgen05 kernels): 1 skip epilogue math, 2 skip MMAs, 4 / 8 skip
    // the B / A bulk copies, 16 skip the epilogue, 32 record block 0's phase clocks (pc_debug_read*)
    int debug_flags;
    int host_chunks;     // option "host_chunks": cap on the transfer pipeline depth (0 = PC_MAX_CHUNKS)
    int64_t launches;    // kernels launched through this handle (bench: gpu_launches)
    // workspace for the host-buffer entry point
    void *ws;
    size_t ws_bytes;
    void *pinned;
    size_t pinned_bytes;
    cudaStream_t copy_stream;                 // host-buffer entry point: frames travel on their own stream
    cudaEvent_t chunk_ev[PC_MAX_CHUNKS];
    cudaEvent_t start_ev;
    cudaEvent_t fork_ev, join_ev;             // transition reductions beside the accumulation kernel
    // device counters: [PC_CNT_CLAMPED] standardised features clamped by the frame preparation
    int *dev_counters;
    int k1_kernel;       // option "k1_kernel": 1 = wide accumulators for <= 16 mixtures (score_tc_wide.cu), 0 = score_tc.cu only
    int k3_kernel;       // option "k3_kernel": 1 = Gaussians on the accumulator lanes, gathered frame blocks (accumulate_tcx.cu), 0 = accumulate_tc.cu
    int kmeans_cluster;  // option "kmeans_cluster": 1 = problems beyond 20 480 points run on a thread-block cluster (kmeans.cu)
    int k2_kernel;       // option "k2_kernel": 1 = one warp per utterance (fwdbwd_warp.cu), 0 = three warps (fwdbwd.cu)
    // cross-rank reduction hook of the host-buffer entry point (pc_set_reduce_hook)
    pc_reduce_hook hook;
    void *hook_user;
    double *hook_tmax, *hook_flat;
    int64_t hook_flat_len;
    // exchange blocks of the peer-memory reduction (peer.cu): [rank] = this rank's own allocation
    int peer_n, peer_rank;
    void *peer_block[PC_MAX_PEERS];
    int64_t peer_n_acc;
    int peer_n_units;
    int64_t peer_epoch;  // M-steps run so far; buffer parity of the current iteration = peer_epoch & 1
};
#define PC_CNT_CLAMPED 0
#define PC_CNT_PEER_TIMEOUT 1
#define PC_CNT_N 8

struct pc_corpus_s {
    pc_handle h;
    int device;
    CorpusView v;
    int64_t total_frames;
    int64_t emis_floats;
    int64_t total_states;
    void *dev_block;     // one allocation holding every table
    int64_t *host_frame_off, *host_emis_off, *host_pair_off, *host_state_off;
    int32_t items_per_chunk;
    const float *flags_lgam;  // log-gamma buffer whose K3 activity flags K2 left in v.tile_active (consumed by pc_accumulate)
    // runs of consecutive utterances the host-buffer entry point pipelines (copy k+1 under score k)
    int32_t n_chunks;
    int64_t chunk_frame[PC_MAX_CHUNKS + 1];  // first frame of each chunk
    int64_t chunk_xtile[PC_MAX_CHUNKS + 1];  // first frame-tile image
    int32_t chunk_sitem[PC_MAX_CHUNKS + 1];  // first K1 work item
};

// Kernel launchers implemented in the per-kernel translation units.
int launch_pack_gmm(pc_handle h, const double *mean, const double *var, const double *alpha,
                    const double *shift, const double *inv_scale, int n_gauss, int dim, int mix,
                    float *W, cudaStream_t st);
int launch_prepare_rows(pc_handle h, const void *x, int is_f64, int64_t n, int dim,
                        const double *shift, const double *inv_scale, float *X, cudaStream_t st);
// [xtile_lo, xtile_hi): range of frame-tile images (whole corpus: 0, cv.n_xtiles)
int launch_prepare_frames(pc_handle h, const CorpusView &cv, const void *x, int is_f64, int dim,
                          const double *shift, const double *inv_scale, float *X, int64_t xtile_lo,
                          int64_t xtile_hi, cudaStream_t st);
int launch_score_simt(pc_handle h, const CorpusView &v, const float *X, const float *W, int mix,
                      float *b, cudaStream_t st);
bool score_tc_supported(int mix);
// [item_lo, item_hi): range of utterance-major work items (whole corpus: 0, v.n_sitems)
int launch_score_tc(pc_handle h, const CorpusView &v, const float *X, const float *W, int mix,
                    float *b, int item_lo, int item_hi, cudaStream_t st);
bool score_tc_wide_supported(int mix);
bool score_tc_big_supported(int mix);
int launch_score_tc_big(pc_handle h, const CorpusView &v, const float *X, const float *W, int mix, float *b,
                        int item_lo, int item_hi, cudaStream_t st);
int launch_score_tc_wide(pc_handle h, const CorpusView &v, const float *X, const float *W, int mix,
                         float *b, int item_lo, int item_hi, cudaStream_t st);
bool accumulate_tc_supported(int mix);
// flags_fresh: the activity flags of `lgam` were set by launch_forward_backward (skip the pre-pass)
int launch_accumulate_tc(pc_handle h, const CorpusView &v, const float *X, const float *W, int mix,
                         const float *b, const float *lgam, double *acc, bool flags_fresh, cudaStream_t st);
bool accumulate_tcx_supported(int mix);
int launch_accumulate_tcx(pc_handle h, const CorpusView &v, const float *X, const float *W, int mix,
                          const float *b, const float *lgam, double *acc, bool flags_fresh, cudaStream_t st);
int launch_score_dense_simt(pc_handle h, const float *X, int64_t n, const float *W, int n_states,
                            int mix, float *out, cudaStream_t st);
int launch_accumulate_simt(pc_handle h, const CorpusView &v, const float *X, const float *W,
                           int mix, const float *b, const float *lgam, double *acc,
                           cudaStream_t st);
int launch_forward_backward(pc_handle h, const CorpusView &v, const float *b,
                            const double *log_self, const double *log_next, float *lgam,
                            float *scratch0, double *utt_logp, int32_t *utt_iters,
                            float *pair_trans, cudaStream_t st);
int launch_forward_backward_warp(pc_handle h, const CorpusView &v, const float *b,
                                 const double *log_self, const double *log_next, float *lgam,
                                 float *scratch0, double *utt_logp, int32_t *utt_iters,
                                 float *pair_trans, cudaStream_t st);
int launch_transitions_max(pc_handle h, const CorpusView &v, const double *utt_logp,
                           const float *pair_trans, double *tmax, cudaStream_t st);
int launch_transitions_sum(pc_handle h, const CorpusView &v, const double *utt_logp,
                           const float *pair_trans, const double *tmax, double *tsum,
                           cudaStream_t st);
int launch_update_params_peer(pc_handle h, int mix, int dim, const double *shift, const double *inv_scale, double c_cov,
                              int fix_code, double *mean, double *var, double *alpha, double *transmat, cudaStream_t st);
int launch_update_params(pc_handle h, int n_units, int mix, int dim, const double *acc,
                         const double *tmax, const double *tsum, const double *shift,
                         const double *inv_scale, double c_cov, int fix_code, double *mean,
                         double *var, double *alpha, double *transmat, cudaStream_t st);
int launch_viterbi(pc_handle h, const CorpusView &v, const float *b, const double *b64,
                   const double *log_self, const double *log_next, const double *utt_logpi,
                   const double *state_logpi, int32_t *path, int32_t *unit_path, double *score,
                   cudaStream_t st);
int launch_segment_keys(pc_handle h, const CorpusView &v, int mode, const int32_t *path, int32_t *key,
                        int32_t *kept, cudaStream_t st);
int64_t group_workspace_bytes(int64_t n_frames, int n_keys);
int launch_group_frames(pc_handle h, const int32_t *key, int64_t n, int n_keys, void *ws, int64_t *key_off,
                        int32_t *order, cudaStream_t st);
int launch_gather_rows(pc_handle h, const int32_t *order, int64_t n_rows, int row_bytes, const void *src,
                       void *dst, cudaStream_t st);
