/*
 * poccala_b200.h — C ABI of the B200-native GMM-HMM E-step engine.
 *
 * This is the drop-in boundary for Poccala's embedded Baum-Welch hot path.  The reference
 * (pure Python/numpy, /root/reference) has no FFI layer of its own: its boundary is the Python
 * class surface (StatisticalModel/LHMM.py, StatisticalModel/Clustering.py,
 * AcousticModel/AcousticModel.py).  poccala_b200/ mirrors those classes and binds the entry
 * points below with ctypes; INTEGRATION.md shows the stub a maintainer of the reference would add.
 * Every entry point cites the reference code it replaces (paths relative to the reference root).
 *
 * Conventions
 *   - plain C types only; all `dev` pointers are CUDA device pointers owned by the caller;
 *     `host` pointers are ordinary host memory.  The library never frees caller memory.
 *   - every call returns 0 (PC_OK) or a negative PC_ERR_* code; pc_last_error() gives the message
 *     for the calling thread.  No C++ exception crosses the boundary.
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream).  Device-pointer calls
 *     are asynchronous on that stream and perform no hidden host synchronisation unless stated.
 *   - a pc_handle is bound to one CUDA device and is not thread-safe.
 *   - there is NO CPU fallback: without a CUDA device pc_create fails with PC_ERR_CUDA.
 *
 * Data layout in HBM
 *   X      pc_corpus_frames_bytes(c) bytes: float [F][PC_XS] standardised frames (cols [0,D) data,
 *          col 39 = 1.0) followed by one fp16 (hi, lo) tensor-core operand image per 128-frame tile
 *   W      pc_gmm_bytes(G) bytes: float [G][PC_KA] packed Gaussians (mu/var (39), k_hi |
 *          -1/(2 var) (39), k_lo), float scale[G], int32 flags[G], one fp16 operand image per unit
 *   b,lgam float per utterance [T][SP] time-major emissions, s = 3*position + state, SP = 3L up to 8
 *   acc    double[n_gauss][PC_KA]          sum gamma*x (39), sum gamma | sum gamma*x^2 (39), sum gamma
 *   Gaussian index g = (unit*3 + state)*mix + m.
 */
#ifndef POCCALA_B200_H
#define POCCALA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PC_ABI_VERSION 3

#define PC_OK 0
#define PC_ERR_INVALID (-1)     /* bad argument (shape, NULL pointer, dimension mismatch) */
#define PC_ERR_CUDA (-2)        /* CUDA runtime error or no usable device */
#define PC_ERR_UNSUPPORTED (-3) /* valid request outside what the kernels cover */
#define PC_ERR_NOMEM (-4)

#define PC_DIM_MAX 39 /* feature dimension limit (39-dim MFCC, init.py:27-43) */
#define PC_XS 40      /* floats per frame row of X */
#define PC_KA 80      /* augmented contraction length [x (39), 1 | x^2 (39), 1] */
#define PC_EMIT 3     /* emitting states per unit HMM (state_num 5, AcousticModel.py:39) */
#define PC_STATES 5
#define PC_TRANS_SLOTS 9 /* per unit: 3 x (self, next, gamma) log-domain transition accumulators */
#define PC_W_BYTES_PER_GAUSS 1190 /* upper bound: 320 (fp32 row) + 8 (scale, flags) + 2 x <= 427 (fp16 images) */

typedef struct pc_handle_s *pc_handle;
typedef struct pc_corpus_s *pc_corpus;

int pc_abi_version(void);
int64_t pc_rows_bytes(int64_t n_frames); /* size of a rows-only X buffer (pc_prepare_rows_*) */
int64_t pc_gmm_bytes(int64_t n_gauss);   /* size of a W buffer */
const char *pc_last_error(void);

/* One handle per device.  Fails (PC_ERR_CUDA) when the device is absent or not sm_100. */
int pc_create(int device, pc_handle *out);
int pc_destroy(pc_handle h);
/* Options: "tensor_core" 0 = CUDA-core kernels, 1 = tcgen05/TMA kernels (default when the shape is
 * covered; the tests cross-check both on the same inputs); "host_chunks" caps the transfer pipeline
 * depth of pc_em_iteration_host (0 = 8); "debug_flags" is a tuning aid for the tcgen05 kernels (skip
 * stages, record block 0's phase clocks); "k1_kernel" 1 (default) = scoring kernel with wide accumulators
 * (several label positions per tcgen05 accumulator) for units of <= 16 mixtures and resident frame tiles with the
 * unit images in pieces for 64 mixtures, 0 = the per-position-pair kernel for every mixture count; "k3_kernel" 1 (default) = accumulation kernel with the Gaussians on the
 * accumulator lanes and gathered 32-frame blocks, 0 = one (128-frame tile, unit) pair at a time; "k2_kernel" 1 (default) = one warp per utterance, posteriors
 * normalised by the utterance likelihood; 0 = three warps per utterance with a per-frame normaliser (the
 * tests cross-check both); "kmeans_cluster" 1 (default) = k-means problems run on a thread-block cluster each when
 * they are few and large, 0 = always one CTA per problem, 2 = always a cluster (same results, bit for bit);
 * read-only: "launches" (kernels launched), "sm_count", "clamped" (standardised
 * feature values the frame preparation had to clamp to +-240 since the counter was last read: synchronises
 * the device and clears the counter; anything but 0 means the frames were not standardised - see
 * pc_frame_moments_host). */
int pc_set_option(pc_handle h, const char *key, int64_t value);
int64_t pc_get_option(pc_handle h, const char *key);

/* ---- corpus descriptors -------------------------------------------------------------------
 * Replaces the per-utterance Python lists that AcousticModel.embedded_training walks
 * (AcousticModel.py:842-870: one task per (label, data)).  Host arrays in, device descriptor
 * tables (owned by the corpus object) out.  labels are unit indices in [0, n_units). */
int pc_corpus_create(pc_handle h, int32_t n_utt, const int32_t *host_n_frames,
                     const int32_t *host_n_labels, const int32_t *host_labels, int32_t n_units,
                     pc_corpus *out);
int pc_corpus_destroy(pc_corpus c);
int64_t pc_corpus_total_frames(pc_corpus c);
int64_t pc_corpus_total_tiles(pc_corpus c);
/* (tile, unit) pairs the last pc_accumulate on this corpus actually contracted: pairs whose
 * posteriors all vanish below the resolution of the kernel's fp16 P tile contribute exactly zero
 * and are skipped (synchronises the device; diagnostics / bench accounting). */
int64_t pc_corpus_active_tiles(pc_corpus c);
int64_t pc_corpus_emission_floats(pc_corpus c); /* length of the b / lgam buffers */
int64_t pc_corpus_total_pairs(pc_corpus c);     /* number of (utterance, label position) pairs */
int64_t pc_corpus_total_states(pc_corpus c);    /* sum over utterances of 3L+2 */
int64_t pc_corpus_frames_bytes(pc_corpus c);    /* size of the corpus' X buffer */
/* host_out[n_utt+1]: first frame / first float in b / first pair / first composite state. */
int pc_corpus_offsets(pc_corpus c, int64_t *host_frame_off, int64_t *host_emis_off,
                      int64_t *host_pair_off, int64_t *host_state_off);

/* ---- model packing -------------------------------------------------------------------------
 * util.gaussian_function (util.py:20-36, log branch incl. the -1/2*sum(var) normaliser, Q1) and
 * the log(alpha) term of Clustering.GMM.point (Clustering.py:753-757), folded into one row per
 * Gaussian so that score = <[x, x^2, 1, 1], W_g>.  mean/var/alpha: dev double [n_gauss][dim] /
 * [n_gauss]; shift/inv_scale: dev double [dim] or NULL (identity). */
/* mix = components per state when the Gaussians are the [unit][3][mix] set of an acoustic model
 * (builds the per-unit tensor-core operand images); mix = 0 for a flat list of Gaussians. */
int pc_pack_gmm(pc_handle h, const double *dev_mean, const double *dev_var,
                const double *dev_alpha, const double *dev_shift, const double *dev_inv_scale,
                int32_t n_gauss, int32_t dim, int32_t mix, float *dev_W, void *stream);

/* Corpus frames [F][dim] (double or float, device, utterances concatenated in corpus order) ->
 * X buffer: standardised fp32 rows + per-tile fp16 operand images. */
int pc_prepare_frames_f64(pc_handle h, pc_corpus c, const double *dev_x, int32_t dim,
                          const double *dev_shift, const double *dev_inv_scale, float *dev_X,
                          void *stream);
int pc_prepare_frames_f32(pc_handle h, pc_corpus c, const float *dev_x, int32_t dim,
                          const double *dev_shift, const double *dev_inv_scale, float *dev_X,
                          void *stream);
/* Frames without a corpus (dense scoring sweep): fp32 rows [n][PC_XS] only. */
int pc_prepare_rows_f64(pc_handle h, const double *dev_x, int64_t n, int32_t dim,
                        const double *dev_shift, const double *dev_inv_scale, float *dev_X,
                        void *stream);
int pc_prepare_rows_f32(pc_handle h, const float *dev_x, int64_t n, int32_t dim,
                        const double *dev_shift, const double *dev_inv_scale, float *dev_X,
                        void *stream);

/* ---- K1: GMM scoring -----------------------------------------------------------------------
 * LHMM.cal_observation_pro -> GMM.point -> util.gaussian_function + util.log_sum_exp
 * (LHMM.py:163-187, Clustering.py:740-767, util.py:20-36,54-77) for every (utterance, label
 * position, emitting state, frame).  Writes b (emission log-likelihoods). */
int pc_gmm_score(pc_handle h, pc_corpus c, const float *dev_X, const float *dev_W, int32_t mix,
                 float *dev_b, void *stream);

/* Dense scoring sweep (BASELINE config 3): n frames against n_states GMMs of `mix` components,
 * out[n][n_states] = log-sum-exp over each state's components. */
int pc_gmm_score_dense(pc_handle h, const float *dev_X, int64_t n, const float *dev_W,
                       int32_t n_states, int32_t mix, float *dev_out, void *stream);

/* ---- K2: forward-backward ------------------------------------------------------------------
 * LHMM.baulm_welch loop on the sentence HMM assembled by AcousticModel.embedded
 * (AcousticModel.py:957-1014; LHMM.py:335-366 forward/backward, :426-471 ksai/gamma/pi,
 * :412-422 likelihood, :526-544 pi-iteration with the 0.64 threshold, :486-500 per-frame
 * normalised log gamma).  log_self/log_next: dev double [n_units][5] = log of the unit HMM's
 * transmat diagonal / super-diagonal (host-computed np.log).  Outputs: lgam (same layout as b),
 * utt_logp double [n_utt], utt_iters int32 [n_utt] (Baum-Welch iterations the reference would
 * run), pair_trans float [n_pairs][9]: log expected (self, next, occupancy) counts over t<T-1 per
 * emitting state, RELATIVE to utt_logp (reference value = utt_logp + pair_trans, Q6).
 * Side effect: the corpus object remembers, for dev_lgam, which (128-frame tile, label position)
 * pairs carry posterior mass (every log gamma of the others lies below log 2^-41); the next
 * pc_accumulate on this corpus with the same dev_lgam pointer uses that once instead of re-reading
 * the rows.  A caller that edits dev_lgam in place between the two calls must pass it through a
 * different pointer, or call pc_forward_backward again. */
int pc_forward_backward(pc_handle h, pc_corpus c, const float *dev_b, const double *dev_log_self,
                        const double *dev_log_next, float *dev_lgam, double *dev_utt_logp,
                        int32_t *dev_utt_iters, float *dev_pair_trans, void *stream);

/* log of the diagonal / super-diagonal of every unit transmat [n_units][5][5] -> [n_units][5] each
 * (np.log(transmat) of LHMM.py:340,358 restricted to the two bands; the bit-exact Viterbi path takes
 * host-computed logs instead). */
int pc_log_bands(pc_handle h, const double *dev_transmat, int32_t n_units, double *dev_log_self,
                 double *dev_log_next, void *stream);

/* ---- K3: Baum-Welch accumulation -----------------------------------------------------------
 * LHMM.update_acc -> Clustering.GMM.update_acc (LHMM.py:473-507, Clustering.py:653-680) in the
 * linear-equivalent form of SURVEY A.4: acc[g] += sum_t gamma_t(j,m) * [x, x^2, 1, 1].
 * dev_acc is accumulated into (caller zeroes it at the start of an EM iteration).  Pairs without
 * posterior mass are skipped exactly (their fp16 posterior tiles are all zero): the flags come from
 * the preceding pc_forward_backward on the same dev_lgam, else from a pre-pass over dev_lgam. */
int pc_accumulate(pc_handle h, pc_corpus c, const float *dev_X, const float *dev_W, int32_t mix,
                  const float *dev_b, const float *dev_lgam, double *dev_acc, void *stream);

/* ---- K6 local part: log-domain transition accumulators -------------------------------------
 * LHMM.add_acc / init_acc (LHMM.py:149-161,256-290): log-sum-exp over every (utterance, position)
 * of a unit.  Two steps so that a cross-rank allreduce(max) / allreduce(sum) can sit between them:
 *   pc_transitions_max : dev_max[n_units][9] = max(dev_max, utt_logp + pair_trans)
 *   pc_transitions_sum : dev_sum[n_units][9] += exp(utt_logp + pair_trans - dev_max)
 * accumulator = dev_max + log(dev_sum). */
int pc_transitions_max(pc_handle h, pc_corpus c, const double *dev_utt_logp,
                       const float *dev_pair_trans, double *dev_max, void *stream);
int pc_transitions_sum(pc_handle h, pc_corpus c, const double *dev_utt_logp,
                       const float *dev_pair_trans, const double *dev_max, double *dev_sum,
                       void *stream);

/* ---- M-step --------------------------------------------------------------------------------
 * LHMM.update_param + Clustering.GMM.update_param (LHMM.py:509-524, Clustering.py:682-693):
 * alpha = occ/socc, mean = sx/occ, var = max(E[(x-mu_old)^2], c_covariance) (Q8), transmat rows
 * 1..3 = exp(ksai_acc - gamma_acc).  Parameters are updated IN PLACE (dev double).  fix_code bit 4
 * locks transmat, bit 2 locks the GMMs (LHMM.py:140-145).  Units with zero occupancy keep their
 * parameters. */
int pc_update_params(pc_handle h, int32_t n_units, int32_t mix, int32_t dim, const double *dev_acc,
                     const double *dev_trans_max, const double *dev_trans_sum,
                     const double *dev_shift, const double *dev_inv_scale, double c_covariance,
                     int32_t fix_code, double *dev_mean, double *dev_var, double *dev_alpha,
                     double *dev_transmat /* [n_units][5][5] */, void *stream);

/* ---- K4: Viterbi forced alignment ----------------------------------------------------------
 * LHMM.viterbi (LHMM.py:546-609) on the banded sentence HMM: fp64 max-plus recurrence, ties to
 * the lower state index, end state = first argmax over all states.  Emissions: the float buffer
 * written by pc_gmm_score (dev_b) or a double buffer of the same layout (dev_b64); exactly one
 * is non-NULL.  logpi: dev double [n_utt] (uniform value per utterance, np.log(1/N)) or
 * dev_state_logpi [total_states] (general); exactly one is non-NULL.  Outputs: path int32
 * [n_frames_total] composite state index per frame, optional unit label per frame
 * (AcousticModel.py:1016-1027 convert=True), score double [n_utt]. */
int pc_viterbi(pc_handle h, pc_corpus c, const float *dev_b, const double *dev_b64,
               const double *dev_log_self, const double *dev_log_next, const double *dev_utt_logpi,
               const double *dev_state_logpi, int32_t *dev_path, int32_t *dev_unit_path,
               double *dev_score, void *stream);

/* ---- K5: k-means initialisation ------------------------------------------------------------
 * ClusterInitialization.kmeans(algorithm=1) greedy passes (Clustering.py:894-940) with the
 * dimension-0 metric of cal_distance (:796-801, Q2), run to convergence for `n_problems`
 * independent problems (one per HMM state, AcousticModel.py:553-554) in one launch.
 *   host_point_off [n_problems+1]  HOST array: rows of dev_x belonging to each problem
 *   dev_x          double [total_points][dim]
 *   dev_seed_points int32 [n_problems][k]: point index (within the problem) of each seed, drawn by
 *                  the caller with Python's `random` exactly like Clustering.py:975-1020
 *   dev_workspace  pc_kmeans_workspace_bytes(...) bytes of scratch
 * Outputs: dev_owner int32 [total_points] (cluster owning each point, -1 = never claimed);
 * dev_member_list int32 [total_points + n_problems*k]: problem p's region starts at
 * host_point_off[p] + p*k and holds its clusters back to back, each in the reference's dict
 * insertion order (seed first); dev_member_count int32 [n_problems][k]; dev_passes int32
 * [n_problems]; dev_moves int64 [n_problems] (= number of random.random() key draws the reference
 * makes, Clustering.py:932).  See DESIGN.md "K5". */
int64_t pc_kmeans_workspace_bytes(int32_t n_problems, const int64_t *host_point_off, int32_t k);
int pc_kmeans_run(pc_handle h, int32_t n_problems, const int64_t *host_point_off,
                  const double *dev_x, int32_t dim, int32_t k, const int32_t *dev_seed_points,
                  void *dev_workspace, int32_t *dev_owner, int32_t *dev_member_list,
                  int32_t *dev_member_count, int32_t *dev_passes, int64_t *dev_moves,
                  int64_t max_passes, void *stream);
/* Cluster statistics the reference returns (Clustering.py:880-891 cal_center, :807-832
 * cal_variance(algorithm='kmeans') with its 1e-4 floor, :947 alpha = len/n): sequential fp64 sums
 * in insertion order.  dev_mean / dev_var double [n_problems][k][dim], dev_alpha [n_problems][k].
 * dev_workspace is the buffer pc_kmeans_run used (it holds the problem table). */
int pc_kmeans_finish(pc_handle h, int32_t n_problems, const int64_t *host_point_off,
                     const double *dev_x, int32_t dim, int32_t k, const void *dev_workspace,
                     const int32_t *dev_member_list, const int32_t *dev_member_count,
                     double *dev_mean, double *dev_var, double *dev_alpha, void *stream);

/* ---- alignment post-processing ---------------------------------------------
 * The steps either side of Viterbi / k-means in the reference's mode-1 training.
 * pc_segment_keys: per frame the (unit, emitting state) whose data set the frame joins, as
 * key = unit * 3 + state, or -1 when the frame is dropped.
 *   mode 0  uniform segmentation, AcousticModel.__eq_segment(mode='e') (AcousticModel.py:605-612):
 *           L chunks of T // L frames per utterance, the remainder dropped; dev_path is ignored
 *   mode 1  after forced alignment, multi_process_data (AcousticModel.py:750-764): dev_path is the
 *           composite-state path pc_viterbi wrote; the per-frame unit sequence is cut into maximal
 *           runs of one unit (discriminate, :937-955); an utterance whose path visits fewer
 *           distinct units than its label holds is dropped (dev_utt_kept = 0, :753-757)
 * Every segment of n frames is then cut into 3 parts of n // 3 frames, the last taking the
 * remainder (__eq_segment(mode='g') :613-626, __get_gmmdata :630-644).
 * dev_frame_key int32 [total_frames]; dev_utt_kept int32 [n_utt] or NULL. */
int pc_segment_keys(pc_handle h, pc_corpus c, int32_t mode, const int32_t *dev_path,
                    int32_t *dev_frame_key, int32_t *dev_utt_kept, void *stream);

/* Stable counting sort of the frames by key: frames of one key become contiguous, in ascending
 * frame order (= utterance, then time: the order __get_gmmdata concatenates segments in).
 *   dev_key_off int64 [n_keys + 2]: first slot of each key; [n_keys] = kept frames (keys outside
 *               [0, n_keys) sort last); [n_keys + 1] = n_frames
 *   dev_order   int32 [n_frames]: frame index held by each slot
 * n_frames < 2^31, n_keys <= 12000.  Workspace: pc_group_workspace_bytes (-1 on bad sizes). */
int64_t pc_group_workspace_bytes(int64_t n_frames, int32_t n_keys);
int pc_group_frames(pc_handle h, const int32_t *dev_frame_key, int64_t n_frames, int32_t n_keys,
                    void *dev_workspace, int64_t *dev_key_off, int32_t *dev_order, void *stream);

/* dst[i] = src[order[i]] for rows of row_bytes bytes (a multiple of 4): the per-state data sets
 * (AcousticModel.__get_gmmdata) that pc_kmeans_run and the stand-alone GMM EM consume. */
int pc_gather_rows(pc_handle h, const int32_t *dev_order, int64_t n_rows, int32_t row_bytes,
                   const void *dev_src, void *dev_dst, void *stream);

/* ---- host-buffer entry point (end-to-end) --------------------------------------------------
 * Standardisation constants of a corpus: per-dimension sum and sum of squares (fp64) of host frames
 * [n_frames][dim] float, reduced on the device.  The caller forms shift = sum/n and inv_scale =
 * 1/sqrt(sumsq/n - shift^2) - across ranks after adding up the three quantities - once per corpus
 * (the expanded quadratic of the scoring contraction cancels when |mu| >> sigma). */
int pc_frame_moments_host(pc_handle h, const float *host_frames, int64_t n_frames, int32_t dim,
                          double *host_sum, double *host_sumsq, void *stream);

/* Cross-rank reduction hook of pc_em_iteration_host (the reference merges accumulator files of all
 * machines, LHMM.py:256-290, Clustering.py:314-367).  With a hook installed the entry point keeps its
 * exchange buffers in caller-owned device memory - dev_tmax double [n_units][9], dev_flat double
 * [n_gauss*PC_KA + n_units*9] (GMM statistics, then the transition sums) - and calls
 *     fn(user, 0, stream)   after the local transition maxima are in dev_tmax  (all-reduce MAX)
 *     fn(user, 1, stream)   after the local statistics are in dev_flat          (all-reduce SUM)
 * on the host thread, between kernel launches; `stream` is the CUDA stream the collective must be
 * queued on.  A non-zero return aborts the iteration (PC_ERR_CUDA).  fn = NULL removes the hook. */
typedef int (*pc_reduce_hook)(void *user, int32_t op, void *stream);
int pc_set_reduce_hook(pc_handle h, pc_reduce_hook fn, void *user, double *dev_tmax, double *dev_flat,
                       int64_t flat_len);

/* Front end.  pc_mfcc: AudioProcessing.MFCC.mfcc (AudioProcessing.py:416-448: pre-emphasis
 * :195-198, framing :215-227, the per-frame Hamming factor :243-246, |rfft| :262-263, filter bank + frame energy
 * :328-343, log + DCT :355-368, deltas :405-412) on device buffers, fp64.  dev_signal double [n_samples] (the samples
 * after AudioProcessing.MFCC.init_audio's zero removal); frame geometry as frame_blocking computes it (framesize =
 * int(rate * sampletime), step = int(framesize * overlap), n_frames = 1 + ceil((n_samples - framesize) / step) >= 2);
 * nfft a power of two in [64, 2048]; dev_fbank double [n_filters][nfft/2 + 1] the triangular responses (host-built
 * with the reference's expression); dev_out double [n_frames][n_ceps * (1 + n_delta)], n_delta in {0, 1, 2}.
 * pc_vad_distance: VAD.mel_distance + VAD.osf (:462-506): dev_dist / dev_dist_osf double [n_frames]. */
int pc_mfcc(pc_handle h, const double *dev_signal, int64_t n_samples, int32_t framesize, int32_t step,
            int32_t n_frames, int32_t nfft, const double *dev_fbank, int32_t n_filters, int32_t n_ceps,
            int32_t cal_energy, int32_t n_delta, double *dev_out, void *stream);
int pc_vad_distance(pc_handle h, const double *dev_mfcc, int32_t n_frames, int32_t dim, int32_t sample_size,
                    double alpha, double beta, double *dev_dist, double *dev_dist_osf, void *stream);

/* Cross-rank reduction over PEER MEMORY (one process per GPU on one NVLink / NVSwitch node), the
 * device-side replacement of the accumulator-file merge (LHMM.py:256-290, Clustering.py:314-367;
 * AcousticModel.py:842-882): every rank keeps its statistics in an exchange block the other ranks map
 * through CUDA IPC; the M-step of a state runs on its owner (state mod n_ranks), whose kernel adds the N copies of
 * the state's rows in rank order while it reads them over NVLink and publishes the new parameters, which the other
 * ranks copy - no collective library call, replicas bit-identical.
 *   pc_peer_create   allocates this rank's block (two alternating statistic sets + the reduced set + arrival
 *                    flags) and returns its 64-byte IPC handle;
 *   pc_peer_connect  takes the handles of ALL ranks in rank order ([n_ranks][64], own entry ignored); call a
 *                    host barrier between pc_peer_connect and the first iteration;
 *   pc_peer_buffers  device pointers of a statistic set: which = 0 / 1 the set of an even / odd iteration
 *                    (option "peer_epoch" & 1 names the current one), 2 = the reduced sums of the last M-step:
 *                    acc double [n_gauss][PC_KA], tsum / tmax double [n_units][9] (tsum relative to the SAME
 *                    rank's tmax - pc_transitions_max / _sum on the local pairs, no collective in between);
 *   pc_update_params_peer  pc_update_params on the sum over the ranks of the current set; every rank must call
 *                    it once per iteration.  Read-only option "peer_timeouts": arrival waits that gave up (a rank
 *                    that never called) - 0 in a healthy run.
 * With a connected block and no reduce hook pc_em_iteration_host uses this path by itself. */
#define PC_MAX_PEERS 16
#define PC_IPC_HANDLE_BYTES 64
int pc_peer_create(pc_handle h, int32_t rank, int32_t n_ranks, int64_t n_gauss, int32_t n_units,
                   uint8_t *handle_out);
int pc_peer_connect(pc_handle h, const uint8_t *handles);
int pc_peer_buffers(pc_handle h, int32_t which, double **acc, double **tsum, double **tmax);
int pc_update_params_peer(pc_handle h, int32_t mix, int32_t dim, const double *dev_shift,
                          const double *dev_inv_scale, double c_covariance, int32_t fix_code, double *dev_mean,
                          double *dev_var, double *dev_alpha, double *dev_transmat, void *stream);
int pc_peer_destroy(pc_handle h);

/* One full EM iteration the way AcousticModel.embedded_training runs it (AcousticModel.py:842-882)
 * with HOST inputs and outputs: frames [total_frames][dim] float (pinned or pageable), parameters
 * double, updated in place on return.  Copies host->device, runs K1,K2,K3,K6 (+ the reduce hook),
 * the M-step, copies the new parameters and sum log-likelihood back, and synchronises the stream.
 * host_shift / host_inv_scale double [dim]: the corpus' standardisation constants
 * (pc_frame_moments_host); both NULL = computed from this call's frames (the scoring then waits
 * for the whole copy instead of overlapping it).  Fails with PC_ERR_INVALID when a standardised
 * feature exceeds +-240 (wrong constants). */
int pc_em_iteration_host(pc_handle h, pc_corpus c, const float *host_frames, int32_t dim,
                         int32_t n_units, int32_t mix, double *host_mean, double *host_var,
                         double *host_alpha, double *host_transmat, const double *host_shift,
                         const double *host_inv_scale, double c_covariance, int32_t fix_code,
                         double *host_sum_logp, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* POCCALA_B200_H */
