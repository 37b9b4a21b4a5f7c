// K6 (local part) and the M-step.
//
// Transition accumulators: LHMM.add_acc / init_acc (LHMM.py:149-161,256-290) keep
// ksai_acc[3][5] / gamma_acc[3] per unit as log-sum-exp over every (utterance, label position) of
// UNNORMALISED log values (Q6): value = utt_logp + log(expected count).  Magnitudes are 1e4..1e5,
// so this reduction is done in fp64 as (max, sum of exp(value - max)); the split lets a cross-rank
// allreduce(max) / allreduce(sum) sit between the two kernels.
//
// M-step: LHMM.update_param (LHMM.py:509-524) + Clustering.GMM.update_param (Clustering.py:682-693)
// from the linear statistics.
#include "common.cuh"

// PC_TR_CHUNKS blocks per unit: the unit's (utterance, position) pairs are contiguous in the unit-major pair list and are
// cut into PC_TR_CHUNKS equal runs; every block reduces its run (block reduction: a fixed order), leaves the partial in
// the corpus scratch, and the block of a unit that finishes last combines the partials IN CHUNK ORDER - no floating-point
// atomics, and a summation order that depends on the corpus only (the partial sums of a rank are reproducible run to
// run).  One block per unit (round 1) left 57 blocks with 17 500 pairs each at configs[4]: 1.5 ms, and the kernels do not
// overlap the accumulation kernel, whose persistent CTAs leave no shared memory on any SM.
#define TR_THREADS 128

__device__ __forceinline__ double block_reduce(double v, bool is_max, double *sh) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double other = __shfl_xor_sync(0xffffffffu, v, o);
        v = is_max ? fmax(v, other) : v + other;
    }
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    double r = sh[0];
    for (int w = 1; w < TR_THREADS / 32; ++w) r = is_max ? fmax(r, sh[w]) : r + sh[w];
    return r;
}

// the partial of (unit, chunk) is in place: the last block of the unit combines all of them in chunk order
__device__ __forceinline__ void combine_chunks(const CorpusView &v, int unit, bool is_max, double *out) {
    __shared__ int is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = atomicAdd(v.trans_cnt + unit, 1) == PC_TR_CHUNKS - 1;
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    if (threadIdx.x < PC_TRANS_SLOTS) {
        const double *p = v.trans_tmp + (size_t)unit * PC_TR_CHUNKS * PC_TRANS_SLOTS + threadIdx.x;
        double r = is_max ? -INFINITY : 0.0;
        for (int c = 0; c < PC_TR_CHUNKS; ++c) {
            const double x = __ldcg(p + (size_t)c * PC_TRANS_SLOTS);
            r = is_max ? fmax(r, x) : r + x;
        }
        out[(size_t)unit * PC_TRANS_SLOTS + threadIdx.x] = r;
    }
    if (threadIdx.x == 0) v.trans_cnt[unit] = 0;  // clean for the next launch
}

__device__ __forceinline__ void chunk_range(const CorpusView &v, int unit, int chunk, int64_t &lo, int64_t &hi) {
    const int64_t a = v.unit_pair_off[unit], b = v.unit_pair_off[unit + 1];
    const int64_t len = (b - a + PC_TR_CHUNKS - 1) / PC_TR_CHUNKS;
    lo = min(b, a + (int64_t)chunk * len);
    hi = min(b, lo + len);
}

__global__ void __launch_bounds__(TR_THREADS)
transitions_max_kernel(CorpusView v, const double *__restrict__ utt_logp,
                       const float *__restrict__ pair_trans, double *tmax) {
    __shared__ double sh[TR_THREADS / 32];
    const int unit = blockIdx.x, chunk = blockIdx.y;
    int64_t lo, hi;
    chunk_range(v, unit, chunk, lo, hi);
    double m[PC_TRANS_SLOTS];
#pragma unroll
    for (int s = 0; s < PC_TRANS_SLOTS; ++s) m[s] = -INFINITY;
    for (int64_t i = lo + threadIdx.x; i < hi; i += TR_THREADS) {
        const int64_t pair = v.sorted_pair[i];
        const double lp = utt_logp[v.pair_utt[pair]];
#pragma unroll
        for (int s = 0; s < PC_TRANS_SLOTS; ++s) {
            const double val = lp + (double)pair_trans[pair * PC_TRANS_SLOTS + s];
            if (val > m[s]) m[s] = val;  // NaN and -inf never win
        }
    }
    double *part = v.trans_tmp + ((size_t)unit * PC_TR_CHUNKS + chunk) * PC_TRANS_SLOTS;
#pragma unroll
    for (int s = 0; s < PC_TRANS_SLOTS; ++s) {
        const double r = block_reduce(m[s], true, sh);
        if (threadIdx.x == 0) part[s] = r;
    }
    combine_chunks(v, unit, true, tmax);
}

__global__ void __launch_bounds__(TR_THREADS)
transitions_sum_kernel(CorpusView v, const double *__restrict__ utt_logp,
                       const float *__restrict__ pair_trans, const double *__restrict__ tmax,
                       double *tsum) {
    __shared__ double sh[TR_THREADS / 32];
    const int unit = blockIdx.x, chunk = blockIdx.y;
    int64_t lo, hi;
    chunk_range(v, unit, chunk, lo, hi);
    double acc[PC_TRANS_SLOTS], mx[PC_TRANS_SLOTS];
#pragma unroll
    for (int s = 0; s < PC_TRANS_SLOTS; ++s) {
        acc[s] = 0.0;
        mx[s] = tmax[(size_t)unit * PC_TRANS_SLOTS + s];
    }
    for (int64_t i = lo + threadIdx.x; i < hi; i += TR_THREADS) {
        const int64_t pair = v.sorted_pair[i];
        const double lp = utt_logp[v.pair_utt[pair]];
#pragma unroll
        for (int s = 0; s < PC_TRANS_SLOTS; ++s) {
            const double val = lp + (double)pair_trans[pair * PC_TRANS_SLOTS + s];
            if (val == -INFINITY || isnan(val)) continue;
            const double e = exp(val - mx[s]);
            if (e > 0.0) acc[s] += e;
        }
    }
    double *part = v.trans_tmp + ((size_t)unit * PC_TR_CHUNKS + chunk) * PC_TRANS_SLOTS;
#pragma unroll
    for (int s = 0; s < PC_TRANS_SLOTS; ++s) {
        const double r = block_reduce(acc[s], false, sh);
        if (threadIdx.x == 0) part[s] = r;
    }
    combine_chunks(v, unit, false, tsum);
}

// GMM part of the M-step.  One block per state: the occupancies of the state's components are read once into shared
// memory (a thread per (Gaussian, dimension) that summed them itself cost 64 loads per thread at 64 mixtures: 1.4 ms per
// iteration at configs[4]), then one thread per (component, dimension).
__global__ void __launch_bounds__(256)
update_gmm_kernel(int n_units, int mix, int dim, const double *__restrict__ acc, const double *__restrict__ shift,
                  const double *__restrict__ inv_scale, double c_cov, double *mean, double *var, double *alpha) {
    __shared__ double occ_s[64];
    const int64_t state = blockIdx.x;
    const double *a0 = acc + (size_t)state * mix * PC_KA;
    if ((int)threadIdx.x < mix) occ_s[threadIdx.x] = a0[(size_t)threadIdx.x * PC_KA + PC_XS - 1];
    __syncthreads();
    double socc = 0.0;
    for (int m = 0; m < mix; ++m) socc += occ_s[m];  // component order: the same sum in every thread
    if (!(socc > 0.0)) return;  // unseen state: parameters stay (deviation D1)
    for (int i = threadIdx.x; i < mix * dim; i += blockDim.x) {
        const int m = i / dim, d = i - m * dim;
        const int64_t g = state * mix + m;
        const double *a = a0 + (size_t)m * PC_KA;
        const double occ = occ_s[m];
        if (!(occ > 0.0)) {  // component without any posterior mass: weight 0, mean / variance stay (D1)
            if (d == 0) alpha[g] = 0.0;
            continue;
        }
        const double sx = a[d], sxx = a[PC_XS + d];
        const double sh = shift ? shift[d] : 0.0;
        const double is = inv_scale ? inv_scale[d] : 1.0;
        const double mu_old_s = (mean[g * dim + d] - sh) * is;  // old mean in the standardised space of X
        const double mu_s = sx / occ;
        const double var_s = (sxx - 2.0 * mu_old_s * sx + mu_old_s * mu_old_s * occ) / occ;  // Q8: old mean
        double v_new = var_s / (is * is);
        if (v_new < c_cov) v_new = c_cov;  // Clustering.py:690-691
        mean[g * dim + d] = sh + mu_s / is;
        var[g * dim + d] = v_new;
        if (d == 0) alpha[g] = occ / socc;
    }
}

__global__ void update_transmat_kernel(int n_units, const double *__restrict__ tmax,
                                       const double *__restrict__ tsum, double *transmat) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;  // (unit, emitting row r)
    if (i >= n_units * PC_EMIT) return;
    int unit = i / PC_EMIT, r = i - unit * PC_EMIT;
    const double *mx = tmax + (size_t)unit * PC_TRANS_SLOTS + r * 3;
    const double *sm = tsum + (size_t)unit * PC_TRANS_SLOTS + r * 3;
    double g = mx[2] + log(sm[2]);
    if (!(sm[2] > 0.0) || mx[2] == -INFINITY) return;  // unit never observed: keep (deviation D1)
    double ks = (sm[0] > 0.0) ? mx[0] + log(sm[0]) : -INFINITY;
    double kn = (sm[1] > 0.0) ? mx[1] + log(sm[1]) : -INFINITY;
    double *row = transmat + ((size_t)unit * PC_STATES + 1 + r) * PC_STATES;
    for (int c = 0; c < PC_STATES; ++c) row[c] = 0.0;  // exp(-inf - g), LHMM.py:520
    row[1 + r] = exp(ks - g);
    row[2 + r] = exp(kn - g);
}

int launch_transitions_max(pc_handle h, const CorpusView &v, const double *utt_logp,
                           const float *pair_trans, double *tmax, cudaStream_t st) {
    if (v.n_units == 0) return PC_OK;
    transitions_max_kernel<<<dim3(v.n_units, PC_TR_CHUNKS), TR_THREADS, 0, st>>>(v, utt_logp, pair_trans, tmax);
    PC_LAUNCH_CHECK();
    h->launches++;
    return PC_OK;
}

int launch_transitions_sum(pc_handle h, const CorpusView &v, const double *utt_logp,
                           const float *pair_trans, const double *tmax, double *tsum,
                           cudaStream_t st) {
    if (v.n_units == 0) return PC_OK;
    transitions_sum_kernel<<<dim3(v.n_units, PC_TR_CHUNKS), TR_THREADS, 0, st>>>(v, utt_logp, pair_trans, tmax, tsum);
    PC_LAUNCH_CHECK();
    h->launches++;
    return PC_OK;
}

int launch_update_params(pc_handle h, int n_units, int mix, int dim, const double *acc,
                         const double *tmax, const double *tsum, const double *shift,
                         const double *inv_scale, double c_cov, int fix_code, double *mean,
                         double *var, double *alpha, double *transmat, cudaStream_t st) {
    if (n_units == 0) return PC_OK;
    if (!(fix_code & 2)) {
        update_gmm_kernel<<<(unsigned)(n_units * PC_EMIT), 256, 0, st>>>(n_units, mix, dim, acc, shift, inv_scale, c_cov,
                                                                         mean, var, alpha);
        PC_LAUNCH_CHECK();
        h->launches++;
    }
    if (!(fix_code & 4)) {
        int n = n_units * PC_EMIT;
        update_transmat_kernel<<<(n + 127) / 128, 128, 0, st>>>(n_units, tmax, tsum, transmat);
        PC_LAUNCH_CHECK();
        h->launches++;
    }
    return PC_OK;
}
