// K6 over peer memory: the cross-rank reduction of the accumulators fused into the M-step.
//
// The reference merges the accumulator files of all workers before update_param (LHMM.py:256-290,
// Clustering.py:314-367; AcousticModel.py:842-882).  With one process per GPU on an NVSwitch node the
// merge needs no collective library: every rank keeps its statistics in an EXCHANGE BLOCK that the other
// ranks map through CUDA IPC.  The M-step is split by OWNER (state s belongs to rank s mod N):
//   1. signal     : a release store of the iteration number into every peer's flag array;
//   2. reduce + M : the owner's blocks wait for all N flags, read the N copies of their states' rows
//                   straight over NVLink, add them in rank order, re-estimate the states and publish the new
//                   parameters (and the summed rows) in the owner's block;
//   3. signal 2, gather : every rank copies the states it does not own from their owners' blocks.
// Per rank that moves (N-1)/N of the statistics once and (N-1)/N of the parameters once - a
// reduce-scatter and an all-gather whose "reduce" is the M-step's own input read.  (A first version let every
// rank add all N copies of everything: N-1 full reads per rank, 41 us slower than NCCL's in-switch all-reduce
// at 8 GPUs.)  Every state is summed by exactly one rank, so replicas are bit-identical by construction.
// The transition accumulators (U x 9 pairs) are tiny: every rank combines all N copies itself.
//
// Block layout (doubles): [parity 0: acc G*80 | tsum U*9 | tmax U*9] [parity 1: same] [reduced: same]
// [published parameters: G rows of 80 = mean (39), alpha | var (39), -] then int32 arrive[PC_MAX_PEERS],
// arrive2[PC_MAX_PEERS].  The parity alternates per iteration: a rank overwrites set p again two
// iterations later, which it can only reach after every peer has signalled the iteration in between -
// i.e. after every peer has finished reading p; the published parameters of iteration k are overwritten in
// step 2 of iteration k+1, behind the wait for every peer's signal of k+1, which a peer sends after its
// gather of k.  `reduced` holds the summed statistics of the last M-step (linear_stats() / save_acc).
#include "common.cuh"

namespace {

__device__ __forceinline__ int ld_acquire_sys(const int *p) {
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(int *p, int v) {
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// statistics of a peer: read past the (incoherent) L1; no ordering of its own - every read sits behind the block's
// arrival wait (acquire load + barrier) - so the compiler is free to keep many of them in flight
__device__ __forceinline__ double ld_peer(const double *p) { return __ldcg(p); }

struct PeerView {
    char *block[PC_MAX_PEERS];
    int n, rank;
    size_t set_doubles;   // doubles per [acc | tsum | tmax] set
    size_t flag_off;      // byte offset of arrive[]; arrive2[] follows
    int64_t n_acc;        // G * 80
    int n_units;
};

__device__ __forceinline__ const double *peer_set(const PeerView &pv, int r, int which) {
    return reinterpret_cast<const double *>(pv.block[r]) + (size_t)which * pv.set_doubles;
}
__device__ __forceinline__ double *own_set(const PeerView &pv, int which) {
    return reinterpret_cast<double *>(pv.block[pv.rank]) + (size_t)which * pv.set_doubles;
}
__device__ __forceinline__ double *own_params(const PeerView &pv) { return own_set(pv, 3); }

__global__ void peer_signal_kernel(PeerView pv, int epoch, int second) {
    if ((int)threadIdx.x < pv.n) {
        __threadfence_system();
        st_release_sys(reinterpret_cast<int *>(pv.block[threadIdx.x] + pv.flag_off) + second * PC_MAX_PEERS + pv.rank, epoch);
    }
}

// wait until the ranks [lo, hi) have signalled this iteration (bounded: a rank that never arrives must not hang
// the device - the kernel then counts a timeout and goes on with whatever it reads)
__device__ __forceinline__ void peer_wait(const PeerView &pv, int epoch, int second, int lo, int hi, int *timeouts) {
    const int t = lo + (int)threadIdx.x;
    if (t < hi) {
        const int *flag = reinterpret_cast<const int *>(pv.block[pv.rank] + pv.flag_off) + second * PC_MAX_PEERS + t;
        const long long t0 = clock64();
        while (ld_acquire_sys(flag) < epoch) {
            if (clock64() - t0 > 6000000000ll) {  // ~3 s
                atomicAdd(timeouts, 1);
                break;
            }
            __nanosleep(100);
        }
    }
    __syncthreads();
}

constexpr int PEER_MPB = 16;  // mixture components per block

// Step 2.  Block = (state, part of <= 16 components), run by the state's owner only: sum the rows over the ranks
// (rank order), keep the sums in the `reduced` set, re-estimate the components (same arithmetic as
// update_gmm_kernel, reduce.cu) into the local model AND the published-parameter rows.
__global__ void __launch_bounds__(512)
update_gmm_peer_kernel(PeerView pv, int parity, int epoch, int mix, int dim, const double *__restrict__ shift,
                       const double *__restrict__ inv_scale, double c_cov, double *mean, double *var, double *alpha,
                       int *timeouts) {
    __shared__ double sm[PEER_MPB * PC_KA];
    __shared__ double occ_all[64];
    const int parts = (mix + PEER_MPB - 1) / PEER_MPB;
    const int64_t state = blockIdx.x / parts;
    const int part = blockIdx.x - (int)state * parts;
    if ((int)(state % pv.n) != pv.rank) return;
    peer_wait(pv, epoch, 0, 0, pv.n, timeouts);
    const int m_lo = part * PEER_MPB, m_n = min(PEER_MPB, mix - m_lo);
    const size_t base = ((size_t)state * mix + m_lo) * PC_KA;
    const int n_el = m_n * PC_KA;
    // NVLink loads are latency-bound: each thread keeps up to 4 elements x 4 ranks in flight, ranks added in order
    for (int i0 = threadIdx.x; i0 < n_el; i0 += 4 * blockDim.x) {
        double s[4] = {0.0, 0.0, 0.0, 0.0};
        for (int r0 = 0; r0 < pv.n; r0 += 4) {
            double v[4][4];
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) {
                const double *src = peer_set(pv, min(r0 + rr, pv.n - 1), parity) + base;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int i = i0 + e * blockDim.x;
                    v[rr][e] = (r0 + rr < pv.n && i < n_el) ? ld_peer(src + i) : 0.0;
                }
            }
#pragma unroll
            for (int rr = 0; rr < 4; ++rr)
#pragma unroll
                for (int e = 0; e < 4; ++e) s[e] += v[rr][e];
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int i = i0 + e * blockDim.x;
            if (i < n_el) sm[i] = s[e];
        }
    }
    // the occupancies of ALL components of the state (the weights' denominator)
    if ((int)threadIdx.x < mix) {
        double s = 0.0;
        for (int r = 0; r < pv.n; ++r)
            s += ld_peer(peer_set(pv, r, parity) + ((size_t)state * mix + threadIdx.x) * PC_KA + PC_XS - 1);
        occ_all[threadIdx.x] = s;
    }
    __syncthreads();
    double *reduced = own_set(pv, 2) + base;
    for (int i = threadIdx.x; i < n_el; i += blockDim.x) reduced[i] = sm[i];
    double socc = 0.0;
    for (int m = 0; m < mix; ++m) socc += occ_all[m];
    double *pub = own_params(pv) + base;
    for (int i = threadIdx.x; i < m_n * PC_XS; i += blockDim.x) {
        const int m = i / PC_XS, d = i - m * PC_XS;  // d == 39: the weight
        const int64_t g = state * mix + m_lo + m;
        const double *a = sm + m * PC_KA;
        const double occ = a[PC_XS - 1];
        double mu = 0.0, vv = 0.0, al = 0.0;
        if (d < dim) { mu = mean[g * dim + d]; vv = var[g * dim + d]; }
        if (d == PC_XS - 1) al = alpha[g];
        if (socc > 0.0) {  // (an unseen state keeps its parameters, deviation D1)
            if (!(occ > 0.0)) {
                if (d == PC_XS - 1) al = 0.0;  // component without posterior mass: weight 0, mean / variance stay
            } else if (d < dim) {
                const double sx = a[d], sxx = a[PC_XS + d];
                const double sh = shift ? shift[d] : 0.0;
                const double is = inv_scale ? inv_scale[d] : 1.0;
                const double mu_old_s = (mu - sh) * is;
                const double mu_s = sx / occ;
                const double var_s = (sxx - 2.0 * mu_old_s * sx + mu_old_s * mu_old_s * occ) / occ;  // Q8: old mean
                vv = var_s / (is * is);
                if (vv < c_cov) vv = c_cov;
                mu = sh + mu_s / is;
            } else if (d == PC_XS - 1) {
                al = occ / socc;
            }
        }
        if (d < dim) {
            mean[g * dim + d] = mu;
            var[g * dim + d] = vv;
            pub[m * PC_KA + d] = mu;
            pub[m * PC_KA + PC_XS + d] = vv;
        } else if (d == PC_XS - 1) {
            alpha[g] = al;
            pub[m * PC_KA + PC_XS - 1] = al;
        }
    }
}

// Step 3.  The states of other owners: new parameters and summed rows from the owner's block.
__global__ void __launch_bounds__(512)
gather_params_peer_kernel(PeerView pv, int epoch, int mix, int dim, double *mean, double *var, double *alpha,
                          int *timeouts) {
    const int parts = (mix + PEER_MPB - 1) / PEER_MPB;
    const int64_t state = blockIdx.x / parts;
    const int part = blockIdx.x - (int)state * parts;
    const int owner = (int)(state % pv.n);
    if (owner == pv.rank) return;
    peer_wait(pv, epoch, 1, owner, owner + 1, timeouts);
    const int m_lo = part * PEER_MPB, m_n = min(PEER_MPB, mix - m_lo);
    const size_t base = ((size_t)state * mix + m_lo) * PC_KA;
    const double *pub = peer_set(pv, owner, 3) + base, *red = peer_set(pv, owner, 2) + base;
    double *reduced = own_set(pv, 2) + base;
    for (int i = threadIdx.x; i < m_n * PC_KA; i += blockDim.x) {
        const double p = ld_peer(pub + i), r = ld_peer(red + i);
        reduced[i] = r;
        const int m = i / PC_KA, c = i - m * PC_KA;
        const int64_t g = state * mix + m_lo + m;
        if (c < dim) mean[g * dim + c] = p;
        else if (c == PC_XS - 1) alpha[g] = p;
        else if (c >= PC_XS && c - PC_XS < dim) var[g * dim + c - PC_XS] = p;
    }
}

// Transition accumulators: every rank holds (max_r, sum_r of exp(value - max_r)); the global pair is
// (max over r, sum_r sum_r * exp(max_r - max)).  One thread per (unit, emitting row); every rank combines all N.
__global__ void update_transmat_peer_kernel(PeerView pv, int parity, int epoch, int update, double *transmat, int *timeouts) {
    peer_wait(pv, epoch, 0, 0, pv.n, timeouts);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= pv.n_units * PC_EMIT) return;
    const int unit = i / PC_EMIT, r = i - unit * PC_EMIT;
    double *red = own_set(pv, 2) + pv.n_acc;
    double mx[3], sm[3];
#pragma unroll
    for (int s = 0; s < 3; ++s) {
        const size_t o = (size_t)unit * PC_TRANS_SLOTS + r * 3 + s;
        double gmax = -INFINITY;
        for (int q = 0; q < pv.n; ++q) {
            const double m = ld_peer(peer_set(pv, q, parity) + pv.n_acc + (size_t)pv.n_units * PC_TRANS_SLOTS + o);
            if (m > gmax) gmax = m;
        }
        double sum = 0.0;
        for (int q = 0; q < pv.n; ++q) {
            const double *set = peer_set(pv, q, parity) + pv.n_acc;
            const double m = ld_peer(set + (size_t)pv.n_units * PC_TRANS_SLOTS + o);
            const double t = ld_peer(set + o);
            if (m > -INFINITY && t > 0.0) sum += t * exp(m - gmax);
        }
        mx[s] = gmax;
        sm[s] = sum;
        red[o] = sum;
        red[(size_t)pv.n_units * PC_TRANS_SLOTS + o] = gmax;
    }
    if (!update) return;
    const double g = mx[2] + log(sm[2]);
    if (!(sm[2] > 0.0) || mx[2] == -INFINITY) return;  // unit never observed: keep (deviation D1)
    const double ks = (sm[0] > 0.0) ? mx[0] + log(sm[0]) : -INFINITY;
    const double kn = (sm[1] > 0.0) ? mx[1] + log(sm[1]) : -INFINITY;
    double *row = transmat + ((size_t)unit * PC_STATES + 1 + r) * PC_STATES;
    for (int c = 0; c < PC_STATES; ++c) row[c] = 0.0;
    row[1 + r] = exp(ks - g);
    row[2 + r] = exp(kn - g);
}

PeerView make_view(pc_handle h) {
    PeerView pv;
    for (int r = 0; r < PC_MAX_PEERS; ++r) pv.block[r] = r < h->peer_n ? (char *)h->peer_block[r] : nullptr;
    pv.n = h->peer_n;
    pv.rank = h->peer_rank;
    pv.n_acc = h->peer_n_acc;
    pv.n_units = h->peer_n_units;
    pv.set_doubles = (size_t)h->peer_n_acc + 2 * (size_t)h->peer_n_units * PC_TRANS_SLOTS;
    pv.flag_off = 4 * pv.set_doubles * sizeof(double);
    return pv;
}

}  // namespace

int launch_update_params_peer(pc_handle h, int mix, int dim, const double *shift, const double *inv_scale, double c_cov,
                              int fix_code, double *mean, double *var, double *alpha, double *transmat, cudaStream_t st) {
    const PeerView pv = make_view(h);
    const int parity = (int)(h->peer_epoch & 1);
    const int epoch = (int)(++h->peer_epoch);
    peer_signal_kernel<<<1, 32, 0, st>>>(pv, epoch, 0);
    PC_LAUNCH_CHECK();
    const int n_states = pv.n_units * PC_EMIT;
    const int parts = (mix + PEER_MPB - 1) / PEER_MPB;
    int *timeouts = h->dev_counters + PC_CNT_PEER_TIMEOUT;
    if (!(fix_code & 2)) {
        update_gmm_peer_kernel<<<n_states * parts, 512, 0, st>>>(pv, parity, epoch, mix, dim, shift, inv_scale, c_cov, mean,
                                                                  var, alpha, timeouts);
        PC_LAUNCH_CHECK();
        h->launches++;
    }
    update_transmat_peer_kernel<<<(n_states + 127) / 128, 128, 0, st>>>(pv, parity, epoch, (fix_code & 4) ? 0 : 1, transmat,
                                                                       timeouts);
    PC_LAUNCH_CHECK();
    h->launches += 2;
    if (!(fix_code & 2) && pv.n > 1) {
        peer_signal_kernel<<<1, 32, 0, st>>>(pv, epoch, 1);
        PC_LAUNCH_CHECK();
        gather_params_peer_kernel<<<n_states * parts, 512, 0, st>>>(pv, epoch, mix, dim, mean, var, alpha, timeouts);
        PC_LAUNCH_CHECK();
        h->launches += 2;
    }
    return PC_OK;
}
