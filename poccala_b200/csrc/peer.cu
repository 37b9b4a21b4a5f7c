// K6 over peer memory: the cross-rank reduction of the accumulators fused into the M-step.
//
// The reference merges the accumulator files of all workers before update_param (LHMM.py:256-290,
// Clustering.py:314-367; AcousticModel.py:842-882).  With one process per GPU on an NVSwitch node the
// merge needs no collective library: every rank keeps its statistics in an EXCHANGE BLOCK that the other
// ranks map through CUDA IPC, and the M-step kernel of every rank reads the N copies of a state's rows
// straight over NVLink, adds them in rank order (the same order everywhere: replicas stay bit-identical)
// and re-estimates the state.  Per iteration and rank: one signal kernel (a release store of the iteration
// number into every peer's flag array), then the M-step kernels, whose blocks first wait until all N
// flags show the iteration.  No all-reduce, no second pass over the statistics, no host involvement.
//
// Block layout (doubles): [parity 0: acc G*80 | tsum U*9 | tmax U*9] [parity 1: same] [reduced: same]
// then int32 arrive[PC_MAX_PEERS].  The parity alternates per iteration: a rank overwrites buffer p
// again two iterations later, which it can only reach after every peer has signalled the iteration in
// between - i.e. after every peer has finished reading p.  `reduced` holds the summed statistics of the
// last M-step (what linear_stats() / save_acc report).
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
    size_t flag_off;      // byte offset of arrive[]
    int64_t n_acc;        // G * 80
    int n_units;
};

__global__ void peer_signal_kernel(PeerView pv, int epoch) {
    if ((int)threadIdx.x < pv.n) {
        __threadfence_system();
        st_release_sys(reinterpret_cast<int *>(pv.block[threadIdx.x] + pv.flag_off) + pv.rank, epoch);
    }
}

// every block waits for the N arrival flags of this iteration (bounded: a rank that never arrives must not
// hang the device - the kernel then counts a timeout and goes on with whatever it reads)
__device__ __forceinline__ void peer_wait(const PeerView &pv, int epoch, int *timeouts) {
    if ((int)threadIdx.x < pv.n) {
        const int *flag = reinterpret_cast<const int *>(pv.block[pv.rank] + pv.flag_off) + threadIdx.x;
        const long long t0 = clock64();
        while (ld_acquire_sys(flag) < epoch) {
            if (clock64() - t0 > 6000000000ll) {  // ~3 s
                atomicAdd(timeouts, 1);
                break;
            }
            __nanosleep(200);
        }
    }
    __syncthreads();
}

// One block per state: sum the state's mix x 80 statistics over the ranks (rank order), keep the sum in the
// `reduced` set, re-estimate the state (same arithmetic as update_gmm_kernel, reduce.cu).
__global__ void __launch_bounds__(512)
update_gmm_peer_kernel(PeerView pv, int parity, int epoch, int mix, int dim, const double *__restrict__ shift,
                       const double *__restrict__ inv_scale, double c_cov, double *mean, double *var, double *alpha,
                       int *timeouts) {
    extern __shared__ double sm[];  // [mix][80]
    peer_wait(pv, epoch, timeouts);
    const int64_t state = blockIdx.x;
    const size_t base = (size_t)state * mix * PC_KA;
    double *reduced = reinterpret_cast<double *>(pv.block[pv.rank]) + 2 * pv.set_doubles + base;
    // NVLink loads are latency-bound: each thread keeps 4 elements x up to 4 ranks in flight, ranks added in order
    const size_t off = (size_t)parity * pv.set_doubles + base;
    const int n_el = mix * PC_KA;
    for (int i0 = threadIdx.x; i0 < n_el; i0 += 4 * blockDim.x) {
        double s[4] = {0.0, 0.0, 0.0, 0.0};
        for (int r0 = 0; r0 < pv.n; r0 += 4) {
            double v[4][4];
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) {
                const double *src = reinterpret_cast<const double *>(pv.block[min(r0 + rr, pv.n - 1)]) + off;
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
            if (i < n_el) {
                sm[i] = s[e];
                reduced[i] = s[e];
            }
        }
    }
    __syncthreads();
    double socc = 0.0;
    for (int m = 0; m < mix; ++m) socc += sm[m * PC_KA + PC_XS - 1];
    if (!(socc > 0.0)) return;  // unseen state: parameters stay (deviation D1)
    for (int i = threadIdx.x; i < mix * dim; i += blockDim.x) {
        const int m = i / dim, d = i - m * dim;
        const int64_t g = state * mix + m;
        const double *a = sm + m * PC_KA;
        const double occ = a[PC_XS - 1];
        if (!(occ > 0.0)) {
            if (d == 0) alpha[g] = 0.0;
            continue;
        }
        const double sx = a[d], sxx = a[PC_XS + d];
        const double sh = shift ? shift[d] : 0.0;
        const double is = inv_scale ? inv_scale[d] : 1.0;
        const double mu_old_s = (mean[g * dim + d] - sh) * is;
        const double mu_s = sx / occ;
        const double var_s = (sxx - 2.0 * mu_old_s * sx + mu_old_s * mu_old_s * occ) / occ;  // Q8: old mean
        double v_new = var_s / (is * is);
        if (v_new < c_cov) v_new = c_cov;
        mean[g * dim + d] = sh + mu_s / is;
        var[g * dim + d] = v_new;
        if (d == 0) alpha[g] = occ / socc;
    }
}

// Transition accumulators: every rank holds (max_r, sum_r of exp(value - max_r)); the global pair is
// (max over r, sum_r sum_r * exp(max_r - max)).  One thread per (unit, emitting row).
__global__ void update_transmat_peer_kernel(PeerView pv, int parity, int epoch, int update, double *transmat, int *timeouts) {
    peer_wait(pv, epoch, timeouts);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= pv.n_units * PC_EMIT) return;
    const int unit = i / PC_EMIT, r = i - unit * PC_EMIT;
    double *red = reinterpret_cast<double *>(pv.block[pv.rank]) + 2 * pv.set_doubles + pv.n_acc;
    double mx[3], sm[3];
#pragma unroll
    for (int s = 0; s < 3; ++s) {
        const size_t o = (size_t)unit * PC_TRANS_SLOTS + r * 3 + s;
        double gmax = -INFINITY;
        for (int q = 0; q < pv.n; ++q) {
            const double *set = reinterpret_cast<const double *>(pv.block[q]) + (size_t)parity * pv.set_doubles + pv.n_acc;
            const double m = ld_peer(set + (size_t)pv.n_units * PC_TRANS_SLOTS + o);
            if (m > gmax) gmax = m;
        }
        double sum = 0.0;
        for (int q = 0; q < pv.n; ++q) {
            const double *set = reinterpret_cast<const double *>(pv.block[q]) + (size_t)parity * pv.set_doubles + pv.n_acc;
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
    pv.flag_off = 3 * pv.set_doubles * sizeof(double);
    return pv;
}

}  // namespace

int launch_update_params_peer(pc_handle h, int mix, int dim, const double *shift, const double *inv_scale, double c_cov,
                              int fix_code, double *mean, double *var, double *alpha, double *transmat, cudaStream_t st) {
    const PeerView pv = make_view(h);
    const int parity = (int)(h->peer_epoch & 1);
    const int epoch = (int)(++h->peer_epoch);
    peer_signal_kernel<<<1, 32, 0, st>>>(pv, epoch);
    PC_LAUNCH_CHECK();
    const int n_states = pv.n_units * PC_EMIT;
    int *timeouts = h->dev_counters + PC_CNT_PEER_TIMEOUT;
    if (!(fix_code & 2)) {
        const size_t smem = (size_t)mix * PC_KA * sizeof(double);
        PC_CUDA_TRY(cudaFuncSetAttribute(update_gmm_peer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        update_gmm_peer_kernel<<<n_states, 512, smem, st>>>(pv, parity, epoch, mix, dim, shift, inv_scale, c_cov, mean,
                                                             var, alpha, timeouts);
        PC_LAUNCH_CHECK();
        h->launches++;
    }
    update_transmat_peer_kernel<<<(n_states + 127) / 128, 128, 0, st>>>(pv, parity, epoch, (fix_code & 4) ? 0 : 1, transmat,
                                                                       timeouts);
    PC_LAUNCH_CHECK();
    h->launches += 2;
    return PC_OK;
}
