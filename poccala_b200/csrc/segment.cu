// Alignment post-processing: the steps either side of Viterbi / k-means in the
// reference's mode-1 training, as device kernels over the resident corpus.
//
//   segment_keys_kernel  per frame -> (unit, emitting state) it trains, or -1 (dropped)
//       mode 0: AcousticModel.__eq_segment(mode='e') (AcousticModel.py:605-612): an utterance of T
//               frames and L labels is cut into L chunks of T // L frames (the remainder is dropped)
//       mode 1: multi_process_data after Viterbi (AcousticModel.py:750-764): the per-frame unit
//               sequence is cut into maximal runs of one unit (discriminate, :937-955); an
//               utterance whose path visits fewer distinct units than its label holds is dropped
//       both:   every segment of n frames is cut into 3 parts of n // 3 frames, the last part
//               taking the remainder (__eq_segment(mode='g') :613-626 via __get_gmmdata :630-644)
//   key_hist / key_scan / key_scatter  stable counting sort of the frames by key: the frames of
//       one (unit, state) end up contiguous, in (utterance, time) order - the concatenation
//       __get_gmmdata builds with np.append
//   gather_rows_kernel   copies the rows in that order (the per-state data sets k-means and
//       GMM.em consume)
//
// All of it is HBM-bound index work: 4 B read + 4 B written per frame for the keys, ~12 B per frame
// for the sort, 2 x row bytes per kept frame for the gather.
#include <limits.h>

#include "common.cuh"

namespace {

constexpr int SEG_WARPS = 4;
constexpr int GRP_THREADS = 256;
constexpr int GRP_CHUNK = 4096;  // frames per block of the counting sort

__device__ __forceinline__ int state_of_part(int off, int n) {
    int c = n / PC_EMIT;
    if (c == 0) return PC_EMIT - 1;  // n < 3: the first parts are empty slices, the last takes all
    int r = off / c;
    return r < PC_EMIT - 1 ? r : PC_EMIT - 1;
}

__global__ void __launch_bounds__(SEG_WARPS * 32)
segment_keys_kernel(CorpusView v, int mode, const int32_t *__restrict__ path, int32_t *__restrict__ key,
                    int32_t *__restrict__ kept) {
    extern __shared__ int32_t seg_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int u = blockIdx.x * SEG_WARPS + warp;
    if (u >= v.n_utt) return;
    const int64_t f0 = v.frame_off[u];
    const int T = (int)(v.frame_off[u + 1] - f0);
    const int64_t p0 = v.pair_off[u];
    const int L = (int)(v.pair_off[u + 1] - p0);
    const int32_t *lab = v.labels + p0;
    int32_t *out = key + f0;
    if (L == 0) {
        for (int t = lane; t < T; t += 32) out[t] = -1;
        if (lane == 0 && kept) kept[u] = 0;
        return;
    }
    if (mode == 0) {
        const int chunk = L > 0 ? T / L : 0;
        for (int t = lane; t < T; t += 32) {
            int k = -1;
            if (chunk > 0) {
                int p = t / chunk;
                if (p < L) k = lab[p] * PC_EMIT + state_of_part(t - p * chunk, chunk);
            }
            out[t] = k;
        }
        if (lane == 0 && kept) kept[u] = 1;
        return;
    }
    // ---- mode 1: runs of one unit in the aligned sequence
    int32_t *visited = seg_smem + warp * v.max_labels;
    const int32_t *pth = path + f0;
    auto pos_of = [&](int t) {
        int s = pth[t];
        int p = s <= 0 ? 0 : (s - 1) / PC_EMIT;
        return p < L ? p : L - 1;  // the exit state belongs to the last unit (AcousticModel.py:968-976)
    };
    for (int p = lane; p < L; p += 32) visited[p] = 0;
    __syncwarp();
    // pass 1: run start of every frame (kept in the output array until pass 2 replaces it)
    int carry = 0;
    for (int tb = 0; tb < T; tb += 32) {
        int t = tb + lane;
        int start = -1;
        if (t < T) {
            int p = pos_of(t);
            visited[p] = 1;
            if (t == 0 || lab[p] != lab[pos_of(t - 1)]) start = t;
        }
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int n = __shfl_up_sync(0xffffffffu, start, o);
            if (lane >= o) start = max(start, n);
        }
        start = max(start, carry);
        if (t < T) out[t] = start;
        carry = __shfl_sync(0xffffffffu, start, 31);
    }
    __syncwarp();
    // fewer distinct units on the path than in the label: the alignment failed, drop the utterance
    int n_all = 0, n_vis = 0;
    for (int pb = 0; pb < L; pb += 32) {
        int p = pb + lane;
        bool first_all = false, first_vis = false;
        if (p < L) {
            first_all = true;
            first_vis = visited[p] != 0;
            for (int q = 0; q < p; ++q)
                if (lab[q] == lab[p]) {
                    first_all = false;
                    if (visited[q]) first_vis = false;
                }
        }
        n_all += __popc(__ballot_sync(0xffffffffu, first_all));
        n_vis += __popc(__ballot_sync(0xffffffffu, first_vis));
    }
    const bool ok = n_vis >= n_all;
    if (lane == 0 && kept) kept[u] = ok ? 1 : 0;
    // pass 2 (backwards): run end of every frame, then the key
    carry = T;
    for (int tb = ((T - 1) / 32) * 32; tb >= 0; tb -= 32) {
        int t = tb + lane;
        int end = INT_MAX;
        int unit = -1, start = 0;
        if (t < T) {
            unit = lab[pos_of(t)];
            start = out[t];
            if (t + 1 < T && lab[pos_of(t + 1)] != unit) end = t + 1;
        }
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int n = __shfl_down_sync(0xffffffffu, end, o);
            if (lane + o < 32) end = min(end, n);
        }
        end = min(end, carry);
        carry = __shfl_sync(0xffffffffu, end, 0);
        if (t < T) out[t] = ok ? unit * PC_EMIT + state_of_part(t - start, end - start) : -1;
    }
}

__device__ __forceinline__ int bucket_of(int k, int n_keys) { return (k < 0 || k >= n_keys) ? n_keys : k; }

// hist[k][block] = frames of the block's chunk carrying key k (bucket n_keys = dropped frames)
__global__ void __launch_bounds__(GRP_THREADS)
key_hist_kernel(const int32_t *__restrict__ key, int64_t n, int n_keys, int32_t *__restrict__ hist) {
    extern __shared__ int32_t cnt[];
    for (int k = threadIdx.x; k <= n_keys; k += GRP_THREADS) cnt[k] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * GRP_CHUNK;
    for (int i = threadIdx.x; i < GRP_CHUNK; i += GRP_THREADS) {
        int64_t f = base + i;
        if (f < n) atomicAdd(&cnt[bucket_of(key[f], n_keys)], 1);
    }
    __syncthreads();
    for (int k = threadIdx.x; k <= n_keys; k += GRP_THREADS) hist[(size_t)k * gridDim.x + blockIdx.x] = cnt[k];
}

template <typename T>
__device__ __forceinline__ T block_exclusive_scan(T v, T *total, T *warp_tot /* [GRP_THREADS/32] shared */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        T n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += n;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    T before = 0, all = 0;
    for (int w = 0; w < GRP_THREADS / 32; ++w) {
        T x = warp_tot[w];
        if (w < warp) before += x;
        all += x;
    }
    __syncthreads();
    *total = all;
    return before + inc - v;
}

// one block per key: exclusive prefix of the key's row over the chunks (in place), row total out
__global__ void __launch_bounds__(GRP_THREADS)
key_scan_rows_kernel(int32_t *__restrict__ hist, int n_blocks, int64_t *__restrict__ key_total) {
    __shared__ int32_t wt[GRP_THREADS / 32];
    int32_t *row = hist + (size_t)blockIdx.x * n_blocks;
    int32_t carry = 0;
    for (int b0 = 0; b0 < n_blocks; b0 += GRP_THREADS) {
        int b = b0 + threadIdx.x;
        int32_t v = b < n_blocks ? row[b] : 0, tot;
        int32_t ex = block_exclusive_scan<int32_t>(v, &tot, wt);
        if (b < n_blocks) row[b] = carry + ex;
        carry += tot;
    }
    if (threadIdx.x == 0) key_total[blockIdx.x] = carry;
}

// key_off[k] = first slot of key k; key_off[n_keys] = kept frames; key_off[n_keys + 1] = all frames
__global__ void __launch_bounds__(GRP_THREADS)
key_scan_totals_kernel(const int64_t *__restrict__ key_total, int n_keys, int64_t *__restrict__ key_off) {
    __shared__ int64_t wt[GRP_THREADS / 32];
    int64_t carry = 0;
    for (int k0 = 0; k0 <= n_keys; k0 += GRP_THREADS) {
        int k = k0 + threadIdx.x;
        int64_t v = k <= n_keys ? key_total[k] : 0, tot;
        int64_t ex = block_exclusive_scan<int64_t>(v, &tot, wt);
        if (k <= n_keys) key_off[k] = carry + ex;
        carry += tot;
    }
    if (threadIdx.x == 0) key_off[n_keys + 1] = carry;
}

// order[slot] = frame, slots of one key in ascending frame order (stable)
__global__ void __launch_bounds__(GRP_THREADS)
key_scatter_kernel(const int32_t *__restrict__ key, int64_t n, int n_keys, const int32_t *__restrict__ hist,
                   const int64_t *__restrict__ key_off, int32_t *__restrict__ order) {
    extern __shared__ int32_t cnt[];
    for (int k = threadIdx.x; k <= n_keys; k += GRP_THREADS)
        cnt[k] = (int32_t)key_off[k] + hist[(size_t)k * gridDim.x + blockIdx.x];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t base = (int64_t)blockIdx.x * GRP_CHUNK;
    for (int i0 = 0; i0 < GRP_CHUNK; i0 += GRP_THREADS) {
        if (base + i0 >= n) break;  // uniform over the block
        const int64_t f = base + i0 + threadIdx.x;
        const bool valid = f < n;
        const int k = valid ? bucket_of(key[f], n_keys) : n_keys + 1 + lane;  // idle lanes match nobody
        const unsigned peers = __match_any_sync(0xffffffffu, k);
        const int rank = __popc(peers & ((1u << lane) - 1u));
        for (int w = 0; w < GRP_THREADS / 32; ++w) {  // warps take their slots in frame order
            if (warp == w && valid) {
                int slot = cnt[k] + rank;
                __syncwarp(peers);
                if (rank == 0) cnt[k] += __popc(peers);
                order[slot] = (int32_t)f;
            }
            __syncthreads();
        }
    }
}

// W = 8- or 4-byte words; 32-bit index arithmetic (the launcher slices to < 2^31 words),
// four independent row fetches in flight per thread
template <typename W>
__global__ void __launch_bounds__(256)
gather_rows_kernel(const int32_t *__restrict__ order, uint32_t n_rows, uint32_t words, const W *__restrict__ src,
                   W *__restrict__ dst) {
    const uint32_t total = n_rows * words;
    const uint32_t stride = gridDim.x * blockDim.x;
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    for (; i < total && total - i > 3 * stride; i += 4 * stride) {
        W v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint32_t e = i + j * stride, row = e / words;
            v[j] = src[(size_t)order[row] * words + (e - row * words)];
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) dst[i + j * stride] = v[j];
    }
    for (; i < total; i += stride) {
        uint32_t row = i / words;
        dst[i] = src[(size_t)order[row] * words + (i - row * words)];
    }
}

}  // namespace

int launch_segment_keys(pc_handle h, const CorpusView &v, int mode, const int32_t *path, int32_t *key,
                        int32_t *kept, cudaStream_t st) {
    if (v.n_utt == 0) return PC_OK;
    size_t smem = (size_t)SEG_WARPS * (v.max_labels > 0 ? v.max_labels : 1) * sizeof(int32_t);
    segment_keys_kernel<<<(v.n_utt + SEG_WARPS - 1) / SEG_WARPS, SEG_WARPS * 32, smem, st>>>(v, mode, path, key, kept);
    PC_LAUNCH_CHECK();
    h->launches++;
    return PC_OK;
}

int64_t group_workspace_bytes(int64_t n_frames, int n_keys) {
    int64_t nb = (n_frames + GRP_CHUNK - 1) / GRP_CHUNK;
    if (nb < 1) nb = 1;
    return (int64_t)(n_keys + 1) * 8 + (int64_t)(n_keys + 1) * nb * 4;
}

int launch_group_frames(pc_handle h, const int32_t *key, int64_t n, int n_keys, void *ws, int64_t *key_off,
                        int32_t *order, cudaStream_t st) {
    int nb = (int)((n + GRP_CHUNK - 1) / GRP_CHUNK);
    if (nb < 1) nb = 1;
    int64_t *key_total = (int64_t *)ws;
    int32_t *hist = (int32_t *)((char *)ws + (size_t)(n_keys + 1) * 8);
    size_t smem = (size_t)(n_keys + 1) * sizeof(int32_t);
    key_hist_kernel<<<nb, GRP_THREADS, smem, st>>>(key, n, n_keys, hist);
    PC_LAUNCH_CHECK();
    key_scan_rows_kernel<<<n_keys + 1, GRP_THREADS, 0, st>>>(hist, nb, key_total);
    PC_LAUNCH_CHECK();
    key_scan_totals_kernel<<<1, GRP_THREADS, 0, st>>>(key_total, n_keys, key_off);
    PC_LAUNCH_CHECK();
    key_scatter_kernel<<<nb, GRP_THREADS, smem, st>>>(key, n, n_keys, hist, key_off, order);
    PC_LAUNCH_CHECK();
    h->launches += 4;
    return PC_OK;
}

int launch_gather_rows(pc_handle h, const int32_t *order, int64_t n_rows, int row_bytes, const void *src, void *dst,
                       cudaStream_t st) {
    if (n_rows == 0) return PC_OK;
    const bool wide = row_bytes % 8 == 0 && ((uintptr_t)src | (uintptr_t)dst) % 8 == 0;
    const int words = row_bytes / (wide ? 8 : 4);
    // slices of < 2^31 words per launch
    const int64_t rows_per_launch = ((1ll << 31) - 1) / words;  // 32-bit index arithmetic with headroom
    for (int64_t r0 = 0; r0 < n_rows; r0 += rows_per_launch) {
        int64_t nr = n_rows - r0 < rows_per_launch ? n_rows - r0 : rows_per_launch;
        int64_t want = (nr * words + 1023) / 1024;
        int grid = (int)(want < (int64_t)h->sm_count * 8 ? want : (int64_t)h->sm_count * 8);
        char *d = (char *)dst + (size_t)r0 * row_bytes;
        if (wide)
            gather_rows_kernel<uint64_t><<<grid, 256, 0, st>>>(order + r0, (uint32_t)nr, (uint32_t)words,
                                                               (const uint64_t *)src, (uint64_t *)d);
        else
            gather_rows_kernel<uint32_t><<<grid, 256, 0, st>>>(order + r0, (uint32_t)nr, (uint32_t)words,
                                                               (const uint32_t *)src, (uint32_t *)d);
        PC_LAUNCH_CHECK();
        h->launches++;
    }
    return PC_OK;
}
