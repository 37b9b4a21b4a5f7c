// C ABI (include/poccala_b200.h): handle, corpus descriptors, argument checking, dispatch.
#include <stdarg.h>
#include <string.h>

#include <algorithm>
#include <numeric>
#include <vector>

#include "common.cuh"

static thread_local char g_err[512] = "";

const char *pc_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return g_err;
}

extern "C" {

int pc_abi_version(void) { return PC_ABI_VERSION; }
int64_t pc_rows_bytes(int64_t n) { return n < 0 ? -1 : n * PC_XS * 4; }
int64_t pc_gmm_bytes(int64_t n) { return n < 0 ? -1 : n * PC_W_BYTES_PER_GAUSS + 8192; }
const char *pc_last_error(void) { return g_err; }

int pc_create(int device, pc_handle *out) {
    PC_REQUIRE(out != nullptr, "pc_create: out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        pc_set_error("pc_create: no CUDA device (%s); this engine has no CPU fallback",
                     e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        return PC_ERR_CUDA;
    }
    PC_REQUIRE(device >= 0 && device < n, "pc_create: device %d out of range [0,%d)", device, n);
    cudaDeviceProp prop;
    PC_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        pc_set_error("pc_create: device %d is sm_%d%d; the kernels are built for sm_100a only",
                     device, prop.major, prop.minor);
        return PC_ERR_CUDA;
    }
    PC_CUDA_TRY(cudaSetDevice(device));
    pc_handle h = new pc_handle_s();
    memset(h, 0, sizeof(*h));
    h->device = device;
    h->sm_count = prop.multiProcessorCount;
    h->use_tc = 1;
    h->k2_kernel = 1;
    h->k1_kernel = 1;
    h->k3_kernel = 1;
    h->kmeans_cluster = 1;
    if (cudaMalloc((void **)&h->dev_counters, PC_CNT_N * sizeof(int)) != cudaSuccess ||
        cudaMemset(h->dev_counters, 0, PC_CNT_N * sizeof(int)) != cudaSuccess) {
        pc_set_error("pc_create: device allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
        delete h;
        return PC_ERR_CUDA;
    }
    *out = h;
    return PC_OK;
}

int pc_destroy(pc_handle h) {
    if (!h) return PC_OK;
    cudaSetDevice(h->device);
    pc_peer_destroy(h);
    if (h->ws) cudaFree(h->ws);
    if (h->dev_counters) cudaFree(h->dev_counters);
    if (h->pinned) cudaFreeHost(h->pinned);
    if (h->copy_stream) {
        cudaStreamDestroy(h->copy_stream);
        cudaEventDestroy(h->start_ev);
        cudaEventDestroy(h->fork_ev);
        cudaEventDestroy(h->join_ev);
        for (int k = 0; k < PC_MAX_CHUNKS; ++k) cudaEventDestroy(h->chunk_ev[k]);
    }
    delete h;
    return PC_OK;
}

int pc_set_option(pc_handle h, const char *key, int64_t value) {
    PC_REQUIRE(h && key, "pc_set_option: NULL argument");
    if (!strcmp(key, "tensor_core")) { h->use_tc = (int)value; return PC_OK; }
    if (!strcmp(key, "debug_flags")) { h->debug_flags = (int)value; return PC_OK; }
    if (!strcmp(key, "host_chunks")) { h->host_chunks = (int)value; return PC_OK; }
    if (!strcmp(key, "launches")) { h->launches = value; return PC_OK; }
    if (!strcmp(key, "k2_kernel")) { h->k2_kernel = (int)value; return PC_OK; }
    if (!strcmp(key, "k1_kernel")) { h->k1_kernel = (int)value; return PC_OK; }
    if (!strcmp(key, "k3_kernel")) { h->k3_kernel = (int)value; return PC_OK; }
    if (!strcmp(key, "kmeans_cluster")) { h->kmeans_cluster = (int)value; return PC_OK; }
    pc_set_error("pc_set_option: unknown key '%s'", key);
    return PC_ERR_INVALID;
}

int64_t pc_get_option(pc_handle h, const char *key) {
    if (!h || !key) return -1;
    if (!strcmp(key, "tensor_core")) return h->use_tc;
    if (!strcmp(key, "debug_flags")) return h->debug_flags;
    if (!strcmp(key, "host_chunks")) return h->host_chunks;
    if (!strcmp(key, "launches")) return h->launches;
    if (!strcmp(key, "sm_count")) return h->sm_count;
    if (!strcmp(key, "k2_kernel")) return h->k2_kernel;
    if (!strcmp(key, "k1_kernel")) return h->k1_kernel;
    if (!strcmp(key, "k3_kernel")) return h->k3_kernel;
    if (!strcmp(key, "kmeans_cluster")) return h->kmeans_cluster;
    if (!strcmp(key, "peer_epoch")) return h->peer_epoch;
    if (!strcmp(key, "clamped") || !strcmp(key, "peer_timeouts")) {
        const int idx = !strcmp(key, "clamped") ? PC_CNT_CLAMPED : PC_CNT_PEER_TIMEOUT;
        int n = 0;
        if (cudaSetDevice(h->device) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess ||
            cudaMemcpy(&n, h->dev_counters + idx, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess)
            return -1;
        if (n != 0) cudaMemset(h->dev_counters + idx, 0, sizeof(int));
        return n;
    }
    return -1;
}

// --------------------------------------------------------------------------------- corpus
int pc_corpus_create(pc_handle h, int32_t n_utt, const int32_t *n_frames, const int32_t *n_labels,
                     const int32_t *labels, int32_t n_units, pc_corpus *out) {
    PC_REQUIRE(h && out, "pc_corpus_create: NULL handle/out");
    PC_REQUIRE(n_utt >= 0 && n_units > 0, "pc_corpus_create: n_utt=%d n_units=%d", n_utt, n_units);
    PC_REQUIRE(n_utt == 0 || (n_frames && n_labels && labels), "pc_corpus_create: NULL host array");
    *out = nullptr;
    std::vector<int64_t> frame_off(n_utt + 1, 0), emis_off(n_utt + 1, 0), pair_off(n_utt + 1, 0),
        state_off(n_utt + 1, 0);
    int max_frames = 0, max_labels = 0;
    for (int u = 0; u < n_utt; ++u) {
        PC_REQUIRE(n_frames[u] >= 1, "pc_corpus_create: utterance %d has %d frames", u, n_frames[u]);
        PC_REQUIRE(n_labels[u] >= 1, "pc_corpus_create: utterance %d has %d labels", u, n_labels[u]);
        frame_off[u + 1] = frame_off[u] + n_frames[u];
        pair_off[u + 1] = pair_off[u] + n_labels[u];
        emis_off[u + 1] = emis_off[u] + (int64_t)n_frames[u] * pc_spad(n_labels[u]);
        state_off[u + 1] = state_off[u] + PC_EMIT * n_labels[u] + 2;
        max_frames = std::max(max_frames, n_frames[u]);
        max_labels = std::max(max_labels, n_labels[u]);
    }
    std::vector<int64_t> xtile_off(n_utt + 1, 0);
    std::vector<int32_t> xtile_utt, xtile_t0;
    for (int u = 0; u < n_utt; ++u) {
        for (int t0 = 0; t0 < n_frames[u]; t0 += PC_TILE_ROWS) {
            xtile_utt.push_back(u);
            xtile_t0.push_back(t0);
        }
        xtile_off[u + 1] = (int64_t)xtile_utt.size();
    }
    const int64_t n_xtiles = xtile_off[n_utt];
    const int64_t n_pairs = pair_off[n_utt];
    std::vector<int32_t> pair_utt(n_pairs);
    for (int u = 0; u < n_utt; ++u)
        for (int64_t p = pair_off[u]; p < pair_off[u + 1]; ++p) {
            PC_REQUIRE(labels[p] >= 0 && labels[p] < n_units,
                       "pc_corpus_create: label %d of utterance %d is outside [0,%d)", labels[p], u,
                       n_units);
            pair_utt[p] = u;
        }
    // forward-backward order: heaviest utterances first
    std::vector<int32_t> fb_order(n_utt);
    std::iota(fb_order.begin(), fb_order.end(), 0);
    std::stable_sort(fb_order.begin(), fb_order.end(), [&](int a, int b) {
        int64_t wa = (int64_t)n_frames[a] * (PC_EMIT * n_labels[a] + 1);
        int64_t wb = (int64_t)n_frames[b] * (PC_EMIT * n_labels[b] + 1);
        return wa > wb;
    });
    // unit-major decomposition: counting sort of the pairs by unit (stable)
    std::vector<int64_t> unit_pair_off(n_units + 1, 0);
    for (int64_t p = 0; p < n_pairs; ++p) unit_pair_off[labels[p] + 1]++;
    for (int k = 0; k < n_units; ++k) unit_pair_off[k + 1] += unit_pair_off[k];
    std::vector<int64_t> sorted_pair(n_pairs);
    {
        std::vector<int64_t> cur(unit_pair_off.begin(), unit_pair_off.end() - 1);
        for (int64_t p = 0; p < n_pairs; ++p) sorted_pair[cur[labels[p]]++] = p;
    }
    // K3 walks (tile, unit) pairs unit-major so that a unit's statistics stay in tensor memory, but
    // every frame tile is then fetched once per label position.  The walk is therefore split into
    // runs of consecutive utterances whose frame-tile images (40 KiB each) fit in a quarter of the
    // L2 (PC_L2_RUN_BYTES): inside a run the re-reads hit L2 instead of HBM.
    std::vector<int64_t> tile_pair, tile_xrow, tile_boff, tile_xblk;
    std::vector<int32_t> tile_t0, tile_rows, tile_tp;
    std::vector<int64_t> run_tile_off;  // first tile of each (run, unit) block, then n_tiles
    std::vector<int64_t> pair_tile0(n_pairs, 0);
    {
        std::vector<int> run_utt(1, 0);
        int64_t bytes = 0;
        for (int u = 0; u < n_utt; ++u) {
            const int64_t add_b = (xtile_off[u + 1] - xtile_off[u]) * (int64_t)PC_XTILE_BYTES;
            if (bytes > 0 && bytes + add_b > PC_L2_RUN_BYTES) {
                run_utt.push_back(u);
                bytes = 0;
            }
            bytes += add_b;
        }
        run_utt.push_back(n_utt);
        const int n_runs = (int)run_utt.size() - 1;
        // sorted_pair is stable: inside a unit the pairs are in utterance order, so each run is a
        // contiguous slice of the unit's pair list
        std::vector<int64_t> cur(unit_pair_off.begin(), unit_pair_off.end() - 1);
        for (int r = 0; r < n_runs; ++r) {
            for (int k = 0; k < n_units; ++k) {
                run_tile_off.push_back((int64_t)tile_pair.size());
                int64_t &i = cur[k];
                for (; i < unit_pair_off[k + 1] && pair_utt[sorted_pair[i]] < run_utt[r + 1]; ++i) {
                    const int64_t p = sorted_pair[i];
                    const int u = pair_utt[p];
                    const int T = n_frames[u];
                    const int tp = pc_spad(n_labels[u]);
                    const int64_t pos = p - pair_off[u];
                    pair_tile0[p] = (int64_t)tile_pair.size();
                    for (int t0 = 0; t0 < T; t0 += PC_TILE_ROWS) {
                        tile_pair.push_back(p);
                        tile_t0.push_back(t0);
                        tile_rows.push_back(std::min(PC_TILE_ROWS, T - t0));
                        tile_tp.push_back(tp);
                        tile_xrow.push_back(frame_off[u] + t0);
                        tile_xblk.push_back(xtile_off[u] + t0 / PC_TILE_ROWS);
                        tile_boff.push_back(emis_off[u] + (int64_t)t0 * tp + PC_EMIT * pos);
                    }
                }
            }
        }
        run_tile_off.push_back((int64_t)tile_pair.size());
    }
    const int64_t n_tiles = (int64_t)tile_pair.size();
    // work items: runs of <= chunk tiles inside one unit; aim at >= 8 items per SM
    int64_t chunk = n_tiles / ((int64_t)h->sm_count * 8);
    // <= 64 tiles: the accumulation kernel keeps an item's sums in fp32 (TMEM) before the fp64 flush
    // and addresses an item's tiles through two 32-bit activity masks (typically under half are
    // active); every item costs a pipeline drain / fill and a flush of ~10 000 clk
    chunk = std::max<int64_t>(4, std::min<int64_t>(64, 2 * chunk));
    std::vector<int64_t> item_tile_lo;
    std::vector<int32_t> item_unit;
    for (size_t blk = 0; blk + 1 < run_tile_off.size(); ++blk)
        for (int64_t lo = run_tile_off[blk]; lo < run_tile_off[blk + 1]; lo += chunk) {
            item_tile_lo.push_back(lo);
            item_unit.push_back((int32_t)(blk % (size_t)n_units));
        }
    item_tile_lo.push_back(n_tiles);
    std::vector<int32_t> tile_item((size_t)n_tiles);
    for (size_t it = 0; it + 1 < item_tile_lo.size(); ++it)
        for (int64_t t = item_tile_lo[it]; t < item_tile_lo[it + 1]; ++t) tile_item[(size_t)t] = (int32_t)it;
    const int64_t n_items = (int64_t)item_unit.size();
    PC_REQUIRE(n_items < 2147483647LL, "pc_corpus_create: too many work items");
    // transfer chunks for the host-buffer entry point: <= PC_MAX_CHUNKS runs of consecutive
    // utterances with about the same number of frames each (>= 32k frames per chunk)
    int n_chunks = (int)std::max<int64_t>(1, std::min<int64_t>(PC_MAX_CHUNKS, frame_off[n_utt] / 32768));
    std::vector<int32_t> chunk_utt(n_chunks + 1, n_utt);
    chunk_utt[0] = 0;
    for (int k = 1, u = 0; k < n_chunks; ++k) {
        const int64_t want = frame_off[n_utt] * k / n_chunks;
        while (u < n_utt && frame_off[u] < want) ++u;
        chunk_utt[k] = u;
    }
    std::vector<int32_t> utt_chunk(n_utt, 0);
    for (int k = 0; k < n_chunks; ++k)
        for (int u = chunk_utt[k]; u < chunk_utt[k + 1]; ++u) utt_chunk[u] = k;
    // utterance-major groups of <= 3 tiles for the scoring kernel: by chunk, heaviest first inside
    std::vector<int32_t> sitem_utt, sitem_t0, sitem_nt;
    std::vector<int32_t> chunk_sitem(n_chunks + 1, 0);
    {
        std::vector<int32_t> su, st0, snt;
        for (int u = 0; u < n_utt; ++u) {
            const int nt = (n_frames[u] + PC_TILE_ROWS - 1) / PC_TILE_ROWS;
            for (int j = 0; j < nt; j += 3) {
                su.push_back(u);
                st0.push_back(j * PC_TILE_ROWS);
                snt.push_back(std::min(3, nt - j));
            }
        }
        std::vector<int32_t> ord(su.size());
        std::iota(ord.begin(), ord.end(), 0);
        std::stable_sort(ord.begin(), ord.end(), [&](int a, int b) {
            if (utt_chunk[su[a]] != utt_chunk[su[b]]) return utt_chunk[su[a]] < utt_chunk[su[b]];
            return (int64_t)snt[a] * n_labels[su[a]] > (int64_t)snt[b] * n_labels[su[b]];
        });
        for (int32_t i : ord) {
            sitem_utt.push_back(su[i]);
            sitem_t0.push_back(st0[i]);
            sitem_nt.push_back(snt[i]);
            chunk_sitem[utt_chunk[su[i]] + 1]++;
        }
        for (int k = 0; k < n_chunks; ++k) chunk_sitem[k + 1] += chunk_sitem[k];
    }
    const int64_t n_sitems = (int64_t)sitem_utt.size();

    // one device block for every table
    struct Seg { const void *src; size_t bytes; size_t off; };
    std::vector<Seg> segs;
    size_t total = 0;
    auto add = [&](const void *src, size_t bytes) {
        size_t off = total;
        segs.push_back({src, bytes, off});
        total += (bytes + 255) & ~(size_t)255;
        return off;
    };
    size_t o_frame = add(frame_off.data(), frame_off.size() * 8);
    size_t o_emis = add(emis_off.data(), emis_off.size() * 8);
    size_t o_pair = add(pair_off.data(), pair_off.size() * 8);
    size_t o_state = add(state_off.data(), state_off.size() * 8);
    size_t o_labels = add(labels, (size_t)n_pairs * 4);
    size_t o_putt = add(pair_utt.data(), (size_t)n_pairs * 4);
    size_t o_fb = add(fb_order.data(), (size_t)n_utt * 4);
    size_t o_sorted = add(sorted_pair.data(), (size_t)n_pairs * 8);
    size_t o_upo = add(unit_pair_off.data(), unit_pair_off.size() * 8);
    size_t o_tpair = add(tile_pair.data(), (size_t)n_tiles * 8);
    size_t o_tt0 = add(tile_t0.data(), (size_t)n_tiles * 4);
    size_t o_trows = add(tile_rows.data(), (size_t)n_tiles * 4);
    size_t o_ttp = add(tile_tp.data(), (size_t)n_tiles * 4);
    size_t o_txrow = add(tile_xrow.data(), (size_t)n_tiles * 8);
    size_t o_tboff = add(tile_boff.data(), (size_t)n_tiles * 8);
    size_t o_pt0 = add(pair_tile0.data(), (size_t)n_pairs * 8);
    size_t o_ilo = add(item_tile_lo.data(), item_tile_lo.size() * 8);
    size_t o_iunit = add(item_unit.data(), (size_t)n_items * 4);
    size_t o_scratch = add(nullptr, (size_t)frame_off[n_utt] * 16);  // float4 per frame (K2)
    size_t o_scratch1 = add(nullptr, (size_t)emis_off[n_utt] * 4);    // beta_hat rows (K2)
    size_t o_tact = add(nullptr, (size_t)n_tiles * 4);                 // active-tile flags (K3)
    size_t o_iact = add(nullptr, (size_t)(n_items + 1) * 4);           // directly behind: item counts + a ticket counter
    size_t o_titem = add(tile_item.data(), (size_t)n_tiles * 4);
    size_t o_iord = add(nullptr, (size_t)n_items * 4);
    size_t o_ttmp = add(nullptr, (size_t)n_units * PC_TR_CHUNKS * PC_TRANS_SLOTS * 8);
    size_t o_tcnt = add(nullptr, (size_t)n_units * 4);
    size_t o_sutt = add(sitem_utt.data(), (size_t)n_sitems * 4);
    size_t o_st0 = add(sitem_t0.data(), (size_t)n_sitems * 4);
    size_t o_snt = add(sitem_nt.data(), (size_t)n_sitems * 4);
    size_t o_xoff = add(xtile_off.data(), xtile_off.size() * 8);
    size_t o_xutt = add(xtile_utt.data(), (size_t)n_xtiles * 4);
    size_t o_xt0 = add(xtile_t0.data(), (size_t)n_xtiles * 4);
    size_t o_txblk = add(tile_xblk.data(), (size_t)n_tiles * 8);
    PC_CUDA_TRY(cudaSetDevice(h->device));
    char *dev = nullptr;
    PC_CUDA_TRY(cudaMalloc((void **)&dev, std::max<size_t>(total, 256)));
    if (cudaMemset(dev, 0, std::max<size_t>(total, 256)) != cudaSuccess) {  // scratch regions start clean (K3's ticket counter)
        cudaFree(dev);
        pc_set_error("pc_corpus_create: cudaMemset failed");
        return PC_ERR_CUDA;
    }
    for (const Seg &s : segs)
        if (s.src && s.bytes) {
            cudaError_t e = cudaMemcpy(dev + s.off, s.src, s.bytes, cudaMemcpyHostToDevice);
            if (e != cudaSuccess) {
                cudaFree(dev);
                pc_set_error("pc_corpus_create: upload failed: %s", cudaGetErrorString(e));
                return PC_ERR_CUDA;
            }
        }
    pc_corpus c = new pc_corpus_s();
    memset(c, 0, sizeof(*c));
    c->h = h;
    c->device = h->device;
    c->dev_block = dev;
    c->total_frames = frame_off[n_utt];
    c->emis_floats = emis_off[n_utt];
    c->total_states = state_off[n_utt];
    c->items_per_chunk = (int32_t)chunk;
    c->n_chunks = n_chunks;
    for (int k = 0; k <= n_chunks; ++k) {
        const int u = chunk_utt[k];
        c->chunk_frame[k] = frame_off[u];
        c->chunk_xtile[k] = xtile_off[u];
        c->chunk_sitem[k] = chunk_sitem[k];
    }
    CorpusView &v = c->v;
    v.n_utt = n_utt;
    v.n_units = n_units;
    v.n_pairs = n_pairs;
    v.n_tiles = n_tiles;
    v.n_items = (int32_t)n_items;
    v.max_frames = max_frames;
    v.max_labels = max_labels;
    v.frame_off = (const int64_t *)(dev + o_frame);
    v.emis_off = (const int64_t *)(dev + o_emis);
    v.pair_off = (const int64_t *)(dev + o_pair);
    v.state_off = (const int64_t *)(dev + o_state);
    v.labels = (const int32_t *)(dev + o_labels);
    v.pair_utt = (const int32_t *)(dev + o_putt);
    v.fb_order = (const int32_t *)(dev + o_fb);
    v.sorted_pair = (const int64_t *)(dev + o_sorted);
    v.unit_pair_off = (const int64_t *)(dev + o_upo);
    v.tile_pair = (const int64_t *)(dev + o_tpair);
    v.tile_t0 = (const int32_t *)(dev + o_tt0);
    v.tile_rows = (const int32_t *)(dev + o_trows);
    v.tile_tp = (const int32_t *)(dev + o_ttp);
    v.tile_xrow = (const int64_t *)(dev + o_txrow);
    v.tile_boff = (const int64_t *)(dev + o_tboff);
    v.pair_tile0 = (const int64_t *)(dev + o_pt0);
    v.item_tile_lo = (const int64_t *)(dev + o_ilo);
    v.item_unit = (const int32_t *)(dev + o_iunit);
    v.scratch0 = (float *)(dev + o_scratch);
    v.scratch1 = (float *)(dev + o_scratch1);
    v.tile_active = (int32_t *)(dev + o_tact);
    v.tile_item = (const int32_t *)(dev + o_titem);
    v.item_act = (int32_t *)(dev + o_iact);
    v.trans_tmp = (double *)(dev + o_ttmp);
    v.trans_cnt = (int32_t *)(dev + o_tcnt);
    v.item_order = (int32_t *)(dev + o_iord);
    v.total_frames = frame_off[n_utt];
    v.n_sitems = (int32_t)n_sitems;
    v.sitem_utt = (const int32_t *)(dev + o_sutt);
    v.sitem_t0 = (const int32_t *)(dev + o_st0);
    v.sitem_nt = (const int32_t *)(dev + o_snt);
    v.n_xtiles = n_xtiles;
    v.xtile_off = (const int64_t *)(dev + o_xoff);
    v.xtile_utt = (const int32_t *)(dev + o_xutt);
    v.xtile_t0 = (const int32_t *)(dev + o_xt0);
    v.tile_xblk = (const int64_t *)(dev + o_txblk);
    c->host_frame_off = new int64_t[4 * (size_t)(n_utt + 1)];
    c->host_emis_off = c->host_frame_off + (n_utt + 1);
    c->host_pair_off = c->host_emis_off + (n_utt + 1);
    c->host_state_off = c->host_pair_off + (n_utt + 1);
    memcpy(c->host_frame_off, frame_off.data(), (size_t)(n_utt + 1) * 8);
    memcpy(c->host_emis_off, emis_off.data(), (size_t)(n_utt + 1) * 8);
    memcpy(c->host_pair_off, pair_off.data(), (size_t)(n_utt + 1) * 8);
    memcpy(c->host_state_off, state_off.data(), (size_t)(n_utt + 1) * 8);
    *out = c;
    return PC_OK;
}

int pc_corpus_destroy(pc_corpus c) {
    if (!c) return PC_OK;
    cudaSetDevice(c->device);  // not c->h: the handle may already be gone
    if (c->dev_block) cudaFree(c->dev_block);
    delete[] c->host_frame_off;
    delete c;
    return PC_OK;
}

int64_t pc_corpus_active_tiles(pc_corpus c) {
    if (!c) return -1;
    std::vector<int32_t> a((size_t)c->v.n_tiles);
    if (cudaSetDevice(c->device) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess ||
        cudaMemcpy(a.data(), c->v.tile_active, a.size() * 4, cudaMemcpyDeviceToHost) != cudaSuccess)
        return -1;
    int64_t n = 0;
    for (int32_t x : a) n += x != 0;
    return n;
}

int64_t pc_corpus_total_tiles(pc_corpus c) { return c ? c->v.n_tiles : -1; }
int64_t pc_corpus_total_frames(pc_corpus c) { return c ? c->total_frames : -1; }
int64_t pc_corpus_emission_floats(pc_corpus c) { return c ? c->emis_floats : -1; }
int64_t pc_corpus_total_pairs(pc_corpus c) { return c ? c->v.n_pairs : -1; }
int64_t pc_corpus_total_states(pc_corpus c) { return c ? c->total_states : -1; }
int64_t pc_corpus_frames_bytes(pc_corpus c) {
    return c ? (int64_t)(pc_x32_offset(c->total_frames, c->v.n_xtiles) + (size_t)c->v.n_xtiles * PC_X32TILE_BYTES) : -1;
}

int pc_corpus_offsets(pc_corpus c, int64_t *frame_off, int64_t *emis_off, int64_t *pair_off,
                      int64_t *state_off) {
    PC_REQUIRE(c, "pc_corpus_offsets: NULL corpus");
    size_t n = (size_t)(c->v.n_utt + 1) * 8;
    if (frame_off) memcpy(frame_off, c->host_frame_off, n);
    if (emis_off) memcpy(emis_off, c->host_emis_off, n);
    if (pair_off) memcpy(pair_off, c->host_pair_off, n);
    if (state_off) memcpy(state_off, c->host_state_off, n);
    return PC_OK;
}

// --------------------------------------------------------------------------------- kernels
// (a stale error left behind by an unrelated runtime call must not be blamed on this call's launches)
#define PC_ENTER(h)                                           \
    PC_REQUIRE((h) != nullptr, "%s: NULL handle", __func__);  \
    PC_CUDA_TRY(cudaSetDevice((h)->device));                  \
    (void)cudaGetLastError()

static int check_dim_mix(const char *fn, int dim, int mix) {
    if (dim < 1 || dim > PC_DIM_MAX) {
        pc_set_error("%s: dimension %d outside [1,%d] (DataDimensionError in the reference, "
                     "Clustering.py:749-751)", fn, dim, PC_DIM_MAX);
        return PC_ERR_INVALID;
    }
    if (mix < 1) {
        pc_set_error("%s: mix=%d", fn, mix);
        return PC_ERR_INVALID;
    }
    return PC_OK;
}

int pc_pack_gmm(pc_handle h, const double *mean, const double *var, const double *alpha,
                const double *shift, const double *inv_scale, int32_t n_gauss, int32_t dim,
                int32_t mix, float *W, void *stream) {
    PC_ENTER(h);
    PC_REQUIRE(n_gauss >= 0, "pc_pack_gmm: n_gauss=%d", n_gauss);
    PC_REQUIRE(n_gauss == 0 || (mean && var && alpha && W), "pc_pack_gmm: NULL pointer");
    int rc = check_dim_mix("pc_pack_gmm", dim, 1);
    if (rc) return rc;
    PC_REQUIRE(mix >= 0, "pc_pack_gmm: mix=%d", mix);
    PC_REQUIRE(mix == 0 || n_gauss % (PC_EMIT * mix) == 0,
               "pc_pack_gmm: %d Gaussians is not a multiple of 3*mix (mix=%d)", n_gauss, mix);
    return launch_pack_gmm(h, mean, var, alpha, shift, inv_scale, n_gauss, dim, mix, W,
                           (cudaStream_t)stream);
}

int pc_prepare_frames_f64(pc_handle h, pc_corpus c, const double *x, int32_t dim,
                          const double *shift, const double *inv_scale, float *X, void *stream) {
    PC_ENTER(h);
    PC_REQUIRE(c && (c->total_frames == 0 || (x && X)), "pc_prepare_frames_f64: bad arguments");
    int rc = check_dim_mix("pc_prepare_frames_f64", dim, 1);
    if (rc) return rc;
    return launch_prepare_frames(h, c->v, x, 1, dim, shift, inv_scale, X, 0, c->v.n_xtiles, (cudaStream_t)stream);
}

int pc_prepare_frames_f32(pc_handle h, pc_corpus c, const float *x, int32_t dim, const double *shift,
                          const double *inv_scale, float *X, void *stream) {
    PC_ENTER(h);
    PC_REQUIRE(c && (c->total_frames == 0 || (x && X)), "pc_prepare_frames_f32: bad arguments");
    int rc = check_dim_mix("pc_prepare_frames_f32", dim, 1);
    if (rc) return rc;
    return launch_prepare_frames(h, c->v, x, 0, dim, shift, inv_scale, X, 0, c->v.n_xtiles, (cudaStream_t)stream);
}

int pc_prepare_rows_f64(pc_handle h, const double *x, int64_t n, int32_t dim, const double *shift,
                        const double *inv_scale, float *X, void *stream) {
    PC_ENTER(h);
    PC_REQUIRE(n >= 0 && (n == 0 || (x && X)), "pc_prepare_rows_f64: bad arguments");
    int rc = check_dim_mix("pc_prepare_rows_f64", dim, 1);
    if (rc) return rc;
    return launch_prepare_rows(h, x, 1, n, dim, shift, inv_scale, X, (cudaStream_t)stream);
}

int pc_prepare_rows_f32(pc_handle h, const float *x, int64_t n, int32_t dim, const double *shift,
                        const double *inv_scale, float *X, void *stream) {
    PC_ENTER(h);
    PC_REQUIRE(n >= 0 && (n == 0 || (x && X)), "pc_prepare_rows_f32: bad arguments");
    int rc = check_dim_mix("pc_prepare_rows_f32", dim, 1);
    if (rc) return rc;
    return launch_prepare_rows(h, x, 0, n, dim, shift, inv_scale, X, (cudaStream_t)stream);
}

int pc_gmm_score(pc_handle h, pc_corpus c, const float *X, const float *W, int32_t mix, float *b,
                 void *stream) {
    PC_ENTER(h);
    PC_REQUIRE(c && X && W && b, "pc_gmm_score: NULL argument");
    int rc = check_dim_mix("pc_gmm_score", 1, mix);
    if (rc) return rc;
    if (h->use_tc && score_tc_supported(mix))
        return launch_score_tc(h, c->v, X, W, mix, b, 0, c->v.n_sitems, (cudaStream_t)stream);
    return launch_score_simt(h, c->v, X, W, mix, b, (cudaStream_t)stream);
}

int pc_gmm_score_dense(pc_handle h, const float *X, int64_t n, const float *W, int32_t n_states,
                       int32_t mix, float *out, void *stream) {
    PC_ENTER(h);
    PC_REQUIRE(n >= 0 && n_states >= 0, "pc_gmm_score_dense: negative size");
    PC_REQUIRE((n == 0 || n_states == 0) || (X && W && out), "pc_gmm_score_dense: NULL argument");
    int rc = check_dim_mix("pc_gmm_score_dense", 1, mix);
    if (rc) return rc;
    return launch_score_dense_simt(h, X, n, W, n_states, mix, out, (cudaStream_t)stream);
}

int pc_forward_backward(pc_handle h, pc_corpus c, const float *b, const double *log_self,
                        const double *log_next, float *lgam, double *utt_logp, int32_t *utt_iters,
                        float *pair_trans, void *stream) {
    PC_ENTER(h);
    PC_REQUIRE(c && b && log_self && log_next && lgam && utt_logp && utt_iters && pair_trans,
               "pc_forward_backward: NULL argument");
    int rc = launch_forward_backward(h, c->v, b, log_self, log_next, lgam, c->v.scratch0, utt_logp,
                                     utt_iters, pair_trans, (cudaStream_t)stream);
    c->flags_lgam = rc == PC_OK ? lgam : nullptr;
    return rc;
}

int pc_accumulate(pc_handle h, pc_corpus c, const float *X, const float *W, int32_t mix,
                  const float *b, const float *lgam, double *acc, void *stream) {
    PC_ENTER(h);
    PC_REQUIRE(c && X && W && b && lgam && acc, "pc_accumulate: NULL argument");
    int rc = check_dim_mix("pc_accumulate", 1, mix);
    if (rc) return rc;
    // the flags K2 left behind are used once, for the buffer it wrote, on the stream order the caller
    // gives; log gamma from anywhere else (or a second pass) goes through the pre-pass
    const bool fresh = c->flags_lgam == lgam && !(h->debug_flags & 2048);
    c->flags_lgam = nullptr;
    if (h->use_tc && h->k3_kernel && accumulate_tcx_supported(mix))
        return launch_accumulate_tcx(h, c->v, X, W, mix, b, lgam, acc, fresh, (cudaStream_t)stream);
    if (h->use_tc && accumulate_tc_supported(mix))
        return launch_accumulate_tc(h, c->v, X, W, mix, b, lgam, acc, fresh, (cudaStream_t)stream);
    return launch_accumulate_simt(h, c->v, X, W, mix, b, lgam, acc, (cudaStream_t)stream);
}

int pc_transitions_max(pc_handle h, pc_corpus c, const double *utt_logp, const float *pair_trans,
                       double *tmax, void *stream) {
    PC_ENTER(h);
    PC_REQUIRE(c && utt_logp && pair_trans && tmax, "pc_transitions_max: NULL argument");
    return launch_transitions_max(h, c->v, utt_logp, pair_trans, tmax, (cudaStream_t)stream);
}

int pc_transitions_sum(pc_handle h, pc_corpus c, const double *utt_logp, const float *pair_trans,
                       const double *tmax, double *tsum, void *stream) {
    PC_ENTER(h);
    PC_REQUIRE(c && utt_logp && pair_trans && tmax && tsum, "pc_transitions_sum: NULL argument");
    return launch_transitions_sum(h, c->v, utt_logp, pair_trans, tmax, tsum, (cudaStream_t)stream);
}

int pc_update_params(pc_handle h, int32_t n_units, int32_t mix, int32_t dim, const double *acc,
                     const double *tmax, const double *tsum, const double *shift,
                     const double *inv_scale, double c_cov, int32_t fix_code, double *mean,
                     double *var, double *alpha, double *transmat, void *stream) {
    PC_ENTER(h);
    PC_REQUIRE(n_units >= 0, "pc_update_params: n_units=%d", n_units);
    PC_REQUIRE(acc && tmax && tsum && mean && var && alpha && transmat,
               "pc_update_params: NULL argument");
    int rc = check_dim_mix("pc_update_params", dim, mix);
    if (rc) return rc;
    return launch_update_params(h, n_units, mix, dim, acc, tmax, tsum, shift, inv_scale, c_cov,
                                fix_code, mean, var, alpha, transmat, (cudaStream_t)stream);
}

int pc_viterbi(pc_handle h, pc_corpus c, const float *b, const double *b64, const double *log_self,
               const double *log_next, const double *utt_logpi, const double *state_logpi,
               int32_t *path, int32_t *unit_path, double *score, void *stream) {
    PC_ENTER(h);
    PC_REQUIRE(c && log_self && log_next && path && score, "pc_viterbi: NULL argument");
    PC_REQUIRE((b != nullptr) != (b64 != nullptr), "pc_viterbi: exactly one of b / b64 must be set");
    PC_REQUIRE((utt_logpi != nullptr) != (state_logpi != nullptr),
               "pc_viterbi: exactly one of utt_logpi / state_logpi must be set");
    return launch_viterbi(h, c->v, b, b64, log_self, log_next, utt_logpi, state_logpi, path,
                          unit_path, score, (cudaStream_t)stream);
}

// ------------------------------------------------------------------ alignment post-processing
int pc_segment_keys(pc_handle h, pc_corpus c, int32_t mode, const int32_t *path, int32_t *frame_key,
                    int32_t *utt_kept, void *stream) {
    PC_ENTER(h);
    PC_REQUIRE(c && frame_key, "pc_segment_keys: NULL argument");
    PC_REQUIRE(mode == 0 || mode == 1, "pc_segment_keys: mode %d (0 = uniform, 1 = aligned path)", mode);
    PC_REQUIRE(mode == 0 || path, "pc_segment_keys: mode 1 needs the path pc_viterbi wrote");
    return launch_segment_keys(h, c->v, mode, path, frame_key, utt_kept, (cudaStream_t)stream);
}

int64_t pc_group_workspace_bytes(int64_t n_frames, int32_t n_keys) {
    if (n_frames < 0 || n_frames >= (1ll << 31) || n_keys < 1 || n_keys > 12000) return -1;
    return group_workspace_bytes(n_frames, n_keys);
}

int pc_group_frames(pc_handle h, const int32_t *frame_key, int64_t n_frames, int32_t n_keys, void *ws,
                    int64_t *key_off, int32_t *order, void *stream) {
    PC_ENTER(h);
    PC_REQUIRE(pc_group_workspace_bytes(n_frames, n_keys) >= 0,
               "pc_group_frames: n_frames %lld / n_keys %d outside the supported range", (long long)n_frames, n_keys);
    PC_REQUIRE(ws && key_off && (n_frames == 0 || (frame_key && order)), "pc_group_frames: NULL argument");
    return launch_group_frames(h, frame_key, n_frames, n_keys, ws, key_off, order, (cudaStream_t)stream);
}

int pc_gather_rows(pc_handle h, const int32_t *order, int64_t n_rows, int32_t row_bytes, const void *src,
                   void *dst, void *stream) {
    PC_ENTER(h);
    PC_REQUIRE(n_rows >= 0 && row_bytes > 0 && row_bytes % 4 == 0,
               "pc_gather_rows: row_bytes %d must be a positive multiple of 4", row_bytes);
    PC_REQUIRE(n_rows == 0 || (order && src && dst), "pc_gather_rows: NULL argument");
    return launch_gather_rows(h, order, n_rows, row_bytes, src, dst, (cudaStream_t)stream);
}

// --------------------------------------------------------------------------------- host e2e
static int ensure_ws(pc_handle h, size_t bytes) {
    if (h->ws_bytes >= bytes) return PC_OK;
    if (h->ws) cudaFree(h->ws);
    h->ws = nullptr;
    h->ws_bytes = 0;
    PC_CUDA_TRY(cudaMalloc(&h->ws, bytes));
    h->ws_bytes = bytes;
    return PC_OK;
}

__global__ void fill_double_kernel(double *p, int64_t n, double v) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

__global__ void log_bands_kernel(const double *transmat, int n_units, double *log_self,
                                 double *log_next) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;  // (unit, state)
    if (i >= n_units * PC_STATES) return;
    int unit = i / PC_STATES, s = i - unit * PC_STATES;
    const double *row = transmat + ((size_t)unit * PC_STATES + s) * PC_STATES;
    log_self[i] = log(row[s]);
    log_next[i] = (s + 1 < PC_STATES) ? log(row[s + 1]) : -INFINITY;
}

int pc_log_bands(pc_handle h, const double *transmat, int32_t n_units, double *log_self, double *log_next,
                 void *stream) {
    PC_ENTER(h);
    PC_REQUIRE(n_units >= 0 && (n_units == 0 || (transmat && log_self && log_next)), "pc_log_bands: bad arguments");
    if (n_units == 0) return PC_OK;
    log_bands_kernel<<<(n_units * PC_STATES + 127) / 128, 128, 0, (cudaStream_t)stream>>>(transmat, n_units, log_self,
                                                                                         log_next);
    PC_LAUNCH_CHECK();
    h->launches++;
    return PC_OK;
}

__global__ void sum_double_kernel(const double *p, int n, double *out) {
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += 32) s += p[i];
    s = warp_sum_d(s);
    if (threadIdx.x == 0) *out = s;
}

// per-dimension sum / sum of squares of float frames [n][dim]: a block strides over rows, thread =
// (row slot, dimension); fp64 partial sums, one atomicAdd per block and dimension
__global__ void __launch_bounds__(256)
frame_moments_kernel(const float *__restrict__ x, int64_t n, int dim, double *__restrict__ mom) {
    __shared__ double sh[2][256];
    const int d = threadIdx.x % 64, slot = threadIdx.x / 64;  // 4 row slots x 64 dimension lanes
    double s = 0.0, q = 0.0;
    if (d < dim)
        for (int64_t r = (int64_t)blockIdx.x * 4 + slot; r < n; r += (int64_t)gridDim.x * 4) {
            const double v = (double)x[r * dim + d];
            s += v;
            q += v * v;
        }
    sh[0][threadIdx.x] = s;
    sh[1][threadIdx.x] = q;
    __syncthreads();
    if (slot == 0 && d < dim) {
        for (int k = 1; k < 4; ++k) { s += sh[0][k * 64 + d]; q += sh[1][k * 64 + d]; }
        atomicAdd(mom + d, s);
        atomicAdd(mom + PC_XS + d, q);
    }
}

// shift = mean, inv_scale = 1 / std from the moments (one thread per dimension)
__global__ void moments_to_affine_kernel(const double *__restrict__ mom, int64_t n, int dim, double *shift,
                                         double *inv_scale) {
    const int d = threadIdx.x;
    if (d >= dim) return;
    const double mu = mom[d] / (double)n;
    double var = mom[PC_XS + d] / (double)n - mu * mu;
    if (!(var > 0.0)) var = 0.0;
    double sd = sqrt(var);
    if (sd < 1e-12) sd = 1e-12;
    shift[d] = mu;
    inv_scale[d] = 1.0 / sd;
}

static int ensure_streams(pc_handle h) {
    if (h->copy_stream) return PC_OK;
    PC_CUDA_TRY(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    PC_CUDA_TRY(cudaEventCreateWithFlags(&h->start_ev, cudaEventDisableTiming));
    PC_CUDA_TRY(cudaEventCreateWithFlags(&h->fork_ev, cudaEventDisableTiming));
    PC_CUDA_TRY(cudaEventCreateWithFlags(&h->join_ev, cudaEventDisableTiming));
    for (int k = 0; k < PC_MAX_CHUNKS; ++k)
        PC_CUDA_TRY(cudaEventCreateWithFlags(&h->chunk_ev[k], cudaEventDisableTiming));
    return PC_OK;
}

int pc_frame_moments_host(pc_handle h, const float *host_frames, int64_t n_frames, int32_t dim,
                          double *host_sum, double *host_sumsq, void *stream) {
    PC_ENTER(h);
    PC_REQUIRE(n_frames >= 0 && host_sum && host_sumsq && (n_frames == 0 || host_frames),
               "pc_frame_moments_host: bad arguments");
    int rc = check_dim_mix("pc_frame_moments_host", dim, 1);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    // frames go through the workspace in slabs of <= 64 MiB, two in flight
    const int64_t slab_rows = std::max<int64_t>(1, (64ll << 20) / ((int64_t)dim * 4));
    const size_t slab_bytes = (size_t)slab_rows * dim * 4;
    if ((rc = ensure_ws(h, 2 * slab_bytes + 4096))) return rc;
    double *mom = (double *)((char *)h->ws + 2 * slab_bytes);
    PC_CUDA_TRY(cudaMemsetAsync(mom, 0, 2 * PC_XS * 8, st));
    for (int64_t r0 = 0, k = 0; r0 < n_frames; r0 += slab_rows, ++k) {
        const int64_t rows = std::min(slab_rows, n_frames - r0);
        float *dst = (float *)((char *)h->ws + (k & 1) * slab_bytes);
        PC_CUDA_TRY(cudaMemcpyAsync(dst, host_frames + (size_t)r0 * dim, (size_t)rows * dim * 4,
                                    cudaMemcpyHostToDevice, st));
        const int blocks = (int)std::min<int64_t>((rows + 3) / 4, (int64_t)h->sm_count * 8);
        frame_moments_kernel<<<blocks, 256, 0, st>>>(dst, rows, dim, mom);
        PC_LAUNCH_CHECK();
        h->launches++;
    }
    double out[2 * PC_XS];
    PC_CUDA_TRY(cudaMemcpyAsync(out, mom, sizeof(out), cudaMemcpyDeviceToHost, st));
    PC_CUDA_TRY(cudaStreamSynchronize(st));
    for (int d = 0; d < dim; ++d) {
        host_sum[d] = out[d];
        host_sumsq[d] = out[PC_XS + d];
    }
    return PC_OK;
}

int pc_set_reduce_hook(pc_handle h, pc_reduce_hook fn, void *user, double *dev_tmax, double *dev_flat,
                       int64_t flat_len) {
    PC_REQUIRE(h, "pc_set_reduce_hook: NULL handle");
    PC_REQUIRE(fn == nullptr || (dev_tmax && dev_flat && flat_len > 0),
               "pc_set_reduce_hook: a hook needs its exchange buffers");
    h->hook = fn;
    h->hook_user = user;
    h->hook_tmax = fn ? dev_tmax : nullptr;
    h->hook_flat = fn ? dev_flat : nullptr;
    h->hook_flat_len = fn ? flat_len : 0;
    return PC_OK;
}

// --------------------------------------------------------------------------------- peer-memory reduction
static size_t peer_set_doubles(pc_handle h) { return (size_t)h->peer_n_acc + 2 * (size_t)h->peer_n_units * PC_TRANS_SLOTS; }
static size_t peer_block_bytes(pc_handle h) { return 4 * peer_set_doubles(h) * sizeof(double) + 2 * PC_MAX_PEERS * sizeof(int) + 256; }

int pc_peer_destroy(pc_handle h) {
    if (!h || h->peer_n == 0) return PC_OK;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    for (int r = 0; r < h->peer_n; ++r) {
        if (!h->peer_block[r]) continue;
        if (r == h->peer_rank) cudaFree(h->peer_block[r]);
        else cudaIpcCloseMemHandle(h->peer_block[r]);
        h->peer_block[r] = nullptr;
    }
    h->peer_n = 0;
    h->peer_epoch = 0;
    return PC_OK;
}

int pc_peer_create(pc_handle h, int32_t rank, int32_t n_ranks, int64_t n_gauss, int32_t n_units,
                   uint8_t *handle_out) {
    PC_ENTER(h);
    PC_REQUIRE(handle_out && n_ranks >= 1 && n_ranks <= PC_MAX_PEERS && rank >= 0 && rank < n_ranks && n_gauss > 0 &&
                   n_units > 0, "pc_peer_create: rank %d of %d (at most %d), %lld Gaussians, %d units", rank, n_ranks,
               PC_MAX_PEERS, (long long)n_gauss, n_units);
    static_assert(sizeof(cudaIpcMemHandle_t) == PC_IPC_HANDLE_BYTES, "IPC handle size");
    pc_peer_destroy(h);
    h->peer_n = n_ranks;
    h->peer_rank = rank;
    h->peer_n_acc = n_gauss * PC_KA;
    h->peer_n_units = n_units;
    h->peer_epoch = 0;
    for (int r = 0; r < PC_MAX_PEERS; ++r) h->peer_block[r] = nullptr;
    const size_t bytes = peer_block_bytes(h);
    if (cudaMalloc(&h->peer_block[rank], bytes) != cudaSuccess) {
        h->peer_n = 0;
        pc_set_error("pc_peer_create: cudaMalloc of %zu bytes failed", bytes);
        return PC_ERR_CUDA;
    }
    PC_CUDA_TRY(cudaMemset(h->peer_block[rank], 0, bytes));
    cudaIpcMemHandle_t ipc;
    PC_CUDA_TRY(cudaIpcGetMemHandle(&ipc, h->peer_block[rank]));
    memcpy(handle_out, &ipc, sizeof(ipc));
    return PC_OK;
}

int pc_peer_connect(pc_handle h, const uint8_t *handles) {
    PC_ENTER(h);
    PC_REQUIRE(handles && h->peer_n > 0, "pc_peer_connect: pc_peer_create first");
    for (int r = 0; r < h->peer_n; ++r) {
        if (r == h->peer_rank) continue;
        cudaIpcMemHandle_t ipc;
        memcpy(&ipc, handles + (size_t)r * PC_IPC_HANDLE_BYTES, sizeof(ipc));
        PC_CUDA_TRY(cudaIpcOpenMemHandle(&h->peer_block[r], ipc, cudaIpcMemLazyEnablePeerAccess));
    }
    return PC_OK;
}

int pc_peer_buffers(pc_handle h, int32_t which, double **acc, double **tsum, double **tmax) {
    PC_REQUIRE(h && h->peer_n > 0 && which >= 0 && which <= 2, "pc_peer_buffers: no exchange block / which=%d", which);
    double *set = (double *)h->peer_block[h->peer_rank] + (size_t)which * peer_set_doubles(h);
    if (acc) *acc = set;
    if (tsum) *tsum = set + h->peer_n_acc;
    if (tmax) *tmax = set + h->peer_n_acc + (size_t)h->peer_n_units * PC_TRANS_SLOTS;
    return PC_OK;
}

int pc_update_params_peer(pc_handle h, int32_t mix, int32_t dim, const double *shift, const double *inv_scale,
                          double c_cov, int32_t fix_code, double *mean, double *var, double *alpha, double *transmat,
                          void *stream) {
    PC_ENTER(h);
    PC_REQUIRE(h->peer_n > 0, "pc_update_params_peer: no exchange block (pc_peer_create / pc_peer_connect)");
    PC_REQUIRE(mean && var && alpha && transmat, "pc_update_params_peer: NULL parameter pointer");
    int rc = check_dim_mix("pc_update_params_peer", dim, mix);
    if (rc) return rc;
    PC_REQUIRE((int64_t)h->peer_n_units * PC_EMIT * mix * PC_KA == h->peer_n_acc,
               "pc_update_params_peer: the block was created for %lld statistics, mix=%d needs %lld",
               (long long)h->peer_n_acc, mix, (long long)h->peer_n_units * PC_EMIT * mix * PC_KA);
    return launch_update_params_peer(h, mix, dim, shift, inv_scale, c_cov, fix_code, mean, var, alpha, transmat,
                                     (cudaStream_t)stream);
}

int pc_em_iteration_host(pc_handle h, pc_corpus c, const float *host_frames, int32_t dim,
                         int32_t n_units, int32_t mix, double *host_mean, double *host_var,
                         double *host_alpha, double *host_transmat, const double *host_shift,
                         const double *host_inv_scale, double c_cov, int32_t fix_code,
                         double *host_sum_logp, void *stream) {
    PC_ENTER(h);
    PC_REQUIRE(c && host_frames && host_mean && host_var && host_alpha && host_transmat,
               "pc_em_iteration_host: NULL argument");
    PC_REQUIRE(n_units == c->v.n_units, "pc_em_iteration_host: n_units %d != corpus %d", n_units,
               c->v.n_units);
    PC_REQUIRE((host_shift != nullptr) == (host_inv_scale != nullptr),
               "pc_em_iteration_host: host_shift / host_inv_scale go together");
    int rc = check_dim_mix("pc_em_iteration_host", dim, mix);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t F = c->total_frames;
    const int64_t G = (int64_t)n_units * PC_EMIT * mix;
    const int64_t flat_len = G * PC_KA + (int64_t)n_units * PC_TRANS_SLOTS;
    PC_REQUIRE(!h->hook || h->hook_flat_len == flat_len,
               "pc_em_iteration_host: the reduce hook's flat buffer holds %lld doubles, this model needs %lld",
               (long long)h->hook_flat_len, (long long)flat_len);
    // workspace carve-up (256-byte aligned)
    size_t off = 0;
    auto carve = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    size_t o_raw = carve((size_t)F * dim * 4), o_X = carve((size_t)pc_corpus_frames_bytes(c));
    size_t o_b = carve((size_t)c->emis_floats * 4), o_lg = carve((size_t)c->emis_floats * 4);
    size_t o_W = carve((size_t)pc_gmm_bytes(G)), o_flat = carve((size_t)flat_len * 8);
    size_t o_mean = carve((size_t)G * dim * 8), o_var = carve((size_t)G * dim * 8);
    size_t o_alpha = carve((size_t)G * 8), o_tm = carve((size_t)n_units * 25 * 8);
    size_t o_ls = carve((size_t)n_units * 5 * 8), o_ln = carve((size_t)n_units * 5 * 8);
    size_t o_logp = carve((size_t)c->v.n_utt * 8), o_it = carve((size_t)c->v.n_utt * 4);
    size_t o_pt = carve((size_t)c->v.n_pairs * PC_TRANS_SLOTS * 4);
    size_t o_tmax = carve((size_t)n_units * PC_TRANS_SLOTS * 8);
    size_t o_sum = carve(8), o_aff = carve(2 * PC_XS * 8), o_mom = carve(2 * PC_XS * 8);
    rc = ensure_ws(h, off);
    if (rc) return rc;
    char *ws = (char *)h->ws;
    float *raw = (float *)(ws + o_raw), *X = (float *)(ws + o_X), *b = (float *)(ws + o_b);
    float *lg = (float *)(ws + o_lg), *W = (float *)(ws + o_W), *pt = (float *)(ws + o_pt);
    // statistics: the hook's exchange buffers, the current set of the peer-memory block, or the workspace
    const bool peer = !h->hook && h->peer_n > 1;
    PC_REQUIRE(!peer || (h->peer_n_acc == G * PC_KA && h->peer_n_units == n_units),
               "pc_em_iteration_host: the peer exchange block was created for another model");
    double *flat = h->hook ? h->hook_flat : (double *)(ws + o_flat);
    double *tmax = h->hook ? h->hook_tmax : (double *)(ws + o_tmax);
    if (peer) pc_peer_buffers(h, (int32_t)(h->peer_epoch & 1), &flat, nullptr, &tmax);
    double *acc = flat, *tsum = flat + G * PC_KA;
    double *mean = (double *)(ws + o_mean);
    double *var = (double *)(ws + o_var), *alpha = (double *)(ws + o_alpha);
    double *tm = (double *)(ws + o_tm), *ls = (double *)(ws + o_ls), *ln = (double *)(ws + o_ln);
    double *logp = (double *)(ws + o_logp);
    double *sum = (double *)(ws + o_sum);
    double *shift = (double *)(ws + o_aff), *inv_scale = shift + PC_XS, *mom = (double *)(ws + o_mom);
    int32_t *iters = (int32_t *)(ws + o_it);

    // Frames travel on their own stream in <= PC_MAX_CHUNKS runs of utterances; the compute stream
    // prepares and scores run k as soon as it has landed, under the copy of run k+1 (pinned host
    // memory makes the copies asynchronous; pageable memory still works, without the overlap).
    if ((rc = ensure_streams(h))) return rc;
    // the model first: host-to-device copies are served in the order they were issued, and the packing of the model
    // must not queue behind 47 MB of frames
    PC_CUDA_TRY(cudaMemcpyAsync(mean, host_mean, (size_t)G * dim * 8, cudaMemcpyHostToDevice, st));
    PC_CUDA_TRY(cudaMemcpyAsync(var, host_var, (size_t)G * dim * 8, cudaMemcpyHostToDevice, st));
    PC_CUDA_TRY(cudaMemcpyAsync(alpha, host_alpha, (size_t)G * 8, cudaMemcpyHostToDevice, st));
    PC_CUDA_TRY(cudaMemcpyAsync(tm, host_transmat, (size_t)n_units * 25 * 8, cudaMemcpyHostToDevice, st));
    if (host_shift) {
        PC_CUDA_TRY(cudaMemcpyAsync(shift, host_shift, (size_t)dim * 8, cudaMemcpyHostToDevice, st));
        PC_CUDA_TRY(cudaMemcpyAsync(inv_scale, host_inv_scale, (size_t)dim * 8, cudaMemcpyHostToDevice, st));
    }
    PC_CUDA_TRY(cudaEventRecord(h->start_ev, st));  // whatever the caller queued before us
    PC_CUDA_TRY(cudaStreamWaitEvent(h->copy_stream, h->start_ev, 0));
    // option "host_chunks" merges neighbouring runs (1 = one copy, no overlap)
    const int cap = h->host_chunks > 0 ? h->host_chunks : PC_MAX_CHUNKS;
    const int kstep = (c->n_chunks + cap - 1) / cap;
    for (int k = 0; k < c->n_chunks; k += kstep) {
        const int k1 = k + kstep < c->n_chunks ? k + kstep : c->n_chunks;
        const int64_t f_lo = c->chunk_frame[k], f_hi = c->chunk_frame[k1];
        PC_CUDA_TRY(cudaMemcpyAsync(raw + (size_t)f_lo * dim, host_frames + (size_t)f_lo * dim,
                                    (size_t)(f_hi - f_lo) * dim * 4, cudaMemcpyHostToDevice, h->copy_stream));
        PC_CUDA_TRY(cudaEventRecord(h->chunk_ev[k], h->copy_stream));
    }
    PC_CUDA_TRY(cudaMemsetAsync(acc, 0, (size_t)G * PC_KA * 8, st));
    {
        log_bands_kernel<<<(n_units * PC_STATES + 127) / 128, 128, 0, st>>>(tm, n_units, ls, ln);
        PC_LAUNCH_CHECK();
        h->launches += 1;
    }
    // Standardisation: the caller's corpus constants, or - without them - the
    // moments of this call's frames, which every chunk has to land for
    if (!host_shift) {
        PC_CUDA_TRY(cudaMemsetAsync(mom, 0, 2 * PC_XS * 8, st));
        for (int k = 0; k < c->n_chunks; k += kstep) {
            const int k1 = k + kstep < c->n_chunks ? k + kstep : c->n_chunks;
            const int64_t f_lo = c->chunk_frame[k], rows = c->chunk_frame[k1] - f_lo;
            PC_CUDA_TRY(cudaStreamWaitEvent(st, h->chunk_ev[k], 0));
            if (rows <= 0) continue;
            const int blocks = (int)std::min<int64_t>((rows + 3) / 4, (int64_t)h->sm_count * 8);
            frame_moments_kernel<<<blocks, 256, 0, st>>>(raw + (size_t)f_lo * dim, rows, dim, mom);
            PC_LAUNCH_CHECK();
            h->launches++;
        }
        moments_to_affine_kernel<<<1, 64, 0, st>>>(mom, F, dim, shift, inv_scale);
        PC_LAUNCH_CHECK();
        h->launches++;
    }
    if ((rc = launch_pack_gmm(h, mean, var, alpha, shift, inv_scale, (int)G, dim, mix, W, st))) return rc;
    const bool tc_score = h->use_tc && score_tc_supported(mix);
    for (int k = 0; k < c->n_chunks; k += kstep) {
        const int k1 = k + kstep < c->n_chunks ? k + kstep : c->n_chunks;
        PC_CUDA_TRY(cudaStreamWaitEvent(st, h->chunk_ev[k], 0));
        if ((rc = launch_prepare_frames(h, c->v, raw, 0, dim, shift, inv_scale, X, c->chunk_xtile[k],
                                        c->chunk_xtile[k1], st))) return rc;
        if (tc_score && (rc = launch_score_tc(h, c->v, X, W, mix, b, c->chunk_sitem[k], c->chunk_sitem[k1], st)))
            return rc;
    }
    if (!tc_score && (rc = launch_score_simt(h, c->v, X, W, mix, b, st))) return rc;
    if ((rc = launch_forward_backward(h, c->v, b, ls, ln, lg, c->v.scratch0, logp, iters, pt, st))) return rc;
    c->flags_lgam = nullptr;  // consumed below
    // the transition reductions need K2's outputs only: they run on the (by now idle) copy stream,
    // beside the accumulation kernel - and so does the MAX collective of the reduce hook
    PC_CUDA_TRY(cudaEventRecord(h->fork_ev, st));
    PC_CUDA_TRY(cudaStreamWaitEvent(h->copy_stream, h->fork_ev, 0));
    if ((rc = launch_transitions_max(h, c->v, logp, pt, tmax, h->copy_stream))) return rc;
    if (h->hook && h->hook(h->hook_user, 0, (void *)h->copy_stream) != 0) {
        pc_set_error("pc_em_iteration_host: the reduce hook failed (MAX)");
        return PC_ERR_CUDA;
    }
    if ((rc = launch_transitions_sum(h, c->v, logp, pt, tmax, tsum, h->copy_stream))) return rc;
    PC_CUDA_TRY(cudaEventRecord(h->join_ev, h->copy_stream));
    if (!(fix_code & 2))
    {
        if (h->use_tc && h->k3_kernel && accumulate_tcx_supported(mix)) {
            if ((rc = launch_accumulate_tcx(h, c->v, X, W, mix, b, lg, acc, true, st))) return rc;
        } else if (h->use_tc && accumulate_tc_supported(mix)) {
            if ((rc = launch_accumulate_tc(h, c->v, X, W, mix, b, lg, acc, true, st))) return rc;
        } else if ((rc = launch_accumulate_simt(h, c->v, X, W, mix, b, lg, acc, st))) {
            return rc;
        }
    }
    PC_CUDA_TRY(cudaStreamWaitEvent(st, h->join_ev, 0));
    if (h->hook && h->hook(h->hook_user, 1, (void *)st) != 0) {
        pc_set_error("pc_em_iteration_host: the reduce hook failed (SUM)");
        return PC_ERR_CUDA;
    }
    if (peer) {
        if ((rc = launch_update_params_peer(h, mix, dim, shift, inv_scale, c_cov, fix_code, mean, var, alpha, tm, st)))
            return rc;
    } else if ((rc = launch_update_params(h, n_units, mix, dim, acc, tmax, tsum, shift, inv_scale, c_cov,
                                          fix_code, mean, var, alpha, tm, st))) {
        return rc;
    }
    sum_double_kernel<<<1, 32, 0, st>>>(logp, c->v.n_utt, sum);
    PC_LAUNCH_CHECK();
    h->launches++;
    PC_CUDA_TRY(cudaMemcpyAsync(host_mean, mean, (size_t)G * dim * 8, cudaMemcpyDeviceToHost, st));
    PC_CUDA_TRY(cudaMemcpyAsync(host_var, var, (size_t)G * dim * 8, cudaMemcpyDeviceToHost, st));
    PC_CUDA_TRY(cudaMemcpyAsync(host_alpha, alpha, (size_t)G * 8, cudaMemcpyDeviceToHost, st));
    PC_CUDA_TRY(cudaMemcpyAsync(host_transmat, tm, (size_t)n_units * 25 * 8, cudaMemcpyDeviceToHost, st));
    double s = 0.0;
    int clamped = 0;
    PC_CUDA_TRY(cudaMemcpyAsync(&s, sum, 8, cudaMemcpyDeviceToHost, st));
    PC_CUDA_TRY(cudaMemcpyAsync(&clamped, h->dev_counters + PC_CNT_CLAMPED, sizeof(int), cudaMemcpyDeviceToHost, st));
    PC_CUDA_TRY(cudaStreamSynchronize(st));
    if (host_sum_logp) *host_sum_logp = s;
    if (clamped) {
        cudaMemset(h->dev_counters + PC_CNT_CLAMPED, 0, sizeof(int));
        pc_set_error("pc_em_iteration_host: %d standardised feature values exceeded +-240 and were clamped: "
                     "host_shift / host_inv_scale do not describe these frames (pc_frame_moments_host)", clamped);
        return PC_ERR_INVALID;
    }
    return PC_OK;
}

}  // extern "C"
