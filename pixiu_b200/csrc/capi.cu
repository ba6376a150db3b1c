// C ABI (include/pixiu_b200.h) over the Store: argument checking, staging of host batches,
// error translation.  No compute happens here.
#include <algorithm>
#include <chrono>
#include <cstring>

#include "index.h"
#include "store.h"

namespace pixiu {

Store::~Store() {
    if (mg_comm) mg_comm_free(mg_comm);
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    if (ev_nodes) cudaEventDestroy(ev_nodes);
    if (st) cudaStreamDestroy(st);
}

void Store::init(const pixiu_config &c) {
    cfg = c;
    knobs.from_env();
    PX_CUDA(cudaSetDevice(cfg.device));
    PX_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    PX_CUDA(cudaEventCreate(&ev0));
    PX_CUDA(cudaEventCreate(&ev1));
    PX_CUDA(cudaEventCreateWithFlags(&ev_nodes, cudaEventDisableTiming));
    PX_CUDA(cudaFree(nullptr));  // make sure the primary context exists before the driver-API calls
    d_enc.init(cfg.device, 256ull << 30);
    index.reset(new HostIndex());
}

void Store::grow_record_tables(size_t n_total, uint64_t enc_total, uint64_t tiles_total) {
    const size_t n_old = n_records();
    d_enc.ensure(enc_total + 4096);  // maps more physical memory behind the reserved range; nothing moves
    // record tables start at 1 M records / 4 M tiles (a few MB): every re-allocation is a cudaMalloc + copy +
    // stream sync + cudaFree, which stalls ingest for milliseconds
    const size_t rmin = d_enc_off.cap ? 0 : (1u << 20), tmin = d_tile_desc.cap ? 0 : (4u << 20);
    d_enc_off.reserve_keep(std::max(n_total + 1, rmin), n_old, st);
    d_enc_len.reserve_keep(std::max(n_total + 1, rmin), n_old, st);
    d_dec_len.reserve_keep(std::max(n_total + 1, rmin), n_old, st);
    d_first.reserve_keep(std::max(n_total + 1, rmin), n_old, st);
    d_tile_base.reserve_keep(std::max(n_total + 1, rmin), n_old, st);
    d_tile_desc.reserve_keep(std::max<size_t>(tiles_total + 2, tmin), n_tiles, st);
}

void Store::open_window() {
    win_open = true;
    win_R = 0;
    win_N = 0;
    h_win_rec_start.assign(1, 0u);
    pool_nth = 1;  // SuffixTree::init_prop allocates the root: first pool, 5 blocks (SuffixTree.cpp:69)
    pool_used = 5;
    for (int w = 0; w < 8; w++) win_present[w] = 0;
    win_long251 = false;
    win_present[0] |= 1u | (1u << 2);  // terminator bytes 0 and 2 ...
    win_present[251 >> 5] |= 1u << (251 & 31);  // ... and 251 are in every record
    chunk_first.push_back((uint32_t) n_records());
    chunk_count.push_back(0);
}

void Store::close_window() {
    // SuffixTree::free_prop + init_prop (PiXiuCtrl.cpp:13-17): the uncompressed window is dropped
    win_open = false;
    win_R = 0;
    win_N = 0;
    h_win_rec_start.clear();
}

}  // namespace pixiu

using pixiu::Store;

struct pixiu_store {
    Store s;
};

namespace {
// what an entry point does to the store decides when it may run (see Store::dirty / poisoned, and the multi-GPU
// phases, which keep a half-encoded batch between calls):
//   G_READ    lookups / decodes / exports: refused while a multi-GPU batch is between its phases (they would reuse its
//             scratch) and on a poisoned store
//   G_INDEX   delitem: the same (it changes the replicated index only)
//   G_PLAIN   single-GPU setitem / import / rotate / reinsert: also refused once the store is a multi-GPU shard
//             (they would desynchronise the global record numbering of the ranks)
//   G_MG      the multi-GPU phases themselves
enum GuardKind { G_READ = 0, G_INDEX = 1, G_PLAIN = 2, G_MG = 3 };

template <typename F>
int guarded(pixiu_store *h, F &&f, GuardKind kind = G_READ) {
    if (!h) return PIXIU_EINVAL;
    Store &S = h->s;
    if (S.poisoned) {
        if (S.err.rfind("store poisoned", 0) != 0) S.err = "store poisoned by an earlier failed update: " + S.err;
        return PIXIU_EPOISONED;
    }
    if (kind != G_MG && S.mg_pending) {
        S.err = "a multi-GPU setitem batch is between its phases (pixiu_mg_setitem_begin without _end)";
        return PIXIU_EINVAL;
    }
    if (kind == G_PLAIN && S.mg_world > 0) {
        S.err = "store is a multi-GPU window shard (pixiu_mg_config): use the pixiu_mg_* entry points to change it";
        return PIXIU_EINVAL;
    }
    int rc;
    try {
        cudaSetDevice(S.cfg.device);
        rc = f(S);
    } catch (const pixiu::CudaError &e) {
        S.err = e.what();
        rc = PIXIU_ECUDA;
    } catch (const std::exception &e) {
        S.err = e.what();
        rc = PIXIU_EINTERNAL;
    }
    if (rc < 0 && S.dirty) S.poisoned = true;  // the update stopped half way: nothing may run on this state
    if (rc >= 0) S.dirty = S.mg_pending != 0;  // (a multi-GPU batch between its phases is a half-built window)
    return rc;
}

// stage a packed host batch on the device: returns rebased offsets on the device
void stage(Store &S, int64_t n, const uint8_t *data, const int64_t *off, pixiu::DevBuf<uint8_t> &d_data,
           pixiu::DevBuf<int64_t> &d_off) {
    const int64_t bytes = off[n] - off[0];
    d_data.reserve_discard((size_t) bytes + 16);
    d_off.reserve_discard((size_t) n + 1);
    if (bytes) PX_CUDA(cudaMemcpyAsync(d_data.p, data + off[0], (size_t) bytes, cudaMemcpyHostToDevice, S.st));
    if (off[0] == 0) {
        PX_CUDA(cudaMemcpyAsync(d_off.p, off, (size_t) (n + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, S.st));
    } else {
        std::vector<int64_t> rel((size_t) n + 1);
        for (int64_t i = 0; i <= n; i++) rel[i] = off[i] - off[0];
        PX_CUDA(cudaMemcpyAsync(d_off.p, rel.data(), (size_t) (n + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, S.st));
        PX_CUDA(cudaStreamSynchronize(S.st));
    }
}

bool offsets_ok(int64_t n, const int64_t *off) {
    if (!off) return false;
    for (int64_t i = 0; i < n; i++)
        if (off[i + 1] < off[i]) return false;
    return true;
}

// decode `recs` into a device buffer laid out by out_off; shared by getitem / iter / decode_chunk
int decode_to(Store &S, const std::vector<uint32_t> &recs, const std::vector<uint64_t> &offs, uint8_t *out,
              int64_t out_cap, bool out_is_device, int64_t *need) {
    const uint64_t total = offs.back();
    if (need) *need = (int64_t) total;
    if ((int64_t) total > out_cap) return PIXIU_ENOSPC;
    if (recs.empty() || total == 0) return PIXIU_OK;
    if (out_is_device) {
        S.decode_records(recs, out, offs);
    } else {
        S.out_stage.reserve_discard(total + 64);
        S.decode_records(recs, S.out_stage.p, offs);
        PX_CUDA(cudaMemcpyAsync(out, S.out_stage.p, total, cudaMemcpyDeviceToHost, S.st));
        PX_CUDA(cudaStreamSynchronize(S.st));
    }
    return PIXIU_OK;
}

int getitem_common(Store &S, int64_t n, const uint8_t *keys, const int64_t *key_off, uint8_t *out, int64_t out_cap,
                   int64_t *out_off, uint8_t *found, int64_t *need, bool dev) {
    if (n < 0 || (n && (!keys || !offsets_ok(n, key_off))) || !out_off) return PIXIU_EINVAL;
    std::vector<uint32_t> rec;
    const auto t0 = std::chrono::steady_clock::now();
    pixiu::lookup_batch(S, n, keys, key_off, rec);
    const auto t1 = std::chrono::steady_clock::now();
    std::vector<uint32_t> recs;
    std::vector<uint64_t> offs(1, 0);
    recs.reserve((size_t) n);
    offs.reserve((size_t) n + 1);
    out_off[0] = 0;
    for (int64_t i = 0; i < n; i++) {
        bool f = rec[i] != 0xFFFFFFFFu;
        if (found) found[i] = f;
        if (f) {
            recs.push_back(rec[i]);
            offs.push_back(offs.back() + S.h_dec_len[rec[i]]);
        }
        out_off[i + 1] = (int64_t) offs.back();
    }
    const auto t2 = std::chrono::steady_clock::now();
    const int r = decode_to(S, recs, offs, out, out_cap, dev, need);
    if (S.knobs.trace)
        fprintf(stderr, "[getitem] n=%lld lookup %.3f ms, lists %.3f ms, decode_to %.3f ms\n", (long long) n,
                std::chrono::duration<double, std::milli>(t1 - t0).count(), std::chrono::duration<double, std::milli>(t2 - t1).count(),
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t2).count());
    return r;
}

// PiXiuCtrl::reinsert (PiXiuCtrl.cpp:88-114): the live records of a closed chunk are decoded (GPU), inserted again
// through the ordinary setitem path (they compress against the open window and replace themselves in the index,
// which tombstones the old copies) and the chunk is dropped.  Unlike the reference (bug B4: it walks all 65,535 slots
// and dereferences the NULL ones) this works for chunks of any size.  Returns the number of records moved.
int64_t reinsert_chunk(Store &S, int64_t c) {
    if (c < 0 || c >= (int64_t) S.n_chunks()) return PIXIU_EINVAL;
    if (S.win_open && c == (int64_t) S.n_chunks() - 1) return PIXIU_EINVAL;  // never the open chunk (PiXiuCtrl.cpp:26)
    if (S.chunk_dropped.size() < S.n_chunks()) S.chunk_dropped.resize(S.n_chunks(), 0);
    if (S.chunk_dropped[c]) return 0;
    // the tombstones of the moved records make c the candidate again: another chunk that was waiting for its
    // compaction must not be forgotten (the reference saves and restores Glob_Reinsert_Chunk, PiXiuCtrl.cpp:90,:112)
    const int64_t saved_candidate = S.reinsert_candidate;
    std::vector<uint32_t> recs;
    std::vector<uint64_t> offs(1, 0);
    const uint32_t g0 = S.chunk_first[c], cnt = S.chunk_count[c];
    for (uint32_t r = 0; r < cnt; r++)
        if (S.h_live[g0 + r]) {
            recs.push_back(g0 + r);
            offs.push_back(offs.back() + S.h_dec_len[g0 + r]);
        }
    const int64_t n = (int64_t) recs.size();
    if (n) {
        std::vector<uint8_t> docs(offs.back() + 1);
        int d = decode_to(S, recs, offs, docs.data(), (int64_t) docs.size(), false, nullptr);
        if (d != PIXIU_OK) return d;
        // doc = esc(k) 251 0 [esc(v) 251 2]  ->  raw key, raw value
        std::vector<uint8_t> keys, vals;
        std::vector<int64_t> koff(1, 0), voff(1, 0);
        keys.reserve(docs.size());
        vals.reserve(docs.size());
        for (int64_t i = 0; i < n; i++) {
            const uint8_t *p = docs.data() + offs[i];
            const uint32_t len = (uint32_t) (offs[i + 1] - offs[i]);
            bool in_val = false, closed = false;
            for (uint32_t k = 0; k < len;) {
                if (p[k] != 251) {
                    (in_val ? vals : keys).push_back(p[k++]);
                    continue;
                }
                if (k + 1 >= len) return PIXIU_ECORRUPT;
                const uint8_t nx = p[k + 1];
                if (nx == 251) (in_val ? vals : keys).push_back(251);
                else if (nx == 0 && !in_val) in_val = true;
                else if (nx == 2 && in_val) closed = true;
                else return PIXIU_ECORRUPT;
                k += 2;
            }
            if (!in_val || (closed != ((int64_t) vals.size() > voff.back()))) return PIXIU_ECORRUPT;
            koff.push_back((int64_t) keys.size());
            voff.push_back((int64_t) vals.size());
        }
        keys.push_back(0);
        vals.push_back(0);
        std::vector<int32_t> rc((size_t) n);
        stage(S, n, keys.data(), koff.data(), S.in_keys, S.in_koff);
        stage(S, n, vals.data(), voff.data(), S.in_vals, S.in_voff);
        const int64_t raw0 = S.raw_bytes, doc0 = S.doc_bytes;
        int r = S.setitem_batch(n, S.in_keys.p, S.in_koff.p, S.in_vals.p, S.in_voff.p, keys.data(), koff.data(), voff.data(),
                                rc.data(), nullptr);
        if (r != PIXIU_OK) return r;
        S.raw_bytes = raw0;  // moved, not added
        S.doc_bytes = doc0;
        for (int64_t i = 0; i < n; i++)
            if (rc[i] != PIXIU_CBT_SET_REPLACE) {
                S.err = "reinsert: a live record did not replace itself";
                return PIXIU_EINTERNAL;
            }
    }
    if (S.chunk_dropped.size() < S.n_chunks()) S.chunk_dropped.resize(S.n_chunks(), 0);
    S.chunk_dropped[c] = 1;
    S.reinserted_records += n;
    const uint32_t gl = g0 + cnt - 1;
    if (cnt) S.reclaimable_bytes += (int64_t) (S.h_enc_off[gl] + S.h_enc_len[gl] - S.h_enc_off[g0]);
    S.reinsert_candidate = saved_candidate == c ? -1 : saved_candidate;
    return n;
}

// the reference's trigger (PiXiuCtrl.cpp:7-8,:26-29,:64-67): a chunk that a tombstone left below 0.8 x 65,535 live
// records is remembered; it is re-inserted as soon as it is not the open chunk and less than half of it is live
int64_t maybe_reinsert(Store &S) {
    if (S.mg_world > 0) return 0;  // (a multi-GPU shard is only changed through the pixiu_mg_* phases)
    const int64_t c = S.reinsert_candidate;
    if (c < 0 || c >= (int64_t) S.n_chunks()) return 0;
    if (S.win_open && c == (int64_t) S.n_chunks() - 1) return 0;
    if (c < (int64_t) S.chunk_dropped.size() && S.chunk_dropped[c]) return 0;
    const uint32_t live = c < (int64_t) S.chunk_live.size() ? S.chunk_live[c] : 0;
    if (!(live < 0.5 * S.chunk_count[c])) return 0;
    return reinsert_chunk(S, c);
}
}  // namespace

extern "C" {

void pixiu_default_config(pixiu_config *cfg) {
    if (!cfg) return;
    memset(cfg, 0, sizeof(*cfg));
    cfg->device = 0;
    cfg->rotate_policy = PIXIU_ROTATE_REFERENCE;
    cfg->window_bytes = 12500000;
    cfg->strict251 = 0;
}

pixiu_store *pixiu_create(const pixiu_config *cfg) {
    pixiu_config c;
    if (cfg) c = *cfg;
    else pixiu_default_config(&c);
    if (c.window_bytes <= 0) c.window_bytes = 12500000;
    pixiu_store *h = nullptr;
    try {
        h = new pixiu_store();
        h->s.init(c);
        return h;
    } catch (const std::exception &e) {
        fprintf(stderr, "pixiu_create: %s\n", e.what());
        delete h;
        return nullptr;
    }
}

void pixiu_destroy(pixiu_store *h) {
    if (!h) return;
    cudaSetDevice(h->s.cfg.device);
    cudaDeviceSynchronize();
    delete h;
}

const char *pixiu_last_error(const pixiu_store *h) { return h ? h->s.err.c_str() : "null store"; }

int64_t pixiu_reinsert_chunk(pixiu_store *h, int64_t chunk) {
    int64_t moved = 0;
    int rc = guarded(h, [&](Store &S) -> int {
        moved = reinsert_chunk(S, chunk);
        return moved < 0 ? (int) moved : PIXIU_OK;
    }, G_PLAIN);
    return rc == PIXIU_OK ? moved : rc;
}

int pixiu_chunk_info(pixiu_store *h, int64_t chunk, int64_t *total, int64_t *live, int32_t *dropped) {
    if (!h || chunk < 0 || chunk >= (int64_t) h->s.n_chunks()) return PIXIU_EINVAL;
    Store &S = h->s;
    if (total) *total = S.chunk_count[chunk];
    if (live) *live = chunk < (int64_t) S.chunk_live.size() ? S.chunk_live[chunk] : 0;
    if (dropped) *dropped = chunk < (int64_t) S.chunk_dropped.size() ? S.chunk_dropped[chunk] : 0;
    return PIXIU_OK;
}

int pixiu_get_stats(pixiu_store *h, pixiu_stats *o) {
    if (!h || !o) return PIXIU_EINVAL;
    Store &S = h->s;
    o->records = (int64_t) S.n_records();
    o->live_records = S.live_records;
    o->chunks = (int64_t) S.n_chunks();
    o->raw_bytes = S.raw_bytes;
    o->doc_bytes = S.doc_bytes;
    o->encoded_bytes = (int64_t) S.enc_bytes;
    o->window_bytes = S.win_open ? S.win_N : 0;
    o->kernel_launches = S.launches;
    o->last_setitem_gpu_ms = S.last_set_ms;
    o->last_getitem_gpu_ms = S.last_get_ms;
    o->last_lookup_gpu_ms = S.last_lookup_ms;
    o->reinserted_records = S.reinserted_records;
    o->reclaimable_bytes = S.reclaimable_bytes;
    o->index_key_arena_bytes = (int64_t) S.index->key_arena_bytes();
    o->index_host_bytes = (int64_t) S.index->host_bytes();
    o->index_device_bytes = (int64_t) S.index->device_bytes();
    o->table_device_bytes = (int64_t) (S.d_enc_off.cap * 8 + (S.d_enc_len.cap + S.d_dec_len.cap + S.d_first.cap + S.d_tile_base.cap +
                                                               S.d_tile_desc.cap) * 4);
    return PIXIU_OK;
}

int pixiu_setitem_batch(pixiu_store *h, int64_t n, const uint8_t *keys, const int64_t *key_off, const uint8_t *vals,
                        const int64_t *val_off, int32_t *rc, int32_t *saved) {
    return guarded(h, [&](Store &S) -> int {
        if (n < 0 || (n && (!keys || !offsets_ok(n, key_off) || !offsets_ok(n, val_off)))) return PIXIU_EINVAL;
        if (n == 0) return PIXIU_OK;
        if (val_off[n] > val_off[0] && !vals) return PIXIU_EINVAL;
        if (S.cfg.auto_reinsert) {
            int64_t m = maybe_reinsert(S);
            if (m < 0) return (int) m;
        }
        stage(S, n, keys, key_off, S.in_keys, S.in_koff);
        stage(S, n, vals, val_off, S.in_vals, S.in_voff);
        return S.setitem_batch(n, S.in_keys.p, S.in_koff.p, S.in_vals.p, S.in_voff.p, keys, key_off, val_off, rc, saved);
    }, G_PLAIN);
}

int pixiu_setitem_batch_dev(pixiu_store *h, int64_t n, const uint8_t *d_keys, const int64_t *d_key_off,
                            const uint8_t *d_vals, const int64_t *d_val_off, int32_t *rc, int32_t *saved) {
    return guarded(h, [&](Store &S) -> int {
        if (n < 0 || (n && (!d_keys || !d_key_off || !d_val_off))) return PIXIU_EINVAL;
        if (n == 0) return PIXIU_OK;
        // the host index needs the keys (small next to the values): copy keys + offsets back
        std::vector<int64_t> koff((size_t) n + 1), voff((size_t) n + 1);
        PX_CUDA(cudaMemcpyAsync(koff.data(), d_key_off, (size_t) (n + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, S.st));
        PX_CUDA(cudaMemcpyAsync(voff.data(), d_val_off, (size_t) (n + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, S.st));
        PX_CUDA(cudaStreamSynchronize(S.st));
        if (koff[0] != 0 || voff[0] != 0 || !offsets_ok(n, koff.data()) || !offsets_ok(n, voff.data())) return PIXIU_EINVAL;
        if (S.cfg.auto_reinsert) {
            int64_t m = maybe_reinsert(S);
            if (m < 0) return (int) m;
        }
        std::vector<uint8_t> hk((size_t) koff[n] + 1);
        PX_CUDA(cudaMemcpyAsync(hk.data(), d_keys, (size_t) koff[n], cudaMemcpyDeviceToHost, S.st));
        PX_CUDA(cudaStreamSynchronize(S.st));
        return S.setitem_batch(n, d_keys, d_key_off, d_vals, d_val_off, hk.data(), koff.data(), voff.data(), rc, saved);
    }, G_PLAIN);
}

int pixiu_contains_batch(pixiu_store *h, int64_t n, const uint8_t *keys, const int64_t *key_off, uint8_t *found) {
    return guarded(h, [&](Store &S) -> int {
        if (n < 0 || (n && (!keys || !offsets_ok(n, key_off) || !found))) return PIXIU_EINVAL;
        std::vector<uint32_t> rec;
        pixiu::lookup_batch(S, n, keys, key_off, rec);
        for (int64_t i = 0; i < n; i++) found[i] = rec[i] != 0xFFFFFFFFu;
        return PIXIU_OK;
    });
}

int pixiu_contains_batch_dev(pixiu_store *h, int64_t n, const uint8_t *d_keys, const int64_t *d_key_off, uint8_t *d_found) {
    return guarded(h, [&](Store &S) -> int {
        if (n < 0 || n > 0x7fffffff || (n && (!d_keys || !d_key_off || !d_found))) return PIXIU_EINVAL;
        pixiu::contains_batch_dev(S, n, d_keys, d_key_off, d_found);
        return PIXIU_OK;
    });
}

int pixiu_debug_index_depth(pixiu_store *h, int64_t n, const uint8_t *keys, const int64_t *key_off, int32_t *depth) {
    return guarded(h, [&](Store &S) -> int {
        if (n < 0 || (n && (!keys || !offsets_ok(n, key_off) || !depth))) return PIXIU_EINVAL;
        pixiu::index_depths(S, n, keys, key_off, depth);
        return PIXIU_OK;
    });
}

int pixiu_delitem_batch(pixiu_store *h, int64_t n, const uint8_t *keys, const int64_t *key_off, int32_t *rc) {
    return guarded(h, [&](Store &S) -> int {
        if (n < 0 || (n && (!keys || !offsets_ok(n, key_off)))) return PIXIU_EINVAL;
        std::vector<uint8_t> q;
        for (int64_t i = 0; i < n; i++) {
            pixiu::escape_key(keys + key_off[i], (size_t) (key_off[i + 1] - key_off[i]), q);
            // (PiXiuCtrl::delitem, PiXiuCtrl.cpp:63-69: the compaction trigger is looked at before every delete)
            if (S.cfg.auto_reinsert) {
                int64_t m = maybe_reinsert(S);
                if (m < 0) return (int) m;
            }
            int64_t r = S.index->del(q.data(), (uint32_t) q.size());
            if (r >= 0) S.tombstone((uint32_t) r);
            if (rc) rc[i] = r >= 0 ? 0 : PIXIU_CBT_DEL_NOT_FOUND;
        }
        return PIXIU_OK;
    }, G_INDEX);
}

int pixiu_getitem_batch(pixiu_store *h, int64_t n, const uint8_t *keys, const int64_t *key_off, uint8_t *out,
                        int64_t out_cap, int64_t *out_off, uint8_t *found, int64_t *need) {
    return guarded(h, [&](Store &S) -> int {
        return getitem_common(S, n, keys, key_off, out, out_cap, out_off, found, need, false);
    });
}

int pixiu_getitem_batch_dev(pixiu_store *h, int64_t n, const uint8_t *keys, const int64_t *key_off, uint8_t *d_out,
                            int64_t out_cap, int64_t *out_off, uint8_t *found, int64_t *need) {
    return guarded(h, [&](Store &S) -> int {
        return getitem_common(S, n, keys, key_off, d_out, out_cap, out_off, found, need, true);
    });
}

int pixiu_iter(pixiu_store *h, const uint8_t *prefix, int64_t prefix_len, uint8_t *out, int64_t out_cap,
               int64_t *out_off, int64_t off_cap, int64_t *count, int64_t *need) {
    return guarded(h, [&](Store &S) -> int {
        if (prefix_len < 0 || (prefix_len && !prefix) || !count) return PIXIU_EINVAL;
        std::vector<uint8_t> q;
        pixiu::escape_key(prefix, (size_t) prefix_len, q, false);
        std::vector<uint32_t> recs;
        pixiu::iter_prefix(S, q.data(), (uint32_t) q.size(), recs);   // the walk runs on the device (k_iter_prefix)
        *count = (int64_t) recs.size();
        std::vector<uint64_t> offs(1, 0);
        for (uint32_t g : recs) offs.push_back(offs.back() + S.h_dec_len[g]);
        if (need) *need = (int64_t) offs.back();
        if ((int64_t) recs.size() + 1 > off_cap || (int64_t) offs.back() > out_cap) return PIXIU_ENOSPC;
        if (!out_off) return PIXIU_EINVAL;
        for (size_t i = 0; i < offs.size(); i++) out_off[i] = (int64_t) offs[i];
        return decode_to(S, recs, offs, out, out_cap, false, need);
    });
}

int pixiu_encoded_view(pixiu_store *h, int64_t chunk, int64_t idx, uint8_t *out, int64_t out_cap) {
    return guarded(h, [&](Store &S) -> int {
        if (chunk < 0 || chunk >= (int64_t) S.n_chunks() || idx < 0 || idx >= (int64_t) S.chunk_count[chunk]) return PIXIU_EINVAL;
        uint32_t g = S.chunk_first[chunk] + (uint32_t) idx;
        uint32_t len = S.h_enc_len[g];
        if ((int64_t) len > out_cap) return PIXIU_ENOSPC;
        PX_CUDA(cudaMemcpyAsync(out, S.d_enc.ptr() + S.h_enc_off[g], len, cudaMemcpyDeviceToHost, S.st));
        PX_CUDA(cudaStreamSynchronize(S.st));
        return (int) len;
    });
}

int pixiu_record_location(pixiu_store *h, int64_t record, int64_t *chunk, int64_t *idx) {
    if (!h || record < 0 || record >= (int64_t) h->s.n_records() || !chunk || !idx) return PIXIU_EINVAL;
    Store &S = h->s;
    uint32_t f = S.h_first[record];
    size_t c = std::upper_bound(S.chunk_first.begin(), S.chunk_first.end(), (uint32_t) record) - S.chunk_first.begin() - 1;
    *chunk = (int64_t) c;
    *idx = record - f;
    return PIXIU_OK;
}

int pixiu_decode_chunk(pixiu_store *h, int64_t chunk, uint8_t *out, int64_t out_cap, int64_t *out_off, int64_t *need) {
    return guarded(h, [&](Store &S) -> int {
        if (chunk < 0 || chunk >= (int64_t) S.n_chunks() || !out_off) return PIXIU_EINVAL;
        std::vector<uint32_t> recs;
        std::vector<uint64_t> offs(1, 0);
        for (uint32_t r = 0; r < S.chunk_count[chunk]; r++) {
            uint32_t g = S.chunk_first[chunk] + r;
            recs.push_back(g);
            offs.push_back(offs.back() + S.h_dec_len[g]);
        }
        for (size_t i = 0; i < offs.size(); i++) out_off[i] = (int64_t) offs[i];
        return decode_to(S, recs, offs, out, out_cap, false, need);
    });
}

int64_t pixiu_import_chunk(pixiu_store *h, int64_t n, const uint8_t *enc, const int64_t *enc_off) {
    int64_t chunk_id = -1;
    int rc = guarded(h, [&](Store &S) -> int {
        if (!enc || !offsets_ok(n, enc_off)) return PIXIU_EINVAL;
        int64_t c = S.import_chunk(n, enc, enc_off);
        if (c < 0) return (int) c;
        chunk_id = c;
        // index the keys: decode the chunk once and read each record's escaped key (up to 251,0)
        std::vector<uint32_t> recs;
        std::vector<uint64_t> offs(1, 0);
        for (uint32_t r = 0; r < S.chunk_count[c]; r++) {
            recs.push_back(S.chunk_first[c] + r);
            offs.push_back(offs.back() + S.h_dec_len[recs.back()]);
        }
        std::vector<uint8_t> dec(offs.back() + 1);
        int d = decode_to(S, recs, offs, dec.data(), (int64_t) dec.size(), false, nullptr);
        if (d != PIXIU_OK) return d;
        for (size_t r = 0; r < recs.size(); r++) {
            const uint8_t *p = dec.data() + offs[r];
            uint32_t len = (uint32_t) (offs[r + 1] - offs[r]), k = 0;
            while (k + 1 < len && !(p[k] == 251 && p[k + 1] == 0)) k += (p[k] == 251) ? 2 : 1;
            if (k + 1 >= len) return PIXIU_ECORRUPT;
            int64_t old = S.index->set(p, k + 2, recs[r]);
            if (old >= 0) S.tombstone((uint32_t) old);
            S.note_live(recs[r]);
            S.doc_bytes += len;
        }
        return PIXIU_OK;
    }, G_PLAIN);
    return rc == PIXIU_OK ? chunk_id : rc;
}

int64_t pixiu_import_chunk_raw(pixiu_store *h, int64_t n, const uint8_t *enc, const int64_t *enc_off) {
    int64_t chunk_id = -1;
    int rc = guarded(h, [&](Store &S) -> int {
        if (!enc || !offsets_ok(n, enc_off)) return PIXIU_EINVAL;
        const size_t g0 = S.n_records();
        int64_t c = S.import_chunk(n, enc, enc_off);
        if (c < 0) return (int) c;
        chunk_id = c;
        for (size_t g = g0; g < S.n_records(); g++) {   // nothing is live until an index says so
            S.h_live[g] = 0;
            S.doc_bytes += S.h_dec_len[g];
        }
        return PIXIU_OK;
    }, G_PLAIN);
    return rc == PIXIU_OK ? chunk_id : rc;
}

int pixiu_export_index(pixiu_store *h, uint8_t *out, int64_t out_cap, int64_t *need) {
    return guarded(h, [&](Store &S) -> int {
        std::vector<uint8_t> blob;
        S.index->save(blob);
        if (need) *need = (int64_t) blob.size();
        if ((int64_t) blob.size() > out_cap || !out) return PIXIU_ENOSPC;
        memcpy(out, blob.data(), blob.size());
        return PIXIU_OK;
    });
}

int pixiu_import_index(pixiu_store *h, const uint8_t *blob, int64_t size) {
    return guarded(h, [&](Store &S) -> int {
        if (!blob || size <= 0 || S.index->size() != 0) return PIXIU_EINVAL;
        std::vector<uint32_t> live;
        if (!S.index->load(blob, (size_t) size, (uint32_t) S.n_records(), live)) return PIXIU_ECORRUPT;
        S.dirty = true;
        for (uint32_t g : live) {
            if (S.h_live[g]) return PIXIU_ECORRUPT;   // two leaves on one record
            S.h_live[g] = 1;
            S.note_live(g);
        }
        return PIXIU_OK;
    }, G_PLAIN);
}

int pixiu_mg_config(pixiu_store *h, int rank, int world) {
    if (!h || world < 1 || rank < 0 || rank >= world || h->s.n_records() != 0) return PIXIU_EINVAL;
    h->s.mg_rank = rank;
    h->s.mg_world = world;
    return PIXIU_OK;
}

int pixiu_mg_setitem_begin(pixiu_store *h, int64_t n, const uint8_t *keys, const int64_t *key_off, const uint8_t *vals,
                           const int64_t *val_off, uint32_t **d_m, int64_t *count) {
    return guarded(h, [&](Store &S) -> int {
        if (n <= 0 || !keys || !offsets_ok(n, key_off) || !offsets_ok(n, val_off) || !d_m || !count) return PIXIU_EINVAL;
        if (val_off[n] > val_off[0] && !vals) return PIXIU_EINVAL;
        stage(S, n, keys, key_off, S.mg_in_keys, S.mg_in_koff);
        stage(S, n, vals, val_off, S.mg_in_vals, S.mg_in_voff);
        return S.mg_begin(n, S.mg_in_keys.p, S.mg_in_koff.p, S.mg_in_vals.p, S.mg_in_voff.p, keys, key_off, val_off, d_m, count);
    }, G_MG);
}

int pixiu_mg_setitem_mid(pixiu_store *h, uint32_t **d_cand, int64_t *count) {
    return guarded(h, [&](Store &S) -> int {
        if (!d_cand || !count) return PIXIU_EINVAL;
        return S.mg_mid(d_cand, count);
    }, G_MG);
}

int pixiu_mg_setitem_end(pixiu_store *h, int32_t *rc, int32_t *saved) {
    return guarded(h, [&](Store &S) -> int { return S.mg_end(rc, saved); }, G_MG);
}

int pixiu_mg_unique_id(pixiu_store *h, uint8_t *id) {
    if (!id) return PIXIU_EINVAL;
    std::string e;
    const int r = pixiu::mg_unique_id(id, e);
    if (r != PIXIU_OK) {
        if (h) h->s.err = e;
        else fprintf(stderr, "pixiu_mg_unique_id: %s\n", e.c_str());
    }
    return r;
}

int pixiu_mg_comm_init(pixiu_store *h, int rank, int world, const uint8_t *id) {
    return guarded(h, [&](Store &S) -> int {
        if (!id) return PIXIU_EINVAL;
        return pixiu::mg_comm_init(S, rank, world, id);
    }, G_MG);
}

int pixiu_mg_setitem_batch(pixiu_store *h, int64_t n, const uint8_t *keys, const int64_t *key_off, const uint8_t *vals,
                           const int64_t *val_off, int32_t *rc, int32_t *saved) {
    return guarded(h, [&](Store &S) -> int {
        if (n <= 0 || !keys || !offsets_ok(n, key_off) || !offsets_ok(n, val_off)) return PIXIU_EINVAL;
        if (val_off[n] > val_off[0] && !vals) return PIXIU_EINVAL;
        stage(S, n, keys, key_off, S.mg_in_keys, S.mg_in_koff);
        stage(S, n, vals, val_off, S.mg_in_vals, S.mg_in_voff);
        return pixiu::mg_setitem_nccl(S, n, S.mg_in_keys.p, S.mg_in_koff.p, S.mg_in_vals.p, S.mg_in_voff.p, keys, key_off,
                                      val_off, rc, saved);
    }, G_MG);
}

int pixiu_mg_get_stats(pixiu_store *h, pixiu_mg_stats *o) {
    if (!h || !o) return PIXIU_EINVAL;
    pixiu::mg_comm_stats(h->s, o);
    return PIXIU_OK;
}

int pixiu_profile_enable(pixiu_store *h, int on) {
    if (!h) return PIXIU_EINVAL;
    h->s.prof.reset();
    h->s.prof.on = on != 0;
    return PIXIU_OK;
}

int pixiu_profile_get(pixiu_store *h, int cls, const char **name, double *ms, double *bytes, int64_t *launches) {
    if (!h || cls < 0) return PIXIU_EINVAL;
    if (cls >= pixiu::PC_COUNT) return 1;  /* past the end */
    if (name) *name = pixiu::prof_class_name(cls);
    if (ms) *ms = h->s.prof.ms[cls];
    if (bytes) *bytes = h->s.prof.bytes[cls];
    if (launches) *launches = h->s.prof.launches[cls];
    return PIXIU_OK;
}

void *pixiu_stream(pixiu_store *h) { return h ? (void *) h->s.st : nullptr; }

int pixiu_export_chunk(pixiu_store *h, int64_t chunk, uint8_t *out, int64_t out_cap, int64_t *out_off, int64_t *count,
                       int64_t *need) {
    return guarded(h, [&](Store &S) -> int {
        if (chunk < 0 || chunk >= (int64_t) S.n_chunks() || !count) return PIXIU_EINVAL;
        const uint32_t g0 = S.chunk_first[chunk], n = S.chunk_count[chunk];
        *count = n;
        uint64_t total = 0;
        for (uint32_t r = 0; r < n; r++) total += S.h_enc_len[g0 + r];
        if (need) *need = (int64_t) total;
        if ((int64_t) total > out_cap || !out || !out_off) return PIXIU_ENOSPC;
        out_off[0] = 0;
        for (uint32_t r = 0; r < n; r++) out_off[r + 1] = out_off[r] + S.h_enc_len[g0 + r];
        // the records of a chunk are contiguous in the compressed arena
        if (total) PX_CUDA(cudaMemcpyAsync(out, S.d_enc.ptr() + S.h_enc_off[g0], total, cudaMemcpyDeviceToHost, S.st));
        PX_CUDA(cudaStreamSynchronize(S.st));
        return PIXIU_OK;
    });
}

int pixiu_rotate(pixiu_store *h) {
    return guarded(h, [&](Store &S) -> int {
        if (S.win_open) S.close_window();
        return PIXIU_OK;
    }, G_PLAIN);
}

int pixiu_reserve(pixiu_store *h, int64_t encoded_bytes) {
    return guarded(h, [&](Store &S) -> int {
        if (encoded_bytes < 0) return PIXIU_EINVAL;
        S.d_enc.ensure(S.enc_bytes + (uint64_t) encoded_bytes + 4096);
        return PIXIU_OK;
    }, G_READ);
}

}  // extern "C"

// ---------------------------------------------------------------------------------
// debug / test hooks (declared in include/pixiu_b200_debug.h; used by tests/ only)
// ---------------------------------------------------------------------------------
extern "C" {

// GPU radix sort of host (key,value) pairs on bits [0,end_bit); vals == NULL sorts indices.
int pixiu_debug_sort_pairs(int device, uint64_t *keys, uint32_t *vals, int64_t n, int end_bit, uint32_t *vals_out) {
    try {
        PX_CUDA(cudaSetDevice(device));
        pixiu::DevBuf<uint64_t> k0, k1;
        pixiu::DevBuf<uint32_t> v0, v1, err;
        pixiu::RadixSortTemp tmp;
        k0.reserve_discard(n);
        k1.reserve_discard(n);
        v0.reserve_discard(n);
        v1.reserve_discard(n);
        err.reserve_discard(4);
        PX_CUDA(cudaMemset(err.p, 0, 16));
        PX_CUDA(cudaMemcpy(k0.p, keys, n * sizeof(uint64_t), cudaMemcpyHostToDevice));
        if (vals) PX_CUDA(cudaMemcpy(v0.p, vals, n * sizeof(uint32_t), cudaMemcpyHostToDevice));
        int cur = pixiu::radix_sort_pairs<uint64_t>(k0.p, k1.p, v0.p, v1.p, (uint32_t) n, 0, end_bit, vals == nullptr,
                                                    tmp, err.p, 0);
        PX_CUDA(cudaDeviceSynchronize());
        PX_CUDA(cudaMemcpy(keys, cur ? k1.p : k0.p, n * sizeof(uint64_t), cudaMemcpyDeviceToHost));
        PX_CUDA(cudaMemcpy(vals_out, cur ? v1.p : v0.p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
        uint32_t e = 0;
        PX_CUDA(cudaMemcpy(&e, err.p, 4, cudaMemcpyDeviceToHost));
        return e ? PIXIU_EINTERNAL : PIXIU_OK;
    } catch (const std::exception &e) {
        fprintf(stderr, "pixiu_debug_sort_pairs: %s\n", e.what());
        return PIXIU_ECUDA;
    }
}

// plain cudaMemcpy for tests that play the role of the collective: kind 1 = device->host, 2 = host->device
int pixiu_debug_memcpy(void *dst, const void *src, int64_t bytes, int kind) {
    cudaError_t e = cudaMemcpy(dst, src, (size_t) bytes, kind == 1 ? cudaMemcpyDeviceToHost : cudaMemcpyHostToDevice);
    return e == cudaSuccess ? PIXIU_OK : PIXIU_ECUDA;
}

// change a tuning / test knob of a live store (Knobs::set, store.h); returns PIXIU_EINVAL for an unknown name
int pixiu_debug_set_knob(pixiu_store *h, const char *name, int64_t value) {
    if (!h || !name) return PIXIU_EINVAL;
    return h->s.knobs.set(name, value) ? PIXIU_OK : PIXIU_EINVAL;
}

// pending (polled) pieces and drain passes of the last decode call
int pixiu_debug_decode_counters(pixiu_store *h, int64_t *pending_pieces, int64_t *drains) {
    if (!h) return PIXIU_EINVAL;
    if (pending_pieces) *pending_pieces = (int64_t) h->s.last_pending_pieces;
    if (drains) *drains = (int64_t) h->s.last_drains;
    return PIXIU_OK;
}

// arena state the reference's suffix tree would have for the open window (MemPool::nth, used_num)
int pixiu_debug_pool_state(pixiu_store *h, int32_t *nth, int32_t *used) {
    if (!h || !nth || !used) return PIXIU_EINVAL;
    *nth = (int32_t) h->s.pool_nth;
    *used = (int32_t) h->s.pool_used;
    return PIXIU_OK;
}

// copies an internal array of the last encode (open window) to the host:
// name in {"sa","rank","lcp","reach","off","prevp","nextp"} -> u32[win_N]; {"text","flagp","flagc"} -> u8[win_N]
int64_t pixiu_debug_window_array(pixiu_store *h, const char *name, void *out, int64_t cap_bytes) {
    if (!h || !name) return PIXIU_EINVAL;
    Store &S = h->s;
    std::string nm(name);
    const void *src = nullptr;
    size_t esz = 4;
    if (nm == "sa") src = S.es.sa.p;
    else if (nm == "rank") src = S.es.rank.p;
    else if (nm == "lcp") {
        if (S.es.tree.leaf) {   // the lcp halves of the interleaved {sa, lcp} leaves
            if ((int64_t) ((size_t) S.win_N * 4) > cap_bytes) return PIXIU_ENOSPC;
            cudaSetDevice(S.cfg.device);
            if (cudaMemcpy2D(out, 4, reinterpret_cast<const char *>(S.es.tree.leaf) + 4, 8, 4, S.win_N, cudaMemcpyDeviceToHost) != cudaSuccess)
                return PIXIU_ECUDA;
            for (uint32_t i = 0; i < S.win_N; i++) static_cast<uint32_t *>(out)[i] &= 0xFFFFu;   // (upper half: dist[sa])
            return (int64_t) S.win_N;
        }
        src = S.es.lcp.p;
    }
    else if (nm == "reach") src = S.es.reach.p;
    else if (nm == "off") src = S.es.off.p;
    else if (nm == "prevp") src = S.es.prevp.p;
    else if (nm == "nextp") src = S.es.nextp.p;
    else if (nm == "text") { src = S.w_text.p; esz = 1; }
    else if (nm == "flagp") { src = S.es.flagp.p; esz = 1; }
    else if (nm == "flagc") { src = S.es.flagc.p; esz = 1; }
    else if (nm == "dist") { src = S.w_dist.p; esz = 2; }
    if (!src) return PIXIU_EINVAL;
    size_t bytes = (size_t) S.win_N * esz;
    if ((int64_t) bytes > cap_bytes) return PIXIU_ENOSPC;
    cudaSetDevice(S.cfg.device);
    if (cudaMemcpy(out, src, bytes, cudaMemcpyDeviceToHost) != cudaSuccess) return PIXIU_ECUDA;
    return (int64_t) S.win_N;
}

}  // extern "C"
