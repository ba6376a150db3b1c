// CritBit index: host maintenance + GPU batched lookup (see index.h).
#include "index.h"

#include <algorithm>
#include <cstring>

#include "scan.cuh"
#include "store.h"

namespace pixiu {

void escape_key(const uint8_t *k, size_t n, std::vector<uint8_t> &out, bool terminator) {
    out.clear();
    out.reserve(n + 8);
    for (size_t i = 0; i < n; i++) {
        out.push_back(k[i]);
        if (k[i] == 251) out.push_back(251);
    }
    if (terminator) {
        out.push_back(251);
        out.push_back(0);
    }
}

// ---------------------------------------------------------------------------------
// host tree
// ---------------------------------------------------------------------------------
int32_t HostIndex::new_leaf(const uint8_t *q, uint32_t qlen, uint32_t rec) {
    int32_t s;
    if (!free_leaf.empty()) {
        s = free_leaf.back();
        free_leaf.pop_back();
    } else {
        s = (int32_t) leaf_rec.size();
        leaf_rec.push_back(0);
        leaf_klen.push_back(0);
        leaf_koff.push_back(0);
    }
    leaf_rec[s] = rec;
    leaf_klen[s] = qlen;
    leaf_koff[s] = arena.size();
    arena.insert(arena.end(), q, q + qlen);
    return s;
}

int32_t HostIndex::new_inner() {
    if (!free_inner.empty()) {
        int32_t s = free_inner.back();
        free_inner.pop_back();
        return s;
    }
    child[0].push_back(0);
    child[1].push_back(0);
    diff_at.push_back(0);
    mask.push_back(0);
    return (int32_t) diff_at.size() - 1;
}

int64_t HostIndex::get(const uint8_t *q, uint32_t qlen) const {
    if (!has_root) return -1;
    int32_t p = root;
    while (p >= 0) p = child[dir_of(p, q, qlen)][p];
    int32_t s = ~p;
    if (leaf_klen[s] == qlen && memcmp(arena.data() + leaf_koff[s], q, qlen) == 0) return leaf_rec[s];
    return -1;
}

int64_t HostIndex::set(const uint8_t *q, uint32_t qlen, uint32_t rec) {
    dirty = true;
    if (!has_root) {
        root = ~new_leaf(q, qlen, rec);
        has_root = true;
        n_live = 1;
        return -1;
    }
    int32_t p = root;
    while (p >= 0) p = child[dir_of(p, q, qlen)][p];
    int32_t s = ~p;
    const uint8_t *lk = arena.data() + leaf_koff[s];
    uint32_t ll = leaf_klen[s], m = std::min(ll, qlen), diff = 0;
    while (diff < m && lk[diff] == q[diff]) diff++;
    if (diff == qlen && diff == ll) {  // same key: repoint the leaf (replace(), CritBitTree.cpp:32-43)
        int64_t old = leaf_rec[s];
        leaf_rec[s] = rec;
        return old;
    }
    // insert() (CritBitTree.cpp:45-92); reference bug B5 (record lost when the first difference
    // directly follows a matched 251) is not reproduced.
    uint8_t a = diff < ll ? lk[diff] : 0, b = diff < qlen ? q[diff] : 0;
    uint8_t mk = (uint8_t) (a ^ b);
    mk |= mk >> 1;
    mk |= mk >> 2;
    mk |= mk >> 4;
    mk = (uint8_t) ((mk & ~(mk >> 1)) ^ 0xFF);
    int dir = (1 + (mk | b)) >> 8;
    int32_t nl = new_leaf(q, qlen, rec);
    int32_t ni = new_inner();
    diff_at[ni] = (uint16_t) diff;
    mask[ni] = mk;
    child[dir][ni] = ~nl;
    int32_t parent = -1, pdir = 0, cur = root;
    while (cur >= 0) {
        if (diff_at[cur] > diff || (diff_at[cur] == diff && mask[cur] > mk)) break;
        pdir = dir_of(cur, q, qlen);
        parent = cur;
        cur = child[pdir][cur];
    }
    child[1 - dir][ni] = cur;
    if (parent < 0) root = ni;
    else child[pdir][parent] = ni;
    n_live++;
    return -1;
}

int64_t HostIndex::del(const uint8_t *q, uint32_t qlen) {
    if (!has_root) return -1;
    int32_t p = root, pa = -1, gr = -1, pdir = 0;
    while (p >= 0) {
        gr = pa;
        pa = p;
        pdir = dir_of(p, q, qlen);
        p = child[pdir][p];
    }
    int32_t s = ~p;
    if (!(leaf_klen[s] == qlen && memcmp(arena.data() + leaf_koff[s], q, qlen) == 0)) return -1;
    dirty = true;
    int64_t rec = leaf_rec[s];
    if (pa < 0) {
        has_root = false;
    } else {
        int32_t sib = child[1 - pdir][pa];
        if (gr < 0) root = sib;
        else child[child[0][gr] == pa ? 0 : 1][gr] = sib;
        free_inner.push_back(pa);
    }
    free_leaf.push_back(s);
    n_live--;
    return rec;
}

void HostIndex::iter_rec(int32_t p, const uint8_t *pre, uint32_t plen, bool include_all, bool &harvest, bool &stop,
                         std::vector<uint32_t> &out) const {
    if (stop) return;
    if (p < 0) {
        int32_t s = ~p;
        if (!harvest) {
            if (leaf_klen[s] < plen || memcmp(arena.data() + leaf_koff[s], pre, plen) != 0) {
                stop = true;  // CBTGHelper yields NULL and CBTGen stops (CritBitTree.h:76-79,:145-147)
                return;
            }
            harvest = true;
        }
        out.push_back(leaf_rec[s]);
        return;
    }
    if (!include_all && diff_at[p] >= plen) include_all = true;
    if (include_all) {
        iter_rec(child[0][p], pre, plen, true, harvest, stop, out);
        iter_rec(child[1][p], pre, plen, true, harvest, stop, out);
    } else {
        iter_rec(child[dir_of(p, pre, plen)][p], pre, plen, false, harvest, stop, out);
    }
}

void HostIndex::iter(const uint8_t *prefix, uint32_t plen, std::vector<uint32_t> &out) const {
    out.clear();
    if (!has_root) return;
    bool harvest = false, stop = false;
    // explicit stack would be safer for 65k-byte keys; tree depth is bounded by key bits
    iter_rec(root, prefix, plen, false, harvest, stop, out);
}

HostIndex::DeviceView HostIndex::device_view(cudaStream_t st) {
    if (dirty) {
        size_t ni = diff_at.size(), nl = leaf_rec.size();
        d_child0.reserve_discard(ni + 1);
        d_child1.reserve_discard(ni + 1);
        d_diff.reserve_discard(ni + 1);
        d_mask.reserve_discard(ni + 1);
        d_leaf_rec.reserve_discard(nl + 1);
        if (ni) {
            PX_CUDA(cudaMemcpyAsync(d_child0.p, child[0].data(), ni * sizeof(int32_t), cudaMemcpyHostToDevice, st));
            PX_CUDA(cudaMemcpyAsync(d_child1.p, child[1].data(), ni * sizeof(int32_t), cudaMemcpyHostToDevice, st));
            PX_CUDA(cudaMemcpyAsync(d_diff.p, diff_at.data(), ni * sizeof(uint16_t), cudaMemcpyHostToDevice, st));
            PX_CUDA(cudaMemcpyAsync(d_mask.p, mask.data(), ni * sizeof(uint8_t), cudaMemcpyHostToDevice, st));
        }
        if (nl) {
            d_leaf_klen.reserve_discard(nl + 1);
            d_leaf_koff.reserve_discard(nl + 1);
            PX_CUDA(cudaMemcpyAsync(d_leaf_rec.p, leaf_rec.data(), nl * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
            PX_CUDA(cudaMemcpyAsync(d_leaf_klen.p, leaf_klen.data(), nl * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
            PX_CUDA(cudaMemcpyAsync(d_leaf_koff.p, leaf_koff.data(), nl * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
        }
        // the key arena only ever grows: upload what is new
        if (arena.size() > keys_uploaded) {
            d_keys.reserve_keep(arena.size() + 16, keys_uploaded, st);
            PX_CUDA(cudaMemcpyAsync(d_keys.p + keys_uploaded, arena.data() + keys_uploaded, arena.size() - keys_uploaded,
                                    cudaMemcpyHostToDevice, st));
            keys_uploaded = arena.size();
        }
        PX_CUDA(cudaStreamSynchronize(st));
        dirty = false;
    }
    return DeviceView{d_child0.p, d_child1.p, d_diff.p, d_mask.p, d_leaf_rec.p, d_leaf_klen.p, d_leaf_koff.p, d_keys.p, root, has_root ? 1 : 0};
}

// ---------------------------------------------------------------------------------
// device: query escaping, walk, verify
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_query_len(uint32_t n, const uint8_t *__restrict__ keys, const int64_t *__restrict__ koff, uint32_t *__restrict__ qlen) {
    uint32_t i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    uint32_t c = 0;
    for (int64_t j = koff[i]; j < koff[i + 1]; j++) c += keys[j] == 251;
    qlen[i] = (uint32_t) (koff[i + 1] - koff[i]) + c + 2;
}

__global__ void __launch_bounds__(256)
k_query_write(uint32_t n, const uint8_t *__restrict__ keys, const int64_t *__restrict__ koff,
              const uint64_t *__restrict__ qoff, uint8_t *__restrict__ q) {
    uint32_t i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    uint8_t *o = q + qoff[i];
    for (int64_t j = koff[i]; j < koff[i + 1]; j++) {
        uint8_t b = keys[j];
        *o++ = b;
        if (b == 251) *o++ = 251;
    }
    o[0] = 251;
    o[1] = 0;
}

// one query per thread: level-by-level walk over the SoA nodes (find_best_match, CritBitTree.cpp:253-269), then
// verification of the candidate leaf against its escaped key in the index's key arena (key_eq, PiXiuStr.cpp:129-143:
// the key is the decoded prefix "esc(k) 251 0" of the record the leaf points to)
__global__ void __launch_bounds__(128)
k_lookup(uint32_t n, HostIndex::DeviceView T, const uint8_t *__restrict__ q, const uint64_t *__restrict__ qoff,
         const uint32_t *__restrict__ qlen, uint32_t *__restrict__ rec_out) {
    uint32_t i = blockIdx.x * 128 + threadIdx.x;
    if (i >= n) return;
    uint32_t res = 0xFFFFFFFFu;
    if (T.has_root) {
        const uint8_t *key = q + qoff[i];
        const uint32_t kl = qlen[i];
        int32_t p = T.root;
        while (p >= 0) {
            uint32_t da = T.diff_at[p];
            uint32_t b = da < kl ? key[da] : 0u;
            uint32_t dir = (1u + (T.mask[p] | b)) >> 8;
            p = dir ? T.child1[p] : T.child0[p];
        }
        const uint32_t s = (uint32_t) ~p;
        if (T.leaf_klen[s] == kl) {
            const uint8_t *lk = T.keys + T.leaf_koff[s];
            uint32_t j = 0;
            // 8 bytes per step once both sides are 8-byte aligned relative to each other; bytes otherwise
            if ((((uintptr_t) lk ^ (uintptr_t) key) & 7) == 0) {
                while (j < kl && (((uintptr_t) (key + j)) & 7) && lk[j] == key[j]) j++;
                if (j == kl || (((uintptr_t) (key + j)) & 7) == 0)
                    while (j + 8 <= kl && *reinterpret_cast<const uint64_t *>(lk + j) == *reinterpret_cast<const uint64_t *>(key + j)) j += 8;
            }
            while (j < kl && lk[j] == key[j]) j++;
            if (j == kl) res = T.leaf_rec[s];
        }
    }
    rec_out[i] = res;
}

void lookup_batch(Store &S, int64_t n, const uint8_t *h_keys, const int64_t *h_koff, std::vector<uint32_t> &rec_out) {
    rec_out.assign((size_t) n, 0xFFFFFFFFu);
    if (n == 0) return;
    cudaStream_t st = S.st;
    const uint32_t nn = (uint32_t) n;
    const int64_t kbytes = h_koff[n] - h_koff[0];
    S.in_keys.reserve_discard((size_t) kbytes + 16);
    S.in_koff.reserve_discard(nn + 1);
    std::vector<int64_t> rel(nn + 1);
    for (uint32_t i = 0; i <= nn; i++) rel[i] = h_koff[i] - h_koff[0];
    PX_CUDA(cudaMemcpyAsync(S.in_keys.p, h_keys + h_koff[0], (size_t) kbytes, cudaMemcpyHostToDevice, st));
    PX_CUDA(cudaMemcpyAsync(S.in_koff.p, rel.data(), (nn + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, st));
    HostIndex::DeviceView T = S.index->device_view(st);
    PX_CUDA(cudaEventRecord(S.ev0, st));
    S.prof.begin(PC_LOOKUP, st);
    S.doc_len.reserve_discard(nn + 1);
    DevBuf<uint64_t> &qoff = S.es.qoff;
    qoff.reserve_discard((size_t) nn + 2);
    uint64_t *d_qoff = qoff.p;
    k_query_len<<<div_up<uint32_t>(nn, 256), 256, 0, st>>>(nn, S.in_keys.p, S.in_koff.p, S.doc_len.p);
    {
        const uint32_t *ql = S.doc_len.p;
        device_scan<uint64_t>(
            (size_t) nn + 1, [=] __device__(size_t i) -> uint64_t { return i < nn ? (uint64_t) ql[i] : 0ull; },
            [=] __device__(size_t i, uint64_t v) { d_qoff[i] = v; }, OpSum(), 0ull, true, S.es.scanws, st);
    }
    uint64_t qbytes = 0;
    PX_CUDA(cudaMemcpyAsync(&qbytes, d_qoff + nn, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    PX_CUDA(cudaStreamSynchronize(st));
    S.in_vals.reserve_discard(qbytes + 16);
    k_query_write<<<div_up<uint32_t>(nn, 256), 256, 0, st>>>(nn, S.in_keys.p, S.in_koff.p, d_qoff, S.in_vals.p);
    S.doc_off.reserve_discard(nn + 1);
    k_lookup<<<div_up<uint32_t>(nn, 128), 128, 0, st>>>(nn, T, S.in_vals.p, d_qoff, S.doc_len.p, S.doc_off.p);
    PX_LAUNCH_CHECK();
    S.prof.end(st, 0.0, 4);
    S.launches += 4;
    PX_CUDA(cudaEventRecord(S.ev1, st));
    PX_CUDA(cudaMemcpyAsync(rec_out.data(), S.doc_off.p, nn * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    PX_CUDA(cudaStreamSynchronize(st));
    float ms = 0;
    PX_CUDA(cudaEventElapsedTime(&ms, S.ev0, S.ev1));
    S.last_lookup_ms = ms;
    S.prof.collect();
}

}  // namespace pixiu
