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
        if (nl) PX_CUDA(cudaMemcpyAsync(d_leaf_rec.p, leaf_rec.data(), nl * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        PX_CUDA(cudaStreamSynchronize(st));
        dirty = false;
    }
    return DeviceView{d_child0.p, d_child1.p, d_diff.p, d_mask.p, d_leaf_rec.p, root, has_root ? 1 : 0};
}

// ---------------------------------------------------------------------------------
// device: partial decode straight from the compressed store
// ---------------------------------------------------------------------------------
struct ChaseView {
    const uint8_t *enc;
    const uint64_t *enc_off;
    const uint32_t *enc_len, *dec_len, *first, *tile_base, *tile_desc;
};

// byte reader over an encoded record that fetches aligned 16-byte blocks (one global load per 16 encoded
// bytes instead of one per byte: the token walk is a chain of dependent loads)
struct EncReader {
    const uint8_t *base;
    uint64_t blk_addr = ~0ull;
    uint64_t lo = 0, hi = 0;
    __device__ explicit EncReader(const uint8_t *b) : base(b) {}
    __device__ __forceinline__ uint32_t get(uint32_t e) {
        uint64_t a = (uint64_t) (uintptr_t) (base + e), ba = a & ~15ull;
        if (ba != blk_addr) {
            uint4 v = __ldg(reinterpret_cast<const uint4 *>(ba));
            lo = (uint64_t) v.x | ((uint64_t) v.y << 32);
            hi = (uint64_t) v.z | ((uint64_t) v.w << 32);
            blk_addr = ba;
        }
        uint32_t o = (uint32_t) (a & 15);
        return (uint32_t) (((o < 8 ? lo : hi) >> (8 * (o & 7))) & 0xff);
    }
};

// decoded byte `o` of record g, following back references to a literal (iterative, no stack)
__device__ int resolve_byte(const ChaseView &V, uint32_t g, uint32_t o) {
    for (int hops = 0; hops < (1 << 20); hops++) {
        if (o >= V.dec_len[g]) return -1;
        uint32_t t = o / TILE;
        uint32_t desc = V.tile_desc[V.tile_base[g] + t];
        uint32_t e = desc & 0xffff, skip = desc >> 16;
        EncReader enc(V.enc + V.enc_off[g]);
        const uint32_t el = V.enc_len[g];
        bool raw = skip == 0xFFFF;
        uint32_t cur = t * TILE - (raw ? 0u : skip);
        bool jumped = false;
        while (e < el) {
            uint32_t b = enc.get(e);
            if (b != 251 || raw) {
                if (cur == o) return b;
                cur++;
                e++;
                raw = false;
                continue;
            }
            uint32_t nx = enc.get(e + 1);
            if (nx == 0 || nx == 251 || nx == 2) {
                if (cur == o) return 251;
                if (cur + 1 == o) return nx;
                cur += 2;
                e += 2;
                continue;
            }
            uint32_t idx = enc.get(e + 2) | (enc.get(e + 3) << 8), to = enc.get(e + 4) | (enc.get(e + 5) << 8), from, adv;
            if (nx == 1) {
                from = enc.get(e + 6) | (enc.get(e + 7) << 8);
                adv = 8;
            } else {
                from = to - nx;
                adv = 6;
            }
            uint32_t tl = to - from;
            if (o < cur + tl) {
                uint32_t k = o - cur, sg = V.first[g] + idx;
                if (sg == g) {
                    uint32_t period = cur - from;
                    o = from + (k % period);
                } else {
                    o = from + k;
                    g = sg;
                }
                jumped = true;
                break;
            }
            cur += tl;
            e += adv;
        }
        if (!jumped) return -1;
    }
    return -1;
}

// decoded[g0][0..qlen) == q ?   Ranges are followed through back references with a small stack;
// overlapping self references and stack overflow fall back to resolve_byte.
__device__ bool compare_prefix(const ChaseView &V, uint32_t g0, const uint8_t *q, uint32_t qlen) {
    if (V.dec_len[g0] < qlen) return false;
    struct Item {
        uint32_t g;
        uint16_t a, b, qo;
    };
    constexpr int STACK = 24;
    Item stack[STACK];
    int sp = 0;
    stack[sp++] = Item{g0, 0, (uint16_t) qlen, 0};
    while (sp > 0) {
        Item it = stack[--sp];
        uint32_t t = it.a / TILE;
        uint32_t desc = V.tile_desc[V.tile_base[it.g] + t];
        uint32_t e = desc & 0xffff, skip = desc >> 16;
        EncReader enc(V.enc + V.enc_off[it.g]);
        const uint32_t el = V.enc_len[it.g];
        bool raw = skip == 0xFFFF;
        uint32_t cur = t * TILE - (raw ? 0u : skip);
        uint32_t pos = it.a;
        while (pos < it.b) {
            if (e >= el) return false;
            uint32_t b = enc.get(e);
            if (b != 251 || raw) {
                if (cur == pos) {
                    if (b != q[it.qo + (pos - it.a)]) return false;
                    pos++;
                }
                cur++;
                e++;
                raw = false;
                continue;
            }
            uint32_t nx = enc.get(e + 1);
            if (nx == 0 || nx == 251 || nx == 2) {
                if (cur == pos) {
                    if (q[it.qo + (pos - it.a)] != 251) return false;
                    pos++;
                }
                cur++;
                if (pos < it.b && cur == pos) {
                    if (q[it.qo + (pos - it.a)] != nx) return false;
                    pos++;
                }
                cur++;
                e += 2;
                continue;
            }
            uint32_t idx = enc.get(e + 2) | (enc.get(e + 3) << 8), to = enc.get(e + 4) | (enc.get(e + 5) << 8), from, adv;
            if (nx == 1) {
                from = enc.get(e + 6) | (enc.get(e + 7) << 8);
                adv = 8;
            } else {
                from = to - nx;
                adv = 6;
            }
            uint32_t tl = to - from;
            if (cur + tl > pos) {
                uint32_t hi = min((uint32_t) it.b, cur + tl);
                uint32_t k0 = pos - cur, cnt = hi - pos;
                uint32_t sg = V.first[it.g] + idx;
                bool overlap = sg == it.g && from + k0 + cnt > cur;
                if (overlap || sp >= STACK) {
                    uint32_t period = cur - from;
                    for (uint32_t x = 0; x < cnt; x++) {
                        uint32_t so = sg == it.g ? from + ((k0 + x) % period) : from + k0 + x;
                        if (resolve_byte(V, sg, so) != (int) q[it.qo + (pos - it.a) + x]) return false;
                    }
                } else {
                    stack[sp++] = Item{sg, (uint16_t) (from + k0), (uint16_t) (from + k0 + cnt),
                                       (uint16_t) (it.qo + (pos - it.a))};
                }
                pos = hi;
            }
            cur += tl;
            e += adv;
        }
    }
    return true;
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

// one query per thread: level-by-level walk over the SoA nodes, then verification
__global__ void __launch_bounds__(128)
k_lookup(uint32_t n, HostIndex::DeviceView T, ChaseView V, const uint8_t *__restrict__ q,
         const uint64_t *__restrict__ qoff, const uint32_t *__restrict__ qlen, uint32_t *__restrict__ rec_out) {
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
        uint32_t g = T.leaf_rec[~p];
        if (compare_prefix(V, g, key, kl)) res = g;
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
    ChaseView V{S.d_enc.ptr(), S.d_enc_off.p, S.d_enc_len.p, S.d_dec_len.p, S.d_first.p, S.d_tile_base.p, S.d_tile_desc.p};
    k_lookup<<<div_up<uint32_t>(nn, 128), 128, 0, st>>>(nn, T, V, S.in_vals.p, d_qoff, S.doc_len.p, S.doc_off.p);
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
