// CritBit index: host maintenance + GPU batched lookup (see index.h).
#include "index.h"

#include <algorithm>
#include <chrono>
#include <cstring>

#include "scan.cuh"
#include "store.h"
#include "mintree.cuh"

namespace pixiu {

void escape_key(const uint8_t *k, size_t n, std::vector<uint8_t> &out, bool terminator) {
    if (!memchr(k, 251, n)) {  // nothing to double (every ASCII key)
        out.resize(n + (terminator ? 2 : 0));
        memcpy(out.data(), k, n);
    } else {
        out.clear();
        out.reserve(2 * n + 2);
        for (size_t i = 0; i < n; i++) {
            out.push_back(k[i]);
            if (k[i] == 251) out.push_back(251);
        }
        if (terminator) out.resize(out.size() + 2);
    }
    if (terminator) {
        out[out.size() - 2] = 251;
        out[out.size() - 1] = 0;
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
        dirty = true;  // a mirrored slot changes wholesale
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
        dirty = true;  // a mirrored slot changes wholesale
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

int32_t HostIndex::depth(const uint8_t *q, uint32_t qlen) const {
    if (!has_root) return 0;
    int32_t p = root, d = 0;
    while (p >= 0) {
        p = child[dir_of(p, q, qlen)][p];
        d++;
    }
    return d;
}

int64_t HostIndex::set(const uint8_t *q, uint32_t qlen, uint32_t rec) { return set_below(-1, 0, q, qlen, rec); }

// CritBitTree::setitem restricted to the subtree hanging on the edge (top, tdir) (top < 0: the whole tree).  The
// caller guarantees that the key's path from the root passes through that edge and that the new node belongs on
// it or below it.
int64_t HostIndex::set_below(int32_t top, int tdir, const uint8_t *q, uint32_t qlen, uint32_t rec) {
    if (!has_root) {
        root = ~new_leaf(q, qlen, rec);
        has_root = true;
        n_live = 1;
        return -1;
    }
    const int32_t start = top < 0 ? root : child[tdir][top];
    int32_t p = start;
    while (p >= 0) p = child[dir_of(p, q, qlen)][p];
    int32_t s = ~p;
    const uint8_t *lk = arena.data() + leaf_koff[s];
    uint32_t ll = leaf_klen[s], m = std::min(ll, qlen), diff = 0;
    while (diff < m && lk[diff] == q[diff]) diff++;
    if (diff == qlen && diff == ll) {  // same key: repoint the leaf (replace(), CritBitTree.cpp:32-43)
        int64_t old = leaf_rec[s];
        leaf_rec[s] = rec;
        note_leaf(s);
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
    int32_t parent = top, pdir = tdir, cur = start;
    while (cur >= 0) {
        if (diff_at[cur] > diff || (diff_at[cur] == diff && mask[cur] > mk)) break;
        pdir = dir_of(cur, q, qlen);
        parent = cur;
        cur = child[pdir][cur];
    }
    Probe pr{s, parent, cur, diff | ((uint32_t) mk << 16) | ((uint32_t) dir << 24) | ((uint32_t) pdir << 25)};
    return splice(pr, q, qlen, rec);
}

// put the new leaf and its inner node on the edge (pr.parent, pdir) that currently holds pr.cur
int64_t HostIndex::splice(const Probe &pr, const uint8_t *q, uint32_t qlen, uint32_t rec) {
    const int dir = (pr.info >> 24) & 1, pdir = (pr.info >> 25) & 1;
    int32_t nl = new_leaf(q, qlen, rec);
    int32_t ni = new_inner();
    diff_at[ni] = (uint16_t) (pr.info & 0xFFFF);
    mask[ni] = (uint8_t) (pr.info >> 16);
    child[dir][ni] = ~nl;
    child[1 - dir][ni] = pr.cur;
    if (pr.parent < 0) {
        root = ni;
    } else {
        child[pdir][pr.parent] = ni;
        note_child(pr.parent, pdir);
    }
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

// In-order walk below the prefix (CBTGen / CBTGHelper, CritBitTree.h:55-157) with an explicit stack: the depth of a
// CritBit tree is bounded only by the key bits (keys a, aa, aaa, ... chain one node per key), so recursion could
// overflow the host stack on ~100k chained keys.
void HostIndex::iter(const uint8_t *prefix, uint32_t plen, std::vector<uint32_t> &out) const {
    out.clear();
    if (!has_root) return;
    bool harvest = false;
    std::vector<std::pair<int32_t, bool>> stack;  // (node or ~leaf, whole subtree lies below the prefix)
    stack.emplace_back(root, false);
    while (!stack.empty()) {
        const int32_t p = stack.back().first;
        bool include_all = stack.back().second;
        stack.pop_back();
        if (p < 0) {
            const int32_t s = ~p;
            if (!harvest) {
                // CBTGHelper yields NULL and CBTGen stops (CritBitTree.h:76-79,:145-147)
                if (leaf_klen[s] < plen || memcmp(arena.data() + leaf_koff[s], prefix, plen) != 0) return;
                harvest = true;
            }
            out.push_back(leaf_rec[s]);
            continue;
        }
        if (!include_all && diff_at[p] >= plen) include_all = true;
        if (include_all) {
            stack.emplace_back(child[1][p], true);  // (popped after the whole left subtree)
            stack.emplace_back(child[0][p], true);
        } else {
            stack.emplace_back(child[dir_of(p, prefix, plen)][p], false);
        }
    }
}

// ---- wire format: "PXCB" u32 version | u64 n_inner, n_leaf, arena_bytes | i32 root, has_root | arrays ----
namespace {
template <typename T>
void put_vec(std::vector<uint8_t> &out, const std::vector<T> &v) {
    const uint8_t *p = reinterpret_cast<const uint8_t *>(v.data());
    out.insert(out.end(), p, p + v.size() * sizeof(T));
}
template <typename T>
bool get_vec(const uint8_t *&p, const uint8_t *end, size_t n, std::vector<T> &v) {
    if ((size_t) (end - p) < n * sizeof(T)) return false;
    v.resize(n);
    if (n) memcpy(v.data(), p, n * sizeof(T));
    p += n * sizeof(T);
    return true;
}
}  // namespace

void HostIndex::save(std::vector<uint8_t> &out) const {
    // free slots are compacted away: the saved tree only holds what is reachable
    std::vector<int32_t> inner_map(diff_at.size(), -1), leaf_map(leaf_rec.size(), -1);
    std::vector<int32_t> order_inner, order_leaf, stack;
    if (has_root) stack.push_back(root);
    while (!stack.empty()) {
        const int32_t p = stack.back();
        stack.pop_back();
        if (p < 0) {
            leaf_map[~p] = (int32_t) order_leaf.size();
            order_leaf.push_back(~p);
        } else {
            inner_map[p] = (int32_t) order_inner.size();
            order_inner.push_back(p);
            stack.push_back(child[1][p]);
            stack.push_back(child[0][p]);
        }
    }
    auto remap = [&](int32_t c) { return c < 0 ? ~leaf_map[~c] : inner_map[c]; };
    const uint64_t ni = order_inner.size(), nl = order_leaf.size();
    std::vector<int32_t> c0(ni), c1(ni);
    std::vector<uint16_t> da(ni);
    std::vector<uint8_t> mk(ni), ar;
    std::vector<uint32_t> lr(nl), lk(nl);
    std::vector<uint64_t> lo(nl);
    for (uint64_t i = 0; i < ni; i++) {
        const int32_t p = order_inner[i];
        c0[i] = remap(child[0][p]);
        c1[i] = remap(child[1][p]);
        da[i] = diff_at[p];
        mk[i] = mask[p];
    }
    for (uint64_t i = 0; i < nl; i++) {   // (DFS order = key order: the saved arena is sorted by key)
        const int32_t s = order_leaf[i];
        lr[i] = leaf_rec[s];
        lk[i] = leaf_klen[s];
        lo[i] = ar.size();
        ar.insert(ar.end(), arena.begin() + (ptrdiff_t) leaf_koff[s], arena.begin() + (ptrdiff_t) (leaf_koff[s] + leaf_klen[s]));
    }
    const uint64_t ab = ar.size();
    const uint32_t magic = 0x42435850u, version = 1;
    const int32_t r = has_root ? remap(root) : 0, hr = has_root ? 1 : 0;
    out.clear();
    auto put = [&](const void *p, size_t n) { out.insert(out.end(), (const uint8_t *) p, (const uint8_t *) p + n); };
    put(&magic, 4);
    put(&version, 4);
    put(&ni, 8);
    put(&nl, 8);
    put(&ab, 8);
    put(&r, 4);
    put(&hr, 4);
    put_vec(out, c0);
    put_vec(out, c1);
    put_vec(out, da);
    put_vec(out, mk);
    put_vec(out, lr);
    put_vec(out, lk);
    put_vec(out, lo);
    put_vec(out, ar);
}

bool HostIndex::load(const uint8_t *blob, size_t size, uint32_t n_records, std::vector<uint32_t> &live_out) {
    const uint8_t *p = blob, *end = blob + size;
    if (size < 40) return false;
    uint32_t magic, version;
    uint64_t ni, nl, ab;
    int32_t r, hr;
    memcpy(&magic, p, 4);
    memcpy(&version, p + 4, 4);
    memcpy(&ni, p + 8, 8);
    memcpy(&nl, p + 16, 8);
    memcpy(&ab, p + 24, 8);
    memcpy(&r, p + 32, 4);
    memcpy(&hr, p + 36, 4);
    p += 40;
    if (magic != 0x42435850u || version != 1 || ni > 0x7fffffffull || nl > 0x7fffffffull || (nl && ni + 1 != nl) || (hr != 0) != (nl != 0))
        return false;
    std::vector<int32_t> c0, c1;
    std::vector<uint16_t> da;
    std::vector<uint8_t> mk, ar;
    std::vector<uint32_t> lr, lk;
    std::vector<uint64_t> lo;
    if (!get_vec(p, end, ni, c0) || !get_vec(p, end, ni, c1) || !get_vec(p, end, ni, da) || !get_vec(p, end, ni, mk) ||
        !get_vec(p, end, nl, lr) || !get_vec(p, end, nl, lk) || !get_vec(p, end, nl, lo) || !get_vec(p, end, ab, ar) || p != end)
        return false;
    auto child_ok = [&](int32_t c) { return c < 0 ? (uint64_t) ~c < nl : (uint64_t) c < ni; };
    for (uint64_t i = 0; i < ni; i++)
        if (!child_ok(c0[i]) || !child_ok(c1[i])) return false;
    for (uint64_t i = 0; i < nl; i++)
        if (lr[i] >= n_records || lo[i] + lk[i] > ab || lk[i] < 2) return false;
    if (hr && !child_ok(r)) return false;
    child[0].swap(c0);
    child[1].swap(c1);
    diff_at.swap(da);
    mask.swap(mk);
    leaf_rec.swap(lr);
    leaf_klen.swap(lk);
    leaf_koff.swap(lo);
    arena.swap(ar);
    free_inner.clear();
    free_leaf.clear();
    root = r;
    has_root = hr != 0;
    n_live = nl;
    dirty = true;
    synced_inner = synced_leaf = 0;
    keys_uploaded = 0;
    mod_child.clear();
    mod_leaf.clear();
    live_out = leaf_rec;
    return true;
}

size_t HostIndex::host_bytes() const {
    return child[0].capacity() * 8 + diff_at.capacity() * 2 + mask.capacity() + leaf_rec.capacity() * 4 + leaf_klen.capacity() * 4 +
           leaf_koff.capacity() * 8 + arena.capacity();
}
size_t HostIndex::device_bytes() const {
    return d_child0.cap * 4 + d_child1.cap * 4 + d_diff.cap * 2 + d_mask.cap + d_leaf_rec.cap * 4 + d_leaf_klen.cap * 4 +
           d_leaf_koff.cap * 8 + d_keys.cap + d_nodes.cap * 16 + d_leaves.cap * 16;
}

// packed copy of the inner nodes for the walks: {child0, child1, diff_at | mask << 16, 0} - one 16-byte load (one
// sector) per level instead of three loads from three arrays
__global__ void __launch_bounds__(256)
k_pack_nodes(uint32_t a, uint32_t b, const int32_t *__restrict__ child0, const int32_t *__restrict__ child1,
             const uint16_t *__restrict__ diff_at, const uint8_t *__restrict__ mask, int4 *__restrict__ nodes) {
    uint32_t i = a + blockIdx.x * 256 + threadIdx.x;
    if (i < b) nodes[i] = make_int4(child0[i], child1[i], (int32_t) ((uint32_t) diff_at[i] | ((uint32_t) mask[i] << 16)), 0);
}

// packed leaves {key offset (64 bit), key length, record id}
__global__ void __launch_bounds__(256)
k_pack_leaves(uint32_t a, uint32_t b, const uint64_t *__restrict__ koff, const uint32_t *__restrict__ klen,
              const uint32_t *__restrict__ rec, uint4 *__restrict__ leaves) {
    uint32_t i = a + blockIdx.x * 256 + threadIdx.x;
    if (i < b) leaves[i] = make_uint4((uint32_t) koff[i], (uint32_t) (koff[i] >> 32), klen[i], rec[i]);
}

__global__ void __launch_bounds__(256)
k_apply_index_mods(uint32_t n_child, uint32_t n_leaf, const int32_t *__restrict__ mods, int32_t *__restrict__ child0,
                   int32_t *__restrict__ child1, uint32_t *__restrict__ leaf_rec, int4 *__restrict__ nodes,
                   uint4 *__restrict__ leaves) {
    // mods: n_child pairs (node * 2 + dir, value), then n_leaf pairs (slot, record id)
    uint32_t i = blockIdx.x * 256 + threadIdx.x;
    if (i < n_child) {
        int32_t k = mods[2 * i], v = mods[2 * i + 1];
        ((k & 1) ? child1 : child0)[k >> 1] = v;
        (reinterpret_cast<int32_t *>(nodes + (k >> 1)))[k & 1] = v;
    } else if (i < n_child + n_leaf) {
        leaf_rec[mods[2 * i]] = (uint32_t) mods[2 * i + 1];
        leaves[mods[2 * i]].w = (uint32_t) mods[2 * i + 1];
    }
}

HostIndex::DeviceView HostIndex::device_view(cudaStream_t st) {
    const size_t ni = diff_at.size(), nl = leaf_rec.size();
    if (dirty) {
        synced_inner = synced_leaf = 0;
        mod_child.clear();
        mod_leaf.clear();
    }
    if (ni > synced_inner) {  // appended inner nodes
        const size_t a = synced_inner, c = ni - a;
        d_child0.reserve_keep(ni + 1, a, st);
        d_child1.reserve_keep(ni + 1, a, st);
        d_diff.reserve_keep(ni + 1, a, st);
        d_mask.reserve_keep(ni + 1, a, st);
        PX_CUDA(cudaMemcpyAsync(d_child0.p + a, child[0].data() + a, c * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        PX_CUDA(cudaMemcpyAsync(d_child1.p + a, child[1].data() + a, c * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        PX_CUDA(cudaMemcpyAsync(d_diff.p + a, diff_at.data() + a, c * sizeof(uint16_t), cudaMemcpyHostToDevice, st));
        PX_CUDA(cudaMemcpyAsync(d_mask.p + a, mask.data() + a, c * sizeof(uint8_t), cudaMemcpyHostToDevice, st));
        d_nodes.reserve_keep(ni + 1, a, st);
        k_pack_nodes<<<(unsigned) div_up<size_t>(c, 256), 256, 0, st>>>((uint32_t) a, (uint32_t) ni, d_child0.p, d_child1.p, d_diff.p,
                                                                       d_mask.p, d_nodes.p);
    }
    if (nl > synced_leaf) {  // appended leaves
        const size_t a = synced_leaf, c = nl - a;
        d_leaf_rec.reserve_keep(nl + 1, a, st);
        d_leaf_klen.reserve_keep(nl + 1, a, st);
        d_leaf_koff.reserve_keep(nl + 1, a, st);
        PX_CUDA(cudaMemcpyAsync(d_leaf_rec.p + a, leaf_rec.data() + a, c * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        PX_CUDA(cudaMemcpyAsync(d_leaf_klen.p + a, leaf_klen.data() + a, c * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        PX_CUDA(cudaMemcpyAsync(d_leaf_koff.p + a, leaf_koff.data() + a, c * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
        d_leaves.reserve_keep(nl + 1, a, st);
        k_pack_leaves<<<(unsigned) div_up<size_t>(c, 256), 256, 0, st>>>((uint32_t) a, (uint32_t) nl, d_leaf_koff.p, d_leaf_klen.p,
                                                                        d_leaf_rec.p, d_leaves.p);
    }
    // the key arena only ever grows: upload what is new
    if (arena.size() > keys_uploaded) {
        d_keys.reserve_keep(arena.size() + 16, keys_uploaded, st);
        PX_CUDA(cudaMemcpyAsync(d_keys.p + keys_uploaded, arena.data() + keys_uploaded, arena.size() - keys_uploaded,
                                cudaMemcpyHostToDevice, st));
        keys_uploaded = arena.size();
    }
    // scattered changes of entries that were already mirrored
    const size_t nc = mod_child.size(), nm = mod_leaf.size();
    if (nc + nm) {
        std::vector<int32_t> pairs(2 * (nc + nm));
        for (size_t i = 0; i < nc; i++) {
            pairs[2 * i] = mod_child[i];
            pairs[2 * i + 1] = child[mod_child[i] & 1][mod_child[i] >> 1];
        }
        for (size_t i = 0; i < nm; i++) {
            pairs[2 * (nc + i)] = mod_leaf[i];
            pairs[2 * (nc + i) + 1] = (int32_t) leaf_rec[mod_leaf[i]];
        }
        d_mod.reserve_discard(pairs.size());
        PX_CUDA(cudaMemcpyAsync(d_mod.p, pairs.data(), pairs.size() * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        k_apply_index_mods<<<(unsigned) div_up<size_t>(nc + nm, 256), 256, 0, st>>>((uint32_t) nc, (uint32_t) nm, d_mod.p, d_child0.p,
                                                                                   d_child1.p, d_leaf_rec.p, d_nodes.p, d_leaves.p);
        PX_CUDA(cudaStreamSynchronize(st));  // pairs is a local
        mod_child.clear();
        mod_leaf.clear();
    }
    PX_CUDA(cudaStreamSynchronize(st));
    synced_inner = ni;
    synced_leaf = nl;
    dirty = false;
    return DeviceView{d_child0.p, d_child1.p, d_diff.p, d_mask.p, d_leaf_rec.p, d_leaf_klen.p, d_leaf_koff.p, d_keys.p, d_leaves.p, d_nodes.p, root, has_root ? 1 : 0};
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
            const int4 nd = __ldg(T.nodes + p);  // child0, child1, diff_at | mask << 16
            const uint32_t da = (uint32_t) nd.z & 0xFFFFu;
            const uint32_t b = da < kl ? key[da] : 0u;
            const uint32_t dir = (1u + ((((uint32_t) nd.z >> 16) & 0xFFu) | b)) >> 8;
            p = dir ? nd.y : nd.x;
        }
        const uint4 lf = __ldg(T.leaves + (uint32_t) ~p);  // key offset lo/hi, key length, record id
        if (lf.z == kl) {
            const uint8_t *lk = T.keys + (((uint64_t) lf.y << 32) | lf.x);
            uint32_t j = 0;
            // 8 bytes per step once both sides are 8-byte aligned relative to each other; bytes otherwise
            if ((((uintptr_t) lk ^ (uintptr_t) key) & 7) == 0) {
                while (j < kl && (((uintptr_t) (key + j)) & 7) && lk[j] == key[j]) j++;
                if (j == kl || (((uintptr_t) (key + j)) & 7) == 0)
                    while (j + 8 <= kl && *reinterpret_cast<const uint64_t *>(lk + j) == *reinterpret_cast<const uint64_t *>(key + j)) j += 8;
            }
            while (j < kl && lk[j] == key[j]) j++;
            if (j == kl) res = lf.w;
        }
    }
    rec_out[i] = res;
}

// one key per thread: the two read-only walks of CritBitTree::setitem (best-match leaf and critical position,
// CritBitTree.cpp:13-30,:45-66; then the edge where the new node belongs, :67-81)
__global__ void __launch_bounds__(128)
k_insert_probe(uint32_t n, HostIndex::DeviceView T, const uint8_t *__restrict__ q, const uint64_t *__restrict__ qoff,
               const uint32_t *__restrict__ qlen, HostIndex::Probe *__restrict__ out) {
    uint32_t i = blockIdx.x * 128 + threadIdx.x;
    if (i >= n) return;
    const uint8_t *key = q + qoff[i];
    const uint32_t kl = qlen[i];
    int32_t p = T.root;
    while (p >= 0) {
        uint32_t da = T.diff_at[p];
        uint32_t b = da < kl ? key[da] : 0u;
        p = ((1u + (T.mask[p] | b)) >> 8) ? T.child1[p] : T.child0[p];
    }
    const uint32_t s = (uint32_t) ~p;
    const uint8_t *lk = T.keys + T.leaf_koff[s];
    const uint32_t ll = T.leaf_klen[s], m = min(ll, kl);
    uint32_t diff = 0;
    while (diff < m && lk[diff] == key[diff]) diff++;
    HostIndex::Probe r;
    r.leaf = (int32_t) s;
    if (diff == kl && diff == ll) {
        r.parent = -1;
        r.cur = 0;
        r.info = 1u << 26;
        out[i] = r;
        return;
    }
    const uint32_t a = diff < ll ? lk[diff] : 0u, b = diff < kl ? key[diff] : 0u;
    uint32_t mk = a ^ b;
    mk |= mk >> 1;
    mk |= mk >> 2;
    mk |= mk >> 4;
    mk = ((mk & ~(mk >> 1)) ^ 0xFFu) & 0xFFu;
    const uint32_t dir = (1u + (mk | b)) >> 8;
    int32_t parent = -1, cur = T.root;
    uint32_t pdir = 0;
    while (cur >= 0) {
        const uint32_t da = T.diff_at[cur], cm = T.mask[cur];
        if (da > diff || (da == diff && cm > mk)) break;
        const uint32_t kb = da < kl ? key[da] : 0u;
        pdir = (1u + (cm | kb)) >> 8;
        parent = cur;
        cur = pdir ? T.child1[cur] : T.child0[cur];
    }
    r.parent = parent;
    r.cur = cur;
    r.info = diff | (mk << 16) | (dir << 24) | (pdir << 25);
    out[i] = r;
}

// ---------------------------------------------------------------------------------
// Bulk build of the tree on the GPU: when a batch is at least as large as the tree, the tree is not
// spliced key by key on the host (CritBitTree::setitem, CritBitTree.cpp:13-105: 0.35 us per key, serial) but rebuilt
// from ALL keys - the live leaves and the batch - on the device:
//   1. LSD radix sort of the keys by their bytes, 8 bytes per round from the last chunk to the first (the existing
//      onesweep sort; stable, so equal keys keep their order: tree first, then the batch in insertion order);
//   2. equal neighbours: the last of a group survives, every batch element whose key was there before it replaces its
//      predecessor (in-order semantics of n setitem calls: rc = CBT_SET_REPLACE and the old record is tombstoned);
//   3. critical position of every pair of neighbouring distinct keys (first differing byte, highest differing bit);
//   4. the tree over the sorted leaves is the Cartesian tree of those positions (the pair with the smallest position
//      splits a range: it is unique there): nearest smaller position to the left and to the right of every pair
//      (block-min tree search), the parent is the nearer-in-value of the two, children are set by their parents' slots;
//   5. the flat arrays come back to the host mirror in one copy; the leaves are in key order.
// ---------------------------------------------------------------------------------
constexpr uint32_t BULK_MAX_KEY = 248;   // escaped key bytes the radix rounds cover (longer keys: the host path)

// 8 key bytes [8c, 8c + 8) of entry perm[j], big endian, zero padded: the radix key of round c
__global__ void __launch_bounds__(256)
k_bulk_chunk(uint32_t n, const uint32_t *__restrict__ perm, const uint64_t *__restrict__ eoff, const uint32_t *__restrict__ elen,
             const uint8_t *__restrict__ keys, uint32_t c, uint64_t *__restrict__ out) {
    const uint32_t j = blockIdx.x * 256 + threadIdx.x;
    if (j >= n) return;
    const uint32_t e = perm ? perm[j] : j;
    const uint8_t *k = keys + eoff[e];
    const uint32_t len = elen[e];
    uint64_t v = 0;
#pragma unroll
    for (uint32_t b = 0; b < 8; b++) {
        const uint32_t p = 8 * c + b;
        v = (v << 8) | (p < len ? (uint64_t) k[p] : 0ull);
    }
    out[j] = v;
}

// neighbours in sorted order: dup[j] = same key as j - 1; crit[j] = (first differing byte << 8) | mask otherwise
__global__ void __launch_bounds__(256)
k_bulk_adjacent(uint32_t n, const uint32_t *__restrict__ perm, const uint64_t *__restrict__ eoff, const uint32_t *__restrict__ elen,
                const uint8_t *__restrict__ keys, uint8_t *__restrict__ dup, uint32_t *__restrict__ crit) {
    const uint32_t j = blockIdx.x * 256 + threadIdx.x;
    if (j >= n) return;
    if (j == 0) {
        dup[0] = 0;
        crit[0] = 0;
        return;
    }
    const uint32_t a = perm[j - 1], b = perm[j];
    const uint8_t *ka = keys + eoff[a], *kb = keys + eoff[b];
    const uint32_t la = elen[a], lb = elen[b], m = min(la, lb);
    uint32_t d = 0;
    while (d < m && ka[d] == kb[d]) d++;
    if (d == la && d == lb) {
        dup[j] = 1;
        crit[j] = 0;
        return;
    }
    // (keys end with 251,0 and are prefix free: d < m)
    uint32_t x = (uint32_t) (ka[d] ^ kb[d]);
    x |= x >> 1;
    x |= x >> 2;
    x |= x >> 4;
    const uint32_t mk = ((x & ~(x >> 1)) ^ 0xFFu) & 0xFFu;
    dup[j] = 0;
    crit[j] = (d << 8) | mk;
}

// survivors -> leaves (key order), replaced predecessors -> old_out, crit of unique pair (u - 1, u) -> cu[u - 1]
__global__ void __launch_bounds__(256)
k_bulk_leaves(uint32_t n, uint32_t n_old, const uint32_t *__restrict__ perm, const uint8_t *__restrict__ dup,
              const uint32_t *__restrict__ crit, const uint32_t *__restrict__ uidx /* exclusive count of group heads */,
              const uint64_t *__restrict__ eoff, const uint32_t *__restrict__ elen, const uint32_t *__restrict__ erec,
              uint64_t *__restrict__ leaf_koff, uint32_t *__restrict__ leaf_klen, uint32_t *__restrict__ leaf_rec,
              uint32_t *__restrict__ cu, long long *__restrict__ old_out) {
    const uint32_t j = blockIdx.x * 256 + threadIdx.x;
    if (j >= n) return;
    const uint32_t e = perm[j];
    const uint32_t u = uidx[j] + (dup[j] ? 0u : 1u) - 1u;   // index of this element's group among the unique keys
    if (dup[j] && e >= n_old) old_out[e - n_old] = (long long) erec[perm[j - 1]];   // it replaces its predecessor
    if (!dup[j] && u > 0) cu[u - 1] = crit[j];
    if (j + 1 == n || !dup[j + 1]) {   // the last of its group survives
        leaf_koff[u] = eoff[e];
        leaf_klen[u] = elen[e];
        leaf_rec[u] = erec[e];
    }
}

// inner node i sits between leaves i and i + 1; its parent is the nearer in value of the nearest smaller positions to
// its left and right; a side without inner nodes holds the leaf
__global__ void __launch_bounds__(256)
k_bulk_tree(uint32_t ni, MinTree T, const uint32_t *__restrict__ cu, int32_t *__restrict__ child0, int32_t *__restrict__ child1,
            uint16_t *__restrict__ diff_at, uint8_t *__restrict__ mask, int32_t *__restrict__ root) {
    const uint32_t i = blockIdx.x * 256 + threadIdx.x;
    if (i >= ni) return;
    const uint32_t c = cu[i];
    uint32_t acc = 0xFFFFFFFFu;
    const int64_t L = i > 0 ? tree_search<true, false, true>(T, i, c, acc, -1) : -1;
    const int64_t R = i + 1 < ni ? tree_search<false, false, true>(T, i, c, acc, -1) : (int64_t) ni;
    diff_at[i] = (uint16_t) (c >> 8);
    mask[i] = (uint8_t) (c & 0xFFu);
    if (L == (int64_t) i - 1) child0[i] = ~(int32_t) i;          // no inner node between L and i: leaf i
    if (R == (int64_t) i + 1) child1[i] = ~(int32_t) (i + 1);    // ... leaf i + 1
    if (L < 0 && R >= (int64_t) ni) {
        *root = (int32_t) i;
    } else {
        const int64_t p = L < 0 ? R : (R >= (int64_t) ni ? L : (cu[L] > cu[R] ? L : R));
        if ((int64_t) i < p) child0[p] = (int32_t) i;
        else child1[p] = (int32_t) i;
    }
}

bool HostIndex::bulk_build(Store &S, uint32_t n, const uint8_t *d_keys_raw, const int64_t *d_koff, uint32_t first_rec,
                           int64_t *old_out) {
    cudaStream_t st = S.st;
    auto tnow = [] { return std::chrono::steady_clock::now(); };
    auto t_prev = tnow();
    auto lap = [&](const char *what) {   // (PIXIU_TRACE: where the time of a bulk build goes)
        if (!S.knobs.trace) return;
        cudaStreamSynchronize(st);
        const auto t = tnow();
        fprintf(stderr, "[index bulk] %-28s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(t - t_prev).count());
        t_prev = t;
    };
    // live leaves of the tree (slot list) and their longest key
    std::vector<uint32_t> live_slots;
    live_slots.reserve(n_live);
    {
        std::vector<uint8_t> is_free(leaf_rec.size(), 0);
        for (int32_t f : free_leaf) is_free[(size_t) f] = 1;
        for (size_t sidx = 0; sidx < leaf_rec.size(); sidx++)
            if (!is_free[sidx]) {
                if (leaf_klen[sidx] > BULK_MAX_KEY) return false;
                live_slots.push_back((uint32_t) sidx);
            }
    }
    const uint32_t n_old = (uint32_t) live_slots.size();
    const uint64_t E64 = (uint64_t) n_old + n;
    if (E64 >= 0x7fffffffull) return false;
    const uint32_t E = (uint32_t) E64;
    // escaped batch keys on the device
    d_qlen.reserve_discard(n + 1);
    d_qoff.reserve_discard((size_t) n + 2);
    uint32_t *ql = d_qlen.p;
    uint64_t *qo = d_qoff.p;
    k_query_len<<<div_up<uint32_t>(n, 256), 256, 0, st>>>(n, d_keys_raw, d_koff, ql);
    device_scan<uint64_t>(
        (size_t) n + 1, [=] __device__(size_t i) -> uint64_t { return i < n ? (uint64_t) ql[i] : 0ull; },
        [=] __device__(size_t i, uint64_t v) { qo[i] = v; }, OpSum(), 0ull, true, S.es.scanws, st);
    // (longest batch key: a max-scan's last element)
    DevBuf<uint32_t> &tmp32 = S.iter_buf;
    tmp32.reserve_discard(4 * (size_t) E + 64);
    uint32_t *d_max = tmp32.p;
    device_scan<uint32_t>(
        (size_t) n, [=] __device__(size_t i) -> uint32_t { return ql[i]; },
        [=] __device__(size_t i, uint32_t v) {
            if (i + 1 == n) d_max[0] = v;
        },
        OpMax(), 0u, false, S.es.scanws, st);
    uint64_t qbytes = 0;
    uint32_t qmax = 0;
    PX_CUDA(cudaMemcpyAsync(&qbytes, qo + n, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    PX_CUDA(cudaMemcpyAsync(&qmax, d_max, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    PX_CUDA(cudaStreamSynchronize(st));
    if (qmax > BULK_MAX_KEY) return false;
    uint32_t lmax = qmax;
    for (uint32_t sidx : live_slots) lmax = std::max(lmax, leaf_klen[sidx]);
    // the batch keys join the key arena (device: appended behind the mirrored arena; host: one copy back)
    device_view(st);   // the mirror is current (arena uploaded)
    const uint64_t abase = arena.size();
    this->d_keys.reserve_keep(abase + qbytes + 16, keys_uploaded, st);
    k_query_write<<<div_up<uint32_t>(n, 256), 256, 0, st>>>(n, d_keys_raw, d_koff, qo, this->d_keys.p + abase);
    arena.resize(abase + qbytes);
    PX_CUDA(cudaMemcpyAsync(arena.data() + abase, this->d_keys.p + abase, qbytes, cudaMemcpyDeviceToHost, st));
    keys_uploaded = arena.size();
    // entries: live leaves first, then the batch in order
    DevBuf<uint64_t> &eoff = S.es.qoff;          // (free here: the lookups' query offsets)
    eoff.reserve_discard((size_t) E + 2);
    DevBuf<uint32_t> elen, erec;
    elen.reserve_discard(E + 1);
    erec.reserve_discard(E + 1);
    {
        std::vector<uint64_t> ho(n_old);
        std::vector<uint32_t> hl(n_old), hr(n_old);
        for (uint32_t k = 0; k < n_old; k++) {
            ho[k] = leaf_koff[live_slots[k]];
            hl[k] = leaf_klen[live_slots[k]];
            hr[k] = leaf_rec[live_slots[k]];
        }
        if (n_old) {
            PX_CUDA(cudaMemcpyAsync(eoff.p, ho.data(), n_old * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
            PX_CUDA(cudaMemcpyAsync(elen.p, hl.data(), n_old * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
            PX_CUDA(cudaMemcpyAsync(erec.p, hr.data(), n_old * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
            PX_CUDA(cudaStreamSynchronize(st));
        }
        uint64_t *eo = eoff.p;
        uint32_t *el = elen.p, *er = erec.p;
        device_scan<uint32_t>(
            (size_t) n, [=] __device__(size_t i) -> uint32_t { return 0u; },
            [=] __device__(size_t i, uint32_t) {
                eo[n_old + i] = abase + qo[i];
                el[n_old + i] = ql[i];
                er[n_old + i] = first_rec + (uint32_t) i;
            },
            OpSum(), 0u, true, S.es.scanws, st);
    }
    lap("entries + key arena");
    // ---- 1. LSD radix sort by 8-byte chunks ----
    EncodeScratch &X = S.es;
    X.keys0.reserve_discard(E);
    X.keys1.reserve_discard(E);
    X.vals0.reserve_discard(E);
    X.vals1.reserve_discard(E);
    if (!X.counters.p) {
        X.counters.reserve_discard(16);
        PX_CUDA(cudaMemsetAsync(X.counters.p, 0, 16 * sizeof(uint32_t), st));
    }
    const uint32_t nchunks = div_up<uint32_t>(lmax, 8);
    uint32_t *perm = nullptr;   // nullptr: identity
    int launches_ = 0;
    for (int c = (int) nchunks - 1; c >= 0; c--) {
        uint32_t *vin = perm ? perm : X.vals0.p;
        uint32_t *vother = vin == X.vals0.p ? X.vals1.p : X.vals0.p;
        k_bulk_chunk<<<div_up<uint32_t>(E, 256), 256, 0, st>>>(E, perm, eoff.p, elen.p, this->d_keys.p, (uint32_t) c, X.keys0.p);
        int cur;
        if (!perm) {
            cur = radix_sort_pairs<uint64_t>(X.keys0.p, X.keys1.p, X.vals0.p, X.vals1.p, E, 0, 64, true, X.rs, X.counters.p + 2, st, &launches_);
            perm = cur ? X.vals1.p : X.vals0.p;
        } else {
            // values: the permutation so far (vin), ping-ponged with the other buffer
            cur = radix_sort_pairs<uint64_t>(X.keys0.p, X.keys1.p, vin, vother, E, 0, 64, false, X.rs, X.counters.p + 2, st, &launches_);
            perm = cur ? vother : vin;
        }
        launches_++;
    }
    if (!perm) {   // (no key bytes at all cannot happen: keys end with 251,0)
        return false;
    }
    lap("radix sort");
    // ---- 2./3. neighbours ----
    DevBuf<uint8_t> dup;
    dup.reserve_discard(E + 1);
    uint32_t *crit = tmp32.p, *uidx = tmp32.p + E, *cu = tmp32.p + 2 * (size_t) E + 8;
    k_bulk_adjacent<<<div_up<uint32_t>(E, 256), 256, 0, st>>>(E, perm, eoff.p, elen.p, this->d_keys.p, dup.p, crit);
    {
        const uint8_t *dp = dup.p;
        uint32_t *ui = uidx;
        device_scan<uint32_t>(
            (size_t) E + 1, [=] __device__(size_t j) -> uint32_t { return j < E && !dp[j] ? 1u : 0u; },
            [=] __device__(size_t j, uint32_t v) {
                if (j < E) ui[j] = v;
                else ui[E] = v;   // (slot E: the number of unique keys)
            },
            OpSum(), 0u, true, S.es.scanws, st);
    }
    uint32_t m = 0;
    PX_CUDA(cudaMemcpyAsync(&m, uidx + E, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    DevBuf<long long> d_old;
    d_old.reserve_discard(n + 1);
    PX_CUDA(cudaMemsetAsync(d_old.p, 0xFF, (size_t) n * sizeof(long long), st));   // -1
    PX_CUDA(cudaStreamSynchronize(st));
    if (m == 0) return false;
    // ---- leaves in key order ----
    d_leaf_koff.reserve_discard(m + 1);
    d_leaf_klen.reserve_discard(m + 1);
    d_leaf_rec.reserve_discard(m + 1);
    k_bulk_leaves<<<div_up<uint32_t>(E, 256), 256, 0, st>>>(E, n_old, perm, dup.p, crit, uidx, eoff.p, elen.p, erec.p, d_leaf_koff.p,
                                                           d_leaf_klen.p, d_leaf_rec.p, cu, d_old.p);
    // ---- 4. Cartesian tree of the critical positions ----
    const uint32_t ni = m - 1;
    d_child0.reserve_discard(ni + 1);
    d_child1.reserve_discard(ni + 1);
    d_diff.reserve_discard(ni + 1);
    d_mask.reserve_discard(ni + 1);
    DevBuf<int32_t> d_root;
    d_root.reserve_discard(4);
    int32_t h_root = ~0;   // a single leaf
    if (ni) {
        MinTree T{};
        size_t total = 0;
        uint32_t sz = ni;
        int nlev = 1;
        while (sz > 1 && nlev < TREE_MAX_LEVELS) {
            sz = div_up<uint32_t>(sz, TREE_B);
            total += sz;
            nlev++;
        }
        X.tree_a.reserve_discard(total + 1);
        T.a[0] = cu;
        T.l[0] = cu;
        T.size[0] = ni;
        T.nlev = 1;
        sz = ni;
        size_t o = 0;
        while (sz > 1 && T.nlev < TREE_MAX_LEVELS) {
            const uint32_t so = div_up<uint32_t>(sz, TREE_B);
            k_tree_level<<<div_up<uint32_t>(so, 256), 256, 0, st>>>(T.a[T.nlev - 1], T.a[T.nlev - 1], sz, X.tree_a.p + o, X.tree_a.p + o, so);
            T.a[T.nlev] = X.tree_a.p + o;
            T.l[T.nlev] = X.tree_a.p + o;
            T.size[T.nlev] = so;
            T.nlev++;
            o += so;
            sz = so;
        }
        k_bulk_tree<<<div_up<uint32_t>(ni, 256), 256, 0, st>>>(ni, T, cu, d_child0.p, d_child1.p, d_diff.p, d_mask.p, d_root.p);
        PX_CUDA(cudaMemcpyAsync(&h_root, d_root.p, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    }
    PX_LAUNCH_CHECK();
    lap("neighbours + leaves + tree");
    // ---- 5. back to the host mirror ----
    child[0].resize(ni);
    child[1].resize(ni);
    diff_at.resize(ni);
    mask.resize(ni);
    leaf_rec.resize(m);
    leaf_klen.resize(m);
    leaf_koff.resize(m);
    static_assert(sizeof(long long) == sizeof(int64_t), "old_out copy");
    if (ni) {
        PX_CUDA(cudaMemcpyAsync(child[0].data(), d_child0.p, ni * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        PX_CUDA(cudaMemcpyAsync(child[1].data(), d_child1.p, ni * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        PX_CUDA(cudaMemcpyAsync(diff_at.data(), d_diff.p, ni * sizeof(uint16_t), cudaMemcpyDeviceToHost, st));
        PX_CUDA(cudaMemcpyAsync(mask.data(), d_mask.p, ni * sizeof(uint8_t), cudaMemcpyDeviceToHost, st));
    }
    PX_CUDA(cudaMemcpyAsync(leaf_rec.data(), d_leaf_rec.p, m * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    PX_CUDA(cudaMemcpyAsync(leaf_klen.data(), d_leaf_klen.p, m * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    PX_CUDA(cudaMemcpyAsync(leaf_koff.data(), d_leaf_koff.p, m * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    PX_CUDA(cudaMemcpyAsync(old_out, d_old.p, (size_t) n * sizeof(long long), cudaMemcpyDeviceToHost, st));
    PX_CUDA(cudaStreamSynchronize(st));
    free_inner.clear();
    free_leaf.clear();
    root = h_root;
    has_root = true;
    n_live = m;
    // the device arrays ARE the mirror now; the packed walk copies follow
    d_nodes.reserve_discard(ni + 1);
    d_leaves.reserve_discard(m + 1);
    if (ni) k_pack_nodes<<<div_up<uint32_t>(ni, 256), 256, 0, st>>>(0u, ni, d_child0.p, d_child1.p, d_diff.p, d_mask.p, d_nodes.p);
    k_pack_leaves<<<div_up<uint32_t>(m, 256), 256, 0, st>>>(0u, m, d_leaf_koff.p, d_leaf_klen.p, d_leaf_rec.p, d_leaves.p);
    synced_inner = ni;
    synced_leaf = m;
    dirty = false;
    mod_child.clear();
    mod_leaf.clear();
    S.launches += launches_ + 12;
    lap("host mirror");
    return true;
}

void HostIndex::insert_batch(Store &S, uint32_t n, const uint8_t *d_keys, const int64_t *d_koff, const uint8_t *h_keys,
                             const int64_t *h_koff, uint32_t first_rec, int64_t *old_out) {
    cudaStream_t st = S.st;
    std::vector<uint8_t> q;
    uint32_t a = 0;
    const bool trace = S.knobs.trace;
    double t_probe = 0, t_apply = 0, t_sync = 0;
    bool reserved = false;
    uint64_t n_fallback = 0, n_rounds = 0;
    auto now = [] { return std::chrono::steady_clock::now(); };
    // a batch at least as large as the tree: rebuild the whole tree on the GPU from the sorted keys
    if (n >= S.knobs.bulk_min && (size_t) n >= n_live) {
        const auto t0 = now();
        if (bulk_build(S, n, d_keys, d_koff, first_rec, old_out)) {
            if (trace)
                fprintf(stderr, "[index] %u keys: bulk build on the GPU, %zu leaves, %.1f ms\n", n, n_live,
                        std::chrono::duration<double, std::milli>(now() - t0).count());
            return;
        }
    }
    while (a < n) {
        // while the tree is small (or the rest of the batch is), plain host inserts; afterwards sub-batches of
        // at most a quarter of the tree, so that few keys of a sub-batch meet on the same edge
        const size_t live = n_live;
        uint32_t b;
        bool probe = live >= 4096 && n - a >= 1024;
        if (!probe) {
            b = (uint32_t) std::min<uint64_t>(n, (uint64_t) a + (live < 4096 ? 4096 : n - a));
            for (uint32_t i = a; i < b; i++) {
                escape_key(h_keys + h_koff[i], (size_t) (h_koff[i + 1] - h_koff[i]), q);
                old_out[i] = set(q.data(), (uint32_t) q.size(), first_rec + i);
            }
            a = b;
            continue;
        }
        b = (uint32_t) std::min<uint64_t>(n, (uint64_t) a + live / 4);
        const uint32_t m = b - a;
        const auto t0 = now();
        if (!reserved) {
            // one growth step for the whole batch instead of one per sub-batch (cudaMalloc/cudaFree synchronise)
            const size_t rest = n - a, ni = diff_at.size() + rest + 1, nl = leaf_rec.size() + rest + 1;
            const size_t kb = arena.size() + (size_t) (h_koff[n] - h_koff[a]) + 4 * rest + 64;  // room for doubled 251s is taken on demand
            if (dirty) device_view(st);
            d_child0.reserve_keep(ni, synced_inner, st);
            d_child1.reserve_keep(ni, synced_inner, st);
            d_diff.reserve_keep(ni, synced_inner, st);
            d_mask.reserve_keep(ni, synced_inner, st);
            d_leaf_rec.reserve_keep(nl, synced_leaf, st);
            d_leaf_klen.reserve_keep(nl, synced_leaf, st);
            d_leaf_koff.reserve_keep(nl, synced_leaf, st);
            this->d_keys.reserve_keep(kb, keys_uploaded, st);
            reserved = true;
        }
        DeviceView T = device_view(st);
        t_sync += std::chrono::duration<double, std::milli>(now() - t0).count();
        S.prof.begin(PC_LOOKUP, st);
        d_qlen.reserve_discard(m + 1);
        d_qoff.reserve_discard((size_t) m + 2);
        uint32_t *ql = d_qlen.p;
        uint64_t *qo = d_qoff.p;
        k_query_len<<<div_up<uint32_t>(m, 256), 256, 0, st>>>(m, d_keys, d_koff + a, ql);
        device_scan<uint64_t>(
            (size_t) m + 1, [=] __device__(size_t i) -> uint64_t { return i < m ? (uint64_t) ql[i] : 0ull; },
            [=] __device__(size_t i, uint64_t v) { qo[i] = v; }, OpSum(), 0ull, true, S.es.scanws, st);
        uint64_t qbytes = 0;
        PX_CUDA(cudaMemcpyAsync(&qbytes, qo + m, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
        PX_CUDA(cudaStreamSynchronize(st));
        d_q.reserve_discard(qbytes + 16);
        k_query_write<<<div_up<uint32_t>(m, 256), 256, 0, st>>>(m, d_keys, d_koff + a, qo, d_q.p);
        d_probe.reserve_discard(m);
        k_insert_probe<<<div_up<uint32_t>(m, 128), 128, 0, st>>>(m, T, d_q.p, qo, ql, d_probe.p);
        PX_LAUNCH_CHECK();
        S.prof.end(st, 0.0, 4);
        S.launches += 4;
        h_probe.resize(m);
        PX_CUDA(cudaMemcpyAsync(h_probe.data(), d_probe.p, (size_t) m * sizeof(Probe), cudaMemcpyDeviceToHost, st));
        PX_CUDA(cudaStreamSynchronize(st));
        const auto t1 = now();
        for (uint32_t i = 0; i < m; i++) {
            const Probe &pr = h_probe[i];
            escape_key(h_keys + h_koff[a + i], (size_t) (h_koff[a + i + 1] - h_koff[a + i]), q);
            const uint32_t rec = first_rec + a + i;
            if (pr.info & (1u << 26)) {  // the key exists: repoint its leaf (replace(), CritBitTree.cpp:32-43)
                old_out[a + i] = leaf_rec[pr.leaf];
                leaf_rec[pr.leaf] = rec;
                note_leaf(pr.leaf);
                continue;
            }
            const int pdir = (pr.info >> 25) & 1;
            const int32_t now = pr.parent < 0 ? root : child[pdir][pr.parent];
            if (now == pr.cur) old_out[a + i] = splice(pr, q.data(), (uint32_t) q.size(), rec);
            else {
                // the edge changed within this sub-batch (earlier keys were spliced onto it): the key still reaches
                // this edge, so only the few new nodes below it are walked again
                old_out[a + i] = set_below(pr.parent, pdir, q.data(), (uint32_t) q.size(), rec);
                n_fallback++;
            }
        }
        t_probe += std::chrono::duration<double, std::milli>(t1 - t0).count();
        t_apply += std::chrono::duration<double, std::milli>(now() - t1).count();
        n_rounds++;
        a = b;
    }
    if (trace)
        fprintf(stderr, "[index] %u keys: %llu probe rounds %.1f ms (mirror sync %.1f ms), host splices %.1f ms, %llu keys re-walked below their edge\n", n,
                (unsigned long long) n_rounds, t_probe, t_sync, t_apply, (unsigned long long) n_fallback);
}

// the walk itself: d_keys/d_koff are the packed raw keys on the device (koff[0] == 0); the record id of every key
// (0xFFFFFFFF = absent) is left in S.doc_off on the device
void lookup_core(Store &S, uint32_t nn, const uint8_t *d_keys, const int64_t *d_koff) {
    cudaStream_t st = S.st;
    HostIndex::DeviceView T = S.index->device_view(st);
    PX_CUDA(cudaEventRecord(S.ev0, st));
    S.prof.begin(PC_LOOKUP, st);
    S.doc_len.reserve_discard(nn + 1);
    DevBuf<uint64_t> &qoff = S.es.qoff;
    qoff.reserve_discard((size_t) nn + 2);
    uint64_t *d_qoff = qoff.p;
    k_query_len<<<div_up<uint32_t>(nn, 256), 256, 0, st>>>(nn, d_keys, d_koff, S.doc_len.p);
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
    k_query_write<<<div_up<uint32_t>(nn, 256), 256, 0, st>>>(nn, d_keys, d_koff, d_qoff, S.in_vals.p);
    S.doc_off.reserve_discard(nn + 1);
    k_lookup<<<div_up<uint32_t>(nn, 128), 128, 0, st>>>(nn, T, S.in_vals.p, d_qoff, S.doc_len.p, S.doc_off.p);
    PX_LAUNCH_CHECK();
    S.last_lookup_qbytes = qbytes;
    S.prof.end(st, 2.0 * (double) qbytes, 4);  // escaped query read for the walk and for the verification (+ depth x 7 B: bench.py)
    S.launches += 4;
    PX_CUDA(cudaEventRecord(S.ev1, st));
}

static void lookup_finish(Store &S) {
    PX_CUDA(cudaStreamSynchronize(S.st));
    float ms = 0;
    PX_CUDA(cudaEventElapsedTime(&ms, S.ev0, S.ev1));
    S.last_lookup_ms = ms;
    S.prof.collect();
}

void lookup_batch(Store &S, int64_t n, const uint8_t *h_keys, const int64_t *h_koff, std::vector<uint32_t> &rec_out) {
    rec_out.assign((size_t) n, 0xFFFFFFFFu);
    if (n == 0) return;
    cudaStream_t st = S.st;
    const uint32_t nn = (uint32_t) n;
    const int64_t kbytes = h_koff[n] - h_koff[0];
    S.in_keys.reserve_discard((size_t) kbytes + 16);
    S.in_koff.reserve_discard(nn + 1);
    PX_CUDA(cudaMemcpyAsync(S.in_keys.p, h_keys + h_koff[0], (size_t) kbytes, cudaMemcpyHostToDevice, st));
    if (h_koff[0] == 0) {
        PX_CUDA(cudaMemcpyAsync(S.in_koff.p, h_koff, (nn + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, st));
    } else {
        std::vector<int64_t> rel(nn + 1);
        for (uint32_t i = 0; i <= nn; i++) rel[i] = h_koff[i] - h_koff[0];
        PX_CUDA(cudaMemcpyAsync(S.in_koff.p, rel.data(), (nn + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, st));
        PX_CUDA(cudaStreamSynchronize(st));
    }
    lookup_core(S, nn, S.in_keys.p, S.in_koff.p);
    PX_CUDA(cudaMemcpyAsync(rec_out.data(), S.doc_off.p, nn * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    lookup_finish(S);
}

__global__ void __launch_bounds__(256) k_found(uint32_t n, const uint32_t *__restrict__ rec, uint8_t *__restrict__ found) {
    const uint32_t i = blockIdx.x * 256 + threadIdx.x;
    if (i < n) found[i] = rec[i] != 0xFFFFFFFFu;
}

// contains for keys already resident in HBM: found[] is written on the device
void contains_batch_dev(Store &S, int64_t n, const uint8_t *d_keys, const int64_t *d_koff, uint8_t *d_found) {
    if (n == 0) return;
    const uint32_t nn = (uint32_t) n;
    lookup_core(S, nn, d_keys, d_koff);
    k_found<<<div_up<uint32_t>(nn, 256), 256, 0, S.st>>>(nn, S.doc_off.p, d_found);
    S.launches++;
    lookup_finish(S);
}

// iter(prefix) on the device (replaces the coroutine walk of CBTGen / CBTGHelper, CritBitTree.h:55-157): one CTA.
//   1. thread 0 walks from the root along the prefix's bits until the critical position of a node lies behind the
//      prefix (from there on the whole subtree qualifies) or a leaf is reached;
//   2. the subtree is unfolded level by level in ORDER: the frontier (inner nodes and leaves, left to right) is
//      rewritten with every inner node replaced by its two children (a block-wide scan gives the positions), until it
//      holds leaves only - their order is the ascending byte order of esc(key) 251 0 (the tree order, Appendix A.20);
//   3. the first leaf must start with the prefix (keys the walk never compared; CritBitTree.h:76-79), else the result
//      is empty; the leaves' record ids are the output.
// out[0] = number of records, out[1 ...] = record ids in key order; buf0 / buf1 hold at least (inner + leaves) entries.
constexpr int ITER_THREADS = 512;
__global__ void __launch_bounds__(ITER_THREADS)
k_iter_prefix(HostIndex::DeviceView T, const uint8_t *__restrict__ prefix, uint32_t plen, int32_t *__restrict__ buf0,
              int32_t *__restrict__ buf1, uint32_t *__restrict__ out) {
    __shared__ uint32_t wsum[ITER_THREADS / 32];
    __shared__ uint32_t s_n, s_inner, s_ok;
    const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) {
        uint32_t n = 0;
        if (T.has_root) {
            int32_t p = T.root;
            while (p >= 0) {
                const int4 nd = __ldg(T.nodes + p);
                const uint32_t da = (uint32_t) nd.z & 0xFFFFu;
                if (da >= plen) break;
                p = ((1u + ((((uint32_t) nd.z >> 16) & 0xFFu) | prefix[da])) >> 8) ? nd.y : nd.x;
            }
            buf0[0] = p;
            n = 1;
        }
        s_n = n;
        s_inner = 1;
    }
    __syncthreads();
    int32_t *cur = buf0, *nxt = buf1;
    uint32_t n = s_n;
    while (n && s_inner) {
        __syncthreads();
        if (tid == 0) s_inner = 0;
        __syncthreads();
        uint32_t base = 0;
        bool any_inner = false;
        for (uint32_t i0 = 0; i0 < n; i0 += ITER_THREADS) {
            const uint32_t i = i0 + tid;
            const int32_t e = i < n ? cur[i] : -1;
            const bool inner = i < n && e >= 0;
            const uint32_t cnt = i < n ? (inner ? 2u : 1u) : 0u;
            uint32_t inc = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
                if ((int) lane >= d) inc += o;
            }
            if (lane == 31) wsum[wid] = inc;
            __syncthreads();
            uint32_t woff = 0, total = 0;
#pragma unroll
            for (int k = 0; k < ITER_THREADS / 32; k++) {
                const uint32_t v = wsum[k];
                if ((uint32_t) k < wid) woff += v;
                total += v;
            }
            const uint32_t pos = base + woff + inc - cnt;
            if (inner) {
                const int4 nd = __ldg(T.nodes + e);
                nxt[pos] = nd.x;
                nxt[pos + 1] = nd.y;
                if (nd.x >= 0 || nd.y >= 0) any_inner = true;
            } else if (i < n) {
                nxt[pos] = e;
            }
            base += total;
            __syncthreads();
        }
        if (any_inner) s_inner = 1;   // (benign race: every writer stores 1)
        n = base;
        int32_t *t = cur;
        cur = nxt;
        nxt = t;
        __syncthreads();
    }
    // the entries of `cur` are leaves now
    if (tid == 0) {
        uint32_t ok = 0;
        if (n) {
            const uint4 lf = __ldg(T.leaves + (uint32_t) ~cur[0]);
            if (lf.z >= plen) {
                const uint8_t *lk = T.keys + (((uint64_t) lf.y << 32) | lf.x);
                uint32_t j = 0;
                while (j < plen && lk[j] == prefix[j]) j++;
                ok = j == plen;
            }
        }
        s_ok = ok;
        out[0] = ok ? n : 0u;
    }
    __syncthreads();
    if (s_ok)
        for (uint32_t i = tid; i < n; i += ITER_THREADS) out[1 + i] = __ldg(T.leaves + (uint32_t) ~cur[i]).w;
}

// record ids of all live keys starting with the escaped prefix, ascending key order - walked on the device
void iter_prefix(Store &S, const uint8_t *h_prefix, uint32_t plen, std::vector<uint32_t> &out) {
    out.clear();
    if (S.index->size() == 0) return;
    cudaStream_t st = S.st;
    HostIndex::DeviceView T = S.index->device_view(st);
    const size_t cap = S.index->size() * 2 + 64;   // leaves + inner nodes
    S.iter_buf.reserve_discard(3 * cap + 64);
    S.in_keys.reserve_discard((size_t) plen + 16);
    if (plen) PX_CUDA(cudaMemcpyAsync(S.in_keys.p, h_prefix, plen, cudaMemcpyHostToDevice, st));
    int32_t *b0 = reinterpret_cast<int32_t *>(S.iter_buf.p), *b1 = b0 + cap;
    uint32_t *o = S.iter_buf.p + 2 * cap;
    k_iter_prefix<<<1, ITER_THREADS, 0, st>>>(T, S.in_keys.p, plen, b0, b1, o);
    PX_LAUNCH_CHECK();
    S.launches++;
    uint32_t n = 0;
    PX_CUDA(cudaMemcpyAsync(&n, o, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    PX_CUDA(cudaStreamSynchronize(st));
    out.resize(n);
    if (n) {
        PX_CUDA(cudaMemcpyAsync(out.data(), o + 1, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        PX_CUDA(cudaStreamSynchronize(st));
    }
}

// depth of the leaf each (raw) key's walk ends in: inner nodes visited (measurement only)
void index_depths(Store &S, int64_t n, const uint8_t *h_keys, const int64_t *h_koff, int32_t *out) {
    std::vector<uint8_t> q;
    for (int64_t i = 0; i < n; i++) {
        escape_key(h_keys + h_koff[i], (size_t) (h_koff[i + 1] - h_koff[i]), q);
        out[i] = S.index->depth(q.data(), (uint32_t) q.size());
    }
}

}  // namespace pixiu
