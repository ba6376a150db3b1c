// CritBit key index (replaces data_struct/CritBitTree.{h,cpp}).
//
// Host side: the tree lives in flat arrays (no per-node malloc, no tagged pointers):
//   inner node i : child[0][i], child[1][i] (>= 0 inner node, < 0 leaf ~slot), diff_at[i], mask[i]
//   leaf slot s  : leaf_rec[s] = global record id; its escaped key (with the 251,0 terminator)
//                  sits in a host key arena so inserts need no decode of stored records.
// Device side: the same arrays (and the key arena) mirrored as SoA for the batched level-synchronous
// walk (find_best_match, CritBitTree.cpp:253-269); a candidate leaf is verified against its escaped
// key in the arena, which is byte for byte the decoded prefix "esc(k) 251 0" of the record the leaf
// points to (key_eq / contains, PiXiuStr.cpp:129-143, CritBitTree.cpp:154-178) - the reference decodes
// that prefix from the compressed record instead, which costs a token walk per back-reference hop.
#pragma once
#include <cstdint>
#include <vector>

#include "common.cuh"

namespace pixiu {

// esc(k) 251 0   (PiXiuStr_init_key, proj/PiXiuStr.cpp:12-14)
void escape_key(const uint8_t *k, size_t n, std::vector<uint8_t> &out, bool terminator = true);

struct Store;

class HostIndex {
   public:
    // returns the record id that was replaced, or -1 when the key is new (CritBitTree.cpp:13-105)
    int64_t set(const uint8_t *q, uint32_t qlen, uint32_t rec);
    int64_t get(const uint8_t *q, uint32_t qlen) const;  // record id or -1
    int32_t depth(const uint8_t *q, uint32_t qlen) const;  // inner nodes on the key's walk
    int64_t del(const uint8_t *q, uint32_t qlen);        // record id or -1 (CritBitTree.cpp:107-152)
    // record ids of all keys starting with the escaped prefix, ascending key order (CritBitTree.h:55-157)
    void iter(const uint8_t *prefix, uint32_t plen, std::vector<uint32_t> &out) const;
    size_t size() const { return n_live; }
    // memory the index holds beside the compressed store: host arrays (nodes, leaves, escaped-key arena) and their
    // device mirror (SoA + packed walk copy + key arena)
    size_t host_bytes() const;
    size_t device_bytes() const;
    size_t key_arena_bytes() const { return arena.size(); }
    // wire format of the whole tree (flat SoA arrays, leaf -> record ids, escaped-key arena): the index of a store
    // whose chunks were exported with pixiu_export_chunk; load() replaces the tree (the device mirror is rebuilt on
    // the next lookup).  live_out: the record ids the leaves point to.
    void save(std::vector<uint8_t> &out) const;
    bool load(const uint8_t *blob, size_t size, uint32_t n_records, std::vector<uint32_t> &live_out);

    // ---- device mirror ----
    struct DeviceView {
        const int32_t *child0, *child1;
        const uint16_t *diff_at;
        const uint8_t *mask;
        const uint32_t *leaf_rec, *leaf_klen;
        const uint64_t *leaf_koff;
        const uint8_t *keys;  // escaped keys of the leaves
        const uint4 *leaves;  // packed leaves {key offset lo, hi, key length, record id}
        const int4 *nodes;    // packed inner nodes {child0, child1, diff_at | mask << 16, 0}: what the lookup walks read
        int32_t root;
        int32_t has_root;
    };
    DeviceView device_view(cudaStream_t st);  // brings the mirror up to date (appended ranges + scattered changes)

    // Batched insert (the batch form of CritBitTree::setitem, CritBitTree.cpp:13-105): the GPU finds, for every
    // key, its best-match leaf, the critical position against that leaf and the edge the new node goes on (two
    // read-only walks per key); the host then splices in batch order, O(1) per key.  Splices on distinct edges
    // commute; a key whose edge was already changed in this batch takes the ordinary set().  d_keys/d_koff: the
    // raw keys on the device (koff indexed [0, n]); h_keys/h_koff the same on the host.  old_out[i] = record id
    // that key i replaced, or -1.
    void insert_batch(Store &S, uint32_t n, const uint8_t *d_keys, const int64_t *d_koff, const uint8_t *h_keys,
                      const int64_t *h_koff, uint32_t first_rec, int64_t *old_out);
    // the whole tree rebuilt on the GPU from the sorted keys (live leaves + batch); false: not applicable (long keys)
    bool bulk_build(Store &S, uint32_t n, const uint8_t *d_keys_raw, const int64_t *d_koff, uint32_t first_rec, int64_t *old_out);
    struct Probe {
        int32_t leaf;     // best-match leaf slot
        int32_t parent;   // node above the new inner node (-1: the root edge)
        int32_t cur;      // what hangs on that edge now (>= 0 inner node, < 0 leaf ~slot)
        uint32_t info;    // diff_at (16) | mask << 16 (8) | dir << 24 | pdir << 25 | exists << 26
    };

   private:
    std::vector<int32_t> child[2];
    std::vector<uint16_t> diff_at;
    std::vector<uint8_t> mask;
    std::vector<uint32_t> leaf_rec, leaf_klen;
    std::vector<uint64_t> leaf_koff;
    std::vector<uint8_t> arena;
    std::vector<int32_t> free_inner, free_leaf;
    int32_t root = 0;
    bool has_root = false;
    size_t n_live = 0;
    bool dirty = true;            // the device mirror needs a full upload
    size_t synced_inner = 0, synced_leaf = 0;   // entries of the mirror that are current (up to scattered changes)
    std::vector<int32_t> mod_child;             // changed child pointers of mirrored nodes: node * 2 + dir
    std::vector<int32_t> mod_leaf;              // mirrored leaves whose record id changed
    DevBuf<int32_t> d_mod;
    DevBuf<Probe> d_probe;
    DevBuf<uint8_t> d_q;          // escaped keys of a sub-batch
    DevBuf<uint64_t> d_qoff;
    DevBuf<uint32_t> d_qlen;
    std::vector<Probe> h_probe;
    void note_child(int32_t node, int dir) {
        if ((size_t) node < synced_inner) mod_child.push_back(node * 2 + dir);
    }
    void note_leaf(int32_t slot) {
        if ((size_t) slot < synced_leaf) mod_leaf.push_back(slot);
    }
    int64_t splice(const Probe &pr, const uint8_t *q, uint32_t qlen, uint32_t rec);
    int64_t set_below(int32_t top, int tdir, const uint8_t *q, uint32_t qlen, uint32_t rec);
    DevBuf<int32_t> d_child0, d_child1;
    DevBuf<uint16_t> d_diff;
    DevBuf<uint8_t> d_mask;
    DevBuf<uint32_t> d_leaf_rec, d_leaf_klen;
    DevBuf<uint64_t> d_leaf_koff;
    DevBuf<uint8_t> d_keys;
    DevBuf<int4> d_nodes;
    DevBuf<uint4> d_leaves;
    size_t keys_uploaded = 0;

    int32_t new_leaf(const uint8_t *q, uint32_t qlen, uint32_t rec);
    int32_t new_inner();
    int dir_of(int32_t node, const uint8_t *q, uint32_t qlen) const {
        uint8_t b = qlen > diff_at[node] ? q[diff_at[node]] : 0;
        return (1 + (mask[node] | b)) >> 8;
    }
};

struct Store;
// GPU batched lookup: escapes the n packed keys on the device, walks the SoA tree and verifies the
// candidates against the compressed store.  rec_out[i] = record id or 0xFFFFFFFF.
void lookup_batch(Store &S, int64_t n, const uint8_t *h_keys, const int64_t *h_koff, std::vector<uint32_t> &rec_out);
void contains_batch_dev(Store &S, int64_t n, const uint8_t *d_keys, const int64_t *d_koff, uint8_t *d_found);
void index_depths(Store &S, int64_t n, const uint8_t *h_keys, const int64_t *h_koff, int32_t *out);
// iter(prefix) walked on the device (k_iter_prefix): record ids in ascending key order
void iter_prefix(Store &S, const uint8_t *h_prefix, uint32_t plen, std::vector<uint32_t> &out);

}  // namespace pixiu
