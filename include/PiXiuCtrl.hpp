// Header-only C++ façade with the reference's API over the C ABI (pixiu_b200.h).
//
// Mirrors /root/reference/src/proj/PiXiuCtrl.h:7-26 — same method names, argument meaning and
// return codes — and the generator types of proj/PiXiuStr.h:110-212 (PXSGen) and
// data_struct/CritBitTree.h:130-157 (CBTGen), so code written against the reference
// (README.md:104-149) compiles against this header unchanged:
//
//     PiXiuCtrl ctrl;  ctrl.init_prop();
//     ctrl.setitem(k, kl, v, vl);                       // 0, or CBT_SET_REPLACE
//     PXSGen *gen = ctrl.getitem(k, kl);                // NULL when absent
//     uint8_t rv;  while (gen->operator()(rv)) {...}    // esc(k) 251 0 esc(v) 251 2
//     PXSGen_free(gen);                                 // or gen->consume_repr()
//     CBTGen *it = ctrl.iter(prefix, pl);  PXSGen *g;  while (it->operator()(g)) {...}  CBTGen_free(it);
//     ctrl.free_prop();
//
// Differences from the reference: batched forms are added (`*_batch`), an oversize record is an
// error code instead of an assert, and several PiXiuCtrl objects may live in one process.
#ifndef PIXIU_CTRL_HPP
#define PIXIU_CTRL_HPP
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "pixiu_b200.h"

#define CBT_SET_REPLACE PIXIU_CBT_SET_REPLACE     /* data_struct/CritBitTree.h:7 */
#define CBT_DEL_NOT_FOUND PIXIU_CBT_DEL_NOT_FOUND /* data_struct/CritBitTree.h:8 */
#define PXS_UNIQUE 251
#define PXS_KEY 0
#define PXS_KEY_SEC 2

// byte generator over one decoded record (`$gen(PXSGen)`, proj/PiXiuStr.h:110-212)
struct PXSGen {
    uint8_t *data;
    int len, pos;
    bool operator()(uint8_t &rv) {
        if (pos >= len) return false;
        rv = data[pos++];
        return true;
    }
    // visible bytes (33..126) of the rest, NUL terminated, malloc'd; frees the generator (PiXiuStr.h:200-211)
    char *consume_repr(void) {
        char *out = (char *) malloc((size_t) (len - pos) + 1);
        int n = 0;
        for (; pos < len; pos++)
            if (33 <= data[pos] && data[pos] <= 126) out[n++] = (char) data[pos];
        out[n] = '\0';
        free(data);
        free(this);
        return out;
    }
};

inline void PXSGen_free(PXSGen *gen) {
    if (!gen) return;
    free(gen->data);
    free(gen);
}

// generator of PXSGen* in ascending key order (`$gen(CBTGen)`, data_struct/CritBitTree.h:130-157)
struct CBTGen {
    uint8_t *buf;
    int64_t *off;
    int64_t count, pos;
    bool operator()(PXSGen *&rv) {
        if (pos >= count) return false;
        int64_t n = off[pos + 1] - off[pos];
        PXSGen *g = (PXSGen *) malloc(sizeof(PXSGen));
        g->data = (uint8_t *) malloc(n ? (size_t) n : 1);
        memcpy(g->data, buf + off[pos], (size_t) n);
        g->len = (int) n;
        g->pos = 0;
        pos++;
        rv = g;
        return true;
    }
};

inline void CBTGen_free(CBTGen *gen) {
    if (!gen) return;
    free(gen->buf);
    free(gen->off);
    free(gen);
}

struct PiXiuCtrl {
    pixiu_store *store = nullptr;
    pixiu_config config;

    PiXiuCtrl() { pixiu_default_config(&config); }

    void init_prop(void) {  // PiXiuCtrl.cpp:77-81
        if (store) free_prop();
        store = pixiu_create(&config);
    }
    void free_prop(void) {  // PiXiuCtrl.cpp:83-86
        pixiu_destroy(store);
        store = nullptr;
    }

    // PiXiuCtrl.cpp:12-47.  v_len == 0 stores a key-only record.  Returns 0, CBT_SET_REPLACE, or a
    // negative PIXIU_E* code (the reference asserts).
    int setitem(uint8_t k[], int k_len, uint8_t v[], int v_len, bool = false) {
        int64_t ko[2] = {0, k_len}, vo[2] = {0, v_len};
        int32_t rc = 0;
        int e = pixiu_setitem_batch(store, 1, k, ko, v_len ? v : k, vo, &rc, nullptr);
        return e < 0 ? e : rc;
    }
    bool contains(uint8_t k[], int k_len) {  // PiXiuCtrl.cpp:55-57
        int64_t ko[2] = {0, k_len};
        uint8_t f = 0;
        pixiu_contains_batch(store, 1, k, ko, &f);
        return f != 0;
    }
    PXSGen *getitem(uint8_t k[], int k_len) {  // PiXiuCtrl.cpp:59-61
        int64_t ko[2] = {0, k_len}, oo[2] = {0, 0}, need = 0;
        uint8_t f = 0;
        uint8_t *buf = (uint8_t *) malloc(65536);
        int e = pixiu_getitem_batch(store, 1, k, ko, buf, 65536, oo, &f, &need);
        if (e < 0 || !f) {
            free(buf);
            return NULL;
        }
        PXSGen *g = (PXSGen *) malloc(sizeof(PXSGen));
        g->data = buf;
        g->len = (int) oo[1];
        g->pos = 0;
        return g;
    }
    int delitem(uint8_t k[], int k_len) {  // PiXiuCtrl.cpp:63-69
        int64_t ko[2] = {0, k_len};
        int32_t rc = 0;
        int e = pixiu_delitem_batch(store, 1, k, ko, &rc);
        return e < 0 ? e : rc;
    }
    CBTGen *iter(uint8_t prefix[], int prefix_len) {  // PiXiuCtrl.cpp:71-75; NULL on an empty tree
        int64_t count = 0, need = 0;
        int e = pixiu_iter(store, prefix, prefix_len, nullptr, 0, nullptr, 0, &count, &need);
        if ((e < 0 && e != PIXIU_ENOSPC)) return NULL;
        pixiu_stats st;
        pixiu_get_stats(store, &st);
        if (st.live_records == 0) return NULL;
        CBTGen *g = (CBTGen *) malloc(sizeof(CBTGen));
        g->buf = (uint8_t *) malloc(need ? (size_t) need : 1);
        g->off = (int64_t *) calloc((size_t) count + 1, sizeof(int64_t));
        g->count = count;
        g->pos = 0;
        if (count) pixiu_iter(store, prefix, prefix_len, g->buf, need, g->off, count + 1, &count, &need);
        return g;
    }

    // capacity hint (std::vector::reserve): map room for this many more compressed bytes now, not inside a batch
    int reserve(int64_t encoded_bytes) { return pixiu_reserve(store, encoded_bytes); }

    // encoded length of the record stored last - what the reference's REPL reads through
    // `ctrl.st.cbt_chunk->getitem(ctrl.st.local_chunk.used_num - 1)->len` to print the bytes saved (main.cpp:67)
    int last_encoded_len(void) {
        pixiu_stats st;
        int64_t chunk = 0, idx = 0;
        if (pixiu_get_stats(store, &st) != PIXIU_OK || st.records == 0) return -1;
        if (pixiu_record_location(store, st.records - 1, &chunk, &idx) != PIXIU_OK) return -1;
        uint8_t *buf = (uint8_t *) malloc(65536);
        const int n = pixiu_encoded_view(store, chunk, idx, buf, 65536);
        free(buf);
        return n;
    }

    // PiXiuCtrl.cpp:88-114, by chunk id instead of `PiXiuChunk *&` (set config.auto_reinsert = 1 before init_prop for
    // the reference's trigger).  Returns the number of records moved or a negative PIXIU_E* code.
    int64_t reinsert(int64_t chunk) { return pixiu_reinsert_chunk(store, chunk); }

    // ---- batched forms (packed data + int64 offsets), in-order semantics of n single calls ----
    int setitem_batch(int64_t n, const uint8_t *keys, const int64_t *key_off, const uint8_t *vals,
                      const int64_t *val_off, int32_t *rc, int32_t *saved = nullptr) {
        return pixiu_setitem_batch(store, n, keys, key_off, vals, val_off, rc, saved);
    }
    int contains_batch(int64_t n, const uint8_t *keys, const int64_t *key_off, uint8_t *found) {
        return pixiu_contains_batch(store, n, keys, key_off, found);
    }
    int getitem_batch(int64_t n, const uint8_t *keys, const int64_t *key_off, std::vector<uint8_t> &out,
                      std::vector<int64_t> &out_off, std::vector<uint8_t> &found) {
        out_off.assign((size_t) n + 1, 0);
        found.assign((size_t) n, 0);
        int64_t need = 0;
        int e = pixiu_getitem_batch(store, n, keys, key_off, out.data(), (int64_t) out.size(), out_off.data(),
                                    found.data(), &need);
        if (e == PIXIU_ENOSPC) {
            out.resize((size_t) need);
            e = pixiu_getitem_batch(store, n, keys, key_off, out.data(), (int64_t) out.size(), out_off.data(),
                                    found.data(), &need);
        }
        return e;
    }
};

#endif
