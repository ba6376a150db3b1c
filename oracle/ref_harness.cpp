// TEST INFRASTRUCTURE ONLY — not part of the product path.
//
// Thin C-ABI harness around the UNMODIFIED reference implementation
// (/root/reference/src, compiled in place by oracle/Makefile into
// oracle/_ref/libpixiu_ref.so).  It lets the tests and bench.py's
// cpu_baseline / --impl reference legs drive the reference's own
// PiXiuCtrl (PiXiuCtrl.h:7-26) and read back what it stored.
//
// The reference is not re-entrant (file-scope globals SuffixTree.cpp:5-6,
// PiXiuStr.cpp:4, function-static encoder state PiXiuStr.cpp:17-26), so the
// harness owns exactly one PiXiuCtrl per process.
#include "proj/PiXiuCtrl.h"

#include <chrono>
#include <cstdint>
#include <cstdlib>
#include <cstring>

namespace {
PiXiuCtrl *g_ctrl = nullptr;
long long g_chunk_serial = 0;  // number of window rotations seen so far

double now_s() {
    using namespace std::chrono;
    return duration<double>(steady_clock::now().time_since_epoch()).count();
}

// setitem + rotation bookkeeping (rotation = SuffixTree re-init, PiXiuCtrl.cpp:13-17)
int do_setitem(const uint8_t *k, int kl, const uint8_t *v, int vl) {
    int used_before = g_ctrl->st.local_chunk.used_num;
    int rc = g_ctrl->setitem((uint8_t *) k, kl, (uint8_t *) v, vl);
    int idx = g_ctrl->st.local_chunk.used_num - 1;
    if (idx < used_before) g_chunk_serial++;
    return rc;
}
}  // namespace

extern "C" {

int ref_init(void) {
    if (g_ctrl) return -1;
    g_ctrl = (PiXiuCtrl *) calloc(1, sizeof(PiXiuCtrl));
    new (g_ctrl) PiXiuCtrl();
    g_ctrl->init_prop();
    g_chunk_serial = 0;
    return 0;
}

void ref_free(void) {
    if (!g_ctrl) return;
    g_ctrl->free_prop();
    free(g_ctrl);
    g_ctrl = nullptr;
}

int ref_setitem(const uint8_t *k, int kl, const uint8_t *v, int vl) { return do_setitem(k, kl, v, vl); }

// State of the record inserted last: chunk serial, idx in chunk, encoded length,
// arena pools allocated (MemPool::nth) and blocks used in the current pool.
void ref_last_info(long long *chunk_serial, int *idx, int *enc_len, int *pools, int *pool_used) {
    int i = g_ctrl->st.local_chunk.used_num - 1;
    *chunk_serial = g_chunk_serial;
    *idx = i;
    *enc_len = g_ctrl->st.cbt_chunk->getitem(i)->len;
    *pools = g_ctrl->st.local_pool.nth;
    *pool_used = g_ctrl->st.local_pool.used_num;
}

// Encoded bytes of record `idx` of the CURRENT chunk; returns length or -1.
int ref_encoded(int idx, uint8_t *out, int cap) {
    if (idx < 0 || idx >= g_ctrl->st.local_chunk.used_num) return -1;
    PiXiuStr *p = g_ctrl->st.cbt_chunk->getitem(idx);
    if (p->len > cap) return -1;
    memcpy(out, p->data, p->len);
    return p->len;
}

int ref_contains(const uint8_t *k, int kl) { return g_ctrl->contains((uint8_t *) k, kl) ? 1 : 0; }

int ref_delitem(const uint8_t *k, int kl) { return g_ctrl->delitem((uint8_t *) k, kl); }

// Drains the reference generator (PiXiuStr.h:110-198). Returns decoded length, -1 if
// absent.  NOTE: the reference decoder has bugs B1/B2 — its
// output is only trusted on small records without self references.
int ref_getitem(const uint8_t *k, int kl, uint8_t *out, int cap) {
    PXSGen *gen = g_ctrl->getitem((uint8_t *) k, kl);
    if (!gen) return -1;
    int n = 0;
    uint8_t rv;
    while (gen->operator()(rv)) {
        if (n < cap) out[n] = rv;
        n++;
        if (n > 200000) break;  // bug B1 can derail the reference decoder
    }
    PXSGen_free(gen);
    return n;
}

// Drains iter(prefix) (CritBitTree.h:55-157): decoded records concatenated into
// `out`, offs[i]..offs[i+1]; returns number of records (or -1 on overflow).
int ref_iter(const uint8_t *prefix, int pl, uint8_t *out, long long cap, long long *offs, int max_n) {
    CBTGen *it = g_ctrl->iter((uint8_t *) prefix, pl);
    if (!it) return 0;
    int n = 0;
    long long pos = 0;
    offs[0] = 0;
    PXSGen *gen;
    while (it->operator()(gen)) {
        uint8_t rv;
        int m = 0;
        while (gen->operator()(rv)) {
            if (pos < cap) out[pos] = rv;
            pos++;
            if (++m > 200000) break;
        }
        PXSGen_free(gen);
        if (n < max_n) offs[++n] = pos;
        else { n = -1; break; }
    }
    CBTGen_free(it);
    return pos > cap ? -1 : n;
}

// Timed batch insert (timing only the setitem loop). Per record: rc, encoded
// length, chunk serial, idx, pools (MemPool::nth), blocks used in the current pool. Any out pointer may be NULL.
double ref_setitem_batch(int n, const uint8_t *keys, const long long *koff, const uint8_t *vals,
                         const long long *voff, int *rc, int *enc_len, long long *chunk, int *idx,
                         int *pools, int *pool_used) {
    double t0 = now_s();
    for (int i = 0; i < n; i++) {
        int r = do_setitem(keys + koff[i], (int) (koff[i + 1] - koff[i]), vals + voff[i],
                           (int) (voff[i + 1] - voff[i]));
        if (rc) rc[i] = r;
        int j = g_ctrl->st.local_chunk.used_num - 1;
        if (enc_len) enc_len[i] = g_ctrl->st.cbt_chunk->getitem(j)->len;
        if (chunk) chunk[i] = g_chunk_serial;
        if (idx) idx[i] = j;
        if (pools) pools[i] = g_ctrl->st.local_pool.nth;
        if (pool_used) pool_used[i] = g_ctrl->st.local_pool.used_num;
    }
    return now_s() - t0;
}

// Timed batch getitem + drain; returns seconds, *total = decoded bytes yielded.
double ref_getitem_batch(int n, const uint8_t *keys, const long long *koff, long long *total,
                         int *found) {
    long long tot = 0;
    int nf = 0;
    double t0 = now_s();
    for (int i = 0; i < n; i++) {
        PXSGen *gen = g_ctrl->getitem((uint8_t *) (keys + koff[i]), (int) (koff[i + 1] - koff[i]));
        if (!gen) continue;
        nf++;
        uint8_t rv;
        int m = 0;
        while (gen->operator()(rv)) {
            tot++;
            if (++m > 200000) break;
        }
        PXSGen_free(gen);
    }
    double dt = now_s() - t0;
    *total = tot;
    if (found) *found = nf;
    return dt;
}

double ref_contains_batch(int n, const uint8_t *keys, const long long *koff, uint8_t *found) {
    double t0 = now_s();
    for (int i = 0; i < n; i++) {
        bool f = g_ctrl->contains((uint8_t *) (keys + koff[i]), (int) (koff[i + 1] - koff[i]));
        if (found) found[i] = f;
    }
    return now_s() - t0;
}

// Direct access to the reference stream encoder (PiXiuStr.cpp:16-118) for the
// t_PiXiuStr known-answer vectors (PiXiuStr.cpp:327-352): cmd = -1 ON, -2 OFF,
// -3 PASS, >=0 COMPRESS(chunk idx). On OFF copies the record out and returns its length.
int ref_stream(int cmd, int pxs_idx, int val, uint8_t *out, int cap) {
    PXSMsg m;
    m.chunk_idx_Cmd = cmd;
    m.pxs_idx = pxs_idx;
    m.val = (uint8_t) val;
    PiXiuStr *r = PiXiuStr_init_stream(m);
    if (!r) return -1;
    int n = r->len;
    if (n <= cap) memcpy(out, r->data, n);
    PiXiuStr_free(r);
    return n;
}

// escape_unique via PiXiuStr_init / PiXiuStr_init_key (PiXiuStr.cpp:8-14,:228-271)
int ref_escape(const uint8_t *src, int len, int is_key, uint8_t *out, int cap) {
    PiXiuStr *p = is_key ? PiXiuStr_init_key((uint8_t *) src, len) : PiXiuStr_init((uint8_t *) src, len);
    int n = p->len;
    if (n <= cap) memcpy(out, p->data, n);
    PiXiuStr_free(p);
    return n;
}

}  // extern "C"
