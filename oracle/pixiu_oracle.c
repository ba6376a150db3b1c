/* TEST INFRASTRUCTURE ONLY — see pixiu_oracle.h for the contract and parity status.
 *
 * CPU restatement of the PiXiu hot path in plain C.  Nothing here is shared with,
 * linked into, or called by the product library (pixiu_b200/csrc).  Citations are
 * relative to /root/reference/src/.
 */
#include "pixiu_oracle.h"

#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ */
/* codec: escape_unique (proj/PiXiuStr.cpp:228-271)                     */
/* ------------------------------------------------------------------ */
int pxo_escape(const uint8_t *src, int n, int is_key, uint8_t *out) {
    int o = 0;
    for (int i = 0; i < n; i++) {
        out[o++] = src[i];
        if (src[i] == PXO_UNIQUE) out[o++] = PXO_UNIQUE; /* 251 -> 251,251 */
    }
    if (is_key) { /* terminator 251,0 (PiXiuStr.cpp:262-266) */
        out[o++] = PXO_UNIQUE;
        out[o++] = PXO_KEY;
    }
    return o;
}

/* doc assembly of PiXiuCtrl::setitem (proj/PiXiuCtrl.cpp:31-44) */
int pxo_make_doc(const uint8_t *k, int kl, const uint8_t *v, int vl, uint8_t *out) {
    long long need = 2;
    for (int i = 0; i < kl; i++) need += 1 + (k[i] == PXO_UNIQUE);
    if (vl > 0) {
        need += 2;
        for (int i = 0; i < vl; i++) need += 1 + (v[i] == PXO_UNIQUE);
    }
    if (need > PXO_MAX_LEN) return -1; /* asserts PiXiuStr.cpp:123,:238 */
    int o = pxo_escape(k, kl, 1, out);
    if (vl > 0) {
        o += pxo_escape(v, vl, 1, out + o);
        out[o - 1] = PXO_KEY_SEC; /* PiXiuCtrl.cpp:36 */
    }
    return o;
}

/* ------------------------------------------------------------------ */
/* stream encoder (proj/PiXiuStr.cpp:16-118), offline over arrays       */
/* ------------------------------------------------------------------ */
int pxo_stream_encode(int n, const int32_t *cmd, const int32_t *pos, const uint8_t *val,
                      uint8_t *out, int strict251) {
    uint8_t *isc = (uint8_t *) malloc((size_t) n + 1);
    for (int i = 0; i < n; i++) isc[i] = cmd[i] >= 0;
    /* escape-pair coherence (PiXiuStr.cpp:34-54): a 251 opens a pair with the next
     * message; unless both are COMPRESS or both are PASS, both become PASS. */
    for (int i = 0; i < n;) {
        if (val[i] == PXO_UNIQUE && i + 1 < n) {
            if (isc[i] != isc[i + 1]) isc[i] = isc[i + 1] = 0;
            i += 2;
        } else {
            i += 1;
        }
    }
    int o = 0;
    int run = 0, r_idx = 0, r_to = 0;
    for (int i = 0; i <= n; i++) {
        if (i < n && isc[i]) { /* set_record (PiXiuStr.cpp:84-88) */
            r_idx = cmd[i];
            r_to = pos[i] + 1;
            run++;
            out[o++] = val[i];
            continue;
        }
        /* PASS or OFF: try_explode (PiXiuStr.cpp:56-82) */
        if (run > 6) {
            o -= run;
            int big = run > 255 || (!strict251 && run == PXO_UNIQUE);
            out[o++] = PXO_UNIQUE;
            if (big) {
                out[o++] = PXO_COMPRESS;
                out[o++] = (uint8_t) (r_idx & 255);
                out[o++] = (uint8_t) (r_idx >> 8);
                out[o++] = (uint8_t) (r_to & 255);
                out[o++] = (uint8_t) (r_to >> 8);
                out[o++] = (uint8_t) ((r_to - run) & 255);
                out[o++] = (uint8_t) ((r_to - run) >> 8);
            } else {
                out[o++] = (uint8_t) run;
                out[o++] = (uint8_t) (r_idx & 255);
                out[o++] = (uint8_t) (r_idx >> 8);
                out[o++] = (uint8_t) (r_to & 255);
                out[o++] = (uint8_t) (r_to >> 8);
            }
        }
        run = 0;
        if (i < n) out[o++] = val[i];
    }
    free(isc);
    return o;
}

/* ------------------------------------------------------------------ */
/* match finder: generalized suffix automaton                          */
/* ------------------------------------------------------------------ */
/* Restates the observable behaviour of s_insert_char (SuffixTree.cpp:144-289):
 *  - byte i is COMPRESS iff the Ukkonen active point extends by D[i] without a
 *    mismatch, i.e. iff the longest suffix of D[0..i] that already occurs in the
 *    window (earlier records, or D[0..i-1]; never across records) has length
 *    E(i) = E(i-1)+1;
 *  - the pointer is the END of the first occurrence, in insertion order, of that
 *    suffix (edge labels keep the coordinates of the leaf that created them,
 *    SuffixTree.cpp:154-156,:196-217).
 * A suffix automaton gives both online: the matched state after a transition and
 * its first end position. */
struct pxo_window {
    int ndocs;
    int ns, cap;
    int *len, *link, *fp_doc, *fp_pos, *head; /* head: first edge of a state */
    /* edge pool */
    int ne, ecap;
    int *e_state, *e_to, *e_next;
    uint8_t *e_ch;
    /* (state,ch) -> edge index, open addressing */
    int *h;
    size_t hcap;
};

static size_t h_slot(const pxo_window *w, int s, int c) {
    uint64_t k = ((uint64_t) (uint32_t) s << 8) | (uint32_t) c;
    k *= 0x9E3779B97F4A7C15ull;
    return (size_t) (k >> 20) & (w->hcap - 1);
}

static int edge_find(const pxo_window *w, int s, int c) {
    size_t i = h_slot(w, s, c);
    for (;;) {
        int e = w->h[i];
        if (e < 0) return -1;
        if (w->e_state[e] == s && w->e_ch[e] == c) return e;
        i = (i + 1) & (w->hcap - 1);
    }
}

static void h_insert(pxo_window *w, int e) {
    size_t i = h_slot(w, w->e_state[e], w->e_ch[e]);
    while (w->h[i] >= 0) i = (i + 1) & (w->hcap - 1);
    w->h[i] = e;
}

static void edge_set(pxo_window *w, int s, int c, int to) {
    int e = edge_find(w, s, c);
    if (e >= 0) {
        w->e_to[e] = to;
        return;
    }
    if (w->ne == w->ecap) {
        w->ecap *= 2;
        w->e_state = (int *) realloc(w->e_state, sizeof(int) * w->ecap);
        w->e_to = (int *) realloc(w->e_to, sizeof(int) * w->ecap);
        w->e_next = (int *) realloc(w->e_next, sizeof(int) * w->ecap);
        w->e_ch = (uint8_t *) realloc(w->e_ch, w->ecap);
    }
    if ((size_t) (w->ne + 1) * 2 > w->hcap) {
        free(w->h);
        w->hcap *= 2;
        w->h = (int *) malloc(sizeof(int) * w->hcap);
        memset(w->h, 0xff, sizeof(int) * w->hcap);
        for (int j = 0; j < w->ne; j++) h_insert(w, j);
    }
    e = w->ne++;
    w->e_state[e] = s;
    w->e_ch[e] = (uint8_t) c;
    w->e_to[e] = to;
    w->e_next[e] = w->head[s];
    w->head[s] = e;
    h_insert(w, e);
}

static int trans(const pxo_window *w, int s, int c) {
    int e = edge_find(w, s, c);
    return e < 0 ? -1 : w->e_to[e];
}

static int new_state(pxo_window *w, int len, int link, int fd, int fpos) {
    if (w->ns == w->cap) {
        w->cap *= 2;
        w->len = (int *) realloc(w->len, sizeof(int) * w->cap);
        w->link = (int *) realloc(w->link, sizeof(int) * w->cap);
        w->fp_doc = (int *) realloc(w->fp_doc, sizeof(int) * w->cap);
        w->fp_pos = (int *) realloc(w->fp_pos, sizeof(int) * w->cap);
        w->head = (int *) realloc(w->head, sizeof(int) * w->cap);
    }
    int s = w->ns++;
    w->len[s] = len;
    w->link[s] = link;
    w->fp_doc[s] = fd;
    w->fp_pos[s] = fpos;
    w->head[s] = -1;
    return s;
}

pxo_window *pxo_window_new(void) {
    pxo_window *w = (pxo_window *) calloc(1, sizeof(*w));
    w->cap = 1024;
    w->len = (int *) malloc(sizeof(int) * w->cap);
    w->link = (int *) malloc(sizeof(int) * w->cap);
    w->fp_doc = (int *) malloc(sizeof(int) * w->cap);
    w->fp_pos = (int *) malloc(sizeof(int) * w->cap);
    w->head = (int *) malloc(sizeof(int) * w->cap);
    w->ecap = 1024;
    w->e_state = (int *) malloc(sizeof(int) * w->ecap);
    w->e_to = (int *) malloc(sizeof(int) * w->ecap);
    w->e_next = (int *) malloc(sizeof(int) * w->ecap);
    w->e_ch = (uint8_t *) malloc(w->ecap);
    w->hcap = 4096;
    w->h = (int *) malloc(sizeof(int) * w->hcap);
    memset(w->h, 0xff, sizeof(int) * w->hcap);
    new_state(w, 0, -1, -1, -1); /* root */
    return w;
}

void pxo_window_free(pxo_window *w) {
    if (!w) return;
    free(w->len); free(w->link); free(w->fp_doc); free(w->fp_pos); free(w->head);
    free(w->e_state); free(w->e_to); free(w->e_next); free(w->e_ch); free(w->h);
    free(w);
}

int pxo_window_count(const pxo_window *w) { return w->ndocs; }

static int clone_state(pxo_window *w, int q, int newlen) {
    int cl = new_state(w, newlen, w->link[q], w->fp_doc[q], w->fp_pos[q]);
    for (int e = w->head[q]; e >= 0; e = w->e_next[e]) edge_set(w, cl, w->e_ch[e], w->e_to[e]);
    return cl;
}

/* generalized-automaton extension; *cl_from/*cl_to report a clone (or -1) */
static int sam_extend(pxo_window *w, int last, int c, int doc, int pos, int *cl_from, int *cl_to) {
    *cl_from = *cl_to = -1;
    int q = trans(w, last, c);
    if (q >= 0) {
        if (w->len[q] == w->len[last] + 1) return q;
        int cl = clone_state(w, q, w->len[last] + 1);
        for (int p = last; p >= 0 && trans(w, p, c) == q; p = w->link[p]) edge_set(w, p, c, cl);
        w->link[q] = cl;
        *cl_from = q;
        *cl_to = cl;
        return cl;
    }
    int cur = new_state(w, w->len[last] + 1, 0, doc, pos);
    int p = last;
    while (p >= 0 && trans(w, p, c) < 0) {
        edge_set(w, p, c, cur);
        p = w->link[p];
    }
    if (p < 0) {
        w->link[cur] = 0;
    } else {
        q = trans(w, p, c);
        if (w->len[p] + 1 == w->len[q]) {
            w->link[cur] = q;
        } else {
            int cl = clone_state(w, q, w->len[p] + 1);
            for (; p >= 0 && trans(w, p, c) == q; p = w->link[p]) edge_set(w, p, c, cl);
            w->link[q] = cl;
            w->link[cur] = cl;
            *cl_from = q;
            *cl_to = cl;
        }
    }
    return cur;
}

int pxo_window_encode(pxo_window *w, const uint8_t *doc, int n, uint8_t *out, int strict251,
                      int32_t *cmd_out, int32_t *pos_out) {
    int32_t *cmd = (int32_t *) malloc(sizeof(int32_t) * (size_t) (n + 1));
    int32_t *pos = (int32_t *) malloc(sizeof(int32_t) * (size_t) (n + 1));
    int d = w->ndocs;
    int last = 0;      /* automaton state of D[0..i-1] (restarts per record: no cross-record strings) */
    int v = 0, l = 0;  /* active point: state/length of the longest already-seen suffix */
    for (int i = 0; i < n; i++) {
        int c = doc[i];
        int t = trans(w, v, c);
        if (t >= 0) { /* active point extends: MSG_COMPRESS (SuffixTree.cpp:166,:183,:186) */
            v = t;
            l++;
            cmd[i] = w->fp_doc[v];
            pos[i] = w->fp_pos[v];
        } else { /* mismatch: MSG_NO_COMPRESS, then the suffix-link walk (:189,:252-285) */
            cmd[i] = PXO_PASS;
            pos[i] = 0;
            while (v > 0 && trans(w, v, c) < 0) {
                v = w->link[v];
                l = w->len[v];
            }
            t = trans(w, v, c);
            if (t >= 0) {
                v = t;
                l++;
            } else {
                v = 0;
                l = 0;
            }
        }
        int cf, ct;
        last = sam_extend(w, last, c, d, i, &cf, &ct);
        if (cf >= 0 && v == cf && l <= w->len[ct]) v = ct;
    }
    w->ndocs++;
    if (cmd_out) memcpy(cmd_out, cmd, sizeof(int32_t) * (size_t) n);
    if (pos_out) memcpy(pos_out, pos, sizeof(int32_t) * (size_t) n);
    int o = pxo_stream_encode(n, cmd, pos, doc, out, strict251);
    free(cmd);
    free(pos);
    return o;
}

/* ------------------------------------------------------------------ */
/* decoder (proj/PiXiuStr.h:129-198; Appendix A rules 12-13)           */
/* ------------------------------------------------------------------ */
struct pxo_chunk {
    int n, cap;
    uint8_t **dec;
    int *dlen;
};

pxo_chunk *pxo_chunk_new(void) {
    pxo_chunk *c = (pxo_chunk *) calloc(1, sizeof(*c));
    c->cap = 64;
    c->dec = (uint8_t **) malloc(sizeof(uint8_t *) * c->cap);
    c->dlen = (int *) malloc(sizeof(int) * c->cap);
    return c;
}

void pxo_chunk_free(pxo_chunk *c) {
    if (!c) return;
    for (int i = 0; i < c->n; i++) free(c->dec[i]);
    free(c->dec);
    free(c->dlen);
    free(c);
}

int pxo_chunk_count(const pxo_chunk *c) { return c->n; }
int pxo_chunk_declen(const pxo_chunk *c, int idx) { return (idx < 0 || idx >= c->n) ? -1 : c->dlen[idx]; }

int pxo_chunk_append(pxo_chunk *c, const uint8_t *enc, int enc_len) {
    int self = c->n;
    uint8_t *buf = (uint8_t *) malloc(PXO_MAX_LEN + 8);
    int o = 0;
    for (int i = 0; i < enc_len; i++) {
        uint8_t b = enc[i];
        if (b != PXO_UNIQUE) {
            if (o >= PXO_MAX_LEN) goto bad;
            buf[o++] = b;
            continue;
        }
        if (i + 1 >= enc_len) goto bad;
        uint8_t nx = enc[i + 1];
        if (nx == PXO_KEY || nx == PXO_UNIQUE || nx == PXO_KEY_SEC) { /* PiXiuStr.h:142-147 */
            if (o + 2 > PXO_MAX_LEN) goto bad;
            buf[o++] = b;
            buf[o++] = nx;
            i++;
            continue;
        }
        int idx, to, from;
        if (nx == PXO_COMPRESS) { /* big record, PiXiuStr.h:149-153 */
            if (i + 8 > enc_len) goto bad;
            idx = enc[i + 2] | (enc[i + 3] << 8);
            to = enc[i + 4] | (enc[i + 5] << 8);
            from = enc[i + 6] | (enc[i + 7] << 8);
            i += 7;
        } else if (nx > 6) { /* small record, :154-159 */
            if (i + 6 > enc_len) goto bad;
            idx = enc[i + 2] | (enc[i + 3] << 8);
            to = enc[i + 4] | (enc[i + 5] << 8);
            from = to - nx;
            i += 5;
        } else {
            goto bad; /* 3..6: assert(false), PiXiuStr.h:193 */
        }
        if (from < 0 || to < from || o + (to - from) > PXO_MAX_LEN) goto bad;
        if (idx == self) { /* self reference, overlap allowed (LZ77-style), :168-181 */
            if (from >= o) goto bad;
            for (int k = from; k < to; k++) buf[o++] = buf[k];
        } else {
            if (idx > self || to > c->dlen[idx]) goto bad;
            memcpy(buf + o, c->dec[idx] + from, (size_t) (to - from));
            o += to - from;
        }
    }
    if (c->n == c->cap) {
        c->cap *= 2;
        c->dec = (uint8_t **) realloc(c->dec, sizeof(uint8_t *) * c->cap);
        c->dlen = (int *) realloc(c->dlen, sizeof(int) * c->cap);
    }
    c->dec[c->n] = (uint8_t *) realloc(buf, o ? (size_t) o : 1);
    c->dlen[c->n] = o;
    return c->n++;
bad:
    free(buf);
    return -1;
}

int pxo_chunk_decode(const pxo_chunk *c, int idx, int from, int to, uint8_t *out) {
    if (idx < 0 || idx >= c->n || from < 0) return -1;
    if (to > c->dlen[idx]) to = c->dlen[idx];
    if (to <= from) return 0;
    memcpy(out, c->dec[idx] + from, (size_t) (to - from));
    return to - from;
}

/* ------------------------------------------------------------------ */
/* CritBit index (data_struct/CritBitTree.cpp)                         */
/* ------------------------------------------------------------------ */
typedef struct {
    long long child[2]; /* >=0: leaf slot, <0: inner ~slot */
    uint16_t diff_at;
    uint8_t mask;
} cb_inner;

typedef struct {
    uint8_t *key;
    int len;
    long long id;
} cb_leaf;

struct pxo_cbt {
    int has_root;
    long long root;
    cb_inner *in;
    long long nin, incap, in_free; /* freed inner slots chained through child[0] */
    cb_leaf *lf;
    long long nlf, lfcap, lf_free;
    long long size;
};

pxo_cbt *pxo_cbt_new(void) {
    pxo_cbt *t = (pxo_cbt *) calloc(1, sizeof(*t));
    t->incap = t->lfcap = 64;
    t->in = (cb_inner *) malloc(sizeof(cb_inner) * t->incap);
    t->lf = (cb_leaf *) malloc(sizeof(cb_leaf) * t->lfcap);
    t->in_free = t->lf_free = -1;
    return t;
}

void pxo_cbt_free(pxo_cbt *t) {
    if (!t) return;
    for (long long i = 0; i < t->nlf; i++) free(t->lf[i].key);
    free(t->in);
    free(t->lf);
    free(t);
}

long long pxo_cbt_size(const pxo_cbt *t) { return t->size; }

static long long leaf_new(pxo_cbt *t, const uint8_t *k, int n, long long id) {
    long long s;
    if (t->lf_free >= 0) {
        s = t->lf_free;
        t->lf_free = t->lf[s].id;
    } else {
        if (t->nlf == t->lfcap) {
            t->lfcap *= 2;
            t->lf = (cb_leaf *) realloc(t->lf, sizeof(cb_leaf) * t->lfcap);
        }
        s = t->nlf++;
    }
    t->lf[s].key = (uint8_t *) malloc(n ? n : 1);
    memcpy(t->lf[s].key, k, n);
    t->lf[s].len = n;
    t->lf[s].id = id;
    return s;
}

static void leaf_drop(pxo_cbt *t, long long s) {
    free(t->lf[s].key);
    t->lf[s].key = NULL;
    t->lf[s].len = 0;
    t->lf[s].id = t->lf_free;
    t->lf_free = s;
}

static long long inner_new(pxo_cbt *t) {
    if (t->in_free >= 0) {
        long long s = t->in_free;
        t->in_free = t->in[s].child[0];
        return s;
    }
    if (t->nin == t->incap) {
        t->incap *= 2;
        t->in = (cb_inner *) realloc(t->in, sizeof(cb_inner) * t->incap);
    }
    return t->nin++;
}

/* direction rule of find_best_match (CritBitTree.cpp:261-264) */
static int cb_dir(const cb_inner *nd, const uint8_t *q, int qlen) {
    uint8_t b = qlen > nd->diff_at ? q[nd->diff_at] : 0;
    return (1 + (nd->mask | b)) >> 8;
}

static long long cb_walk(const pxo_cbt *t, const uint8_t *q, int qlen, long long *pa, int *pa_dir,
                         long long *grand, long long *depth) {
    long long p = t->root, par = -1, gr = -1, dep = 0;
    int d = 0;
    while (p < 0) {
        const cb_inner *nd = &t->in[~p];
        gr = par;
        par = ~p;
        d = cb_dir(nd, q, qlen);
        p = nd->child[d];
        dep++;
    }
    if (pa) *pa = par;
    if (pa_dir) *pa_dir = d;
    if (grand) *grand = gr;
    if (depth) *depth = dep;
    return p;
}

long long pxo_cbt_depth(const pxo_cbt *t, const uint8_t *q, int qlen) {
    long long dep = 0;
    if (!t->has_root) return 0;
    cb_walk(t, q, qlen, NULL, NULL, NULL, &dep);
    return dep;
}

long long pxo_cbt_get(const pxo_cbt *t, const uint8_t *q, int qlen) {
    if (!t->has_root) return -1;
    long long lf = cb_walk(t, q, qlen, NULL, NULL, NULL, NULL);
    const cb_leaf *L = &t->lf[lf];
    /* key_eq / contains compare up to and including the 251,0 terminator
     * (PiXiuStr.cpp:129-143, CritBitTree.cpp:154-178); keys are prefix-free */
    if (L->len == qlen && memcmp(L->key, q, (size_t) qlen) == 0) return L->id;
    return -1;
}

long long pxo_cbt_set(pxo_cbt *t, const uint8_t *q, int qlen, long long id) {
    if (!t->has_root) { /* CritBitTree.cpp:15-17 */
        t->root = leaf_new(t, q, qlen, id);
        t->has_root = 1;
        t->size = 1;
        return -1;
    }
    long long pa;
    int pa_dir;
    long long lf = cb_walk(t, q, qlen, &pa, &pa_dir, NULL, NULL);
    cb_leaf *L = &t->lf[lf];
    int m = L->len < qlen ? L->len : qlen, diff = 0;
    while (diff < m && L->key[diff] == q[diff]) diff++;
    if (diff == qlen && diff == L->len) { /* replace() (:32-43) */
        long long old = L->id;
        L->id = id;
        return old;
    }
    /* insert() (:45-92). NOTE: the reference skips this when the first differing
     * byte directly follows a matched 251 (`if (!spec_mode)`, :100) and silently
     * loses the record — reference bug B5, not reproduced. */
    uint8_t a = diff < L->len ? L->key[diff] : 0, b = diff < qlen ? q[diff] : 0;
    uint8_t mask = (uint8_t) (a ^ b);
    mask |= mask >> 1;
    mask |= mask >> 2;
    mask |= mask >> 4;
    mask = (uint8_t) ((mask & ~(mask >> 1)) ^ 0xFF);
    int dir = (1 + (mask | b)) >> 8;
    long long nl = leaf_new(t, q, qlen, id);
    long long ni = inner_new(t);
    cb_inner *nd = &t->in[ni];
    nd->diff_at = (uint16_t) diff;
    nd->mask = mask;
    nd->child[dir] = nl;
    long long rp_parent = -1;
    int rp_dir = 0;
    long long p = t->root;
    while (p < 0) {
        cb_inner *r = &t->in[~p];
        if (r->diff_at > diff || (r->diff_at == diff && r->mask > mask)) break;
        rp_dir = cb_dir(r, q, qlen);
        rp_parent = ~p;
        p = r->child[rp_dir];
    }
    nd->child[1 - dir] = p;
    if (rp_parent < 0) t->root = ~ni;
    else t->in[rp_parent].child[rp_dir] = ~ni;
    t->size++;
    return -1;
}

long long pxo_cbt_del(pxo_cbt *t, const uint8_t *q, int qlen) {
    if (!t->has_root) return -1;
    long long pa, grand;
    int pa_dir;
    long long lf = cb_walk(t, q, qlen, &pa, &pa_dir, &grand, NULL);
    cb_leaf *L = &t->lf[lf];
    if (!(L->len == qlen && memcmp(L->key, q, (size_t) qlen) == 0)) return -1;
    long long id = L->id;
    if (pa < 0) { /* case_del (:122-139) */
        t->has_root = 0;
    } else {
        long long sib = t->in[pa].child[1 - pa_dir];
        if (grand < 0) t->root = sib;
        else {
            int gd = t->in[grand].child[0] == ~pa ? 0 : 1;
            t->in[grand].child[gd] = sib;
        }
        t->in[pa].child[0] = t->in_free;
        t->in_free = pa;
    }
    leaf_drop(t, lf);
    t->size--;
    return id;
}

/* CBTGHelper (data_struct/CritBitTree.h:55-128) */
static int cb_iter_rec(const pxo_cbt *t, long long p, const uint8_t *pre, int plen, int include_all,
                       int *harvest, long long *ids, long long cap, long long *n) {
    if (p >= 0) {
        const cb_leaf *L = &t->lf[p];
        if (!*harvest) {
            if (L->len < plen || memcmp(L->key, pre, (size_t) plen) != 0) return 0; /* yield NULL -> stop */
            *harvest = 1;
        }
        if (*n < cap) ids[*n] = L->id;
        (*n)++;
        return 1;
    }
    const cb_inner *nd = &t->in[~p];
    int d = cb_dir(nd, pre, plen), d_end;
    if (!include_all && nd->diff_at >= plen) include_all = 1;
    if (include_all) {
        d = 0;
        d_end = 2;
    } else {
        d_end = d + 1;
    }
    for (; d < d_end; d++)
        if (!cb_iter_rec(t, nd->child[d], pre, plen, include_all, harvest, ids, cap, n)) return 0;
    return 1;
}

long long pxo_cbt_iter(const pxo_cbt *t, const uint8_t *prefix, int plen, long long *ids, long long cap) {
    long long n = 0;
    int harvest = 0;
    if (!t->has_root) return 0;
    cb_iter_rec(t, t->root, prefix, plen, 0, &harvest, ids, cap, &n);
    return n;
}
