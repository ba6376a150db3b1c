/* TEST INFRASTRUCTURE ONLY — CPU restatement ("port") of the PiXiu hot path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library; the product (pixiu_b200/) never does.
 *
 * Parity status: PINNED.  Every function here is checked by tests/test_oracle*.py
 * against (1) the reference's own known-answer vectors (t_PiXiuStr,
 * /root/reference/src/proj/PiXiuStr.cpp:302-352; README.md:70-96), (2) golden
 * fixtures under tests/golden/ produced by the real reference compiled into
 * oracle/_ref/ (tests/golden/make_golden.py), and (3) when oracle/_ref is
 * present, live differential fuzzing against the real reference.
 *
 * All file:line citations are relative to /root/reference/src/.
 */
#ifndef PIXIU_ORACLE_H
#define PIXIU_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PXO_UNIQUE 251   /* proj/PiXiuStr.h:11 */
#define PXO_KEY 0        /* :12 */
#define PXO_KEY_SEC 2    /* :13 */
#define PXO_COMPRESS 1   /* :14 */
#define PXO_MAX_LEN 65535 /* PXSG_MAX_TO, :21 */
#define PXO_CHUNK_RECS 65535 /* PXC_STR_NUM, :20 */

/* message commands of the match finder (proj/PiXiuStr.h:16-18) */
#define PXO_PASS (-3)

/* ---- codec: escaping and doc assembly (PiXiuStr.cpp:228-271, PiXiuCtrl.cpp:31-44) ---- */
int pxo_escape(const uint8_t *src, int n, int is_key, uint8_t *out);
/* doc = esc(k) 251 0 [esc(v) 251 2]; returns length, or -1 when > 65535 */
int pxo_make_doc(const uint8_t *k, int kl, const uint8_t *v, int vl, uint8_t *out);

/* ---- stream encoder from an explicit message stream (PiXiuStr.cpp:16-118) ----
 * cmd[i] = PXO_PASS or the source record index (>=0); pos[i] = matched source
 * offset; val[i] = the byte.  strict251 != 0 reproduces reference bug B1
 * (run of exactly 251 emitted as the ambiguous FB FB .. small form). */
int pxo_stream_encode(int n, const int32_t *cmd, const int32_t *pos, const uint8_t *val,
                      uint8_t *out, int strict251);

/* ---- match finder + encoder over one window (= one chunk) ----
 * Restates SuffixTree::setitem / s_insert_char (data_struct/SuffixTree.cpp:144-304)
 * through the closed form of SURVEY.md §8a-A2, with a generalized suffix automaton. */
typedef struct pxo_window pxo_window;
pxo_window *pxo_window_new(void);
void pxo_window_free(pxo_window *w);
int pxo_window_count(const pxo_window *w);
/* Appends `doc` as the next record; writes its encoded form to `out` (cap 65535);
 * optionally also the per-byte messages (cmd/pos, each n entries, may be NULL).
 * Returns encoded length. */
int pxo_window_encode(pxo_window *w, const uint8_t *doc, int n, uint8_t *out, int strict251,
                      int32_t *cmd_out, int32_t *pos_out);

/* ---- decoder (PiXiuStr.h:110-198; clean spec Appendix A 12-13) ---- */
typedef struct pxo_chunk pxo_chunk;
pxo_chunk *pxo_chunk_new(void);
void pxo_chunk_free(pxo_chunk *c);
/* append an encoded record; it is decoded immediately. returns idx or <0 on malformed */
int pxo_chunk_append(pxo_chunk *c, const uint8_t *enc, int enc_len);
int pxo_chunk_count(const pxo_chunk *c);
/* Decoded[from:to) of record idx -> out; returns bytes written */
int pxo_chunk_decode(const pxo_chunk *c, int idx, int from, int to, uint8_t *out);
int pxo_chunk_declen(const pxo_chunk *c, int idx);

/* ---- CritBit index (data_struct/CritBitTree.cpp:13-269) ----
 * Keys are escaped keys with the 251,0 terminator; leaves carry an opaque id. */
typedef struct pxo_cbt pxo_cbt;
pxo_cbt *pxo_cbt_new(void);
void pxo_cbt_free(pxo_cbt *t);
/* returns the leaf id that was replaced (>=0) or -1 when the key was new */
long long pxo_cbt_set(pxo_cbt *t, const uint8_t *qkey, int qlen, long long leaf_id);
long long pxo_cbt_get(const pxo_cbt *t, const uint8_t *qkey, int qlen); /* id or -1 */
long long pxo_cbt_del(pxo_cbt *t, const uint8_t *qkey, int qlen);       /* id or -1 */
/* ids of all leaves whose key starts with the escaped prefix (no terminator), key order */
long long pxo_cbt_iter(const pxo_cbt *t, const uint8_t *prefix, int plen, long long *ids, long long cap);
long long pxo_cbt_size(const pxo_cbt *t);
/* number of inner nodes visited by find_best_match (CritBitTree.cpp:253-269) */
long long pxo_cbt_depth(const pxo_cbt *t, const uint8_t *qkey, int qlen);

#ifdef __cplusplus
}
#endif
#endif
