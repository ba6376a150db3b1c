/* pixiu_b200 — C ABI of the B200-native PiXiu hot path.
 *
 * Drop-in boundary for the reference's `struct PiXiuCtrl`
 * (/root/reference/src/proj/PiXiuCtrl.h:7-26; public API README.md:104-119).  The
 * reference has no FFI of its own: its only API is that C++ struct, whose methods
 * these entry points replace one for one, batched:
 *
 *   PiXiuCtrl::init_prop  (PiXiuCtrl.cpp:77-81)  -> pixiu_create
 *   PiXiuCtrl::free_prop  (PiXiuCtrl.cpp:83-86)  -> pixiu_destroy
 *   PiXiuCtrl::setitem    (PiXiuCtrl.cpp:12-47)  -> pixiu_setitem_batch   (in-order semantics of n calls)
 *   PiXiuCtrl::contains   (PiXiuCtrl.cpp:55-57)  -> pixiu_contains_batch
 *   PiXiuCtrl::getitem    (PiXiuCtrl.cpp:59-61)  -> pixiu_getitem_batch   (drained PXSGen streams)
 *   PiXiuCtrl::delitem    (PiXiuCtrl.cpp:63-69)  -> pixiu_delitem_batch
 *   PiXiuCtrl::iter       (PiXiuCtrl.cpp:71-75)  -> pixiu_iter
 *   PiXiuCtrl::reinsert   (PiXiuCtrl.cpp:88-114) -> pixiu_reinsert_chunk (+ pixiu_config.auto_reinsert for the trigger)
 *   ctrl.st.cbt_chunk->getitem(i)->{len,data} (main.cpp:67) -> pixiu_encoded_view
 *
 * Conventions: plain pointers and sizes; inputs are borrowed; outputs go to caller
 * buffers (query the size first by passing cap = 0); return 0 on success, a negative
 * PIXIU_E* code otherwise (the reference only asserts — oversize records are an explicit
 * error here).  Batches are *packed*: `data` holds the items back to back and
 * `off[n+1]` (int64) their boundaries.  A store may be used by one host thread at a
 * time; several stores per process are fine (no globals).  All compute runs on the
 * GPU chosen at create time; there is no CPU fallback.
 */
#ifndef PIXIU_B200_H
#define PIXIU_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pixiu_store pixiu_store;

#define PIXIU_OK 0
#define PIXIU_EINVAL (-1)      /* bad argument */
#define PIXIU_ETOOLONG (-2)    /* record longer than 65,535 escaped bytes (PiXiuStr.cpp:123,:238) */
#define PIXIU_ECUDA (-3)       /* CUDA failure; see pixiu_last_error */
#define PIXIU_ENOSPC (-4)      /* caller buffer too small; required size reported */
#define PIXIU_ECORRUPT (-5)    /* malformed encoded record */
#define PIXIU_EINTERNAL (-6)
#define PIXIU_EPOISONED (-7)   /* an earlier update failed half way (CUDA error, ...): the store refuses further calls; destroy it */

#define PIXIU_CBT_SET_REPLACE 1   /* data_struct/CritBitTree.h:7 */
#define PIXIU_CBT_DEL_NOT_FOUND 1 /* data_struct/CritBitTree.h:8 */

/* window (= chunk) rotation policy, PiXiuCtrl.cpp:13-17 */
#define PIXIU_ROTATE_REFERENCE 0 /* the reference's rule: arena pools >= 2048 or 65,535 records */
#define PIXIU_ROTATE_BYTES 1     /* start a new chunk when the window text would exceed window_bytes */
#define PIXIU_ROTATE_RECORDS 2   /* only the format limit of 65,535 records per chunk */

typedef struct pixiu_config {
    int32_t device;          /* CUDA device ordinal */
    int32_t rotate_policy;   /* PIXIU_ROTATE_* */
    int64_t window_bytes;    /* for PIXIU_ROTATE_BYTES */
    int32_t strict251;       /* 1: reproduce reference bug B1 (run of 251 -> ambiguous FB FB form) */
    int32_t auto_reinsert;   /* 1: the reference's compaction trigger (PiXiuCtrl.cpp:7-8,:26-29,:64-67): a closed chunk
                                with < 50 % live records is re-inserted before the next setitem batch / delitem */
} pixiu_config;

typedef struct pixiu_stats {
    int64_t records;        /* records ever stored (tombstoned included) */
    int64_t live_records;
    int64_t chunks;
    int64_t raw_bytes;      /* sum of key+value bytes stored */
    int64_t doc_bytes;      /* sum of escaped doc lengths */
    int64_t encoded_bytes;  /* sum of encoded record lengths */
    int64_t window_bytes;   /* text currently resident in the open window */
    int64_t kernel_launches;/* kernels launched by this store so far */
    double last_setitem_gpu_ms; /* device time of the last setitem batch (CUDA events) */
    double last_getitem_gpu_ms;
    double last_lookup_gpu_ms;
    int64_t reinserted_records; /* records moved by pixiu_reinsert_chunk / the auto_reinsert trigger so far */
    int64_t reclaimable_bytes;  /* encoded bytes of dropped chunks (still resident; freed by export + import) */
    /* memory the CritBit index holds BESIDE the compressed store (the reference verifies a key by decoding the record,
     * CritBitTree.cpp:154-178; this index keeps esc(k) 251 0 of every leaf instead - these bytes are the price) */
    int64_t index_key_arena_bytes; /* escaped keys of the leaves (one copy on the host, one in HBM) */
    int64_t index_host_bytes;      /* host arrays: nodes, leaves, key arena */
    int64_t index_device_bytes;    /* their device mirror (allocated capacity) */
    int64_t table_device_bytes;    /* record tables + decode tile descriptors in HBM (allocated capacity) */
} pixiu_stats;

void pixiu_default_config(pixiu_config *cfg);
pixiu_store *pixiu_create(const pixiu_config *cfg);   /* NULL on failure */
void pixiu_destroy(pixiu_store *s);
const char *pixiu_last_error(const pixiu_store *s);
int pixiu_get_stats(pixiu_store *s, pixiu_stats *out);

/* n sequential setitem calls. val_off[i]==val_off[i+1] stores a key-only record
 * (PiXiuCtrl.cpp:41-44).  rc[i] = 0 or PIXIU_CBT_SET_REPLACE; saved[i] = doc_len - encoded_len
 * (main.cpp:67); either may be NULL.  Host pointers. */
int pixiu_setitem_batch(pixiu_store *s, int64_t n, const uint8_t *keys, const int64_t *key_off,
                        const uint8_t *vals, const int64_t *val_off, int32_t *rc, int32_t *saved);
/* same with DEVICE pointers for keys/key_off/vals/val_off (inputs already in HBM) */
int pixiu_setitem_batch_dev(pixiu_store *s, int64_t n, const uint8_t *d_keys, const int64_t *d_key_off,
                            const uint8_t *d_vals, const int64_t *d_val_off, int32_t *rc, int32_t *saved);

int pixiu_contains_batch(pixiu_store *s, int64_t n, const uint8_t *keys, const int64_t *key_off, uint8_t *found);
/* same with DEVICE pointers: keys/key_off (key_off[0] == 0) already in HBM, found[] written on the device */
int pixiu_contains_batch_dev(pixiu_store *s, int64_t n, const uint8_t *d_keys, const int64_t *d_key_off, uint8_t *d_found);
int pixiu_delitem_batch(pixiu_store *s, int64_t n, const uint8_t *keys, const int64_t *key_off, int32_t *rc);

/* Decoded records `esc(k) 251 0 [esc(v) 251 2]` (what draining PXSGen yields, README.md:157)
 * packed into out[0..out_off[n]); absent keys get an empty slot and found[i]=0.
 * Returns PIXIU_ENOSPC with *need set when out_cap is too small (pass out_cap=0 to query). */
int pixiu_getitem_batch(pixiu_store *s, int64_t n, const uint8_t *keys, const int64_t *key_off,
                        uint8_t *out, int64_t out_cap, int64_t *out_off, uint8_t *found, int64_t *need);
/* same, output left in device memory (d_out: device pointer, out_off/found: host) */
int pixiu_getitem_batch_dev(pixiu_store *s, int64_t n, const uint8_t *keys, const int64_t *key_off,
                            uint8_t *d_out, int64_t out_cap, int64_t *out_off, uint8_t *found, int64_t *need);

/* iter(prefix): decoded records of all live keys starting with `prefix`, ascending key
 * order (CritBitTree.h:55-157). *count = number of records. */
int pixiu_iter(pixiu_store *s, const uint8_t *prefix, int64_t prefix_len, uint8_t *out, int64_t out_cap,
               int64_t *out_off, int64_t off_cap, int64_t *count, int64_t *need);

/* encoded bytes of record (chunk, idx) copied to out; returns length or negative error */
int pixiu_encoded_view(pixiu_store *s, int64_t chunk, int64_t idx, uint8_t *out, int64_t out_cap);
/* (chunk, idx) of the i-th record ever inserted */
int pixiu_record_location(pixiu_store *s, int64_t record, int64_t *chunk, int64_t *idx);

/* PiXiuCtrl::reinsert (PiXiuCtrl.cpp:88-114): decode the live records of a CLOSED chunk, insert them again through
 * the setitem path (they compress against the open window and replace themselves in the index) and drop the chunk.
 * Works for chunks of any size (the reference dereferences NULL slots of non-full chunks, bug B4).  Returns the
 * number of records moved or a negative error. */
int64_t pixiu_reinsert_chunk(pixiu_store *s, int64_t chunk);
/* records ever stored in the chunk, records still live, and whether it was dropped by a re-insertion */
int pixiu_chunk_info(pixiu_store *s, int64_t chunk, int64_t *total, int64_t *live, int32_t *dropped);

/* Import one pre-encoded chunk (array of PiXiu-encoded records, e.g. produced by the
 * reference) as a new closed chunk and index its keys; used by parity tests and as the
 * load path of a serialised store.  Returns the chunk id or a negative error. */
int64_t pixiu_import_chunk(pixiu_store *s, int64_t n, const uint8_t *enc, const int64_t *enc_off);

/* Export one chunk in its wire format — the PiXiu-encoded records back to back plus offsets[count+1] — the
 * inverse of pixiu_import_chunk (a chunk is self-contained: back references never leave it).  Pass out_cap = 0
 * to query *count and *need; out_off must hold *count + 1 entries. */
int pixiu_export_chunk(pixiu_store *s, int64_t chunk, uint8_t *out, int64_t out_cap, int64_t *out_off, int64_t *count,
                       int64_t *need);

/* The CritBit index in its wire format (flat SoA node arrays, leaf -> record ids, escaped keys; compacted, leaves in key
 * order).  Together with pixiu_export_chunk of every chunk this is the whole store: load the chunks with
 * pixiu_import_chunk_raw (no indexing, no decode) in the same order into a fresh store, then pixiu_import_index -
 * lookups work at once, nothing is rebuilt.  Pass out_cap = 0 to query *need. */
int pixiu_export_index(pixiu_store *s, uint8_t *out, int64_t out_cap, int64_t *need);
int64_t pixiu_import_chunk_raw(pixiu_store *s, int64_t n, const uint8_t *enc, const int64_t *enc_off);
int pixiu_import_index(pixiu_store *s, const uint8_t *blob, int64_t size);

/* Decode every record of one chunk (tombstoned included) in idx order. */
int pixiu_decode_chunk(pixiu_store *s, int64_t chunk, uint8_t *out, int64_t out_cap, int64_t *out_off, int64_t *need);

/* ---- multi-GPU extended window (one store per GPU / rank) ----
 * The open window is sharded by record over `world` stores (record idx -> rank idx % world); a batch is
 * given to every rank.  setitem is split in three phases around the two collectives the caller issues
 * (torch.distributed / NCCL all_reduce on the returned DEVICE buffers, in place):
 *     begin -> d_m[count]    uint32 longest earlier match per batch position   -> all_reduce(MAX)
 *     mid   -> d_cand[count] uint32 (idx << 16 | to) leftmost source per run   -> all_reduce(MIN)
 *     end   -> every rank stores the identical encoded batch (compressed store and index are replicated)
 * Requires PIXIU_ROTATE_BYTES (window_bytes is per GPU) or PIXIU_ROTATE_RECORDS; a batch holds <= 65,535 records. */
int pixiu_mg_config(pixiu_store *s, int rank, int world);
int pixiu_mg_setitem_begin(pixiu_store *s, int64_t n, const uint8_t *keys, const int64_t *key_off, const uint8_t *vals,
                           const int64_t *val_off, uint32_t **d_m, int64_t *count);
int pixiu_mg_setitem_mid(pixiu_store *s, uint32_t **d_cand, int64_t *count);
int pixiu_mg_setitem_end(pixiu_store *s, int32_t *rc, int32_t *saved);

/* The same three phases with the two collectives issued INSIDE the library: ncclAllReduce(MAX) / ncclAllReduce(MIN) on
 * the store's own stream, no host synchronisation between a phase and its collective, no PyTorch.  NCCL is bound at
 * run time (libnccl.so.2, or the path in PIXIU_NCCL_LIB; a copy already loaded by the process is shared).
 *   rank 0:      pixiu_mg_unique_id(s, id)  -> broadcast the 128 bytes to the other ranks by any means
 *   every rank:  pixiu_mg_comm_init(s, rank, world, id)   (collective: ncclCommInitRank; implies pixiu_mg_config)
 *   every rank:  pixiu_mg_setitem_batch(s, <the same batch>)   (collective)
 * Replaces the single window of SuffixTree::setitem (SuffixTree.cpp:291-304) for BASELINE config 5. */
#define PIXIU_NCCL_UNIQUE_ID_BYTES 128
int pixiu_mg_unique_id(pixiu_store *s, uint8_t *id /* [PIXIU_NCCL_UNIQUE_ID_BYTES] */);
int pixiu_mg_comm_init(pixiu_store *s, int rank, int world, const uint8_t *id);
int pixiu_mg_setitem_batch(pixiu_store *s, int64_t n, const uint8_t *keys, const int64_t *key_off, const uint8_t *vals,
                           const int64_t *val_off, int32_t *rc, int32_t *saved);
typedef struct pixiu_mg_stats {
    int32_t rank, world;
    int32_t nccl_version;       /* ncclGetVersion, 0 when the collectives are played by the caller */
    int32_t pad;
    int64_t batches;            /* pixiu_mg_setitem_batch calls so far */
    int64_t max_reduce_bytes;   /* payload of the MAX all-reduces so far (4 B per batch position) */
    int64_t min_reduce_bytes;   /* payload of the MIN all-reduces so far (4 B per long run) */
    double max_reduce_ms;       /* their time on the store's stream (CUDA events; includes waiting for the slowest rank) */
    double min_reduce_ms;
} pixiu_mg_stats;
int pixiu_mg_get_stats(pixiu_store *s, pixiu_mg_stats *out);

/* Per-kernel-class device timing (CUDA events on the store's stream), for the roofline report.
 * enable(1) resets the counters; get() returns 1 once cls is past the last class. `bytes` are the
 * ALGORITHMIC bytes of the launches (DESIGN.md states the per-unit figures). */
int pixiu_profile_enable(pixiu_store *s, int on);
int pixiu_profile_get(pixiu_store *s, int cls, const char **name, double *ms, double *bytes, int64_t *launches);
/* the CUDA stream (cudaStream_t) all of this store's work is issued on */
void *pixiu_stream(pixiu_store *s);

/* force the open window to close (next setitem starts a new chunk) */
int pixiu_rotate(pixiu_store *s);

/* Capacity hint (the std::vector::reserve of the store; the reference's MemPool mallocs 64 KiB blocks as it goes,
 * common/MemPool.cpp): make room NOW for `encoded_bytes` more bytes of compressed records.  The compressed arena is a
 * reserved virtual range; physical HBM is mapped behind it in 256 MiB steps when a batch needs it, and on a freshly
 * booted GPU a single cuMemCreate / cuMemSetAccess was measured at up to 75 ms (memory not scrubbed yet) - an ingest
 * that knows its size pays that here instead of inside a batch.  Never needed for correctness. */
int pixiu_reserve(pixiu_store *s, int64_t encoded_bytes);

#ifdef __cplusplus
}
#endif
#endif
