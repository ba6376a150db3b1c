// Internal definition of the store behind the C ABI (include/pixiu_b200.h).
//
// HBM layout
//   compressed store (all chunks):  enc u8[]          encoded records back to back
//                                   rec_enc_off u64[] start of record g in enc
//                                   rec_enc_len u32[] / rec_dec_len u32[]
//                                   rec_first   u32[] global id of the first record of g's chunk
//                                                     (back references carry chunk-local idx)
//                                   rec_tile_base u32[] first decode tile of record g
//                                   tile_desc u32[]   per 2 KiB of decoded bytes: enc offset (lo16)
//                                                     of the token holding the tile's first byte and
//                                                     the bytes of that token to skip (hi16)
//   open window (= current chunk):  text u8[N]  escaped docs, one 0 separator after each
//                                   dist u16[N] bytes to the end of the own record (0 at separator)
//                                   recid u16[N] chunk-local record index
//                                   rec_start u32[R+1]
//   index mirror:                   CritBit nodes as SoA (child0/child1 i32, diff_at u16, mask u8)
#pragma once
#include <algorithm>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/pixiu_b200.h"
#include "common.cuh"
#include "prof.h"
#include "radix_sort.cuh"
#include "scan.cuh"

namespace pixiu {

constexpr uint32_t TILE = 2048;          // decoded bytes per decode tile
constexpr uint32_t MAX_DOC = 65535;      // PXSG_MAX_TO (proj/PiXiuStr.h:21)
constexpr uint32_t MAX_CHUNK_RECS = 65535;  // PXC_STR_NUM (proj/PiXiuStr.h:20)
constexpr int TREE_B = 16;               // branching of the block-min trees
constexpr int TREE_MAX_LEVELS = 9;

struct MinTree {
    const uint32_t *a[TREE_MAX_LEVELS];   // block minima of the suffix array (level 0 = sa)
    const uint32_t *l[TREE_MAX_LEVELS];   // block minima of the LCP array   (level 0 = lcp)
    uint32_t size[TREE_MAX_LEVELS];
    int nlev;
    // level 0 interleaved, {sa[j], lcp[j] | dist[sa[j]] << 16} in one 8-byte entry: a neighbour probe of the match finder
    // touches ONE 32-byte sector where the two arrays cost two, and "does that occurrence continue past the match"
    // (the node rule of the rotation law) needs no further lookup (nullptr: level 0 is read from a[0] / l[0])
    const uint2 *leaf;
};

class HostIndex;  // CritBit (index.cu)
struct MgComm;    // NCCL communicator + collective statistics of a multi-GPU shard (mgcomm.cu)
void mg_comm_free(MgComm *c);

struct EncodeScratch {
    DevBuf<uint64_t> keys0, keys1;
    DevBuf<uint32_t> vals0, vals1, slot0, slot1, gk, sa, rank, lcp, reach, lastnon, prevp, nextp, off;
    DevBuf<uint64_t> qoff;
    DevBuf<uint2> leaf;   // {sa, lcp} per suffix-array slot (MinTree::leaf)
    DevBuf<uint32_t> goff, glarge;  // segmented group sort: group offsets, list of big groups
    // groups no CTA can sort (> GS_MAX members): their ids, their offsets in the compact buffers, and the compact
    // (key, value) pairs a radix sort orders while every other group is sorted in place
    DevBuf<uint32_t> gmedium, ghuge, hoff, hv0, hv1;   // (gmedium: groups of 33 .. 128 members, a warp each)
    DevBuf<uint64_t> hk1;
    ScanWorkspace scanws;
    DevBuf<uint32_t> tree_a, tree_l;
    DevBuf<uint8_t> flagp, flagc, symmap;
    DevBuf<uint32_t> leafmask, splitmask, wordpre, longmap;
    PinnedBuf<uint32_t> h_leafmask, h_splitmask, h_wordpre;
    PinnedBuf<uint32_t> h_round;   // [0..7] counters of a suffix-array round, [8] sequence number (polled by the host)
    DevBuf<uint32_t> counters;  // [0] active count, [1] group count, [2] error flag, [3..] misc
    RadixSortTemp rs;
    MinTree tree{};  // block-min trees of the last phase A
    // multi-GPU exchange buffers
    DevBuf<uint32_t> mg_m, mg_cand, runidx;
    DevBuf<uint16_t> gidx;
};

// Tuning / test knobs: read from the environment ONCE, when the store is created (never on a hot path); tests change
// them afterwards through pixiu_debug_set_knob.
struct Knobs {
    bool trace = false;            // PIXIU_TRACE: phase trace lines on stderr
    bool lcp_kasai = false;        // PIXIU_LCP_KASAI: the one-pass Kasai walk (A/B measurements)
    bool no_spec_emit = false;     // PIXIU_NO_SPEC_EMIT: emit after the rotation cut instead of beside it (A/B)
    bool no_segsort = false;       // PIXIU_NO_SEGSORT: radix sort only (A/B)
    uint32_t lastnon_mode = 0;     // tests: 0 the batch decides (k_doc_len), 1 always run the last-non-251 scan, 2 never (only
                                   // valid when no run of 251s is longer than the walk-back of the pair rule may go)
    uint64_t dec_arena_limit = 7ull << 29;  // PIXIU_DEC_ARENA_LIMIT: decoded bytes of one decode pass (3.5 GiB)
    uint32_t piece_cap = 0xFFFFFFFFu;       // PIXIU_PIECE_CAP: pending pieces a decode tile keeps before it drains them
    uint32_t sleep_after = 16, sleep_ns = 64;  // PIXIU_SLEEP_AFTER / _NS: back-off of the decoder's polls
    uint32_t bulk_min = 65536;     // PIXIU_BULK_MIN: smallest batch the index rebuilds on the GPU instead of splicing on the host
    uint32_t copy_ctas = 8;        // PIXIU_COPY_CTAS: resident CTAs per SM of the copy kernel (persistent warps)
    uint32_t sweep_gap = 0;        // PIXIU_SWEEP_GAP: ns a decode warp sleeps between two sweeps over its open pieces
    std::string dec_trace_file;    // PIXIU_DEC_TRACE_FILE: per-tile timestamps of a decode call
    void from_env() {
        auto num = [](const char *n, uint64_t dflt) -> uint64_t {
            const char *v = getenv(n);
            return v ? (uint64_t) atoll(v) : dflt;
        };
        trace = getenv("PIXIU_TRACE") != nullptr;
        lcp_kasai = getenv("PIXIU_LCP_KASAI") != nullptr;
        no_spec_emit = getenv("PIXIU_NO_SPEC_EMIT") != nullptr;
        no_segsort = getenv("PIXIU_NO_SEGSORT") != nullptr;
        dec_arena_limit = num("PIXIU_DEC_ARENA_LIMIT", dec_arena_limit);
        piece_cap = (uint32_t) num("PIXIU_PIECE_CAP", piece_cap);
        sleep_after = (uint32_t) num("PIXIU_SLEEP_AFTER", sleep_after);
        sleep_ns = (uint32_t) num("PIXIU_SLEEP_NS", sleep_ns);
        sweep_gap = (uint32_t) num("PIXIU_SWEEP_GAP", sweep_gap);
        copy_ctas = (uint32_t) num("PIXIU_COPY_CTAS", copy_ctas);
        bulk_min = (uint32_t) num("PIXIU_BULK_MIN", bulk_min);
        if (const char *f = getenv("PIXIU_DEC_TRACE_FILE")) dec_trace_file = f;
    }
    bool set(const std::string &name, int64_t v) {
        if (name == "dec_arena_limit") dec_arena_limit = (uint64_t) v;
        else if (name == "piece_cap") piece_cap = (uint32_t) v;
        else if (name == "sleep_after") sleep_after = (uint32_t) v;
        else if (name == "sleep_ns") sleep_ns = (uint32_t) v;
        else if (name == "sweep_gap") sweep_gap = (uint32_t) v;
        else if (name == "copy_ctas") copy_ctas = (uint32_t) v;
        else if (name == "bulk_min") bulk_min = (uint32_t) v;
        else if (name == "trace") trace = v != 0;
        else if (name == "lcp_kasai") lcp_kasai = v != 0;
        else if (name == "no_spec_emit") no_spec_emit = v != 0;
        else if (name == "no_segsort") no_segsort = v != 0;
        else if (name == "lastnon_mode") lastnon_mode = (uint32_t) v;
        else return false;
        return true;
    }
};

struct Store {
    pixiu_config cfg{};
    Knobs knobs;
    cudaStream_t st = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_nodes = nullptr;
    std::string err;
    // A mutating call that fails after its first change (CUDA error, sticky scan time-out, late argument error) leaves
    // records that are not indexed, or a half-built window: `dirty` is raised at the first change of such a call and
    // lowered when it completes; a call that ends with `dirty` still raised poisons the store, and every later entry
    // point answers PIXIU_EPOISONED instead of running on that state.
    bool dirty = false, poisoned = false;
    int64_t launches = 0;
    Profiler prof;
    double last_set_ms = 0, last_get_ms = 0, last_lookup_ms = 0;
    uint64_t last_lookup_qbytes = 0;  // escaped query bytes of the last lookup

    // ---- record tables (host mirrors) ----
    std::vector<uint64_t> h_enc_off;
    std::vector<uint32_t> h_enc_len, h_dec_len, h_first, h_tile_base;
    std::vector<uint8_t> h_live;
    std::vector<uint32_t> chunk_first;  // global id of the first record of chunk c
    std::vector<uint32_t> chunk_count;  // records in chunk c
    uint64_t enc_bytes = 0, n_tiles = 0;
    int64_t raw_bytes = 0, doc_bytes = 0, live_records = 0;
    // occupancy per chunk and the reference's compaction trigger (PiXiuStr.cpp:178-187, PiXiuCtrl.cpp:7-8,:26-29)
    std::vector<uint32_t> chunk_live;    // live (not tombstoned) records of chunk c
    std::vector<uint8_t> chunk_dropped;  // chunk c was re-inserted: none of its records is live or reachable
    int64_t reinsert_candidate = -1;     // Glob_Reinsert_Chunk
    int64_t reinserted_records = 0, reclaimable_bytes = 0;
    size_t chunk_of(uint32_t g) const {
        return (size_t) (std::upper_bound(chunk_first.begin(), chunk_first.end(), g) - chunk_first.begin()) - 1;
    }
    void note_live(uint32_t g) {
        const size_t c = chunk_of(g);
        if (chunk_live.size() <= c) chunk_live.resize(c + 1, 0);
        chunk_live[c]++;
        live_records++;
    }
    void tombstone(uint32_t g) {  // PiXiuChunk::delitem (PiXiuStr.cpp:178-187): the bytes stay, later records may reference them
        h_live[g] = 0;
        live_records--;
        const size_t c = chunk_of(g);
        if (chunk_live.size() <= c) chunk_live.resize(c + 1, 0);
        chunk_live[c]--;
        if (chunk_live[c] < 0.8 * MAX_CHUNK_RECS) reinsert_candidate = (int64_t) c;
    }

    // ---- device store ----
    VmArena d_enc;  // compressed arena: 256 GiB of address space, physical memory mapped as it fills
    DevBuf<uint64_t> d_enc_off;
    DevBuf<uint32_t> d_enc_len, d_dec_len, d_first, d_tile_base, d_tile_desc;

    // ---- open window ----
    bool win_open = false;
    uint32_t win_R = 0, win_N = 0;
    DevBuf<uint8_t> w_text;
    DevBuf<uint16_t> w_dist, w_recid;
    DevBuf<uint32_t> w_rec_start;
    std::vector<uint32_t> h_win_rec_start;  // R+1
    // arena state of the reference's suffix tree for the open window (MemPool::nth / used_num)
    uint32_t pool_nth = 1, pool_used = 5;
    double rho = 1.35;  // running estimate of suffix-tree nodes per window byte
    double rho_err = 0.02;  // recent relative error of that estimate
    uint32_t win_present[8] = {0}, batch_present[9] = {0};  // byte values present in the open window / last batch ([8]: long 251 run)
    bool win_long251 = false;   // the open window may hold a run of more than 65 bytes 251 (then the flag phase scans for it)

    EncodeScratch es;
    std::unique_ptr<HostIndex> index;

    // ---- decode scratch ----
    DevBuf<uint8_t> dec_scratch;   // decoded arena when the caller's layout differs from arena order
    DevBuf<uint64_t> dec_loc;      // output offsets of the requested records
    DevBuf<uint32_t> dec_flags;    // zero-byte bitmap of the arena (1 bit per byte; all-zero between calls)
    DevBuf<uint32_t> dec_dirty;    // words of that bitmap the running call has set bits in
    DevBuf<uint64_t> dec_pieces;   // packed table of copy pieces {source, meta} written by K10, read by K11
    DevBuf<uint64_t> dec_phead;    // per tile: {first piece, pieces}
    uint32_t dec_sms = 0;          // SM count of the store's device (grid of the persistent decode kernel)
    uint64_t last_pending_pieces = 0, last_drains = 0;  // copy pieces of the last decode (and spare counter)
    DevBuf<uint32_t> dec_aoff;     // per record: offset in the arena
    DevBuf<uint32_t> dec_reqs;     // requested record ids
    DevBuf<uint32_t> dec_work;     // work list: tile ids, record ids, range ids (built by k_dec_work)
    DevBuf<uint32_t> dec_ranges;   // touched chunk ranges of the call (DecRange[], decode.cu)
    std::vector<uint64_t> h_dec_prefix;  // running sum of h_dec_len (NR + 1 entries, extended lazily)
    DevBuf<uint64_t> d_dec_prefix;
    size_t dec_prefix_synced = 0;
    DevBuf<uint32_t> dec_ctr;      // counters of a decode call (enum DC_* in decode.cu)

    // ---- staging for batches ----
    DevBuf<uint8_t> in_keys, in_vals, out_stage;
    DevBuf<int64_t> in_koff, in_voff;
    // the multi-GPU phases keep their batch across calls: own staging, never touched by lookups (which reuse in_*)
    DevBuf<uint8_t> mg_in_keys, mg_in_vals;
    DevBuf<int64_t> mg_in_koff, mg_in_voff;
    DevBuf<uint32_t> doc_len, doc_off;
    DevBuf<uint32_t> iter_buf;     // frontier ping-pong buffers + output of the device-side prefix walk

    Store() = default;
    ~Store();

    size_t n_records() const { return h_enc_len.size(); }
    size_t n_chunks() const { return chunk_first.size(); }

    void init(const pixiu_config &c);
    void grow_record_tables(size_t n_total, uint64_t enc_total, uint64_t tiles_total);
    void open_window();
    void close_window();

    // encode.cu
    int setitem_batch(int64_t n, const uint8_t *d_keys, const int64_t *d_koff, const uint8_t *d_vals,
                      const int64_t *d_voff, const uint8_t *h_keys, const int64_t *h_koff,
                      const int64_t *h_voff, int32_t *rc, int32_t *saved);
    uint32_t encode_window_records(uint32_t first_new);  // returns the number of records accepted
    void enc_phase_a(uint32_t first_new, bool fuse_flags = false);   // fuse_flags: k_lpf also scatters the PASS flags
    bool ep_flags_done = false;
    uint32_t round_seq = 0;   // hand-overs of round counters so far (k_publish_counts)
    void enc_phase_b();
    uint32_t enc_phase_c(const uint32_t *cand, const uint32_t *runidx, const uint16_t *gidx);
    uint32_t ep_first_new = 0, ep_s0 = 0, ep_N = 0, ep_n_new = 0;  // state shared by the phases
    bool ep_emitted = false;  // k_emit already ran for the candidates (enc_emit_all)
    void enc_emit_all();
    void finish_index(uint32_t nn, size_t g_batch_first, const uint8_t *d_keys, const int64_t *d_koff, const uint8_t *h_keys,
                      const int64_t *h_koff, const int64_t *h_voff,
                      const uint32_t *h_doc_len, int32_t *rc, int32_t *saved);
    // multi-GPU extended window (encode.cu, "Multi-GPU extended window")
    int mg_rank = 0, mg_world = 0, mg_pending = 0;
    MgComm *mg_comm = nullptr;   // set by pixiu_mg_comm_init: the collectives run inside the library (mgcomm.cu)
    uint32_t mg_gR = 0;          // records of the open chunk over all ranks
    uint64_t mg_gbytes = 0, mg_batch_bytes = 0;
    std::vector<uint16_t> mg_h_gidx;  // chunk index of every local window record
    std::vector<uint32_t> mg_doc_len;
    std::vector<uint8_t> mg_h_keys;
    std::vector<int64_t> mg_h_koff, mg_h_voff;
    const uint8_t *mg_d_keys = nullptr, *mg_d_vals = nullptr;
    const int64_t *mg_d_koff = nullptr, *mg_d_voff = nullptr;
    int mg_begin(int64_t n, const uint8_t *d_keys, const int64_t *d_koff, const uint8_t *d_vals, const int64_t *d_voff,
                 const uint8_t *h_keys, const int64_t *h_koff, const int64_t *h_voff, uint32_t **d_m, int64_t *count,
                 bool sync = true);
    // (sync: the caller issues the collective from another stream / library, so the phase ends synchronised)
    int mg_mid(uint32_t **d_cand, int64_t *count, bool sync = true);
    int mg_end(int32_t *rc, int32_t *saved);
    void count_nodes_enqueue(uint32_t s0, uint32_t N);
    uint32_t count_nodes_and_cut(uint32_t first_new, uint32_t s0, uint32_t N);
    void apply_rotation_cut();
    void flush_mirrors();
    size_t mirror_from = 0;  // first record whose host mirrors are not fetched yet
    // decode.cu
    void decode_records(const std::vector<uint32_t> &recs, uint8_t *d_out, const std::vector<uint64_t> &out_off);
    void decode_pass(const std::vector<uint32_t> &recs, uint8_t *d_out, const std::vector<uint64_t> &out_off,
                     const std::map<uint32_t, uint32_t> *known_max);
    int64_t import_chunk(int64_t n, const uint8_t *enc, const int64_t *enc_off);
};

// mgcomm.cu
int mg_unique_id(uint8_t *id, std::string &err);
int mg_comm_init(Store &S, int rank, int world, const uint8_t *id);
int mg_setitem_nccl(Store &S, int64_t n, const uint8_t *d_keys, const int64_t *d_koff, const uint8_t *d_vals,
                    const int64_t *d_voff, const uint8_t *h_keys, const int64_t *h_koff, const int64_t *h_voff,
                    int32_t *rc, int32_t *saved);
void mg_comm_stats(const Store &S, pixiu_mg_stats *o);

}  // namespace pixiu
