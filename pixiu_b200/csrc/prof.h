// Optional per-kernel-class timing with CUDA events on the store's stream (bench.py's roofline leg).
// Disabled by default: begin/end are no-ops and add no events to the stream.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <vector>

namespace pixiu {

enum ProfClass {
    PC_DOCS = 0,     // k_doc_len, k_write_docs
    PC_INIT_KEYS,    // k_init_keys
    PC_SORT_HIST,    // k_rs_histogram + k_rs_scan_bins
    PC_SORT_PASS,    // k_rs_onesweep
    PC_RANK_SCAN,    // head/rank scan + compaction scan of a doubling round
    PC_ROUND_KEYS,   // k_round_keys / k_round_key2
    PC_SEG_SORT,     // k_group_sort_small / k_group_sort_large
    PC_LCP,          // k_lcp
    PC_TREE,         // k_tree_level
    PC_LPF,          // k_lpf
    PC_NODES,        // k_nodes (reference rotation rule)
    PC_FLAGS,        // k_flag_scatter, k_pair_rule, run/offset scans
    PC_EMIT,         // k_emit
    PC_TABLES,       // k_record_tables, k_tile_desc
    PC_DECODE,       // the decode kernels together (k_decode_literals, k_decode_copies, k_fin_clean [, k_copy_records])
    PC_DECODE_LIT,   // k_decode_literals alone
    PC_DECODE_COPY,  // k_decode_copies alone
    PC_LOOKUP,       // k_query_*, k_lookup
    PC_COUNT
};

inline const char *prof_class_name(int c) {
    static const char *names[PC_COUNT] = {"docs", "init_keys", "sort_hist", "sort_pass", "rank_scan", "round_keys", "seg_sort", "lcp",
                                          "tree", "lpf", "nodes", "flags", "emit", "tables", "decode", "decode_lit", "decode_copy", "lookup"};
    return c >= 0 && c < PC_COUNT ? names[c] : "?";
}

struct Profiler {
    bool on = false;
    struct Span {
        cudaEvent_t a, b;
        int cls;
        double bytes;
        int launches;
        bool open;
    };
    std::vector<Span> spans;
    std::vector<cudaEvent_t> pool;
    double ms[PC_COUNT] = {0}, bytes[PC_COUNT] = {0};
    int64_t launches[PC_COUNT] = {0};

    cudaEvent_t get() {
        if (!pool.empty()) {
            cudaEvent_t e = pool.back();
            pool.pop_back();
            return e;
        }
        cudaEvent_t e;
        cudaEventCreate(&e);
        return e;
    }
    void begin(int cls, cudaStream_t st) {
        if (!on) return;
        Span s{get(), get(), cls, 0, 0, true};
        cudaEventRecord(s.a, st);
        spans.push_back(s);
    }
    void end(cudaStream_t st, double nbytes, int nlaunch) {
        if (!on) return;
        // (spans nest: an end closes the innermost span that is still open)
        int k = (int) spans.size() - 1;
        while (k >= 0 && !spans[k].open) k--;
        if (k < 0) return;
        Span &s = spans[k];
        s.open = false;
        s.bytes = nbytes;
        s.launches = nlaunch;
        cudaEventRecord(s.b, st);
    }
    // call after the stream has been synchronised
    void collect() {
        for (Span &s : spans) {
            float t = 0;
            if (cudaEventElapsedTime(&t, s.a, s.b) == cudaSuccess) {
                ms[s.cls] += t;
                bytes[s.cls] += s.bytes;
                launches[s.cls] += s.launches;
            }
            pool.push_back(s.a);
            pool.push_back(s.b);
        }
        spans.clear();
    }
    void reset() {
        collect();
        for (int i = 0; i < PC_COUNT; i++) ms[i] = bytes[i] = 0, launches[i] = 0;
    }
    ~Profiler() {
        collect();
        for (cudaEvent_t e : pool) cudaEventDestroy(e);
    }
};

}  // namespace pixiu
