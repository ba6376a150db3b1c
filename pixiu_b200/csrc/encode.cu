// setitem hot path: batched compression of incoming records against the open window.
//
// Replaces, with flat-array kernels (all HBM-bound integer work, no tensor cores):
//   doc assembly + 251 escaping   PiXiuCtrl.cpp:31-44, PiXiuStr.cpp:228-271      -> k_doc_len, k_write_docs
//   online suffix tree            SuffixTree.cpp:144-304 (+ScapegoatTree.h)       -> suffix array by prefix
//                                 doubling (one radix sort of the first 8 symbols, then every round sorts its groups
//                                 where they stand: a warp per 32 consecutive suffixes / per group up to 256 members,
//                                 a CTA up to 2,048, a radix sort of their own for the members of larger groups),
//                                 LCP, block-min trees over {sa, lcp, dist} leaves, longest previous factor
//   stream encoder                PiXiuStr.cpp:16-118                              -> PASS flags (set by the match
//                                 finder), pair rule (inside the run-start scan), run scans, output-size scan, token
//                                 emission (run ends regrouped per CTA)
// The closed form (checked against the reference by the oracle and by
// tests/pipeline_model.py):  M(s) = longest prefix of D[s..] starting earlier in the window,
// reach(s) = s+M(s);  byte i is PASS iff i is a value of reach;  the pointer of a run ending at j
// is the end of the leftmost occurrence of D[s*(j)..j],  s*(j) = min{s : reach(s) > j}.
#include <algorithm>
#include <cmath>
#include <chrono>
#include <cstring>

#include "index.h"
#include "scan.cuh"
#include "store.h"
#include "mintree.cuh"

namespace pixiu {

// ---------------------------------------------------------------------------------
// K1: doc assembly
// ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t warp_sum(uint32_t v) {
#pragma unroll
    for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

// doc_len[r] = |esc(k)| + 2 [+ |esc(v)| + 2]; 0xFFFFFFFF marks an empty key (invalid)
__global__ void __launch_bounds__(256)
k_doc_len(uint32_t n, const uint8_t *__restrict__ keys, const int64_t *__restrict__ koff,
          const uint8_t *__restrict__ vals, const int64_t *__restrict__ voff, uint32_t *__restrict__ doc_len,
          uint32_t *__restrict__ present /* [0..7] 256-bit set of the byte values seen (for the dense symbol map),
                                             [8] != 0: some key or value may hold more than 32 consecutive 251s */) {
    uint32_t r = (blockIdx.x * 256 + threadIdx.x) >> 5;
    if (r >= n) return;
    const int lane = lane_id();
    int64_t k0 = koff[r], k1 = koff[r + 1], v0 = voff[r], v1 = voff[r + 1];
    uint32_t c = 0;
    uint32_t seen[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    // a lane that meets 251 in two consecutive rounds (bytes 32 apart): what every run of more than 32 raw 251s produces
    // in some lane (present[8]; a coincidence only costs the scan)
    bool longrun = false, prev251 = false;
    for (int64_t i = k0 + lane; i < k1; i += 32) {
        const uint8_t b = keys[i];
        c += b == 251;
        longrun |= prev251 && b == 251;
        prev251 = b == 251;
#pragma unroll
        for (int w = 0; w < 8; w++) seen[w] |= (b >> 5) == w ? 1u << (b & 31) : 0u;
    }
    prev251 = false;
    for (int64_t i = v0 + lane; i < v1; i += 32) {
        const uint8_t b = vals[i];
        c += b == 251;
        longrun |= prev251 && b == 251;
        prev251 = b == 251;
#pragma unroll
        for (int w = 0; w < 8; w++) seen[w] |= (b >> 5) == w ? 1u << (b & 31) : 0u;
    }
    longrun = __any_sync(0xffffffffu, longrun) != 0;
    if (present && longrun && lane == 0) atomicOr(&present[8], 1u);
    if (present) {
#pragma unroll
        for (int w = 0; w < 8; w++) {
            uint32_t v = __reduce_or_sync(0xffffffffu, seen[w]);
            if (lane == 0 && v) atomicOr(&present[w], v);
        }
    }
    c = warp_sum(c);
    if (lane == 0) {
        uint64_t len = (uint64_t) (k1 - k0) + c + 2 + (v1 > v0 ? (uint64_t) (v1 - v0) + 2 : 0);
        doc_len[r] = (k1 <= k0) ? 0xFFFFFFFFu : (len > 0xFFFFFFF0ull ? 0xFFFFFFF0u : (uint32_t) len);
    }
}

// CTA-cooperative escaped copy of src[a, b) to text[dst..]: 16 bytes per thread and iteration, the
// positions shifted by the number of 251s before them (block scan).  Returns the new dst.
__device__ __forceinline__ uint32_t block_write_escaped(const uint8_t *__restrict__ src, int64_t a, int64_t b,
                                                        uint8_t *__restrict__ text, uint32_t dst, uint32_t *sm) {
    constexpr int PER = 16;
    for (int64_t base = a; base < b; base += 256 * PER) {
        int64_t i0 = base + (int64_t) threadIdx.x * PER;
        uint8_t v[PER];
        uint32_t cnt = 0;
#pragma unroll
        for (int k = 0; k < PER; k++) {
            v[k] = i0 + k < b ? src[i0 + k] : 0;
            cnt += (i0 + k < b) && v[k] == 251;
        }
        uint32_t total;
        uint32_t pre = block_scan_exclusive<uint32_t>(cnt, OpSum(), 0u, sm, &total);
        uint32_t p = dst + threadIdx.x * PER + pre;
        if (total == 0) {
#pragma unroll
            for (int k = 0; k < PER; k++)
                if (i0 + k < b) text[p + k] = v[k];
        } else {
#pragma unroll
            for (int k = 0; k < PER; k++)
                if (i0 + k < b) {
                    text[p++] = v[k];
                    if (v[k] == 251) text[p++] = 251;
                }
        }
        int64_t chunk = b - base < 256 * PER ? b - base : 256 * PER;
        dst += (uint32_t) chunk + total;
    }
    return dst;
}

// one CTA per new record: text = esc(k) 251 0 [esc(v) 251 2] 0(separator); dist; recid
__global__ void __launch_bounds__(256)
k_write_docs(uint32_t n_new, uint32_t batch_first, uint32_t win_first, const uint8_t *__restrict__ keys,
             const int64_t *__restrict__ koff, const uint8_t *__restrict__ vals, const int64_t *__restrict__ voff,
             const uint32_t *__restrict__ rec_start, uint8_t *__restrict__ text, uint16_t *__restrict__ dist,
             uint16_t *__restrict__ recid, const uint32_t *__restrict__ src_list) {
    __shared__ uint32_t sm[33];
    const uint32_t w = blockIdx.x;
    if (w >= n_new) return;
    const uint32_t src = src_list ? src_list[w] : batch_first + w, idx = win_first + w;
    const uint32_t base = rec_start[idx], len = rec_start[idx + 1] - base - 1;
    uint32_t dst = block_write_escaped(keys, koff[src], koff[src + 1], text, base, sm);
    if (threadIdx.x == 0) {
        text[dst] = 251;
        text[dst + 1] = 0;
    }
    dst += 2;
    int64_t v0 = voff[src], v1 = voff[src + 1];
    if (v1 > v0) {
        dst = block_write_escaped(vals, v0, v1, text, dst, sm);
        if (threadIdx.x == 0) {
            text[dst] = 251;
            text[dst + 1] = 2;
        }
        dst += 2;
    }
    if (threadIdx.x == 0) text[base + len] = 0;
    for (uint32_t p = threadIdx.x; p <= len; p += 256) {
        dist[base + p] = (uint16_t) (len - p);
        recid[base + p] = (uint16_t) idx;
    }
}

// ---------------------------------------------------------------------------------
// K2-K4: suffix array by prefix doubling
// ---------------------------------------------------------------------------------
// initial key = the first `nsym` symbols, `bits` bits each: symbol = 1 + rank of the byte among the byte
// values present in the window (dense, order preserving), 0 from the record end on
__global__ void __launch_bounds__(256)
k_init_keys(const uint8_t *__restrict__ text, const uint16_t *__restrict__ dist, uint32_t n, int nsym, int bits,
            const uint8_t *__restrict__ symmap, uint64_t *__restrict__ keys) {
    __shared__ uint8_t sm[256];
    sm[threadIdx.x] = symmap[threadIdx.x];
    __syncthreads();
    uint32_t i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    uint32_t d = dist[i];
    uint64_t k = 0;
    for (int j = 0; j < nsym; j++) k = (k << bits) | (uint64_t) (j < (int) d ? (uint32_t) sm[text[i + j]] + 1u : 0u);
    keys[i] = k;
}

struct HeadFn {
    const uint64_t *keys;
    uint32_t n;
    uint64_t initial;  // mask of the last symbol slot in the round that follows the initial sort, else 0
    __device__ __forceinline__ bool operator()(uint32_t a) const {
        if (a == 0 || a >= n) return true;
        uint64_t k = keys[a];
        if (k != keys[a - 1]) return true;
        return initial && (k & initial) == 0;  // a suffix that hit its record end is its own group
    }
};

// key'[x] = (group << kb) | rank[val[x] + h]
__global__ void __launch_bounds__(256)
k_round_keys(uint32_t A, const uint32_t *__restrict__ vals, const uint32_t *__restrict__ gk,
             const uint32_t *__restrict__ rank, uint32_t h, int kb, uint64_t *__restrict__ keys) {
    uint32_t x = blockIdx.x * 256 + threadIdx.x;
    if (x >= A) return;
    keys[x] = ((uint64_t) gk[x] << kb) | (uint64_t) rank[vals[x] + h];
}

// ---------------------------------------------------------------------------------
// Segmented sort of the active groups by rank[i+h] (one launch instead of 5-6 radix passes): once no
// group is larger than GS_MAX the groups are sorted independently — a warp per group of <= 32 members
// (bitonic network over shuffles), a CTA per larger group (bitonic in shared memory).
// ---------------------------------------------------------------------------------
constexpr uint32_t GS_MAX = 2048;
constexpr uint32_t GS_WARP = 256;   // a warp sorts up to eight members per lane in registers

// per group: size; max size -> cnt[3]; ids of the groups with 33 .. GS_WARP members -> medium[] (count cnt[7]), with
// GS_WARP + 1 .. GS_MAX members -> large[] (count cnt[4])
__global__ void __launch_bounds__(256)
k_group_stats(const uint32_t *__restrict__ cnt_in, const uint32_t *__restrict__ goff, uint32_t *__restrict__ cnt,
              uint32_t *__restrict__ medium, uint32_t *__restrict__ large, uint32_t *__restrict__ huge) {
    const uint32_t G = cnt_in[1];
    const uint32_t g = blockIdx.x * 256 + threadIdx.x;
    uint32_t size = g < G ? goff[g + 1] - goff[g] : 0u;
    if (size > GS_MAX) {   // no CTA can sort it: members -> cnt[5], id -> huge[], count cnt[6]
        atomicAdd(&cnt[5], size);
        huge[atomicAdd(&cnt[6], 1u)] = g;
    } else if (size > GS_WARP) {
        large[atomicAdd(&cnt[4], 1u)] = g;
    } else if (size > 32) {
        medium[atomicAdd(&cnt[7], 1u)] = g;
    }
    uint32_t m = __reduce_max_sync(0xffffffffu, size);
    if ((threadIdx.x & 31) == 0 && m) atomicMax(&cnt[3], m);
}

// the counters of a round go to the host through pinned memory the host polls: a copy and a stream synchronisation per
// round cost twice as much (measured: 16.5 against 7.3 us per hand-over, nine rounds per window)
__global__ void k_publish_counts(const uint32_t *__restrict__ cnt, volatile uint32_t *host, uint32_t seq) {
    if (threadIdx.x < 8) host[threadIdx.x] = cnt[threadIdx.x];
    __syncwarp();
    __threadfence_system();
    if (threadIdx.x == 0) host[8] = seq;
}

// warp per group (size <= 32)
__global__ void __launch_bounds__(256)
k_group_sort_small(uint32_t G, const uint32_t *__restrict__ goff, const uint32_t *__restrict__ rank, uint32_t h,
                   const uint32_t *__restrict__ vals, int kb, uint64_t *__restrict__ keys_out,
                   uint32_t *__restrict__ vals_out) {
    const uint32_t g = (blockIdx.x * 256 + threadIdx.x) >> 5;
    if (g >= G) return;
    const uint32_t off = goff[g], size = goff[g + 1] - off;
    if (size > 32) return;
    const uint32_t lane = lane_id();
    uint64_t e = lane < size ? ((uint64_t) rank[vals[off + lane] + h] << 32) | vals[off + lane] : ~0ull;
#pragma unroll
    for (uint32_t k = 2; k <= 32; k <<= 1) {
#pragma unroll
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            uint64_t o = __shfl_xor_sync(0xffffffffu, e, j);
            bool up = (lane & k) == 0, lower = (lane & j) == 0;
            e = (up == lower) ? (e < o ? e : o) : (e > o ? e : o);
        }
    }
    if (lane < size) {
        keys_out[off + lane] = ((uint64_t) g << kb) | (e >> 32);
        vals_out[off + lane] = (uint32_t) e;
    }
}

// Groups of up to 32 members, a warp per 32 CONSECUTIVE active suffixes (coalesced, and one sorting network serves all the
// groups that lie inside the block instead of one network per group of a handful of members): the block is sorted by
// (group, rank[i+h], suffix) with the members of a group that began before the block pinned in front and those of a
// group that runs past its end pinned behind - groups occupy fixed positions, so every member lands inside its own
// group.  A second network sorts the group that starts in the block and ends in the next one.
// (key layout: 6 bits of group, 29 + 29 bits of rank and suffix: windows below 2^29 positions)
__global__ void __launch_bounds__(256)
k_group_sort_blocks(uint32_t A, const uint32_t *__restrict__ gk, const uint32_t *__restrict__ goff,
                    const uint32_t *__restrict__ rank, uint32_t h, const uint32_t *__restrict__ vals, int kb, uint64_t *__restrict__ keys_out,
                    uint32_t *__restrict__ vals_out) {
    const uint32_t x = blockIdx.x * 256 + threadIdx.x;
    const uint32_t lane = lane_id(), base = x - lane, bend = base + 32;
    if (base >= A) return;
    const uint32_t FULL = 0xffffffffu;
    const bool valid = x < A;
    const uint32_t g = valid ? gk[x] : 0u;
    const uint32_t off = valid ? goff[g] : 0u, end = valid ? goff[g + 1] : 0u;
    const uint32_t g0 = __shfl_sync(FULL, g, 0);
    const bool interior = valid && off >= base && end <= bend;
    constexpr uint64_t M29 = (1ull << 29) - 1;
    uint64_t e = ~0ull;
    if (interior) e = ((uint64_t) (g - g0) << 58) | ((uint64_t) rank[vals[x] + h] << 29) | vals[x];
    else if (valid && off < base) e = 0;
    auto network = [&](uint64_t v) {
#pragma unroll
        for (uint32_t k = 2; k <= 32; k <<= 1) {
#pragma unroll
            for (uint32_t j = k >> 1; j > 0; j >>= 1) {
                const uint64_t o = __shfl_xor_sync(FULL, v, j);
                const bool up = (lane & k) == 0, lower = (lane & j) == 0;
                v = (up == lower) ? (v < o ? v : o) : (v > o ? v : o);
            }
        }
        return v;
    };
    if (__any_sync(FULL, interior && end - off > 1)) e = network(e);
    if (interior) {
        keys_out[x] = ((uint64_t) g << kb) | ((e >> 29) & M29);
        vals_out[x] = (uint32_t) (e & M29);
    }
    // the group that starts in this block and runs into the next one
    const uint32_t last = min(31u, A - 1 - base);
    const uint32_t gs = __shfl_sync(FULL, g, last), so = __shfl_sync(FULL, off, last), se = __shfl_sync(FULL, end, last);
    if (se > bend && so >= base && se - so <= 32) {
        const uint32_t size = se - so;
        uint64_t v = lane < size ? ((uint64_t) rank[vals[so + lane] + h] << 32) | vals[so + lane] : ~0ull;
        v = network(v);
        if (lane < size) {
            keys_out[so + lane] = ((uint64_t) gs << kb) | (v >> 32);
            vals_out[so + lane] = (uint32_t) v;
        }
    }
}

// warp per group of 33 .. GS_WARP members: R = 2, 4 or 8 members per lane (member i sits in lane i % 32, register i / 32), a
// bitonic network whose exchanges at distance >= 32 stay inside the lane - no shared memory, no barrier
// steps j = JSTART, JSTART / 2, ..., 1 of stage k of the bitonic network over the warp's 32 * R consecutive members
// (member with local index lane + 32 r, global index base + that; base is a multiple of 32 * R)
template <int R, uint32_t JSTART>
__device__ __forceinline__ void warp_bitonic_steps(uint64_t (&e)[R], uint32_t lane, uint32_t base, uint32_t k) {
#pragma unroll
    for (uint32_t j = JSTART; j > 0; j >>= 1) {
        if (j >= 32) {
            const uint32_t jr = j >> 5;
#pragma unroll
            for (uint32_t r = 0; r < (uint32_t) R; r++) {
                if ((r & jr) == 0) {
                    const uint32_t i = base + lane + 32 * r;
                    const bool up = (i & k) == 0;
                    const uint64_t a = e[r], b = e[r | jr];
                    if ((a > b) == up) {
                        e[r] = b;
                        e[r | jr] = a;
                    }
                }
            }
        } else {
#pragma unroll
            for (uint32_t r = 0; r < (uint32_t) R; r++) {
                const uint32_t i = base + lane + 32 * r;
                const uint64_t o = __shfl_xor_sync(0xffffffffu, e[r], j);
                const bool up = (i & k) == 0, lower = (lane & j) == 0;
                e[r] = (up == lower) ? (e[r] < o ? e[r] : o) : (e[r] > o ? e[r] : o);
            }
        }
    }
}

// all stages k = 2 .. 32 * R (a full sort of the warp's members when base == 0; with a base, the last stage follows the
// direction of the enclosing network)
template <int R, uint32_t K = 2>
__device__ __forceinline__ void warp_bitonic(uint64_t (&e)[R], uint32_t lane, uint32_t base = 0) {
    if constexpr (K <= 32u * R) {
        warp_bitonic_steps<R, K / 2>(e, lane, base, K);
        warp_bitonic<R, K * 2>(e, lane, base);
    }
}

__global__ void __launch_bounds__(128)
k_group_sort_medium(uint32_t nm, const uint32_t *__restrict__ medium, const uint32_t *__restrict__ goff,
                    const uint32_t *__restrict__ rank, uint32_t h, const uint32_t *__restrict__ vals, int kb, uint64_t *__restrict__ keys_out,
                    uint32_t *__restrict__ vals_out) {
    const uint32_t w = (blockIdx.x * 128 + threadIdx.x) >> 5;
    if (w >= nm) return;
    const uint32_t g = medium[w];
    const uint32_t off = goff[g], size = goff[g + 1] - off;
    const uint32_t lane = lane_id();
    if (size <= 64) {
        uint64_t e[2];
#pragma unroll
        for (int r = 0; r < 2; r++) {
            const uint32_t i = lane + 32 * r;
            e[r] = i < size ? ((uint64_t) rank[vals[off + i] + h] << 32) | vals[off + i] : ~0ull;
        }
        warp_bitonic<2>(e, lane);
#pragma unroll
        for (int r = 0; r < 2; r++) {
            const uint32_t i = lane + 32 * r;
            if (i < size) {
                keys_out[off + i] = ((uint64_t) g << kb) | (e[r] >> 32);
                vals_out[off + i] = (uint32_t) e[r];
            }
        }
    } else if (size <= 128) {
        uint64_t e[4];
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const uint32_t i = lane + 32 * r;
            e[r] = i < size ? ((uint64_t) rank[vals[off + i] + h] << 32) | vals[off + i] : ~0ull;
        }
        warp_bitonic<4>(e, lane);
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const uint32_t i = lane + 32 * r;
            if (i < size) {
                keys_out[off + i] = ((uint64_t) g << kb) | (e[r] >> 32);
                vals_out[off + i] = (uint32_t) e[r];
            }
        }
    } else {
        uint64_t e[8];
#pragma unroll
        for (int r = 0; r < 8; r++) {
            const uint32_t i = lane + 32 * r;
            e[r] = i < size ? ((uint64_t) rank[vals[off + i] + h] << 32) | vals[off + i] : ~0ull;
        }
        warp_bitonic<8>(e, lane);
#pragma unroll
        for (int r = 0; r < 8; r++) {
            const uint32_t i = lane + 32 * r;
            if (i < size) {
                keys_out[off + i] = ((uint64_t) g << kb) | (e[r] >> 32);
                vals_out[off + i] = (uint32_t) e[r];
            }
        }
    }
}

// CTA per group (GS_WARP < size <= GS_MAX): every warp keeps 256 consecutive members in registers; the stages of the
// bitonic network up to k = 256 and the steps at distance < 256 of the later stages run there (shuffles, no barrier),
// only the steps at distance >= 256 go through shared memory: 13 barriers for 2,048 members instead of 66
__global__ void __launch_bounds__(256)
k_group_sort_large(const uint32_t *__restrict__ large, const uint32_t *__restrict__ goff, const uint32_t *__restrict__ rank, uint32_t h,
                   const uint32_t *__restrict__ vals, int kb, uint64_t *__restrict__ keys_out,
                   uint32_t *__restrict__ vals_out) {
    __shared__ uint64_t sm[GS_MAX];
    const uint32_t g = large[blockIdx.x];
    const uint32_t off = goff[g], size = goff[g + 1] - off;
    uint32_t n2 = 512;
    while (n2 < size) n2 <<= 1;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31, base = warp * 256;
    const bool mine = base < n2;   // (warps beyond the padded size only take part in the barriers and the shared steps)
    uint64_t e[8];
#pragma unroll
    for (int r = 0; r < 8; r++) {
        const uint32_t i = base + lane + 32 * r;
        e[r] = (mine && i < size) ? ((uint64_t) rank[vals[off + i] + h] << 32) | vals[off + i] : ~0ull;
    }
    if (mine) warp_bitonic<8>(e, lane, base);
    for (uint32_t k = 512; k <= n2; k <<= 1) {
        if (mine) {
#pragma unroll
            for (int r = 0; r < 8; r++) sm[base + lane + 32 * r] = e[r];
        }
        __syncthreads();
        for (uint32_t j = k >> 1; j >= 256; j >>= 1) {
            for (uint32_t t = threadIdx.x; t < n2 / 2; t += 256) {
                const uint32_t i = 2 * t - (t & (j - 1));  // lower index of the pair (bit j clear)
                const uint32_t p = i + j;
                const uint64_t a = sm[i], b = sm[p];
                const bool up = (i & k) == 0;
                if ((a > b) == up) {
                    sm[i] = b;
                    sm[p] = a;
                }
            }
            __syncthreads();
        }
        if (mine) {
#pragma unroll
            for (int r = 0; r < 8; r++) e[r] = sm[base + lane + 32 * r];
            warp_bitonic_steps<8, 128>(e, lane, base, k);
        }
    }
    if (mine) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
            const uint32_t i = base + lane + 32 * r;
            if (i < size) {
                keys_out[off + i] = ((uint64_t) g << kb) | (e[r] >> 32);
                vals_out[off + i] = (uint32_t) e[r];
            }
        }
    }
}

// The few groups with more than GS_MAX members (a dozen per cent of the active suffixes after the first round, next
// to none after the second) are gathered into compact buffers, ordered there by one radix sort on
// (index in huge[], rank[i+h]) and written back in place, while every other group is sorted where it stands.
// offsets of the huge groups in the compact buffers (one CTA; a few hundred groups at most per round)
__global__ void __launch_bounds__(256)
k_huge_offsets(uint32_t nh, const uint32_t *__restrict__ huge, const uint32_t *__restrict__ goff, uint32_t *__restrict__ hoff) {
    __shared__ uint32_t part[256];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t b = 0; b < nh; b += 256) {
        const uint32_t j = b + threadIdx.x;
        const uint32_t g = j < nh ? huge[j] : 0u;
        const uint32_t size = j < nh ? goff[g + 1] - goff[g] : 0u;
        part[threadIdx.x] = size;
        __syncthreads();
        for (uint32_t d = 1; d < 256; d <<= 1) {
            uint32_t v = threadIdx.x >= d ? part[threadIdx.x - d] : 0u;
            __syncthreads();
            part[threadIdx.x] += v;
            __syncthreads();
        }
        if (j < nh) hoff[j] = carry + part[threadIdx.x] - size;
        __syncthreads();
        if (threadIdx.x == 255) carry += part[255];
        __syncthreads();
    }
    if (threadIdx.x == 0) hoff[nh] = carry;
}

__global__ void __launch_bounds__(256)
k_huge_gather(const uint32_t *__restrict__ huge, const uint32_t *__restrict__ hoff, const uint32_t *__restrict__ goff,
              const uint32_t *__restrict__ rank, uint32_t h, const uint32_t *__restrict__ vals, int kb, uint64_t *__restrict__ ck,
              uint32_t *__restrict__ cv) {
    const uint32_t j = blockIdx.y, g = huge[j];
    const uint32_t off = goff[g], size = goff[g + 1] - off, ho = hoff[j];
    for (uint32_t i = blockIdx.x * 256 + threadIdx.x; i < size; i += gridDim.x * 256) {
        ck[ho + i] = ((uint64_t) j << kb) | rank[vals[off + i] + h];
        cv[ho + i] = vals[off + i];
    }
}

__global__ void __launch_bounds__(256)
k_huge_scatter(const uint32_t *__restrict__ huge, const uint32_t *__restrict__ hoff, const uint32_t *__restrict__ goff,
               const uint64_t *__restrict__ ck, const uint32_t *__restrict__ cv, int kb, uint64_t *__restrict__ keys_out,
               uint32_t *__restrict__ vals_out) {
    const uint32_t j = blockIdx.y, g = huge[j];
    const uint32_t off = goff[g], size = goff[g + 1] - off, ho = hoff[j];
    const uint64_t low = (1ull << kb) - 1;
    for (uint32_t i = blockIdx.x * 256 + threadIdx.x; i < size; i += gridDim.x * 256) {
        keys_out[off + i] = ((uint64_t) g << kb) | (ck[ho + i] & low);
        vals_out[off + i] = cv[ho + i];
    }
}

// ---------------------------------------------------------------------------------
// K5: LCP (Kasai over 32-position segments; restarts cost one direct compare per segment)
// ---------------------------------------------------------------------------------
constexpr int LCP_SEG = 16;

// 8 text bytes starting at an arbitrary address: two aligned 8-byte loads + funnel shift
__device__ __forceinline__ uint64_t load8_unaligned(const uint8_t *__restrict__ p) {
    uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint64_t *q = reinterpret_cast<const uint64_t *>(a & ~(uintptr_t) 7);
    uint32_t sh = (uint32_t) (a & 7) * 8;
    uint64_t lo = q[0];
    if (sh == 0) return lo;
    uint64_t hi = q[1];
    return (lo >> sh) | (hi << (64 - sh));
}

__global__ void __launch_bounds__(128)
k_lcp(const uint8_t *__restrict__ text, const uint16_t *__restrict__ dist, const uint32_t *__restrict__ sa,
      const uint32_t *__restrict__ rank, uint32_t n, uint32_t *__restrict__ lcp) {
    uint32_t t = blockIdx.x * 128 + threadIdx.x;
    uint64_t i0 = (uint64_t) t * LCP_SEG;
    if (i0 >= n) return;
    uint32_t i1 = (uint32_t) (i0 + LCP_SEG < n ? i0 + LCP_SEG : n);
    uint32_t h = 0;
    for (uint32_t i = (uint32_t) i0; i < i1; i++) {
        uint32_t r = rank[i];
        if (r == 0) {
            lcp[0] = 0;
            h = 0;
            continue;
        }
        uint32_t p = sa[r - 1];
        uint32_t lim = min((uint32_t) dist[i], (uint32_t) dist[p]);
        if (h > lim) h = lim;
        // extend 8 bytes at a time (the text buffer has 16 bytes of slack past its end)
        while (h < lim) {
            uint64_t x = load8_unaligned(text + i + h) ^ load8_unaligned(text + p + h);
            if (x) {
                h += (uint32_t) (__ffsll((long long) x) - 1) >> 3;
                break;
            }
            h += 8;
        }
        if (h > lim) h = lim;
        lcp[r] = h;
        if (h) h--;
    }
}

// K5a/K5b: the same LCP array in two passes that keep the random accesses inside the L2.
// Kasai's walk (k_lcp above) visits the text in order and, per position, touches sa[rank-1], dist[p], text[p+h] and
// lcp[r] at random in ~200 MB of arrays: six 32-byte DRAM sectors per position (ncu: 2.6 GB read for 214 MB of
// algorithmic bytes).  Here pass a walks the SUFFIX ARRAY in order - sa and lcp stream, only text and dist (40 MB,
// L2-resident) are touched at random - and compares neighbours from scratch, but only up to LCP_CAP bytes; the few
// positions whose match is longer are marked in a bitmap and pass b finishes them in text order with Kasai's carry
// (matches of consecutive text positions shrink by at most one), so long repeats still cost O(length), not O(length^2).
constexpr uint32_t LCP_CAP = 64;

__global__ void __launch_bounds__(256)
k_lcp_sa(const uint8_t *__restrict__ text, const uint16_t *__restrict__ dist, const uint32_t *__restrict__ sa, uint32_t n,
         uint2 *__restrict__ leaf, uint32_t *__restrict__ longmap) {
    const uint32_t r = blockIdx.x * 256 + threadIdx.x;
    if (r >= n) return;
    uint32_t h = 0;
    const uint32_t a = __ldcs(sa + r);
    if (r > 0) {
        const uint32_t b = __ldcs(sa + r - 1);
        const uint32_t lim = min((uint32_t) dist[a], (uint32_t) dist[b]);
        const uint32_t stop = min(lim, LCP_CAP);
        bool diff = false;
        while (h < stop) {
            const uint64_t x = load8_unaligned(text + a + h) ^ load8_unaligned(text + b + h);
            if (x) {
                h += (uint32_t) (__ffsll((long long) x) - 1) >> 3;
                diff = true;
                break;
            }
            h += 8;
        }
        if (h > lim) h = lim;
        if (!diff && h >= LCP_CAP && lim > LCP_CAP) {  // the match may go on: finished by k_lcp_long
            h = LCP_CAP;
            atomicOr(longmap + (a >> 5), 1u << (a & 31));
        }
    }
    leaf[r] = make_uint2(a, h | ((uint32_t) dist[a] << 16));   // the leaf level of the block-min trees (MinTree::leaf)
}

__global__ void __launch_bounds__(128)
k_lcp_long(const uint8_t *__restrict__ text, const uint16_t *__restrict__ dist, const uint32_t *__restrict__ sa,
           const uint32_t *__restrict__ rank, const uint32_t *__restrict__ longmap, uint32_t n, uint2 *__restrict__ leaf) {
    // one thread per 32 text positions (one word of the bitmap); words without a long match cost one load
    const uint32_t w = blockIdx.x * 128 + threadIdx.x;
    if ((uint64_t) w * 32 >= n) return;
    uint32_t bits = longmap[w];
    uint32_t h = 0, prev = 0xFFFFFFFFu;
    while (bits) {
        const uint32_t k = __ffs(bits) - 1;
        bits &= bits - 1;
        const uint32_t i = w * 32 + k;
        const uint32_t r = rank[i];
        const uint32_t p = sa[r - 1];  // (r > 0: position i was marked by a comparison with its predecessor)
        const uint32_t lim = min((uint32_t) dist[i], (uint32_t) dist[p]);
        // Kasai: the match of position i is at least the match of i - 1 minus one (when i - 1 was long as well)
        h = (prev + 1 == i && h > LCP_CAP) ? h - 1 : LCP_CAP;
        if (h > lim) h = lim;
        while (h < lim) {
            const uint64_t x = load8_unaligned(text + i + h) ^ load8_unaligned(text + p + h);
            if (x) {
                h += (uint32_t) (__ffsll((long long) x) - 1) >> 3;
                break;
            }
            h += 8;
        }
        if (h > lim) h = lim;
        leaf[r].y = h | ((uint32_t) dist[i] << 16);
        prev = i;
    }
}

// ---------------------------------------------------------------------------------
// K6: longest previous factor. reach[s] = s + M(s) for the new positions s >= s0.
//
// With NODES the same two neighbour searches also answer the arena law of the reference's window
// rotation (PiXiuCtrl.cpp:13; MemPool.cpp:7-37) — which suffixes of the new records create a
// suffix-tree leaf and which of those also split an edge (verified against the reference's MemPool
// counters, DESIGN.md "Window rotation"):
//   leaf(s)   <=>  M(s) ends before the record does (suffixes still pending at SuffixTree::reset()
//                  are dropped, SuffixTree.cpp:302);
//   split(s)  <=>  leaf(s), M > 0 and the locus of w = T[s..s+M) is not yet an explicit node, i.e. the
//                  leftmost occurrence of w does not end its record (it would be an ex-leaf that simply
//                  gains a child, SuffixTree.cpp:223-230) and all earlier occurrences of w that continue
//                  do so with the same byte.
// Earlier occurrences of w with a smaller next byte sit above rank[s] in the suffix array, those with a
// larger one below; occurrences that end their record ("terminal") sort first.  One bit per position.
// ---------------------------------------------------------------------------------
#ifndef PIXIU_LPF_FAST
#define PIXIU_LPF_FAST 12
#endif
template <bool NODES>
__global__ void __launch_bounds__(256)
k_lpf(MinTree T, const uint8_t *__restrict__ text, const uint32_t *__restrict__ rank, const uint16_t *__restrict__ dist,
      uint32_t s0, uint32_t n, uint32_t *__restrict__ reach, uint32_t *__restrict__ leafmask,
      uint32_t *__restrict__ splitmask, uint8_t *__restrict__ flagp /* nullptr: k_flag_scatter does it later */) {
    const uint32_t k = blockIdx.x * 256 + threadIdx.x;
    const uint32_t s = s0 + k;
    bool leaf = false, split = false;
    if (s < n) {
        uint32_t best = 0;
        const uint32_t d = dist[s];
        if (d != 0) {
            // level 0 of the trees: {sa, lcp} interleaved, or the two arrays (knob lcp_kasai)
            const uint2 *LF = T.leaf;
            const uint32_t *A0 = T.a[0], *L0 = T.l[0];
            // entry j as {sa, lcp | dist[sa] << 16}
            auto AL = [&](uint32_t j) {
                if (LF) return LF[j];
                const uint32_t a = A0[j];
                return make_uint2(a, L0[j] | ((uint32_t) dist[a] << 16));
            };
            const uint32_t r = rank[s];
            // nearest smaller text position above r: lcp = min L[j+1..r].  Most searches end within a few
            // entries: a plain scan of up to LPF_FAST neighbours first, the block-min tree only for the rest
            constexpr int LPF_FAST = PIXIU_LPF_FAST;
            uint32_t l1 = AL(r).y & 0xFFFFu;
            uint32_t d_jl = 0;   // dist[sa[jl]]: bytes left in the record of the occurrence above
            int64_t jl = -1;
            if (l1) {
                uint32_t pos = r;
                bool open = true;
                for (int k = 0; k < LPF_FAST && pos > 0; k++) {
                    pos--;
                    const uint2 e = AL(pos);
                    if (e.x < s) {
                        jl = pos;
                        d_jl = e.y >> 16;
                        open = false;
                        break;
                    }
                    l1 = min(l1, e.y & 0xFFFFu);
                    if (l1 == 0) {
                        open = false;
                        break;
                    }
                }
                if (open) {
                    jl = pos > 0 ? tree_search<true, false, true>(T, pos, s, l1, 0) : -1;
                    if (NODES && jl >= 0) d_jl = AL((uint32_t) jl).y >> 16;
                }
            }
            if (jl < 0) l1 = 0;
            // nearest smaller text position below r: lcp = min L[r+1..j]; ties with l1 matter for NODES
            const int64_t floor2 = (NODES && l1 > 0) ? (int64_t) l1 - 1 : (int64_t) l1;
            uint32_t l2 = 0xFFFFFFFFu;
            int64_t jr = n;
            {
                uint32_t pos = r;
                bool open = true;
                for (int k = 0; k < LPF_FAST && pos + 1 < n; k++) {
                    pos++;
                    const uint2 e = AL(pos);
                    l2 = min(l2, e.y & 0xFFFFu);
                    if (e.x < s) {
                        jr = pos;
                        open = false;
                        break;
                    }
                    if ((int64_t) l2 <= floor2) {
                        open = false;
                        break;
                    }
                }
                if (open) jr = pos + 1 < n ? tree_search<false, true, true>(T, pos, s, l2, floor2) : (int64_t) n;
            }
            if (jr >= (int64_t) n) l2 = 0;
            best = max(l1, l2);
            if (NODES) {
                const uint32_t M = best;
                leaf = M < d;
                if (leaf && M > 0) {
                    bool up = l1 == M && d_jl > M;          // a continuing earlier occurrence above
                    bool dn = l2 == M;                      // ... below (never terminal)
                    bool is_explicit = (up && dn) || (!up && !dn);  // two next bytes | only ex-leaf occurrences
                    if (!is_explicit) {
                        // exactly one side: a second distinct next byte needs an earlier occurrence beyond the
                        // child interval of (w + that side's byte), still inside w's interval (lcp >= M)
                        if (up) {
                            uint32_t a2 = 0xFFFFFFFFu;
                            int64_t x1 = tree_search<true, true, false>(T, (uint32_t) jl + 1, M + 1, a2, -1);
                            uint32_t acc2 = AL((uint32_t) x1).y & 0xFFFFu;
                            if (acc2 >= M) {
                                int64_t j2 = tree_search<true, false, true>(T, (uint32_t) x1, s, acc2, (int64_t) M - 1);
                                if (j2 >= 0 && acc2 >= M && (AL((uint32_t) j2).y >> 16) > M) is_explicit = true;
                            }
                        } else {
                            uint32_t a2 = 0xFFFFFFFFu;
                            int64_t y1 = tree_search<false, false, false>(T, (uint32_t) jr, M + 1, a2, -1);
                            if (y1 < (int64_t) n) {
                                uint32_t acc2 = 0xFFFFFFFFu;
                                int64_t j2 = tree_search<false, true, true>(T, (uint32_t) y1 - 1, s, acc2, (int64_t) M - 1);
                                if (j2 < (int64_t) n && acc2 >= M) is_explicit = true;
                            }
                        }
                        // ex-leaf: the leftmost occurrence of w ends its record.  Every record ends with
                        // 251,0 or 251,2, so this needs w to end in 0 or 2 — rare; only then pay for the
                        // SA interval of w and its minimum.
                        if (!is_explicit) {
                            uint8_t lastb = text[s + M - 1];
                            if (lastb == 0 || lastb == 2) {
                                uint32_t accl = 0xFFFFFFFFu, accr = 0xFFFFFFFFu;
                                tree_search<true, true, false>(T, r + 1, M, accl, 0);
                                tree_search<false, false, false>(T, r, M, accr, 0);
                                uint32_t t0 = min(accl, accr);
                                if (dist[t0] == M) is_explicit = true;
                            }
                        }
                    }
                    split = !is_explicit;
                }
            }
        }
        reach[s] = s + best;
        if (flagp) flagp[s + best] = 1;   // PASS flags = image of reach, plus the separators (d == 0: best == 0)
    }
    if (NODES) {
        uint32_t lm = __ballot_sync(0xffffffffu, leaf), sm = __ballot_sync(0xffffffffu, split);
        if ((threadIdx.x & 31) == 0) {
            leafmask[k >> 5] = lm;
            splitmask[k >> 5] = sm;
        }
    }
}

// ---------------------------------------------------------------------------------
// K7: flags
// ---------------------------------------------------------------------------------
// PASS flags = image of reach, plus the separators
__global__ void __launch_bounds__(256)
k_flag_scatter(const uint32_t *__restrict__ reach, const uint16_t *__restrict__ dist, uint32_t s0, uint32_t n,
               uint8_t *__restrict__ flagp) {
    uint32_t s = s0 + blockIdx.x * 256 + threadIdx.x;
    if (s >= n) return;
    if (dist[s] == 0) flagp[s] = 1;
    else flagp[reach[s]] = 1;
}

// 1-based position of byte j (a 251) inside its run of 251s: from the max-scan of "last non-251 position" when the
// window may hold long runs, else by walking back to `lo` at most (the runs are then at most 65 long: 32 escaped pairs
// and a terminator; no run crosses a record start)
__device__ __forceinline__ uint32_t run_pos251(const uint8_t *__restrict__ text, const uint32_t *__restrict__ lastnon, uint32_t lo,
                                               uint32_t j) {
    if (lastnon) return j - lastnon[j];
    uint32_t p = j;
    while (p > lo && text[p - 1] == 251) p--;
    return j - p + 1;
}

// escape-pair coherence (PiXiuStr.cpp:34-54): 1 iff byte i stays COMPRESS
__device__ __forceinline__ uint8_t pair_rule(const uint8_t *__restrict__ text, const uint16_t *__restrict__ dist,
                                             const uint8_t *__restrict__ flagp, const uint32_t *__restrict__ lastnon, uint32_t s0,
                                             uint32_t n, uint32_t i) {
    if (dist[i] == 0) return 0;
    bool c = !flagp[i];
    int64_t partner = -1;
    if (text[i] == 251) {
        const uint32_t k = run_pos251(text, lastnon, s0, i);
        partner = (k & 1) ? (int64_t) i + 1 : (int64_t) i - 1;
    } else if (i > s0 && text[i - 1] == 251 && (run_pos251(text, lastnon, s0, i - 1) & 1)) {
        partner = (int64_t) i - 1;
    }
    if (partner >= 0 && partner < (int64_t) n && dist[partner] != 0 && flagp[partner]) c = false;
    return c ? 1 : 0;
}

// output bytes contributed by position i: PASS / short run byte -> 1; last byte of a long run -> 6 or 8
__device__ __forceinline__ uint32_t contrib_at(uint32_t i, const uint8_t *flagc, const uint16_t *dist,
                                               const uint32_t *prevp, const uint32_t *nextp, int strict251) {
    if (dist[i] == 0) return 0;
    if (!flagc[i]) return 1;
    uint32_t rl = nextp[i] - prevp[i];  // prevp is +1 biased: run = [prevp, nextp)
    if (rl <= 6) return 1;
    if (i != nextp[i] - 1) return 0;
    return (rl > 255 || (rl == 251 && !strict251)) ? 8 : 6;
}

// ---------------------------------------------------------------------------------
// K8+K9: run-end pointer query and emission
// ---------------------------------------------------------------------------------
// (idx << 16 | to) of the reference's pointer for the long run ending at i (length rl), or 0xFFFFFFFF when the
// window holds no earlier occurrence:  s* = min{s in record : reach(s) > i} by binary search (reach is
// monotone), then the leftmost occurrence of D[s*..i] = minimum text position in its SA interval.
__device__ __forceinline__ uint32_t run_pointer(const MinTree &T, const uint16_t *__restrict__ recid,
                                                const uint32_t *__restrict__ rec_start, const uint32_t *__restrict__ rank,
                                                const uint32_t *__restrict__ reach, const uint16_t *__restrict__ gidx,
                                                uint32_t i, uint32_t rl) {
    uint32_t rec = recid[i];
    uint32_t lo = rec_start[rec], hi = i + 1;
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (reach[mid] > i) hi = mid;
        else lo = mid + 1;
    }
    uint32_t sstar = lo, E = i + 1 - sstar;
    uint32_t r = rank[sstar];
    uint32_t accl = 0xFFFFFFFFu, accr = 0xFFFFFFFFu;
    // SA interval of D[s*..i]: [x, y) with L[x] < E and L[y] < E; leftmost occurrence = min sa over it
    tree_search<true, true, false>(T, r + 1, E, accl, 0);
    tree_search<false, false, false>(T, r, E, accr, 0);
    uint32_t left = min(accl, accr);
    if (left >= sstar || E < rl) return 0xFFFFFFFFu;
    uint32_t src = recid[left];
    uint32_t to = left + E - rec_start[src];
    return ((gidx ? (uint32_t) gidx[src] : src) << 16) | to;
}

// multi-GPU: this shard's candidate pointer of every long run (slot = number of long runs before it)
__global__ void __launch_bounds__(256)
k_candidates(MinTree T, const uint16_t *__restrict__ dist, const uint16_t *__restrict__ recid,
             const uint32_t *__restrict__ rec_start, const uint32_t *__restrict__ rank, const uint32_t *__restrict__ reach,
             const uint8_t *__restrict__ flagc, const uint32_t *__restrict__ prevp, const uint32_t *__restrict__ nextp,
             const uint32_t *__restrict__ runidx, const uint16_t *__restrict__ gidx, uint32_t s0, uint32_t n, int strict251,
             uint32_t *__restrict__ cand) {
    // (the run ends of a CTA are handled by consecutive threads, see k_emit)
    __shared__ uint32_t task[256];
    __shared__ uint32_t ntask;
    if (threadIdx.x == 0) ntask = 0;
    __syncthreads();
    {
        const uint32_t i0 = s0 + blockIdx.x * 256 + threadIdx.x;
        if (i0 < n && contrib_at(i0, flagc, dist, prevp, nextp, strict251) > 1) task[atomicAdd(&ntask, 1u)] = i0;
    }
    __syncthreads();
    if (threadIdx.x >= ntask) return;
    const uint32_t i = task[threadIdx.x];
    cand[runidx[i]] = run_pointer(T, recid, rec_start, rank, reach, gidx, i, nextp[i] - prevp[i]);
}

__global__ void __launch_bounds__(256)
k_emit(MinTree T, const uint8_t *__restrict__ text, const uint16_t *__restrict__ dist,
       const uint16_t *__restrict__ recid, const uint32_t *__restrict__ rec_start, const uint32_t *__restrict__ rank,
       const uint32_t *__restrict__ reach, const uint8_t *__restrict__ flagc, const uint32_t *__restrict__ prevp,
       const uint32_t *__restrict__ nextp, const uint32_t *__restrict__ off, uint32_t s0, uint32_t n, int strict251,
       uint8_t *__restrict__ enc_out /* already offset so that off[] indexes it directly */, uint32_t *__restrict__ err,
       const uint32_t *__restrict__ cand, const uint32_t *__restrict__ runidx, const uint16_t *__restrict__ gidx) {
    // Literals are written by the thread of their position.  The ends of the long runs - one position in ten, each a
    // binary search plus two walks of the block-min trees, all dependent random loads - are first collected per CTA and
    // then handled by consecutive threads: a warp with 32 searches in flight instead of three
    __shared__ uint32_t task[256];
    __shared__ uint32_t ntask;
    if (threadIdx.x == 0) ntask = 0;
    __syncthreads();
    {
        const uint32_t i0 = s0 + blockIdx.x * 256 + threadIdx.x;
        if (i0 < n) {
            const uint32_t c0 = contrib_at(i0, flagc, dist, prevp, nextp, strict251);
            if (c0 == 1) enc_out[off[i0]] = text[i0];
            else if (c0 > 1) task[atomicAdd(&ntask, 1u)] = i0;
        }
    }
    __syncthreads();
    if (threadIdx.x >= ntask) return;
    const uint32_t i = task[threadIdx.x];
    const uint32_t rl = nextp[i] - prevp[i];   // long run ending at i
    const uint32_t c = (rl > 255 || (rl == 251 && !strict251)) ? 8u : 6u;
    const uint32_t o = off[i];
    uint32_t src, to;
    if (cand) {  // multi-GPU: the pointer was min-reduced across the shards
        uint32_t key = cand[runidx[i]];
        if (key == 0xFFFFFFFFu) {
            atomicExch(err, 8u);
            return;
        }
        src = key >> 16;
        to = key & 0xffff;
    } else {
        uint32_t key = run_pointer(T, recid, rec_start, rank, reach, gidx, i, rl);
        if (key == 0xFFFFFFFFu) {
            atomicExch(err, 2u);
            return;
        }
        src = key >> 16;
        to = key & 0xffff;
    }
    enc_out[o] = 251;
    if (c == 8) {
        uint32_t from = to - rl;
        enc_out[o + 1] = 1;
        enc_out[o + 2] = (uint8_t) src;
        enc_out[o + 3] = (uint8_t) (src >> 8);
        enc_out[o + 4] = (uint8_t) to;
        enc_out[o + 5] = (uint8_t) (to >> 8);
        enc_out[o + 6] = (uint8_t) from;
        enc_out[o + 7] = (uint8_t) (from >> 8);
    } else {
        enc_out[o + 1] = (uint8_t) rl;
        enc_out[o + 2] = (uint8_t) src;
        enc_out[o + 3] = (uint8_t) (src >> 8);
        enc_out[o + 4] = (uint8_t) to;
        enc_out[o + 5] = (uint8_t) (to >> 8);
    }
}

// per new record: encoded offset/length, decoded length, tile descriptors
__global__ void __launch_bounds__(256)
k_record_tables(uint32_t n_new, uint32_t win_first, uint32_t g_first, uint32_t g_chunk_first,
                const uint32_t *__restrict__ rec_start, const uint32_t *__restrict__ off, uint32_t off_s0,
                uint64_t enc_base, uint64_t *__restrict__ rec_enc_off, uint32_t *__restrict__ rec_enc_len,
                uint32_t *__restrict__ rec_dec_len, uint32_t *__restrict__ rec_first) {
    uint32_t r = blockIdx.x * 256 + threadIdx.x;
    if (r >= n_new) return;
    uint32_t a = rec_start[win_first + r], b = rec_start[win_first + r + 1] - 1;
    uint32_t g = g_first + r;
    rec_enc_off[g] = enc_base + (off[a] - off_s0);
    rec_enc_len[g] = off[b] - off[a];
    rec_dec_len[g] = b - a;
    rec_first[g] = g_chunk_first;
}

// tile descriptor: enc offset (within the record) of the token holding decoded byte t*TILE,
// and how many bytes of that token lie before it
__global__ void __launch_bounds__(256)
k_tile_desc(uint32_t n_tiles_new, uint32_t tile_first, uint32_t n_new, uint32_t win_first, uint32_t g_first,
            const uint32_t *__restrict__ rec_tile_base, const uint32_t *__restrict__ rec_start,
            const uint32_t *__restrict__ off, const uint8_t *__restrict__ flagc, const uint32_t *__restrict__ prevp,
            const uint32_t *__restrict__ nextp, const uint8_t *__restrict__ text, const uint32_t *__restrict__ lastnon,
            uint32_t *__restrict__ tile_desc) {
    uint32_t t = blockIdx.x * 256 + threadIdx.x;
    if (t >= n_tiles_new) return;
    uint32_t gt = tile_first + t;
    // record owning tile gt: binary search over rec_tile_base[g_first .. g_first+n_new)
    uint32_t lo = 0, hi = n_new;
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (rec_tile_base[g_first + mid] <= gt) lo = mid;
        else hi = mid;
    }
    uint32_t a = rec_start[win_first + lo];
    uint32_t i = a + (gt - rec_tile_base[g_first + lo]) * TILE;
    uint32_t skip = 0;
    bool in_ref = false;
    if (flagc[i]) {
        uint32_t rl = nextp[i] - prevp[i];
        if (rl > 6) {
            skip = i - prevp[i];
            in_ref = true;
        }
    }
    if (!in_ref) {
        // a tile that starts on the 2nd byte of an escape pair must take that byte literally
        uint8_t b = text[i];
        bool second = false;
        if (b == 251) second = (run_pos251(text, lastnon, a, i) & 1) == 0;
        else if (i > a && text[i - 1] == 251) second = (run_pos251(text, lastnon, a, i - 1) & 1) != 0;
        if (second) skip = 0xFFFF;
    }
    tile_desc[gt] = (off[i] - off[a]) | (skip << 16);
}

// ---------------------------------------------------------------------------------
// host orchestration
// ---------------------------------------------------------------------------------
static int bits_for(uint64_t v) {  // bits needed to represent values in [0, v]
    int b = 0;
    while (v) {
        b++;
        v >>= 1;
    }
    return b ? b : 1;
}

// Builds sa/rank over the window text [0, N)
static void build_suffix_array(Store &S, uint32_t N) {
    EncodeScratch &E = S.es;
    cudaStream_t st = S.st;
    E.keys0.reserve_discard(N);
    E.keys1.reserve_discard(N);
    E.vals0.reserve_discard(N);
    E.vals1.reserve_discard(N);
    E.slot0.reserve_discard(N);
    E.slot1.reserve_discard(N);
    E.gk.reserve_discard(N);
    E.sa.reserve_discard(N);
    E.rank.reserve_discard(N + 8);
    // [0] active count, [1] group count, [2] sticky error flag (cleared by flush_mirrors), [3] largest group,
    // [4] number of big groups, [8..15] byte presence set of the batch
    PX_CUDA(cudaMemsetAsync(E.counters.p, 0, 2 * sizeof(uint32_t), st));
    PX_CUDA(cudaMemsetAsync(E.counters.p + 3, 0, 5 * sizeof(uint32_t), st));
    E.goff.reserve_discard((size_t) N / 2 + 4);
    E.glarge.reserve_discard((size_t) N / 32 + 4);
    E.gmedium.reserve_discard((size_t) N / 32 + 4);
    E.ghuge.reserve_discard((size_t) N / GS_MAX + 4);
    E.hoff.reserve_discard((size_t) N / GS_MAX + 5);
    int L = 0;

    Profiler *PF = S.prof.on ? &S.prof : nullptr;
    // dense symbol map of the byte values present in the window (always 0, 2 and 251: the terminators)
    uint8_t h_map[256];
    int nsymbols = 0;
    for (int b = 0; b < 256; b++) {
        h_map[b] = (uint8_t) nsymbols;
        if ((S.win_present[b >> 5] >> (b & 31)) & 1u) nsymbols++;
    }
    int bits = 1;
    while ((1 << bits) <= nsymbols) bits++;       // symbols 1..nsymbols, 0 = past the record end
    const int nsym = bits <= 7 ? 8 : 7;            // 56 bits (7 passes) up to 8-bit symbols, else 63 bits (8 passes)
    E.symmap.reserve_discard(256);
    PX_CUDA(cudaMemcpyAsync(E.symmap.p, h_map, 256, cudaMemcpyHostToDevice, st));
    S.prof.begin(PC_INIT_KEYS, st);
    k_init_keys<<<div_up<uint32_t>(N, 256), 256, 0, st>>>(S.w_text.p, S.w_dist.p, N, nsym, bits, E.symmap.p, E.keys0.p);
    S.prof.end(st, 11.0 * N, 1);
    L++;
    int cur = radix_sort_pairs<uint64_t>(E.keys0.p, E.keys1.p, E.vals0.p, E.vals1.p, N, 0, nsym * bits, true, E.rs, E.counters.p + 2, st, &L, PF);
    uint64_t *skeys = cur ? E.keys1.p : E.keys0.p;
    uint32_t *svals = cur ? E.vals1.p : E.vals0.p;
    uint32_t *slot_cur = nullptr;  // nullptr: slot[a] = a (first round)
    uint32_t *slot_next = E.slot0.p;
    uint32_t A = N;
    uint32_t h = (uint32_t) nsym;
    const int kb = bits_for(N);
    bool initial = true;
    uint32_t *d_cnt = E.counters.p;
    uint32_t *sa = E.sa.p, *rank = E.rank.p, *gk = E.gk.p;

    while (true) {
        HeadFn head{skeys, A, initial ? ((1ull << bits) - 1) : 0ull};
        S.prof.begin(PC_RANK_SCAN, st);
        // (1)+(2) one pass, two scans: the inclusive MAX of "index+1 of the latest group head" gives every
        // element the slot of its group head (its rank; written with sa), the exclusive SUM of
        // (active, active head) compacts the members of groups larger than one and numbers those groups
        {
            const uint32_t *sl = slot_cur;
            const uint32_t *sv = svals;
            uint32_t *sl_out = slot_next;
            uint32_t *vals_out = (svals == E.vals0.p) ? E.vals1.p : E.vals0.p;
            const uint32_t An = A;
            uint32_t *goff = E.goff.p;
            device_scan_dual(
                A,
                [=] __device__(size_t a, uint32_t &m, unsigned long long &sm) {
                    bool hd = head((uint32_t) a), hn = head((uint32_t) a + 1);
                    bool act = !(hd && hn);
                    m = hd ? (uint32_t) a + 1u : 0u;
                    sm = act ? (1ull | ((unsigned long long) hd << 32)) : 0ull;
                },
                [=] __device__(size_t a, uint32_t hp1, unsigned long long ex, uint32_t m, unsigned long long sm) {
                    const uint32_t hp = hp1 - 1;  // index of the group head (element 0 is always a head)
                    const uint32_t v = sv[a];
                    sa[sl ? sl[a] : (uint32_t) a] = v;
                    rank[v] = sl ? sl[hp] : hp;
                    const bool act = sm & 1ull, hd = m != 0;
                    const uint32_t dst = (uint32_t) ex, g = (uint32_t) (ex >> 32);
                    if (act) {
                        sl_out[dst] = sl ? sl[a] : (uint32_t) a;
                        vals_out[dst] = v;
                        gk[dst] = hd ? g : g - 1;
                        if (hd) goff[g] = dst;  // first member of surviving group g
                    }
                    if (a == An - 1) {
                        uint32_t At = dst + (act ? 1u : 0u), Gt = g + ((act && hd) ? 1u : 0u);
                        d_cnt[0] = At;
                        d_cnt[1] = Gt;
                        goff[Gt] = At;
                    }
                },
                E.scanws, st);
            L += 1;
            svals = vals_out;  // compacted values (unsorted for the next key) live here now
        }
        // group sizes: largest group and the list of groups too big for one warp
        k_group_stats<<<div_up<uint32_t>(A / 2 + 1, 256), 256, 0, st>>>(d_cnt, E.goff.p, d_cnt, E.gmedium.p, E.glarge.p, E.ghuge.p);
        L++;
        S.prof.end(st, 36.0 * A, 2);
        uint32_t h_cnt[8];
        {
            if (!E.h_round.p) {
                E.h_round.reserve_discard(16);
                memset(E.h_round.p, 0, 16 * sizeof(uint32_t));
            }
            volatile uint32_t *hr = E.h_round.p;
            if (++S.round_seq == 0) S.round_seq = 1;   // (0 is what a fresh buffer holds)
            const uint32_t seq = S.round_seq;
            k_publish_counts<<<1, 32, 0, st>>>(d_cnt, hr, seq);
            L++;
            for (uint32_t spins = 0; hr[8] != seq; spins++) {
                if ((spins & 0xFFFu) == 0xFFFu) {   // (a failed launch would never publish)
                    const cudaError_t q = cudaStreamQuery(st);
                    if (q != cudaSuccess && q != cudaErrorNotReady) PX_CUDA(q);
                    if (q == cudaSuccess && hr[8] != seq) throw std::runtime_error("suffix array: round counters never arrived");
                }
            }
            for (int i = 0; i < 8; i++) h_cnt[i] = hr[i];
        }
        if (h_cnt[2]) throw std::runtime_error("radix sort look-back timed out");
        uint32_t An = h_cnt[0], G = h_cnt[1];
        if (S.knobs.trace)
            fprintf(stderr, "[sa] N=%u sorted_by=%u active=%u groups=%u largest=%u huge_groups=%u huge_members=%u\n", N, h, An, G, h_cnt[3],
                    h_cnt[6], h_cnt[5]);
        if (An == 0) break;
        if (h > 65535u) throw std::runtime_error("suffix array: groups left after h > 65535");
        const uint32_t nlarge = h_cnt[4], hmem = h_cnt[5], nhuge = h_cnt[6], nmedium = h_cnt[7];
        PX_CUDA(cudaMemsetAsync(d_cnt + 3, 0, 5 * sizeof(uint32_t), st));
        if (hmem <= An / 2 && nhuge <= 65535u && !S.knobs.no_segsort) {   // (nhuge is a grid dimension below)
            // (3a) the groups are sorted independently by rank[i+h], each where it stands: a warp or a CTA per group of
            // up to GS_MAX members; the members of the few larger groups go through one radix sort of their own
            uint32_t *vout = (svals == E.vals0.p) ? E.vals1.p : E.vals0.p;
            // (the keys rank[i + h] are gathered by the sort kernels themselves: no pass that writes them out first)
            S.prof.begin(PC_SEG_SORT, st);
            if (kb <= 29)
                k_group_sort_blocks<<<div_up<uint32_t>(An, 256), 256, 0, st>>>(An, gk, E.goff.p, rank, h, svals, kb, E.keys0.p, vout);
            else
                k_group_sort_small<<<(unsigned) div_up<uint64_t>((uint64_t) G * 32, 256), 256, 0, st>>>(G, E.goff.p, rank, h, svals, kb,
                                                                                                  E.keys0.p, vout);
            if (nmedium)
                k_group_sort_medium<<<div_up<uint32_t>(nmedium, 4), 128, 0, st>>>(nmedium, E.gmedium.p, E.goff.p, rank, h, svals, kb,
                                                                                  E.keys0.p, vout);
            if (nlarge) k_group_sort_large<<<nlarge, 256, 0, st>>>(E.glarge.p, E.goff.p, rank, h, svals, kb, E.keys0.p, vout);
            S.prof.end(st, 28.0 * (An - hmem), 1 + (nmedium ? 1 : 0) + (nlarge ? 1 : 0));
            L += 1 + (nmedium ? 1 : 0) + (nlarge ? 1 : 0);
            if (nhuge) {
                E.hv0.reserve_discard(hmem);
                E.hv1.reserve_discard(hmem);
                E.hk1.reserve_discard(hmem);
                S.prof.begin(PC_ROUND_KEYS, st);
                k_huge_offsets<<<1, 256, 0, st>>>(nhuge, E.ghuge.p, E.goff.p, E.hoff.p);
                const dim3 hgrid(std::min<uint32_t>(div_up<uint32_t>(h_cnt[3], 256), 64), nhuge);
                k_huge_gather<<<hgrid, 256, 0, st>>>(E.ghuge.p, E.hoff.p, E.goff.p, rank, h, svals, kb, E.keys1.p, E.hv0.p);
                S.prof.end(st, 20.0 * hmem, 2);
                const int hb = bits_for(nhuge - 1);
                const int c3 = radix_sort_pairs<uint64_t>(E.keys1.p, E.hk1.p, E.hv0.p, E.hv1.p, hmem, 0, kb + hb, false, E.rs,
                                                          E.counters.p + 2, st, &L, PF);
                S.prof.begin(PC_ROUND_KEYS, st);
                k_huge_scatter<<<hgrid, 256, 0, st>>>(E.ghuge.p, E.hoff.p, E.goff.p, c3 ? E.hk1.p : E.keys1.p, c3 ? E.hv1.p : E.hv0.p, kb,
                                                      E.keys0.p, vout);
                S.prof.end(st, 24.0 * hmem, 1);
                L += 3;
            }
            skeys = E.keys0.p;
            svals = vout;
            slot_cur = slot_next;
            slot_next = (slot_next == E.slot0.p) ? E.slot1.p : E.slot0.p;
            A = An;
            h *= 2;
            initial = false;
            continue;
        }
        // (3b) next keys: (group, rank[i+h]) and radix sort
        uint64_t *kin = E.keys0.p, *kalt = E.keys1.p;
        S.prof.begin(PC_ROUND_KEYS, st);
        k_round_keys<<<div_up<uint32_t>(An, 256), 256, 0, st>>>(An, svals, gk, rank, h, kb, kin);
        S.prof.end(st, 20.0 * An, 1);
        L++;
        uint32_t *vin = svals, *valt = (svals == E.vals0.p) ? E.vals1.p : E.vals0.p;
        int gb = bits_for(G ? G - 1 : 0);
        int c2 = radix_sort_pairs<uint64_t>(kin, kalt, vin, valt, An, 0, kb + gb, false, E.rs, E.counters.p + 2, st, &L, PF);
        skeys = c2 ? kalt : kin;
        svals = c2 ? valt : vin;
        slot_cur = slot_next;
        slot_next = (slot_next == E.slot0.p) ? E.slot1.p : E.slot0.p;
        A = An;
        h *= 2;
        initial = false;
    }
    S.launches += L;
}

// Replays the reference's arena allocations (MemPool::p_malloc, MemPool.cpp:7-37) for the leaf /
// split events of the candidate records and returns how many of them fit before the rotation
// trigger `nth >= 2048` (PiXiuCtrl.cpp:13).  Commits pool_nth/pool_used for the accepted ones.
// (a) enqueue: word-level prefix of the arena blocks on the GPU and the copies of the masks to pinned memory
void Store::count_nodes_enqueue(uint32_t s0, uint32_t N) {
    EncodeScratch &E = es;
    const uint32_t M = N - s0;
    const size_t words = (size_t) div_up<uint32_t>(M, 256) * 8;  // masks written by k_lpf<true>
    // blocks per 32-position word (8 per leaf + 8 per split), exclusive prefix: the host then finds every
    // pool boundary by binary search instead of walking all words
    E.wordpre.reserve_discard(words + 1);
    {
        const uint32_t *lmk = E.leafmask.p, *smk = E.splitmask.p;
        uint32_t *wp = E.wordpre.p;
        device_scan<uint32_t>(
            words + 1,
            [=] __device__(size_t w) -> uint32_t {
                if (w >= words) return 0u;
                uint32_t l = lmk[w];
                return 8u * (uint32_t) (__popc(l) + __popc(smk[w] & l));
            },
            [=] __device__(size_t w, uint32_t v) { wp[w] = v; }, OpSum(), 0u, true, E.scanws, st);
        launches++;
    }
    E.h_leafmask.reserve_discard(words);
    E.h_splitmask.reserve_discard(words);
    E.h_wordpre.reserve_discard(words + 1);
    PX_CUDA(cudaMemcpyAsync(E.h_leafmask.p, E.leafmask.p, words * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    PX_CUDA(cudaMemcpyAsync(E.h_splitmask.p, E.splitmask.p, words * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    PX_CUDA(cudaMemcpyAsync(E.h_wordpre.p, E.wordpre.p, (words + 1) * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    PX_CUDA(cudaEventRecord(ev_nodes, st));
}

// (b) replay on the host (while the GPU already runs phase B on the candidate superset): returns the number
// of candidate records accepted
uint32_t Store::count_nodes_and_cut(uint32_t first_new, uint32_t s0, uint32_t N) {
    EncodeScratch &E = es;
    (void) N;
    auto t_a = std::chrono::steady_clock::now();
    PX_CUDA(cudaEventSynchronize(ev_nodes));
    auto t_b = std::chrono::steady_clock::now();
    const uint32_t *lm = E.h_leafmask.p, *sm = E.h_splitmask.p, *wp = E.h_wordpre.p;
    constexpr uint32_t C = 65535;  // POOL_BLOCK_NUM (MemPool.h:6)
    uint32_t nth = pool_nth, used = pool_used;
    auto alloc = [&](uint32_t blocks) {
        if (blocks > C - used) {  // does not fit: next pool (MemPool.cpp:23-29)
            nth++;
            used = 0;
        }
        used += blocks;
    };
    // blocks of the events at bit positions [0, k)
    auto P = [&](uint32_t k) -> uint64_t {
        uint32_t w = k >> 5, b = k & 31;
        uint64_t v = wp[w];
        if (b) {
            uint32_t mask = (1u << b) - 1, l = lm[w] & mask;
            v += 8u * (uint32_t) (__builtin_popcount(l) + __builtin_popcount(sm[w] & l));
        }
        return v;
    };
    // allocation by allocation over bits [k, k1) of one word: leaf 5,3 / leaf+split 5,5,3,3 (SuffixTree.cpp:193-231)
    auto walk_bits = [&](uint32_t k, uint32_t k1) {
        uint32_t w = k >> 5, b0 = k & 31, b1 = b0 + (k1 - k);
        uint32_t mask = (b1 >= 32 ? 0xFFFFFFFFu : ((1u << b1) - 1)) & ~((1u << b0) - 1);
        uint32_t ml = lm[w] & mask, ms = sm[w] & ml;
        while (ml) {
            uint32_t bit = ml & (0u - ml);
            if (ms & bit) {
                alloc(5);
                alloc(5);
                alloc(3);
                alloc(3);
            } else {
                alloc(5);
                alloc(3);
            }
            ml ^= bit;
        }
    };
    uint32_t accepted = 0;
    uint64_t node_blocks = 0;
    for (uint32_t r = first_new; r < win_R; r++) {
        if (r > first_new && nth >= 2048) break;  // PiXiuCtrl.cpp:13: checked before inserting record r
        uint32_t k = h_win_rec_start[r] - s0;
        const uint32_t kb = h_win_rec_start[r + 1] - 1 - s0;  // bit range [k, kb) of the record
        const uint64_t Pend = P(kb);
        node_blocks += Pend - P(k);
        while (k < kb) {
            uint64_t Pk = P(k);
            if (Pend - Pk <= C - used) {  // the rest of the record fits the current pool: no tail waste possible
                used += (uint32_t) (Pend - Pk);
                break;
            }
            // first word whose end exceeds the room left: everything before it fits as a whole
            uint64_t limit = Pk + (C - used);
            uint32_t lo = k >> 5, hi = (kb - 1) >> 5;  // find the smallest w in [lo, hi] with blocks up to end of w > limit
            {
                // the answer lies about (room / average blocks per word) words ahead: bracket it by galloping from
                // that guess instead of bisecting the whole record (a 40 KB page spans ~1,300 words and six pools)
                auto over = [&](uint32_t w) { return P(std::min<uint32_t>((w + 1) << 5, kb)) > limit; };
                const uint64_t span_blocks = Pend - Pk;
                const uint32_t span_words = hi - lo + 1;
                uint32_t g = lo + (uint32_t) std::min<uint64_t>((uint64_t) (C - used) * span_words / (span_blocks ? span_blocks : 1), hi - lo);
                if (over(g)) {
                    hi = g;
                    uint32_t step = 1;
                    while (hi > lo) {  // walk down until a word that is not over: the answer is in (that word, hi]
                        const uint32_t c = hi - lo > step ? hi - step : lo;
                        if (!over(c)) {
                            lo = c + 1;
                            break;
                        }
                        hi = c;
                        step <<= 1;
                    }
                } else {
                    lo = g + 1;
                    uint32_t step = 1;
                    while (lo < hi) {  // walk up until a word that is over: the answer is in [lo, that word]
                        const uint32_t c = hi - lo > step ? lo + step : hi;
                        if (over(c)) {
                            hi = c;
                            break;
                        }
                        lo = c + 1;
                        step <<= 1;
                    }
                }
            }
            while (lo < hi) {
                uint32_t mid = (lo + hi) >> 1;
                uint32_t wend = std::min<uint32_t>((mid + 1) << 5, kb);
                if (P(wend) > limit) hi = mid;
                else lo = mid + 1;
            }
            uint32_t wstart = std::max<uint32_t>(lo << 5, k), wend = std::min<uint32_t>((lo + 1) << 5, kb);
            used += (uint32_t) (P(wstart) - Pk);
            walk_bits(wstart, wend);
            k = wend;
        }
        accepted++;
        pool_nth = nth;
        pool_used = used;
    }
    if (knobs.trace) {
        auto t_c = std::chrono::steady_clock::now();
        static const auto t_origin = std::chrono::steady_clock::now();
        fprintf(stderr, "[rot] cand=%u accepted=%u nth=%u wait+d2h=%.3f ms replay=%.3f ms rho=%.3f at=%.3f ms\n", win_R - first_new, accepted,
                pool_nth, std::chrono::duration<double, std::milli>(t_b - t_a).count(),
                std::chrono::duration<double, std::milli>(t_c - t_b).count(), rho,
                std::chrono::duration<double, std::milli>(t_c - t_origin).count());
    }
    // running estimate of nodes per byte (drives the next candidate selection only; the cut is exact)
    uint32_t bytes = h_win_rec_start[first_new + accepted] - s0;
    if (bytes > (2u << 20)) {
        const double rho_new = std::max(0.02, (double) node_blocks / 8.0 / (double) bytes);
        // how far the previous estimate was off (decaying maximum): sizes the slack of the next candidate selection
        rho_err = std::max(0.7 * rho_err, std::fabs(rho_new / rho - 1.0));
        rho = rho_new;
    }
    return accepted;
}

// Phase A: suffix array, LCP, trees, longest previous factor of the records [first_new, win_R) of the open
// window (and, under the reference policy, the cut of the candidates at the rotation record).
void Store::enc_phase_a(uint32_t first_new, bool fuse_flags) {
    EncodeScratch &E = es;
    uint32_t N = win_N, R = win_R;
    const uint32_t s0 = h_win_rec_start[first_new];
    uint32_t n_new = R - first_new;
    uint32_t M = N - s0;  // new positions
    int L = 0;

    build_suffix_array(*this, N);

    // ---- LCP + trees ----
    prof.begin(PC_LCP, st);
    MinTree T{};
    if (knobs.lcp_kasai) {  // (knob: the one-pass Kasai walk into a plain lcp array, for A/B measurements)
        E.lcp.reserve_discard(N);
        k_lcp<<<div_up<uint32_t>(div_up<uint32_t>(N, LCP_SEG), 128), 128, 0, st>>>(w_text.p, w_dist.p, E.sa.p, E.rank.p, N, E.lcp.p);
        L++;
    } else {
        const uint32_t words = div_up<uint32_t>(N, 32);
        E.longmap.reserve_discard(words + 1);
        E.leaf.reserve_discard(N);
        PX_CUDA(cudaMemsetAsync(E.longmap.p, 0, (size_t) words * sizeof(uint32_t), st));
        k_lcp_sa<<<div_up<uint32_t>(N, 256), 256, 0, st>>>(w_text.p, w_dist.p, E.sa.p, N, E.leaf.p, E.longmap.p);
        k_lcp_long<<<div_up<uint32_t>(words, 128), 128, 0, st>>>(w_text.p, w_dist.p, E.sa.p, E.rank.p, E.longmap.p, N, E.leaf.p);
        L += 2;
        T.leaf = E.leaf.p;
    }
    prof.end(st, 20.0 * N, 2);
    {
        size_t total = 0;
        uint32_t sz = N;
        int nlev = 1;
        while (sz > 1 && nlev < TREE_MAX_LEVELS) {
            sz = div_up<uint32_t>(sz, TREE_B);
            total += sz;
            nlev++;
        }
        E.tree_a.reserve_discard(total + 1);
        E.tree_l.reserve_discard(total + 1);
        prof.begin(PC_TREE, st);
        T.a[0] = E.sa.p;
        T.l[0] = T.leaf ? nullptr : E.lcp.p;
        T.size[0] = N;
        T.nlev = 1;
        sz = N;
        size_t o = 0;
        while (sz > 1 && T.nlev < TREE_MAX_LEVELS) {
            uint32_t so = div_up<uint32_t>(sz, TREE_B);
            if (T.nlev == 1 && T.leaf)
                k_tree_level_leaf<<<div_up<uint32_t>(so, 256), 256, 0, st>>>(T.leaf, sz, E.tree_a.p + o, E.tree_l.p + o, so);
            else
                k_tree_level<<<div_up<uint32_t>(so, 256), 256, 0, st>>>(T.a[T.nlev - 1], T.l[T.nlev - 1], sz,
                                                                         E.tree_a.p + o, E.tree_l.p + o, so);
            L++;
            T.a[T.nlev] = E.tree_a.p + o;
            T.l[T.nlev] = E.tree_l.p + o;
            T.size[T.nlev] = so;
            T.nlev++;
            o += so;
            sz = so;
        }
        prof.end(st, 8.5 * N, T.nlev - 1);
    }

    // ---- M / reach, flags ----
    E.reach.reserve_discard(N + 1);
    E.flagp.reserve_discard(N + 2);
    E.flagc.reserve_discard(N + 2);
    E.lastnon.reserve_discard(N + 1);
    E.prevp.reserve_discard(N + 1);
    E.nextp.reserve_discard(N + 1);
    E.off.reserve_discard(N + 2);
    PX_CUDA(cudaMemsetAsync(E.flagp.p + s0, 0, (size_t) M + 2, st));
    uint32_t gridM = div_up<uint32_t>(M, 256);
    prof.begin(PC_LPF, st);
    if (cfg.rotate_policy == PIXIU_ROTATE_REFERENCE) {
        E.leafmask.reserve_discard((size_t) gridM * 8);
        E.splitmask.reserve_discard((size_t) gridM * 8);
        k_lpf<true><<<gridM, 256, 0, st>>>(T, w_text.p, E.rank.p, w_dist.p, s0, N, E.reach.p, E.leafmask.p, E.splitmask.p,
                                           fuse_flags ? E.flagp.p : nullptr);
    } else {
        k_lpf<false><<<gridM, 256, 0, st>>>(T, w_text.p, E.rank.p, w_dist.p, s0, N, E.reach.p, nullptr, nullptr,
                                            fuse_flags ? E.flagp.p : nullptr);
    }
    ep_flags_done = fuse_flags;
    prof.end(st, 18.0 * M, 1);
    L++;
    // reference rotation rule: the arena replay needs the node masks on the host; they are copied
    // asynchronously and replayed while phase B already runs (apply_rotation_cut)
    if (cfg.rotate_policy == PIXIU_ROTATE_REFERENCE) count_nodes_enqueue(s0, N);
    (void) R;
    E.tree = T;
    ep_first_new = first_new;
    ep_s0 = s0;
    ep_N = N;
    ep_n_new = n_new;
    launches += L;
}

// Phase B: PASS/COMPRESS flags from reach[], escape-pair rule, runs, output offsets.
void Store::enc_phase_b() {
    EncodeScratch &E = es;
    const uint32_t s0 = ep_s0, N = ep_N, M = N - s0;
    const uint32_t gridM = div_up<uint32_t>(M, 256);
    int L = 0;
    prof.begin(PC_FLAGS, st);
    if (!ep_flags_done) {   // (multi-GPU: reach has just been MAX-reduced over the shards)
        k_flag_scatter<<<gridM, 256, 0, st>>>(E.reach.p, w_dist.p, s0, N, E.flagp.p);
        L++;
    }
    if (win_long251) {
        // lastnon[i] = index of the last non-251 byte at or before i (max-scan of index+1, stored -1).
        // Position s0-1 is a separator (or the text start), i.e. "non-251": seed element 0 with it.
        const uint8_t *text = w_text.p;
        uint32_t *ln = E.lastnon.p;
        device_scan<uint32_t>(
            M,
            [=] __device__(size_t k) -> uint32_t {
                return text[s0 + k] != 251 ? (uint32_t) (s0 + k) + 1u : (k == 0 ? s0 : 0u);
            },
            [=] __device__(size_t k, uint32_t v) { ln[s0 + k] = v - 1u; }, OpMax(), 0u, false, E.scanws, st);
        L += 1;
    }
    {
        uint8_t *fcw = E.flagc.p;
        const uint8_t *fc = E.flagc.p;
        const uint8_t *text = w_text.p, *fp = E.flagp.p;
        const uint16_t *dist = w_dist.p;
        const uint32_t *ln = win_long251 ? E.lastnon.p : nullptr;   // (nullptr: pair_rule walks the short runs itself)
        uint32_t *pp = E.prevp.p, *np = E.nextp.p;
        // prevp[i] = 1 + index of the last non-COMPRESS position at or before i  (run start if i is COMPRESS); the
        // input functor also applies the pair rule and leaves flagc for the passes that follow
        device_scan<uint32_t>(
            M,
            [=] __device__(size_t k) -> uint32_t {
                const uint8_t c = pair_rule(text, dist, fp, ln, s0, N, (uint32_t) (s0 + k));
                fcw[s0 + k] = c;
                return c ? (k == 0 ? s0 : 0u) : (uint32_t) (s0 + k) + 1u;
            },
            [=] __device__(size_t k, uint32_t v) { pp[s0 + k] = v; }, OpMax(), 0u, false, E.scanws, st);
        // nextp[i] = index of the first non-COMPRESS position at or after i (suffix min-scan, reversed index)
        const uint32_t last = N - 1;
        device_scan<uint32_t>(
            M, [=] __device__(size_t k) -> uint32_t { return fc[last - k] ? 0xFFFFFFFFu : (uint32_t) (last - k); },
            [=] __device__(size_t k, uint32_t v) { np[last - k] = v; }, OpMin(), 0xFFFFFFFFu, false, E.scanws, st);
        L += 2;
    }
    {
        const uint8_t *fc = E.flagc.p;
        const uint16_t *dist = w_dist.p;
        const uint32_t *pp = E.prevp.p, *np = E.nextp.p;
        uint32_t *off = E.off.p;
        const int strict = cfg.strict251;
        // exclusive scan of the per-position output sizes; off[N] = total (one extra slot)
        device_scan<uint32_t>(
            (size_t) M + 1,
            [=] __device__(size_t k) -> uint32_t { return k < M ? contrib_at(s0 + (uint32_t) k, fc, dist, pp, np, strict) : 0u; },
            [=] __device__(size_t k, uint32_t v) { off[s0 + k] = v; }, OpSum(), 0u, true, E.scanws, st);
        L += 1;
    }
    prof.end(st, 45.0 * M, L);
    launches += L;
}

// Phase C: grow the store, emit the encoded records and their tables.  `cand`/`runidx` (multi-GPU mode)
// carry the globally reduced (idx << 16 | to) of every long run; `gidx` maps window records to chunk indices.
uint32_t Store::enc_phase_c(const uint32_t *cand, const uint32_t *runidx, const uint16_t *gidx) {
    EncodeScratch &E = es;
    const MinTree &T = E.tree;
    const uint32_t first_new = ep_first_new, s0 = ep_s0, N = ep_N, n_new = ep_n_new, M = N - s0;
    const uint32_t gridM = div_up<uint32_t>(M, 256);
    int L = 0;
    uint32_t enc_total = 0;
    PX_CUDA(cudaMemcpyAsync(&enc_total, E.off.p + N, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    PX_CUDA(cudaStreamSynchronize(st));

    // ---- grow the store, emit ----
    const size_t g_first = n_records();
    uint64_t new_tiles = 0;
    std::vector<uint32_t> tile_base(n_new);
    for (uint32_t r = 0; r < n_new; r++) {
        uint32_t dl = h_win_rec_start[first_new + r + 1] - h_win_rec_start[first_new + r] - 1;
        tile_base[r] = (uint32_t) (n_tiles + new_tiles);
        new_tiles += div_up<uint32_t>(dl, TILE);
    }
    grow_record_tables(g_first + n_new, enc_bytes + enc_total, n_tiles + new_tiles);
    PX_CUDA(cudaMemcpyAsync(d_tile_base.p + g_first, tile_base.data(), n_new * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    uint32_t g_chunk_first = chunk_first.back();
    if (!ep_emitted) {
        prof.begin(PC_EMIT, st);
        k_emit<<<gridM, 256, 0, st>>>(T, w_text.p, w_dist.p, w_recid.p, w_rec_start.p, E.rank.p, E.reach.p, E.flagc.p,
                                      E.prevp.p, E.nextp.p, E.off.p, s0, N, cfg.strict251, d_enc.ptr() + enc_bytes,
                                      E.counters.p + 2, cand, runidx, gidx);
        prof.end(st, 20.0 * M, 1);
        L++;
    }
    ep_emitted = false;
    prof.begin(PC_TABLES, st);
    k_record_tables<<<div_up<uint32_t>(n_new, 256), 256, 0, st>>>(n_new, first_new, (uint32_t) g_first, g_chunk_first,
                                                                  w_rec_start.p, E.off.p, 0u, enc_bytes, d_enc_off.p,
                                                                  d_enc_len.p, d_dec_len.p, d_first.p);
    if (new_tiles)
        k_tile_desc<<<div_up<uint32_t>((uint32_t) new_tiles, 256), 256, 0, st>>>(
            (uint32_t) new_tiles, (uint32_t) n_tiles, n_new, first_new, (uint32_t) g_first, d_tile_base.p,
            w_rec_start.p, E.off.p, E.flagc.p, E.prevp.p, E.nextp.p, w_text.p, win_long251 ? E.lastnon.p : nullptr, d_tile_desc.p);
    prof.end(st, 24.0 * n_new, 2);
    L += 2;
    // host mirrors
    size_t old = h_enc_len.size();
    h_enc_off.resize(old + n_new);
    h_enc_len.resize(old + n_new);
    h_dec_len.resize(old + n_new);
    h_first.resize(old + n_new, g_chunk_first);
    h_tile_base.insert(h_tile_base.end(), tile_base.begin(), tile_base.end());
    h_live.resize(old + n_new, 1);
    // the host mirrors of the new records and the error flags are fetched once per batch (flush_mirrors)
    enc_bytes += enc_total;
    n_tiles += new_tiles;
    chunk_count.back() += n_new;
    launches += L;
    return n_new;
}

// Reference rotation rule: keep only the candidate records that fit the arena budget.  Phase B ran on the
// whole candidate set; its per-position results for the accepted prefix do not depend on what follows
// (flags are local, the run scans stop at separators, the output offsets are a prefix sum).
void Store::apply_rotation_cut() {
    if (cfg.rotate_policy != PIXIU_ROTATE_REFERENCE) return;
    uint32_t acc = count_nodes_and_cut(ep_first_new, ep_s0, ep_N);
    if (acc < ep_n_new) {
        ep_n_new = acc;
        win_R = ep_first_new + acc;
        win_N = ep_N = h_win_rec_start[win_R];
        h_win_rec_start.resize(win_R + 1);
    }
}

// Host mirrors (encoded offset / length, decoded length) of the records emitted since the last flush, and the
// device-side error flags: one copy and one synchronisation per batch instead of per window.
void Store::flush_mirrors() {
    EncodeScratch &E = es;
    const size_t from = mirror_from, n = n_records() - from;
    uint32_t errflag = 0, scanerr = 0;
    if (n) {
        PX_CUDA(cudaMemcpyAsync(h_enc_off.data() + from, d_enc_off.p + from, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
        PX_CUDA(cudaMemcpyAsync(h_enc_len.data() + from, d_enc_len.p + from, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        PX_CUDA(cudaMemcpyAsync(h_dec_len.data() + from, d_dec_len.p + from, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    }
    if (E.counters.p) PX_CUDA(cudaMemcpyAsync(&errflag, E.counters.p + 2, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    if (E.scanws.ctl.p) PX_CUDA(cudaMemcpyAsync(&scanerr, E.scanws.ctl.p + 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    if (E.counters.p) PX_CUDA(cudaMemsetAsync(E.counters.p + 2, 0, sizeof(uint32_t), st));
    PX_CUDA(cudaStreamSynchronize(st));
    mirror_from = n_records();
    if (scanerr) throw std::runtime_error("encode: scan look-back timed out");
    if (errflag) throw std::runtime_error("encode: internal inconsistency (err=" + std::to_string(errflag) + ")");
}

// The encoded bytes of ALL candidate records are emitted before the rotation cut is known: the cut only ever drops
// the last few candidates (their bytes land beyond the committed end of the arena and are overwritten by the next
// window), and k_emit then runs while the host replays the reference's arena allocations instead of after it.
void Store::enc_emit_all() {
    EncodeScratch &E = es;
    const uint32_t s0 = ep_s0, N = ep_N, M = N - s0;
    d_enc.ensure(enc_bytes + (uint64_t) M + 4096);  // an encoded record is never longer than its document
    prof.begin(PC_EMIT, st);
    k_emit<<<div_up<uint32_t>(M, 256), 256, 0, st>>>(E.tree, w_text.p, w_dist.p, w_recid.p, w_rec_start.p, E.rank.p, E.reach.p,
                                                     E.flagc.p, E.prevp.p, E.nextp.p, E.off.p, s0, N, cfg.strict251,
                                                     d_enc.ptr() + enc_bytes, E.counters.p + 2, nullptr, nullptr, nullptr);
    prof.end(st, 20.0 * M, 1);
    launches++;
    ep_emitted = true;
}

uint32_t Store::encode_window_records(uint32_t first_new) {
    enc_phase_a(first_new, true);
    enc_phase_b();
    if (!knobs.no_spec_emit) enc_emit_all();  // (knob: A/B measurement)
    apply_rotation_cut();
    return enc_phase_c(nullptr, nullptr, nullptr);
}

// index maintenance (host CritBit; in-order semantics of n sequential setitem calls) + rc / saved outputs
void Store::finish_index(uint32_t nn, size_t g_batch_first, const uint8_t *d_keys, const int64_t *d_koff, const uint8_t *h_keys,
                         const int64_t *h_koff, const int64_t *h_voff, const uint32_t *h_doc_len, int32_t *rc, int32_t *saved) {
    std::vector<int64_t> old(nn);
    index->insert_batch(*this, nn, d_keys, d_koff, h_keys, h_koff, (uint32_t) g_batch_first, old.data());
    for (uint32_t i = 0; i < nn; i++) {
        const uint32_t g = (uint32_t) (g_batch_first + i);
        if (old[i] >= 0) tombstone((uint32_t) old[i]);
        note_live(g);
        if (rc) rc[i] = old[i] >= 0 ? PIXIU_CBT_SET_REPLACE : 0;
        if (saved) saved[i] = (int32_t) h_doc_len[i] - (int32_t) h_enc_len[g];
        raw_bytes += (h_koff[i + 1] - h_koff[i]) + (h_voff[i + 1] - h_voff[i]);
        doc_bytes += h_doc_len[i];
    }
}

int Store::setitem_batch(int64_t n, const uint8_t *d_keys, const int64_t *d_koff, const uint8_t *d_vals,
                         const int64_t *d_voff, const uint8_t *h_keys, const int64_t *h_koff,
                         const int64_t *h_voff, int32_t *rc, int32_t *saved) {
    if (n == 0) return PIXIU_OK;
    if (n > 0x7fffffff) return PIXIU_EINVAL;
    const uint32_t nn = (uint32_t) n;
    const auto t_begin = std::chrono::steady_clock::now();
    PX_CUDA(cudaEventRecord(ev0, st));
    doc_len.reserve_discard(nn);
    if (!es.counters.p) {
        es.counters.reserve_discard(24);
        PX_CUDA(cudaMemsetAsync(es.counters.p, 0, 24 * sizeof(uint32_t), st));
    }
    PX_CUDA(cudaMemsetAsync(es.counters.p + 8, 0, 9 * sizeof(uint32_t), st));
    prof.begin(PC_DOCS, st);
    k_doc_len<<<(unsigned) div_up<uint64_t>((uint64_t) nn * 32u, 256), 256, 0, st>>>(nn, d_keys, d_koff, d_vals, d_voff, doc_len.p,
                                                                                    es.counters.p + 8);
    prof.end(st, 0.0, 1);
    launches++;
    std::vector<uint32_t> h_doc_len(nn);
    PX_CUDA(cudaMemcpyAsync(h_doc_len.data(), doc_len.p, nn * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    PX_CUDA(cudaMemcpyAsync(batch_present, es.counters.p + 8, 9 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    PX_CUDA(cudaStreamSynchronize(st));
    for (uint32_t i = 0; i < nn; i++) {
        if (h_doc_len[i] == 0xFFFFFFFFu) {
            err = "setitem: empty key at record " + std::to_string(i);
            return PIXIU_EINVAL;
        }
        if (h_doc_len[i] > MAX_DOC) {
            err = "setitem: record " + std::to_string(i) + " longer than 65535 escaped bytes";
            return PIXIU_ETOOLONG;
        }
    }
    // the sort's look-back words carry 30-bit counts: a window never exceeds 2^30 - 2^17 positions
    const int64_t hard_cap = (1ll << 30) - (1ll << 17);
    const int64_t budget = cfg.rotate_policy == PIXIU_ROTATE_RECORDS ? hard_cap : std::min<int64_t>(cfg.window_bytes, hard_cap);
    const size_t g_batch_first = n_records();
    const bool ref_policy = cfg.rotate_policy == PIXIU_ROTATE_REFERENCE;
    uint32_t a = 0;
    dirty = true;  // from here on the window, the record tables and the index change: a failure poisons the store
    while (a < nn) {
        if (win_open) {
            bool full = win_R >= MAX_CHUNK_RECS;
            if (ref_policy) full = full || pool_nth >= 2048;                       // PiXiuCtrl.cpp:13
            else full = full || (int64_t) win_N + h_doc_len[a] + 1 > budget;
            if (full) close_window();
        }
        if (!win_open) open_window();
        for (int w8 = 0; w8 < 8; w8++) win_present[w8] |= batch_present[w8];
        win_long251 = knobs.lastnon_mode == 2 ? false : (win_long251 || batch_present[8] != 0 || knobs.lastnon_mode == 1);
        int64_t lim = budget;
        if (ref_policy) {
            // candidates: what the remaining arena is expected to hold (+4% and one record); the exact cut is
            // found by count_nodes_and_cut.  At least a quarter of the window so re-sorting stays geometric.
            double blocks_left = (2048.0 - pool_nth) * 65535.0 + (65535.0 - pool_used);
            // slack over the estimate: candidates the cut then drops were sorted for nothing, a shortfall costs a second
            // pass over the window.  1 % .. 4 %, three times the recent error of the estimate, plus one large record
            const double slack = 1.0 + std::min(0.04, std::max(0.01, 3.0 * rho_err));
            double est = blocks_left / (8.0 * rho) * slack + 70000.0;
            lim = std::min<int64_t>(hard_cap, (int64_t) win_N + (int64_t) std::max<double>(est, (double) win_N / 4));
        }
        // records [a, b) go into the open window
        uint32_t b = a;
        uint64_t bytes = win_N;
        std::vector<uint32_t> &rs = h_win_rec_start;
        while (b < nn && win_R + (b - a) < MAX_CHUNK_RECS &&
               (b == a ? true : (int64_t) (bytes + h_doc_len[b] + 1) <= lim)) {
            bytes += h_doc_len[b] + 1;
            rs.push_back((uint32_t) bytes);
            b++;
        }
        const uint32_t first_new = win_R, n_new = b - a;
        const uint32_t newN = (uint32_t) bytes;
        w_text.reserve_keep(newN + 16, win_N, st);
        w_dist.reserve_keep(newN + 16, win_N, st);
        w_recid.reserve_keep(newN + 16, win_N, st);
        w_rec_start.reserve_discard(rs.size() + 1);
        PX_CUDA(cudaMemcpyAsync(w_rec_start.p, rs.data(), rs.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        prof.begin(PC_DOCS, st);
        k_write_docs<<<n_new, 256, 0, st>>>(n_new, a, first_new, d_keys, d_koff, d_vals, d_voff,
                                                                         w_rec_start.p, w_text.p, w_dist.p, w_recid.p, nullptr);
        prof.end(st, 7.0 * (newN - win_N), 1);
        launches++;
        win_R += n_new;
        win_N = newN;
        uint32_t acc = encode_window_records(first_new);
        a += acc;
        if (acc < n_new) close_window();  // the arena budget was reached inside the candidates: rotate
    }
    flush_mirrors();
    const auto t_gpu_done = std::chrono::steady_clock::now();
    PX_CUDA(cudaEventRecord(ev1, st));
    finish_index(nn, g_batch_first, d_keys, d_koff, h_keys, h_koff, h_voff, h_doc_len.data(), rc, saved);
    PX_CUDA(cudaEventSynchronize(ev1));
    float ms = 0;
    PX_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
    last_set_ms = ms;
    prof.collect();
    if (knobs.trace) {
        const auto t_end = std::chrono::steady_clock::now();
        fprintf(stderr, "[setitem] n=%u encode(wall)=%.3f ms index=%.3f ms gpu(events)=%.3f ms\n", nn,
                std::chrono::duration<double, std::milli>(t_gpu_done - t_begin).count(),
                std::chrono::duration<double, std::milli>(t_end - t_gpu_done).count(), ms);
    }
    return PIXIU_OK;
}

// ---------------------------------------------------------------------------------
// Multi-GPU extended window: the open window is sharded by record over `world` stores (one per GPU,
// record idx -> rank idx % world); an incoming batch is replicated.  Every rank builds the suffix
// array of (its shard + the batch) and computes its local M(s); the global M is the element-wise MAX
// over ranks (collective 1), from which every rank derives identical flags and runs; each rank's
// leftmost-occurrence candidate (idx << 16 | to) per long run is MIN-reduced (collective 2: lowest
// record, then lowest offset = leftmost in insertion order); every rank then emits the identical
// encoded batch.  The collectives themselves are issued by the caller (torch.distributed / NCCL) on
// the device buffers these phases hand out.
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_extract_m(const uint32_t *__restrict__ reach, uint32_t s0, uint32_t m, uint32_t *__restrict__ out) {
    uint32_t k = blockIdx.x * 256 + threadIdx.x;
    if (k < m) out[k] = reach[s0 + k] - (s0 + k);
}

__global__ void __launch_bounds__(256)
k_apply_m(const uint32_t *__restrict__ in, uint32_t s0, uint32_t m, uint32_t *__restrict__ reach) {
    uint32_t k = blockIdx.x * 256 + threadIdx.x;
    if (k < m) reach[s0 + k] = s0 + k + in[k];
}

int Store::mg_begin(int64_t n, const uint8_t *d_keys, const int64_t *d_koff, const uint8_t *d_vals, const int64_t *d_voff,
                    const uint8_t *h_keys, const int64_t *h_koff, const int64_t *h_voff, uint32_t **d_m, int64_t *count,
                    bool sync) {
    if (mg_world < 1 || mg_pending) return PIXIU_EINVAL;
    if (n <= 0 || n > (int64_t) MAX_CHUNK_RECS) return PIXIU_EINVAL;  // a batch must fit one chunk
    if (cfg.rotate_policy == PIXIU_ROTATE_REFERENCE) {
        err = "multi-GPU mode extends the window beyond the reference's arena rule: use PIXIU_ROTATE_BYTES or _RECORDS";
        return PIXIU_EINVAL;
    }
    const uint32_t nn = (uint32_t) n;
    PX_CUDA(cudaEventRecord(ev0, st));
    doc_len.reserve_discard(nn);
    if (!es.counters.p) {
        es.counters.reserve_discard(24);
        PX_CUDA(cudaMemsetAsync(es.counters.p, 0, 24 * sizeof(uint32_t), st));
    }
    PX_CUDA(cudaMemsetAsync(es.counters.p + 8, 0, 9 * sizeof(uint32_t), st));
    k_doc_len<<<(unsigned) div_up<uint64_t>((uint64_t) nn * 32u, 256), 256, 0, st>>>(nn, d_keys, d_koff, d_vals, d_voff, doc_len.p,
                                                                                    es.counters.p + 8);
    launches++;
    mg_doc_len.resize(nn);
    PX_CUDA(cudaMemcpyAsync(mg_doc_len.data(), doc_len.p, nn * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    PX_CUDA(cudaMemcpyAsync(batch_present, es.counters.p + 8, 9 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    PX_CUDA(cudaStreamSynchronize(st));
    uint64_t batch_bytes = 0;
    for (uint32_t i = 0; i < nn; i++) {
        if (mg_doc_len[i] == 0xFFFFFFFFu) return PIXIU_EINVAL;
        if (mg_doc_len[i] > MAX_DOC) return PIXIU_ETOOLONG;
        batch_bytes += mg_doc_len[i] + 1;
    }
    // rotation is decided from global counters only, so every rank takes the same decision
    const uint64_t byte_budget = cfg.rotate_policy == PIXIU_ROTATE_BYTES ? (uint64_t) cfg.window_bytes * mg_world : ~0ull;
    const bool rotate = win_open && (mg_gR + nn > MAX_CHUNK_RECS || (mg_gbytes && mg_gbytes + batch_bytes > byte_budget));
    // the sort's look-back words carry 30-bit counts (checked before anything changes)
    if ((rotate || !win_open ? 0ull : (uint64_t) win_N) + batch_bytes >= (1ull << 30) - (1ull << 17)) {
        err = "multi-GPU setitem: shard + batch exceed 2^30 window positions";
        return PIXIU_EINVAL;
    }
    dirty = true;
    if (rotate) close_window();
    if (!win_open) {
        open_window();
        mg_gR = 0;
        mg_gbytes = 0;
        mg_h_gidx.clear();
    }
    for (int w8 = 0; w8 < 8; w8++) win_present[w8] |= batch_present[w8];
    win_long251 = knobs.lastnon_mode == 2 ? false : (win_long251 || batch_present[8] != 0 || knobs.lastnon_mode == 1);
    // local window = shard records + the whole batch
    const uint32_t first_new = win_R;
    uint64_t bytes = win_N;
    for (uint32_t i = 0; i < nn; i++) {
        bytes += mg_doc_len[i] + 1;
        h_win_rec_start.push_back((uint32_t) bytes);
        mg_h_gidx.push_back((uint16_t) (mg_gR + i));
    }
    const uint32_t newN = (uint32_t) bytes;
    w_text.reserve_keep(newN + 16, win_N, st);
    w_dist.reserve_keep(newN + 16, win_N, st);
    w_recid.reserve_keep(newN + 16, win_N, st);
    w_rec_start.reserve_discard(h_win_rec_start.size() + 1);
    es.gidx.reserve_discard(mg_h_gidx.size() + 1);
    PX_CUDA(cudaMemcpyAsync(w_rec_start.p, h_win_rec_start.data(), h_win_rec_start.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    PX_CUDA(cudaMemcpyAsync(es.gidx.p, mg_h_gidx.data(), mg_h_gidx.size() * sizeof(uint16_t), cudaMemcpyHostToDevice, st));
    k_write_docs<<<nn, 256, 0, st>>>(nn, 0, first_new, d_keys, d_koff, d_vals, d_voff, w_rec_start.p, w_text.p, w_dist.p,
                                     w_recid.p, nullptr);
    launches++;
    win_R += nn;
    win_N = newN;
    enc_phase_a(first_new);
    const uint32_t m = ep_N - ep_s0;
    es.mg_m.reserve_discard(m + 1);
    k_extract_m<<<div_up<uint32_t>(m, 256), 256, 0, st>>>(es.reach.p, ep_s0, m, es.mg_m.p);
    launches++;
    if (sync) PX_CUDA(cudaStreamSynchronize(st));
    // keep what the later phases need
    mg_d_keys = d_keys;
    mg_d_koff = d_koff;
    mg_d_vals = d_vals;
    mg_d_voff = d_voff;
    mg_h_keys.assign(h_keys + h_koff[0], h_keys + h_koff[nn]);
    mg_h_koff.resize(nn + 1);
    mg_h_voff.resize(nn + 1);
    for (uint32_t i = 0; i <= nn; i++) {
        mg_h_koff[i] = h_koff[i] - h_koff[0];
        mg_h_voff[i] = h_voff[i] - h_voff[0];
    }
    mg_pending = 1;
    mg_batch_bytes = batch_bytes;
    *d_m = es.mg_m.p;
    *count = m;
    return PIXIU_OK;
}

int Store::mg_mid(uint32_t **d_cand, int64_t *count, bool sync) {
    if (mg_pending != 1) return PIXIU_EINVAL;
    EncodeScratch &E = es;
    const uint32_t s0 = ep_s0, N = ep_N, m = N - s0;
    k_apply_m<<<div_up<uint32_t>(m, 256), 256, 0, st>>>(E.mg_m.p, s0, m, E.reach.p);
    launches++;
    enc_phase_b();
    // number the long runs: runidx[i] = long runs ending before i
    E.runidx.reserve_discard((size_t) N + 2);
    {
        const uint8_t *fc = E.flagc.p;
        const uint16_t *dist = w_dist.p;
        const uint32_t *pp = E.prevp.p, *np = E.nextp.p;
        uint32_t *ri = E.runidx.p;
        const int strict = cfg.strict251;
        device_scan<uint32_t>(
            (size_t) m + 1,
            [=] __device__(size_t k) -> uint32_t { return k < m && contrib_at(s0 + (uint32_t) k, fc, dist, pp, np, strict) > 1 ? 1u : 0u; },
            [=] __device__(size_t k, uint32_t v) { ri[s0 + k] = v; }, OpSum(), 0u, true, E.scanws, st);
        launches++;
    }
    uint32_t nruns = 0;
    PX_CUDA(cudaMemcpyAsync(&nruns, E.runidx.p + N, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    PX_CUDA(cudaStreamSynchronize(st));
    E.mg_cand.reserve_discard((size_t) nruns + 1);
    if (nruns) {
        k_candidates<<<div_up<uint32_t>(m, 256), 256, 0, st>>>(E.tree, w_dist.p, w_recid.p, w_rec_start.p, E.rank.p, E.reach.p,
                                                             E.flagc.p, E.prevp.p, E.nextp.p, E.runidx.p, E.gidx.p, s0, N,
                                                             cfg.strict251, E.mg_cand.p);
        launches++;
    }
    if (sync) PX_CUDA(cudaStreamSynchronize(st));
    mg_pending = 2;
    *d_cand = E.mg_cand.p;
    *count = nruns;
    return PIXIU_OK;
}

int Store::mg_end(int32_t *rc, int32_t *saved) {
    if (mg_pending != 2) return PIXIU_EINVAL;
    EncodeScratch &E = es;
    const uint32_t first_new = ep_first_new, nn = ep_n_new;
    const size_t g_batch_first = n_records();
    enc_phase_c(E.mg_cand.p, E.runidx.p, E.gidx.p);
    // shard maintenance: drop the batch from the local window, keep only the records this rank owns
    std::vector<uint32_t> own;
    for (uint32_t b = 0; b < nn; b++)
        if ((mg_gR + b) % (uint32_t) mg_world == (uint32_t) mg_rank) own.push_back(b);
    h_win_rec_start.resize(first_new + 1);
    mg_h_gidx.resize(first_new);
    uint64_t bytes = h_win_rec_start.back();
    for (uint32_t b : own) {
        bytes += mg_doc_len[b] + 1;
        h_win_rec_start.push_back((uint32_t) bytes);
        mg_h_gidx.push_back((uint16_t) (mg_gR + b));
    }
    win_R = first_new + (uint32_t) own.size();
    win_N = (uint32_t) bytes;
    if (!own.empty()) {
        DevBuf<uint32_t> &lst = E.mg_m;  // free after the reduce: reuse as the source list
        lst.reserve_discard(own.size() + 1);
        PX_CUDA(cudaMemcpyAsync(lst.p, own.data(), own.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        PX_CUDA(cudaMemcpyAsync(w_rec_start.p, h_win_rec_start.data(), h_win_rec_start.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        k_write_docs<<<(unsigned) own.size(), 256, 0, st>>>((uint32_t) own.size(), 0, first_new, mg_d_keys, mg_d_koff, mg_d_vals,
                                                          mg_d_voff, w_rec_start.p, w_text.p, w_dist.p, w_recid.p, lst.p);
        launches++;
    }
    flush_mirrors();
    PX_CUDA(cudaEventRecord(ev1, st));
    mg_gR += nn;
    mg_gbytes += mg_batch_bytes;
    finish_index(nn, g_batch_first, mg_d_keys, mg_d_koff, mg_h_keys.data(), mg_h_koff.data(), mg_h_voff.data(), mg_doc_len.data(), rc, saved);
    PX_CUDA(cudaEventSynchronize(ev1));
    float ms = 0;
    PX_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
    last_set_ms = ms;
    prof.collect();
    mg_pending = 0;
    return PIXIU_OK;
}

}  // namespace pixiu
