// Hand-written LSD radix sort of (key, u32 value) pairs for sm_100a.
//
// One up-front histogram launch counts every 8-bit digit of every pass; each pass
// is then ONE launch ("onesweep"): a CTA takes a tile of 256 x 8 keys in ticket
// order, ranks them stably with warp match-any + shared-memory digit counters,
// resolves its global digit offsets by decoupled look-back over the tile-status
// table, and scatters keys and values.  Per pass each element is read once and
// written once (8|4 B key + 4 B value), the histogram launch reads the keys once.
//
// Used by the suffix-array construction (encode.cu): the initial 56-bit keys of 8 symbols (63-bit / 7 symbols for rich
// alphabets), the members of the groups too large for a CTA in a prefix-doubling round, and the (group, rank[i+h]) keys
// of a whole round when most suffixes sit in such groups; by the bulk build of the CritBit tree (index.cu).
#pragma once
#include "common.cuh"
#include "prof.h"

namespace pixiu {

constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
#ifndef PIXIU_RS_ITEMS
#define PIXIU_RS_ITEMS 8
#endif
#ifndef PIXIU_RS_MINB
#define PIXIU_RS_MINB 5
#endif
constexpr int RS_ITEMS = PIXIU_RS_ITEMS;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;  // keys per CTA
constexpr int RS_BINS = 256;
constexpr int RS_MAX_PASSES = 8;
constexpr uint32_t RS_FLAG_AGG = 1u << 30;
constexpr uint32_t RS_FLAG_PREFIX = 2u << 30;
constexpr uint32_t RS_VALUE_MASK = (1u << 30) - 1;
constexpr uint32_t RS_SPIN_LIMIT = 1u << 27;
#ifndef PIXIU_RS_LOOK
#define PIXIU_RS_LOOK 8
#endif
constexpr int RS_LOOK = PIXIU_RS_LOOK;

// hist[pass][bin] += count, for passes [0, npass) covering bits [begin_bit + 8*pass, ..)
template <typename KeyT>
__global__ void __launch_bounds__(RS_THREADS)
k_rs_histogram(const KeyT *__restrict__ keys, uint32_t n, int begin_bit, int npass, uint32_t *__restrict__ hist) {
    __shared__ uint32_t sh[RS_MAX_PASSES * RS_BINS];
    for (int i = threadIdx.x; i < npass * RS_BINS; i += RS_THREADS) sh[i] = 0;
    __syncthreads();
    const uint32_t stride = gridDim.x * RS_THREADS;
    for (uint32_t i = blockIdx.x * RS_THREADS + threadIdx.x; i < n; i += stride) {
        KeyT k = keys[i];
        for (int p = 0; p < npass; p++) atomicAdd(&sh[p * RS_BINS + (uint32_t) ((k >> (begin_bit + 8 * p)) & 0xff)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < npass * RS_BINS; i += RS_THREADS)
        if (sh[i]) atomicAdd(&hist[i], sh[i]);
}

// exclusive scan of each pass's 256 bins, in place: one CTA of 256 threads per pass
static __global__ void __launch_bounds__(RS_BINS) k_rs_scan_bins(uint32_t *hist) {
    __shared__ uint32_t sm[RS_BINS];
    uint32_t *h = hist + blockIdx.x * RS_BINS;
    uint32_t v = h[threadIdx.x];
    sm[threadIdx.x] = v;
    __syncthreads();
    for (int d = 1; d < RS_BINS; d <<= 1) {
        uint32_t o = threadIdx.x >= d ? sm[threadIdx.x - d] : 0;
        __syncthreads();
        sm[threadIdx.x] += o;
        __syncthreads();
    }
    h[threadIdx.x] = sm[threadIdx.x] - v;
}

// One onesweep pass.  status: [tiles][256] zero-initialised; ticket: zero-initialised counter.
// vals_in == nullptr means "value = input index" (first pass of an index sort).
// Dynamic shared memory: RS_TILE keys + RS_TILE values (the tile is re-ordered by digit in shared
// memory so that the global stores of one digit are consecutive: coalesced 128-byte runs on average).
template <typename KeyT>
__global__ void __launch_bounds__(RS_THREADS, PIXIU_RS_MINB)
k_rs_onesweep(const KeyT *__restrict__ keys_in, KeyT *__restrict__ keys_out, const uint32_t *__restrict__ vals_in,
              uint32_t *__restrict__ vals_out, uint32_t n, int shift, const uint32_t *__restrict__ bin_base,
              uint32_t *__restrict__ status, uint32_t *__restrict__ ticket, uint32_t *__restrict__ err) {
    extern __shared__ __align__(16) uint8_t rs_smem[];
    KeyT *keys_s = reinterpret_cast<KeyT *>(rs_smem);
    uint32_t *vals_s = reinterpret_cast<uint32_t *>(rs_smem + sizeof(KeyT) * RS_TILE);
    __shared__ uint32_t warp_hist[RS_WARPS][RS_BINS];  // per-warp digit counts, then exclusive warp prefixes
    __shared__ uint32_t digit_start[RS_BINS];          // first tile-local slot of each digit
    __shared__ uint32_t adj[RS_BINS];                  // global position of slot j of digit d = adj[d] + j
    __shared__ uint32_t warp_tot[RS_BINS / 32];
    __shared__ uint32_t s_tile;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    for (int i = threadIdx.x; i < RS_WARPS * RS_BINS; i += RS_THREADS) (&warp_hist[0][0])[i] = 0;
    __syncthreads();
    const uint32_t tile = s_tile;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t lt_mask = (1u << lane) - 1;
    const uint32_t wbase = tile * RS_TILE + warp * (32 * RS_ITEMS);
    const uint32_t tile_count = min((uint32_t) RS_TILE, n - tile * RS_TILE);

    KeyT key[RS_ITEMS];
    uint16_t off[RS_ITEMS];
#pragma unroll
    for (int k = 0; k < RS_ITEMS; k++) {
        uint32_t i = wbase + k * 32 + lane;
        key[k] = i < n ? keys_in[i] : (KeyT) ~(KeyT) 0;
    }
    // stable ranking: the warp walks its 512 consecutive keys 32 at a time
#pragma unroll
    for (int k = 0; k < RS_ITEMS; k++) {
        uint32_t i = wbase + k * 32 + lane;
        bool valid = i < n;
        uint32_t d = (uint32_t) ((key[k] >> shift) & 0xff);
        uint32_t active = __ballot_sync(0xffffffffu, valid);
        uint32_t m = __match_any_sync(0xffffffffu, valid ? d : 0x100u + (uint32_t) lane) & active;
        uint32_t c = valid ? warp_hist[warp][d] : 0;
        __syncwarp();
        if (valid && (m & lt_mask) == 0) warp_hist[warp][d] = c + __popc(m);
        __syncwarp();
        off[k] = (uint16_t) (c + __popc(m & lt_mask));
    }
    // the values are pulled into L2 while the digits are scanned and the look-back runs (no register held for them)
    if (vals_in && lane < RS_ITEMS) {
        uint32_t i = wbase + lane * 32;
        if (i < n) asm volatile("prefetch.global.L2 [%0];" ::"l"(vals_in + i));
    }
    __syncthreads();
    // digit `t` (threads 0..255): exclusive prefix over the warps, tile total, tile-local digit start,
    // decoupled look-back.  The other threads only take part in the barriers.
    {
        const int t = threadIdx.x;
        const bool dig = t < RS_BINS;
        uint32_t run = 0;
        if (dig) {
#pragma unroll
            for (int w = 0; w < RS_WARPS; w++) {
                uint32_t c = warp_hist[w][t];
                warp_hist[w][t] = run;
                run += c;
            }
        }
        // exclusive scan of `run` over the 256 digits
        uint32_t inc = run;
#pragma unroll
        for (int dd = 1; dd < 32; dd <<= 1) {
            uint32_t o = __shfl_up_sync(0xffffffffu, inc, dd);
            if (lane >= dd) inc += o;
        }
        if (dig && lane == 31) warp_tot[warp] = inc;
        __syncthreads();
        if (dig) {
            uint32_t wpre = 0;
#pragma unroll
            for (int w = 0; w < RS_BINS / 32; w++) wpre += (w < warp) ? warp_tot[w] : 0u;
            const uint32_t dstart = wpre + inc - run;
            digit_start[t] = dstart;

            uint32_t *mine = status + (size_t) tile * RS_BINS + t;
            uint32_t excl = 0;
            if (tile == 0) {
                st_relaxed_u32(mine, RS_FLAG_PREFIX | run);
            } else {
                st_relaxed_u32(mine, RS_FLAG_AGG | run);
                // look back RS_LOOK predecessors at a time (independent relaxed loads overlap): the first wave
                // of tiles runs in lock step and would otherwise walk hundreds of rows one round trip each
                int64_t look = (int64_t) tile - 1;
                uint32_t spins = 0;
                bool done = false;
                while (!done && look >= 0) {
                    uint32_t sv[RS_LOOK];
#pragma unroll
                    for (int b = 0; b < RS_LOOK; b++)
                        sv[b] = look - b >= 0 ? ld_relaxed_u32(status + (size_t) (look - b) * RS_BINS + t) : RS_FLAG_PREFIX;
                    int used = 0;
#pragma unroll
                    for (int b = 0; b < RS_LOOK; b++) {
                        if (done || used != b) continue;
                        uint32_t flag = sv[b] & ~RS_VALUE_MASK;
                        if (flag == 0) continue;  // not published yet: retry from here
                        excl += sv[b] & RS_VALUE_MASK;
                        used = b + 1;
                        if (flag == RS_FLAG_PREFIX) done = true;
                    }
                    look -= used;
                    if (used == 0 && ++spins > RS_SPIN_LIMIT) {
                        atomicExch(err, 1u);
                        break;
                    }
                }
                st_relaxed_u32(mine, RS_FLAG_PREFIX | (excl + run));
            }
            adj[t] = bin_base[t] + excl - dstart;
        }
    }
    __syncthreads();
    // re-order the tile by digit in shared memory
#pragma unroll
    for (int k = 0; k < RS_ITEMS; k++) {
        uint32_t i = wbase + k * 32 + lane;
        if (i < n) {
            uint32_t d = (uint32_t) ((key[k] >> shift) & 0xff);
            uint32_t lp = digit_start[d] + warp_hist[warp][d] + off[k];
            keys_s[lp] = key[k];
            vals_s[lp] = vals_in ? vals_in[i] : i;
        }
    }
    __syncthreads();
    for (uint32_t j = threadIdx.x; j < tile_count; j += RS_THREADS) {
        KeyT kk = keys_s[j];
        uint32_t pos = adj[(uint32_t) ((kk >> shift) & 0xff)] + j;
        keys_out[pos] = kk;
        vals_out[pos] = vals_s[j];
    }
}

template <typename KeyT>
constexpr size_t rs_dyn_smem() { return (sizeof(KeyT) + sizeof(uint32_t)) * RS_TILE; }

struct RadixSortTemp {
    DevBuf<uint32_t> status;   // [hist: 8 x 256][tickets: 8][status: passes x tiles x 256]
};

// Sorts n (key,value) pairs on bits [begin_bit, end_bit) ascending, stable.
// Ping-pongs between (k0,v0) and (k1,v1); returns 0 if the result is in (k0,v0), 1 otherwise.
// iota => the input values are taken to be the input indices (v0 is only used as a buffer).
template <typename KeyT>
int radix_sort_pairs(KeyT *k0, KeyT *k1, uint32_t *v0, uint32_t *v1, uint32_t n, int begin_bit, int end_bit,
                     bool iota, RadixSortTemp &tmp, uint32_t *d_err, cudaStream_t st, int *launches = nullptr,
                     Profiler *prof = nullptr) {
    if (n == 0) return 0;
    if (end_bit <= begin_bit) throw std::runtime_error("radix_sort_pairs: empty bit range");
    int npass = (end_bit - begin_bit + 7) / 8;
    if (npass > RS_MAX_PASSES) throw std::runtime_error("radix_sort_pairs: too many passes");
    uint32_t tiles = div_up<uint32_t>(n, RS_TILE);
    // one buffer, one memset: [hist: 8 x 256][tickets: 8][status: npass x tiles x 256]
    const size_t head_words = RS_MAX_PASSES * RS_BINS + RS_MAX_PASSES;
    const size_t total_words = head_words + (size_t) npass * tiles * RS_BINS;
    tmp.status.reserve_discard(total_words);
    PX_CUDA(cudaMemsetAsync(tmp.status.p, 0, total_words * sizeof(uint32_t), st));
    uint32_t *hist = tmp.status.p, *ticket = tmp.status.p + RS_MAX_PASSES * RS_BINS, *status = tmp.status.p + head_words;
    uint32_t hgrid = tiles < 148u * 8u ? tiles : 148u * 8u;
    if (prof) prof->begin(PC_SORT_HIST, st);
    k_rs_histogram<KeyT><<<hgrid, RS_THREADS, 0, st>>>(k0, n, begin_bit, npass, hist);
    k_rs_scan_bins<<<npass, RS_BINS, 0, st>>>(hist);
    if (prof) prof->end(st, (double) n * sizeof(KeyT), 2);
    PX_CUDA(cudaFuncSetAttribute(k_rs_onesweep<KeyT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) rs_dyn_smem<KeyT>()));
    int cur = 0;
    for (int p = 0; p < npass; p++) {
        KeyT *ki = cur ? k1 : k0, *ko = cur ? k0 : k1;
        uint32_t *vi = cur ? v1 : v0, *vo = cur ? v0 : v1;
        if (prof) prof->begin(PC_SORT_PASS, st);
        k_rs_onesweep<KeyT><<<tiles, RS_THREADS, rs_dyn_smem<KeyT>(), st>>>(ki, ko, (p == 0 && iota) ? nullptr : vi, vo, n,
                                                          begin_bit + 8 * p, hist + p * RS_BINS,
                                                          status + (size_t) p * tiles * RS_BINS, ticket + p, d_err);
        if (prof) prof->end(st, (double) n * (2.0 * sizeof(KeyT) + ((p == 0 && iota) ? 4.0 : 8.0)), 1);
        cur ^= 1;
    }
    PX_LAUNCH_CHECK();
    if (launches) *launches += 2 + npass;
    return cur;
}

}  // namespace pixiu
