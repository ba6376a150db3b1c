// Device-wide scans (sum / max / min) with fused input and output transforms — ONE launch.
//
// Single-pass "decoupled look-back": CTAs take tiles of 256 threads x 8 items by ticket, publish
// their tile aggregate in a 64-bit status word (2 flag bits + value), and warp 0 resolves the
// exclusive prefix by looking back over up to 32 predecessor tiles at a time.  Input and output are
// functors, so callers fuse flag tests, index reversal (suffix scans) and scatters into the scan.
// HBM traffic: input read once, output written once.  Values must fit 62 bits.
#pragma once
#include "common.cuh"

namespace pixiu {

constexpr int SCAN_THREADS = 256;
#ifndef PIXIU_SCAN_ITEMS
#define PIXIU_SCAN_ITEMS 8
#endif
constexpr int SCAN_ITEMS = PIXIU_SCAN_ITEMS;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;
// the dual scan of the suffix-array rounds scatters (rank[sa] = ...) from its output functor: fewer items per thread
// (more tiles in flight, fewer registers) hide that latency better - measured 8 -> 4: 25.7 -> 22.4 ms per setitem step
#ifndef PIXIU_SCAN_DUAL_ITEMS
#define PIXIU_SCAN_DUAL_ITEMS 4
#endif
constexpr int SCAN_DUAL_ITEMS = PIXIU_SCAN_DUAL_ITEMS;
constexpr int SCAN_DUAL_TILE = SCAN_THREADS * SCAN_DUAL_ITEMS;
constexpr unsigned long long SCAN_FLAG_AGG = 1ull << 62;
constexpr unsigned long long SCAN_FLAG_PREFIX = 2ull << 62;
constexpr unsigned long long SCAN_VALUE_MASK = (1ull << 62) - 1;
constexpr uint32_t SCAN_SPIN_LIMIT = 1u << 27;

struct OpSum {
    template <typename T>
    __host__ __device__ __forceinline__ T operator()(T a, T b) const { return a + b; }
};
struct OpMax {
    template <typename T>
    __host__ __device__ __forceinline__ T operator()(T a, T b) const { return a > b ? a : b; }
};
struct OpMin {
    template <typename T>
    __host__ __device__ __forceinline__ T operator()(T a, T b) const { return a < b ? a : b; }
};

#ifdef __CUDACC__
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

template <typename T, typename Op>
__device__ __forceinline__ T warp_scan_inclusive(T v, Op op) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        T o = __shfl_up_sync(0xffffffffu, v, d);
        if ((int) lane_id() >= d) v = op(o, v);
    }
    return v;
}

// exclusive scan of one value per thread over the CTA; smem must hold 33 T's.
// Returns the exclusive prefix; *total = reduction over the whole CTA.
template <typename T, typename Op>
__device__ __forceinline__ T block_scan_exclusive(T v, Op op, T identity, T *smem, T *total) {
    const int warp = threadIdx.x >> 5, lane = lane_id(), nwarp = blockDim.x >> 5;
    T inc = warp_scan_inclusive(v, op);
    if (lane == 31) smem[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        T w = lane < nwarp ? smem[lane] : identity;
        T winc = warp_scan_inclusive(w, op);
        smem[lane] = winc;  // inclusive over warps
    }
    __syncthreads();
    T warp_prefix = warp ? smem[warp - 1] : identity;
    *total = smem[nwarp - 1];
    T exc = __shfl_up_sync(0xffffffffu, inc, 1);
    if (lane == 0) exc = identity;
    __syncthreads();
    return op(warp_prefix, exc);
}

template <typename T, typename Op, typename InFn, typename OutFn>
__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_lookback(size_t n, InFn in, OutFn out, Op op, T identity, int exclusive, unsigned long long *__restrict__ status,
                uint32_t *__restrict__ ticket, uint32_t *__restrict__ err) {
    __shared__ T sm[33];
    __shared__ T s_prefix;
    __shared__ uint32_t s_tile;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const size_t base = (size_t) tile * SCAN_TILE + (size_t) threadIdx.x * SCAN_ITEMS;
    T v[SCAN_ITEMS];
    T acc = identity;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        v[k] = base + k < n ? in(base + k) : identity;
        acc = op(acc, v[k]);
    }
    T total;
    T pre = block_scan_exclusive(acc, op, identity, sm, &total);
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        T excl = identity;
        if (tile == 0) {
            if (lane == 0) st_relaxed_u64(status, SCAN_FLAG_PREFIX | (unsigned long long) total);
        } else {
            if (lane == 0) st_relaxed_u64(status + tile, SCAN_FLAG_AGG | (unsigned long long) total);
            int64_t look = (int64_t) tile - 1 - lane;
            while (true) {
                unsigned long long sv = SCAN_FLAG_PREFIX | (unsigned long long) identity;  // before tile 0
                if (look >= 0) {
                    uint32_t spins = 0;
                    while (((sv = ld_relaxed_u64(status + look)) >> 62) == 0) {
                        if (++spins > SCAN_SPIN_LIMIT) {
                            atomicExch(err, 1u);
                            sv = SCAN_FLAG_PREFIX | (unsigned long long) identity;
                            break;
                        }
                    }
                }
                const bool is_prefix = (sv >> 62) == 2;
                const uint32_t pm = __ballot_sync(0xffffffffu, is_prefix);
                const int first = pm ? __ffs(pm) - 1 : 31;
                T contrib = lane <= first ? (T) (sv & SCAN_VALUE_MASK) : identity;
#pragma unroll
                for (int d = 16; d; d >>= 1) contrib = op(contrib, (T) __shfl_xor_sync(0xffffffffu, contrib, d));
                excl = op(contrib, excl);
                if (pm) break;
                look -= 32;
            }
            if (lane == 0) st_relaxed_u64(status + tile, SCAN_FLAG_PREFIX | (unsigned long long) op(excl, total));
        }
        if (lane == 0) s_prefix = excl;
    }
    __syncthreads();
    pre = op(s_prefix, pre);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        T inc = op(pre, v[k]);
        if (base + k < n) out(base + k, exclusive ? pre : inc);
        pre = inc;
    }
}

// Two scans over the same input pass: an inclusive MAX of a u32 and an exclusive SUM of a u64 (the round
// finish of the suffix-array construction needs both from the same predicate).  in(i, m, s) produces the
// two inputs, out(i, max_incl, sum_excl, m, s) consumes the results together with the element's own inputs.
template <typename InFn, typename OutFn>
__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_dual(size_t n, InFn in, OutFn out, unsigned long long *__restrict__ status_m, unsigned long long *__restrict__ status_s,
            uint32_t *__restrict__ ticket, uint32_t *__restrict__ err) {
    __shared__ uint32_t sm_m[33];
    __shared__ unsigned long long sm_s[33];
    __shared__ uint32_t s_pm;
    __shared__ unsigned long long s_ps;
    __shared__ uint32_t s_tile;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const size_t base = (size_t) tile * SCAN_DUAL_TILE + (size_t) threadIdx.x * SCAN_DUAL_ITEMS;
    uint32_t vm[SCAN_DUAL_ITEMS];
    unsigned long long vs[SCAN_DUAL_ITEMS];
    uint32_t am = 0;
    unsigned long long as = 0;
#pragma unroll
    for (int k = 0; k < SCAN_DUAL_ITEMS; k++) {
        vm[k] = 0;
        vs[k] = 0;
        if (base + k < n) in(base + k, vm[k], vs[k]);
        am = max(am, vm[k]);
        as += vs[k];
    }
    uint32_t tot_m;
    unsigned long long tot_s;
    uint32_t pre_m = block_scan_exclusive(am, OpMax(), 0u, sm_m, &tot_m);
    unsigned long long pre_s = block_scan_exclusive(as, OpSum(), 0ull, sm_s, &tot_s);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp < 2) {
        // warp 0 resolves the MAX prefix, warp 1 the SUM prefix (same look-back as k_scan_lookback)
        unsigned long long *status = warp == 0 ? status_m : status_s;
        const unsigned long long mine = warp == 0 ? (unsigned long long) tot_m : tot_s;
        unsigned long long excl = 0;
        if (tile == 0) {
            if (lane == 0) st_relaxed_u64(status, SCAN_FLAG_PREFIX | mine);
        } else {
            if (lane == 0) st_relaxed_u64(status + tile, SCAN_FLAG_AGG | mine);
            int64_t look = (int64_t) tile - 1 - lane;
            while (true) {
                unsigned long long sv = SCAN_FLAG_PREFIX;  // before tile 0: identity (0 for both scans)
                if (look >= 0) {
                    uint32_t spins = 0;
                    while (((sv = ld_relaxed_u64(status + look)) >> 62) == 0) {
                        if (++spins > SCAN_SPIN_LIMIT) {
                            atomicExch(err, 1u);
                            sv = SCAN_FLAG_PREFIX;
                            break;
                        }
                    }
                }
                const uint32_t pm = __ballot_sync(0xffffffffu, (sv >> 62) == 2);
                const int first = pm ? __ffs(pm) - 1 : 31;
                unsigned long long contrib = lane <= first ? (sv & SCAN_VALUE_MASK) : 0ull;
#pragma unroll
                for (int d = 16; d; d >>= 1) {
                    unsigned long long o = __shfl_xor_sync(0xffffffffu, contrib, d);
                    contrib = warp == 0 ? (contrib > o ? contrib : o) : contrib + o;
                }
                excl = warp == 0 ? (excl > contrib ? excl : contrib) : excl + contrib;
                if (pm) break;
                look -= 32;
            }
            const unsigned long long incl = warp == 0 ? (excl > mine ? excl : mine) : excl + mine;
            if (lane == 0) st_relaxed_u64(status + tile, SCAN_FLAG_PREFIX | incl);
        }
        if (lane == 0) {
            if (warp == 0) s_pm = (uint32_t) excl;
            else s_ps = excl;
        }
    }
    __syncthreads();
    pre_m = max(pre_m, s_pm);
    pre_s += s_ps;
#pragma unroll
    for (int k = 0; k < SCAN_DUAL_ITEMS; k++) {
        pre_m = max(pre_m, vm[k]);
        if (base + k < n) out(base + k, pre_m, pre_s, vm[k], vs[k]);
        pre_s += vs[k];
    }
}
#endif  // __CUDACC__

// workspace shared by all scans of a store (status words + ticket + sticky error flag)
struct ScanWorkspace {
    DevBuf<unsigned long long> status;  // [ticket word][status words ...]
    DevBuf<uint32_t> ctl;               // [1] sticky error flag
};

#ifdef __CUDACC__
template <typename InFn, typename OutFn>
void device_scan_dual(size_t n, InFn in, OutFn out, ScanWorkspace &ws, cudaStream_t st) {
    if (n == 0) return;
    unsigned tiles = (unsigned) div_up<size_t>(n, SCAN_DUAL_TILE);
    // one buffer, one memset: [ticket word][status of scan 1: tiles][status of scan 2: tiles]
    ws.status.reserve_discard(2 * (size_t) tiles + 2);
    if (!ws.ctl.p) {
        ws.ctl.reserve_discard(4);
        PX_CUDA(cudaMemsetAsync(ws.ctl.p, 0, 4 * sizeof(uint32_t), st));
    }
    PX_CUDA(cudaMemsetAsync(ws.status.p, 0, (2 * (size_t) tiles + 1) * sizeof(unsigned long long), st));
    k_scan_dual<<<tiles, SCAN_THREADS, 0, st>>>(n, in, out, ws.status.p + 1, ws.status.p + 1 + tiles,
                                                reinterpret_cast<uint32_t *>(ws.status.p), ws.ctl.p + 1);
    PX_LAUNCH_CHECK();
}

template <typename T, typename Op, typename InFn, typename OutFn>
void device_scan(size_t n, InFn in, OutFn out, Op op, T identity, bool exclusive, ScanWorkspace &ws, cudaStream_t st) {
    if (n == 0) return;
    unsigned tiles = (unsigned) div_up<size_t>(n, SCAN_TILE);
    // one buffer, one memset: [ticket word][status: tiles]
    ws.status.reserve_discard((size_t) tiles + 2);
    if (!ws.ctl.p) {
        ws.ctl.reserve_discard(4);
        PX_CUDA(cudaMemsetAsync(ws.ctl.p, 0, 4 * sizeof(uint32_t), st));
    }
    PX_CUDA(cudaMemsetAsync(ws.status.p, 0, ((size_t) tiles + 1) * sizeof(unsigned long long), st));
    k_scan_lookback<T><<<tiles, SCAN_THREADS, 0, st>>>(n, in, out, op, identity, exclusive ? 1 : 0, ws.status.p + 1,
                                                      reinterpret_cast<uint32_t *>(ws.status.p), ws.ctl.p + 1);
    PX_LAUNCH_CHECK();
}
#endif

}  // namespace pixiu
