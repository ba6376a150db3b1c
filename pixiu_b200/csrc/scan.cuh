// Device-wide scans (sum / max / min) with fused input and output transforms.
//
// Three launches (tile reduce -> scan of tile partials -> tile down-sweep), tiles of
// 256 threads x 8 items.  Input and output are functors so the callers fuse flag
// tests, index reversal (for suffix scans) and scattering into the scan itself.
// HBM traffic: input read twice, output written once.
#pragma once
#include "common.cuh"

namespace pixiu {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

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

template <typename T, typename Op, typename InFn>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_reduce(size_t n, InFn in, Op op, T identity, T *partials) {
    __shared__ T sm[33];
    size_t base = (size_t) blockIdx.x * SCAN_TILE + (size_t) threadIdx.x * SCAN_ITEMS;
    T acc = identity;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++)
        if (base + k < n) acc = op(acc, in(base + k));
    T total;
    block_scan_exclusive(acc, op, identity, sm, &total);
    if (threadIdx.x == 0) partials[blockIdx.x] = total;
}

template <typename T, typename Op>
__global__ void __launch_bounds__(1024) k_scan_partials(size_t m, Op op, T identity, T *partials) {
    __shared__ T sm[33];
    T carry = identity;
    for (size_t base = 0; base < m; base += 1024) {
        size_t i = base + threadIdx.x;
        T v = i < m ? partials[i] : identity;
        T total;
        T exc = block_scan_exclusive(v, op, identity, sm, &total);
        if (i < m) partials[i] = op(carry, exc);
        carry = op(carry, total);
        __syncthreads();
    }
}

template <typename T, typename Op, typename InFn, typename OutFn>
__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_down(size_t n, InFn in, OutFn out, Op op, T identity, const T *partials, int exclusive) {
    __shared__ T sm[33];
    size_t base = (size_t) blockIdx.x * SCAN_TILE + (size_t) threadIdx.x * SCAN_ITEMS;
    T v[SCAN_ITEMS];
    T acc = identity;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        v[k] = base + k < n ? in(base + k) : identity;
        acc = op(acc, v[k]);
    }
    T total;
    T pre = block_scan_exclusive(acc, op, identity, sm, &total);
    pre = op(partials[blockIdx.x], pre);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        T inc = op(pre, v[k]);
        if (base + k < n) out(base + k, exclusive ? pre : inc);
        pre = inc;
    }
}

struct OpSum {
    template <typename T>
    __device__ __forceinline__ T operator()(T a, T b) const { return a + b; }
};
struct OpMax {
    template <typename T>
    __device__ __forceinline__ T operator()(T a, T b) const { return a > b ? a : b; }
};
struct OpMin {
    template <typename T>
    __device__ __forceinline__ T operator()(T a, T b) const { return a < b ? a : b; }
};

// tmp must hold div_up(n, SCAN_TILE) elements of T
template <typename T, typename Op, typename InFn, typename OutFn>
void device_scan(size_t n, InFn in, OutFn out, Op op, T identity, bool exclusive, T *tmp, cudaStream_t st) {
    if (n == 0) return;
    unsigned tiles = (unsigned) div_up<size_t>(n, SCAN_TILE);
    k_scan_reduce<T><<<tiles, SCAN_THREADS, 0, st>>>(n, in, op, identity, tmp);
    k_scan_partials<T><<<1, 1024, 0, st>>>((size_t) tiles, op, identity, tmp);
    k_scan_down<T><<<tiles, SCAN_THREADS, 0, st>>>(n, in, out, op, identity, tmp, exclusive ? 1 : 0);
    PX_LAUNCH_CHECK();
}

inline size_t scan_tmp_elems(size_t n) { return div_up<size_t>(n, SCAN_TILE) + 1; }

}  // namespace pixiu
