// Common host/device helpers for the pixiu_b200 CUDA library (sm_100a only).
#pragma once
#include <chrono>
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <vector>

namespace pixiu {

struct CudaError : std::runtime_error {
    cudaError_t code;
    CudaError(cudaError_t c, const char *what_, const char *file, int line)
        : std::runtime_error(std::string(what_) + ": " + cudaGetErrorString(c) + " @" + file + ":" + std::to_string(line)),
          code(c) {}
};

#define PX_CUDA(expr)                                                        \
    do {                                                                     \
        cudaError_t _e = (expr);                                             \
        if (_e != cudaSuccess) throw ::pixiu::CudaError(_e, #expr, __FILE__, __LINE__); \
    } while (0)

#define PX_LAUNCH_CHECK() PX_CUDA(cudaGetLastError())

template <typename T>
static inline T div_up(T a, T b) { return (a + b - 1) / b; }

// PIXIU_TRACE, looked at once per process (allocation / phase trace lines on stderr)
static inline bool trace_on() {
    static const bool on = getenv("PIXIU_TRACE") != nullptr;
    return on;
}

// Growable device array (capacity doubling).  HBM is 180 GB: slack is cheap, realloc is not.
template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t cap = 0;   // elements
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { release(); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    // ensure capacity >= n elements; contents are NOT preserved
    void reserve_discard(size_t n) {
        if (n <= cap) return;
        release();
        size_t want = n + n / 4 + 256;
        if (trace_on()) fprintf(stderr, "[mem] alloc %zu bytes\n", want * sizeof(T));
        PX_CUDA(cudaMalloc(&p, want * sizeof(T)));
        cap = want;
    }
    // ensure capacity >= n elements, preserving the first `keep` elements
    void reserve_keep(size_t n, size_t keep, cudaStream_t st) {
        if (n <= cap) return;
        size_t want = std::max<size_t>(2 * cap, n + n / 2 + 256);
        T *q = nullptr;
        if (trace_on()) fprintf(stderr, "[mem] grow %zu -> %zu bytes (keep %zu)\n", cap * sizeof(T), want * sizeof(T), keep * sizeof(T));
        const auto t0 = std::chrono::steady_clock::now();
        PX_CUDA(cudaMalloc(&q, want * sizeof(T)));
        if (p && keep) PX_CUDA(cudaMemcpyAsync(q, p, keep * sizeof(T), cudaMemcpyDeviceToDevice, st));
        if (p) {
            PX_CUDA(cudaStreamSynchronize(st));
            cudaFree(p);
        }
        p = q;
        cap = want;
        if (trace_on()) fprintf(stderr, "[mem] grow took %.3f ms\n", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    }
};

// Growable device arena without re-allocation: a large virtual range is reserved once and physical memory is
// mapped behind it in 256 MiB steps (CUDA virtual memory management).  Growing never copies, never frees and
// never moves the data, so pointers handed to kernels stay valid and ingest never stalls on a multi-GB realloc.
struct VmArena {
    // driver entry points are resolved through the runtime (cudaGetDriverEntryPoint): the library does not
    // link libcuda, so it still loads (and reports its symbols) on a machine without a driver
    struct Drv {
        CUresult (*GetGran)(size_t *, const CUmemAllocationProp *, CUmemAllocationGranularity_flags) = nullptr;
        CUresult (*Reserve)(CUdeviceptr *, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
        CUresult (*Create)(CUmemGenericAllocationHandle *, size_t, const CUmemAllocationProp *, unsigned long long) = nullptr;
        CUresult (*Map)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
        CUresult (*SetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc *, size_t) = nullptr;
        CUresult (*Unmap)(CUdeviceptr, size_t) = nullptr;
        CUresult (*Release)(CUmemGenericAllocationHandle) = nullptr;
        CUresult (*AddressFree)(CUdeviceptr, size_t) = nullptr;
        template <typename F>
        static void get(const char *name, F &fn) {
            void *p = nullptr;
            cudaDriverEntryPointQueryResult q;
            PX_CUDA(cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q));
            if (!p || q != cudaDriverEntryPointSuccess) throw std::runtime_error(std::string("driver entry point not found: ") + name);
            fn = reinterpret_cast<F>(p);
        }
        void load() {
            if (Reserve) return;
            get("cuMemGetAllocationGranularity", GetGran);
            get("cuMemAddressReserve", Reserve);
            get("cuMemCreate", Create);
            get("cuMemMap", Map);
            get("cuMemSetAccess", SetAccess);
            get("cuMemUnmap", Unmap);
            get("cuMemRelease", Release);
            get("cuMemAddressFree", AddressFree);
        }
    } drv;
    CUdeviceptr base = 0;
    size_t reserved = 0, gran = 0;
    std::atomic<size_t> mapped{0};       // bytes usable; only ever grows, published by whoever mapped them
    int device = 0;
    std::vector<CUmemGenericAllocationHandle> handles;   // (guarded by mu)
    static constexpr size_t STEP = 256ull << 20;
    // Growth runs AHEAD of use on a helper thread: on a freshly booted GPU one cuMemCreate / cuMemSetAccess of 256 MiB
    // was measured at 20 - 200 ms (profiles/README.md) while kernel launches and stream synchronisation of another
    // thread go on undisturbed, so the ingest path only ever finds the memory already there.
    std::mutex mu;
    std::condition_variable cv_work, cv_done;
    std::thread helper;
    size_t target = 0;                   // the helper maps until mapped >= target (guarded by mu)
    bool stop = false, helper_on = false;
    std::string helper_err;              // first failure of the helper (guarded by mu)
    VmArena() = default;
    VmArena(const VmArena &) = delete;
    VmArena &operator=(const VmArena &) = delete;
    static void check(CUresult r, const char *what) {
        if (r != CUDA_SUCCESS) throw std::runtime_error(std::string(what) + " failed with CUresult " + std::to_string((int) r));
    }
    uint8_t *ptr() const { return reinterpret_cast<uint8_t *>(base); }
    CUmemAllocationProp prop() const {
        CUmemAllocationProp p = {};
        p.type = CU_MEM_ALLOCATION_TYPE_PINNED;
        p.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
        p.location.id = device;
        return p;
    }
    void init(int dev, size_t reserve_bytes) {
        device = dev;
        drv.load();
        CUmemAllocationProp p = prop();
        check(drv.GetGran(&gran, &p, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED), "cuMemGetAllocationGranularity");
        reserved = (reserve_bytes + STEP - 1) / STEP * STEP;
        check(drv.Reserve(&base, reserved, 0, 0, 0), "cuMemAddressReserve");
    }
    // maps the next 256 MiB behind the range (one caller at a time: the helper, or ensure() while the helper is idle)
    void map_step() {
        CUmemAllocationProp p = prop();
        CUmemAccessDesc acc = {};
        acc.location = p.location;
        acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
        const size_t at = mapped.load(std::memory_order_relaxed);
        CUmemGenericAllocationHandle h;
        const auto t0 = std::chrono::steady_clock::now();
        check(drv.Create(&h, STEP, &p, 0), "cuMemCreate");
        const auto t1 = std::chrono::steady_clock::now();
        CUresult r = drv.Map(base + at, STEP, 0, h, 0);
        if (r == CUDA_SUCCESS) r = drv.SetAccess(base + at, STEP, &acc, 1);
        if (r != CUDA_SUCCESS) {
            drv.Release(h);
            check(r, "cuMemMap / cuMemSetAccess");
        }
        const auto t2 = std::chrono::steady_clock::now();
        {
            std::lock_guard<std::mutex> g(mu);
            handles.push_back(h);
        }
        mapped.store(at + STEP, std::memory_order_release);
        if (trace_on())
            fprintf(stderr, "[mem] arena mapped %zu MiB: create %.3f map+access %.3f ms\n", (at + STEP) >> 20,
                    std::chrono::duration<double, std::milli>(t1 - t0).count(), std::chrono::duration<double, std::milli>(t2 - t1).count());
    }
    void helper_loop() {
        cudaSetDevice(device);
        std::unique_lock<std::mutex> lk(mu);
        for (;;) {
            cv_work.wait(lk, [&] { return stop || (helper_err.empty() && mapped.load() < target); });
            if (stop) return;
            lk.unlock();
            std::string e;
            try {
                map_step();
            } catch (const std::exception &ex) {
                e = ex.what();
            }
            lk.lock();
            if (!e.empty()) helper_err = e;
            cv_done.notify_all();
        }
    }
    // make bytes [0, n) usable now, and ask for room ahead of them (a quarter of n, 256 MiB .. 4 GiB) in the background
    void ensure(size_t n) {
        if (n > reserved) throw std::runtime_error("VmArena: reserved range exhausted");
        size_t ahead = 0;
        if (n > STEP / 2) ahead = std::min<size_t>(std::max<size_t>(n / 4, STEP), 16 * STEP);   // (small stores never start the helper)
        const size_t want = std::min(reserved, (n + ahead + STEP - 1) / STEP * STEP);
        if (want <= mapped.load(std::memory_order_acquire)) return;
        std::unique_lock<std::mutex> lk(mu);
        if (!helper_on && ahead == 0) {   // first steps of a small store: inline
            lk.unlock();
            while (mapped.load() < n) map_step();
            return;
        }
        if (!helper_on) {
            helper = std::thread([this] { helper_loop(); });
            helper_on = true;
        }
        if (want > target) target = want;
        cv_work.notify_one();
        cv_done.wait(lk, [&] { return mapped.load() >= n || !helper_err.empty(); });
        if (mapped.load() < n) throw std::runtime_error("VmArena: " + helper_err);
    }
    ~VmArena() {
        if (helper_on) {
            {
                std::lock_guard<std::mutex> g(mu);
                stop = true;
            }
            cv_work.notify_all();
            helper.join();
        }
        if (!base) return;
        if (mapped.load()) drv.Unmap(base, mapped.load());
        for (auto h : handles) drv.Release(h);
        drv.AddressFree(base, reserved);
    }
};

// Growable pinned host buffer (D2H targets that the host then walks)
template <typename T>
struct PinnedBuf {
    T *p = nullptr;
    size_t cap = 0;
    PinnedBuf() = default;
    PinnedBuf(const PinnedBuf &) = delete;
    PinnedBuf &operator=(const PinnedBuf &) = delete;
    ~PinnedBuf() {
        if (p) cudaFreeHost(p);
    }
    void reserve_discard(size_t n) {
        if (n <= cap) return;
        if (p) cudaFreeHost(p);
        p = nullptr;
        size_t want = n + n / 4 + 256;
        PX_CUDA(cudaMallocHost(&p, want * sizeof(T)));
        cap = want;
    }
};

#ifdef __CUDACC__
__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31; }

// acquire-release fence at device scope (cheaper than the sequentially consistent __threadfence)
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long v;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(v));
    return v;
}
__device__ __forceinline__ void fence_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// OR into a word with release semantics: the thread's earlier writes are visible before the bits are
__device__ __forceinline__ void red_release_or_u32(uint32_t *p, uint32_t v) {
    asm volatile("red.release.gpu.global.or.b32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_relaxed_or_u32(uint32_t *p, uint32_t v) {
    asm volatile("red.relaxed.gpu.global.or.b32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_release_u32(uint32_t *p, uint32_t v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// relaxed (no fence) device-scope accesses for look-back status words: the word itself carries the
// whole message (flag + value), so no ordering with other memory is needed and loads can overlap
__device__ __forceinline__ uint32_t ld_relaxed_u32(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
// poll of a flag word whose set bits license later (control-dependent, L2-coherent) loads of other data: the
// "memory" clobber keeps the compiler from moving those loads above the poll
__device__ __forceinline__ uint32_t ld_poll_u32(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u32(uint32_t *p, uint32_t v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v));
}
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v));
}
#endif

}  // namespace pixiu
