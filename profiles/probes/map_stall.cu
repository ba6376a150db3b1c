// does a slow cuMemCreate / cuMemSetAccess in a helper thread stall kernel launches + syncs of the main thread?
#include <cuda.h>
#include <cuda_runtime.h>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <thread>
#include <vector>
#include <algorithm>
__global__ void k_tick(unsigned v, unsigned *dev) { dev[0] = v; }
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main() {
  cudaSetDevice(0); cudaFree(0);
  cudaStream_t st; cudaStreamCreate(&st);
  unsigned *dev; cudaMalloc(&dev, 64);
  std::atomic<int> stop{0};
  std::vector<std::pair<double,double>> maps;  // (start, duration)
  const size_t STEP = 256ull << 20;
  std::thread th([&] {
    cudaSetDevice(0);
    CUdeviceptr base; cuMemAddressReserve(&base, 64ull << 30, 0, 0, 0);
    CUmemAllocationProp p = {}; p.type = CU_MEM_ALLOCATION_TYPE_PINNED; p.location.type = CU_MEM_LOCATION_TYPE_DEVICE; p.location.id = 0;
    CUmemAccessDesc acc = {}; acc.location = p.location; acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    size_t mapped = 0;
    while (!stop.load() && mapped < (60ull << 30)) {
      double a = now();
      CUmemGenericAllocationHandle h;
      CUresult r1 = cuMemCreate(&h, STEP, &p, 0); double b = now();
      CUresult r2 = cuMemMap(base + mapped, STEP, 0, h, 0);
      CUresult r3 = cuMemSetAccess(base + mapped, STEP, &acc, 1); double c = now();
      if (r1 || r2 || r3) { printf("map failed %d %d %d\n", r1, r2, r3); break; }
      mapped += STEP;
      maps.push_back({a, c - a});
      if (c - a > 5e-3) printf("  map #%zu: create %.1f ms, map+access %.1f ms\n", mapped / STEP, 1e3 * (b - a), 1e3 * (c - b));
      std::this_thread::sleep_for(std::chrono::milliseconds(20));
    }
  });
  std::vector<std::pair<double,double>> lat;
  double t0 = now(); unsigned i = 0;
  while (now() - t0 < 6.0) { double a = now(); k_tick<<<1, 1, 0, st>>>(++i, dev); cudaStreamSynchronize(st); lat.push_back({a, now() - a}); }
  stop = 1; th.join();
  // main-thread iterations that overlapped a slow map
  double worst = 0; size_t slow = 0; for (auto &l : lat) { worst = std::max(worst, l.second); if (l.second > 1e-3) slow++; }
  double slow_map = 0; size_t nslow = 0; for (auto &m : maps) if (m.second > 5e-3) { slow_map += m.second; nslow++; }
  printf("maps %zu (slow >5ms: %zu, total %.1f ms); main loop iterations %zu, >1 ms: %zu, worst %.2f ms\n", maps.size(), nslow, 1e3 * slow_map, lat.size(), slow, 1e3 * worst);
  double blocked = 0; for (auto &l : lat) if (l.second > 1e-3) blocked += l.second; printf("main thread time in >1ms iterations: %.1f ms\n", 1e3 * blocked);
  return 0;
}
