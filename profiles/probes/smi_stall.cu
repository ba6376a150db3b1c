// which CUDA calls stall while nvidia-smi samples the GPU?  (measurement aid, not part of the product)
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <string>
__global__ void k_tick(volatile unsigned *flag, unsigned v, unsigned *dev) { dev[0] = v; if (flag) { *flag = v; __threadfence_system(); } }
__global__ void k_spin(unsigned ns) { unsigned long long t0; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0)); unsigned long long t = t0; while (t - t0 < ns) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); }
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
struct Stat { std::vector<double> v; void add(double x) { v.push_back(x); }
  void report(const char *name) { std::sort(v.begin(), v.end()); double s = 0, big = 0; for (double x : v) { s += x; if (x > 1e-3) big += x; }
    printf("%-34s n=%7zu mean=%8.2f us p50=%8.2f p99=%8.2f max=%9.2f us  time in >1ms iterations: %5.1f %%\n", name, v.size(), 1e6 * s / v.size(), 1e6 * v[v.size() / 2], 1e6 * v[v.size() * 99 / 100], 1e6 * v.back(), 100 * big / s); } };
int main(int argc, char **argv) {
  double secs = argc > 1 ? atof(argv[1]) : 3.0;
  cudaSetDevice(0); cudaStream_t st; cudaStreamCreate(&st);
  unsigned *dev; cudaMalloc(&dev, 64);
  unsigned *hflag; cudaHostAlloc(&hflag, 64, cudaHostAllocMapped); volatile unsigned *vflag = hflag; unsigned *dflag; cudaHostGetDevicePointer(&dflag, hflag, 0);
  unsigned *pinned; cudaHostAlloc(&pinned, 64, cudaHostAllocDefault);
  unsigned pageable[16];
  cudaEvent_t ev; cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
  for (int i = 0; i < 100; i++) k_tick<<<1, 1, 0, st>>>(nullptr, i, dev); cudaStreamSynchronize(st);
  { Stat s; double t0 = now(); unsigned i = 0; while (now() - t0 < secs) { double a = now(); k_tick<<<1, 1, 0, st>>>(nullptr, ++i, dev); cudaStreamSynchronize(st); s.add(now() - a); } s.report("A launch + streamSynchronize"); }
  { Stat s; double t0 = now(); unsigned i = 0; *vflag = 0; while (now() - t0 < secs) { double a = now(); k_tick<<<1, 1, 0, st>>>(dflag, ++i, dev); while (*vflag != i) { } s.add(now() - a); } s.report("B launch + mapped flag spin"); }
  { Stat s; double t0 = now(); unsigned i = 0; while (now() - t0 < secs) { double a = now(); k_tick<<<1, 1, 0, st>>>(nullptr, ++i, dev); cudaMemcpyAsync(pageable, dev, 16, cudaMemcpyDeviceToHost, st); cudaStreamSynchronize(st); s.add(now() - a); } s.report("C launch + D2H pageable + sync"); }
  { Stat s; double t0 = now(); unsigned i = 0; while (now() - t0 < secs) { double a = now(); k_tick<<<1, 1, 0, st>>>(nullptr, ++i, dev); cudaMemcpyAsync(pinned, dev, 16, cudaMemcpyDeviceToHost, st); cudaStreamSynchronize(st); s.add(now() - a); } s.report("D launch + D2H pinned + sync"); }
  { Stat s; double t0 = now(); unsigned i = 0; while (now() - t0 < secs) { double a = now(); k_tick<<<1, 1, 0, st>>>(nullptr, ++i, dev); cudaEventRecord(ev, st); while (cudaEventQuery(ev) == cudaErrorNotReady) { } s.add(now() - a); } s.report("E launch + eventQuery spin"); }
  { Stat s; double t0 = now(); while (now() - t0 < secs) { for (int j = 0; j < 200; j++) { double a = now(); k_spin<<<1, 1, 0, st>>>(3000); s.add(now() - a); } cudaStreamSynchronize(st); } s.report("F launch only (3 us kernels)"); }
  { Stat s; double t0 = now(); while (now() - t0 < secs) { double a = now(); void *p; cudaMalloc(&p, 1 << 20); cudaFree(p); s.add(now() - a); } s.report("G cudaMalloc + cudaFree 1 MB"); }
  { Stat s; double t0 = now(); while (now() - t0 < secs) { double a = now(); cudaMemsetAsync(dev, 0, 64, st); cudaStreamSynchronize(st); s.add(now() - a); } s.report("H memsetAsync + sync"); }
  return 0;
}
