// Round-2 probe: what a launch that does (almost) nothing costs on this box, by CUDA events and by the host clock, for the
// ways a kernel can hand a 32-byte result to the host.   nvcc -arch=sm_100a -O3 -o scratch/r02_latency_probe scratch/r02_latency_probe.cu
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <algorithm>
#include <vector>
__global__ void k_empty() {}
__global__ void k_dev(double* out) { if (threadIdx.x == 0) out[0] = 1.0; }
__global__ void k_mapped_fence(double* out, volatile unsigned long long* flag, unsigned long long seq) {
  if (threadIdx.x == 0) { out[0] = 1.0; out[1] = 2.0; out[2] = 3.0; __threadfence_system(); *flag = seq; }
}
__global__ void k_mapped_nofence(double* out, volatile unsigned long long* flag, unsigned long long seq) {
  if (threadIdx.x == 0) { out[0] = 1.0; out[1] = 2.0; out[2] = 3.0; *flag = seq; }
}
template <class L, class W> void run(const char* name, L launch, W wait, cudaStream_t s) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  std::vector<float> ev; std::vector<double> host;
  for (int i = 0; i < 300; ++i) {
    auto t0 = std::chrono::steady_clock::now();
    cudaEventRecord(e0, s); launch(i + 1); cudaEventRecord(e1, s); wait(i + 1);
    auto t1 = std::chrono::steady_clock::now();
    cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (i >= 50) { ev.push_back(ms * 1e3f); host.push_back(std::chrono::duration<double, std::micro>(t1 - t0).count()); }
  }
  std::sort(ev.begin(), ev.end()); std::sort(host.begin(), host.end());
  printf("%-34s events min %.2f med %.2f us | host (launch..result) min %.2f med %.2f us\n", name, ev[0], ev[ev.size() / 2], host[0], host[host.size() / 2]);
}
int main() {
  cudaStream_t s; cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
  double* dev; cudaMalloc(&dev, 64);
  char* mh; cudaHostAlloc((void**)&mh, 4096, cudaHostAllocMapped); char* md; cudaHostGetDevicePointer((void**)&md, mh, 0);
  double* pin; cudaMallocHost((void**)&pin, 64);
  volatile unsigned long long* flag = (volatile unsigned long long*)(mh + 64);
  run("empty kernel + stream sync", [&](int) { k_empty<<<1, 256, 0, s>>>(); }, [&](int) { cudaStreamSynchronize(s); }, s);
  run("device write + D2H + stream sync", [&](int) { k_dev<<<1, 256, 0, s>>>(dev); cudaMemcpyAsync(pin, dev, 32, cudaMemcpyDeviceToHost, s); }, [&](int) { cudaStreamSynchronize(s); }, s);
  run("mapped write + fence.sys + flag, poll", [&](int i) { k_mapped_fence<<<1, 256, 0, s>>>((double*)md, (volatile unsigned long long*)(md + 64), i); },
      [&](int i) { while (*flag != (unsigned long long)i) {} }, s);
  run("mapped write, no fence, poll", [&](int i) { k_mapped_nofence<<<1, 256, 0, s>>>((double*)md, (volatile unsigned long long*)(md + 64), i); },
      [&](int i) { while (*flag != (unsigned long long)i) {} }, s);
  run("391 empty CTAs + stream sync", [&](int) { k_empty<<<391, 256, 0, s>>>(); }, [&](int) { cudaStreamSynchronize(s); }, s);
  return 0;
}
