// Micro-probe (not product code): host->device copy rate for Arrow-batch-sized pieces (512 KiB columns in
// separate pinned allocations) issued from compiled code: one stream, several streams, cudaMemcpyBatchAsync,
// and a staging memcpy into one large pinned buffer followed by large DMAs.
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main() {
  const size_t piece = 512 << 10, n = 768 * 2;  // two panes' worth of 64 Ki-row columns
  std::vector<void*> h(n);
  for (auto& p : h) { CK(cudaHostAlloc(&p, piece, cudaHostAllocDefault)); memset(p, 1, piece); }
  char* d; CK(cudaMalloc(&d, piece * n));
  cudaStream_t st[8]; for (auto& s : st) CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  const double total = (double)piece * n;
  for (int ns : {1, 2, 4, 8}) {
    for (int rep = 0; rep < 2; ++rep) {
      CK(cudaDeviceSynchronize());
      double t0 = now();
      for (size_t i = 0; i < n; ++i) CK(cudaMemcpyAsync(d + i * piece, h[i], piece, cudaMemcpyHostToDevice, st[i % ns]));
      double t1 = now();
      CK(cudaDeviceSynchronize());
      double t2 = now();
      if (rep) printf("cudaMemcpyAsync x %zu pieces of 512 KiB, %d stream(s): %6.1f GB/s (enqueue %.2f ms, total %.2f ms)\n", n, ns, total / (t2 - t0) / 1e9, (t1 - t0) * 1e3, (t2 - t0) * 1e3);
    }
  }
  // streams taking whole batches (3 consecutive columns) instead of alternating columns
  for (int ns : {2, 4}) {
    CK(cudaDeviceSynchronize());
    double t0 = now();
    for (size_t i = 0; i < n; ++i) CK(cudaMemcpyAsync(d + i * piece, h[i], piece, cudaMemcpyHostToDevice, st[(i / 3) % ns]));
    CK(cudaDeviceSynchronize());
    printf("batches round-robin over %d streams: %6.1f GB/s\n", ns, total / (now() - t0) / 1e9);
  }
#if CUDART_VERSION >= 12080
  for (size_t group : {(size_t)3, (size_t)48, (size_t)768}) {
    std::vector<void*> dsts(n), srcs(n); std::vector<size_t> sizes(n, piece);
    for (size_t i = 0; i < n; ++i) { dsts[i] = d + i * piece; srcs[i] = h[i]; }
    cudaMemcpyAttributes at{}; at.srcAccessOrder = cudaMemcpySrcAccessOrderStream; at.flags = cudaMemcpyFlagPreferOverlapWithCompute;
    size_t idx = 0, fail = 0;
    for (int rep = 0; rep < 2; ++rep) {
      CK(cudaDeviceSynchronize());
      double t0 = now();
      for (size_t o = 0; o < n; o += group) {
        size_t c = std::min(group, n - o);
        CK(cudaMemcpyBatchAsync(dsts.data() + o, srcs.data() + o, sizes.data() + o, c, &at, &idx, 1, &fail, st[0]));
      }
      double t1 = now();
      CK(cudaDeviceSynchronize());
      double t2 = now();
      if (rep) printf("cudaMemcpyBatchAsync groups of %zu: %6.1f GB/s (enqueue %.2f ms, total %.2f ms)\n", group, total / (t2 - t0) / 1e9, (t1 - t0) * 1e3, (t2 - t0) * 1e3);
    }
  }
#endif
  // staging: T host threads memcpy the pieces into one pinned buffer, DMA in 16 MiB pieces behind them
  char* stage; CK(cudaHostAlloc((void**)&stage, piece * n, cudaHostAllocDefault));
  for (int T : {1, 2, 4, 8}) {
    CK(cudaDeviceSynchronize());
    double t0 = now();
    const size_t per_dma = 32;  // pieces per DMA (16 MiB)
    for (size_t o = 0; o < n; o += per_dma) {
      std::vector<std::thread> th;
      for (int t = 0; t < T; ++t) th.emplace_back([&, t] { for (size_t i = o + t; i < o + per_dma && i < n; i += T) memcpy(stage + i * piece, h[i], piece); });
      for (auto& x : th) x.join();
      CK(cudaMemcpyAsync(d + o * piece, stage + o * piece, piece * per_dma, cudaMemcpyHostToDevice, st[0]));
    }
    CK(cudaDeviceSynchronize());
    printf("stage with %d host thread(s) + 16 MiB DMAs: %6.1f GB/s\n", T, total / (now() - t0) / 1e9);
  }
  // one big copy for reference
  CK(cudaDeviceSynchronize());
  double t0 = now();
  CK(cudaMemcpyAsync(d, stage, piece * n, cudaMemcpyHostToDevice, st[0]));
  CK(cudaDeviceSynchronize());
  printf("one %zu MiB copy: %6.1f GB/s\n", (piece * n) >> 20, total / (now() - t0) / 1e9);
  return 0;
}
