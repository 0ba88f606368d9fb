// Device exclusive prefix sum of 32-bit counts into 64-bit offsets (compaction, join output sizing, grouping).
#pragma once

#include "common.cuh"

namespace ab {

// Exclusive scan of 32-bit counts into 64-bit offsets: block sums, serial scan of the (few) block sums,
// then per-block scan.  n is at most a few hundred million: 1024-element blocks.
constexpr int SCAN_BLOCK = 1024;
static __global__ void scan_block_sums(const unsigned int* __restrict__ in, long long n, unsigned long long* __restrict__ sums) {
  __shared__ unsigned long long s[32];
  long long i = (long long)blockIdx.x * SCAN_BLOCK + threadIdx.x;
  unsigned long long v = i < n ? in[i] : 0;
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    unsigned long long t = s[threadIdx.x];
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (threadIdx.x == 0) sums[blockIdx.x] = t;
  }
}
static __global__ void scan_sums_serial(unsigned long long* sums, long long n_blocks, unsigned long long* total) {
  unsigned long long acc = 0;
  for (long long b = 0; b < n_blocks; ++b) {
    unsigned long long v = sums[b];
    sums[b] = acc;
    acc += v;
  }
  *total = acc;
}
static __global__ void scan_apply(const unsigned int* __restrict__ in, long long n, const unsigned long long* __restrict__ sums,
                           unsigned long long* __restrict__ out) {
  __shared__ unsigned long long s_warp[32];
  long long i = (long long)blockIdx.x * SCAN_BLOCK + threadIdx.x;
  unsigned long long v = i < n ? in[i] : 0;
  unsigned long long x = v;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int o = 1; o < 32; o <<= 1) {
    unsigned long long y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) s_warp[w] = x;
  __syncthreads();
  if (w == 0) {
    unsigned long long t = s_warp[lane];
    unsigned long long u = t;
    for (int o = 1; o < 32; o <<= 1) {
      unsigned long long y = __shfl_up_sync(0xffffffffu, u, o);
      if (lane >= o) u += y;
    }
    s_warp[lane] = u - t;
  }
  __syncthreads();
  if (i < n) out[i] = sums[blockIdx.x] + s_warp[w] + x - v;
}


// off[i] = sum(cnt[0..i)), *total_dev = sum(cnt[0..n)); `sums` is scratch that grows as needed.
inline void device_exclusive_scan(const unsigned int* cnt, int64_t n, unsigned long long* off, unsigned long long* total_dev,
                                  DevBuf& sums, cudaStream_t stream) {
  const int64_t nb = (n + SCAN_BLOCK - 1) / SCAN_BLOCK;
  const size_t need = (size_t)(nb > 1 ? nb : 1) * 8;
  if (sums.bytes < need) sums.alloc(need * 2);
  if (nb > 0) {
    scan_block_sums<<<(unsigned)nb, SCAN_BLOCK, 0, stream>>>(cnt, n, sums.as<unsigned long long>());
    AB_CUDA(cudaGetLastError());
  }
  scan_sums_serial<<<1, 1, 0, stream>>>(sums.as<unsigned long long>(), nb, total_dev);
  AB_CUDA(cudaGetLastError());
  if (nb > 0) {
    scan_apply<<<(unsigned)nb, SCAN_BLOCK, 0, stream>>>(cnt, n, sums.as<unsigned long long>(), off);
    AB_CUDA(cudaGetLastError());
  }
}

}  // namespace ab
