// Device exclusive prefix sum of 32-bit counts into 64-bit offsets (compaction, join output sizing, grouping).
#pragma once

#include "common.cuh"

namespace ab {

// Exclusive scan of 32-bit counts into 64-bit offsets: block sums, one-block scan of the block sums,
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
// Exclusive scan of the block sums by one block: 1024 sums per round (warp scans + a scan of the warp totals),
// a running carry between rounds.  (A single thread walking 10 K sums took 0.25 - 0.55 ms per call.)
static __global__ void __launch_bounds__(SCAN_BLOCK) scan_sums_block(unsigned long long* sums, long long n_blocks,
                                                                     unsigned long long* total) {
  __shared__ unsigned long long s_warp[32];
  __shared__ unsigned long long s_carry;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (long long base = 0; base < n_blocks; base += SCAN_BLOCK) {
    const long long i = base + threadIdx.x;
    const unsigned long long v = i < n_blocks ? sums[i] : 0;
    unsigned long long x = v;
    for (int o = 1; o < 32; o <<= 1) {
      unsigned long long y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) s_warp[w] = x;
    __syncthreads();
    if (w == 0) {
      const unsigned long long t = s_warp[lane];
      unsigned long long u = t;
      for (int o = 1; o < 32; o <<= 1) {
        unsigned long long y = __shfl_up_sync(0xffffffffu, u, o);
        if (lane >= o) u += y;
      }
      s_warp[lane] = u - t;
    }
    __syncthreads();
    const unsigned long long carry = s_carry;
    if (i < n_blocks) sums[i] = carry + s_warp[w] + x - v;
    __syncthreads();
    if (threadIdx.x == SCAN_BLOCK - 1) s_carry = carry + s_warp[w] + x;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = s_carry;
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
  scan_sums_block<<<1, SCAN_BLOCK, 0, stream>>>(sums.as<unsigned long long>(), nb, total_dev);
  AB_CUDA(cudaGetLastError());
  if (nb > 0) {
    scan_apply<<<(unsigned)nb, SCAN_BLOCK, 0, stream>>>(cnt, n, sums.as<unsigned long long>(), off);
    AB_CUDA(cudaGetLastError());
  }
}

}  // namespace ab
