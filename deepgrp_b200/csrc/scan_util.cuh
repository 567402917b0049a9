// Ordered stream compaction in three launches (tile counts -> single-block exclusive scan ->
// ordered scatter).  Used by the MSS kernels (event list, surviving candidates) and the FASTA decoder.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dgrp {

constexpr int CP_THREADS = 256;
constexpr int CP_ITEMS = 8;
constexpr int CP_TILE = CP_THREADS * CP_ITEMS;

template <class Pred>
__global__ void cp_count_kernel(int64_t n, Pred pred, unsigned int *tile_counts) {
  __shared__ unsigned int s_cnt;
  if (threadIdx.x == 0) s_cnt = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * CP_TILE;
  unsigned int c = 0;
#pragma unroll
  for (int k = 0; k < CP_ITEMS; ++k) {
    const int64_t i = base + (int64_t)k * CP_THREADS + threadIdx.x;
    if (i < n && pred(i)) ++c;
  }
  for (int off = 16; off > 0; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(&s_cnt, c);
  __syncthreads();
  if (threadIdx.x == 0) tile_counts[blockIdx.x] = s_cnt;
}

// In-place exclusive scan of `a[0..m)` by one block; total[0] receives the sum.
__global__ void cp_scan_kernel(unsigned int *a, int64_t m, unsigned long long *total);

template <class Pred, class Emit>
__global__ void cp_scatter_kernel(int64_t n, Pred pred, Emit emit, const unsigned int *tile_offsets) {
  __shared__ unsigned int s_w[CP_THREADS / 32];
  __shared__ unsigned int s_run;
  if (threadIdx.x == 0) s_run = tile_offsets[blockIdx.x];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t base = (int64_t)blockIdx.x * CP_TILE;
  for (int k = 0; k < CP_ITEMS; ++k) {
    const int64_t i = base + (int64_t)k * CP_THREADS + threadIdx.x;
    const bool f = i < n && pred(i);
    const unsigned int m = __ballot_sync(0xffffffffu, f);
    if (lane == 0) s_w[warp] = __popc(m);
    __syncthreads();
    unsigned int pos = s_run;
    for (int w = 0; w < warp; ++w) pos += s_w[w];
    if (f) emit(i, (int64_t)pos + __popc(m & ((1u << lane) - 1u)));
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned int t = 0;
      for (int w = 0; w < CP_THREADS / 32; ++w) t += s_w[w];
      s_run += t;
    }
    __syncthreads();
  }
}

}  // namespace dgrp
