// K8b: bulk TSV writer.  Replaces the per-segment Python loop of deepgrp/__main__.py:288-292
//   outstream.write("{}\t{}\t{}\t{}\t{}\n".format(filename, header, start, end, label))
// Segment triples (start, end, label) stay on the device; every row's text is produced by one
// thread at the offset given by a 64-bit exclusive scan of the row lengths, so the host only
// receives finished bytes.
#include "dgrp_internal.cuh"
#include "seg_core.cuh"

namespace dgrp {

constexpr int TSV_THREADS = 256;

using seg::fmt_len;
using seg::fmt_put;

__device__ __forceinline__ unsigned row_len(const int64_t *tri, int64_t i, int prefix_len) {
  return (unsigned)(prefix_len + fmt_len(tri[3 * i]) + 1 + fmt_len(tri[3 * i + 1]) + 1 +
                    fmt_len(tri[3 * i + 2]) + 1);
}

__global__ void tsv_len_kernel(const int64_t *__restrict__ tri, int64_t n, int prefix_len,
                               unsigned long long *tile_sum) {
  __shared__ unsigned long long s_sum;
  if (threadIdx.x == 0) s_sum = 0;
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * TSV_THREADS + threadIdx.x;
  unsigned long long v = i < n ? row_len(tri, i, prefix_len) : 0;
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  if ((threadIdx.x & 31) == 0) atomicAdd(&s_sum, v);
  __syncthreads();
  if (threadIdx.x == 0) tile_sum[blockIdx.x] = s_sum;
}

// in-place exclusive scan of 64-bit tile sums by one block; total[0] = sum
__global__ void tsv_scan_kernel(unsigned long long *a, int64_t m, unsigned long long *total) {
  __shared__ unsigned long long s_warp[32];
  __shared__ unsigned long long s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  for (int64_t base = 0; base < m; base += blockDim.x) {
    const int64_t i = base + threadIdx.x;
    const unsigned long long v = i < m ? a[i] : 0;
    unsigned long long x = v;
    for (int off = 1; off < 32; off <<= 1) {
      const unsigned long long t = __shfl_up_sync(0xffffffffu, x, off);
      if (lane >= off) x += t;
    }
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    if (warp == 0) {
      unsigned long long w = lane < nwarp ? s_warp[lane] : 0;
      for (int off = 1; off < 32; off <<= 1) {
        const unsigned long long t = __shfl_up_sync(0xffffffffu, w, off);
        if (lane >= off) w += t;
      }
      s_warp[lane] = w;
    }
    __syncthreads();
    const unsigned long long excl = s_carry + (warp ? s_warp[warp - 1] : 0) + x - v;
    if (i < m) a[i] = excl;
    __syncthreads();
    if (threadIdx.x == 0) s_carry += s_warp[nwarp - 1];
    __syncthreads();
  }
  if (threadIdx.x == 0) total[0] = s_carry;
}

// One thread formats one row.  The block's rows are contiguous in the output, so the text is first
// assembled in shared memory (at the same 16-byte phase as its global destination) and then copied
// out with aligned 16-byte stores; blocks whose text does not fit the staging buffer write directly.
constexpr int TSV_STAGE = 40 * 1024;

__global__ void tsv_write_kernel(const int64_t *__restrict__ tri, int64_t n,
                                 const uint8_t *__restrict__ prefix, int prefix_len,
                                 const unsigned long long *__restrict__ tile_off,
                                 const unsigned long long *__restrict__ total,
                                 uint8_t *__restrict__ out) {
  __shared__ unsigned s_w[TSV_THREADS / 32];
  __shared__ __align__(16) uint8_t s_text[TSV_STAGE + 16];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t i = (int64_t)blockIdx.x * TSV_THREADS + threadIdx.x;
  const unsigned len = i < n ? row_len(tri, i, prefix_len) : 0;
  unsigned incl = len;
  for (int off = 1; off < 32; off <<= 1) {
    const unsigned t = __shfl_up_sync(0xffffffffu, incl, off);
    if (lane >= off) incl += t;
  }
  if (lane == 31) s_w[warp] = incl;
  __syncthreads();
  unsigned before = 0;
  for (int w = 0; w < warp; ++w) before += s_w[w];
  const unsigned long long g0 = tile_off[blockIdx.x];
  const unsigned long long g1 = (blockIdx.x + 1 < gridDim.x) ? tile_off[blockIdx.x + 1] : total[0];
  const unsigned block_len = (unsigned)(g1 - g0);
  const unsigned phase = (unsigned)((reinterpret_cast<uintptr_t>(out) + g0) & 15u);
  const bool staged = block_len + phase <= TSV_STAGE;
  if (i < n) {
    uint8_t *p = staged ? s_text + phase + before + (incl - len) : out + g0 + before + (incl - len);
    for (int k = 0; k < prefix_len; ++k) p[k] = prefix[k];
    p += prefix_len;
    p = fmt_put(p, tri[3 * i]);      *p++ = '\t';
    p = fmt_put(p, tri[3 * i + 1]);  *p++ = '\t';
    p = fmt_put(p, tri[3 * i + 2]);  *p++ = '\n';
  }
  if (!staged) return;
  __syncthreads();
  uint8_t *dst = out + g0 - phase;                  // 16-byte aligned
  const unsigned end = phase + block_len;
  // head and tail bytes that do not fill a 16-byte word, then whole words
  const unsigned first_word = phase ? 16u : 0u, last_word = end & ~15u;
  for (unsigned k = phase + threadIdx.x; k < (first_word < end ? first_word : end); k += TSV_THREADS) dst[k] = s_text[k];
  if (last_word >= first_word) {
    for (unsigned k = first_word / 16 + threadIdx.x; k < last_word / 16; k += TSV_THREADS)
      reinterpret_cast<uint4 *>(dst)[k] = reinterpret_cast<const uint4 *>(s_text)[k];
    for (unsigned k = (last_word > first_word ? last_word : first_word) + threadIdx.x; k < end; k += TSV_THREADS)
      dst[k] = s_text[k];
  }
}

// Format n triples on the device into d_out (reserved by the caller through *need).  Step 1
// (d_out == nullptr): returns the byte count in *need.  Step 2: writes the text.
int run_tsv_measure(dgrp_ctx *c, const int64_t *d_tri, int64_t n, int prefix_len, int64_t *need) {
  *need = 0;
  if (n <= 0) return DGRP_OK;
  const int64_t ntiles = (n + TSV_THREADS - 1) / TSV_THREADS;
  DGRP_CHECK(c->scan.reserve((size_t)ntiles * 8 + 64));
  DGRP_CHECK(c->pin_small.reserve(256));
  unsigned long long *tile = c->scan.as<unsigned long long>();
  unsigned long long *total = tile + ntiles;
  tsv_len_kernel<<<(unsigned)ntiles, TSV_THREADS, 0, c->stream>>>(d_tri, n, prefix_len, tile);
  tsv_scan_kernel<<<1, 1024, 0, c->stream>>>(tile, ntiles, total);
  c->launches += 2;
  unsigned long long *h = c->pin_small.as<unsigned long long>();
  DGRP_CHECK(fetch_small(c, h, total, 8));
  DGRP_CUDA(cudaStreamSynchronize(c->stream));
  *need = (int64_t)h[0];
  return DGRP_OK;
}

int run_tsv_write(dgrp_ctx *c, const int64_t *d_tri, int64_t n, const uint8_t *d_prefix,
                  int prefix_len, uint8_t *d_out) {
  if (n <= 0) return DGRP_OK;
  const int64_t ntiles = (n + TSV_THREADS - 1) / TSV_THREADS;
  const unsigned long long *tile = c->scan.as<unsigned long long>();
  tsv_write_kernel<<<(unsigned)ntiles, TSV_THREADS, 0, c->stream>>>(d_tri, n, d_prefix, prefix_len, tile,
                                                                    tile + ntiles, d_out);
  c->launches++;
  DGRP_CUDA(cudaGetLastError());
  return DGRP_OK;
}

}  // namespace dgrp
