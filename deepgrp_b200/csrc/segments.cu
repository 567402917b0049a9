// K8: run-length segment extraction.  Replaces deepgrp/sequence.pyx:40-53 (get_segments) and
// :79-85 (yield_segments), including the reference's `size - 1` loop bounds: a run that reaches
// the last element is emitted as [s, n-1) plus [n-1, n).
#include "dgrp_internal.cuh"
#include "seg_core.cuh"

namespace dgrp {

constexpr int SEG_THREADS = 256;
constexpr int SEG_ITEMS = 8;
constexpr int SEG_TILE = SEG_THREADS * SEG_ITEMS;

// A non-zero run starts at p when lab[p] != 0 and (p == 0, the label changes, or p == n-1);
// it ends at e (exclusive) when lab[e-1] != 0 and (e == n, the label changes at e, or e == n-1).
// Starts and ends are in one-to-one order, so two ordered compactions pair them up.
// `open`: the labels are a prefix of a record that goes on -- the special case of the record's last element
// (the `n - 1` terms) does not apply; a run that reaches n is still closed there and dropped by the caller.
template <typename L>
__device__ __forceinline__ void seg_flags(const L *__restrict__ lab, int64_t n, int64_t p,
                                          bool &is_start, bool &is_end_after, bool open = false) {
  // is_end_after: a run ends at e = p + 1
  const L cur = lab[p];
  if (cur == 0) { is_start = false; is_end_after = false; return; }
  is_start = (p == 0) || (lab[p - 1] != cur) || (!open && p == n - 1);
  const int64_t e = p + 1;
  is_end_after = (e == n) || (lab[e] != cur) || (!open && e == n - 1);
}

template <typename L>
__global__ void seg_count_kernel(const L *__restrict__ lab, int64_t n, unsigned int *bstart,
                                 unsigned int *bend, bool open) {
  __shared__ unsigned int s_cnt[2];
  if (threadIdx.x == 0) { s_cnt[0] = 0; s_cnt[1] = 0; }
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * SEG_TILE;
  unsigned int cs = 0, ce = 0;
#pragma unroll
  for (int k = 0; k < SEG_ITEMS; ++k) {
    int64_t p = base + (int64_t)k * SEG_THREADS + threadIdx.x;
    if (p < n) {
      bool a, b;
      seg_flags(lab, n, p, a, b, open);
      cs += a; ce += b;
    }
  }
  for (int off = 16; off > 0; off >>= 1) {
    cs += __shfl_xor_sync(0xffffffffu, cs, off);
    ce += __shfl_xor_sync(0xffffffffu, ce, off);
  }
  if ((threadIdx.x & 31) == 0) { atomicAdd(&s_cnt[0], cs); atomicAdd(&s_cnt[1], ce); }
  __syncthreads();
  if (threadIdx.x == 0) { bstart[blockIdx.x] = s_cnt[0]; bend[blockIdx.x] = s_cnt[1]; }
}

// Single-block exclusive scan of two count arrays (in place); totals[0..1] receive the sums.
__global__ void seg_scan_kernel(unsigned int *a, unsigned int *b, int64_t nblk,
                                unsigned long long *totals) {
  __shared__ unsigned long long s_warp[2][32];
  __shared__ unsigned long long s_carry[2];
  if (threadIdx.x == 0) { s_carry[0] = 0; s_carry[1] = 0; }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  for (int64_t base = 0; base < nblk; base += blockDim.x) {
    int64_t i = base + threadIdx.x;
    unsigned long long va = i < nblk ? a[i] : 0, vb = i < nblk ? b[i] : 0;
    unsigned long long xa = va, xb = vb;
    for (int off = 1; off < 32; off <<= 1) {
      unsigned long long ta = __shfl_up_sync(0xffffffffu, xa, off);
      unsigned long long tb = __shfl_up_sync(0xffffffffu, xb, off);
      if (lane >= off) { xa += ta; xb += tb; }
    }
    if (lane == 31) { s_warp[0][warp] = xa; s_warp[1][warp] = xb; }
    __syncthreads();
    if (warp == 0) {
      unsigned long long wa = lane < nwarp ? s_warp[0][lane] : 0, wb = lane < nwarp ? s_warp[1][lane] : 0;
      for (int off = 1; off < 32; off <<= 1) {
        unsigned long long ta = __shfl_up_sync(0xffffffffu, wa, off);
        unsigned long long tb = __shfl_up_sync(0xffffffffu, wb, off);
        if (lane >= off) { wa += ta; wb += tb; }
      }
      s_warp[0][lane] = wa; s_warp[1][lane] = wb;  // inclusive over warps
    }
    __syncthreads();
    unsigned long long pa = s_carry[0] + (warp ? s_warp[0][warp - 1] : 0) + xa - va;
    unsigned long long pb = s_carry[1] + (warp ? s_warp[1][warp - 1] : 0) + xb - vb;
    if (i < nblk) { a[i] = (unsigned int)pa; b[i] = (unsigned int)pb; }
    __syncthreads();
    if (threadIdx.x == 0) { s_carry[0] += s_warp[0][nwarp - 1]; s_carry[1] += s_warp[1][nwarp - 1]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { totals[0] = s_carry[0]; totals[1] = s_carry[1]; }
}

template <typename L>
__global__ void seg_scatter_kernel(const L *__restrict__ lab, int64_t n, const unsigned int *bstart,
                                   const unsigned int *bend, int64_t *__restrict__ triples,
                                   int64_t offset, bool open) {
  __shared__ unsigned int s_ws[SEG_THREADS / 32], s_we[SEG_THREADS / 32];
  __shared__ unsigned int s_run[2];
  if (threadIdx.x == 0) { s_run[0] = bstart[blockIdx.x]; s_run[1] = bend[blockIdx.x]; }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t base = (int64_t)blockIdx.x * SEG_TILE;
  for (int k = 0; k < SEG_ITEMS; ++k) {
    int64_t p = base + (int64_t)k * SEG_THREADS + threadIdx.x;
    bool a = false, b = false;
    if (p < n) seg_flags(lab, n, p, a, b, open);
    unsigned int ma = __ballot_sync(0xffffffffu, a), mb = __ballot_sync(0xffffffffu, b);
    if (lane == 0) { s_ws[warp] = __popc(ma); s_we[warp] = __popc(mb); }
    __syncthreads();
    unsigned int pa = s_run[0], pb = s_run[1];
    for (int w = 0; w < warp; ++w) { pa += s_ws[w]; pb += s_we[w]; }
    unsigned int lower = (1u << lane) - 1u;
    if (a) {
      int64_t idx = pa + __popc(ma & lower);
      triples[3 * idx] = p + offset;
      triples[3 * idx + 2] = (int64_t)lab[p];
    }
    if (b) {
      int64_t idx = pb + __popc(mb & lower);
      triples[3 * idx + 1] = p + 1 + offset;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned int ta = 0, tb = 0;
      for (int w = 0; w < SEG_THREADS / 32; ++w) { ta += s_ws[w]; tb += s_we[w]; }
      s_run[0] += ta; s_run[1] += tb;
    }
    __syncthreads();
  }
}

// ---- uint8 labels (the product path): 16 labels per thread in registers, triples staged through shared memory ----
// The byte-per-thread kernels above read every label three times and store a triple as three scattered
// 8-byte writes from two different threads: 4.8 % / 1.3 % of the HBM peak on the 248 Mbp record
// (profiles/r02n_*).  Here a thread takes 16 consecutive labels with one 128-bit load (the two neighbours
// come from the adjacent lanes), compares four labels per 32-bit operation (a byte-by-byte version of the same
// kernel spent ~25 instructions per label and was no faster than the old one), start / end flags are 16-bit
// masks, and the tile's rows leave as consecutive 8-byte stores.  The label pointer may be any byte address (a record is processed in position slabs, api.cu):
// the tile grid is laid over the 16-byte aligned address below it.
constexpr int SV_THREADS = 256;
constexpr int SV_PER = seg::PER;
constexpr int SV_TILE = SV_THREADS * SV_PER;   // 4096 labels

struct SegWord {
  unsigned ms, me;      // bit k: a run starts at / ends after element k of the thread's 16
  uint32_t q[4];        // the 16 labels (little endian; 0 outside the array)
};

// labels v0 .. v0+15 of the virtual (aligned) array; p = v - mis is the index into lab[0, n)
__device__ __forceinline__ void seg_word(const uint8_t *__restrict__ lab, int64_t n, int mis, int64_t v0, bool open,
                                         SegWord &w) {
  const int lane = threadIdx.x & 31;
  const int64_t p0 = v0 - mis;
  uint32_t q[4] = {0u, 0u, 0u, 0u};
  const bool whole = p0 >= 0 && p0 + SV_PER <= n;
  if (whole) {
    const uint4 t = *reinterpret_cast<const uint4 *>(lab + p0);   // 16-byte aligned
    q[0] = t.x; q[1] = t.y; q[2] = t.z; q[3] = t.w;
  } else if (p0 + SV_PER > 0 && p0 < n) {                // the first / last word of the array: byte by byte
#pragma unroll
    for (int k = 0; k < SV_PER; ++k) {
      const int64_t p = p0 + k;
      if (p >= 0 && p < n) q[k >> 2] |= (uint32_t)lab[p] << (8 * (k & 3));
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) w.q[i] = q[i];
  // neighbours across the thread boundary: the adjacent lanes' edge bytes, global loads at the warp's edges
  unsigned prev = __shfl_up_sync(0xffffffffu, q[3] >> 24, 1);
  unsigned next = __shfl_down_sync(0xffffffffu, q[0] & 0xffu, 1);
  if (lane == 0) prev = (p0 - 1 >= 0 && p0 - 1 < n) ? lab[p0 - 1] : 0u;
  if (lane == 31) next = (p0 + SV_PER >= 0 && p0 + SV_PER < n) ? lab[p0 + SV_PER] : 0u;
  seg::flags16(q, prev, next, p0, n, open, w.ms, w.me);
}

__global__ void __launch_bounds__(SV_THREADS) segv_count_kernel(const uint8_t *__restrict__ lab, int64_t n, int mis,
                                                                int open, unsigned int *bstart, unsigned int *bend) {
  __shared__ unsigned int s_cnt[2];
  if (threadIdx.x == 0) { s_cnt[0] = 0; s_cnt[1] = 0; }
  __syncthreads();
  SegWord w;
  seg_word(lab, n, mis, (int64_t)blockIdx.x * SV_TILE + (int64_t)threadIdx.x * SV_PER, open != 0, w);
  unsigned cs = __popc(w.ms), ce = __popc(w.me);
  for (int off = 16; off > 0; off >>= 1) {
    cs += __shfl_xor_sync(0xffffffffu, cs, off);
    ce += __shfl_xor_sync(0xffffffffu, ce, off);
  }
  if ((threadIdx.x & 31) == 0) { atomicAdd(&s_cnt[0], cs); atomicAdd(&s_cnt[1], ce); }
  __syncthreads();
  if (threadIdx.x == 0) { bstart[blockIdx.x] = s_cnt[0]; bend[blockIdx.x] = s_cnt[1]; }
}

__global__ void __launch_bounds__(SV_THREADS) segv_scatter_kernel(const uint8_t *__restrict__ lab, int64_t n, int mis,
                                                                  int open, const unsigned int *__restrict__ bstart,
                                                                  const unsigned int *__restrict__ bend,
                                                                  int64_t *__restrict__ triples, int64_t offset) {
  __shared__ unsigned short s_spos[SV_TILE], s_epos[SV_TILE];   // tile-local index of a start / of a run's last element
  __shared__ __align__(16) uint8_t s_lab[SV_TILE];              // the tile's labels
  __shared__ unsigned int s_ws[SV_THREADS / 32], s_we[SV_THREADS / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t tile_v0 = (int64_t)blockIdx.x * SV_TILE;
  SegWord w;
  seg_word(lab, n, mis, tile_v0 + (int64_t)threadIdx.x * SV_PER, open != 0, w);
  reinterpret_cast<uint4 *>(s_lab)[threadIdx.x] = make_uint4(w.q[0], w.q[1], w.q[2], w.q[3]);
  // exclusive offsets of this thread's starts / ends inside the tile
  const unsigned cs = __popc(w.ms), ce = __popc(w.me);
  unsigned is = cs, ie = ce;
  for (int off = 1; off < 32; off <<= 1) {
    const unsigned ts = __shfl_up_sync(0xffffffffu, is, off), te = __shfl_up_sync(0xffffffffu, ie, off);
    if (lane >= off) { is += ts; ie += te; }
  }
  if (lane == 31) { s_ws[warp] = is; s_we[warp] = ie; }
  __syncthreads();
  unsigned os = is - cs, oe = ie - ce, ns = 0, ne = 0;
#pragma unroll
  for (int k = 0; k < SV_THREADS / 32; ++k) {
    if (k < warp) { os += s_ws[k]; oe += s_we[k]; }
    ns += s_ws[k]; ne += s_we[k];
  }
  for (unsigned m = w.ms; m; m &= m - 1) s_spos[os++] = (unsigned short)(threadIdx.x * SV_PER + __ffs(m) - 1);
  for (unsigned m = w.me; m; m &= m - 1) s_epos[oe++] = (unsigned short)(threadIdx.x * SV_PER + __ffs(m) - 1);
  __syncthreads();
  // Starts and ends alternate along the record, so the run ordinal of the tile's first end (E0) is the one of
  // its first start (S0) or one less (a run that is open at the tile's first element).  Rows o = E0 .. : field
  // 0 / 2 from this tile's starts, field 1 from its ends; a row that straddles tiles is completed by both.
  const int64_t S0 = bstart[blockIdx.x], E0 = bend[blockIdx.x];
  const int64_t o_end = (S0 + ns > E0 + ne) ? S0 + ns : E0 + ne;
  const int64_t pos0 = tile_v0 - mis + offset;            // record position of the tile's element 0
  const int fields = (int)(o_end - E0) * 3;
  for (int f = threadIdx.x; f < fields; f += SV_THREADS) {
    const int64_t o = E0 + f / 3;
    const int fld = f - (f / 3) * 3;
    const int64_t si = o - S0, ei = o - E0;
    if (fld == 1) {
      if (ei < (int64_t)ne) triples[3 * o + 1] = pos0 + s_epos[ei] + 1;
    } else if (si >= 0 && si < (int64_t)ns) {
      const unsigned at = s_spos[si];
      triples[3 * o + fld] = fld == 0 ? pos0 + at : (int64_t)s_lab[at];
    }
  }
}

template <typename L>
static int run_segments_t(dgrp_ctx *c, const L *d_lab, int64_t n, int64_t offset, bool keep_zero,
                          int64_t **d_triples, int64_t *n_out, int64_t *open_tail) {
  *n_out = 0;
  *d_triples = nullptr;
  if (open_tail) *open_tail = n > 0 ? n : 0;
  if (n <= 0) return DGRP_OK;
  const bool open = open_tail != nullptr;
  constexpr bool VEC = sizeof(L) == 1;
  const int mis = VEC ? (int)(reinterpret_cast<uintptr_t>(d_lab) & 15u) : 0;
  const int64_t nblk = VEC ? (n + mis + SV_TILE - 1) / SV_TILE : (n + SEG_TILE - 1) / SEG_TILE;
  DGRP_CHECK(c->scan.reserve((size_t)nblk * 8 + 64));  // bstart | bend | totals[2]
  DGRP_CHECK(c->pin_small.reserve(256));
  unsigned int *bs = c->scan.as<unsigned int>();
  unsigned int *be = bs + nblk;
  unsigned long long *totals = reinterpret_cast<unsigned long long *>(be + nblk);  // 2*nblk uints: 8-byte aligned
  if (VEC)
    segv_count_kernel<<<(unsigned)nblk, SV_THREADS, 0, c->stream>>>(reinterpret_cast<const uint8_t *>(d_lab), n, mis,
                                                                    open ? 1 : 0, bs, be);
  else
    seg_count_kernel<L><<<(unsigned)nblk, SEG_THREADS, 0, c->stream>>>(d_lab, n, bs, be, open);
  seg_scan_kernel<<<1, 1024, 0, c->stream>>>(bs, be, nblk, totals);
  c->launches += 2;
  // host needs the count (and the last label for the reference's trailing zero segment)
  unsigned long long *h = c->pin_small.as<unsigned long long>();
  DGRP_CHECK(fetch_small(c, h, totals, 16));
  L *h_last = reinterpret_cast<L *>(h + 2);
  DGRP_CHECK(fetch_small(c, h_last, d_lab + (n - 1), sizeof(L)));
  DGRP_CUDA(cudaStreamSynchronize(c->stream));
  const int64_t nrun = (int64_t)h[0];
  const bool tail_zero = keep_zero && !open && (*h_last == 0);
  int64_t total = nrun + (tail_zero ? 1 : 0);
  DGRP_CHECK(c->segs.reserve((size_t)(total > 0 ? total : 1) * 24));
  int64_t *tri = c->segs.as<int64_t>();
  if (nrun > 0) {
    if (VEC)
      segv_scatter_kernel<<<(unsigned)nblk, SV_THREADS, 0, c->stream>>>(reinterpret_cast<const uint8_t *>(d_lab), n,
                                                                        mis, open ? 1 : 0, bs, be, tri, offset);
    else
      seg_scatter_kernel<L><<<(unsigned)nblk, SEG_THREADS, 0, c->stream>>>(d_lab, n, bs, be, tri, offset, open);
    c->launches++;
  }
  if (open && nrun > 0 && *h_last != 0) {
    // the last run reaches the end of the prefix: it may go on, so it is left to the next call
    int64_t *h_tail = reinterpret_cast<int64_t *>(h + 8);
    DGRP_CHECK(fetch_small(c, h_tail, tri + 3 * (nrun - 1), 8));
    DGRP_CUDA(cudaStreamSynchronize(c->stream));
    *open_tail = *h_tail - offset;
    total = nrun - 1;
  }
  if (tail_zero) {
    // sequence.pyx:44-49: skipping zeros stops at index size-1, which is then yielded as
    // (size-1, size, 0)
    int64_t t3[3] = {n - 1 + offset, n + offset, 0};
    int64_t *hp = reinterpret_cast<int64_t *>(h + 4);
    hp[0] = t3[0]; hp[1] = t3[1]; hp[2] = t3[2];
    DGRP_CUDA(cudaMemcpyAsync(tri + 3 * nrun, hp, 24, cudaMemcpyHostToDevice, c->stream));
  }
  DGRP_CUDA(cudaGetLastError());
  *d_triples = tri;
  *n_out = total;
  return DGRP_OK;
}

int run_segments(dgrp_ctx *c, const uint8_t *d_label, const int64_t *d_label64, int64_t n,
                 int64_t offset, bool keep_zero, int64_t **d_triples, int64_t *n_out, int64_t *open_tail) {
  if (d_label) return run_segments_t<uint8_t>(c, d_label, n, offset, keep_zero, d_triples, n_out, open_tail);
  return run_segments_t<int64_t>(c, d_label64, n, offset, keep_zero, d_triples, n_out, open_tail);
}

// get_segments (sequence.pyx:40-53) from `startpos`: one block; two strided searches.
__global__ void get_segments_kernel(const int64_t *__restrict__ cls, int64_t size, int64_t startpos,
                                    int64_t *out3) {
  __shared__ long long s_found;
  const int64_t last = size - 1;
  // phase 1: first p in [startpos, last) with cls[p] != 0, else max(startpos, last)... the
  // reference loop `while startpos < last && cur == 0` stops at `last` when everything is zero,
  // and does not move at all when startpos >= last.
  if (threadIdx.x == 0) s_found = LLONG_MAX;
  __syncthreads();
  int64_t p = startpos;
  if (startpos < last) {
    for (int64_t base = startpos; base < last; base += blockDim.x) {
      int64_t i = base + threadIdx.x;
      if (i < last && cls[i] != 0) atomicMin(&s_found, (long long)i);
      __syncthreads();
      const bool hit = s_found != LLONG_MAX;
      __syncthreads();
      if (hit) break;
    }
    p = s_found == LLONG_MAX ? last : (int64_t)s_found;
  }
  __syncthreads();
  const int64_t cur = cls[p];
  if (threadIdx.x == 0) s_found = LLONG_MAX;
  __syncthreads();
  // phase 2: end = p + 1; while end < last && cls[end] == cur: ++end
  int64_t end = p + 1;
  if (end < last) {
    for (int64_t base = p + 1; base < last; base += blockDim.x) {
      int64_t i = base + threadIdx.x;
      if (i < last && cls[i] != cur) atomicMin(&s_found, (long long)i);
      __syncthreads();
      const bool hit = s_found != LLONG_MAX;
      __syncthreads();
      if (hit) break;
    }
    end = s_found == LLONG_MAX ? last : (int64_t)s_found;
  }
  if (threadIdx.x == 0) { out3[0] = p; out3[1] = end; out3[2] = cur; }
}

int launch_get_segments(dgrp_ctx *c, const int64_t *d_classes, int64_t size, int64_t startpos,
                        int64_t *d_out3) {
  get_segments_kernel<<<1, 1024, 0, c->stream>>>(d_classes, size, startpos, d_out3);
  c->launches++;
  DGRP_CUDA(cudaGetLastError());
  return DGRP_OK;
}

}  // namespace dgrp
