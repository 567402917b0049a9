// K7: chunked maximal-scoring-segment scan + majority gap fill.
// Replaces deepgrp/_mss/mss.c:50-101 (mss_find_all) and deepgrp/_mss/pymss.pyx:31-80
// (_find_mss_labels).  The algorithm and why it is bit-exact are described in mss_core.cuh.
#include <algorithm>

#include "dgrp_internal.cuh"
#include "mss_core.cuh"
#include "scan_util.cuh"

namespace dgrp {

using mss::RunTable;
using mss::ScanState;

__global__ void cp_scan_kernel(unsigned int *a, int64_t m, unsigned long long *total) {
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
    if (i < m) a[i] = (unsigned int)excl;
    __syncthreads();
    if (threadIdx.x == 0) s_carry += s_warp[nwarp - 1];
    __syncthreads();
  }
  if (threadIdx.x == 0) total[0] = s_carry;
}

// count -> scan -> (host reads the total) -> scatter.  Uses c->scan for tile counts and
// c->pin_small for the total.
template <class Pred, class Emit>
static int compact(dgrp_ctx *c, int64_t n, Pred pred, Emit emit, int64_t *count) {
  *count = 0;
  if (n <= 0) return DGRP_OK;
  const int64_t ntiles = (n + CP_TILE - 1) / CP_TILE;
  DGRP_CHECK(c->scan.reserve((size_t)ntiles * 4 + 64));
  DGRP_CHECK(c->pin_small.reserve(256));
  unsigned int *tiles = c->scan.as<unsigned int>();
  unsigned long long *total =
      reinterpret_cast<unsigned long long *>(c->scan.as<unsigned char>() + ((ntiles * 4 + 15) / 16) * 16);
  DGRP_CHECK(c->scan.reserve(((size_t)ntiles * 4 + 15) / 16 * 16 + 64));
  tiles = c->scan.as<unsigned int>();
  total = reinterpret_cast<unsigned long long *>(c->scan.as<unsigned char>() + ((ntiles * 4 + 15) / 16) * 16);
  cp_count_kernel<<<(unsigned)ntiles, CP_THREADS, 0, c->stream>>>(n, pred, tiles);
  cp_scan_kernel<<<1, 1024, 0, c->stream>>>(tiles, ntiles, total);
  c->launches += 2;
  unsigned long long *h = c->pin_small.as<unsigned long long>();
  DGRP_CHECK(fetch_small(c, h, total, 8));
  DGRP_CUDA(cudaStreamSynchronize(c->stream));
  *count = (int64_t)h[0];
  if (*count > 0) {
    cp_scatter_kernel<<<(unsigned)ntiles, CP_THREADS, 0, c->stream>>>(n, pred, emit, tiles);
    c->launches++;
  }
  DGRP_CUDA(cudaGetLastError());
  return DGRP_OK;
}

// ---- stage 0: run ordinals --------------------------------------------------------------------
template <typename T>
__global__ void mss_count_runs_kernel(const T *__restrict__ S, int n, int CH, unsigned int *cnt) {
  const int gs = gridDim.x * blockDim.x;
  for (int i0 = blockIdx.x * blockDim.x; i0 < n; i0 += gs) {   // CH is a multiple of 32
    const int i = i0 + threadIdx.x;
    bool f = false;
    if (i < n) f = ((double)S[i] > 0) && (i == 0 || !((double)S[i - 1] > 0));
    const unsigned int m = __ballot_sync(0xffffffffu, f);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(&cnt[(i0 + (int)threadIdx.x) / CH], __popc(m));
  }
}
// float scores: four per thread (one 128-bit load when the array is 16-byte aligned), the score before them from
// the neighbouring lane; a warp covers 128 scores = one chunk when CH is a multiple of 128
__global__ void mss_count_runs4_kernel(const float *__restrict__ S, int n, int CH, unsigned int *cnt) {
  const int lane = threadIdx.x & 31;
  const bool aligned = (reinterpret_cast<uintptr_t>(S) & 15u) == 0;
  const bool warp_chunk = CH % 128 == 0;
  const int64_t gs = (int64_t)gridDim.x * blockDim.x * 4;
  for (int64_t b0 = (int64_t)blockIdx.x * blockDim.x * 4; b0 < n; b0 += gs) {
    const int64_t i = b0 + (int64_t)threadIdx.x * 4;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (i + 4 <= n && aligned) {
      const float4 t = *reinterpret_cast<const float4 *>(S + i);
      v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) if (i + k < n) v[k] = S[i + k];
    }
    float before = __shfl_up_sync(0xffffffffu, v[3], 1);
    if (lane == 0) before = (i > 0 && i - 1 < n) ? S[i - 1] : 0.f;
    unsigned c = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      c += (i + k < n && v[k] > 0.f && !((k ? v[k - 1] : before) > 0.f)) ? 1u : 0u;
    }
    if (warp_chunk) {
      for (int off = 16; off > 0; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
      if (lane == 0 && c) atomicAdd(&cnt[(b0 + (int64_t)(threadIdx.x & ~31) * 4) / CH], c);
    } else if (c) {
      atomicAdd(&cnt[i / CH], c);   // CH is a multiple of 32, i of 4: the thread's four scores share a chunk
    }
  }
}

// ---- stage 1: reduced scan, chunked (see mss_core.cuh item 4) --------------------------------------
using mss::ChunkSummary;

struct ScanBufs {
  ScanState *used, *out, *pred;
  ChunkSummary *sum;
  uint8_t *dirty;
  const unsigned int *base;   // [NC] first run ordinal of each chunk
  int *n_dirty;
};

// mode 0: chunk 0 from the initial state (running sum L0), every other chunk from the canonical state;
// mode 1: stale chunks from their predicted state.  store = 0: a speculative first round that leaves the run
// records to the second one (which then re-runs every chunk): its 25 bytes per run were the larger part of
// the round's memory traffic.
template <typename T>
__global__ void mss_scan_kernel(const T *__restrict__ S, int n, double xdrop, int CH, int NC,
                                int mode, double L0, int store, ScanBufs b, RunTable rt) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= NC) return;
  ScanState s;
  if (mode == 0) {
    if (c == 0) mss::state_initial(s, L0); else mss::state_canonical(s);
  } else {
    if (!b.dirty[c]) return;
    s = b.pred[c];
  }
  b.used[c] = s;
  const int lo = c * CH, hi = lo + CH < n ? lo + CH : n;
  ChunkSummary sum;
  mss::scan_chunk(S, n, xdrop, lo, hi, (int)b.base[c], s, rt, sum, store != 0);
  b.out[c] = s;
  b.sum[c] = sum;
}

// One warp walks the chunk summaries: predicts every chunk's start state and marks the chunks whose
// last execution started from something else.  Lanes stage 32 chunks at a time in shared memory,
// lane 0 does the (inherently sequential, O(1) per chunk) chain.
__global__ void mss_chain_kernel(int NC, ScanBufs b, double L0, int force) {
  __shared__ ScanState s_used[32], s_out[32], s_pred[32];
  __shared__ ChunkSummary s_sum[32];
  __shared__ uint8_t s_dirty[32];
  const int lane = threadIdx.x;
  ScanState s;
  mss::state_initial(s, L0);
  int n_dirty = 0;
  for (int base = 0; base < NC; base += 32) {
    const int c = base + lane;
    if (c < NC) { s_used[lane] = b.used[c]; s_out[lane] = b.out[c]; s_sum[lane] = b.sum[c]; }
    __syncwarp();
    if (lane == 0) {
      const int m = NC - base < 32 ? NC - base : 32;
      for (int k = 0; k < m; ++k) {
        s_pred[k] = s;
        const bool same = mss::state_equal(s_used[k], s);
        if (same && !force) s_dirty[k] = 0; else { s_dirty[k] = 1; ++n_dirty; }   // force: nothing has been stored yet
        s = same ? s_out[k] : mss::apply_summary(s_sum[k], s_used[k], s_out[k], s);
      }
    }
    __syncwarp();
    if (c < NC) { b.pred[c] = s_pred[lane]; b.dirty[c] = s_dirty[lane]; }
    __syncwarp();
  }
  if (lane == 0) *b.n_dirty = n_dirty;
}

// The same chain in three parallel-friendly steps (used for float32 scores, whose sums are exact so that
// composed predictions equal the sequential ones): (1) one thread per group of MSS_GROUP chunks
// composes the group's chunk effects (mss::compose), (2) one thread walks the group composites -- NC / 32
// steps instead of NC -- and leaves every group's start state, (3) one thread per group walks its own
// chunks from that state, predicting their start states and marking the stale ones.
constexpr int MSS_GROUP = 32;

__global__ void mss_group_compose_kernel(int NC, ScanBufs b, mss::Composite *comp) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  const int c0 = g * MSS_GROUP;
  if (c0 >= NC) return;
  const int c1 = c0 + MSS_GROUP < NC ? c0 + MSS_GROUP : NC;
  mss::Composite acc;
  acc.sum = b.sum[c0]; acc.x_in = b.used[c0]; acc.x_out = b.out[c0];
  for (int c = c0 + 1; c < c1; ++c) {
    mss::Composite nx;
    nx.sum = b.sum[c]; nx.x_in = b.used[c]; nx.x_out = b.out[c];
    acc = mss::compose(acc, nx);
  }
  comp[g] = acc;
}

__global__ void mss_group_chain_kernel(int NG, const mss::Composite *comp, ScanState *gstart, int *n_dirty,
                                       double L0) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  ScanState s;
  mss::state_initial(s, L0);
  for (int g = 0; g < NG; ++g) {
    gstart[g] = s;
    const mss::Composite k = comp[g];
    s = mss::state_equal(k.x_in, s) ? k.x_out : mss::apply_summary(k.sum, k.x_in, k.x_out, s);
  }
  *n_dirty = 0;
}

// second level: composites of MSS_GROUP group composites, and the walk back down to the group starts
__global__ void mss_super_compose_kernel(int NG, const mss::Composite *comp, mss::Composite *super) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  const int g0 = q * MSS_GROUP;
  if (g0 >= NG) return;
  const int g1 = g0 + MSS_GROUP < NG ? g0 + MSS_GROUP : NG;
  mss::Composite acc = comp[g0];
  for (int g = g0 + 1; g < g1; ++g) acc = mss::compose(acc, comp[g]);
  super[q] = acc;
}

__global__ void mss_super_fill_kernel(int NG, const mss::Composite *comp, const ScanState *sstart,
                                      ScanState *gstart) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  const int g0 = q * MSS_GROUP;
  if (g0 >= NG) return;
  const int g1 = g0 + MSS_GROUP < NG ? g0 + MSS_GROUP : NG;
  ScanState s = sstart[q];
  for (int g = g0; g < g1; ++g) {
    gstart[g] = s;
    const mss::Composite k = comp[g];
    s = mss::state_equal(k.x_in, s) ? k.x_out : mss::apply_summary(k.sum, k.x_in, k.x_out, s);
  }
}

__global__ void mss_group_fill_kernel(int NC, ScanBufs b, const ScanState *gstart, int force) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  const int c0 = g * MSS_GROUP;
  if (c0 >= NC) return;
  const int c1 = c0 + MSS_GROUP < NC ? c0 + MSS_GROUP : NC;
  ScanState s = gstart[g];
  int n_dirty = 0;
  for (int c = c0; c < c1; ++c) {
    b.pred[c] = s;
    const ScanState used = b.used[c], out = b.out[c];
    const bool same = mss::state_equal(used, s);
    if (same && !force) b.dirty[c] = 0; else { b.dirty[c] = 1; ++n_dirty; }
    s = same ? out : mss::apply_summary(b.sum[c], used, out, s);
  }
  if (n_dirty) atomicAdd(b.n_dirty, n_dirty);
}

// Parallel verification of the fixed point: every chunk must have started from exactly the state its
// predecessor ended in (chunk 0 from the canonical state).  Counts the chunks for which that fails.
__global__ void mss_verify_kernel(int NC, ScanBufs b, double L0) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= NC) return;
  ScanState expect;
  if (c == 0) mss::state_initial(expect, L0);
  else expect = b.out[c - 1];
  if (!mss::state_equal(b.used[c], expect)) atomicAdd(b.n_dirty, 1);
}

// Sequential completion by one thread (when predictions keep missing, e.g. arbitrary doubles whose
// sums round in every chunk): walk the chain, re-running every chunk that is stale.
template <typename T>
__global__ void mss_complete_kernel(const T *__restrict__ S, int n, double xdrop, int CH, int NC,
                                    double L0, ScanBufs b, RunTable rt) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  ScanState s;
  mss::state_initial(s, L0);
  for (int c = 0; c < NC; ++c) {
    if (mss::state_equal(b.used[c], s)) { s = b.out[c]; continue; }
    b.used[c] = s;
    const int lo = c * CH, hi = lo + CH < n ? lo + CH : n;
    ChunkSummary sum;
    mss::scan_chunk(S, n, xdrop, lo, hi, (int)b.base[c], s, rt, sum);
    b.out[c] = s;
    b.sum[c] = sum;
  }
}

// ---- stage 2: regions ----------------------------------------------------------------------------
struct IsEvent {
  const uint8_t *kind;
  __device__ bool operator()(int64_t k) const { return kind[k] != mss::RUN_PLAIN; }
};
struct EmitEvent {
  int *ev;
  __device__ void operator()(int64_t k, int64_t pos) const { ev[pos] = (int)k; }
};

__global__ void mss_regions_kernel(const int *__restrict__ ev, int n_ev, int NR, int min_sc_int,
                                   RunTable rt, uint8_t *live) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_ev) return;
  const int k0 = ev[r], k1 = r + 1 < n_ev ? ev[r + 1] : NR;
  // the event kind of k1 must be read BEFORE region r+1's thread overwrites slot k1?  It does not:
  // kind[] is never written here, only st/en/L/R/pre of slots inside the own region.
  const int depth = mss::process_region(k0, k1, rt);
  const bool flushed = (k1 == NR) || rt.kind[k1] == mss::RUN_FLUSH;
  for (int j = 0; j < k1 - k0; ++j)
    live[k0 + j] = (j < depth && flushed && (rt.R[k0 + j] - rt.L[k0 + j] >= (double)min_sc_int)) ? 1 : 0;
}

struct IsLive {
  const uint8_t *live;
  __device__ bool operator()(int64_t k) const { return live[k] != 0; }
};
struct EmitSeg {
  RunTable rt;
  dgrp_seg_t *out;
  __device__ void operator()(int64_t k, int64_t pos) const {
    dgrp_seg_t s;
    s.st = rt.st[k]; s.en = rt.en[k]; s.sc = rt.R[k] - rt.L[k];
    out[pos] = s;
  }
};

// ---- open end (a prefix of a record) ---------------------------------------------------------------
// The last FLUSH run of the table: there the reference empties its stack whatever follows (mss.c:78-81), so the
// candidates flushed up to it are final although the record goes on, and the record can be resumed at the
// run's first element with the running sum before it.
struct MssLast {
  int k;       // ordinal of the last FLUSH run, -1: none
  int st;      // its first element
  double L;    // the running sum before it
};
__global__ void mss_last_flush_kernel(const uint8_t *__restrict__ kind, int NR, int *last) {
  int best = -1;
  const int gs = gridDim.x * blockDim.x;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < NR; k += gs)
    if (kind[k] == mss::RUN_FLUSH) best = k;   // k grows along the loop
  for (int off = 16; off > 0; off >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, off));
  if ((threadIdx.x & 31) == 0 && best >= 0) atomicMax(last, best);
}
__global__ void mss_last_record_kernel(const int *last, RunTable rt, MssLast *out) {
  const int k = *last;
  out->k = k;
  out->st = k >= 0 ? rt.st[k] : 0;
  out->L = k >= 0 ? rt.L[k] : 0.0;
}

static inline size_t align256(size_t x) { return (x + 255) / 256 * 256; }

template <typename T>
static int run_mss_t(dgrp_ctx *c, const T *d_S, int n, double min_sc, double xdrop,
                     dgrp_seg_t **d_segs_out, int *n_seg, double L0, MssResume *resume) {
  *n_seg = 0;
  *d_segs_out = nullptr;
  if (resume) { resume->restart = -1; resume->L0 = L0; }
  if (n <= 0) return DGRP_OK;
  int CH = c->mss_chunk;
  if (CH <= 0) {
    // The scan is sequential inside a chunk (~0.25 us per score, twice); the summary pass is sequential
    // over chunks (~0.25 us each, float64 scores) or over groups of 32 chunks (float32 scores): chunk ~
    // sqrt(n) / 2 resp. sqrt(n / 256) balances the two (512 at 46.7 M, 1024 at 248 M).
    // (round 2, with the lean first round: 512 instead of 1024 at 46.7 M is 1.48 instead of 1.66 ms, profiles/r03c_*)
    const int64_t per = sizeof(T) == 4 ? 8 * MSS_GROUP : 4;
    CH = 64;
    while (CH < 16384 && (int64_t)CH * CH * per < (int64_t)n) CH <<= 1;
  }
  CH = (CH + 31) / 32 * 32;
  const int NC = (n + CH - 1) / CH;

  // ---- run ordinals
  DGRP_CHECK(c->mss_a.reserve((size_t)(NC + 1) * 4 + 64));
  unsigned int *base = c->mss_a.as<unsigned int>();
  DGRP_CUDA(cudaMemsetAsync(base, 0, (size_t)(NC + 1) * 4, c->stream));
  {
    const int threads = 256;
    int64_t want = ((int64_t)n + threads - 1) / threads;
    const int blocks = (int)(want < (int64_t)c->sm_count * 16 ? want : (int64_t)c->sm_count * 16);
    if (sizeof(T) == 4) {
      want = ((int64_t)n + threads * 4 - 1) / (threads * 4);
      const int blocks4 = (int)(want < (int64_t)c->sm_count * 16 ? want : (int64_t)c->sm_count * 16);
      mss_count_runs4_kernel<<<blocks4, threads, 0, c->stream>>>(reinterpret_cast<const float *>(d_S), n, CH, base);
    } else {
      mss_count_runs_kernel<T><<<blocks, threads, 0, c->stream>>>(d_S, n, CH, base);
    }
    c->launches++;
  }
  DGRP_CHECK(c->pin_small.reserve(256));
  DGRP_CHECK(c->small.reserve(256));
  unsigned long long *d_total = c->small.as<unsigned long long>() + 8;
  cp_scan_kernel<<<1, 1024, 0, c->stream>>>(base, NC, d_total);
  c->launches++;
  unsigned long long *h = c->pin_small.as<unsigned long long>();
  DGRP_CHECK(fetch_small(c, h, d_total, 8));
  DGRP_CUDA(cudaStreamSynchronize(c->stream));
  int NR = (int)h[0];
  if (NR == 0) return DGRP_OK;   // no positive score anywhere: no segments (and nowhere to resume)

  // ---- buffers
  const size_t nr1 = (size_t)NR + 1;
  size_t off = 0;
  const size_t o_st = off; off = align256(off + nr1 * 4);
  const size_t o_en = off; off = align256(off + nr1 * 4);
  const size_t o_pre = off; off = align256(off + nr1 * 4);
  const size_t o_L = off; off = align256(off + nr1 * 8);
  const size_t o_R = off; off = align256(off + nr1 * 8);
  const size_t o_kind = off; off = align256(off + nr1);
  const size_t o_live = off; off = align256(off + nr1);
  const size_t o_ev = off; off = align256(off + nr1 * 4);
  DGRP_CHECK(c->mss_b.reserve(off));
  unsigned char *pb = c->mss_b.as<unsigned char>();
  RunTable rt;
  rt.st = reinterpret_cast<int *>(pb + o_st);
  rt.en = reinterpret_cast<int *>(pb + o_en);
  rt.pre = reinterpret_cast<int *>(pb + o_pre);
  rt.L = reinterpret_cast<double *>(pb + o_L);
  rt.R = reinterpret_cast<double *>(pb + o_R);
  rt.kind = pb + o_kind;
  uint8_t *live = pb + o_live;
  int *ev = reinterpret_cast<int *>(pb + o_ev);

  off = 0;
  const size_t o_used = off; off = align256(off + (size_t)NC * sizeof(ScanState));
  const size_t o_out = off; off = align256(off + (size_t)NC * sizeof(ScanState));
  const size_t o_pred = off; off = align256(off + (size_t)NC * sizeof(ScanState));
  const size_t o_sum = off; off = align256(off + (size_t)NC * sizeof(ChunkSummary));
  const size_t o_dirty = off; off = align256(off + (size_t)NC);
  const size_t o_cnt = off; off = align256(off + 64);
  DGRP_CHECK(c->mss_c.reserve(off));
  unsigned char *pc = c->mss_c.as<unsigned char>();
  ScanBufs sb;
  sb.used = reinterpret_cast<ScanState *>(pc + o_used);
  sb.out = reinterpret_cast<ScanState *>(pc + o_out);
  sb.pred = reinterpret_cast<ScanState *>(pc + o_pred);
  sb.sum = reinterpret_cast<ChunkSummary *>(pc + o_sum);
  sb.dirty = pc + o_dirty;
  sb.n_dirty = reinterpret_cast<int *>(pc + o_cnt);
  sb.base = base;

  // ---- stage 1
  const int threads = 128;
  const int blocks = (NC + threads - 1) / threads;
  // (Staging the scores through shared memory -- coalesced cp.async tiles, a block in lock step -- was measured
  // SLOWER: 12.2 instead of 7.8 ms for the 248 Mbp finish.  The scan is bound by the instructions of the state
  // machine, a run ending every 2-4 scores with random-init weights, not by its uncoalesced loads.)
  const int max_rounds = c->mss_max_rounds > 0 ? c->mss_max_rounds : 32;   // a round costs ~1 ms per 50 M scores, the sequential completion seconds
  // With more than one chunk nearly every chunk of the first round starts from a guess and is run again: the
  // first round then only produces the summaries ("lean") and the first chain pass marks every chunk stale.
  const bool lean = NC > 1 && max_rounds >= 2;
  auto launch_scan = [&](int mode, int store) {
    mss_scan_kernel<T><<<blocks, threads, 0, c->stream>>>(d_S, n, xdrop, CH, NC, mode, L0, store, sb, rt);
    c->launches++;
  };
  launch_scan(0, lean ? 0 : 1);
  int *h_dirty = reinterpret_cast<int *>(h + 2);
  int rounds = 1;
  bool converged = NC == 1;
  while (!converged) {
    const int force = (lean && rounds == 1) ? 1 : 0;
    // predict all start states from the chunk summaries ...
    if (sizeof(T) == 4 && NC > 4 * MSS_GROUP) {
      // ... float32 scores: composed per group in parallel, a sequential pass over the groups only
      // (two levels: groups of 32 chunks, super-groups of 32 groups; only the super-groups are walked
      // by a single thread -- NC / 1024 steps)
      const int NG = (NC + MSS_GROUP - 1) / MSS_GROUP, NS = (NG + MSS_GROUP - 1) / MSS_GROUP;
      const size_t o_g = align256((size_t)NG * sizeof(mss::Composite));
      const size_t o_s = o_g + align256((size_t)NG * sizeof(ScanState));
      const size_t o_ss = o_s + align256((size_t)NS * sizeof(mss::Composite));
      DGRP_CHECK(c->mss_d.reserve(o_ss + (size_t)NS * sizeof(ScanState) + 512));
      unsigned char *pd = c->mss_d.as<unsigned char>();
      mss::Composite *comp = reinterpret_cast<mss::Composite *>(pd);
      ScanState *gstart = reinterpret_cast<ScanState *>(pd + o_g);
      mss::Composite *super = reinterpret_cast<mss::Composite *>(pd + o_s);
      ScanState *sstart = reinterpret_cast<ScanState *>(pd + o_ss);
      const int gb = (NG + 63) / 64, sbk = (NS + 63) / 64;
      mss_group_compose_kernel<<<gb, 64, 0, c->stream>>>(NC, sb, comp);
      mss_super_compose_kernel<<<sbk, 64, 0, c->stream>>>(NG, comp, super);
      mss_group_chain_kernel<<<1, 32, 0, c->stream>>>(NS, super, sstart, sb.n_dirty, L0);
      mss_super_fill_kernel<<<sbk, 64, 0, c->stream>>>(NG, comp, sstart, gstart);
      mss_group_fill_kernel<<<gb, 64, 0, c->stream>>>(NC, sb, gstart, force);
      c->launches += 5;
    } else {
      // ... sequential over chunks, O(1) each (float64 scores: sums round, composed shifts would miss)
      mss_chain_kernel<<<1, 32, 0, c->stream>>>(NC, sb, L0, force);
      c->launches++;
    }
    DGRP_CHECK(fetch_small(c, h_dirty, sb.n_dirty, 4));
    DGRP_CUDA(cudaStreamSynchronize(c->stream));
    if (*h_dirty == 0) { converged = true; break; }
    if (rounds >= max_rounds) break;
    // ... re-run the stale chunks in parallel ...
    launch_scan(1, 1);
    ++rounds;
    // ... and verify the chain in parallel; only a failed check needs another sequential pass
    DGRP_CUDA(cudaMemsetAsync(sb.n_dirty, 0, 4, c->stream));
    mss_verify_kernel<<<blocks, threads, 0, c->stream>>>(NC, sb, L0);
    c->launches++;
    DGRP_CHECK(fetch_small(c, h_dirty, sb.n_dirty, 4));
    DGRP_CUDA(cudaStreamSynchronize(c->stream));
    if (*h_dirty == 0) { converged = true; break; }
  }
  if (!converged) {
    mss_complete_kernel<T><<<1, 32, 0, c->stream>>>(d_S, n, xdrop, CH, NC, L0, sb, rt);
    c->launches++;
  }
  c->mss_rounds = converged ? rounds : -rounds;

  if (resume) {
    // open end: everything from the last FLUSH run on is left to the next call (see MssLast)
    int *d_last = reinterpret_cast<int *>(c->small.as<unsigned char>() + 128);
    MssLast *d_rec = reinterpret_cast<MssLast *>(c->small.as<unsigned char>() + 144);
    DGRP_CUDA(cudaMemsetAsync(d_last, 0xff, 4, c->stream));
    const int lb = (int)std::min<int64_t>(((int64_t)NR + 255) / 256, (int64_t)c->sm_count * 8);
    mss_last_flush_kernel<<<lb, 256, 0, c->stream>>>(rt.kind, NR, d_last);
    mss_last_record_kernel<<<1, 1, 0, c->stream>>>(d_last, rt, d_rec);
    c->launches += 2;
    MssLast *h_rec = reinterpret_cast<MssLast *>(h + 4);
    DGRP_CHECK(fetch_small(c, h_rec, d_rec, sizeof(MssLast)));
    DGRP_CUDA(cudaStreamSynchronize(c->stream));
    if (h_rec->k < 0) return DGRP_OK;
    resume->restart = h_rec->st;
    resume->L0 = h_rec->L;
    NR = h_rec->k;               // the runs before it: their last region ends at a FLUSH event
    if (NR == 0) return DGRP_OK;
  }

  // ---- stage 2
  int64_t n_ev = 0;
  DGRP_CHECK(compact(c, NR, IsEvent{rt.kind}, EmitEvent{ev}, &n_ev));
  if (n_ev > 0) {
    const int rb = (int)((n_ev + threads - 1) / threads);
    mss_regions_kernel<<<rb, threads, 0, c->stream>>>(ev, (int)n_ev, NR, (int)min_sc, rt, live);
    c->launches++;
  }
  int64_t n_live = 0;
  // the segment list is written behind a generous reservation: count first
  DGRP_CHECK(c->segs.reserve(256));
  {
    // two-step: count (inside compact) needs the output buffer only for scatter, so reserve after
    // counting.  compact() scatters immediately, hence reserve for the worst case NR here.
    DGRP_CHECK(c->mss_e.reserve(nr1 * sizeof(dgrp_seg_t)));
    DGRP_CHECK(compact(c, NR, IsLive{live}, EmitSeg{rt, c->mss_e.as<dgrp_seg_t>()}, &n_live));
  }
  DGRP_CUDA(cudaGetLastError());
  *d_segs_out = c->mss_e.as<dgrp_seg_t>();
  *n_seg = (int)n_live;
  return DGRP_OK;
}

int run_mss_segments(dgrp_ctx *c, const double *d_s64, const float *d_s32, int n, double min_sc,
                     double xdrop, dgrp_seg_t **d_segs_out, int *n_seg, double L0, MssResume *resume) {
  if (d_s64) return run_mss_t<double>(c, d_s64, n, min_sc, xdrop, d_segs_out, n_seg, L0, resume);
  return run_mss_t<float>(c, d_s32, n, min_sc, xdrop, d_segs_out, n_seg, L0, resume);
}

// ---- gap fill (pymss.pyx:59-77) -------------------------------------------------------------------
// One warp per segment: count labels 1..nof-1, majority (ties -> lowest, default 1), rewrite zeros.
constexpr int GF_MAXC = 16;

// Segments longer than GF_LONG positions leave the warp-per-segment kernel (one warp walking a segment
// of millions of positions is the whole step: 385 ms for one 24.8 Mbp segment) and are processed
// position-parallel by gap_long_kernel, GF_PIECE positions per block.  Disjoint segments longer than
// GF_LONG cannot start inside the same aligned window of GF_LONG positions, so st / GF_LONG is a
// collision-free key for their label counts.
constexpr int GF_LONG = 8192;
constexpr int GF_PIECE = 8192;

template <typename LabT>
__global__ void gap_fill_kernel(const dgrp_seg_t *__restrict__ segs, int n_seg,
                                const LabT *__restrict__ lab_in, int nof,
                                uint8_t *__restrict__ lab_out, int *__restrict__ any_long) {
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  for (int g = blockIdx.x * warps_per_block + (threadIdx.x >> 5); g < n_seg;
       g += gridDim.x * warps_per_block) {
    const int st = segs[g].st, en = segs[g].en;
    if (en - st > GF_LONG) {
      if (lane == 0) *any_long = 1;
      continue;
    }
    unsigned int cnt[GF_MAXC];
#pragma unroll
    for (int k = 0; k < GF_MAXC; ++k) cnt[k] = 0;
    for (int i = st + lane; i < en; i += 32) {
      const int l = (int)lab_in[i];
#pragma unroll
      for (int k = 1; k < GF_MAXC; ++k) cnt[k] += (l == k);
    }
#pragma unroll
    for (int k = 1; k < GF_MAXC; ++k)
      for (int off = 16; off > 0; off >>= 1) cnt[k] += __shfl_xor_sync(0xffffffffu, cnt[k], off);
    int best = 1;
    unsigned int bestv = cnt[1];
#pragma unroll
    for (int k = 2; k < GF_MAXC; ++k)
      if (k < nof && bestv < cnt[k]) { best = k; bestv = cnt[k]; }
    for (int i = st + lane; i < en; i += 32)
      if (lab_in[i] == 0) lab_out[i] = (uint8_t)best;
  }
}

// Long segments, position-parallel: block b owns positions [b * GF_PIECE, (b + 1) * GF_PIECE) and visits the
// long segments that intersect them (found by bisection; segments are sorted and disjoint).  FILL = false
// adds the label counts of the intersection to the segment's table entry, FILL = true (a second launch)
// takes the majority over classes 1..nof-1 (ties -> lowest, default 1, pymss.pyx:62-67) and rewrites
// the zeros.  Every condition on the way is block-uniform.
template <typename LabT, bool FILL>
__global__ void gap_long_kernel(const dgrp_seg_t *__restrict__ segs, int n_seg,
                                const LabT *__restrict__ lab_in, int n, int nof,
                                const int *__restrict__ any_long, unsigned int *__restrict__ table,
                                uint8_t *__restrict__ lab_out) {
  if (*any_long == 0) return;
  __shared__ unsigned int s_cnt[GF_MAXC];
  const int64_t r0 = (int64_t)blockIdx.x * GF_PIECE;
  const int64_t r1 = r0 + GF_PIECE < (int64_t)n ? r0 + GF_PIECE : (int64_t)n;
  int lo = 0, hi = n_seg;                      // first segment with en > r0
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if ((int64_t)segs[mid].en > r0) hi = mid; else lo = mid + 1;
  }
  for (int g = lo; g < n_seg; ++g) {
    const int st = segs[g].st, en = segs[g].en;
    if ((int64_t)st >= r1) break;
    if (en - st <= GF_LONG) continue;
    const int a = (int64_t)st > r0 ? st : (int)r0, b = (int64_t)en < r1 ? en : (int)r1;
    unsigned int *entry = table + (size_t)(st / GF_LONG) * GF_MAXC;
    if (!FILL) {
      if (threadIdx.x < GF_MAXC) s_cnt[threadIdx.x] = 0;
      __syncthreads();
      unsigned int cnt[GF_MAXC];
#pragma unroll
      for (int k = 0; k < GF_MAXC; ++k) cnt[k] = 0;
      for (int i = a + (int)threadIdx.x; i < b; i += (int)blockDim.x) {
        const int l = (int)lab_in[i];
#pragma unroll
        for (int k = 1; k < GF_MAXC; ++k) cnt[k] += (l == k);
      }
#pragma unroll
      for (int k = 1; k < GF_MAXC; ++k) {
        for (int off = 16; off > 0; off >>= 1) cnt[k] += __shfl_xor_sync(0xffffffffu, cnt[k], off);
        if ((threadIdx.x & 31) == 0 && cnt[k]) atomicAdd(&s_cnt[k], cnt[k]);
      }
      __syncthreads();
      if (threadIdx.x >= 1 && threadIdx.x < GF_MAXC && s_cnt[threadIdx.x])
        atomicAdd(&entry[threadIdx.x], s_cnt[threadIdx.x]);
      __syncthreads();
    } else {
      int best = 1;
      unsigned int bestv = entry[1];
      for (int k = 2; k < GF_MAXC; ++k) {
        const unsigned int v = entry[k];
        if (k < nof && bestv < v) { best = k; bestv = v; }
      }
      for (int i = a + (int)threadIdx.x; i < b; i += (int)blockDim.x)
        if (lab_in[i] == 0) lab_out[i] = (uint8_t)best;
    }
  }
}

template <typename LabT>
__global__ void copy_labels_kernel(const LabT *__restrict__ in, int64_t n, uint8_t *__restrict__ out) {
  const int64_t gs = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gs)
    out[i] = (uint8_t)in[i];
}
// 16 labels per thread where source and destination share their 16-byte phase; the edges byte by byte
__global__ void copy_labels16_kernel(const uint8_t *__restrict__ in, int64_t n, uint8_t *__restrict__ out) {
  const int64_t head = (16 - (int64_t)(reinterpret_cast<uintptr_t>(in) & 15u)) & 15;
  const int64_t h = head < n ? head : n;
  const int64_t words = (n - h) / 16;
  const int64_t gs = (int64_t)gridDim.x * blockDim.x, t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint4 *src = reinterpret_cast<const uint4 *>(in + h);
  uint4 *dst = reinterpret_cast<uint4 *>(out + h);
  for (int64_t i = t; i < words; i += gs) dst[i] = src[i];
  for (int64_t i = t; i < h; i += gs) out[i] = in[i];
  for (int64_t i = h + words * 16 + t; i < n; i += gs) out[i] = in[i];
}

int run_gap_fill(dgrp_ctx *c, const dgrp_seg_t *d_segs, int n_seg, const uint8_t *d_label_in,
                 const int64_t *d_label64_in, int n, int nof_labels, uint8_t *d_label_out) {
  if (n <= 0) return DGRP_OK;
  DGRP_REQUIRE(nof_labels >= 2 && nof_labels <= GF_MAXC, "nof_labels must be in [2, %d]", GF_MAXC);
  const int threads = 256;
  int64_t want = ((int64_t)n + threads - 1) / threads;
  int blocks = (int)(want < (int64_t)c->sm_count * 16 ? want : (int64_t)c->sm_count * 16);
  if (d_label_in) {
    if (d_label_in != d_label_out) {
      if (((reinterpret_cast<uintptr_t>(d_label_in) ^ reinterpret_cast<uintptr_t>(d_label_out)) & 15u) == 0)
        copy_labels16_kernel<<<blocks, threads, 0, c->stream>>>(d_label_in, n, d_label_out);
      else
        copy_labels_kernel<uint8_t><<<blocks, threads, 0, c->stream>>>(d_label_in, n, d_label_out);
    }
  } else {
    copy_labels_kernel<int64_t><<<blocks, threads, 0, c->stream>>>(d_label64_in, n, d_label_out);
  }
  c->launches++;
  if (n_seg > 0) {
    // [0] "a long segment exists" flag, [64..] label counts of the long segments keyed by st / GF_LONG
    const size_t entries = (size_t)n / GF_LONG + 1;
    const size_t table_bytes = 256 + entries * GF_MAXC * sizeof(unsigned int);
    DGRP_CHECK(c->gapfill.reserve(table_bytes));
    DGRP_CUDA(cudaMemsetAsync(c->gapfill.p, 0, table_bytes, c->stream));
    int *any_long = c->gapfill.as<int>();
    unsigned int *table = c->gapfill.as<unsigned int>() + 64;
    want = ((int64_t)n_seg + 7) / 8;
    blocks = (int)(want < (int64_t)c->sm_count * 8 ? want : (int64_t)c->sm_count * 8);
    const int pieces = (int)(((int64_t)n + GF_PIECE - 1) / GF_PIECE);
    if (d_label_in) {
      gap_fill_kernel<uint8_t><<<blocks, threads, 0, c->stream>>>(d_segs, n_seg, d_label_in, nof_labels, d_label_out, any_long);
      gap_long_kernel<uint8_t, false><<<pieces, threads, 0, c->stream>>>(d_segs, n_seg, d_label_in, n, nof_labels, any_long, table, d_label_out);
      gap_long_kernel<uint8_t, true><<<pieces, threads, 0, c->stream>>>(d_segs, n_seg, d_label_in, n, nof_labels, any_long, table, d_label_out);
    } else {
      gap_fill_kernel<int64_t><<<blocks, threads, 0, c->stream>>>(d_segs, n_seg, d_label64_in, nof_labels, d_label_out, any_long);
      gap_long_kernel<int64_t, false><<<pieces, threads, 0, c->stream>>>(d_segs, n_seg, d_label64_in, n, nof_labels, any_long, table, d_label_out);
      gap_long_kernel<int64_t, true><<<pieces, threads, 0, c->stream>>>(d_segs, n_seg, d_label64_in, n, nof_labels, any_long, table, d_label_out);
    }
    c->launches += 3;
  }
  DGRP_CUDA(cudaGetLastError());
  return DGRP_OK;
}

__global__ void labels_to_onehot_kernel(const uint8_t *__restrict__ lab, int64_t n, int C,
                                        double *__restrict__ out) {
  const int64_t total = n * C;
  const int64_t gs = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gs) {
    const int64_t i = e / C;
    out[e] = ((int)lab[i] == (int)(e - i * C)) ? 1.0 : 0.0;
  }
}

int launch_labels_to_onehot(dgrp_ctx *c, const uint8_t *d_label, int64_t n, int C, double *d_out) {
  if (n <= 0) return DGRP_OK;
  const int threads = 256;
  int64_t want = (n * C + threads - 1) / threads;
  const int blocks = (int)(want < (int64_t)c->sm_count * 16 ? want : (int64_t)c->sm_count * 16);
  labels_to_onehot_kernel<<<blocks, threads, 0, c->stream>>>(d_label, n, C, d_out);
  c->launches++;
  DGRP_CUDA(cudaGetLastError());
  return DGRP_OK;
}

}  // namespace dgrp
