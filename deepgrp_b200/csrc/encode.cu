// K1 (per-record form): edge-'N' trim, one-hot planes, base codes.
// Replaces deepgrp/sequence.pyx:11-36 (ONEHOT table + _one_hot_encode_dna_sequence).
#include "dgrp_internal.cuh"

namespace dgrp {

// A/a->0, C/c->1, G/g->2, T/t->3, everything else -> 4 (the 128-entry ONEHOT table of
// sequence.pyx:11-17; bytes >= 128 are out of the reference's table and map to 4 here).
__device__ __forceinline__ int base_code(unsigned int b) {
  unsigned int u = b & 0xDFu;  // fold ASCII case
  int c = 4;
  c = (u == 'A') ? 0 : c;
  c = (u == 'C') ? 1 : c;
  c = (u == 'G') ? 2 : c;
  c = (u == 'T') ? 3 : c;
  // bytes whose folded value collides ('!' etc. fold to control chars, never to ACGT) are safe:
  // only 0x41/0x61, 0x43/0x63, 0x47/0x67, 0x54/0x74 fold to A, C, G, T.
  return c;
}

// first_last[0] = min index whose byte is not an edge-N candidate, first_last[1] = max such index
// (init: first = n, last = -1).  sequence.pyx:27-30 trims upper-case 'N' only; fold_case adds 'n'
// because the CLI upper-cases the record before encoding (__main__.py:41).
__global__ void trim_kernel(const uint8_t *__restrict__ seq, int64_t n, int fold_case,
                            unsigned long long *first_last) {
  int64_t lo = n, hi = -1;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  // 16 bytes per load over the 16-byte aligned body; the unaligned head and the tail go byte by byte
  int64_t head = (int64_t)((16 - (reinterpret_cast<uintptr_t>(seq) & 15)) & 15);
  head = head < n ? head : n;
  const int64_t nvec = (n - head) / 16;
  const uint4 *body = reinterpret_cast<const uint4 *>(seq + head);
  const unsigned int fold = fold_case ? 0xDFDFDFDFu : 0xFFFFFFFFu;   // 'n' ^ 'N' == 0x20
  for (int64_t v = tid; v < nvec; v += stride) {
    const uint4 q = __ldg(body + v);
    const unsigned int w[4] = {q.x, q.y, q.z, q.w};
    unsigned int m = 0;                                              // bit j: byte j is not an edge-N candidate
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const unsigned int x = (w[k] ^ 0x4E4E4E4Eu) & fold;             // zero byte <=> 'N' (or 'n')
#pragma unroll
      for (int b = 0; b < 4; ++b) m |= (((x >> (8 * b)) & 0xffu) != 0u ? 1u : 0u) << (4 * k + b);
    }
    if (m) {
      const int64_t at = head + v * 16;
      const int64_t f = at + (__ffs(m) - 1), l = at + (31 - __clz(m));
      lo = lo < f ? lo : f;
      hi = hi > l ? hi : l;
    }
  }
  const int64_t tail0 = head + nvec * 16;
  for (int64_t j = tid; j < head + (n - tail0); j += stride) {
    const int64_t i = j < head ? j : tail0 + (j - head);
    unsigned int b = seq[i];
    bool is_n = (b == 'N') || (fold_case && b == 'n');
    if (!is_n) {
      lo = lo < i ? lo : i;
      hi = hi > i ? hi : i;
    }
  }
  for (int off = 16; off > 0; off >>= 1) {
    int64_t olo = __shfl_xor_sync(0xffffffffu, lo, off);
    int64_t ohi = __shfl_xor_sync(0xffffffffu, hi, off);
    lo = lo < olo ? lo : olo;
    hi = hi > ohi ? hi : ohi;
  }
  if ((threadIdx.x & 31) == 0) {
    if (lo < n) atomicMin(first_last, (unsigned long long)lo);
    if (hi >= 0) atomicMax(first_last + 1, (unsigned long long)(hi + 1));  // store last+1 (0 = none)
  }
}

__global__ void trim_init_kernel(unsigned long long *first_last, int64_t n) {
  first_last[0] = (unsigned long long)n;
  first_last[1] = 0ull;
}

int launch_trim(dgrp_ctx *c, const uint8_t *d_seq, int64_t n, int fold_case, int64_t *d_first_last) {
  trim_init_kernel<<<1, 1, 0, c->stream>>>((unsigned long long *)d_first_last, n);
  c->launches++;
  if (n > 0) {
    int threads = 256;
    int64_t want = (n / 16 + threads) / threads;
    int blocks = (int)(want < (int64_t)c->sm_count * 16 ? want : (int64_t)c->sm_count * 16);
    trim_kernel<<<blocks, threads, 0, c->stream>>>(d_seq, n, fold_case,
                                                   (unsigned long long *)d_first_last);
    c->launches++;
  }
  DGRP_CUDA(cudaGetLastError());
  return DGRP_OK;
}

// fwd[c, i] = (code(seq[start+i]) == c), int8 [5, len] C order (sequence.pyx:32-35).
__global__ void onehot_kernel(const uint8_t *__restrict__ seq, int64_t start, int64_t len,
                              int8_t *__restrict__ fwd) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += stride) {
    int code = base_code(seq[start + i]);
#pragma unroll
    for (int ch = 0; ch < 5; ++ch) fwd[(int64_t)ch * len + i] = (int8_t)(code == ch);
  }
}

int launch_onehot(dgrp_ctx *c, const uint8_t *d_seq, int64_t start, int64_t len, int8_t *d_fwd) {
  if (len <= 0) return DGRP_OK;
  int threads = 256;
  int64_t want = (len + threads - 1) / threads;
  int blocks = (int)(want < (int64_t)c->sm_count * 32 ? want : (int64_t)c->sm_count * 32);
  onehot_kernel<<<blocks, threads, 0, c->stream>>>(d_seq, start, len, d_fwd);
  c->launches++;
  DGRP_CUDA(cudaGetLastError());
  return DGRP_OK;
}

// codes[i] = code(seq[start+i]); 16 bases per thread when the source is 16-byte aligned.
__global__ void codes_kernel(const uint8_t *__restrict__ seq, int64_t start, int64_t len,
                             uint8_t *__restrict__ codes) {
  const uint8_t *src = seq + start;
  const bool aligned = ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(codes)) & 15) == 0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t nvec = aligned ? len / 16 : 0;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
    uint4 in = __ldg(reinterpret_cast<const uint4 *>(src) + v);
    unsigned int w[4] = {in.x, in.y, in.z, in.w};
    unsigned int o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      unsigned int r = 0;
#pragma unroll
      for (int b = 0; b < 4; ++b) r |= (unsigned int)base_code((w[k] >> (8 * b)) & 0xffu) << (8 * b);
      o[k] = r;
    }
    reinterpret_cast<uint4 *>(codes)[v] = make_uint4(o[0], o[1], o[2], o[3]);
  }
  for (int64_t i = nvec * 16 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += stride)
    codes[i] = (uint8_t)base_code(src[i]);
}

int launch_codes(dgrp_ctx *c, const uint8_t *d_seq, int64_t start, int64_t len, uint8_t *d_codes) {
  if (len <= 0) return DGRP_OK;
  int threads = 256;
  int64_t want = (len / 16 + threads) / threads;
  int blocks = (int)(want < (int64_t)c->sm_count * 16 ? want : (int64_t)c->sm_count * 16);
  if (blocks < 1) blocks = 1;
  codes_kernel<<<blocks, threads, 0, c->stream>>>(d_seq, start, len, d_codes);
  c->launches++;
  DGRP_CUDA(cudaGetLastError());
  return DGRP_OK;
}

// int8 one-hot planes [5, len] -> codes; flag[0] is set when a column is not exactly one-hot.
__global__ void onehot_to_codes_kernel(const int8_t *__restrict__ fwd, int64_t len,
                                       uint8_t *__restrict__ codes, int *flag) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  bool bad = false;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += stride) {
    int code = 0, ones = 0;
#pragma unroll
    for (int ch = 0; ch < 5; ++ch) {
      int v = fwd[(int64_t)ch * len + i];
      if (v == 1) { code = ch; ++ones; }
      else if (v != 0) bad = true;
    }
    if (ones != 1) bad = true;
    codes[i] = (uint8_t)code;
  }
  if (bad) atomicOr(flag, 1);
}

int launch_onehot_to_codes(dgrp_ctx *c, const int8_t *d_fwd, int64_t len, uint8_t *d_codes) {
  // c->small[0..3] is used as the "not one-hot" flag (zeroed here)
  DGRP_CHECK(c->small.reserve(256));
  DGRP_CUDA(cudaMemsetAsync(c->small.p, 0, 4, c->stream));
  if (len <= 0) return DGRP_OK;
  int threads = 256;
  int64_t want = (len + threads - 1) / threads;
  int blocks = (int)(want < (int64_t)c->sm_count * 32 ? want : (int64_t)c->sm_count * 32);
  onehot_to_codes_kernel<<<blocks, threads, 0, c->stream>>>(d_fwd, len, d_codes, c->small.as<int>());
  c->launches++;
  DGRP_CUDA(cudaGetLastError());
  return DGRP_OK;
}

}  // namespace dgrp
