// K6: max-vote (standalone `_get_max` semantics), score transform, softmax.
// Replaces deepgrp/maxcalc.c:10-24, deepgrp/prediction.py:51-57 and :62-65.
#include "dgrp_internal.cuh"

namespace dgrp {

// Gather form of _get_max: every output element applies `cur > old ? cur : old` over its covering
// windows in ascending window order -- the same operation sequence the reference's nested loops
// apply to that element, so results are bit-identical (NaN behaviour included), with no races.
__global__ void get_max_kernel(float *__restrict__ out, const float *__restrict__ in,
                               int64_t batch, int64_t dim0, int64_t dim1, int64_t stride,
                               int64_t rows) {
  const int64_t total = rows * dim1;
  const int64_t gs = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gs) {
    const int64_t r = e / dim1, col = e - r * dim1;
    int64_t b_lo, b_hi;
    if (stride == 0) {
      b_lo = 0;
      b_hi = batch - 1;
    } else {
      b_lo = r >= dim0 ? (r - dim0) / stride + 1 : 0;
      b_hi = r / stride;
      if (b_hi > batch - 1) b_hi = batch - 1;
    }
    float acc = out[e];
    for (int64_t b = b_lo; b <= b_hi; ++b) {
      const float v = in[(b * dim0 + (r - b * stride)) * dim1 + col];
      acc = acc > v ? acc : v;
    }
    out[e] = acc;
  }
}

int launch_get_max(dgrp_ctx *c, float *d_out, const float *d_in, int64_t batch, int64_t dim0,
                   int64_t dim1, int64_t stride) {
  if (batch <= 0 || dim0 <= 0 || dim1 <= 0) return DGRP_OK;
  const int64_t rows = (batch - 1) * stride + dim0;  // rows touched by the reference loops
  const int64_t total = rows * dim1;
  int threads = 256;
  int64_t want = (total + threads - 1) / threads;
  int blocks = (int)(want < (int64_t)c->sm_count * 32 ? want : (int64_t)c->sm_count * 32);
  get_max_kernel<<<blocks, threads, 0, c->stream>>>(d_out, d_in, batch, dim0, dim1, stride, rows);
  c->launches++;
  DGRP_CUDA(cudaGetLastError());
  return DGRP_OK;
}

// prediction.py:51-57 (float32 arithmetic, then widened):
//   cls = argmax(p) (first maximum); m = max(p) + 1e-6f; m > 0.99f -> 0.99f;
//   t = logf(m / (1 - m)); score = cls > 0 ? t : -10 * t.
// An all-zero (never covered) row gives cls 0, t = log(1e-6/(1-1e-6)) -> score +138.155.
template <int C>
__device__ __forceinline__ void score_row(const float *__restrict__ p, int &cls, float &score) {
  float best = p[0];
  cls = 0;
#pragma unroll
  for (int k = 1; k < C; ++k) {
    float v = p[k];
    if (v > best) { best = v; cls = k; }
  }
  float m = best + 1e-6f;
  if (m > 0.99f) m = 0.99f;
  float t = logf(m / (1.0f - m));
  score = cls > 0 ? t : -10.0f * t;
}

// Max-merge of window probabilities into the prediction rows (the vote of prediction.py:103-111 /
// maxcalc.c:10-24 in gather form): row r takes the maximum over the windows of [w_begin, w_end) whose
// PLACED rows cover it -- windows of complete batches sit at w * step, those of the short last batch
// at tail_base + (w - full_windows) * step (prediction.py:105).  `win` is [w - w_begin][T][C].  One
// thread per row; consecutive rows read consecutive 4C-byte records of the same window.  The row's
// current value takes part in the maximum (zero-initialised, or the other window family's result).
// FUSED form (label != null; all windows of the record are in `win`): the row's vote is final here, so the score
// transform of prediction.py:51-57 is applied at once and `pred` is neither read nor written -- the float32[L, C]
// predictions never touch HBM (K6 of SURVEY.md section 2.3: "emits label u8 + score").
__global__ void vote_gather_kernel(const float *__restrict__ win, int64_t w_begin, int64_t w_end, int T, int C,
                                   int64_t full_windows, int64_t tail_base, int step,
                                   float *__restrict__ pred, int64_t pred_row0, int64_t pred_rows,
                                   uint8_t *__restrict__ label, float *__restrict__ score) {
  const int64_t gs = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < pred_rows; r += gs) {
    const int64_t g = r + pred_row0;   // record position
    float acc[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) acc[k] = (k < C && !label) ? pred[r * C + k] : 0.f;
    bool any = false;
#pragma unroll
    for (int fam = 0; fam < 2; ++fam) {
      // family 0: windows [w_begin, min(w_end, full)) placed at w * step
      // family 1: windows [max(w_begin, full), w_end) placed at tail_base + (w - full) * step
      const int64_t w_lo = fam == 0 ? w_begin : (w_begin > full_windows ? w_begin : full_windows);
      const int64_t w_hi = fam == 0 ? (w_end < full_windows ? w_end : full_windows) : w_end;
      if (w_lo >= w_hi) continue;
      const int64_t origin = fam == 0 ? 0 : tail_base - full_windows * (int64_t)step;   // place(w) = origin + w*step
      const int64_t x = g - origin;                                                      // w*step <= x < w*step + T
      if (x < 0) continue;
      int64_t a = x >= T ? (x - T) / step + 1 : 0, b = x / step;
      if (a < w_lo) a = w_lo;
      if (b > w_hi - 1) b = w_hi - 1;
      for (int64_t w = a; w <= b; ++w) {
        const float *src = win + ((size_t)(w - w_begin) * T + (size_t)(x - w * step)) * C;
#pragma unroll
        for (int k = 0; k < 5; ++k)
          if (k < C) acc[k] = fmaxf(acc[k], src[k]);
        any = true;
      }
    }
    if (label) {
      int cls = 0;
      float best = acc[0];
#pragma unroll
      for (int k = 1; k < 5; ++k)
        if (k < C && acc[k] > best) { best = acc[k]; cls = k; }
      float m = best + 1e-6f;
      if (m > 0.99f) m = 0.99f;
      const float t = logf(m / (1.0f - m));
      label[r] = (uint8_t)cls;
      score[r] = cls > 0 ? t : -10.0f * t;
    } else if (any) {
#pragma unroll
      for (int k = 0; k < 5; ++k)
        if (k < C) pred[r * C + k] = acc[k];
    }
  }
}

int launch_vote_gather(dgrp_ctx *c, const float *d_win, int64_t w_begin, int64_t w_end, int T, int C,
                       int64_t full_windows, int64_t tail_base, int step, float *d_pred, int64_t pred_row0,
                       int64_t pred_rows, uint8_t *d_label, float *d_score) {
  if (pred_rows <= 0 || (w_end <= w_begin && !d_label)) return DGRP_OK;
  const int threads = 256;
  const int64_t want = (pred_rows + threads - 1) / threads;
  const int64_t cap = (int64_t)c->sm_count * 16;
  vote_gather_kernel<<<(unsigned)(want < cap ? want : cap), threads, 0, c->stream>>>(
      d_win, w_begin, w_end, T, C, full_windows, tail_base, step, d_pred, pred_row0, pred_rows, d_label, d_score);
  c->launches++;
  DGRP_CUDA(cudaGetLastError());
  return DGRP_OK;
}

__global__ void score_kernel(const float *__restrict__ pred, int64_t n, int C,
                             uint8_t *__restrict__ label, float *__restrict__ s32,
                             double *__restrict__ s64, int64_t *__restrict__ c64) {
  const int64_t gs = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gs) {
    int cls;
    float sc;
    if (C == 5) {
      float row[5];
#pragma unroll
      for (int k = 0; k < 5; ++k) row[k] = pred[i * 5 + k];
      score_row<5>(row, cls, sc);
    } else {
      const float *p = pred + i * C;
      float best = p[0];
      cls = 0;
      for (int k = 1; k < C; ++k)
        if (p[k] > best) { best = p[k]; cls = k; }
      float m = best + 1e-6f;
      if (m > 0.99f) m = 0.99f;
      float t = logf(m / (1.0f - m));
      sc = cls > 0 ? t : -10.0f * t;
    }
    if (label) label[i] = (uint8_t)cls;
    if (s32) s32[i] = sc;
    if (s64) s64[i] = (double)sc;
    if (c64) c64[i] = cls;
  }
}

int launch_score(dgrp_ctx *c, const float *d_pred, int64_t n, int C, uint8_t *d_label,
                 float *d_score32, double *d_score64, int64_t *d_class64) {
  if (n <= 0) return DGRP_OK;
  int threads = 256;
  int64_t want = (n + threads - 1) / threads;
  int blocks = (int)(want < (int64_t)c->sm_count * 16 ? want : (int64_t)c->sm_count * 16);
  score_kernel<<<blocks, threads, 0, c->stream>>>(d_pred, n, C, d_label, d_score32, d_score64,
                                                  d_class64);
  c->launches++;
  DGRP_CUDA(cudaGetLastError());
  return DGRP_OK;
}

// prediction.py:62-65: e = exp(x - max(x over the WHOLE array)); out = e / rowsum(e).
__device__ __forceinline__ unsigned int f2ord(float f) {
  unsigned int u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned int u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

__global__ void global_max_kernel(const float *__restrict__ in, int64_t total, unsigned int *gmax) {
  unsigned int best = 0u;
  const int64_t gs = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gs) {
    unsigned int o = f2ord(in[i]);
    best = o > best ? o : best;
  }
  for (int off = 16; off > 0; off >>= 1) {
    unsigned int o = __shfl_xor_sync(0xffffffffu, best, off);
    best = o > best ? o : best;
  }
  if ((threadIdx.x & 31) == 0) atomicMax(gmax, best);
}

__global__ void softmax_rows_kernel(const float *__restrict__ in, int64_t n, int C,
                                    const unsigned int *gmax, float *__restrict__ out) {
  const float mx = ord2f(*gmax);
  const int64_t gs = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gs) {
    float sum = 0.f;
    for (int k = 0; k < C; ++k) sum += expf(in[i * C + k] - mx);
    for (int k = 0; k < C; ++k) out[i * C + k] = expf(in[i * C + k] - mx) / sum;
  }
}

int launch_softmax_global(dgrp_ctx *c, const float *d_in, int64_t n, int C, float *d_out) {
  if (n <= 0) return DGRP_OK;
  DGRP_CHECK(c->small.reserve(256));
  DGRP_CUDA(cudaMemsetAsync(c->small.p, 0, 4, c->stream));
  int threads = 256;
  int64_t total = n * C;
  int64_t want = (total + threads - 1) / threads;
  int blocks = (int)(want < (int64_t)c->sm_count * 16 ? want : (int64_t)c->sm_count * 16);
  global_max_kernel<<<blocks, threads, 0, c->stream>>>(d_in, total, c->small.as<unsigned int>());
  want = (n + threads - 1) / threads;
  blocks = (int)(want < (int64_t)c->sm_count * 16 ? want : (int64_t)c->sm_count * 16);
  softmax_rows_kernel<<<blocks, threads, 0, c->stream>>>(d_in, n, C, c->small.as<unsigned int>(), d_out);
  c->launches += 2;
  DGRP_CUDA(cudaGetLastError());
  return DGRP_OK;
}

}  // namespace dgrp
