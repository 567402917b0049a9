/* oracle/oracle_c.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C restatement of the integer / byte / double-precision parts of DeepGRP's prediction
 * path, used only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg as the
 * checker for the CUDA kernels.  Nothing under deepgrp_b200/ may link or call this file.
 *
 * Every function cites the reference lines it restates (paths relative to /root/reference).
 * The restatement is pinned against the reference's own compiled natives (oracle/_ref, built
 * by oracle/build_ref.py) and against the reference's known-answer tests in
 * tests/test_oracle.py.
 *
 * Build: gcc -O2 -shared -fPIC -o oracle/liboracle.so oracle/oracle_c.c -lm   (see Makefile)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------
 * Base -> channel map.  Restates the 128-entry ONEHOT table of deepgrp/sequence.pyx:11-17:
 * A/a->0, C/c->1, G/g->2, T/t->3, every other byte below 128 -> 4.  (Bytes >= 128 index past
 * the reference's table = undefined behaviour there; the oracle maps them to 4.)
 * ---------------------------------------------------------------------------------------- */
static int base_channel(unsigned char b)
{
    switch (b) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return 4;
    }
}

/* deepgrp/sequence.pyx:21-36 (_one_hot_encode_dna_sequence): skip leading and trailing
 * UPPERCASE 'N' only, then fwd[channel(byte), i] = 1 into a zeroed int8[5, len] C-order array.
 * Call with fwd == NULL to obtain (startpos, outlen) first.  Returns outlen, which is NEGATIVE
 * for an all-'N' input exactly as `length - startpos` is in the reference (np.zeros then raises
 * "negative dimensions"). */
long orc_one_hot_encode(const unsigned char *seq, long n, long *startpos, int8_t *fwd)
{
    long st = 0, len = n;
    while (st < len && seq[st] == 'N') ++st;
    while (len > 0 && seq[len - 1] == 'N') --len;
    *startpos = st;
    long out = len - st;
    if (fwd && out > 0) {
        memset(fwd, 0, (size_t)5 * (size_t)out);
        for (long i = 0; i < out; ++i)
            fwd[(size_t)base_channel(seq[st + i]) * (size_t)out + (size_t)i] = 1;
    }
    return out;
}

/* deepgrp/maxcalc.c:10-24 (_get_max): window b of `inputs` (dim0 x dim1 floats) is merged by
 * elementwise maximum into `output` starting at row b*stride. No bounds checks, as there. */
void orc_get_max(float *output, const float *inputs, size_t dim0, size_t dim1, size_t stride,
                 size_t batchsize)
{
    const size_t per_window = dim0 * dim1;
    for (size_t b = 0; b < batchsize; ++b) {
        float *dst = output + b * stride * dim1;
        const float *src = inputs + b * per_window;
        for (size_t e = 0; e < per_window; ++e)
            if (src[e] > dst[e]) dst[e] = src[e];
    }
}

/* deepgrp/sequence.pyx:40-53 (get_segments): note both loops are bounded by size-1. */
void orc_get_segments(const long *classes, long size, long startpos, long out3[3])
{
    const long last = size - 1;
    long cur = classes[startpos];
    while (startpos < last && cur == 0) {
        ++startpos;
        cur = classes[startpos];
    }
    long end = startpos + 1;
    while (end < last && classes[end] == cur) ++end;
    out3[0] = startpos; out3[1] = end; out3[2] = cur;
}

/* deepgrp/sequence.pyx:79-85 (yield_segments), materialised: writes (start+off, end+off, label)
 * triples for EVERY segment (including label 0, as the generator yields them); returns count.
 * `out` needs room for 3*size longs. */
long orc_yield_segments(const long *classes, long size, long start_offset, long *out)
{
    long i = 0, n = 0, seg[3];
    while (i < size) {
        orc_get_segments(classes, size, i, seg);
        i = seg[1];
        out[3 * n] = seg[0] + start_offset;
        out[3 * n + 1] = seg[1] + start_offset;
        out[3 * n + 2] = seg[2];
        ++n;
    }
    return n;
}

/* ------------------------------------------------------------------------------------------
 * Ruzzo-Tompa all maximal scoring segments with x-drop.  Restates deepgrp/_mss/mss.c:50-101
 * (mss_find_all) and :35-47 (move_segs) as an explicit state machine:
 *   state = { cum (running sum "L"), peak ("max"), candidate stack }.
 * Quirks kept on purpose:
 *   - move_segs takes `int min_sc` (mss.c:35): the double threshold is TRUNCATED toward zero
 *     when passed, and then compared as `R - L >= (double)(int)min_sc`;
 *   - indices are 32-bit int; positive means strictly `> 0` (mss.c:59,62);
 *   - x-drop test `xdrop > 0.0 && L + S[i] + xdrop < max` (mss.c:89), NEG_INF = -1e30.
 * ---------------------------------------------------------------------------------------- */
typedef struct { int st, en; double sc; } orc_seg_t;       /* same layout as mss.h:11-14 */

typedef struct { int st, en; double lo, hi; int link; } cand_t;

typedef struct {
    cand_t *cand; size_t ncand, capcand;
    orc_seg_t *out; size_t nout, capout;
} rt_state_t;

static void rt_flush(rt_state_t *s, int min_sc_int)
{
    for (size_t i = 0; i < s->ncand; ++i) {
        const cand_t *c = &s->cand[i];
        if (c->hi - c->lo >= min_sc_int) {
            if (s->nout == s->capout) {
                s->capout = s->capout ? s->capout * 2 : 16;
                s->out = (orc_seg_t *)realloc(s->out, s->capout * sizeof(orc_seg_t));
            }
            s->out[s->nout].st = c->st;
            s->out[s->nout].en = c->en;
            s->out[s->nout].sc = c->hi - c->lo;
            ++s->nout;
        }
    }
    s->ncand = 0;
}

orc_seg_t *orc_mss_find_all(int n, const double *S, double min_sc, double xdrop, int *n_seg)
{
    const int min_sc_int = (int)min_sc;                 /* the C call-site truncation */
    rt_state_t s; memset(&s, 0, sizeof s);
    double cum = 0.0, peak = -1e30;
    int i = 0;
    while (i < n) {
        if (S[i] > 0) {
            /* one maximal run of strictly positive scores [i, k) */
            int k = i + 1;
            double hi = cum + S[i];
            while (k < n && S[k] > 0.) { hi += S[k]; ++k; }
            if (hi > peak) peak = hi;
            cand_t t; t.st = i; t.en = k; t.lo = cum; t.hi = hi; t.link = -1;
            for (;;) {
                /* rightmost candidate whose lo is strictly below t.lo, following links */
                long j = (long)s.ncand - 1;
                while (j >= 0) {
                    const cand_t *p = &s.cand[j];
                    if (p->lo < t.lo) break;
                    j = p->link >= 0 ? p->link : j - 1;
                }
                if (j >= 0 && s.cand[j].hi < t.hi) {     /* absorb candidates j.. into t */
                    t.st = s.cand[j].st; t.lo = s.cand[j].lo; t.link = s.cand[j].link;
                    s.ncand = (size_t)j;
                    continue;
                }
                if (j < 0) { rt_flush(&s, min_sc_int); peak = hi; }
                t.link = (int)j;
                if (s.ncand == s.capcand) {
                    s.capcand = s.capcand ? s.capcand * 2 : 16;
                    s.cand = (cand_t *)realloc(s.cand, s.capcand * sizeof(cand_t));
                }
                s.cand[s.ncand++] = t;
                break;
            }
            cum = hi; i = k;
        } else {
            if (xdrop > 0.0 && cum + S[i] + xdrop < peak) {
                rt_flush(&s, min_sc_int);
                cum = 0.0; peak = -1e30;
            }
            cum += S[i]; ++i;
        }
    }
    rt_flush(&s, min_sc_int);
    free(s.cand);
    *n_seg = (int)s.nout;
    return s.out;                                        /* caller frees with orc_free */
}

void orc_free(void *p) { free(p); }

/* deepgrp/_mss/pymss.pyx:31-80 (_find_mss_labels): s0 = ln(0.99/0.01); xdrop = s0*xdrop_len*10
 * if xdrop_len > 0 else -1; min_sc = s0*min_mss_len; per kept segment the majority label over
 * classes 1..nof-1 (ties -> lowest, default 1) replaces label-0 positions; every other position
 * keeps its own label.  Output: zero-initialised one-hot double[n, nof_labels]. */
void orc_find_mss_labels(const double *inputs, const long *label, int nof_labels, int min_mss_len,
                         int xdrop_len, double *one_hot, int n)
{
    const double s0 = log(0.99 / (1.0 - 0.99));
    const double xdrop = xdrop_len > 0 ? s0 * xdrop_len * 10.0 : -1.0;
    const double min_sc = s0 * min_mss_len;
    int nseg = 0;
    orc_seg_t *segs = orc_mss_find_all(n, inputs, min_sc, xdrop, &nseg);
    long *counts = (long *)malloc((size_t)nof_labels * sizeof(long));
    memset(one_hot, 0, (size_t)n * (size_t)nof_labels * sizeof(double));
    long pos = 0;
    for (int g = 0; g < nseg; ++g) {
        for (int c = 0; c < nof_labels; ++c) counts[c] = 0;
        for (long j = segs[g].st; j < segs[g].en; ++j) counts[label[j]] += 1;
        int best = 1; long bestv = counts[1];
        for (int c = 2; c < nof_labels; ++c)
            if (bestv < counts[c]) { best = c; bestv = counts[c]; }
        for (long j = segs[g].st; j < segs[g].en; ++j)
            one_hot[(size_t)j * nof_labels + (label[j] == 0 ? best : label[j])] = 1.0;
        for (long j = pos; j < segs[g].st; ++j) one_hot[(size_t)j * nof_labels + label[j]] = 1.0;
        pos = segs[g].en;
    }
    for (long j = pos; j < n; ++j) one_hot[(size_t)j * nof_labels + label[j]] = 1.0;
    free(counts);
    free(segs);
}

/* Same as above but returns the relabelled classes directly (argmax of the one-hot), for
 * sizes where a double[n,5] matrix is too large.  Not a reference function: a convenience for
 * parity tests at BASELINE sizes. */
void orc_find_mss_relabel(const double *inputs, const long *label, int nof_labels,
                          int min_mss_len, int xdrop_len, uint8_t *out, int n)
{
    const double s0 = log(0.99 / (1.0 - 0.99));
    const double xdrop = xdrop_len > 0 ? s0 * xdrop_len * 10.0 : -1.0;
    const double min_sc = s0 * min_mss_len;
    int nseg = 0;
    orc_seg_t *segs = orc_mss_find_all(n, inputs, min_sc, xdrop, &nseg);
    long *counts = (long *)malloc((size_t)nof_labels * sizeof(long));
    for (long j = 0; j < n; ++j) out[j] = (uint8_t)label[j];
    for (int g = 0; g < nseg; ++g) {
        for (int c = 0; c < nof_labels; ++c) counts[c] = 0;
        for (long j = segs[g].st; j < segs[g].en; ++j) counts[label[j]] += 1;
        int best = 1; long bestv = counts[1];
        for (int c = 2; c < nof_labels; ++c)
            if (bestv < counts[c]) { best = c; bestv = counts[c]; }
        for (long j = segs[g].st; j < segs[g].en; ++j)
            if (label[j] == 0) out[j] = (uint8_t)best;
    }
    free(counts);
    free(segs);
}
