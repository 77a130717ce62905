/*
 * oracle/nms_oracle.c -- CPU restatement of vdetlib's native NMS arithmetic.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this.  The product path
 * (vdetlib_b200/) never links or calls it.
 *
 * Follows /root/reference/utils/nms.pyx line by line:
 *   oracle_nms            <- nms            nms.pyx:17-68
 *   oracle_vid_nms        <- vid_nms        nms.pyx:71-125
 *   oracle_track_det_nms  <- track_det_nms  nms.pyx:128-189
 *   oracle_pair_iou_f32   <- the pair arithmetic nms.pyx:57-64 (as compiled by
 *                            Cython 3.3: `((xx2 - xx1) + 1.0)` is a float
 *                            subtraction, a DOUBLE add of 1.0, then a narrowing to
 *                            float32 at the call of the inline max(); the threshold
 *                            test widens ovr to double, nms.pyx:65)
 *   oracle_link_f32       <- build-defined frame-to-frame link:
 *                            arg-max over j of the nms.pyx pair IoU, FIRST maximum
 *                            (np.argmax rule, tubelet_cls.py:375-376)
 *
 * Parity pin: checked against the real reference (oracle/_ref/cython_nms*.so built
 * from /root/reference/utils/nms.pyx by oracle/build_ref.py) in
 * tests/test_oracle_pin.py, and against the committed golden vectors in
 * tests/golden/ that were generated from that build.
 *
 * Ordering rule: the reference uses `scores.argsort()[::-1]` (nms.pyx:25,80), an
 * unstable sort whose tie order is numpy-build dependent.  With unique scores every
 * implementation agrees; for ties this oracle (and the CUDA path) define
 * "descending score, then ascending original index".
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared (see oracle/Makefile).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_ZERO_DIVISION (-1)

/* nms.pyx:11-15 */
static inline float f32_max(float a, float b) { return a >= b ? a : b; }
static inline float f32_min(float a, float b) { return a <= b ? a : b; }

/* numpy float32: (x2 - x1 + 1) * (y2 - y1 + 1)   nms.pyx:24,79,136,145 */
static inline float box_area(float x1, float y1, float x2, float y2) {
    float w = (x2 - x1) + 1.0f;
    float h = (y2 - y1) + 1.0f;
    return w * h;
}

/* nms.pyx:57-64.  Returns 0 and sets *zero_div when the divisor is 0. */
static inline float pair_iou(float ix1, float iy1, float ix2, float iy2, float iarea,
                             float jx1, float jy1, float jx2, float jy2, float jarea,
                             int *zero_div) {
    float xx1 = f32_max(ix1, jx1);
    float yy1 = f32_max(iy1, jy1);
    float xx2 = f32_min(ix2, jx2);
    float yy2 = f32_min(iy2, jy2);
    float w = f32_max(0.0f, (float)((double)(xx2 - xx1) + 1.0));
    float h = f32_max(0.0f, (float)((double)(yy2 - yy1) + 1.0));
    float inter = w * h;
    float uni = (iarea + jarea) - inter;
    if (uni == 0) { *zero_div = 1; return 0.0f; }
    return inter / uni;
}

typedef struct { float s; int64_t i; } sort_item;

static int cmp_desc(const void *a, const void *b) {
    const sort_item *x = (const sort_item *)a, *y = (const sort_item *)b;
    if (x->s > y->s) return -1;
    if (x->s < y->s) return 1;
    return (x->i > y->i) - (x->i < y->i);
}

/* order[k] = index of the k-th highest score; ties by ascending index. */
static int64_t *order_desc(const float *scores, int64_t stride, int64_t n) {
    sort_item *it = (sort_item *)malloc(sizeof(sort_item) * (size_t)(n > 0 ? n : 1));
    int64_t *order = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
    for (int64_t k = 0; k < n; ++k) { it[k].s = scores[k * stride]; it[k].i = k; }
    qsort(it, (size_t)n, sizeof(sort_item), cmp_desc);
    for (int64_t k = 0; k < n; ++k) order[k] = it[k].i;
    free(it);
    return order;
}

/*
 * Shared body of nms (ncol=5, no frame column) and vid_nms (ncol=6, frame first).
 * dets: row-major [n, ncol] float32.  keep_out: at least n entries.
 * Returns the number kept, or ORACLE_ZERO_DIVISION.
 */
static int64_t greedy(const float *dets, int64_t n, int ncol, double thresh, int64_t *keep_out) {
    const int has_frame = (ncol == 6);
    const int c0 = has_frame ? 1 : 0;
    float *areas = (float *)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
    for (int64_t k = 0; k < n; ++k) {
        const float *d = dets + k * ncol + c0;
        areas[k] = box_area(d[0], d[1], d[2], d[3]);
    }
    int64_t *order = order_desc(dets + c0 + 4, ncol, n);
    char *suppressed = (char *)calloc((size_t)(n > 0 ? n : 1), 1);
    int64_t nkeep = 0;
    int zero_div = 0;
    for (int64_t _i = 0; _i < n && !zero_div; ++_i) {
        int64_t i = order[_i];
        if (suppressed[i]) continue;
        keep_out[nkeep++] = i;
        const float *bi = dets + i * ncol + c0;
        float ix1 = bi[0], iy1 = bi[1], ix2 = bi[2], iy2 = bi[3], iarea = areas[i];
        for (int64_t _j = _i + 1; _j < n; ++_j) {
            int64_t j = order[_j];
            if (has_frame && dets[i * ncol] != dets[j * ncol]) continue;   /* nms.pyx:110-112 */
            if (suppressed[j]) continue;
            const float *bj = dets + j * ncol + c0;
            float ovr = pair_iou(ix1, iy1, ix2, iy2, iarea, bj[0], bj[1], bj[2], bj[3], areas[j], &zero_div);
            if (zero_div) break;
            if ((double)ovr >= thresh) suppressed[j] = 1;                   /* nms.pyx:65 */
        }
    }
    free(areas); free(order); free(suppressed);
    return zero_div ? ORACLE_ZERO_DIVISION : nkeep;
}

int64_t oracle_nms(const float *dets, int64_t n, double thresh, int64_t *keep_out) {
    return greedy(dets, n, 5, thresh, keep_out);
}

int64_t oracle_vid_nms(const float *dets, int64_t n, double thresh, int64_t *keep_out) {
    return greedy(dets, n, 6, thresh, keep_out);
}

/* nms.pyx:128-189.  tracks [q,5] = (frame,x1,y1,x2,y2); dets [k,6]. */
int64_t oracle_track_det_nms(const float *tracks, int64_t q, const float *dets, int64_t k,
                             double thresh, int64_t *keep_out) {
    float *t_areas = (float *)malloc(sizeof(float) * (size_t)(q > 0 ? q : 1));
    for (int64_t j = 0; j < q; ++j) {
        const float *t = tracks + j * 5 + 1;
        t_areas[j] = box_area(t[0], t[1], t[2], t[3]);
    }
    char *suppressed = (char *)calloc((size_t)(k > 0 ? k : 1), 1);
    int zero_div = 0;
    for (int64_t i = 0; i < k && !zero_div; ++i) {                          /* nms.pyx:163-183 */
        const float *d = dets + i * 6 + 1;
        float iarea = box_area(d[0], d[1], d[2], d[3]);
        for (int64_t j = 0; j < q; ++j) {
            if (dets[i * 6] != tracks[j * 5]) continue;
            const float *t = tracks + j * 5 + 1;
            float ovr = pair_iou(d[0], d[1], d[2], d[3], iarea, t[0], t[1], t[2], t[3], t_areas[j], &zero_div);
            if (zero_div) break;
            if ((double)ovr >= thresh) { suppressed[i] = 1; break; }
        }
    }
    int64_t out = ORACLE_ZERO_DIVISION;
    if (!zero_div) {
        int64_t nrem = 0;
        int64_t *remain = (int64_t *)malloc(sizeof(int64_t) * (size_t)(k > 0 ? k : 1));
        float *sub = (float *)malloc(sizeof(float) * 6 * (size_t)(k > 0 ? k : 1));
        for (int64_t i = 0; i < k; ++i)
            if (!suppressed[i]) { memcpy(sub + nrem * 6, dets + i * 6, 6 * sizeof(float)); remain[nrem++] = i; }
        int64_t *keep2 = (int64_t *)malloc(sizeof(int64_t) * (size_t)(nrem > 0 ? nrem : 1));
        int64_t n2 = greedy(sub, nrem, 6, thresh, keep2);                   /* nms.pyx:186-187 */
        if (n2 >= 0) { for (int64_t i = 0; i < n2; ++i) keep_out[i] = remain[keep2[i]]; out = n2; }
        free(remain); free(sub); free(keep2);
    }
    free(t_areas); free(suppressed);
    return out;
}

/* Dense float32 IoU matrix with the nms.pyx pair arithmetic.  union==0 -> NaN (0/0). */
void oracle_pair_iou_f32(const float *a, int64_t na, const float *b, int64_t nb, float *out) {
    for (int64_t i = 0; i < na; ++i) {
        const float *p = a + i * 4;
        float pa = box_area(p[0], p[1], p[2], p[3]);
        for (int64_t j = 0; j < nb; ++j) {
            const float *r = b + j * 4;
            float ra = box_area(r[0], r[1], r[2], r[3]);
            int zd = 0;
            float v = pair_iou(p[0], p[1], p[2], p[3], pa, r[0], r[1], r[2], r[3], ra, &zd);
            if (zd) { volatile float z = 0.0f; v = z / z; }
            out[i * nb + j] = v;
        }
    }
}

/*
 * Suppression bit matrix of one frame in ORIGINAL index space: bit (i,j) set iff
 * (double)IoU(i,j) >= thresh  (diagonal included).  mask: [n, words] uint32,
 * words = ceil(n/32).  Mirrors what the CUDA path stages in shared memory.
 */
void oracle_iou_bitmask(const float *boxes, int64_t n, double thresh, uint32_t *mask) {
    int64_t words = (n + 31) / 32;
    memset(mask, 0, sizeof(uint32_t) * (size_t)(n * words));
    for (int64_t i = 0; i < n; ++i) {
        const float *p = boxes + i * 4;
        float pa = box_area(p[0], p[1], p[2], p[3]);
        for (int64_t j = 0; j < n; ++j) {
            const float *r = boxes + j * 4;
            float ra = box_area(r[0], r[1], r[2], r[3]);
            int zd = 0;
            float v = pair_iou(p[0], p[1], p[2], p[3], pa, r[0], r[1], r[2], r[3], ra, &zd);
            if (!zd && (double)v >= thresh) mask[i * words + (j >> 5)] |= (1u << (j & 31));
        }
    }
}

/*
 * Frame-to-frame link (build-defined).  For frame t in [0,T-1) and
 * box i < counts[t]: succ = FIRST arg-max over j < counts[t+1] of the float32 pair IoU,
 * best = that IoU.  counts[t+1]==0 -> succ=-1, best=0.  boxes: [T, nmax, 4];
 * succ/best: [T-1, nmax] (entries i >= counts[t] are written as -1 / 0).
 * NaN (union==0) never wins (comparison `v > best` is false), matching the CUDA path.
 */
void oracle_link_f32(const float *boxes, const int32_t *counts, int64_t T, int64_t nmax,
                     int32_t *succ, float *best_out) {
    for (int64_t t = 0; t + 1 < T; ++t) {
        const float *A = boxes + t * nmax * 4, *B = boxes + (t + 1) * nmax * 4;
        for (int64_t i = 0; i < nmax; ++i) {
            int32_t arg = -1; float best = 0.0f;
            if (i < counts[t]) {
                const float *p = A + i * 4;
                float pa = box_area(p[0], p[1], p[2], p[3]);
                for (int64_t j = 0; j < counts[t + 1]; ++j) {
                    const float *r = B + j * 4;
                    float ra = box_area(r[0], r[1], r[2], r[3]);
                    int zd = 0;
                    float v = pair_iou(p[0], p[1], p[2], p[3], pa, r[0], r[1], r[2], r[3], ra, &zd);
                    if (zd) continue;
                    if (arg < 0 || v > best) { arg = (int32_t)j; best = v; }
                }
            }
            succ[t * nmax + i] = arg;
            best_out[t * nmax + i] = best;
        }
    }
}

/*
 * Batched class-shared NMS over frames (what apply_vid_nms does once per class,
 * video_det.py:51-61, restated per (frame, class) problem):
 * boxes [T, nmax, 4], scores [T, nmax, C], counts [T].
 * keep_mask [T, C, nmax] u8; keep_idx [T, C, nmax] int32 (descending score, -1 padded);
 * keep_cnt [T, C].  Returns 0 or ORACLE_ZERO_DIVISION.
 */
int64_t oracle_nms_frames(const float *boxes, const float *scores, const int32_t *counts,
                          int64_t T, int64_t nmax, int64_t C, double thresh,
                          uint8_t *keep_mask, int32_t *keep_idx, int32_t *keep_cnt) {
    float *dets = (float *)malloc(sizeof(float) * 5 * (size_t)(nmax > 0 ? nmax : 1));
    int64_t *keep = (int64_t *)malloc(sizeof(int64_t) * (size_t)(nmax > 0 ? nmax : 1));
    int64_t rc = 0;
    for (int64_t t = 0; t < T && rc == 0; ++t) {
        int64_t n = counts[t];
        for (int64_t c = 0; c < C; ++c) {
            for (int64_t k = 0; k < n; ++k) {
                memcpy(dets + k * 5, boxes + (t * nmax + k) * 4, 4 * sizeof(float));
                dets[k * 5 + 4] = scores[(t * nmax + k) * C + c];
            }
            int64_t nk = greedy(dets, n, 5, thresh, keep);
            if (nk < 0) { rc = ORACLE_ZERO_DIVISION; break; }
            uint8_t *km = keep_mask + (t * C + c) * nmax;
            int32_t *ki = keep_idx + (t * C + c) * nmax;
            memset(km, 0, (size_t)nmax);
            for (int64_t k = 0; k < nmax; ++k) ki[k] = -1;
            for (int64_t k = 0; k < nk; ++k) { km[keep[k]] = 1; ki[k] = (int32_t)keep[k]; }
            keep_cnt[t * C + c] = (int32_t)nk;
        }
    }
    free(dets); free(keep);
    return rc;
}
