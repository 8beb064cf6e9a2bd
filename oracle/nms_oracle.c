/* CPU oracle (C twin) for greedy per-class NMS.  TEST INFRASTRUCTURE ONLY:
 * linked/loaded only by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg; the product library never links it.
 *
 * Restates /root/reference utils/postprocess.py:
 *   y2o_pair_iou   <- iou()               utils/postprocess.py:21-36
 *   y2o_nms_image  <- non_max_suppress()  utils/postprocess.py:39-51
 * Semantics pinned (see oracle/nms_oracle.py and tests/golden/nms_*.npz):
 * float32 arithmetic in the reference's operation order (compile with
 * -ffp-contract=off so no FMA is formed), thresholds rounded to float32
 * (NumPy >= 2 / NEP-50 behaviour of the reference file as executed in the
 * authoring container), Python's stable descending sort carried across classes.
 *
 * Build: see oracle/Makefile  (gcc -O2 -ffp-contract=off -fPIC -shared).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

static inline float f_max(float a, float b) { return a > b ? a : b; }
static inline float f_min(float a, float b) { return a < b ? a : b; }

float y2o_pair_iou(const float *min1, const float *max1, const float *min2, const float *max2)
{
    float a1 = (max1[0] - min1[0]) * (max1[1] - min1[1]);
    float a2 = (max2[0] - min2[0]) * (max2[1] - min2[1]);
    float iw = f_max(f_min(max1[0], max2[0]) - f_max(min1[0], min2[0]), 0.0f);
    float ih = f_max(f_min(max1[1], max2[1]) - f_max(min1[1], min2[1]), 0.0f);
    float inter = iw * ih;
    float d0 = (a1 + a2) - inter;
    /* np.maximum propagates NaN (utils/postprocess.py:36): inf + inf - inf (boxes whose areas overflow float32) gives iou = nan,
     * and nan >= threshold_iou is False -- no suppression.  f_max alone would turn the NaN into 1e-10 and the pair into a hit. */
    float den = (d0 != d0) ? d0 : f_max(d0, (float)1e-10);
    return inter / den;
}

/* stable merge sort of idx[0..n) by key[idx] descending (ties keep order) */
static void merge_sort_desc(int32_t *idx, int32_t *tmp, const float *key, int stride, int n)
{
    for (int width = 1; width < n; width *= 2) {
        for (int lo = 0; lo < n; lo += 2 * width) {
            int mid = lo + width < n ? lo + width : n;
            int hi = lo + 2 * width < n ? lo + 2 * width : n;
            int i = lo, j = mid, k = lo;
            while (i < mid && j < hi) {
                /* take right only if strictly greater: stability */
                if (key[(size_t)idx[j] * stride] > key[(size_t)idx[i] * stride]) tmp[k++] = idx[j++];
                else tmp[k++] = idx[i++];
            }
            while (i < mid) tmp[k++] = idx[i++];
            while (j < hi) tmp[k++] = idx[j++];
        }
        memcpy(idx, tmp, (size_t)n * sizeof(int32_t));
    }
}

/* One image. conf [n, classes] (mutated), xy_min/xy_max [n,2]. order_out [n] (nullable).
 * Returns 0, or -1 if a reference assert (NaN / inverted box) would have fired. */
int y2o_nms_image(float *conf, const float *xy_min, const float *xy_max, int n, int classes,
                  float threshold, float threshold_iou, int32_t *order_out)
{
    int32_t *order = (int32_t *)malloc((size_t)n * sizeof(int32_t));
    int32_t *tmp = (int32_t *)malloc((size_t)n * sizeof(int32_t));
    int rc = 0;
    for (int i = 0; i < n; ++i) order[i] = i;
    for (int c = 0; c < classes; ++c) {
        merge_sort_desc(order, tmp, conf + c, classes, n);
        for (int i = 0; i + 1 < n; ++i) {
            int b = order[i];
            if (conf[(size_t)b * classes + c] <= threshold) continue;
            for (int j = i + 1; j < n; ++j) {
                int o = order[j];
                const float *m1 = xy_min + 2 * (size_t)b, *M1 = xy_max + 2 * (size_t)b;
                const float *m2 = xy_min + 2 * (size_t)o, *M2 = xy_max + 2 * (size_t)o;
                if (isnan(m1[0]) || isnan(m1[1]) || isnan(M1[0]) || isnan(M1[1]) ||
                    isnan(m2[0]) || isnan(m2[1]) || isnan(M2[0]) || isnan(M2[1]) ||
                    !(m1[0] <= M1[0]) || !(m1[1] <= M1[1]) || !(m2[0] <= M2[0]) || !(m2[1] <= M2[1])) {
                    rc = -1;
                    goto done;
                }
                if (y2o_pair_iou(m1, M1, m2, M2) >= threshold_iou) conf[(size_t)o * classes + c] = 0.0f;
            }
        }
    }
    if (order_out) memcpy(order_out, order, (size_t)n * sizeof(int32_t));
done:
    free(order);
    free(tmp);
    return rc;
}

int y2o_nms_batch(float *conf, const float *xy_min, const float *xy_max, int batch, int n, int classes,
                  float threshold, float threshold_iou, int32_t *order_out)
{
    for (int b = 0; b < batch; ++b) {
        int rc = y2o_nms_image(conf + (size_t)b * n * classes, xy_min + (size_t)b * n * 2,
                               xy_max + (size_t)b * n * 2, n, classes, threshold, threshold_iou,
                               order_out ? order_out + (size_t)b * n : 0);
        if (rc) return rc;
    }
    return 0;
}
