/* CPU oracle for SLIC superpixels -- TEST INFRASTRUCTURE ONLY.
 *
 * Restates the algorithm of scikit-image's `skimage.segmentation.slic` as the
 * reference calls it (/root/reference/models/wesup.py:471-476:
 * slic(img_HWC, n_segments=int(H*W/sp_area), compactness=sp_compactness), all
 * other arguments at their defaults: max_iter=10, sigma=0, convert2lab,
 * enforce_connectivity=True, min_size_factor=0.5, max_size_factor=3).
 *
 * scikit-image is a third-party, UNPINNED dependency of the reference
 * (requirements.txt:10; code era 2019 => 0.15/0.16) and is neither vendored
 * under /root/reference nor installable in this image.  This file restates the
 * published algorithm (slic_superpixels.py + _slic.pyx + colorconv.rgb2lab +
 * _regular_grid.py of that era) from SURVEY.md Appendix B.  PARITY UNPINNED:
 * no reference test or golden vector pins SLIC output.
 *
 * All arithmetic is IEEE double, evaluated in the order written (compile with
 * -ffp-contract=off) so the CUDA kernels can reproduce it bit for bit.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

/* ---- rgb -> CIE Lab (D65, 2 degree observer), skimage.color.rgb2lab ---- */
static void rgb_to_lab(const float *rgb_hwc, double *lab_hwc, long n_px)
{
    static const double M[3][3] = {
        {0.412453, 0.357580, 0.180423},
        {0.212671, 0.715160, 0.072169},
        {0.019334, 0.119193, 0.950227}};
    static const double white[3] = {0.95047, 1.0, 1.08883};
    for (long p = 0; p < n_px; ++p) {
        double lin[3], xyz[3], f[3];
        for (int c = 0; c < 3; ++c) {
            double v = (double)rgb_hwc[3 * p + c];
            lin[c] = (v > 0.04045) ? pow((v + 0.055) / 1.055, 2.4) : v / 12.92;
        }
        for (int r = 0; r < 3; ++r) {
            double acc = lin[0] * M[r][0];
            acc = acc + lin[1] * M[r][1];
            acc = acc + lin[2] * M[r][2];
            xyz[r] = acc / white[r];
        }
        for (int r = 0; r < 3; ++r)
            f[r] = (xyz[r] > 0.008856) ? cbrt(xyz[r]) : 7.787 * xyz[r] + 16.0 / 116.0;
        lab_hwc[3 * p + 0] = 116.0 * f[1] - 16.0;
        lab_hwc[3 * p + 1] = 500.0 * (f[0] - f[1]);
        lab_hwc[3 * p + 2] = 200.0 * (f[1] - f[2]);
    }
}

/* skimage.util.regular_grid for a (1,H,W) volume: step and start along y/x.
 * Returns the number of seeds. */
long slic_ref_grid(int H, int W, int n_segments, int *step_out, int *start_out)
{
    double space = (double)H * (double)W;
    if (space <= (double)n_segments) { *step_out = 1; *start_out = 0; return (long)H * W; }
    /* ndim=3 with a unit depth: the cube-root step exceeds the depth, so the
     * depth step collapses to 1 and the in-plane step is the square root. */
    double s = sqrt(space / (double)n_segments);
    int smaller = H < W ? H : W;
    if ((double)smaller < s) {
        /* degenerate strip image (short side below one step): skimage collapses
         * a second axis; the reference never produces such inputs. */
        *step_out = 0; *start_out = 0; return -1;
    }
    int start = (int)floor(s / 2.0);
    int step = (int)nearbyint(s);          /* np.round: half to even */
    if (step < 1) step = 1;
    *step_out = step; *start_out = start;
    long ny = (H - start + step - 1) / step;
    long nx = (W - start + step - 1) / step;
    if (ny < 0) ny = 0;
    if (nx < 0) nx = 0;
    return ny * nx;
}

/* ---- k-means iterations, _slic_cython ---- */
static void slic_iterate(const double *img /* (H,W,3) scaled Lab */, int H, int W,
                         double *cent /* (K,5): y,x,L,a,b */, long K, int step,
                         int max_iter, int32_t *nearest)
{
    long n_px = (long)H * W;
    double *dist = (double *)malloc(sizeof(double) * n_px);
    long *count = (long *)malloc(sizeof(long) * K);
    float stepf = (float)step;
    double spatial_weight = 1.0 / (double)(stepf * stepf);
    for (int it = 0; it < max_iter; ++it) {
        int change = 0;
        for (long p = 0; p < n_px; ++p) dist[p] = DBL_MAX;
        for (long k = 0; k < K; ++k) {
            double cy = cent[5 * k + 0], cx = cent[5 * k + 1];
            if (isnan(cy) || isnan(cx)) continue;       /* empty cluster: never wins */
            double lo;
            lo = cy - 2 * step; long y_min = (long)(lo > 0 ? lo : 0);
            lo = cy + 2 * step + 1; long y_max = (long)(lo < H ? lo : H);
            lo = cx - 2 * step; long x_min = (long)(lo > 0 ? lo : 0);
            lo = cx + 2 * step + 1; long x_max = (long)(lo < W ? lo : W);
            for (long y = y_min; y < y_max; ++y) {
                double dy = (cy - (double)y); dy = dy * dy;
                for (long x = x_min; x < x_max; ++x) {
                    double dx = (cx - (double)x); dx = dx * dx;
                    double d = (dy + dx) * spatial_weight;
                    const double *px = img + 3 * (y * W + x);
                    double dc = 0.0;
                    for (int c = 0; c < 3; ++c) {
                        double t = px[c] - cent[5 * k + 2 + c];
                        dc = dc + t * t;
                    }
                    d = d + dc;
                    if (dist[y * W + x] > d) {
                        nearest[y * W + x] = (int32_t)k;
                        dist[y * W + x] = d;
                        change = 1;
                    }
                }
            }
        }
        if (!change) break;
        memset(count, 0, sizeof(long) * K);
        for (long i = 0; i < 5 * K; ++i) cent[i] = 0.0;
        for (long y = 0; y < H; ++y)
            for (long x = 0; x < W; ++x) {
                long k = nearest[y * W + x];
                count[k] += 1;
                cent[5 * k + 0] += (double)y;
                cent[5 * k + 1] += (double)x;
                const double *px = img + 3 * (y * W + x);
                cent[5 * k + 2] += px[0];
                cent[5 * k + 3] += px[1];
                cent[5 * k + 4] += px[2];
            }
        for (long k = 0; k < K; ++k)
            for (int c = 0; c < 5; ++c)
                cent[5 * k + c] = cent[5 * k + c] / (double)count[k];   /* 0/0 -> NaN */
    }
    free(dist); free(count);
}

/* ---- _enforce_label_connectivity_cython ---- */
static long enforce_connectivity(const int32_t *seg, int H, int W, long min_size, long max_size,
                                 int32_t *out)
{
    static const int ddx[4] = {1, -1, 0, 0};
    static const int ddy[4] = {0, 0, 1, -1};
    long n_px = (long)H * W;
    for (long p = 0; p < n_px; ++p) out[p] = -1;
    long cap = max_size > 1 ? max_size : 1;
    long *qy = (long *)malloc(sizeof(long) * cap);
    long *qx = (long *)malloc(sizeof(long) * cap);
    int32_t next_label = 0;
    for (long y = 0; y < H; ++y)
        for (long x = 0; x < W; ++x) {
            if (out[y * W + x] >= 0) continue;
            int32_t adjacent = 0;
            int32_t label = seg[y * W + x];
            out[y * W + x] = next_label;
            long size = 1, visited = 0;
            qy[0] = y; qx[0] = x;
            while (visited < size && size < max_size) {
                for (int i = 0; i < 4; ++i) {
                    long yy = qy[visited] + ddy[i];
                    long xx = qx[visited] + ddx[i];
                    if (xx < 0 || xx >= W || yy < 0 || yy >= H) continue;
                    long q = yy * W + xx;
                    if (seg[q] == label && out[q] == -1) {
                        out[q] = next_label;
                        qy[size] = yy; qx[size] = xx;
                        size += 1;
                        if (size >= max_size) break;
                    } else if (out[q] >= 0 && out[q] != next_label) {
                        adjacent = out[q];
                    }
                }
                visited += 1;
            }
            if (size < min_size) {
                for (long i = 0; i < size; ++i) out[qy[i] * W + qx[i]] = adjacent;
            } else {
                next_label += 1;
            }
        }
    free(qy); free(qx);
    return next_label;
}

/* Public entry: rgb (H,W,3) float32 in [0,1] -> labels (H,W) int32, 0-based,
 * contiguous, numbered in raster order.  `raw` (optional) receives the k-means
 * assignment before connectivity enforcement; `cent_out` (optional, K*5) the
 * final centres.  Returns the number of output labels, or -1 on bad input. */
long slic_ref(const float *rgb_hwc, int H, int W, int n_segments, double compactness,
              int max_iter, int enforce, int32_t *labels, int32_t *raw, double *cent_out)
{
    if (H <= 0 || W <= 0 || n_segments <= 0 || compactness <= 0) return -1;
    long n_px = (long)H * W;
    int step, start;
    long K = slic_ref_grid(H, W, n_segments, &step, &start);
    if (K <= 0) return -1;
    double *lab = (double *)malloc(sizeof(double) * 3 * n_px);
    rgb_to_lab(rgb_hwc, lab, n_px);
    double ratio = 1.0 / compactness;
    for (long i = 0; i < 3 * n_px; ++i) lab[i] = lab[i] * ratio;
    double *cent = (double *)calloc(5 * K, sizeof(double));
    long k = 0;
    for (long y = start; y < H; y += step)
        for (long x = start; x < W; x += step) {
            cent[5 * k + 0] = (double)y; cent[5 * k + 1] = (double)x; ++k;   /* colour starts at 0 */
        }
    int32_t *nearest = (int32_t *)calloc(n_px, sizeof(int32_t));
    slic_iterate(lab, H, W, cent, K, step, max_iter, nearest);
    if (raw) memcpy(raw, nearest, sizeof(int32_t) * n_px);
    if (cent_out) memcpy(cent_out, cent, sizeof(double) * 5 * K);
    long n_out;
    if (enforce) {
        double segment_size = (double)n_px / (double)n_segments;
        long min_size = (long)(0.5 * segment_size);
        long max_size = (long)(3.0 * segment_size);
        n_out = enforce_connectivity(nearest, H, W, min_size, max_size, labels);
    } else {
        memcpy(labels, nearest, sizeof(int32_t) * n_px);
        n_out = K;
    }
    free(lab); free(cent); free(nearest);
    return n_out;
}

/* Connectivity pass alone (used to test the CUDA connectivity kernels on
 * arbitrary label maps). */
long slic_ref_connectivity(const int32_t *seg, int H, int W, long min_size, long max_size,
                           int32_t *out)
{
    return enforce_connectivity(seg, H, W, min_size, max_size, out);
}

/* Lab conversion alone, scaled by 1/compactness. */
void slic_ref_lab(const float *rgb_hwc, long n_px, double compactness, double *lab)
{
    rgb_to_lab(rgb_hwc, lab, n_px);
    double ratio = 1.0 / compactness;
    for (long i = 0; i < 3 * n_px; ++i) lab[i] = lab[i] * ratio;
}
