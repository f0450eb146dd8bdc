// (d) GPU SLIC.  Replaces the CPU call skimage.segmentation.slic(img,
// n_segments=int(H*W/sp_area), compactness=sp_compactness) and its
// GPU->CPU->GPU round trip (/root/reference/models/wesup.py:471-478).
// Algorithm = scikit-image's (rgb2lab, regular-grid seeds with zero initial
// colour, max_iter k-means sweeps over 2S windows with "lowest cluster index
// wins ties", raster-order connectivity enforcement with the min_size merge AND
// the max_size cut) as restated in oracle/slic_ref.c; distances are IEEE double
// in the oracle's operation order (explicit __dadd_rn/__dmul_rn, nothing
// contracted into FMAs).
//
// Three launches per call, two of them persistent and cooperative, for a BATCH of B same-sized images:
//   slic_lab_kernel      rgb -> Lab (elementwise);
//   slic_kmeans_kernel   seeds and all max_iter sweeps; the phases of a
//                        sweep (assignment + accumulation | centre update + binning)
//                        are separated by grid barriers instead of kernel boundaries;
//   slic_connect_kernel  tile-local union-find in shared memory, border merges,
//                        component sizes, max_size split, raster-order numbering
//                        (device-wide scan), min_size merge, relabel -- six barriers.
// The sequential "for each centre, sweep its window" loop is pixel-centric: a
// 16x16 pixel tile gathers the centres whose window intersects it and every pixel
// takes the arg-min with the lowest-index tie break, which is what the
// centre-ordered strict `<` sweep computes.  Cluster sums are exact integers
// (counts, coordinates) and 2^-s fixed point (colours): integer atomics commute,
// so the result is bit-reproducible run to run and independent of the batch size.
#include "common.cuh"
#include <math_constants.h>
#include <stdlib.h>

namespace wesup {

typedef long long i64;

constexpr int BIN_CAP = 8;          // centres per bin before spilling to the overflow list
constexpr int AT = 16;              // tile edge (one pixel per thread, 256 threads)
constexpr int MAX_CAND = 192;       // candidate centres kept in shared memory per tile
constexpr int KM_BLOCKS_PER_SM = 6;

static inline size_t up256(size_t x) { return (x + 255) / 256 * 256; }

// ---------------------------------------------------------------------------
// grid barrier (the kernels are launched cooperatively: all blocks are resident).
// One monotone counter, zeroed by the host before the launch.
// ---------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// the barrier word sits in a 256-byte slot; block 0 stamps %globaltimer after every barrier into the u64s behind it
// (slot 1 + phase, up to 30): wesup_slic_debug_times reads them back (tools/slic_phases.py)
__device__ __forceinline__ void grid_barrier(unsigned *counter, unsigned &phase) {
    __syncthreads();
    if (threadIdx.x == 0) {
        phase += 1;
        const unsigned target = phase * gridDim.x;
        __threadfence();
        atomicAdd(counter, 1u);
        unsigned seen;
        while (true) {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
            if (seen >= target) break;
            __nanosleep(40);
        }
        __threadfence();
        if (blockIdx.x == 0 && phase < 31) reinterpret_cast<unsigned long long *>(counter)[phase] = global_timer_ns();
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------
// k-means
// ---------------------------------------------------------------------------
struct KmParams {
    const float *rgb;
    int layout, B, H, W, K, nx, step, start, cells_y, cells_x, tiles_y, tiles_x, max_iter;
    long HW;
    double ratio, spatial_weight, qscale, qinv;
    double *lab;            // B * 3 planes * HW (scaled by 1/compactness)
    double *cent;           // B * K * 5: y, x, L, a, b
    i64 *acc;               // B * K * 6: count, sum y, sum x, fixed-point sums of L, a, b
    int32_t *bin_count[2];  // B * cells           (double-buffered: sweep t reads [t&1], its update fills [(t+1)&1])
    int32_t *bin_items[2];  // B * cells * BIN_CAP
    int32_t *ov_count[2];   // B
    int32_t *ov_items[2];   // B * K
    int32_t *nearest;       // B * HW raw assignment (the caller's label buffer when connectivity is off)
    int32_t *n_labels;      // B, written (= K) when write_n
    int write_n;
    unsigned *bar;
};

// Host-side mirror of skimage.util.regular_grid for a (1,H,W) volume
// (oracle/slic_ref.c:slic_ref_grid).
static long slic_grid(int H, int W, int n_segments, int *step, int *start, int *ny, int *nx) {
    double space = (double)H * (double)W;
    if (space <= (double)n_segments) { *step = 1; *start = 0; *ny = H; *nx = W; return (long)H * W; }
    double s = sqrt(space / (double)n_segments);
    if ((double)(H < W ? H : W) < s) return -1;
    *start = (int)floor(s / 2.0);
    *step = (int)nearbyint(s);
    if (*step < 1) *step = 1;
    *ny = (H - *start + *step - 1) / *step;
    *nx = (W - *start + *step - 1) / *step;
    if (*ny <= 0 || *nx <= 0) return -1;
    return (long)(*ny) * (*nx);
}

__device__ __forceinline__ void rgb_to_lab(double r, double g, double b, double ratio, double &L, double &A, double &Bv) {
    double lin[3] = {r, g, b};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const double v = lin[c];
        lin[c] = (v > 0.04045) ? pow(__ddiv_rn(__dadd_rn(v, 0.055), 1.055), 2.4) : __ddiv_rn(v, 12.92);
    }
    const double M[3][3] = {{0.412453, 0.357580, 0.180423}, {0.212671, 0.715160, 0.072169}, {0.019334, 0.119193, 0.950227}};
    const double white[3] = {0.95047, 1.0, 1.08883};
    double f[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        double acc = __dmul_rn(lin[0], M[q][0]);
        acc = __dadd_rn(acc, __dmul_rn(lin[1], M[q][1]));
        acc = __dadd_rn(acc, __dmul_rn(lin[2], M[q][2]));
        const double t = __ddiv_rn(acc, white[q]);
        f[q] = (t > 0.008856) ? cbrt(t) : __dadd_rn(__dmul_rn(7.787, t), 16.0 / 116.0);
    }
    L = __dmul_rn(__dadd_rn(__dmul_rn(116.0, f[1]), -16.0), ratio);
    A = __dmul_rn(__dmul_rn(500.0, __dadd_rn(f[0], -f[1])), ratio);
    Bv = __dmul_rn(__dmul_rn(200.0, __dadd_rn(f[1], -f[2])), ratio);
}

// rgb -> Lab scaled by 1/compactness, three planes per image (its own launch: pow / cbrt would otherwise set the
// register budget of the persistent sweep kernel)
__global__ void __launch_bounds__(256) slic_lab_kernel(const float *__restrict__ rgb, int layout, long HW, long total,
                                                       double ratio, double *__restrict__ lab) {
    const long i = (long)blockIdx.x * 256 + threadIdx.x;
    if (i >= total) return;
    const long b = i / HW, p = i - b * HW;
    const float *img = rgb + b * 3 * HW;
    double r, g, bl;
    if (layout == WESUP_CHW) { r = (double)img[p]; g = (double)img[HW + p]; bl = (double)img[2 * HW + p]; }
    else { r = (double)img[3 * p]; g = (double)img[3 * p + 1]; bl = (double)img[3 * p + 2]; }
    double L, A, Bv;
    rgb_to_lab(r, g, bl, ratio, L, A, Bv);
    double *o = lab + b * 3 * HW;
    o[p] = L; o[HW + p] = A; o[2 * HW + p] = Bv;
}

// candidates of a tile, structure-of-arrays; "near" centres (inside the tile grown by half a
// grid step) fill the list from the front, the others from the back, so that every pixel meets
// its likely winners first and the spatial lower bound prunes most of the rest
struct Cands {
    double cy[MAX_CAND], cx[MAX_CAND], l[MAX_CAND], a[MAX_CAND], b[MAX_CAND];
    double lb[MAX_CAND];                         // spatial term at the tile pixel nearest to the centre: a lower bound for the tile
    int k[MAX_CAND];
    unsigned ywin[MAX_CAND], xwin[MAX_CAND];     // lo | hi << 16 of the centre's clipped 2S window
};

__global__ void __launch_bounds__(256, KM_BLOCKS_PER_SM) slic_kmeans_kernel(const KmParams P) {
    __shared__ Cands cand;
    __shared__ int n_near_s, n_far_s;
    const int tid = threadIdx.x, lane = tid & 31;
    const int H = P.H, W = P.W, K = P.K, step = P.step;
    const long HW = P.HW;
    const long cells = (long)P.cells_y * P.cells_x;
    const long tiles = (long)P.tiles_y * P.tiles_x;
    const long total_tiles = tiles * P.B;
    const long gtid = (long)blockIdx.x * 256 + tid, gthreads = (long)gridDim.x * 256;
    unsigned phase = 0;
    if (gtid == 0) reinterpret_cast<unsigned long long *>(P.bar)[31] = global_timer_ns();

    // ---- phase 0: seeds and the seeds' bins (cell (i,j) holds exactly seed i*nx+j); the raw assignment starts at 0 ----
    for (long i = gtid; i < (long)P.B * HW; i += gthreads) P.nearest[i] = 0;
    for (long i = gtid; i < (long)P.B * K; i += gthreads) {
        const int k = (int)(i % K);
        double *c = P.cent + 5 * i;
        c[0] = (double)(P.start + (k / P.nx) * step);
        c[1] = (double)(P.start + (k % P.nx) * step);
        c[2] = 0.0; c[3] = 0.0; c[4] = 0.0;
        i64 *a = P.acc + 6 * i;
#pragma unroll
        for (int j = 0; j < 6; ++j) a[j] = 0;
    }
    {
        const int ny = K / P.nx;
        for (long i = gtid; i < (long)P.B * cells; i += gthreads) {
            const int c = (int)(i % cells), cy = c / P.cells_x, cx = c % P.cells_x;
            const bool has = cy < ny && cx < P.nx;
            P.bin_count[0][i] = has ? 1 : 0;
            if (has) P.bin_items[0][i * BIN_CAP] = cy * P.nx + cx;
        }
        if (gtid < P.B) { P.ov_count[0][gtid] = 0; if (P.write_n) P.n_labels[gtid] = K; }
    }
    grid_barrier(P.bar, phase);

    const double two_s = (double)(2 * step);
    for (int it = 0; it < P.max_iter; ++it) {
        const int cur = it & 1, nxt = cur ^ 1;
        // the bins the update of this sweep will fill
        for (long i = gtid; i < (long)P.B * cells; i += gthreads) P.bin_count[nxt][i] = 0;
        if (gtid < P.B) P.ov_count[nxt][gtid] = 0;
        // ---- assignment + accumulation of the new cluster sums, tile by tile -------------------
        for (long t = blockIdx.x; t < total_tiles; t += gridDim.x) {
            const int b = (int)(t / tiles), tt = (int)(t % tiles);
            const int tx0 = (tt % P.tiles_x) * AT, ty0 = (tt / P.tiles_x) * AT;
            const int x = tx0 + (tid & (AT - 1)), y = ty0 + (tid >> 4);
            const bool live = x < W && y < H;
            const long p = (long)y * W + x;
            const double *lab = P.lab + (long)b * 3 * HW;
            const double *cent = P.cent + (long)b * 5 * K;
            int32_t *nearest = P.nearest + (long)b * HW;
            double pl = 0, pa = 0, pb = 0;
            if (live) { pl = lab[p]; pa = lab[HW + p]; pb = lab[2 * HW + p]; }
            if (tid == 0) { n_near_s = 0; n_far_s = 0; }
            __syncthreads();
            // gather the centres whose 2S window intersects this tile
            {
                const double half = (double)(step / 2 + 1);
                const double ny0 = (double)ty0 - half, ny1 = (double)(ty0 + AT) + half, nx0 = (double)tx0 - half, nx1 = (double)(tx0 + AT) + half;
                auto consider = [&](int k) {
                    const double cy = cent[5 * k], cx = cent[5 * k + 1];
                    double lo;
                    lo = cy - two_s;       const int y0 = (int)(lo > 0.0 ? lo : 0.0);
                    lo = cy + two_s + 1.0; const int y1 = (int)(lo < (double)H ? lo : (double)H);
                    lo = cx - two_s;       const int x0 = (int)(lo > 0.0 ? lo : 0.0);
                    lo = cx + two_s + 1.0; const int x1 = (int)(lo < (double)W ? lo : (double)W);
                    if (y0 < ty0 + AT && y1 > ty0 && x0 < tx0 + AT && x1 > tx0) {
                        const bool near = cy >= ny0 && cy < ny1 && cx >= nx0 && cx < nx1;
                        const int slot = near ? atomicAdd(&n_near_s, 1) : MAX_CAND - 1 - atomicAdd(&n_far_s, 1);
                        if (slot >= 0 && slot < MAX_CAND) {       // on overflow the list is abandoned (exhaustive scan below)
                            cand.cy[slot] = cy; cand.cx[slot] = cx;
                            cand.l[slot] = cent[5 * k + 2]; cand.a[slot] = cent[5 * k + 3]; cand.b[slot] = cent[5 * k + 4];
                            cand.k[slot] = k;
                            cand.ywin[slot] = (unsigned)y0 | ((unsigned)y1 << 16);
                            cand.xwin[slot] = (unsigned)x0 | ((unsigned)x1 << 16);
                            // the same expression the pixels evaluate, at the tile row / column nearest to the centre: every
                            // pixel of the tile is at least that far on both axes and rounding is monotone
                            const double ye = cy < (double)ty0 ? (double)ty0 : (cy > (double)(ty0 + AT - 1) ? (double)(ty0 + AT - 1) : cy);
                            const double xe = cx < (double)tx0 ? (double)tx0 : (cx > (double)(tx0 + AT - 1) ? (double)(tx0 + AT - 1) : cx);
                            double ey = __dadd_rn(cy, -ye); ey = __dmul_rn(ey, ey);
                            double ex = __dadd_rn(cx, -xe); ex = __dmul_rn(ex, ex);
                            cand.lb[slot] = __dmul_rn(__dadd_rn(ey, ex), P.spatial_weight);
                        }
                    }
                };
                // a centre can reach the tile only from cells within 2S+1 pixels of it
                const int reach = 2 * step + 1;
                const int by0 = max((ty0 - reach) / step - 1, 0), by1 = min((ty0 + AT - 1 + reach) / step, P.cells_y - 1);
                const int bx0 = max((tx0 - reach) / step - 1, 0), bx1 = min((tx0 + AT - 1 + reach) / step, P.cells_x - 1);
                const int nbx = bx1 - bx0 + 1, ncell = (by1 - by0 + 1) * nbx;
                const int32_t *bcount = P.bin_count[cur] + (long)b * cells;
                const int32_t *bitems = P.bin_items[cur] + (long)b * cells * BIN_CAP;
                for (int i = tid; i < ncell * BIN_CAP; i += 256) {
                    const int cell = (by0 + (i / BIN_CAP) / nbx) * P.cells_x + bx0 + (i / BIN_CAP) % nbx;
                    const int slot = i % BIN_CAP;
                    if (slot < min(bcount[cell], BIN_CAP)) consider(bitems[(long)cell * BIN_CAP + slot]);
                }
                const int n_ov = P.ov_count[cur][b];
                const int32_t *ov = P.ov_items[cur] + (long)b * K;
                for (int i = tid; i < n_ov; i += 256) consider(ov[i]);
            }
            __syncthreads();
            const int n_near = n_near_s, n_far = n_far_s;
            const bool overflow = n_near + n_far > MAX_CAND;
            double best = CUDART_INF;
            int best_k = -1;
            auto evaluate = [&](int i) {
                if (cand.lb[i] > best) return;             // no pixel of this tile can be closer to this centre than lb
                const unsigned yw = cand.ywin[i], xw = cand.xwin[i];
                if (y < (int)(yw & 0xffffu) || y >= (int)(yw >> 16) || x < (int)(xw & 0xffffu) || x >= (int)(xw >> 16)) return;
                double dy = __dadd_rn(cand.cy[i], -(double)y); dy = __dmul_rn(dy, dy);
                double dx = __dadd_rn(cand.cx[i], -(double)x); dx = __dmul_rn(dx, dx);
                double d = __dmul_rn(__dadd_rn(dy, dx), P.spatial_weight);
                // d only grows when the (non-negative) colour term is added and rounding is monotone:
                // a centre whose spatial term alone exceeds the best distance can neither win nor tie
                if (d > best) return;
                double t = __dadd_rn(pl, -cand.l[i]);
                double dc = __dmul_rn(t, t);                       // 0 + t*t
                t = __dadd_rn(pa, -cand.a[i]); dc = __dadd_rn(dc, __dmul_rn(t, t));
                t = __dadd_rn(pb, -cand.b[i]); dc = __dadd_rn(dc, __dmul_rn(t, t));
                d = __dadd_rn(d, dc);
                const int k = cand.k[i];
                if (d < best || (d == best && k < best_k)) { best = d; best_k = k; }
            };
            if (live && !overflow) {
                for (int i = 0; i < n_near; ++i) evaluate(i);
                for (int i = MAX_CAND - n_far; i < MAX_CAND; ++i) evaluate(i);
            }
            if (live && overflow) {
                // pathological crowding: the shared list overflowed; exhaustive scan over all centres
                // (same arithmetic, same tie break) so the result stays exact
                for (int k = 0; k < K; ++k) {
                    const double cy = cent[5 * k], cx = cent[5 * k + 1];
                    if (!(cy == cy) || !(cx == cx)) continue;
                    double lo;
                    lo = cy - two_s;       const int y0 = (int)(lo > 0.0 ? lo : 0.0);
                    lo = cy + two_s + 1.0; const int y1 = (int)(lo < (double)H ? lo : (double)H);
                    lo = cx - two_s;       const int x0 = (int)(lo > 0.0 ? lo : 0.0);
                    lo = cx + two_s + 1.0; const int x1 = (int)(lo < (double)W ? lo : (double)W);
                    if (y < y0 || y >= y1 || x < x0 || x >= x1) continue;
                    double dy = __dadd_rn(cy, -(double)y); dy = __dmul_rn(dy, dy);
                    double dx = __dadd_rn(cx, -(double)x); dx = __dmul_rn(dx, dx);
                    double d = __dmul_rn(__dadd_rn(dy, dx), P.spatial_weight);
                    double tq = __dadd_rn(pl, -cent[5 * k + 2]);
                    double dc = __dmul_rn(tq, tq);
                    tq = __dadd_rn(pa, -cent[5 * k + 3]); dc = __dadd_rn(dc, __dmul_rn(tq, tq));
                    tq = __dadd_rn(pb, -cent[5 * k + 4]); dc = __dadd_rn(dc, __dmul_rn(tq, tq));
                    d = __dadd_rn(d, dc);
                    if (d < best || (d == best && k < best_k)) { best = d; best_k = k; }
                }
            }
            // accumulate: warp-level pre-aggregation (a warp is a 16x2 strip of the tile and meets 1-4
            // clusters); the lanes of one cluster are reduced together and one lane issues six integer
            // atomics.  The last sweep's sums are never read: skip them.
            int kf = -1;
            if (live) {
                kf = best_k;
                if (kf >= 0) nearest[p] = kf; else kf = nearest[p];      // uncovered pixel keeps its previous cluster
            }
            if (it + 1 < P.max_iter) {
                const i64 ql = __double2ll_rn(__dmul_rn(pl, P.qscale)), qa = __double2ll_rn(__dmul_rn(pa, P.qscale)),
                          qb = __double2ll_rn(__dmul_rn(pb, P.qscale));
                i64 *acc = P.acc + (long)b * 6 * K;
                unsigned todo = __ballot_sync(0xffffffffu, kf >= 0);
                while (todo) {
                    const int leader = __ffs(todo) - 1;
                    const int key = __shfl_sync(0xffffffffu, kf, leader);
                    const bool mine = kf == key;
                    const unsigned grp = __ballot_sync(0xffffffffu, mine);
                    const unsigned sy = __reduce_add_sync(0xffffffffu, mine ? (unsigned)y : 0u);
                    const unsigned sx = __reduce_add_sync(0xffffffffu, mine ? (unsigned)x : 0u);
                    i64 sl = mine ? ql : 0, sa = mine ? qa : 0, sb = mine ? qb : 0;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        sl += __shfl_xor_sync(0xffffffffu, sl, o);
                        sa += __shfl_xor_sync(0xffffffffu, sa, o);
                        sb += __shfl_xor_sync(0xffffffffu, sb, o);
                    }
                    if (lane == leader) {
                        unsigned long long *a = reinterpret_cast<unsigned long long *>(acc + 6 * (long)key);
                        atomicAdd(a + 0, (unsigned long long)__popc(grp));
                        atomicAdd(a + 1, (unsigned long long)sy);
                        atomicAdd(a + 2, (unsigned long long)sx);
                        atomicAdd(a + 3, (unsigned long long)sl);
                        atomicAdd(a + 4, (unsigned long long)sa);
                        atomicAdd(a + 5, (unsigned long long)sb);
                    }
                    todo &= ~grp;
                }
            }
            __syncthreads();        // the candidate list is rebuilt by the next tile
        }
        if (it + 1 == P.max_iter) break;
        grid_barrier(P.bar, phase);
        // ---- centre update + binning for the next sweep ------------------------------------------
        for (long i = gtid; i < (long)P.B * K; i += gthreads) {
            const int b = (int)(i / K), k = (int)(i % K);
            i64 *a = P.acc + 6 * i;
            const double cnt = (double)a[0];
            // 0/0 -> NaN for an empty cluster, as in the sequential algorithm
            const double cy = __ddiv_rn((double)a[1], cnt), cx = __ddiv_rn((double)a[2], cnt);
            double *c = P.cent + 5 * i;
            c[0] = cy; c[1] = cx;
            c[2] = __ddiv_rn(__dmul_rn((double)a[3], P.qinv), cnt);
            c[3] = __ddiv_rn(__dmul_rn((double)a[4], P.qinv), cnt);
            c[4] = __ddiv_rn(__dmul_rn((double)a[5], P.qinv), cnt);
#pragma unroll
            for (int j = 0; j < 6; ++j) a[j] = 0;
            // bin by the cell that contains the centre; NaN centres (empty clusters) are not binned
            // and therefore never become candidates
            if (cy == cy && cx == cx) {
                int by = (int)(cy / (double)step), bx = (int)(cx / (double)step);
                by = min(max(by, 0), P.cells_y - 1);
                bx = min(max(bx, 0), P.cells_x - 1);
                const long cell = (long)b * cells + (long)by * P.cells_x + bx;
                const int slot = atomicAdd(&P.bin_count[nxt][cell], 1);
                if (slot < BIN_CAP) P.bin_items[nxt][cell * BIN_CAP + slot] = k;
                else P.ov_items[nxt][(long)b * K + atomicAdd(&P.ov_count[nxt][b], 1)] = k;
            }
        }
        grid_barrier(P.bar, phase);
    }
}

// ---------------------------------------------------------------------------
// connectivity enforcement (oracle/slic_ref.c:enforce_connectivity)
// ---------------------------------------------------------------------------
constexpr int SMALL_CAP = 128;      // pieces up to this many pixels run the min_size search level-parallel out of shared memory

struct CcParams {
    const int32_t *seg;     // B * HW raw labels
    int B, H, W, min_size, max_size;
    long HW;
    int32_t *parent;        // B * HW union-find / piece root (min pixel id of the piece)
    int32_t *size;          // B * HW piece size at its root (0 elsewhere once flattened)
    int32_t *keep_scan;     // B * HW at kept roots: raster-order id
    int32_t *adj_root;      // B * HW at small roots: root of the piece it merges into, or -1
    int32_t *link;          // B * HW FIFO links of the max_size split (written before they are read)
    int32_t *mark_small;    // B * HW linked-list queue of the min_size search for pieces above SMALL_CAP (-1 = untouched)
    int32_t *block_sums;    // B * blocks_per_image
    int32_t *labels;        // B * HW output
    int32_t *n_labels;      // B
    unsigned *bar;
    int32_t *flags;         // [0]: some component was cut at max_size (zeroed with the barrier word)
};

constexpr int Q_NONE = -1, Q_END = -2;

__device__ __forceinline__ int uf_find(const int32_t *parent, int a) {
    int r = a;
    const volatile int32_t *vp = parent;      // links change under us: never serve them from a stale L1 line
    while (true) {
        int pr = vp[r];
        if (pr == r) return r;
        r = pr;
    }
}
__device__ __forceinline__ void uf_union(int32_t *parent, int a, int b) {
    while (true) {
        a = uf_find(parent, a);
        b = uf_find(parent, b);
        if (a == b) return;
        if (a > b) { int t = a; a = b; b = t; }
        int old = atomicMin(&parent[b], a);      // hook the larger root under the smaller
        if (old == b) return;
        b = old;
    }
}
__device__ __forceinline__ int sm_find(const volatile int *par, int a) {
    while (true) { int p = par[a]; if (p == a) return a; a = p; }
}
__device__ __forceinline__ void sm_union(int *par, int a, int b) {
    while (true) {
        a = sm_find(par, a);
        b = sm_find(par, b);
        if (a == b) return;
        if (a > b) { int t = a; a = b; b = t; }
        int old = atomicMin(&par[b], a);
        if (old == b) return;
        b = old;
    }
}

// The sequential algorithm caps every breadth-first search at max_size pixels and starts a new piece at
// the next unvisited pixel in raster order.  Replayed here for one raw component larger than max_size by
// ONE WARP executing uniformly (every lane performs the same loads and stores, so each lane observes its
// own writes); only the raster search for the next start uses the lanes in parallel.  "Unvisited" = the
// pixel still points at the raw root r; a visited pixel points at the start of its piece (the first piece,
// whose start IS r, is parked on a negative code and restored at the end).  FIFO queue = links in link[].
__device__ void split_large_component(const CcParams &P, const int32_t *seg, int32_t *parent, int32_t *size, int32_t *link,
                                      int r, int lane) {
    const int W = P.W, H = P.H;
    const int lab = seg[r];
    int remaining = size[r];
    int start = r;
    const long HW = P.HW;
    const int parked = -2 - r;
    while (true) {
        const int code = start == r ? parked : start;
        int cnt = 1, head = start, tail = start;
        parent[start] = code;
        while (cnt < P.max_size) {
            const int y = head / W, x = head - y * W;
            const int nb[4] = {x + 1 < W ? head + 1 : -1, x > 0 ? head - 1 : -1, y + 1 < H ? head + W : -1, y > 0 ? head - W : -1};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int n = nb[i];
                if (n < 0 || cnt >= P.max_size) continue;
                if (seg[n] == lab && parent[n] == r) {
                    link[tail] = n; tail = n;
                    parent[n] = code;
                    cnt += 1;
                }
            }
            if (head == tail) break;              // queue exhausted
            head = link[head];
        }
        size[start] = cnt;
        remaining -= cnt;
        if (remaining <= 0) break;
        // next piece: first pixel after `start` in raster order that still points at the raw root
        long base = (long)start + 1;
        int found = -1;
        while (found < 0 && base < HW) {
            const long q = base + lane;
            const bool hit = q < HW && parent[q] == r && seg[q] == lab;
            const unsigned m = __ballot_sync(0xffffffffu, hit);
            if (m) found = (int)(base + __ffs(m) - 1);
            base += 32;
        }
        if (found < 0) break;            // cannot happen: `remaining` pixels are still unvisited
        start = found;
    }
    // un-park the first piece
    const int n1 = size[r];
    int q = r;
    for (int i = 0; i < n1; ++i) {
        const int nxt = i + 1 < n1 ? link[q] : q;
        parent[q] = r;
        q = nxt;
    }
}

// A piece below min_size takes the label of the LAST already-labelled neighbour the sequential breadth-first
// search saw (neighbour order +x,-x,+y,-y; FIFO), i.e. the last neighbour that belongs to a piece with a smaller
// root (pieces are labelled in root order).  The search order is reproduced level by level by one warp: lane i
// holds the i-th pixel of the current level (in search order) and fetches its four neighbours' roots, so a level
// costs one memory round trip instead of one per pixel.  A neighbour joins the next level through the proposal
// with the smallest key (position of the proposer in search order * 4 + direction) -- exactly the pixel / direction
// that discovers it sequentially -- and the next level is ordered by those keys.  Keys grow from level to level, so
// an atomicMin never disturbs a pixel that was discovered earlier.  Keys live in a 32 x 64 window of shared memory
// (rows below the root, 32 columns either side); a piece that leaves the window reports failure and takes the
// sequential path.
constexpr int WIN_ROWS = 32, WIN_COLS = 64;

__device__ int small_piece_adjacent(const CcParams &P, const int32_t *parent, int r, int lane, int *px, int *slot, bool &ok) {
    const int W = P.W, H = P.H;
    const int ry = r / W, rx = r - ry * W;
    for (int i = lane; i < WIN_ROWS * WIN_COLS; i += 32) slot[i] = INT_MAX;
    __syncwarp();
    if (lane == 0) { px[0] = r; slot[WIN_COLS / 2] = 0; }
    __syncwarp();
    int lvl_start = 0, lvl_end = 1, adj = -1;
    ok = true;
    while (lvl_start < lvl_end) {
        int new_end = lvl_end;
        for (int base = lvl_start; base < lvl_end; base += 32) {
            const int idx = base + lane;
            const bool act = idx < lvl_end;
            const int p = act ? px[idx] : 0;
            const int y = p / W, x = p - y * W;
            int n[4], rn[4], at[4];
            n[0] = act && x + 1 < W ? p + 1 : -1;
            n[1] = act && x > 0 ? p - 1 : -1;
            n[2] = act && y + 1 < H ? p + W : -1;
            n[3] = act && y > 0 ? p - W : -1;
#pragma unroll
            for (int k = 0; k < 4; ++k) rn[k] = n[k] >= 0 ? parent[n[k]] : INT_MAX;
            int hit = -1;
            bool outside = false;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                at[k] = -1;
                if (n[k] < 0) continue;
                if (rn[k] < r) hit = rn[k];                     // the last direction with a labelled neighbour
                else if (rn[k] == r) {
                    const int ny = n[k] / W, dy = ny - ry, dx = n[k] - ny * W - rx + WIN_COLS / 2;
                    if (dy < 0 || dy >= WIN_ROWS || dx < 0 || dx >= WIN_COLS) outside = true;
                    else at[k] = dy * WIN_COLS + dx;
                }
            }
            if (__any_sync(0xffffffffu, outside)) { ok = false; return -1; }
            const unsigned hm = __ballot_sync(0xffffffffu, hit >= 0);
            if (hm) adj = __shfl_sync(0xffffffffu, hit, 31 - __clz(hm));          // the latest pixel in search order wins
#pragma unroll
            for (int k = 0; k < 4; ++k) if (at[k] >= 0) atomicMin(&slot[at[k]], idx * 4 + k);
            __syncwarp();
            int wins = 0;
            bool win[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) { win[k] = at[k] >= 0 && slot[at[k]] == idx * 4 + k; wins += win[k] ? 1 : 0; }
            int incl = wins;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
            const int total = __shfl_sync(0xffffffffu, incl, 31);
            if (new_end + total > SMALL_CAP) { ok = false; return -1; }
            int pos = new_end + incl - wins;
#pragma unroll
            for (int k = 0; k < 4; ++k) if (win[k]) px[pos++] = n[k];
            new_end += total;
            __syncwarp();
        }
        lvl_start = lvl_end;
        lvl_end = new_end;
    }
    return adj;
}

// The same search, sequentially through global memory (FIFO = linked list threaded through mark[]), executed
// uniformly by all lanes of a warp (each lane observes its own writes): pieces above SMALL_CAP pixels or wider
// than the shared-memory window.
__device__ int small_piece_adjacent_sequential(const CcParams &P, const int32_t *parent, int32_t *mark, int r) {
    const int W = P.W, H = P.H;
    int head = r, tail = r, adj = -1;
    mark[r] = Q_END;
    while (head >= 0) {
        const int y = head / W, x = head - y * W;
        const int nb[4] = {x + 1 < W ? head + 1 : -1, x > 0 ? head - 1 : -1, y + 1 < H ? head + W : -1, y > 0 ? head - W : -1};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (nb[k] < 0) continue;
            const int rn = parent[nb[k]];
            if (rn == r) {
                if (mark[nb[k]] == Q_NONE) { mark[nb[k]] = Q_END; mark[tail] = nb[k]; tail = nb[k]; }
            } else if (rn < r) {
                adj = rn;
            }
        }
        const int nxt = mark[head];
        head = nxt == Q_END ? -1 : nxt;
    }
    return adj;
}

__global__ void __launch_bounds__(256, 4) slic_connect_kernel(const CcParams P) {
    __shared__ int s_lab[256], s_par[256], s_cnt[256];
    __shared__ int s_warp[9];
    constexpr int PIECE_WARPS = 4;                       // warps of a block that run the min_size searches
    __shared__ int s_px[PIECE_WARPS][SMALL_CAP];
    __shared__ int s_slot[PIECE_WARPS][WIN_ROWS * WIN_COLS];
    __shared__ int s_small[256], s_small_sz[256], s_nsmall;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int H = P.H, W = P.W;
    const long HW = P.HW;
    const int tiles_x = (W + AT - 1) / AT, tiles_y = (H + AT - 1) / AT;
    const long tiles = (long)tiles_x * tiles_y, total_tiles = tiles * P.B;
    unsigned phase = 0;
    if (blockIdx.x == 0 && tid == 0) reinterpret_cast<unsigned long long *>(P.bar)[31] = global_timer_ns();

    // ---- C1: tile-local components in shared memory: horizontal runs by ballot, union-find over run starts ----
    for (long t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int b = (int)(t / tiles), tt = (int)(t % tiles);
        const int lx = tid & (AT - 1), ly = tid >> 4;
        const int x = (tt % tiles_x) * AT + lx, y = (tt / tiles_x) * AT + ly;
        const bool live = x < W && y < H;
        const long p = (long)b * HW + (long)y * W + x;
        const int lab = live ? P.seg[p] : 0;
        s_lab[tid] = lab; s_cnt[tid] = 0;
        __syncthreads();
        // a live pixel's left / up neighbours inside the tile are live too
        const bool left = live && lx > 0 && s_lab[tid - 1] == lab;
        const bool up = live && ly > 0 && s_lab[tid - AT] == lab;
        const unsigned starts = __ballot_sync(0xffffffffu, !left);           // a warp = two 16-pixel rows; lx == 0 always starts a run
        const int run_lane = 31 - __clz(starts & (0xffffffffu >> (31 - lane)));
        const int run = (tid & ~31) + run_lane;
        s_par[tid] = run;
        __syncthreads();
        // vertical merges, once per pair of touching runs: skipped where the pixel to the left has made the same one
        if (up && !(left && s_lab[tid - AT - 1] == lab)) sm_union(s_par, run, s_par[tid - AT]);
        __syncthreads();
        const int r = sm_find(s_par, tid);
        if (live && run == tid) {                                            // run starts carry their run's length to the root
            const unsigned half = lane < 16 ? 0x0000ffffu : 0xffff0000u;
            const unsigned above = starts & half & ~((2u << lane) - 1u);
            const int next = above ? __ffs(above) - 1 : (lane | 15) + 1;
            atomicAdd(&s_cnt[r], next - lane);
        }
        __syncthreads();
        if (live) {
            const int rx = (tt % tiles_x) * AT + (r & (AT - 1)), ry = (tt / tiles_x) * AT + (r >> 4);
            P.parent[p] = ry * W + rx;              // per-image pixel id of the tile-local root (the smallest id of the piece)
            P.size[p] = r == tid ? s_cnt[tid] : 0;
            P.mark_small[p] = Q_NONE;
        }
        __syncthreads();
    }
    grid_barrier(P.bar, phase);
    // ---- C2: merges across tile borders (global union-find, atomicMin hooking), once per pair of touching runs.
    // The border pixels of all tiles form one flat index space (per image: (tiles_x-1)*H pixels on vertical borders,
    // then (tiles_y-1)*W on horizontal ones), so every thread has work and the dependent find/hook chains of many
    // tiles overlap. ----
    {
        const long nv = (long)(tiles_x - 1) * H, nh = (long)(tiles_y - 1) * W, per_img = nv + nh;
        for (long i = (long)blockIdx.x * 256 + tid; i < per_img * P.B; i += (long)gridDim.x * 256) {
            const int b = (int)(i / per_img);
            const long j = i - (long)b * per_img;
            const int32_t *seg = P.seg + (long)b * HW;
            int32_t *parent = P.parent + (long)b * HW;
            if (j < nv) {
                // pixel (y, x) on the left edge of tile column c >= 1, merged with (y, x-1)
                const int y = (int)(j % H), x = ((int)(j / H) + 1) * AT;
                const int p = y * W + x;
                const int lab = seg[p];
                // implied when the pixels above both carry the label too (they are joined inside their tiles and made
                // this merge themselves)
                if (seg[p - 1] == lab && !((y & (AT - 1)) != 0 && seg[p - W] == lab && seg[p - W - 1] == lab))
                    uf_union(parent, p, p - 1);
            } else {
                const long jj = j - nv;
                const int x = (int)(jj % W), y = ((int)(jj / W) + 1) * AT;
                const int p = y * W + x;
                const int lab = seg[p];
                if (seg[p - W] == lab && !((x & (AT - 1)) != 0 && seg[p - 1] == lab && seg[p - W - 1] == lab))
                    uf_union(parent, p, p - W);
            }
        }
    }
    grid_barrier(P.bar, phase);
    // ---- C3: flatten; tile-local counts flow to the global root -----------------------------------
    for (long i = (long)blockIdx.x * 256 + tid; i < HW * P.B; i += (long)gridDim.x * 256) {
        const long off = (i / HW) * HW;
        const int p = (int)(i - off);
        int32_t *parent = P.parent + off;
        const int r = uf_find(parent, p);
        parent[p] = r;                      // racy-but-monotone path compression: every value written is an ancestor
        const int c = P.size[off + p];
        if (c > 0 && r != p) { atomicAdd(&P.size[off + r], c); P.size[off + p] = 0; }     // only roots carry a size from here on
    }
    grid_barrier(P.bar, phase);
    // ---- C4: components above max_size are cut the way the capped sequential search cuts them; the same pass
    // counts the kept roots of every contiguous chunk (raster-order numbering = exclusive scan of those flags).
    // A cut changes the flags under the counting threads: the count is then redone after the barrier. ----
    const int bpi = max((int)gridDim.x / P.B, 1);                   // blocks per image
    const long chunk = ((HW + bpi - 1) / bpi + 255) / 256 * 256;
    const int my_img = blockIdx.x / bpi, my_chunk = blockIdx.x % bpi;
    const bool scanning = my_img < P.B;                              // leftover blocks idle in the chunked phases
    const int img_step = max((int)gridDim.x / bpi, 1);               // (when B > gridDim.x the images are processed in rounds)
    for (int pass = 0; pass < 2; ++pass) {
        if (pass == 1 && *reinterpret_cast<volatile int32_t *>(P.flags) == 0) break;     // no cut: the counts stand
        for (int img0 = 0; img0 < P.B; img0 += img_step) {
            const int b = img0 + my_img;
            const bool active = scanning && b < P.B;
            int total = 0;
            const long lo = (long)my_chunk * chunk, hi = active ? min(lo + chunk, HW) : lo;
            const long off = (long)b * HW;
            for (long q0 = lo; q0 < hi; q0 += 256) {
                const long q = q0 + tid;
                bool root = false, large = false;
                int sz = 0;
                if (q < hi) { root = P.parent[off + q] == (int)q; sz = root ? P.size[off + q] : 0; large = pass == 0 && sz > P.max_size; }
                unsigned m = __ballot_sync(0xffffffffu, large);
                if (m && lane == 0) *P.flags = 1;
                while (m) {
                    const int src = __ffs(m) - 1;
                    const int r = (int)(q0 + (tid & ~31) + src);
                    split_large_component(P, P.seg + off, P.parent + off, P.size + off, P.link + off, r, lane);
                    m &= m - 1;
                }
                total += (root && !large && sz >= P.min_size) ? 1 : 0;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
            if (lane == 0) s_warp[warp] = total;
            __syncthreads();
            if (tid == 0 && active) {
                int sum = 0;
                for (int w = 0; w < 8; ++w) sum += s_warp[w];
                P.block_sums[(long)b * bpi + my_chunk] = sum;
            }
            __syncthreads();
        }
        grid_barrier(P.bar, phase);
    }
    // ---- C5: the scan itself; C6: every small piece finds the piece it merges into ----------------
    for (int img0 = 0; img0 < P.B; img0 += img_step) {
        const int b = img0 + my_img;
        if (!(scanning && b < P.B)) continue;            // uniform per block
        // offset of this chunk = sum of the earlier chunks' counts
        int part = 0, all = 0;
        for (int c = tid; c < bpi; c += 256) {
            const int v = P.block_sums[(long)b * bpi + c];
            all += v;
            if (c < my_chunk) part += v;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { part += __shfl_xor_sync(0xffffffffu, part, o); all += __shfl_xor_sync(0xffffffffu, all, o); }
        __syncthreads();
        if (lane == 0) { s_lab[warp] = part; s_par[warp] = all; }
        __syncthreads();
        int running = 0, grand = 0;
        for (int w = 0; w < 8; ++w) { running += s_lab[w]; grand += s_par[w]; }
        if (my_chunk == 0 && tid == 0) P.n_labels[b] = grand;
        const long lo = (long)my_chunk * chunk, hi = min(lo + chunk, HW);
        const long off = (long)b * HW;
        const int32_t *parent = P.parent + off, *size = P.size + off;
        int32_t *keep_scan = P.keep_scan + off;
        for (long q0 = lo; q0 < hi; q0 += 256) {
            const long q = q0 + tid;
            const bool root = q < hi && parent[q] == (int)q;
            const int sz = root ? size[q] : 0;
            const bool keep = root && sz >= P.min_size;
            const unsigned m = __ballot_sync(0xffffffffu, keep);
            __syncthreads();
            if (lane == 0) s_warp[warp] = __popc(m);
            __syncthreads();
            int before = running;
            for (int w = 0; w < warp; ++w) before += s_warp[w];
            if (keep) keep_scan[q] = before + __popc(m & ((1u << lane) - 1u));
            for (int w = 0; w < 8; ++w) running += s_warp[w];
            // C6 for the small roots among these pixels: listed, then searched by the first PIECE_WARPS warps
            if (tid == 0) s_nsmall = 0;
            __syncthreads();
            if (root && !keep) { const int slot_i = atomicAdd(&s_nsmall, 1); s_small[slot_i] = (int)q; s_small_sz[slot_i] = sz; }
            __syncthreads();
            if (warp < PIECE_WARPS) {
                const int n_small = s_nsmall;
                for (int i = warp; i < n_small; i += PIECE_WARPS) {
                    const int r = s_small[i];
                    bool ok = false;
                    int adj = -1;
                    if (s_small_sz[i] <= SMALL_CAP) adj = small_piece_adjacent(P, parent, r, lane, s_px[warp], s_slot[warp], ok);
                    if (!ok) adj = small_piece_adjacent_sequential(P, parent, P.mark_small + off, r);
                    if (lane == 0) P.adj_root[off + r] = adj;
                }
            }
        }
        __syncthreads();
    }
    grid_barrier(P.bar, phase);
    // ---- C7: relabel: follow the chain of small pieces down to a kept one (roots strictly decrease) ----
    for (long i = (long)blockIdx.x * 256 + tid; i < HW * P.B; i += (long)gridDim.x * 256) {
        const long off = (i / HW) * HW;
        int r = P.parent[i];
        while (r >= 0 && P.size[off + r] < P.min_size) r = P.adj_root[off + r];
        P.labels[i] = r >= 0 ? P.keep_scan[off + r] : 0;
    }
    if (blockIdx.x == 0 && tid == 0) reinterpret_cast<unsigned long long *>(P.bar)[30] = global_timer_ns();
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
// blocks of `kernel` that are resident at once on the current device (cached per kernel slot and device)
static int coresident_blocks(const void *kernel, int block_threads, int slot, int *out) {
    static int cache[2][64];
    int dev = 0, sms = 0, per_sm = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e == cudaSuccess && dev >= 0 && dev < 64 && cache[slot][dev] > 0) { *out = cache[slot][dev]; return 0; }
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block_threads, 0);
    if (e != cudaSuccess) { set_error("occupancy query: %s", cudaGetErrorString(e)); return (int)e; }
    // Resident blocks per SM: 3 by default (WESUP_SLIC_BLOCKS_PER_SM overrides, 0 = as many as fit).  A cooperative
    // launch that fills the SMs (6 blocks of 256 threads, 61 K registers) can only start once the device has drained
    // and then shuts out the kernels of every other stream, but SLIC runs one image AHEAD of the training graph on a
    // side stream precisely to overlap with it.  Measured (r2, 464^2): alone 0.280 ms per image at full residency,
    // 0.300 ms at 3 blocks per SM (each block loops over two tiles); training step 15.78 -> 15.46 ms.
    {
        const char *env = getenv("WESUP_SLIC_BLOCKS_PER_SM");
        const int lim = env ? atoi(env) : 3;
        if (lim > 0 && lim < per_sm) per_sm = lim;
    }
    *out = sms * per_sm;
    if (*out <= 0) { set_error("kernel does not fit an SM"); return WESUP_E_UNSUPPORTED; }
    if (dev >= 0 && dev < 64) cache[slot][dev] = *out;
    return 0;
}

struct CcLayout { size_t parent, size, keep_scan, adj_root, link, mark_small, block_sums, bar, flags, total; };
static CcLayout cc_layout(long B, long HW) {
    CcLayout L;
    size_t o = 0;
    const size_t px = up256(sizeof(int32_t) * B * HW);
    L.parent = o; o += px;  L.size = o; o += px;  L.keep_scan = o; o += px;  L.adj_root = o; o += px;
    L.link = o; o += px;  L.mark_small = o; o += px;
    L.block_sums = o; o += up256(sizeof(int32_t) * 4096);
    L.bar = o; o += 256;
    L.flags = o; o += 256;
    L.total = o;
    return L;
}

static int launch_connectivity(const int32_t *seg, int B, int H, int W, int min_size, int max_size, int32_t *labels,
                               int32_t *n_labels, char *ws, cudaStream_t stream) {
    const long HW = (long)H * W;
    const CcLayout L = cc_layout(B, HW);
    CcParams P;
    P.seg = seg; P.B = B; P.H = H; P.W = W; P.min_size = min_size; P.max_size = max_size; P.HW = HW;
    P.parent = (int32_t *)(ws + L.parent); P.size = (int32_t *)(ws + L.size); P.keep_scan = (int32_t *)(ws + L.keep_scan);
    P.adj_root = (int32_t *)(ws + L.adj_root); P.link = (int32_t *)(ws + L.link); P.flags = (int32_t *)(ws + L.flags);
    P.mark_small = (int32_t *)(ws + L.mark_small); P.block_sums = (int32_t *)(ws + L.block_sums);
    P.labels = labels; P.n_labels = n_labels; P.bar = (unsigned *)(ws + L.bar);
    int cap = 0;
    int rc = coresident_blocks((const void *)slic_connect_kernel, 256, 1, &cap);
    if (rc) return rc;
    const long work = ((HW + 255) / 256) * B;
    int grid = (int)(work < cap ? work : cap);
    if (grid > 4096) grid = 4096;                 // block_sums capacity
    if (grid < B) {                               // the scan phases want at least one block per image
        grid = B < cap ? B : cap;
    }
    WESUP_REQUIRE(grid >= 1, WESUP_E_UNSUPPORTED, "wesup_slic: no resident blocks");
    cudaError_t e = cudaMemsetAsync(P.bar, 0, 512, stream);             // barrier word + flags
    WESUP_REQUIRE(e == cudaSuccess, (int)e, "wesup_slic: memset: %s", cudaGetErrorString(e));
    void *args[] = {(void *)&P};
    e = cudaLaunchCooperativeKernel((const void *)slic_connect_kernel, dim3(grid), dim3(256), args, 0, stream);
    WESUP_REQUIRE(e == cudaSuccess, (int)e, "wesup_slic: connectivity launch: %s", cudaGetErrorString(e));
    return 0;
}

struct KmLayout { size_t lab, cent, acc, bin_count[2], bin_items[2], ov_count[2], ov_items[2], nearest, bar, cc, total; };
static KmLayout km_layout(long B, long HW, long K, long cells) {
    KmLayout L;
    size_t o = 0;
    L.lab = o; o += up256(sizeof(double) * 3 * B * HW);
    L.cent = o; o += up256(sizeof(double) * 5 * B * K);
    L.acc = o; o += up256(sizeof(i64) * 6 * B * K);
    for (int j = 0; j < 2; ++j) {
        L.bin_count[j] = o; o += up256(sizeof(int32_t) * B * cells);
        L.bin_items[j] = o; o += up256(sizeof(int32_t) * B * cells * BIN_CAP);
        L.ov_count[j] = o; o += up256(sizeof(int32_t) * B);
        L.ov_items[j] = o; o += up256(sizeof(int32_t) * B * K);
    }
    L.nearest = o; o += up256(sizeof(int32_t) * B * HW);
    L.bar = o; o += 256;
    L.cc = o; o += cc_layout(B, HW).total;
    L.total = o;
    return L;
}

}  // namespace wesup

using namespace wesup;

extern "C" size_t wesup_slic_batch_workspace_bytes(int B, int H, int W, int n_segments) {
    if (B <= 0 || H <= 0 || W <= 0 || n_segments <= 0) return 0;
    int step, start, ny, nx;
    long K = slic_grid(H, W, n_segments, &step, &start, &ny, &nx);
    if (K <= 0) return 0;
    const long cells = (long)((H + step - 1) / step) * ((W + step - 1) / step);
    return km_layout(B, (long)H * W, K, cells).total;
}
extern "C" size_t wesup_slic_workspace_bytes(int H, int W, int n_segments) {
    return wesup_slic_batch_workspace_bytes(1, H, W, n_segments);
}
extern "C" size_t wesup_enforce_connectivity_workspace_bytes(int B, int H, int W) {
    if (B <= 0 || H <= 0 || W <= 0) return 0;
    return cc_layout(B, (long)H * W).total;
}

extern "C" int wesup_enforce_connectivity(const int32_t *seg, int B, int H, int W, int min_size, int max_size,
                                          int32_t *labels, int32_t *n_labels, void *ws, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    WESUP_REQUIRE(seg && labels && n_labels && ws, WESUP_E_ARG, "wesup_enforce_connectivity: null pointer");
    WESUP_REQUIRE(B > 0 && B <= 1024 && H > 0 && W > 0 && min_size >= 0 && max_size >= 1, WESUP_E_ARG,
                  "wesup_enforce_connectivity: bad argument B=%d H=%d W=%d min_size=%d max_size=%d", B, H, W, min_size, max_size);
    WESUP_REQUIRE(min_size <= max_size, WESUP_E_UNSUPPORTED,
                  "wesup_enforce_connectivity: min_size %d > max_size %d (the capped search would cut pieces that are then merged)",
                  min_size, max_size);
    WESUP_REQUIRE((long)H * W < (1L << 31), WESUP_E_UNSUPPORTED, "wesup_enforce_connectivity: H*W must fit int32");
    int rc = launch_connectivity(seg, B, H, W, min_size, max_size, labels, n_labels, static_cast<char *>(ws), stream);
    if (rc) return rc;
    WESUP_CHECK_LAUNCH("wesup_enforce_connectivity", 1);
    return 0;
}

extern "C" int wesup_slic_batch(const float *rgb, int rgb_layout, int B, int H, int W, int n_segments, double compactness,
                                int max_iter, int enforce_connectivity, int32_t *labels, int32_t *n_labels, void *ws,
                                void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    WESUP_REQUIRE(rgb && labels && n_labels && ws, WESUP_E_ARG, "wesup_slic: null pointer");
    WESUP_REQUIRE(B > 0 && B <= 1024 && H > 0 && W > 0 && n_segments > 0 && compactness > 0 && max_iter >= 0, WESUP_E_ARG,
                  "wesup_slic: bad argument B=%d H=%d W=%d n_segments=%d compactness=%g", B, H, W, n_segments, compactness);
    WESUP_REQUIRE(rgb_layout == WESUP_CHW || rgb_layout == WESUP_HWC, WESUP_E_ARG, "wesup_slic: bad layout %d", rgb_layout);
    WESUP_REQUIRE((long)H * W < (1L << 31), WESUP_E_UNSUPPORTED, "wesup_slic: H*W must fit int32");
    WESUP_REQUIRE(H < 65536 && W < 65536, WESUP_E_UNSUPPORTED, "wesup_slic: H and W must be below 65536");
    int step, start, ny, nx;
    const long K = slic_grid(H, W, n_segments, &step, &start, &ny, &nx);
    WESUP_REQUIRE(K > 0, WESUP_E_UNSUPPORTED, "wesup_slic: degenerate seed grid for %dx%d / %d segments", H, W, n_segments);
    const long HW = (long)H * W;
    KmParams P;
    P.cells_y = (H + step - 1) / step; P.cells_x = (W + step - 1) / step;
    const long cells = (long)P.cells_y * P.cells_x;
    const KmLayout L = km_layout(B, HW, K, cells);
    char *base = static_cast<char *>(ws);
    P.rgb = rgb; P.layout = rgb_layout; P.B = B; P.H = H; P.W = W; P.K = (int)K; P.nx = nx; P.step = step; P.start = start;
    P.tiles_y = (H + AT - 1) / AT; P.tiles_x = (W + AT - 1) / AT; P.max_iter = max_iter; P.HW = HW;
    P.ratio = 1.0 / compactness;
    const float stepf = (float)step;
    P.spatial_weight = 1.0 / (double)(stepf * stepf);
    // fixed-point scale of the colour sums: |v| <= 256/compactness for rgb in [0,1] (Lab magnitudes stay below 128, 2x
    // margin); a cluster never holds more than max_iter windows' worth of pixels (a pixel joins only from inside the window,
    // uncovered pixels keep their cluster); the running sum must stay below 2^62
    {
        const double win = (double)(4 * step + 1) * (double)(4 * step + 1) * (double)(max_iter > 1 ? max_iter : 1);
        const double cnt = win < (double)HW ? win : (double)HW;
        const int cnt_bits = (int)ceil(log2(cnt + 1.0));
        const int v_bits = (int)ceil(log2(256.0 * P.ratio));
        int s = 62 - cnt_bits - v_bits;
        if (s > 52) s = 52;
        if (s < 0) s = 0;
        P.qscale = ldexp(1.0, s);
        P.qinv = ldexp(1.0, -s);
    }
    P.lab = (double *)(base + L.lab); P.cent = (double *)(base + L.cent); P.acc = (i64 *)(base + L.acc);
    for (int j = 0; j < 2; ++j) {
        P.bin_count[j] = (int32_t *)(base + L.bin_count[j]); P.bin_items[j] = (int32_t *)(base + L.bin_items[j]);
        P.ov_count[j] = (int32_t *)(base + L.ov_count[j]); P.ov_items[j] = (int32_t *)(base + L.ov_items[j]);
    }
    P.nearest = enforce_connectivity ? (int32_t *)(base + L.nearest) : labels;
    P.n_labels = n_labels; P.write_n = enforce_connectivity ? 0 : 1;
    P.bar = (unsigned *)(base + L.bar);
    int cap = 0;
    int rc = coresident_blocks((const void *)slic_kmeans_kernel, 256, 0, &cap);
    if (rc) return rc;
    const long total_tiles = (long)P.tiles_y * P.tiles_x * B;
    const int grid = (int)(total_tiles < cap ? total_tiles : cap);
    cudaError_t e = cudaMemsetAsync(P.bar, 0, 256, stream);
    WESUP_REQUIRE(e == cudaSuccess, (int)e, "wesup_slic: memset: %s", cudaGetErrorString(e));
    slic_lab_kernel<<<cdiv(HW * B, 256), 256, 0, stream>>>(rgb, rgb_layout, HW, HW * B, P.ratio, P.lab);
    void *args[] = {(void *)&P};
    e = cudaLaunchCooperativeKernel((const void *)slic_kmeans_kernel, dim3(grid), dim3(256), args, 0, stream);
    WESUP_REQUIRE(e == cudaSuccess, (int)e, "wesup_slic: k-means launch: %s", cudaGetErrorString(e));
    if (!enforce_connectivity) {
        WESUP_CHECK_LAUNCH("wesup_slic", 2);
        return 0;
    }
    const double segment_size = (double)HW / (double)n_segments;
    const int min_size = (int)(0.5 * segment_size);
    int max_size = (int)(3.0 * segment_size);
    if (max_size < 1) max_size = 1;
    rc = launch_connectivity(P.nearest, B, H, W, min_size, max_size, labels, n_labels, base + L.cc, stream);
    if (rc) return rc;
    WESUP_CHECK_LAUNCH("wesup_slic", 3);
    return 0;
}

// Debug: %globaltimer stamps of block 0 (ns) of the last call that used `ws`: out[0..31] k-means kernel
// (31 = start, 1.. = after each barrier), out[32..63] connectivity kernel (63 = start, 33.. = barriers, 62 = end).
// Synchronises the device; measurement tooling only.
extern "C" int wesup_slic_debug_times(const void *ws, int B, int H, int W, int n_segments, unsigned long long *out_host) {
    int step, start, ny, nx;
    const long K = slic_grid(H, W, n_segments, &step, &start, &ny, &nx);
    WESUP_REQUIRE(ws && out_host && K > 0, WESUP_E_ARG, "wesup_slic_debug_times: bad argument");
    const long HW = (long)H * W;
    const long cells = (long)((H + step - 1) / step) * ((W + step - 1) / step);
    const KmLayout L = km_layout(B, HW, K, cells);
    const char *base = static_cast<const char *>(ws);
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaMemcpy(out_host, base + L.bar, 256, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(out_host + 32, base + L.cc + cc_layout(B, HW).bar, 256, cudaMemcpyDeviceToHost);
    WESUP_REQUIRE(e == cudaSuccess, (int)e, "wesup_slic_debug_times: %s", cudaGetErrorString(e));
    return 0;
}

extern "C" int wesup_slic(const float *rgb, int rgb_layout, int H, int W, int n_segments, double compactness,
                          int max_iter, int enforce_connectivity, int32_t *labels, int32_t *n_labels, void *ws,
                          void *stream_) {
    return wesup_slic_batch(rgb, rgb_layout, 1, H, W, n_segments, compactness, max_iter, enforce_connectivity, labels,
                            n_labels, ws, stream_);
}
