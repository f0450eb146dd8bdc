// (d) GPU SLIC.  Replaces the CPU call skimage.segmentation.slic(img,
// n_segments=int(H*W/sp_area), compactness=sp_compactness) and its
// GPU->CPU->GPU round trip (/root/reference/models/wesup.py:471-478).
// Algorithm = scikit-image's (rgb2lab, regular-grid seeds with zero initial
// colour, max_iter k-means sweeps over 2S windows with "lowest cluster index
// wins ties", raster-order connectivity enforcement) as restated in
// oracle/slic_ref.c; arithmetic is IEEE double in the same operation order
// (explicit __dadd_rn/__dmul_rn so nothing is contracted into FMAs).
//
// The sequential "for each centre, sweep its window" loop becomes
// pixel-centric: a 16x16 pixel tile gathers the centres whose window
// intersects it and every pixel takes the arg-min with the lowest-index tie
// break, which is what the centre-ordered strict `<` sweep computes.
#include "common.cuh"
#include <math_constants.h>

namespace wesup {

typedef unsigned long long u64;

constexpr int BIN_CAP = 8;          // centres per bin before spilling to the overflow list

struct SlicWs {
    double *lab;        // 3 planes, H*W each (already scaled by 1/compactness)
    double *cent;       // K*5: y,x,L,a,b
    // ---- region zeroed by one memset at the start of every call ----
    u64 *acc_n;         // K*3: count, sum y, sum x  (exact integer sums)
    double *acc_c;      // K*3: sum L,a,b
    int32_t *bin_count; // cells: centres binned by floor(c / step)
    int32_t *ov_count;  // 1: length of the overflow list
    uint32_t *ticket;   // 1: blocks that finished the current sweep
    // -----------------------------------------------------------------
    int32_t *bin_items; // cells*BIN_CAP
    int32_t *ov_items;  // K
    int32_t *nearest;   // H*W raw k-means assignment
    int32_t *parent;    // H*W union-find / component root (min pixel id)
    int32_t *size;      // H*W component size at its root
    int32_t *keep_scan; // H*W exclusive scan of "kept root" flags
    int32_t *small_scan;// H*W exclusive scan of small-component sizes at roots
    int32_t *queue;     // H*W BFS scratch for small components
    int32_t *adj_root;  // H*W (at roots) root of the adjacent component or -1
    int32_t *block_sums;// scan scratch
    uint8_t *seen;      // H*W
    size_t zero_bytes;  // length of the zeroed region (starts at acc_n)
    int cells_y, cells_x;
};

static inline size_t up256(size_t x) { return (x + 255) / 256 * 256; }
constexpr int SCAN_ELEMS = 4096;    // per block (1024 threads x 4)

static inline long n_cells(int H, int W, int step, int *cy, int *cx) {
    *cy = (H + step - 1) / step;
    *cx = (W + step - 1) / step;
    return (long)(*cy) * (*cx);
}

static size_t slic_ws_bytes(long HW, long K, long cells) {
    long nblk = (HW + SCAN_ELEMS - 1) / SCAN_ELEMS + 1;
    return up256(sizeof(double) * 3 * HW) + up256(sizeof(double) * 5 * K) + up256(sizeof(u64) * 3 * K) +
           up256(sizeof(double) * 3 * K) + up256(sizeof(int32_t) * cells) + 256 + 256 +
           up256(sizeof(int32_t) * cells * BIN_CAP) + up256(sizeof(int32_t) * K) +
           7 * up256(sizeof(int32_t) * HW) + up256(sizeof(int32_t) * 2 * nblk) + up256(HW);
}

static SlicWs carve_slic(void *ws, long HW, long K, long cells) {
    char *p = static_cast<char *>(ws);
    SlicWs s;
    long nblk = (HW + SCAN_ELEMS - 1) / SCAN_ELEMS + 1;
    s.lab = (double *)p;        p += up256(sizeof(double) * 3 * HW);
    s.cent = (double *)p;       p += up256(sizeof(double) * 5 * K);
    char *z0 = p;
    s.acc_n = (u64 *)p;         p += up256(sizeof(u64) * 3 * K);
    s.acc_c = (double *)p;      p += up256(sizeof(double) * 3 * K);
    s.bin_count = (int32_t *)p; p += up256(sizeof(int32_t) * cells);
    s.ov_count = (int32_t *)p;  p += 256;
    s.ticket = (uint32_t *)p;   p += 256;
    s.zero_bytes = (size_t)(p - z0);
    s.bin_items = (int32_t *)p; p += up256(sizeof(int32_t) * cells * BIN_CAP);
    s.ov_items = (int32_t *)p;  p += up256(sizeof(int32_t) * K);
    s.nearest = (int32_t *)p;   p += up256(sizeof(int32_t) * HW);
    s.parent = (int32_t *)p;    p += up256(sizeof(int32_t) * HW);
    s.size = (int32_t *)p;      p += up256(sizeof(int32_t) * HW);
    s.keep_scan = (int32_t *)p; p += up256(sizeof(int32_t) * HW);
    s.small_scan = (int32_t *)p;p += up256(sizeof(int32_t) * HW);
    s.queue = (int32_t *)p;     p += up256(sizeof(int32_t) * HW);
    s.adj_root = (int32_t *)p;  p += up256(sizeof(int32_t) * HW);
    s.block_sums = (int32_t *)p;p += up256(sizeof(int32_t) * 2 * nblk);
    s.seen = (uint8_t *)p;
    return s;
}

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

// ---------------------------------------------------------------------------
// rgb -> Lab, scaled by 1/compactness
// ---------------------------------------------------------------------------
__global__ void slic_lab_kernel(const float *__restrict__ rgb, int layout, long HW, double ratio, double *__restrict__ lab) {
    long p = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= HW) return;
    double lin[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        double v = (double)(layout == WESUP_CHW ? rgb[(long)c * HW + p] : rgb[3 * p + c]);
        lin[c] = (v > 0.04045) ? pow(__ddiv_rn(__dadd_rn(v, 0.055), 1.055), 2.4) : __ddiv_rn(v, 12.92);
    }
    const double M[3][3] = {{0.412453, 0.357580, 0.180423}, {0.212671, 0.715160, 0.072169}, {0.019334, 0.119193, 0.950227}};
    const double white[3] = {0.95047, 1.0, 1.08883};
    double f[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        double acc = __dmul_rn(lin[0], M[r][0]);
        acc = __dadd_rn(acc, __dmul_rn(lin[1], M[r][1]));
        acc = __dadd_rn(acc, __dmul_rn(lin[2], M[r][2]));
        double t = __ddiv_rn(acc, white[r]);
        f[r] = (t > 0.008856) ? cbrt(t) : __dadd_rn(__dmul_rn(7.787, t), 16.0 / 116.0);
    }
    lab[p] = __dmul_rn(__dadd_rn(__dmul_rn(116.0, f[1]), -16.0), ratio);
    lab[HW + p] = __dmul_rn(__dmul_rn(500.0, __dadd_rn(f[0], -f[1])), ratio);
    lab[2 * HW + p] = __dmul_rn(__dmul_rn(200.0, __dadd_rn(f[1], -f[2])), ratio);
}

// Bin a centre by the cell that contains it (cell edge = step pixels); NaN centres
// (empty clusters) are not binned and therefore never become candidates.
__device__ __forceinline__ void bin_insert(const SlicWs &s, int k, double cy, double cx, int step) {
    if (!(cy == cy) || !(cx == cx)) return;
    int by = (int)(cy / (double)step), bx = (int)(cx / (double)step);
    by = min(max(by, 0), s.cells_y - 1);
    bx = min(max(bx, 0), s.cells_x - 1);
    const int cell = by * s.cells_x + bx;
    const int slot = atomicAdd(&s.bin_count[cell], 1);
    if (slot < BIN_CAP) s.bin_items[cell * BIN_CAP + slot] = k;
    else s.ov_items[atomicAdd(s.ov_count, 1)] = k;
}

// seeds (and their bins) + the per-pixel state of the connectivity pass; the
// accumulators / bin counters were zeroed by the memset that precedes this launch
__global__ void slic_init_kernel(SlicWs s, long K, int nx, int step, int start, long HW) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < K) {
        const double cy = (double)(start + (int)(i / nx) * step), cx = (double)(start + (int)(i % nx) * step);
        s.cent[5 * i + 0] = cy;
        s.cent[5 * i + 1] = cx;
        s.cent[5 * i + 2] = 0.0; s.cent[5 * i + 3] = 0.0; s.cent[5 * i + 4] = 0.0;
        bin_insert(s, (int)i, cy, cx, step);
    }
    if (i < HW) { s.nearest[i] = 0; s.parent[i] = (int)i; s.size[i] = 0; s.seen[i] = 0; s.adj_root[i] = -1; }
}

// ---------------------------------------------------------------------------
// One k-means sweep = ONE launch: assignment, accumulation of the new cluster
// sums (pre-aggregated per tile in shared memory), and -- in the block that
// finishes last -- the centre update and the re-binning for the next sweep.
// ---------------------------------------------------------------------------
constexpr int AT = 16;              // tile edge
constexpr int MAX_CAND = 192;       // candidate centres kept in shared memory per tile

// candidates of a tile, structure-of-arrays; "near" centres (inside the tile grown by half a
// grid step) fill the list from the front, the others from the back, so that every pixel meets
// its likely winners first and the spatial lower bound prunes most of the rest
struct Cands {
    double cy[MAX_CAND], cx[MAX_CAND], l[MAX_CAND], a[MAX_CAND], b[MAX_CAND];
    int k[MAX_CAND];
    unsigned ywin[MAX_CAND], xwin[MAX_CAND];     // lo | hi << 16 of the centre's clipped 2S window
};

__global__ void __launch_bounds__(256, 6) slic_sweep_kernel(SlicWs s, int H, int W, long K, int step, double spatial_weight) {
    __shared__ Cands cand;
    __shared__ int n_near_s, n_far_s, is_last;
    const int tid = threadIdx.x, lane = tid & 31;
    const int tx0 = blockIdx.x * AT, ty0 = blockIdx.y * AT;
    const int x = tx0 + (tid & (AT - 1)), y = ty0 + (tid >> 4);
    const bool live = x < W && y < H;
    const long HW = (long)H * W;
    const long p = (long)y * W + x;
    double pl = 0, pa = 0, pb = 0;
    if (live) { pl = s.lab[p]; pa = s.lab[HW + p]; pb = s.lab[2 * HW + p]; }
    if (tid == 0) { n_near_s = 0; n_far_s = 0; }
    __syncthreads();
    // ---- gather the centres whose 2S window intersects this tile -------------
    const double two_s = (double)(2 * step);
    {
        const double half = (double)(step / 2 + 1);
        const double ny0 = (double)ty0 - half, ny1 = (double)(ty0 + AT) + half, nx0 = (double)tx0 - half, nx1 = (double)(tx0 + AT) + half;
        auto consider = [&](int k) {
            const double cy = s.cent[5 * k], cx = s.cent[5 * k + 1];
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
                    cand.l[slot] = s.cent[5 * k + 2]; cand.a[slot] = s.cent[5 * k + 3]; cand.b[slot] = s.cent[5 * k + 4];
                    cand.k[slot] = k;
                    cand.ywin[slot] = (unsigned)y0 | ((unsigned)y1 << 16);
                    cand.xwin[slot] = (unsigned)x0 | ((unsigned)x1 << 16);
                }
            }
        };
        // a centre can reach the tile only from cells within 2S+1 pixels of it
        const int reach = 2 * step + 1;
        const int by0 = max((ty0 - reach) / step - 1, 0), by1 = min((ty0 + AT - 1 + reach) / step, s.cells_y - 1);
        const int bx0 = max((tx0 - reach) / step - 1, 0), bx1 = min((tx0 + AT - 1 + reach) / step, s.cells_x - 1);
        const int nbx = bx1 - bx0 + 1, ncell = (by1 - by0 + 1) * nbx;
        for (int i = tid; i < ncell * BIN_CAP; i += 256) {
            const int cell = (by0 + (i / BIN_CAP) / nbx) * s.cells_x + bx0 + (i / BIN_CAP) % nbx;
            const int slot = i % BIN_CAP;
            if (slot < min(s.bin_count[cell], BIN_CAP)) consider(s.bin_items[cell * BIN_CAP + slot]);
        }
        const int n_ov = *s.ov_count;
        for (int i = tid; i < n_ov; i += 256) consider(s.ov_items[i]);
    }
    __syncthreads();
    const int n_near = n_near_s, n_far = n_far_s;
    const bool overflow = n_near + n_far > MAX_CAND;
    double best = CUDART_INF;
    int best_k = -1;
    auto evaluate = [&](int i) {
        const unsigned yw = cand.ywin[i], xw = cand.xwin[i];
        if (y < (int)(yw & 0xffffu) || y >= (int)(yw >> 16) || x < (int)(xw & 0xffffu) || x >= (int)(xw >> 16)) return;
        double dy = __dadd_rn(cand.cy[i], -(double)y); dy = __dmul_rn(dy, dy);
        double dx = __dadd_rn(cand.cx[i], -(double)x); dx = __dmul_rn(dx, dx);
        double d = __dmul_rn(__dadd_rn(dy, dx), spatial_weight);
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
    if (overflow) {
        // pathological crowding: the shared list overflowed; exhaustive scan over all centres
        // (same arithmetic, same tie break) so the result stays exact
        if (live) {
            for (long k = 0; k < K; ++k) {
                const double cy = s.cent[5 * k], cx = s.cent[5 * k + 1];
                if (!(cy == cy) || !(cx == cx)) continue;
                double lo;
                lo = cy - two_s;       const int y0 = (int)(lo > 0.0 ? lo : 0.0);
                lo = cy + two_s + 1.0; const int y1 = (int)(lo < (double)H ? lo : (double)H);
                lo = cx - two_s;       const int x0 = (int)(lo > 0.0 ? lo : 0.0);
                lo = cx + two_s + 1.0; const int x1 = (int)(lo < (double)W ? lo : (double)W);
                if (y < y0 || y >= y1 || x < x0 || x >= x1) continue;
                double dy = __dadd_rn(cy, -(double)y); dy = __dmul_rn(dy, dy);
                double dx = __dadd_rn(cx, -(double)x); dx = __dmul_rn(dx, dx);
                double d = __dmul_rn(__dadd_rn(dy, dx), spatial_weight);
                double t = __dadd_rn(pl, -s.cent[5 * k + 2]);
                double dc = __dmul_rn(t, t);
                t = __dadd_rn(pa, -s.cent[5 * k + 3]); dc = __dadd_rn(dc, __dmul_rn(t, t));
                t = __dadd_rn(pb, -s.cent[5 * k + 4]); dc = __dadd_rn(dc, __dmul_rn(t, t));
                d = __dadd_rn(d, dc);
                if (d < best || (d == best && (int)k < best_k)) { best = d; best_k = (int)k; }
            }
        }
    }
    // ---- accumulate the new cluster sums ---------------------------------------
    // Warp-level pre-aggregation (a warp is a 16x2 strip of the tile and meets 1-4 clusters):
    // the lanes of one cluster are reduced together -- integer sums with redux, colour sums with
    // a fixed butterfly -- and one lane issues the six global atomics.  No block barrier.
    int kf = -1;
    if (live) {
        kf = best_k;
        if (kf >= 0) s.nearest[p] = kf; else kf = s.nearest[p];      // uncovered pixel keeps its previous cluster
    }
    unsigned todo = __ballot_sync(0xffffffffu, kf >= 0);
    while (todo) {
        const int leader = __ffs(todo) - 1;
        const int key = __shfl_sync(0xffffffffu, kf, leader);
        const bool mine = kf == key;
        const unsigned grp = __ballot_sync(0xffffffffu, mine);
        const unsigned sy = __reduce_add_sync(0xffffffffu, mine ? (unsigned)y : 0u);
        const unsigned sx = __reduce_add_sync(0xffffffffu, mine ? (unsigned)x : 0u);
        double sl = mine ? pl : 0.0, sa = mine ? pa : 0.0, sb = mine ? pb : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            sl += __shfl_xor_sync(0xffffffffu, sl, o);
            sa += __shfl_xor_sync(0xffffffffu, sa, o);
            sb += __shfl_xor_sync(0xffffffffu, sb, o);
        }
        if (lane == leader) {
            atomicAdd(&s.acc_n[3 * key], (u64)__popc(grp));
            atomicAdd(&s.acc_n[3 * key + 1], (u64)sy);
            atomicAdd(&s.acc_n[3 * key + 2], (u64)sx);
            atomicAdd(&s.acc_c[3 * key], sl);
            atomicAdd(&s.acc_c[3 * key + 1], sa);
            atomicAdd(&s.acc_c[3 * key + 2], sb);
        }
        todo &= ~grp;
    }
    // ---- the last block to finish updates the centres and re-bins them ---------
    __threadfence();
    __syncthreads();
    if (tid == 0) is_last = (atomicAdd(s.ticket, 1u) == gridDim.x * gridDim.y - 1u) ? 1 : 0;
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    const long cells = (long)s.cells_y * s.cells_x;
    for (long i = tid; i < cells; i += 256) s.bin_count[i] = 0;
    if (tid == 0) { *s.ov_count = 0; *s.ticket = 0u; }
    __syncthreads();
    for (long k = tid; k < K; k += 256) {
        volatile u64 *an = s.acc_n + 3 * k;
        volatile double *ac = s.acc_c + 3 * k;
        const double cnt = (double)an[0];
        // 0/0 -> NaN for an empty cluster, as in the sequential algorithm
        const double cy = __ddiv_rn((double)an[1], cnt), cx = __ddiv_rn((double)an[2], cnt);
        s.cent[5 * k + 0] = cy;
        s.cent[5 * k + 1] = cx;
        s.cent[5 * k + 2] = __ddiv_rn(ac[0], cnt);
        s.cent[5 * k + 3] = __ddiv_rn(ac[1], cnt);
        s.cent[5 * k + 4] = __ddiv_rn(ac[2], cnt);
        an[0] = 0; an[1] = 0; an[2] = 0;
        ac[0] = 0.0; ac[1] = 0.0; ac[2] = 0.0;
        bin_insert(s, (int)k, cy, cx, step);
    }
}

// ---------------------------------------------------------------------------
// connectivity enforcement
// ---------------------------------------------------------------------------
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

__global__ void ccl_merge_kernel(SlicWs s, int H, int W) {
    long p = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= (long)H * W) return;
    int y = (int)(p / W), x = (int)(p - (long)y * W);
    int lab = s.nearest[p];
    if (x + 1 < W && s.nearest[p + 1] == lab) uf_union(s.parent, (int)p, (int)p + 1);
    if (y + 1 < H && s.nearest[p + W] == lab) uf_union(s.parent, (int)p, (int)p + W);
}
__global__ void ccl_flatten_kernel(SlicWs s, long HW) {
    long p = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= HW) return;
    int r = uf_find(s.parent, (int)p);
    s.parent[p] = r;             // racy-but-monotone path compression: every value written is an ancestor
    atomicAdd(&s.size[r], 1);
}
// flags to scan: kept roots (for raster-order numbering) and small-root sizes (BFS scratch offsets)
__global__ void ccl_flags_kernel(SlicWs s, long HW, int min_size) {
    long p = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= HW) return;
    bool root = s.parent[p] == (int)p;
    int sz = s.size[p];
    s.keep_scan[p] = (root && sz >= min_size) ? 1 : 0;
    s.small_scan[p] = (root && sz < min_size) ? sz : 0;
}

// device-wide exclusive scan (in place), three launches
__global__ void __launch_bounds__(1024) scan_block_kernel(int32_t *data, long n, int32_t *block_sums) {
    __shared__ int warp_tot[33];
    long base = (long)blockIdx.x * SCAN_ELEMS + (long)threadIdx.x * 4;
    int v[4], sum = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) { v[j] = (base + j < n) ? data[base + j] : 0; sum += v[j]; }
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = warp_tot[lane], wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += t; }
        warp_tot[lane] = wi - w;
        if (lane == 31) warp_tot[32] = wi;
    }
    __syncthreads();
    int excl = warp_tot[warp] + incl - sum;
#pragma unroll
    for (int j = 0; j < 4; ++j) { if (base + j < n) data[base + j] = excl; excl += v[j]; }
    if (threadIdx.x == 0) block_sums[blockIdx.x] = warp_tot[32];
}
__global__ void __launch_bounds__(1024) scan_sums_kernel(int32_t *block_sums, int nblk, int32_t *total_out) {
    __shared__ int warp_tot[33];
    int carry = 0;
    for (int start = 0; start < nblk; start += 1024) {
        int i = start + threadIdx.x;
        int v = i < nblk ? block_sums[i] : 0;
        int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        __syncthreads();
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int w = warp_tot[lane], wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += t; }
            warp_tot[lane] = wi - w;
            if (lane == 31) warp_tot[32] = wi;
        }
        __syncthreads();
        if (i < nblk) block_sums[i] = carry + warp_tot[warp] + incl - v;
        carry += warp_tot[32];
        __syncthreads();
    }
    if (threadIdx.x == 0 && total_out) *total_out = carry;
}
__global__ void __launch_bounds__(1024) scan_add_kernel(int32_t *data, long n, const int32_t *block_sums) {
    long base = (long)blockIdx.x * SCAN_ELEMS + (long)threadIdx.x * 4;
    int add = block_sums[blockIdx.x];
#pragma unroll
    for (int j = 0; j < 4; ++j) if (base + j < n) data[base + j] += add;
}

// One thread per small component: replay the sequential breadth-first search
// (neighbour order +x,-x,+y,-y; FIFO queue) to find the neighbour the
// sequential algorithm would have remembered as `adjacent`: the LAST
// already-labelled neighbour seen, i.e. the last neighbour that belongs to a
// component with a smaller root (components are labelled in root order).
__global__ void ccl_small_adjacent_kernel(SlicWs s, int H, int W, int min_size) {
    long r = (long)blockIdx.x * blockDim.x + threadIdx.x;
    long HW = (long)H * W;
    if (r >= HW) return;
    if (s.parent[r] != (int)r) return;
    int sz = s.size[r];
    if (sz >= min_size) return;
    int32_t *q = s.queue + s.small_scan[r];
    int head = 0, tail = 1;
    q[0] = (int)r;
    s.seen[r] = 1;
    int adj = -1;
    const int ddx[4] = {1, -1, 0, 0}, ddy[4] = {0, 0, 1, -1};
    while (head < tail) {
        int p = q[head++];
        int y = p / W, x = p - y * W;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int yy = y + ddy[i], xx = x + ddx[i];
            if (xx < 0 || xx >= W || yy < 0 || yy >= H) continue;
            int n = yy * W + xx;
            int rn = s.parent[n];
            if (rn == (int)r) {
                if (!s.seen[n]) { s.seen[n] = 1; q[tail++] = n; }
            } else if (rn < (int)r) {
                adj = rn;
            }
        }
    }
    s.adj_root[r] = adj;
}

__global__ void ccl_relabel_kernel(SlicWs s, long HW, int min_size, int32_t *__restrict__ labels) {
    long p = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= HW) return;
    int r = s.parent[p];
    // follow the chain of small components down to a kept one (roots strictly decrease)
    while (r >= 0 && s.size[r] < min_size) r = s.adj_root[r];
    labels[p] = r >= 0 ? s.keep_scan[r] : 0;
}
__global__ void copy_labels_kernel(const int32_t *__restrict__ src, int32_t *__restrict__ dst, long n, int32_t *n_labels, int K) {
    long p = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n) dst[p] = src[p];
    if (p == 0) *n_labels = K;
}

static void device_exclusive_scan(int32_t *data, long n, int32_t *block_sums, int32_t *total_out, cudaStream_t stream) {
    int nblk = (int)((n + SCAN_ELEMS - 1) / SCAN_ELEMS);
    scan_block_kernel<<<nblk, 1024, 0, stream>>>(data, n, block_sums);
    scan_sums_kernel<<<1, 1024, 0, stream>>>(block_sums, nblk, total_out);
    scan_add_kernel<<<nblk, 1024, 0, stream>>>(data, n, block_sums);
}

}  // namespace wesup

using namespace wesup;

extern "C" size_t wesup_slic_workspace_bytes(int H, int W, int n_segments) {
    if (H <= 0 || W <= 0 || n_segments <= 0) return 0;
    int step, start, ny, nx;
    long K = slic_grid(H, W, n_segments, &step, &start, &ny, &nx);
    if (K <= 0) return 0;
    int cy, cx;
    long cells = n_cells(H, W, step, &cy, &cx);
    return slic_ws_bytes((long)H * W, K, cells);
}

extern "C" int wesup_slic(const float *rgb, int rgb_layout, int H, int W, int n_segments, double compactness,
                          int max_iter, int enforce_connectivity, int32_t *labels, int32_t *n_labels, void *ws,
                          void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    WESUP_REQUIRE(rgb && labels && n_labels && ws, WESUP_E_ARG, "wesup_slic: null pointer");
    WESUP_REQUIRE(H > 0 && W > 0 && n_segments > 0 && compactness > 0 && max_iter >= 0, WESUP_E_ARG,
                  "wesup_slic: bad argument H=%d W=%d n_segments=%d compactness=%g", H, W, n_segments, compactness);
    WESUP_REQUIRE(rgb_layout == WESUP_CHW || rgb_layout == WESUP_HWC, WESUP_E_ARG, "wesup_slic: bad layout %d", rgb_layout);
    WESUP_REQUIRE((long)H * W < (1L << 31), WESUP_E_UNSUPPORTED, "wesup_slic: H*W must fit int32");
    WESUP_REQUIRE(H < 65536 && W < 65536, WESUP_E_UNSUPPORTED, "wesup_slic: H and W must be below 65536");
    int step, start, ny, nx;
    long K = slic_grid(H, W, n_segments, &step, &start, &ny, &nx);
    WESUP_REQUIRE(K > 0, WESUP_E_UNSUPPORTED, "wesup_slic: degenerate seed grid for %dx%d / %d segments", H, W, n_segments);
    const long HW = (long)H * W;
    int cells_y, cells_x;
    const long cells = n_cells(H, W, step, &cells_y, &cells_x);
    SlicWs s = carve_slic(ws, HW, K, cells);
    s.cells_y = cells_y; s.cells_x = cells_x;
    const int nb = cdiv(HW, 256);
    cudaError_t me = cudaMemsetAsync(s.acc_n, 0, s.zero_bytes, stream);
    WESUP_REQUIRE(me == cudaSuccess, (int)me, "wesup_slic: memset: %s", cudaGetErrorString(me));
    slic_lab_kernel<<<nb, 256, 0, stream>>>(rgb, rgb_layout, HW, 1.0 / compactness, s.lab);
    slic_init_kernel<<<cdiv(HW > K ? HW : K, 256), 256, 0, stream>>>(s, K, nx, step, start, HW);
    float stepf = (float)step;
    double spatial_weight = 1.0 / (double)(stepf * stepf);
    dim3 tiles(cdiv(W, AT), cdiv(H, AT));
    for (int it = 0; it < max_iter; ++it)
        slic_sweep_kernel<<<tiles, 256, 0, stream>>>(s, H, W, K, step, spatial_weight);
    if (!enforce_connectivity) {
        copy_labels_kernel<<<nb, 256, 0, stream>>>(s.nearest, labels, HW, n_labels, (int)K);
        WESUP_CHECK_LAUNCH("wesup_slic", 3 + max_iter);
        return 0;
    }
    double segment_size = (double)HW / (double)n_segments;
    int min_size = (int)(0.5 * segment_size);
    ccl_merge_kernel<<<nb, 256, 0, stream>>>(s, H, W);
    ccl_flatten_kernel<<<nb, 256, 0, stream>>>(s, HW);
    ccl_flags_kernel<<<nb, 256, 0, stream>>>(s, HW, min_size);
    device_exclusive_scan(s.keep_scan, HW, s.block_sums, n_labels, stream);
    device_exclusive_scan(s.small_scan, HW, s.block_sums, nullptr, stream);
    ccl_small_adjacent_kernel<<<nb, 256, 0, stream>>>(s, H, W, min_size);
    ccl_relabel_kernel<<<nb, 256, 0, stream>>>(s, HW, min_size, labels);
    WESUP_CHECK_LAUNCH("wesup_slic", 2 + max_iter + 11);
    return 0;
}
