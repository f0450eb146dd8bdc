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

struct SlicWs {
    double *lab;        // 3 planes, H*W each (already scaled by 1/compactness)
    double *cent;       // K*5: y,x,L,a,b
    u64 *acc_n;         // K*3: count, sum y, sum x  (exact integer sums)
    double *acc_c;      // K*3: sum L,a,b
    int32_t *nearest;   // H*W raw k-means assignment
    int32_t *parent;    // H*W union-find / component root (min pixel id)
    int32_t *size;      // H*W component size at its root
    int32_t *keep_scan; // H*W exclusive scan of "kept root" flags
    int32_t *small_scan;// H*W exclusive scan of small-component sizes at roots
    int32_t *queue;     // H*W BFS scratch for small components
    int32_t *adj_root;  // H*W (at roots) root of the adjacent component or -1
    int32_t *block_sums;// scan scratch
    uint8_t *seen;      // H*W
};

static inline size_t up256(size_t x) { return (x + 255) / 256 * 256; }
constexpr int SCAN_ELEMS = 4096;    // per block (1024 threads x 4)

static size_t slic_ws_bytes(long HW, long K) {
    long nblk = (HW + SCAN_ELEMS - 1) / SCAN_ELEMS + 1;
    return up256(sizeof(double) * 3 * HW) + up256(sizeof(double) * 5 * K) + up256(sizeof(u64) * 3 * K) +
           up256(sizeof(double) * 3 * K) + 7 * up256(sizeof(int32_t) * HW) + up256(sizeof(int32_t) * 2 * nblk) +
           up256(HW);
}

static SlicWs carve_slic(void *ws, long HW, long K) {
    char *p = static_cast<char *>(ws);
    SlicWs s;
    long nblk = (HW + SCAN_ELEMS - 1) / SCAN_ELEMS + 1;
    s.lab = (double *)p;        p += up256(sizeof(double) * 3 * HW);
    s.cent = (double *)p;       p += up256(sizeof(double) * 5 * K);
    s.acc_n = (u64 *)p;         p += up256(sizeof(u64) * 3 * K);
    s.acc_c = (double *)p;      p += up256(sizeof(double) * 3 * K);
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

__global__ void slic_init_kernel(SlicWs s, long K, int nx, int step, int start, long HW) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < K) {
        s.cent[5 * i + 0] = (double)(start + (int)(i / nx) * step);
        s.cent[5 * i + 1] = (double)(start + (int)(i % nx) * step);
        s.cent[5 * i + 2] = 0.0; s.cent[5 * i + 3] = 0.0; s.cent[5 * i + 4] = 0.0;
        s.acc_n[3 * i] = 0; s.acc_n[3 * i + 1] = 0; s.acc_n[3 * i + 2] = 0;
        s.acc_c[3 * i] = 0.0; s.acc_c[3 * i + 1] = 0.0; s.acc_c[3 * i + 2] = 0.0;
    }
    if (i < HW) s.nearest[i] = 0;
}

// ---------------------------------------------------------------------------
// assignment + accumulation of the new cluster sums
// ---------------------------------------------------------------------------
constexpr int AT = 16;              // tile edge
constexpr int ACHUNK = 256;         // centres examined per round (= block size)

struct Cand {
    double cy, cx, l, a, b;
    int k, y0, y1, x0, x1;
};

__global__ void __launch_bounds__(256) slic_assign_kernel(SlicWs s, int H, int W, long K, int step, double spatial_weight) {
    __shared__ Cand cand[ACHUNK];
    __shared__ int n_cand;
    const int tx0 = blockIdx.x * AT, ty0 = blockIdx.y * AT;
    const int x = tx0 + (threadIdx.x & (AT - 1)), y = ty0 + (threadIdx.x >> 4);
    const bool live = x < W && y < H;
    const long HW = (long)H * W;
    const long p = (long)y * W + x;
    double pl = 0, pa = 0, pb = 0;
    if (live) { pl = s.lab[p]; pa = s.lab[HW + p]; pb = s.lab[2 * HW + p]; }
    double best = CUDART_INF;
    int best_k = -1;
    const double two_s = (double)(2 * step);
    for (long base = 0; base < K; base += ACHUNK) {
        __syncthreads();
        if (threadIdx.x == 0) n_cand = 0;
        __syncthreads();
        long k = base + threadIdx.x;
        if (k < K) {
            double cy = s.cent[5 * k], cx = s.cent[5 * k + 1];
            if (cy == cy && cx == cx) {          // NaN centres (empty clusters) never win
                double lo;
                lo = cy - two_s;       int y0 = (int)(lo > 0.0 ? lo : 0.0);
                lo = cy + two_s + 1.0; int y1 = (int)(lo < (double)H ? lo : (double)H);
                lo = cx - two_s;       int x0 = (int)(lo > 0.0 ? lo : 0.0);
                lo = cx + two_s + 1.0; int x1 = (int)(lo < (double)W ? lo : (double)W);
                if (y0 < ty0 + AT && y1 > ty0 && x0 < tx0 + AT && x1 > tx0) {
                    int slot = atomicAdd(&n_cand, 1);
                    Cand c;
                    c.cy = cy; c.cx = cx; c.l = s.cent[5 * k + 2]; c.a = s.cent[5 * k + 3]; c.b = s.cent[5 * k + 4];
                    c.k = (int)k; c.y0 = y0; c.y1 = y1; c.x0 = x0; c.x1 = x1;
                    cand[slot] = c;
                }
            }
        }
        __syncthreads();
        if (live) {
            const int n = n_cand;
            for (int i = 0; i < n; ++i) {
                const Cand &c = cand[i];
                if (y < c.y0 || y >= c.y1 || x < c.x0 || x >= c.x1) continue;
                double dy = __dadd_rn(c.cy, -(double)y); dy = __dmul_rn(dy, dy);
                double dx = __dadd_rn(c.cx, -(double)x); dx = __dmul_rn(dx, dx);
                double d = __dmul_rn(__dadd_rn(dy, dx), spatial_weight);
                double t = __dadd_rn(pl, -c.l);
                double dc = __dmul_rn(t, t);                       // 0 + t*t
                t = __dadd_rn(pa, -c.a); dc = __dadd_rn(dc, __dmul_rn(t, t));
                t = __dadd_rn(pb, -c.b); dc = __dadd_rn(dc, __dmul_rn(t, t));
                d = __dadd_rn(d, dc);
                if (d < best || (d == best && c.k < best_k)) { best = d; best_k = c.k; }
            }
        }
    }
    if (live) {
        int k = best_k;
        if (k >= 0) s.nearest[p] = k; else k = s.nearest[p];      // uncovered pixel keeps its previous cluster
        atomicAdd(&s.acc_n[3 * k], 1ull);
        atomicAdd(&s.acc_n[3 * k + 1], (u64)y);
        atomicAdd(&s.acc_n[3 * k + 2], (u64)x);
        atomicAdd(&s.acc_c[3 * k], pl);
        atomicAdd(&s.acc_c[3 * k + 1], pa);
        atomicAdd(&s.acc_c[3 * k + 2], pb);
    }
}

__global__ void slic_update_kernel(SlicWs s, long K) {
    long k = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    double n = (double)s.acc_n[3 * k];
    // 0/0 -> NaN for an empty cluster, as in the sequential algorithm
    s.cent[5 * k + 0] = __ddiv_rn((double)s.acc_n[3 * k + 1], n);
    s.cent[5 * k + 1] = __ddiv_rn((double)s.acc_n[3 * k + 2], n);
    s.cent[5 * k + 2] = __ddiv_rn(s.acc_c[3 * k], n);
    s.cent[5 * k + 3] = __ddiv_rn(s.acc_c[3 * k + 1], n);
    s.cent[5 * k + 4] = __ddiv_rn(s.acc_c[3 * k + 2], n);
    s.acc_n[3 * k] = 0; s.acc_n[3 * k + 1] = 0; s.acc_n[3 * k + 2] = 0;
    s.acc_c[3 * k] = 0.0; s.acc_c[3 * k + 1] = 0.0; s.acc_c[3 * k + 2] = 0.0;
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

__global__ void ccl_init_kernel(SlicWs s, long HW) {
    long p = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < HW) { s.parent[p] = (int)p; s.size[p] = 0; s.seen[p] = 0; s.adj_root[p] = -1; }
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
    return slic_ws_bytes((long)H * W, K);
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
    int step, start, ny, nx;
    long K = slic_grid(H, W, n_segments, &step, &start, &ny, &nx);
    WESUP_REQUIRE(K > 0, WESUP_E_UNSUPPORTED, "wesup_slic: degenerate seed grid for %dx%d / %d segments", H, W, n_segments);
    const long HW = (long)H * W;
    SlicWs s = carve_slic(ws, HW, K);
    const int nb = cdiv(HW, 256);
    slic_lab_kernel<<<nb, 256, 0, stream>>>(rgb, rgb_layout, HW, 1.0 / compactness, s.lab);
    slic_init_kernel<<<cdiv(HW > K ? HW : K, 256), 256, 0, stream>>>(s, K, nx, step, start, HW);
    float stepf = (float)step;
    double spatial_weight = 1.0 / (double)(stepf * stepf);
    dim3 tiles(cdiv(W, AT), cdiv(H, AT));
    for (int it = 0; it < max_iter; ++it) {
        slic_assign_kernel<<<tiles, 256, 0, stream>>>(s, H, W, K, step, spatial_weight);
        slic_update_kernel<<<cdiv(K, 256), 256, 0, stream>>>(s, K);
    }
    if (!enforce_connectivity) {
        copy_labels_kernel<<<nb, 256, 0, stream>>>(s.nearest, labels, HW, n_labels, (int)K);
        WESUP_CHECK_LAUNCH("wesup_slic", 3 + 2 * max_iter);
        return 0;
    }
    double segment_size = (double)HW / (double)n_segments;
    int min_size = (int)(0.5 * segment_size);
    ccl_init_kernel<<<nb, 256, 0, stream>>>(s, HW);
    ccl_merge_kernel<<<nb, 256, 0, stream>>>(s, H, W);
    ccl_flatten_kernel<<<nb, 256, 0, stream>>>(s, HW);
    ccl_flags_kernel<<<nb, 256, 0, stream>>>(s, HW, min_size);
    device_exclusive_scan(s.keep_scan, HW, s.block_sums, n_labels, stream);
    device_exclusive_scan(s.small_scan, HW, s.block_sums, nullptr, stream);
    ccl_small_adjacent_kernel<<<nb, 256, 0, stream>>>(s, H, W, min_size);
    ccl_relabel_kernel<<<nb, 256, 0, stream>>>(s, HW, min_size, labels);
    WESUP_CHECK_LAUNCH("wesup_slic", 2 + 2 * max_iter + 12);
    return 0;
}
