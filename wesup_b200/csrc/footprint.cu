// Superpixel pooling over feature levels with PRECOMPUTED footprints.
//
// Reference semantics: bilinear(align_corners=True) upsampling + channel concat
// (/root/reference/models/wesup.py:254-261) followed by torch.mm(sp_maps, x.t()) (:284-285).
// Both are linear, so for every level l and superpixel k
//
//     pooled[k, coff_l + c] = (1/|S_k|) * sum_q  G_k(q) * level_l[q, c]
//     G_k(q) = sum_{(y,x) in S_k} wy(y, q.i) * wx(x, q.j)          (bilinear tap weights)
//
// G depends only on the label map and the level RESOLUTION -- it is the sparse form of the
// reference's dense `sp_maps` (:57-61) composed with the interpolation matrix, and like
// `sp_maps` it is built once per image in preprocessing (`wesup_footprint_build`):
//   * forward lists: per superpixel and resolution, the low-resolution cells it touches with
//     their aggregated weights;
//   * backward lists (the transpose): per low-resolution cell, the superpixels whose pixels
//     tap it, weight already divided by |S_k|, sorted by superpixel row.
// The pooling kernels are then pure streaming gathers with no prologue, no shared-memory
// tables, no barriers on the load path and small register footprints:
//   fwd: pooled[k, chunk]    = 1/|S_k| * sum_e w_e * level[cell_e, chunk]
//   bwd: grad_level[q, slice] = sum_e w_e * grad_pooled[k_e, slice]
// (the in-kernel variant that rebuilds G on every call lives in pool_levels.cu).
//
// Determinism: weights are sums of 32-bit fixed-point tap products (integer adds commute);
// every floating-point sum runs in a fixed order (entry lists are ordered by cell index /
// superpixel row).  WHERE a list is placed inside the entry arrays is decided by an atomic
// cursor and may change run to run; its content does not.
#include "common.cuh"
#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

namespace wesup {

constexpr int FP_MAX_RES = 6;        // distinct non-identity resolutions (VGG16: 4)
constexpr int FP_THREADS = 256;
constexpr int FP_GRID_CAP = 1024;    // cells of a superpixel's low-res weight grids kept in shared memory: the minimum; the
                                     // launch sizes the (dynamic) grids from the expected superpixel extent, up to
constexpr int FP_GRID_CAP_MAX = 16384;   // ... 192 KB.  (r2 config-3 sweep: at 2048^2 / N = 500 a superpixel's boxes hold
                                     // ~3400 cells, nothing fitted 1024, the per-pixel fallback lists were 4x longer and
                                     // scattered, and the forward pooling ran at 0.09 of the HBM rate)
constexpr int FP_SLOTS = 64;         // distinct superpixels per cell footprint on the fast path
constexpr int FP_WARPS = FP_THREADS / 32;

struct FpEnt { int idx; float w; };  // fwd: low-res cell index; bwd: superpixel row (w pre-divided by |S_k|)

struct FpRes {
    int h, w;
    float sy, sx;
    float fscale, finv;               // 2^F and 2^-F of the fixed-point weight sums (0: per-pixel taps instead)
    int2 *fwd_span;                   // (N)    {first entry, number of entries} of superpixel k
    FpEnt *fwd_ent;                   // (cap)
    int2 *bwd_span;                   // (h*w)  same per low-resolution cell
    FpEnt *bwd_ent;                   // (cap)
    int *cursor;                      // [0] fwd entries handed out, [1] bwd entries handed out
    // per-axis footprint tables (bwd builder): first output index, number of output indices and their tap
    // weights for every low-resolution row / column
    int32_t *ylo, *yn, *xlo, *xn;
    float *wy, *wx;
    int ky, kx;
    int blk1;                         // bwd builder: one past the last block of this resolution (coarse first)
};

struct FpPlan {
    int n, H, W, N;
    FpRes r[FP_MAX_RES];
    int *cursors;
    size_t cursor_bytes;
};

static inline size_t up256(size_t x) { return (x + 255) / 256 * 256; }
static inline int table_len(float scale, int out_size) {
    if (!(scale > 0.f)) return out_size;
    const int k = (int)(2.0f / scale) + 7;
    return k < out_size + 2 ? k : out_size + 2;
}

// Lay the footprint blob out (base == nullptr: size only).  Every pixel contributes at most four taps per
// resolution, so 4*H*W entries bound both the forward and the backward lists whatever the label map is.
static long plan_footprint(FpPlan &P, const int *h, const int *w, int n_levels, int H, int W, int N, char *base) {
    P.n = 0; P.H = H; P.W = W; P.N = N;
    size_t off = 0;
    auto take = [&](size_t bytes) { char *p = base ? base + off : nullptr; off += up256(bytes); return p; };
    P.cursor_bytes = up256(sizeof(int) * 2 * FP_MAX_RES);
    P.cursors = (int *)take(P.cursor_bytes);
    const size_t cap = 4 * (size_t)H * (size_t)W;
    for (int l = 0; l < n_levels; ++l) {
        if (h[l] == H && w[l] == W) continue;
        bool seen = false;
        for (int i = 0; i < P.n; ++i) seen = seen || (P.r[i].h == h[l] && P.r[i].w == w[l]);
        if (seen) continue;
        if (P.n == FP_MAX_RES) return -1;
        FpRes &r = P.r[P.n];
        r.h = h[l]; r.w = w[l];
        r.sy = bilinear_scale(h[l], H); r.sx = bilinear_scale(w[l], W);
        // fixed-point format of a weight sum: a cell's sum is at most its footprint area
        const double fy = r.sy > 0.f ? 2.0 / r.sy + 2.0 : (double)H;
        const double fx = r.sx > 0.f ? 2.0 / r.sx + 2.0 : (double)W;
        const int bits = 31 - (int)ceil(log2(fy * fx + 1.0));
        r.fscale = bits >= 16 ? (float)ldexp(1.0, bits) : 0.f;
        r.finv = bits >= 16 ? (float)ldexp(1.0, -bits) : 0.f;
        r.cursor = P.cursors ? P.cursors + 2 * P.n : nullptr;
        r.fwd_span = (int2 *)take(sizeof(int2) * (size_t)N);
        r.fwd_ent = (FpEnt *)take(sizeof(FpEnt) * cap);
        r.bwd_span = (int2 *)take(sizeof(int2) * (size_t)r.h * r.w);
        r.bwd_ent = (FpEnt *)take(sizeof(FpEnt) * cap);
        r.ky = table_len(r.sy, H); r.kx = table_len(r.sx, W);
        r.ylo = (int32_t *)take(sizeof(int32_t) * r.h);
        r.yn = (int32_t *)take(sizeof(int32_t) * r.h);
        r.wy = (float *)take(sizeof(float) * (size_t)r.h * r.ky);
        r.xlo = (int32_t *)take(sizeof(int32_t) * r.w);
        r.xn = (int32_t *)take(sizeof(int32_t) * r.w);
        r.wx = (float *)take(sizeof(float) * (size_t)r.w * r.kx);
        r.blk1 = 0;
        ++P.n;
    }
    return (long)(off > 256 ? off : 256);
}

// ---------------------------------------------------------------------------
// forward lists: one block per superpixel
//
// Every pixel adds its 2x2 tap weights wy*wx to the cells of the superpixel's low-res bounding
// boxes (one per resolution, side by side in shared memory) as 32-bit FIXED-POINT sums --
// native shared-memory integer atomics.  The number of fractional bits is chosen per
// resolution so that the largest possible cell sum (the cell's footprint area) cannot
// overflow: 27 bits at stride 2 ... 21 bits at stride 16 (fp32 itself keeps 24).  The non-zero
// cells are compacted in cell order and written out as (cell index, weight).  A superpixel
// whose bounding boxes do not fit the shared grids (huge / scattered) emits its per-pixel
// taps instead: four entries per pixel, in pixel order.
// ---------------------------------------------------------------------------
struct RInfo { int i_lo, j_lo, gw, cells, goff, lbeg, lend, base; };

__global__ void __launch_bounds__(FP_THREADS) fp_build_fwd_kernel(const FpPlan P, const int32_t *__restrict__ seg_offsets,
                                                                  const int32_t *__restrict__ seg_pixels, int grid_cap) {
    extern __shared__ __align__(16) unsigned char fp_dyn_smem[];
    unsigned *grid = reinterpret_cast<unsigned *>(fp_dyn_smem);                    // grid_cap fixed-point weight sums
    FpEnt *cellw = reinterpret_cast<FpEnt *>(fp_dyn_smem + (size_t)grid_cap * sizeof(unsigned));   // grid_cap compacted entries
    __shared__ RInfo ri[FP_MAX_RES];
    __shared__ int bbx[2];
    __shared__ int warp_cnt[FP_WARPS];
    __shared__ int total_cells;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int k = blockIdx.x;
    const int beg = __ldg(seg_offsets + k), end = __ldg(seg_offsets + k + 1);
    const int n = end - beg;
    if (n <= 0) {
        if (tid < P.n) P.r[tid].fwd_span[k] = make_int2(0, 0);
        return;
    }
    const int W = P.W;
    // bounding box: pixel ids ascend inside a row of the CSR, so rows come from the two ends
    if (tid == 0) { bbx[0] = INT_MAX; bbx[1] = -1; }
    __syncthreads();
    {
        int xmin = INT_MAX, xmax = -1;
        for (int i = beg + tid; i < end; i += FP_THREADS) {
            const int p = __ldg(seg_pixels + i);
            const int x = p - (p / W) * W;
            xmin = min(xmin, x); xmax = max(xmax, x);
        }
        xmin = __reduce_min_sync(0xffffffffu, xmin);
        xmax = __reduce_max_sync(0xffffffffu, xmax);
        if (lane == 0) { atomicMin(&bbx[0], xmin); atomicMax(&bbx[1], xmax); }
    }
    __syncthreads();
    if (tid == 0) {
        const int ymin = __ldg(seg_pixels + beg) / W, ymax = __ldg(seg_pixels + end - 1) / W;
        const int xmin = bbx[0], xmax = bbx[1];
        int off = 0;
        for (int r = 0; r < P.n; ++r) {
            const FpRes &R = P.r[r];
            RInfo q;
            q.i_lo = q.j_lo = 0; q.gw = 1; q.cells = 0; q.goff = off; q.lbeg = q.lend = 0; q.base = 0;
            if (R.fscale > 0.f) {
                q.i_lo = bilinear_tap(ymin, R.sy, R.h).i0;
                q.j_lo = bilinear_tap(xmin, R.sx, R.w).i0;
                const long gh = bilinear_tap(ymax, R.sy, R.h).i1 - q.i_lo + 1;
                q.gw = bilinear_tap(xmax, R.sx, R.w).i1 - q.j_lo + 1;
                if (off + gh * q.gw <= grid_cap) { q.cells = (int)gh * q.gw; off += q.cells; }
            }
            ri[r] = q;
        }
        total_cells = off;
    }
    __syncthreads();
    const int total = total_cells;
    for (int e = tid; e < total; e += FP_THREADS) grid[e] = 0u;
    __syncthreads();
    for (int it = beg + tid; it < end; it += FP_THREADS) {
        const int p = __ldg(seg_pixels + it);
        const int y = p / W, x = p - y * W;
        for (int r = 0; r < P.n; ++r) {
            const RInfo q = ri[r];
            if (q.cells == 0) continue;
            const FpRes &R = P.r[r];
            const Tap ty = bilinear_tap(y, R.sy, R.h), tx = bilinear_tap(x, R.sx, R.w);
            unsigned *gp = grid + q.goff;
            const int r0 = (ty.i0 - q.i_lo) * q.gw, r1 = (ty.i1 - q.i_lo) * q.gw, c0 = tx.i0 - q.j_lo, c1 = tx.i1 - q.j_lo;
            const float fs = R.fscale;
            const float a0 = ty.w0 * fs, a1 = ty.w1 * fs;            // exact: fs is a power of two
            atomicAdd(gp + r0 + c0, __float2uint_rn(a0 * tx.w0));
            atomicAdd(gp + r0 + c1, __float2uint_rn(a0 * tx.w1));
            atomicAdd(gp + r1 + c0, __float2uint_rn(a1 * tx.w0));
            atomicAdd(gp + r1 + c1, __float2uint_rn(a1 * tx.w1));
        }
    }
    __syncthreads();
    // compact the non-zero cells, resolution after resolution (order: cell index -- fixed)
    int n_list = 0;
    for (int e0 = 0; e0 < total; e0 += FP_THREADS) {
        const int e = e0 + tid;
        const unsigned v = e < total ? grid[e] : 0u;
        const unsigned m = __ballot_sync(0xffffffffu, v != 0u);
        if (lane == 0) warp_cnt[wid] = __popc(m);
        __syncthreads();
        int before = n_list, tot = n_list;
#pragma unroll
        for (int q = 0; q < FP_WARPS; ++q) {
            const int c = warp_cnt[q];
            if (q < wid) before += c;
            tot += c;
        }
        if (e < total) {
            int r = 0;
            while (e >= ri[r].goff + ri[r].cells) ++r;
            const int pos = before + __popc(m & ((1u << lane) - 1u));
            const int local = e - ri[r].goff, gw = ri[r].gw;
            if (local == 0) ri[r].lbeg = pos;
            if (v != 0u) {
                const int i = __float2int_rd(((float)local + 0.5f) / (float)gw), j = local - i * gw;
                FpEnt cw;
                cw.idx = (ri[r].i_lo + i) * P.r[r].w + ri[r].j_lo + j;
                cw.w = (float)v * P.r[r].finv;
                cellw[pos] = cw;
            }
        }
        n_list = tot;
        __syncthreads();
    }
    if (tid == 0) {
        int nxt = n_list;
        for (int r = P.n - 1; r >= 0; --r)
            if (ri[r].cells > 0) { ri[r].lend = nxt; nxt = ri[r].lbeg; }
        for (int r = 0; r < P.n; ++r) {
            const int cnt = ri[r].cells > 0 ? ri[r].lend - ri[r].lbeg : 4 * n;
            const int base = atomicAdd(P.r[r].cursor, cnt);
            ri[r].base = base;
            P.r[r].fwd_span[k] = make_int2(base, cnt);
        }
    }
    __syncthreads();
    for (int e = tid; e < n_list; e += FP_THREADS) {
        int r = 0;
        while (ri[r].cells == 0 || e >= ri[r].lend) ++r;
        P.r[r].fwd_ent[ri[r].base + e - ri[r].lbeg] = cellw[e];
    }
    for (int r = 0; r < P.n; ++r) {
        if (ri[r].cells > 0) continue;
        const FpRes &R = P.r[r];
        FpEnt *dst = R.fwd_ent + ri[r].base;
        for (int it = beg + tid; it < end; it += FP_THREADS) {
            const int p = __ldg(seg_pixels + it);
            const int y = p / W, x = p - y * W;
            const Tap ty = bilinear_tap(y, R.sy, R.h), tx = bilinear_tap(x, R.sx, R.w);
            FpEnt *d = dst + 4 * (long)(it - beg);
            d[0].idx = ty.i0 * R.w + tx.i0; d[0].w = ty.w0 * tx.w0;
            d[1].idx = ty.i0 * R.w + tx.i1; d[1].w = ty.w0 * tx.w1;
            d[2].idx = ty.i1 * R.w + tx.i0; d[2].w = ty.w1 * tx.w0;
            d[3].idx = ty.i1 * R.w + tx.i1; d[3].w = ty.w1 * tx.w1;
        }
    }
}

// ---------------------------------------------------------------------------
// backward lists: one warp per low-resolution cell
//
// The warp reads the labels of the cell's high-resolution footprint ONCE and folds the tap
// weights into a 64-slot hash table keyed by superpixel (per warp, shared memory): the key is
// claimed with an atomicCAS, the weight is added as a 32-bit fixed-point integer --
// commutative, so the table's CONTENT does not depend on thread order.  The occupied slots
// are ranked by key, which gives a list sorted by superpixel row.  A footprint that meets
// more than 64 superpixels takes the slow path: distinct labels visited in ascending order,
// the footprint re-read for each.
// ---------------------------------------------------------------------------
__device__ __forceinline__ float tap_weight(int dst, int cell, float scale, int in_size) {
    const Tap t = bilinear_tap(dst, scale, in_size);
    return (t.i0 == cell ? t.w0 : 0.f) + (t.i1 == cell ? t.w1 : 0.f);
}

// blockIdx.y = 2 * resolution + axis; one thread per low-resolution index: the exact support of the cell
// among the output indices (bracket from the inverse map, trimmed with the forward's own tap arithmetic)
// and the tap weights inside it
__global__ void fp_axis_tables_kernel(const FpPlan P) {
    const int r = blockIdx.y >> 1, axis = blockIdx.y & 1;
    const FpRes &R = P.r[r];
    const int in_size = axis ? R.w : R.h, out_size = axis ? P.W : P.H;
    const float scale = axis ? R.sx : R.sy;
    const int K = axis ? R.kx : R.ky;
    int32_t *lo_t = axis ? R.xlo : R.ylo, *n_t = axis ? R.xn : R.yn;
    float *w_t = axis ? R.wx : R.wy;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= in_size) return;
    int lo = 0, hi = out_size - 1;
    if (scale > 0.f) {
        const float inv = 1.0f / scale;
        lo = max((int)floorf((float)(i - 1) * inv) - 1, 0);
        hi = min((int)ceilf((float)(i + 1) * inv) + 1, out_size - 1);
    }
    while (lo < hi && tap_weight(lo, i, scale, in_size) == 0.f) ++lo;
    while (hi > lo && tap_weight(hi, i, scale, in_size) == 0.f) --hi;
    int n = hi - lo + 1;
    if (n > K) n = K;                 // cannot happen: K bounds the bracket
    lo_t[i] = lo;
    n_t[i] = n;
    for (int k = 0; k < n; ++k) w_t[(long)i * K + k] = tap_weight(lo + k, i, scale, in_size);
}

struct Staged { int lab; float w; };

__global__ void __launch_bounds__(FP_THREADS) fp_build_bwd_kernel(const FpPlan P, const int32_t *__restrict__ row_labels,
                                                                  const int32_t *__restrict__ counts) {
    __shared__ int keys_all[FP_WARPS][FP_SLOTS];
    __shared__ unsigned wfix_all[FP_WARPS][FP_SLOTS];
    __shared__ Staged list_all[FP_WARPS][FP_SLOTS];
    __shared__ int s_nl[FP_WARPS];
    __shared__ int s_base;
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int r = P.n - 1;                                    // coarse resolutions (longest footprint walks) own the first blocks
    while (r > 0 && (int)blockIdx.x >= P.r[r].blk1) --r;
    const FpRes &R = P.r[r];
    const int blk = (int)blockIdx.x - (r == P.n - 1 ? 0 : P.r[r + 1].blk1);
    const long q = (long)blk * FP_WARPS + wid;
    const bool active = q < (long)R.h * R.w;            // no early exit: the block allocates its lists together
    const int W = P.W;
    int *keys = keys_all[wid];
    unsigned *wfix = wfix_all[wid];
    Staged *list = list_all[wid];
    const int i = active ? (int)(q / R.w) : 0, j = active ? (int)(q - (long)i * R.w) : 0;
    const int ylo = __ldg(R.ylo + i), xlo = __ldg(R.xlo + j);
    const int nx = __ldg(R.xn + j), nf = active ? __ldg(R.yn + i) * nx : 0;
    const float *__restrict__ wyt = R.wy + (long)i * R.ky;
    const float *__restrict__ wxt = R.wx + (long)j * R.kx;
    const float inv_nx = 1.0f / (float)nx;
    auto fetch = [&](int t) {
        const int a = __float2int_rd(((float)t + 0.5f) * inv_nx), b = t - a * nx;
        Staged s;
        s.w = __ldg(wyt + a) * __ldg(wxt + b);
        s.lab = (s.w != 0.f) ? __ldg(row_labels + (long)(ylo + a) * W + xlo + b) : -1;
        return s;
    };
    auto emit = [&](FpEnt *dst, int lab, float w) {
        const int cnt = __ldg(counts + lab);
        dst->idx = lab;
        dst->w = cnt > 0 ? w / (float)cnt : 0.f;
    };
    // ---- phase 1: the list (fast path) or its length (slow path) -----------------------------------------
    keys[lane] = -1; keys[lane + 32] = -1;
    wfix[lane] = 0u; wfix[lane + 32] = 0u;
    __syncwarp();
    const float fs = R.fscale;
    bool overflow = !(fs > 0.f);
    if (!overflow) {
        auto insert = [&](const Staged &s_) {
            if (s_.lab < 0) return;
            unsigned h = ((unsigned)s_.lab * 2654435761u) >> 26;
            int probes = 0;
            for (; probes < FP_SLOTS; ++probes) {
                const int old = atomicCAS(&keys[h], -1, s_.lab);
                if (old == -1 || old == s_.lab) { atomicAdd(&wfix[h], __float2uint_rn(s_.w * fs)); break; }
                h = (h + 1) & (FP_SLOTS - 1);
            }
            if (probes == FP_SLOTS) overflow = true;
        };
        int t = lane;
        for (; t + 96 < nf; t += 128) {                  // four label loads in flight per lane
            Staged b4[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) b4[u] = fetch(t + 32 * u);
#pragma unroll
            for (int u = 0; u < 4; ++u) insert(b4[u]);
        }
        for (; t < nf; t += 32) insert(fetch(t));
    }
    overflow = __any_sync(0xffffffffu, overflow);
    __syncwarp();
    int nl = 0, first = INT_MAX;
    if (!overflow) {
        // rank the occupied slots by key -> list sorted by superpixel row
        const int k0 = keys[lane], k1 = keys[lane + 32];
        const unsigned occ0 = __ballot_sync(0xffffffffu, k0 >= 0), occ1 = __ballot_sync(0xffffffffu, k1 >= 0);
        int r0 = 0, r1 = 0;
        for (unsigned m = occ0; m; m &= m - 1) {
            const int kk = keys[__ffs(m) - 1];
            r0 += kk < k0; r1 += kk < k1;
        }
        for (unsigned m = occ1; m; m &= m - 1) {
            const int kk = keys[32 + __ffs(m) - 1];
            r0 += kk < k0; r1 += kk < k1;
        }
        const float finv = R.finv;
        if (k0 >= 0) { list[r0].lab = k0; list[r0].w = (float)wfix[lane] * finv; }
        if (k1 >= 0) { list[r1].lab = k1; list[r1].w = (float)wfix[lane + 32] * finv; }
        nl = __popc(occ0) + __popc(occ1);
    } else {
        // count the distinct labels (they are then visited in ascending order, the footprint re-read for each)
        for (int t = lane; t < nf; t += 32) {
            const Staged s_ = fetch(t);
            if (s_.lab >= 0) first = min(first, s_.lab);
        }
        first = __reduce_min_sync(0xffffffffu, first);
        for (int cur = first; cur != INT_MAX; ++nl) {
            int nxt = INT_MAX;
            for (int t = lane; t < nf; t += 32) {
                const Staged s_ = fetch(t);
                if (s_.lab > cur) nxt = min(nxt, s_.lab);
            }
            cur = __reduce_min_sync(0xffffffffu, nxt);
        }
    }
    // ---- phase 2: ONE cursor atomic per block (all its cells belong to one resolution) ------------------------
    if (lane == 0) s_nl[wid] = nl;
    __syncthreads();
    if (threadIdx.x == 0) {
        int tot = 0;
#pragma unroll
        for (int w_ = 0; w_ < FP_WARPS; ++w_) tot += s_nl[w_];
        s_base = tot > 0 ? atomicAdd(R.cursor + 1, tot) : 0;
    }
    __syncthreads();
    if (!active) return;
    int base = s_base;
    for (int w_ = 0; w_ < wid; ++w_) base += s_nl[w_];
    if (lane == 0) R.bwd_span[q] = make_int2(base, nl);
    FpEnt *dst = R.bwd_ent + base;
    // ---- phase 3: write the list ------------------------------------------------------------------------------
    if (!overflow) {
        for (int e = lane; e < nl; e += 32) emit(dst + e, list[e].lab, list[e].w);
    } else {
        int e = 0;
        for (int cur = first; cur != INT_MAX; ++e) {
            float ws = 0.f;
            int nxt = INT_MAX;
            for (int t = lane; t < nf; t += 32) {
                const Staged s_ = fetch(t);
                if (s_.lab == cur) ws += s_.w;
                else if (s_.lab > cur) nxt = min(nxt, s_.lab);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) ws += __shfl_xor_sync(0xffffffffu, ws, o);
            if (lane == 0) emit(dst + e, cur, ws);
            cur = __reduce_min_sync(0xffffffffu, nxt);
        }
    }
}

// ---------------------------------------------------------------------------
// pooling kernels over the lists
// ---------------------------------------------------------------------------
// consecutive levels of equal resolution: their channels are contiguous in the pooled row
struct PoolGroup {
    int res;            // index into FpPlan::r, -1 for full-resolution (identity) levels
    int l0, l1;         // levels [l0, l1)
    int coff, Cg;       // channel offset / channels in the pooled row
    int nchunk;         // fwd: 128-channel chunks; bwd: 256-channel slices
    int split;          // fwd: warps sharing one (superpixel, chunk) unit
    int blk0;           // first block of the group inside the launch
};
struct PoolPlan { int n; PoolGroup g[WESUP_MAX_LEVELS]; };

__device__ __forceinline__ void locate_level(const Levels &L, int l0, int l1, int c, int &l, int &cl) {
    l = l0;
    while (l + 1 < l1 && c >= L.C[l]) { c -= L.C[l]; ++l; }
    cl = c;
}

__device__ __forceinline__ FpEnt ld_ent(const FpEnt *p) {
    const int2 t = __ldg(reinterpret_cast<const int2 *>(p));
    FpEnt e; e.idx = t.x; e.w = __int_as_float(t.y);
    return e;
}

// fwd: unit = (superpixel k, group, 128-channel chunk), one float4 per lane.  `split` warps share a unit
// (entry e goes to warp e mod split) and meet in shared memory in a fixed order; a 256-thread block holds
// 8/split units.  UF independent 128-bit loads per lane are in flight on the long lists.
template <int UF>
__global__ void __launch_bounds__(FP_THREADS, 4) fp_pool_fwd_kernel(const Levels L, const FpPlan P, const PoolPlan F,
                                                                    const int32_t *__restrict__ seg_offsets,
                                                                    const int32_t *__restrict__ seg_pixels, int N,
                                                                    float *__restrict__ pooled) {
    __shared__ float4 part[FP_WARPS][32];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int g = 0;
    while (g + 1 < F.n && (int)blockIdx.x >= F.g[g + 1].blk0) ++g;
    const PoolGroup fg = F.g[g];
    const int split = fg.split;
    const int u_local = wid / split, sub = wid - u_local * split;
    const unsigned unit = (unsigned)((int)blockIdx.x - fg.blk0) * (unsigned)(FP_WARPS / split) + (unsigned)u_local;   // host: < 2^31
    const bool unit_ok = unit < (unsigned)N * (unsigned)fg.nchunk;
    // chunk-major: the units of a block are CONSECUTIVE superpixels (raster neighbours) on the same channels,
    // so the low-resolution cells they share are served by L1 instead of L2
    int k = 0, chunk = 0;
    if (unit_ok) {
        chunk = (int)(unit / (unsigned)N);
        k = (int)(unit - (unsigned)chunk * (unsigned)N);
    }
    const int c = chunk * 128 + lane * 4;
    const bool live = unit_ok && c < fg.Cg;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int n_px = 0;
    if (live) {
        const int beg = __ldg(seg_offsets + k), end = __ldg(seg_offsets + k + 1);
        n_px = end - beg;
        int l, cl;
        locate_level(L, fg.l0, fg.l1, c, l, cl);
        const int Cl = L.C[l];
        const float *__restrict__ src = L.src[l] + cl;
        if (fg.res < 0) {
            const int32_t *__restrict__ px = seg_pixels + beg;
            const int ne = n_px;
            int e = sub;
            for (; e + (UF - 1) * split < ne; e += UF * split) {
                int p[UF];
#pragma unroll
                for (int u = 0; u < UF; ++u) p[u] = __ldg(px + e + u * split);
                float4 v[UF];
#pragma unroll
                for (int u = 0; u < UF; ++u) v[u] = __ldg(reinterpret_cast<const float4 *>(src + (long)p[u] * Cl));
#pragma unroll
                for (int u = 0; u < UF; ++u) acc = acc + v[u];
            }
            for (; e + split < ne; e += 2 * split) {
                const int p0 = __ldg(px + e), p1 = __ldg(px + e + split);
                const float4 v0 = __ldg(reinterpret_cast<const float4 *>(src + (long)p0 * Cl));
                const float4 v1 = __ldg(reinterpret_cast<const float4 *>(src + (long)p1 * Cl));
                acc = acc + v0; acc = acc + v1;
            }
            for (; e < ne; e += split) acc = acc + __ldg(reinterpret_cast<const float4 *>(src + (long)__ldg(px + e) * Cl));
        } else {
            const FpRes &R = P.r[fg.res];                 // an empty superpixel has an empty list
            const int2 span = __ldg(R.fwd_span + k);
            const FpEnt *__restrict__ ent = R.fwd_ent + span.x;
            const int ne = span.y;
            int e = sub;
            for (; e + (UF - 1) * split < ne; e += UF * split) {
                FpEnt a[UF];
#pragma unroll
                for (int u = 0; u < UF; ++u) a[u] = ld_ent(ent + e + u * split);
                float4 v[UF];
#pragma unroll
                for (int u = 0; u < UF; ++u) v[u] = __ldg(reinterpret_cast<const float4 *>(src + (long)a[u].idx * Cl));
#pragma unroll
                for (int u = 0; u < UF; ++u) fma4(acc, a[u].w, v[u]);
            }
            for (; e + split < ne; e += 2 * split) {
                const FpEnt a0 = ld_ent(ent + e), a1 = ld_ent(ent + e + split);
                const float4 v0 = __ldg(reinterpret_cast<const float4 *>(src + (long)a0.idx * Cl));
                const float4 v1 = __ldg(reinterpret_cast<const float4 *>(src + (long)a1.idx * Cl));
                fma4(acc, a0.w, v0); fma4(acc, a1.w, v1);
            }
            for (; e < ne; e += split) {
                const FpEnt a0 = ld_ent(ent + e);
                fma4(acc, a0.w, __ldg(reinterpret_cast<const float4 *>(src + (long)a0.idx * Cl)));
            }
        }
    }
    if (split > 1) {                        // block-uniform
        part[wid][lane] = acc;
        __syncthreads();
        if (sub == 0)
            for (int s_ = 1; s_ < split; ++s_) acc = acc + part[wid + s_][lane];
    }
    if (live && sub == 0) {
        const float inv = n_px > 0 ? 1.0f / (float)n_px : 0.f;
        *reinterpret_cast<float4 *>(pooled + (long)k * L.Ctot + fg.coff + c) = inv * acc;
    }
}

// fwd, whole cells (default): a warp owns (superpixel k, LEVEL l) and reads every listed cell of the level in
// full -- C_l * 4 contiguous bytes: V = C_l / 128 consecutive 128-bit loads per lane, or, for C_l = 64 / 32, EPS = 2 / 4
// consecutive entries side by side in one warp load -- so that the loads a warp has in flight cover 2 - 4 KB of
// contiguous memory (a run of cells of one bounding-box row), not eight scattered 512-byte pieces: tools/l2bw.cu
// measures 6.0 - 6.6 TB/s for gathers of 2 - 4 KB pieces against 2.3 TB/s for 512-byte ones.  The list lives in
// registers, one entry per lane (blocks of 32), index and weight broadcast with shuffles; eight loads per lane in
// flight; no shared memory, no barrier, 64-thread blocks so that a finished warp frees its slot early.
constexpr int FC_THREADS = 64;
constexpr int FC_WARPS = FC_THREADS / 32;
constexpr int FC_INFLIGHT = 8;

// accumulate the 32-entry chunks first_chunk, first_chunk + chunk_stride, ... of one (superpixel, level) list; with
// EPS > 1 entries side by side the entry groups of the warp are folded at the end (fixed order) and lanes < 32 / EPS
// hold the sums in acc[0]
template <int V, int EPS>
__device__ __forceinline__ void fc_accumulate(const float *__restrict__ level, int Cl, const int32_t *__restrict__ px,
                                              const FpEnt *__restrict__ ent, int ne, int lane, int first_chunk, int chunk_stride,
                                              float4 (&acc)[V]) {
    constexpr int UF = FC_INFLIGHT / V;                     // entry slots per batch
    constexpr int LPE = 32 / EPS;                           // lanes per entry
    const int sub = lane / LPE;                             // which entry of a slot this lane reads
    const float *__restrict__ src = level + (lane - sub * LPE) * 4;
#pragma unroll
    for (int q = 0; q < V; ++q) acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int e0 = 32 * first_chunk; e0 < ne; e0 += 32 * chunk_stride) {
        const int eg = e0 + lane;
        const bool valid = eg < ne;
        const int ee = valid ? eg : e0;                     // the tail repeats entry e0 with weight 0
        int my_idx;
        float my_w;
        if (px) { my_idx = __ldg(px + ee); my_w = valid ? 1.0f : 0.f; }
        else { const FpEnt a = ld_ent(ent + ee); my_idx = a.idx; my_w = valid ? a.w : 0.f; }
        const int nb = min(32, ne - e0);
        for (int e = 0; e < nb; e += UF * EPS) {            // padded batches: slots past the list carry weight 0 and a valid index
            float4 v[UF][V];
            float wv[UF];
#pragma unroll
            for (int u = 0; u < UF; ++u) {
                const int slot = e + u * EPS + sub;
                const int idx = __shfl_sync(0xffffffffu, my_idx, slot & 31);
                const float ws = __shfl_sync(0xffffffffu, my_w, slot & 31);
                wv[u] = slot < 32 ? ws : 0.f;
                const float4 *row = reinterpret_cast<const float4 *>(src + (long)idx * Cl);
#pragma unroll
                for (int q = 0; q < V; ++q) v[u][q] = __ldg(row + 32 * q);
            }
#pragma unroll
            for (int u = 0; u < UF; ++u)
#pragma unroll
                for (int q = 0; q < V; ++q) fma4(acc[q], wv[u], v[u][q]);
        }
    }
    if (EPS > 1) {                                          // fold the entry groups of the warp (fixed order)
#pragma unroll
        for (int o = 16; o >= LPE; o >>= 1) {
            acc[0].x += __shfl_down_sync(0xffffffffu, acc[0].x, o);
            acc[0].y += __shfl_down_sync(0xffffffffu, acc[0].y, o);
            acc[0].z += __shfl_down_sync(0xffffffffu, acc[0].z, o);
            acc[0].w += __shfl_down_sync(0xffffffffu, acc[0].w, o);
        }
    }
}

// one (superpixel, level) unit by `split` warps (split == 1: the warp alone; else the warps of the block take the
// list's 32-entry chunks round robin and meet in shared memory, summed in warp order => deterministic)
template <int V, int EPS>
__device__ __forceinline__ void fc_unit(const float *__restrict__ level, int Cl, const int32_t *__restrict__ px, const FpEnt *__restrict__ ent,
                                        int ne, float inv, float *__restrict__ out, int lane, int wid, int split, float4 *red) {
    constexpr int LPE = 32 / EPS;
    float4 acc[V];
    fc_accumulate<V, EPS>(level, Cl, px, ent, ne, lane, split > 1 ? wid : 0, split, acc);
    if (split > 1) {
#pragma unroll
        for (int q = 0; q < V; ++q) red[(wid * V + q) * 32 + lane] = acc[q];
        __syncthreads();
        if (wid != 0) return;
        for (int w = 1; w < split; ++w)
#pragma unroll
            for (int q = 0; q < V; ++q) acc[q] = acc[q] + red[(w * V + q) * 32 + lane];
    }
    if (EPS > 1) {
        if (lane < LPE) *reinterpret_cast<float4 *>(out + lane * 4) = inv * acc[0];
    } else {
#pragma unroll
        for (int q = 0; q < V; ++q) *reinterpret_cast<float4 *>(out + q * 128 + lane * 4) = inv * acc[q];
    }
}

// levels of 32, 64, 128, 256 or 512 channels (host-checked); units are level-major, so blocks are level-homogeneous.
// split == 1: FC_WARPS units per block, a warp each; split > 1 (large superpixels: few, long lists -- r2 config-3 sweep:
// 7.5 ms at 2048^2 / N = 500 with one warp per unit): one unit per block of `split` warps.
__global__ void __launch_bounds__(FC_THREADS, 12) fp_pool_fwd_cells_kernel(const Levels L, const FpPlan P, const PoolPlan F,
                                                                           const int32_t *__restrict__ seg_offsets,
                                                                           const int32_t *__restrict__ seg_pixels, int N,
                                                                           float *__restrict__ pooled) {
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bpl = (N + FC_WARPS - 1) / FC_WARPS;          // blocks per level
    const int l = (int)blockIdx.x / bpl;
    const int k = ((int)blockIdx.x - l * bpl) * FC_WARPS + wid;
    if (k >= N) return;
    int g = 0;
    while (g + 1 < F.n && l >= F.g[g + 1].l0) ++g;
    const int res = F.g[g].res;
    const int beg = __ldg(seg_offsets + k), n_px = __ldg(seg_offsets + k + 1) - beg;
    const int32_t *__restrict__ px = nullptr;
    const FpEnt *__restrict__ ent = nullptr;
    int ne = n_px;
    if (res < 0) {
        px = seg_pixels + beg;
    } else {
        const int2 span = __ldg(P.r[res].fwd_span + k);     // an empty superpixel has an empty list
        ent = P.r[res].fwd_ent + span.x;
        ne = span.y;
    }
    const float inv = n_px > 0 ? 1.0f / (float)n_px : 0.f;
    float *__restrict__ out = pooled + (long)k * L.Ctot + L.coff[l];
    const int Cl = L.C[l];
    const float *__restrict__ level = L.src[l];
    if (Cl == 512) fc_unit<4, 1>(level, Cl, px, ent, ne, inv, out, lane, 0, 1, nullptr);
    else if (Cl == 256) fc_unit<2, 1>(level, Cl, px, ent, ne, inv, out, lane, 0, 1, nullptr);
    else if (Cl == 128) fc_unit<1, 1>(level, Cl, px, ent, ne, inv, out, lane, 0, 1, nullptr);
    else if (Cl == 64) fc_unit<1, 2>(level, Cl, px, ent, ne, inv, out, lane, 0, 1, nullptr);
    else fc_unit<1, 4>(level, Cl, px, ent, ne, inv, out, lane, 0, 1, nullptr);
}

constexpr int FC_MAX_SPLIT = 16;
__global__ void __launch_bounds__(32 * FC_MAX_SPLIT) fp_pool_fwd_cells_split_kernel(const Levels L, const FpPlan P, const PoolPlan F,
                                                                                    const int32_t *__restrict__ seg_offsets,
                                                                                    const int32_t *__restrict__ seg_pixels, int N,
                                                                                    float *__restrict__ pooled) {
    extern __shared__ float4 fc_red[];                      // split x V x 32 partial sums
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31, split = blockDim.x >> 5;
    const int l = (int)blockIdx.x / N;
    const int k = (int)blockIdx.x - l * N;
    int g = 0;
    while (g + 1 < F.n && l >= F.g[g + 1].l0) ++g;
    const int res = F.g[g].res;
    const int beg = __ldg(seg_offsets + k), n_px = __ldg(seg_offsets + k + 1) - beg;
    const int32_t *__restrict__ px = nullptr;
    const FpEnt *__restrict__ ent = nullptr;
    int ne = n_px;
    if (res < 0) {
        px = seg_pixels + beg;
    } else {
        const int2 span = __ldg(P.r[res].fwd_span + k);
        ent = P.r[res].fwd_ent + span.x;
        ne = span.y;
    }
    const float inv = n_px > 0 ? 1.0f / (float)n_px : 0.f;
    float *__restrict__ out = pooled + (long)k * L.Ctot + L.coff[l];
    const int Cl = L.C[l];
    const float *__restrict__ level = L.src[l];
    if (Cl == 512) fc_unit<4, 1>(level, Cl, px, ent, ne, inv, out, lane, wid, split, fc_red);
    else if (Cl == 256) fc_unit<2, 1>(level, Cl, px, ent, ne, inv, out, lane, wid, split, fc_red);
    else if (Cl == 128) fc_unit<1, 1>(level, Cl, px, ent, ne, inv, out, lane, wid, split, fc_red);
    else if (Cl == 64) fc_unit<1, 2>(level, Cl, px, ent, ne, inv, out, lane, wid, split, fc_red);
    else fc_unit<1, 4>(level, Cl, px, ent, ne, inv, out, lane, wid, split, fc_red);
}

// bwd: unit = (low-resolution cell q, group, 256-channel slice), one warp per unit, two float4 per lane;
// the pooled-gradient rows (N x Ctot, L2-resident) of four list entries are in flight together.
template <int U>
__global__ void __launch_bounds__(FP_THREADS) fp_pool_bwd_kernel(const Levels L, const FpPlan P, const PoolPlan B,
                                                                 const float *__restrict__ gp) {
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int g = 0;
    while (g + 1 < B.n && (int)blockIdx.x >= B.g[g + 1].blk0) ++g;
    const PoolGroup bg = B.g[g];
    const FpRes &R = P.r[bg.res];
    const unsigned unit = (unsigned)((int)blockIdx.x - bg.blk0) * FP_WARPS + (unsigned)wid;          // host: < 2^31
    const unsigned cells = (unsigned)(R.h * R.w);
    if (unit >= cells * (unsigned)bg.nchunk) return;
    // slice-major: the warps of a block are consecutive cells on the same channels and share pooled-gradient rows (L1)
    const int slice = (int)(unit / cells);
    const long q = (long)(unit - (unsigned)slice * cells);
    const int c0 = slice * 256 + lane * 4, c1 = c0 + 128;
    const bool live0 = c0 < bg.Cg, live1 = c1 < bg.Cg;
    const int Ctot = L.Ctot;
    const float *__restrict__ gpg = gp + bg.coff;
    const int2 span = __ldg(R.bwd_span + q);
    const FpEnt *__restrict__ ent = R.bwd_ent + span.x;
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 acc0 = zero, acc1 = zero;
    int e = 0;
    for (; e + U <= span.y; e += U) {
        FpEnt a[U];
#pragma unroll
        for (int u = 0; u < U; ++u) a[u] = ld_ent(ent + e + u);
        float4 v0[U], v1[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const float *row = gpg + (long)a[u].idx * Ctot;
            v0[u] = live0 ? __ldg(reinterpret_cast<const float4 *>(row + c0)) : zero;
            v1[u] = live1 ? __ldg(reinterpret_cast<const float4 *>(row + c1)) : zero;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) { fma4(acc0, a[u].w, v0[u]); fma4(acc1, a[u].w, v1[u]); }
    }
    for (; e < span.y; ++e) {
        const FpEnt a0 = ld_ent(ent + e);
        const float *row = gpg + (long)a0.idx * Ctot;
        const float4 v0 = live0 ? __ldg(reinterpret_cast<const float4 *>(row + c0)) : zero;
        const float4 v1 = live1 ? __ldg(reinterpret_cast<const float4 *>(row + c1)) : zero;
        fma4(acc0, a0.w, v0); fma4(acc1, a0.w, v1);
    }
    int l, cl;
    if (live0) {
        locate_level(L, bg.l0, bg.l1, c0, l, cl);
        stg_stream(reinterpret_cast<float4 *>(L.dst[l] + q * L.C[l] + cl), acc0);
    }
    if (live1) {
        locate_level(L, bg.l0, bg.l1, c1, l, cl);
        stg_stream(reinterpret_cast<float4 *>(L.dst[l] + q * L.C[l] + cl), acc1);
    }
}

// bwd, whole cells (default for levels of 128 / 256 / 512 channels): a warp writes complete cell rows of ONE level
// (C_l * 4 contiguous bytes, V = C_l / 128 float4 per lane) and keeps the lists in registers.
//   * short lists (fine levels: a 4x4 / 8x8 footprint meets 1 - 3 superpixels): the warp takes 32 CONSECUTIVE cells;
//     lane i loads the span and the first EPRE entries of cell q0 + i -- one round of coalesced loads for 32 cells
//     instead of a span -> entries chain per cell -- and the cells are then walked CU at a time with their
//     pooled-gradient rows (L2-resident) in flight together; entries beyond EPRE take a plain loop;
//   * long lists (coarse levels): one cell per warp, the list one entry per lane, broadcast with shuffles,
//     8 / V entries in flight -- the forward kernel transposed.
constexpr int BC_THREADS = 128;
constexpr int BC_WARPS = BC_THREADS / 32;
constexpr int BC_MAX = WESUP_MAX_LEVELS;
constexpr int BC_SLICE = 512;            // widest channel slice a warp gathers (r2 measurements: halving it to 256 -- 64
                                         // registers, eight blocks per SM -- LOST 10 %: the piece size of the gathers matters
                                         // more than the occupancy, as in the forward kernel)
struct BwdCellsPlan {
    int n;                               // (level, channel slice) pairs handled here
    short lvl[BC_MAX], res[BC_MAX], batch[BC_MAX], choff[BC_MAX], cw[BC_MAX];
    int blk0[BC_MAX + 1];
};

template <int V, int EPRE, int CU>
__device__ __forceinline__ void bc_batch(const float *__restrict__ gpl, int Ctot, float *__restrict__ dst, int Cl, const FpRes &R,
                                         int q0, int cells, int lane) {
    const int q = q0 + lane;
    const int2 span = q < cells ? __ldg(R.bwd_span + q) : make_int2(0, 0);
    FpEnt pre[EPRE];
#pragma unroll
    for (int u = 0; u < EPRE; ++u) {
        pre[u].idx = 0; pre[u].w = 0.f;                     // row 0 with weight 0: a valid address, no contribution
        if (u < span.y) pre[u] = ld_ent(R.bwd_ent + span.x + u);
    }
    const int nc = min(32, cells - q0);
    for (int c0 = 0; c0 < nc; c0 += CU) {
        float4 v[CU][EPRE][V];
        float wv[CU][EPRE];
        int n_c[CU];
#pragma unroll
        for (int t = 0; t < CU; ++t) {
            const int c = min(c0 + t, nc - 1);              // the last group may repeat its last cell (stores are guarded)
            n_c[t] = __shfl_sync(0xffffffffu, span.y, c);
#pragma unroll
            for (int u = 0; u < EPRE; ++u) {
                const int k = __shfl_sync(0xffffffffu, pre[u].idx, c);
                wv[t][u] = __shfl_sync(0xffffffffu, pre[u].w, c);
                const float4 *row = reinterpret_cast<const float4 *>(gpl + (long)k * Ctot);
#pragma unroll
                for (int qv = 0; qv < V; ++qv) v[t][u][qv] = __ldg(row + 32 * qv);
            }
        }
#pragma unroll
        for (int t = 0; t < CU; ++t) {
            if (c0 + t >= nc) break;                        // warp-uniform
            float4 acc[V];
#pragma unroll
            for (int qv = 0; qv < V; ++qv) acc[qv] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int u = 0; u < EPRE; ++u)
#pragma unroll
                for (int qv = 0; qv < V; ++qv) fma4(acc[qv], wv[t][u], v[t][u][qv]);
            if (n_c[t] > EPRE) {                            // warp-uniform: the rest of a long list
                const int base = __shfl_sync(0xffffffffu, span.x, c0 + t);
                for (int e = EPRE; e < n_c[t]; ++e) {
                    const FpEnt a = ld_ent(R.bwd_ent + base + e);
                    const float4 *row = reinterpret_cast<const float4 *>(gpl + (long)a.idx * Ctot);
#pragma unroll
                    for (int qv = 0; qv < V; ++qv) fma4(acc[qv], a.w, __ldg(row + 32 * qv));
                }
            }
            float4 *o = reinterpret_cast<float4 *>(dst + (long)(q0 + c0 + t) * Cl);
#pragma unroll
            for (int qv = 0; qv < V; ++qv) stg_stream(o + 32 * qv, acc[qv]);
        }
    }
}

template <int V>
__device__ __forceinline__ void bc_cell(const float *__restrict__ gpl, int Ctot, float *__restrict__ dst, int Cl, const FpRes &R, int q,
                                        int lane) {
    constexpr int UF = 8 / V;
    const int2 span = __ldg(R.bwd_span + q);
    const FpEnt *__restrict__ ent = R.bwd_ent + span.x;
    const int ne = span.y;
    float4 acc[V];
#pragma unroll
    for (int qv = 0; qv < V; ++qv) acc[qv] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int e0 = 0; e0 < ne; e0 += 32) {
        const int eg = e0 + lane;
        const bool valid = eg < ne;
        const FpEnt mine = ld_ent(ent + (valid ? eg : e0));  // the tail repeats entry e0 with weight 0
        const float my_w = valid ? mine.w : 0.f;
        const int nb = min(32, ne - e0);
        for (int e = 0; e < nb; e += UF) {                  // padded batches
            float4 v[UF][V];
            float wv[UF];
#pragma unroll
            for (int u = 0; u < UF; ++u) {
                const int k = __shfl_sync(0xffffffffu, mine.idx, (e + u) & 31);
                const float ws = __shfl_sync(0xffffffffu, my_w, (e + u) & 31);
                wv[u] = e + u < 32 ? ws : 0.f;
                const float4 *row = reinterpret_cast<const float4 *>(gpl + (long)k * Ctot);
#pragma unroll
                for (int qv = 0; qv < V; ++qv) v[u][qv] = __ldg(row + 32 * qv);
            }
#pragma unroll
            for (int u = 0; u < UF; ++u)
#pragma unroll
                for (int qv = 0; qv < V; ++qv) fma4(acc[qv], wv[u], v[u][qv]);
        }
    }
    float4 *o = reinterpret_cast<float4 *>(dst + (long)q * Cl);
#pragma unroll
    for (int qv = 0; qv < V; ++qv) stg_stream(o + 32 * qv, acc[qv]);
}

// one warp unit of the cell lists: `wunit` counts 32-cell batches (fine levels) or single cells (coarse levels) of entry i
__device__ __forceinline__ void bc_unit(const Levels &L, const FpPlan &P, const BwdCellsPlan &B, const float *__restrict__ gp, int i,
                                        int wunit, int lane) {
    const int l = B.lvl[i];
    const FpRes &R = P.r[B.res[i]];
    const int cells = R.h * R.w;
    const int Cl = L.C[l], Ctot = L.Ctot, cw = B.cw[i];
    const float *__restrict__ gpl = gp + L.coff[l] + B.choff[i] + lane * 4;
    float *__restrict__ dst = L.dst[l] + B.choff[i] + lane * 4;
    if (B.batch[i]) {
        const int q0 = wunit * 32;
        if (q0 >= cells) return;
        if (cw == 128) bc_batch<1, 2, 4>(gpl, Ctot, dst, Cl, R, q0, cells, lane);
        else if (cw == 256) bc_batch<2, 2, 2>(gpl, Ctot, dst, Cl, R, q0, cells, lane);
        else bc_batch<4, 2, 1>(gpl, Ctot, dst, Cl, R, q0, cells, lane);
    } else {
        if (wunit >= cells) return;
        if (cw == 128) bc_cell<1>(gpl, Ctot, dst, Cl, R, wunit, lane);
        else if (cw == 256) bc_cell<2>(gpl, Ctot, dst, Cl, R, wunit, lane);
        else bc_cell<4>(gpl, Ctot, dst, Cl, R, wunit, lane);
    }
}

__global__ void __launch_bounds__(BC_THREADS, 6) fp_pool_bwd_cells_kernel(const Levels L, const FpPlan P, const BwdCellsPlan B,
                                                                          const float *__restrict__ gp) {
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int i = 0;
    while (i + 1 < B.n && (int)blockIdx.x >= B.blk0[i + 1]) ++i;
    bc_unit(L, P, B, gp, i, ((int)blockIdx.x - B.blk0[i]) * BC_WARPS + wid, lane);
}

// full-resolution levels: grad[p, c] = grad_pooled[row(p), c] / |S_row(p)|.  A warp takes 32 consecutive pixels: label and
// 1/|S| of pixel p0 + lane are loaded once, lane-parallel, and broadcast with shuffles; the pooled-gradient row
// (L2-resident) is re-fetched only when the label changes along the run (a superpixel is ~14 pixels wide), so the
// steady state is shuffle, shuffle, four multiplies, one 128-bit streaming store per pixel.
// One warp, 32 consecutive pixels.  The pixels are walked eight at a time: the (warp-uniform) loads of the rows whose label
// differs from the previous pixel's are issued together, the others reuse the previous value, then eight stores follow.
__device__ __forceinline__ void ident_unit(const Levels &L, const PoolGroup &bg, const float *__restrict__ gp,
                                           const int32_t *__restrict__ row_labels, const int32_t *__restrict__ counts, long HW, long p0,
                                           int lane) {
    constexpr int PU = 8;
    const int nch4 = bg.Cg >> 2, Ctot = L.Ctot;
    const int np = (int)min(32L, HW - p0);
    int my_lab = 0;
    float my_scale = 0.f;
    if (lane < np) {
        my_lab = __ldg(row_labels + p0 + lane);
        const int cnt = __ldg(counts + my_lab);
        my_scale = cnt > 0 ? 1.0f / (float)cnt : 0.f;
    }
    for (int cb = 0; cb < nch4; cb += 32) {
        const int c4 = cb + lane;
        const bool live = c4 < nch4;
        int l = bg.l0, cl = 0;
        if (live) locate_level(L, bg.l0, bg.l1, c4 << 2, l, cl);
        const int Cl = L.C[l];
        float *__restrict__ dst = L.dst[l] + cl + p0 * Cl;
        const float *__restrict__ src = gp + bg.coff + (live ? (c4 << 2) : 0);
        int prev = -1;
        float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int i0 = 0; i0 < np; i0 += PU) {
            int lab[PU];
            float sc[PU];
            float4 v[PU];
#pragma unroll
            for (int t = 0; t < PU; ++t) {
                lab[t] = __shfl_sync(0xffffffffu, my_lab, (i0 + t) & 31);
                sc[t] = __shfl_sync(0xffffffffu, my_scale, (i0 + t) & 31);
            }
#pragma unroll
            for (int t = 0; t < PU; ++t) {
                const bool fresh = lab[t] != (t == 0 ? prev : lab[t - 1]);      // warp-uniform
                v[t] = val;
                if (fresh) v[t] = __ldg(reinterpret_cast<const float4 *>(src + (long)lab[t] * Ctot));
            }
#pragma unroll
            for (int t = 0; t < PU; ++t) {
                const bool fresh = lab[t] != (t == 0 ? prev : lab[t - 1]);
                if (!fresh) v[t] = t == 0 ? val : v[t - 1];
                if (live && i0 + t < np) stg_stream(reinterpret_cast<float4 *>(dst + (long)(i0 + t) * Cl), sc[t] * v[t]);
            }
            prev = lab[PU - 1];
            val = v[PU - 1];
        }
    }
}

__global__ void __launch_bounds__(256) fp_pool_bwd_ident_kernel(const Levels L, const PoolGroup bg, const float *__restrict__ gp,
                                                                const int32_t *__restrict__ row_labels,
                                                                const int32_t *__restrict__ counts, long HW) {
    const long p0 = ((long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32;
    if (p0 >= HW) return;
    ident_unit(L, bg, gp, row_labels, counts, HW, p0, threadIdx.x & 31);
}

// Both kinds of work in ONE launch (default).  As two launches on two streams the write stream of the identity levels
// and the latency-bound list gathers did not overlap: a full wave of either kernel owns the register file.  Here the
// block index decides the role, identity blocks spread evenly between the list blocks (coarse levels first), so both are
// resident on every SM at all times.
struct BwdAllPlan { int ident_blocks, total_blocks, order; };
template <int MINB>
__global__ void __launch_bounds__(BC_THREADS, MINB) fp_pool_bwd_all_kernel(const Levels L, const FpPlan P, const BwdCellsPlan B,
                                                                        const PoolGroup ig, const BwdAllPlan A,
                                                                        const float *__restrict__ gp,
                                                                        const int32_t *__restrict__ row_labels,
                                                                        const int32_t *__restrict__ counts, long HW) {
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long b = blockIdx.x;
    int before;                                                                 // identity blocks among [0, b)
    bool is_ident;
    if (A.order == 1) { is_ident = b < A.ident_blocks; before = is_ident ? (int)b : A.ident_blocks; }
    else if (A.order == 2) { const int nc_ = A.total_blocks - A.ident_blocks; is_ident = b >= nc_; before = is_ident ? (int)b - nc_ : 0; }
    else {
        before = (int)(b * A.ident_blocks / A.total_blocks);
        is_ident = (int)((b + 1) * A.ident_blocks / A.total_blocks) > before;
    }
    if (is_ident) {
        const long p0 = ((long)before * BC_WARPS + wid) * 32;
        if (p0 < HW) ident_unit(L, ig, gp, row_labels, counts, HW, p0, lane);
        return;
    }
    const int cb = (int)b - before;
    int i = 0;
    while (i + 1 < B.n && cb >= B.blk0[i + 1]) ++i;
    bc_unit(L, P, B, gp, i, (cb - B.blk0[i]) * BC_WARPS + wid, lane);
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
// one auxiliary stream owned by the library (created once per process)
struct AuxStream {
    cudaStream_t s;
    cudaEvent_t fork, join;
    bool ok = false;
};
static AuxStream *aux_stream() {
    static AuxStream a;
    static bool tried = false;
    if (!tried) {
        tried = true;
        a.ok = cudaStreamCreateWithFlags(&a.s, cudaStreamNonBlocking) == cudaSuccess &&
               cudaEventCreateWithFlags(&a.fork, cudaEventDisableTiming) == cudaSuccess &&
               cudaEventCreateWithFlags(&a.join, cudaEventDisableTiming) == cudaSuccess;
        if (!a.ok) cudaGetLastError();
    }
    return a.ok ? &a : nullptr;
}

static int fill_levels(Levels &L, const char *who, const void *const *ptrs, bool is_dst, const int *C, const int *h, const int *w,
                       int n_levels, int H, int W) {
    L.n = n_levels; L.H = H; L.W = W;
    int off = 0;
    for (int l = 0; l < n_levels; ++l) {
        WESUP_REQUIRE(C[l] > 0 && h[l] > 0 && w[l] > 0, WESUP_E_ARG, "%s: level %d has empty shape", who, l);
        WESUP_REQUIRE(C[l] % 4 == 0, WESUP_E_ALIGN, "%s: C[%d]=%d must be a multiple of 4", who, l, C[l]);
        WESUP_REQUIRE(ptrs[l] != nullptr && aligned16(ptrs[l]), WESUP_E_ALIGN, "%s: level %d pointer null or unaligned", who, l);
        L.src[l] = is_dst ? nullptr : static_cast<const float *>(ptrs[l]);
        L.dst[l] = is_dst ? static_cast<float *>(const_cast<void *>(ptrs[l])) : nullptr;
        L.C[l] = C[l]; L.h[l] = h[l]; L.w[l] = w[l]; L.coff[l] = off;
        L.sy[l] = bilinear_scale(h[l], H); L.sx[l] = bilinear_scale(w[l], W);
        off += C[l];
    }
    L.Ctot = off;
    return 0;
}

static int build_pool_groups(PoolPlan &F, const Levels &L, const FpPlan &P) {
    F.n = 0;
    for (int l = 0; l < L.n; ++l) {
        const int g = F.n - 1;
        if (g >= 0 && L.h[F.g[g].l0] == L.h[l] && L.w[F.g[g].l0] == L.w[l]) {
            F.g[g].l1 = l + 1;
            F.g[g].Cg += L.C[l];
            continue;
        }
        PoolGroup &G = F.g[F.n++];
        G.l0 = l; G.l1 = l + 1; G.coff = L.coff[l]; G.Cg = L.C[l];
        G.nchunk = 0; G.split = 1; G.blk0 = 0;
        G.res = -1;
        if (!(L.h[l] == L.H && L.w[l] == L.W)) {
            for (int i = 0; i < P.n; ++i)
                if (P.r[i].h == L.h[l] && P.r[i].w == L.w[l]) G.res = i;
            if (G.res < 0) return -1;
        }
    }
    return 0;
}

static inline bool fp_common_args_ok(int n_levels, int H, int W, int N) {
    return n_levels > 0 && n_levels <= WESUP_MAX_LEVELS && H > 0 && W > 0 && N > 0 && (long)H * W < (1L << 29) && H < 65536 && W < 65536;
}

}  // namespace wesup

using namespace wesup;

extern "C" size_t wesup_footprint_bytes(const int *h, const int *w, int n_levels, int H, int W, int N) {
    if (!h || !w || !fp_common_args_ok(n_levels, H, W, N)) return 0;
    for (int l = 0; l < n_levels; ++l)
        if (h[l] <= 0 || w[l] <= 0) return 0;
    FpPlan P;
    const long bytes = plan_footprint(P, h, w, n_levels, H, W, N, nullptr);
    return bytes < 0 ? 0 : (size_t)bytes;
}

extern "C" int wesup_footprint_build(const int *h, const int *w, int n_levels, int H, int W, int N, const int32_t *seg_offsets,
                                     const int32_t *seg_pixels, const int32_t *row_labels, const int32_t *counts, int with_bwd,
                                     void *fp, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    WESUP_REQUIRE(h && w && seg_offsets && seg_pixels && fp, WESUP_E_ARG, "wesup_footprint_build: null pointer");
    WESUP_REQUIRE(!with_bwd || (row_labels && counts), WESUP_E_ARG, "wesup_footprint_build: the backward lists need row_labels and counts");
    WESUP_REQUIRE(fp_common_args_ok(n_levels, H, W, N), WESUP_E_ARG, "wesup_footprint_build: bad size n_levels=%d H=%d W=%d N=%d", n_levels, H, W, N);
    WESUP_REQUIRE(aligned16(fp), WESUP_E_ALIGN, "wesup_footprint_build: fp must be 16-byte aligned");
    for (int l = 0; l < n_levels; ++l)
        WESUP_REQUIRE(h[l] > 0 && w[l] > 0, WESUP_E_ARG, "wesup_footprint_build: level %d has empty shape", l);
    FpPlan P;
    WESUP_REQUIRE(plan_footprint(P, h, w, n_levels, H, W, N, static_cast<char *>(fp)) >= 0, WESUP_E_UNSUPPORTED,
                  "wesup_footprint_build: more than %d distinct level resolutions", FP_MAX_RES);
    if (P.n == 0) return 0;                               // full-resolution levels only: nothing to precompute
    int launched = 0;
    cudaMemsetAsync(P.cursors, 0, P.cursor_bytes, stream);
    // shared grids: 2.5 x the cells a square superpixel of the mean area covers over all resolutions (SLIC superpixels
    // are ragged), between FP_GRID_CAP and FP_GRID_CAP_MAX, a power of two
    int grid_cap = FP_GRID_CAP;
    {
        const double side = sqrt((double)H * W / (double)N);
        double cells = 0.0;
        for (int r = 0; r < P.n; ++r) cells += (side * P.r[r].sy + 3.0) * (side * P.r[r].sx + 3.0);
        while (grid_cap < FP_GRID_CAP_MAX && (double)grid_cap < 2.5 * cells) grid_cap *= 2;
    }
    const size_t build_smem = (size_t)grid_cap * (sizeof(unsigned) + sizeof(FpEnt));
    if (build_smem > 40 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(fp_build_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)build_smem);
        WESUP_REQUIRE(e == cudaSuccess, (int)e, "wesup_footprint_build: smem attribute: %s", cudaGetErrorString(e));
    }
    fp_build_fwd_kernel<<<N, FP_THREADS, build_smem, stream>>>(P, seg_offsets, seg_pixels, grid_cap);
    ++launched;
    if (with_bwd) {
        int in_max = 1, blocks = 0;
        for (int r = P.n - 1; r >= 0; --r) {
            in_max = in_max > P.r[r].h ? in_max : P.r[r].h;
            in_max = in_max > P.r[r].w ? in_max : P.r[r].w;
            blocks += cdiv((long)P.r[r].h * P.r[r].w, FP_WARPS);
            P.r[r].blk1 = blocks;
        }
        fp_axis_tables_kernel<<<dim3(cdiv(in_max, 128), 2 * P.n), 128, 0, stream>>>(P);
        fp_build_bwd_kernel<<<blocks, FP_THREADS, 0, stream>>>(P, row_labels, counts);
        launched += 2;
    }
    WESUP_CHECK_LAUNCH("wesup_footprint_build", launched);
    return 0;
}

extern "C" int wesup_levels_pool_fwd_fp(const void *const *level, const int *C, const int *h, const int *w, int n_levels, int H,
                                        int W, const int32_t *seg_offsets, const int32_t *seg_pixels, int N, const void *fp,
                                        float *pooled, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    WESUP_REQUIRE(level && C && h && w && seg_offsets && seg_pixels && fp && pooled, WESUP_E_ARG, "wesup_levels_pool_fwd_fp: null pointer");
    WESUP_REQUIRE(fp_common_args_ok(n_levels, H, W, N), WESUP_E_ARG, "wesup_levels_pool_fwd_fp: bad size n_levels=%d H=%d W=%d N=%d", n_levels, H, W, N);
    WESUP_REQUIRE(aligned16(pooled) && aligned16(fp), WESUP_E_ALIGN, "wesup_levels_pool_fwd_fp: pooled and fp must be 16-byte aligned");
    Levels L;
    int rc = fill_levels(L, "wesup_levels_pool_fwd_fp", level, false, C, h, w, n_levels, H, W);
    if (rc) return rc;
    FpPlan P;
    WESUP_REQUIRE(plan_footprint(P, h, w, n_levels, H, W, N, static_cast<char *>(const_cast<void *>(fp))) >= 0, WESUP_E_UNSUPPORTED,
                  "wesup_levels_pool_fwd_fp: more than %d distinct level resolutions", FP_MAX_RES);
    PoolPlan F;
    WESUP_REQUIRE(build_pool_groups(F, L, P) == 0, WESUP_E_ARG, "wesup_levels_pool_fwd_fp: level resolution missing from the footprint plan");
    // warps per unit from the expected list length: a superpixel of a pixels covers about (sqrt(a)*scale + 2)^2 cells
    const double side = sqrt((double)H * W / (double)N);
    const double per_warp = 48.0;                         // list entries one warp should own (measured: 12 .. 48 within 8 %)
    int blocks = 0;
    for (int g = 0; g < F.n; ++g) {
        PoolGroup &G = F.g[g];
        G.nchunk = (G.Cg + 127) / 128;
        double expect = side * side;
        if (G.res >= 0) expect = (side * P.r[G.res].sy + 2.0) * (side * P.r[G.res].sx + 2.0);
        G.split = expect >= 8.0 * per_warp ? 8 : expect >= 4.0 * per_warp ? 4 : expect >= 2.0 * per_warp ? 2 : 1;
        G.blk0 = blocks;
        blocks += cdiv((long)N * G.nchunk, FP_WARPS / G.split);
    }
    // default: whole cells per warp (levels of 32 .. 512 channels); any other channel count takes the chunk kernel,
    // which WESUP_FP_FWD=chunks also selects (cross-check in the tests)
    bool cells = getenv("WESUP_FP_FWD") == nullptr;
    for (int l = 0; l < L.n; ++l) cells = cells && (L.C[l] == 32 || L.C[l] == 64 || L.C[l] == 128 || L.C[l] == 256 || L.C[l] == 512);
    if (cells) {
        // warps per (superpixel, level) list: 1 while a superpixel holds < ~768 pixels (its longest list, at the
        // full-resolution levels), else enough to keep a warp's share there, up to 16
        int split = 1;
        while (split < FC_MAX_SPLIT && side * side > 768.0 * split) split *= 2;
        if (const char *e = getenv("WESUP_FP_FWD_SPLIT")) { const int v = atoi(e); if (v >= 1 && v <= FC_MAX_SPLIT) split = v; }
        if (split > 1) {
            const long cb = (long)L.n * N;
            WESUP_REQUIRE(cb < (1L << 31), WESUP_E_UNSUPPORTED, "wesup_levels_pool_fwd_fp: problem too large");
            fp_pool_fwd_cells_split_kernel<<<(int)cb, 32 * split, (size_t)split * 4 * 32 * sizeof(float4), stream>>>(L, P, F, seg_offsets, seg_pixels, N, pooled);
            WESUP_CHECK_LAUNCH("wesup_levels_pool_fwd_fp", 1);
            return 0;
        }
        const long cb = (long)L.n * cdiv(N, FC_WARPS);
        WESUP_REQUIRE(cb < (1L << 31), WESUP_E_UNSUPPORTED, "wesup_levels_pool_fwd_fp: problem too large");
        fp_pool_fwd_cells_kernel<<<(int)cb, FC_THREADS, 0, stream>>>(L, P, F, seg_offsets, seg_pixels, N, pooled);
        WESUP_CHECK_LAUNCH("wesup_levels_pool_fwd_fp", 1);
        return 0;
    }
    WESUP_REQUIRE(blocks >= 0 && (long)blocks * FP_WARPS < (1L << 31), WESUP_E_UNSUPPORTED, "wesup_levels_pool_fwd_fp: problem too large");
    fp_pool_fwd_kernel<8><<<blocks, FP_THREADS, 0, stream>>>(L, P, F, seg_offsets, seg_pixels, N, pooled);
    WESUP_CHECK_LAUNCH("wesup_levels_pool_fwd_fp", 1);
    return 0;
}

extern "C" int wesup_levels_pool_bwd_fp(const float *grad_pooled, const int32_t *row_labels, const int32_t *counts, const int *C,
                                        const int *h, const int *w, int n_levels, int H, int W, int N, const void *fp,
                                        void *const *grad_level, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    WESUP_REQUIRE(grad_pooled && row_labels && counts && C && h && w && fp && grad_level, WESUP_E_ARG, "wesup_levels_pool_bwd_fp: null pointer");
    WESUP_REQUIRE(fp_common_args_ok(n_levels, H, W, N), WESUP_E_ARG, "wesup_levels_pool_bwd_fp: bad size n_levels=%d H=%d W=%d N=%d", n_levels, H, W, N);
    WESUP_REQUIRE(aligned16(grad_pooled) && aligned16(fp), WESUP_E_ALIGN, "wesup_levels_pool_bwd_fp: grad_pooled and fp must be 16-byte aligned");
    Levels L;
    int rc = fill_levels(L, "wesup_levels_pool_bwd_fp", grad_level, true, C, h, w, n_levels, H, W);
    if (rc) return rc;
    FpPlan P;
    WESUP_REQUIRE(plan_footprint(P, h, w, n_levels, H, W, N, static_cast<char *>(const_cast<void *>(fp))) >= 0, WESUP_E_UNSUPPORTED,
                  "wesup_levels_pool_bwd_fp: more than %d distinct level resolutions", FP_MAX_RES);
    PoolPlan G;
    WESUP_REQUIRE(build_pool_groups(G, L, P) == 0, WESUP_E_ARG, "wesup_levels_pool_bwd_fp: level resolution missing from the footprint plan");
    // non-identity levels of 128 / 256 / 512 channels: whole cells per warp, coarse levels (long lists) first; a group
    // with any other channel count (and WESUP_FP_BWD=chunks, the cross-check of the tests) takes the chunk kernel
    const double side = sqrt((double)H * W / (double)N);
    const char *force_env = getenv("WESUP_FP_BWD");                 // "chunks": chunk kernel (cross-check); "split": separate launches
    const bool force_chunks = force_env && strcmp(force_env, "chunks") == 0;
    // experiment knobs (WESUP_FP_X="order,minb")
    int x_order = 0, x_minb = 6;
    if (const char *x = getenv("WESUP_FP_X")) sscanf(x, "%d,%d", &x_order, &x_minb);
    BwdCellsPlan BC;
    BC.n = 0;
    long cblocks = 0;
    bool group_cells[WESUP_MAX_LEVELS];
    for (int g = G.n - 1; g >= 0; --g) {
        group_cells[g] = false;
        if (G.g[g].res < 0) continue;
        bool ok = !force_chunks;
        for (int l = G.g[g].l0; l < G.g[g].l1; ++l) ok = ok && (L.C[l] == 128 || L.C[l] == 256 || L.C[l] == 512);
        group_cells[g] = ok;
        if (!ok) continue;
        const FpRes &R = P.r[G.g[g].res];
        // superpixels met by a cell's footprint (2/scale pixels wide): short lists take the 32-cells-per-warp path
        const double fy = R.sy > 0.f ? 2.0 / R.sy : (double)H, fx = R.sx > 0.f ? 2.0 / R.sx : (double)W;
        const bool batch = (fy / side + 1.0) * (fx / side + 1.0) < 3.5;
        for (int l = G.g[g].l0; l < G.g[g].l1; ++l) {
            for (int off = 0; off < L.C[l]; off += BC_SLICE) {
                const int i = BC.n++;
                BC.lvl[i] = (short)l; BC.res[i] = (short)G.g[g].res; BC.batch[i] = batch ? 1 : 0;
                BC.choff[i] = (short)off; BC.cw[i] = (short)min(BC_SLICE, L.C[l] - off);
                BC.blk0[i] = (int)cblocks;
                const long cells = (long)R.h * R.w;
                cblocks += cdiv(batch ? cdiv(cells, 32) : cells, BC_WARPS);
            }
        }
    }
    BC.blk0[BC.n] = (int)cblocks;
    WESUP_REQUIRE(cblocks < (1L << 28), WESUP_E_UNSUPPORTED, "wesup_levels_pool_bwd_fp: problem too large");
    PoolPlan B;
    B.n = 0;
    int blocks = 0, launched = 0;
    for (int g = G.n - 1; g >= 0; --g) {
        if (G.g[g].res < 0 || group_cells[g]) continue;
        PoolGroup &S = B.g[B.n++];
        S = G.g[g];
        S.nchunk = (S.Cg + 255) / 256;
        S.blk0 = blocks;
        blocks += cdiv((long)P.r[S.res].h * P.r[S.res].w * S.nchunk, FP_WARPS);
    }
    const long HW = (long)H * W;
    int n_ident = 0, g_ident = -1;
    for (int g = 0; g < G.n; ++g)
        if (G.g[g].res < 0) { ++n_ident; g_ident = g; }
    // default: one launch, identity blocks spread between the list blocks (WESUP_FP_BWD=split|chunks: the separate kernels)
    const long iblocks = cdiv(cdiv(HW, 32), BC_WARPS);
    if (!force_env && n_ident == 1 && BC.n > 0 && B.n == 0 && iblocks + cblocks < (1L << 30)) {
        BwdAllPlan A;
        A.ident_blocks = (int)iblocks; A.total_blocks = (int)(iblocks + cblocks);
        A.order = x_order;
        if (x_minb == 8)
            fp_pool_bwd_all_kernel<8><<<A.total_blocks, BC_THREADS, 0, stream>>>(L, P, BC, G.g[g_ident], A, grad_pooled, row_labels, counts, HW);
        else
            fp_pool_bwd_all_kernel<6><<<A.total_blocks, BC_THREADS, 0, stream>>>(L, P, BC, G.g[g_ident], A, grad_pooled, row_labels, counts, HW);
        WESUP_CHECK_LAUNCH("wesup_levels_pool_bwd_fp", 1);
        return 0;
    }
    // separate kernels: the identity groups on the library's auxiliary stream (event dependencies only, which also
    // capture into a CUDA graph as parallel branches)
    cudaStream_t s_ident = stream;
    AuxStream *aux = ((B.n > 0 || BC.n > 0) && n_ident > 0) ? aux_stream() : nullptr;
    if (aux && cudaEventRecord(aux->fork, stream) == cudaSuccess && cudaStreamWaitEvent(aux->s, aux->fork, 0) == cudaSuccess)
        s_ident = aux->s;
    for (int g = 0; g < G.n; ++g) {
        if (G.g[g].res >= 0) continue;
        fp_pool_bwd_ident_kernel<<<cdiv(cdiv(HW, 32), 8), 256, 0, s_ident>>>(L, G.g[g], grad_pooled, row_labels, counts, HW);
        ++launched;
    }
    if (BC.n > 0) {
        fp_pool_bwd_cells_kernel<<<(int)cblocks, BC_THREADS, 0, stream>>>(L, P, BC, grad_pooled);
        ++launched;
    }
    if (B.n > 0) {
        fp_pool_bwd_kernel<2><<<blocks, FP_THREADS, 0, stream>>>(L, P, B, grad_pooled);
        ++launched;
    }
    if (s_ident != stream) {
        cudaEventRecord(aux->join, s_ident);
        cudaStreamWaitEvent(stream, aux->join, 0);
    }
    WESUP_CHECK_LAUNCH("wesup_levels_pool_bwd_fp", launched);
    return 0;
}
