// Superpixel means taken DIRECTLY from low-resolution feature levels, and the
// adjoint -- the footprint formulation of (b) o (a).
//
// Reference semantics: bilinear(align_corners=True) upsampling + channel concat
// (/root/reference/models/wesup.py:254-261) followed by torch.mm(sp_maps, x.t())
// (:284-285).  Both are linear, so for every level l and superpixel k
//
//     pooled[k, coff_l + c] = (1/|S_k|) * sum_q  G_k(q) * level_l[q, c]
//     G_k(q) = sum_{(y,x) in S_k} wy(y, q.i) * wx(x, q.j)          (bilinear tap weights)
//
// where q runs over the LOW-resolution cells of the level.  The per-pixel
// formulation (4 taps per pixel per channel: 8448 FMA/pixel at C = 2112) spends
// its time re-deriving the same few cells; here the weights G_k are aggregated
// ONCE per superpixel and scale (scalar work, shared by all channels and by all
// levels of the same resolution) and the channel work is sum_cells G * level:
// ~250 FMA/pixel.  Nothing of size H*W*C is read or written: traffic is the
// levels themselves (read once from HBM, re-read from L2 by neighbouring
// superpixels) plus the (N, C) result.
//
// The same kernels serve two callers: the 13 side outputs (sum C = 2112, the
// reference's hypercolumn channels) and -- "pool first" -- the 13 backbone conv
// outputs (sum C = 4224), after which the 1x1 side convolutions run on N rows
// instead of H*W pixels (mean and 1x1 conv commute).
//
// Deterministic: floating-point sums have a fixed order, the only atomics are integer
// (fixed-point weight sums in shared memory), so results are bit-reproducible run to run.
#include "common.cuh"
#include <limits.h>
#include <math.h>

namespace wesup {

constexpr int PF_THREADS = 256;
constexpr int PF_GRID_CAP = 1024;    // cells of a superpixel's low-res weight grid kept in shared memory

constexpr int PB_WARPS = 8;          // backward: one warp per low-res cell

// consecutive levels of equal resolution: their channels are contiguous in the pooled row
struct Groups {
    int n;
    int l0[WESUP_MAX_LEVELS], l1[WESUP_MAX_LEVELS];
    int h[WESUP_MAX_LEVELS], w[WESUP_MAX_LEVELS];
    float sy[WESUP_MAX_LEVELS], sx[WESUP_MAX_LEVELS];
    int coff[WESUP_MAX_LEVELS], Cg[WESUP_MAX_LEVELS];
    int ident[WESUP_MAX_LEVELS];
    float fscale[WESUP_MAX_LEVELS], finv[WESUP_MAX_LEVELS];   // 2^F and 2^-F of the fixed-point weight sums (0: not usable)
    int blk1[WESUP_MAX_LEVELS];     // bwd: one past the last block of the group inside its launch (coarse groups first)
    // bwd: per-axis footprint tables in the workspace (built by axis_tables_kernel): first output index,
    // number of output indices and their tap weights for every low-resolution row / column of the group
    int32_t *ylo[WESUP_MAX_LEVELS], *yn[WESUP_MAX_LEVELS], *xlo[WESUP_MAX_LEVELS], *xn[WESUP_MAX_LEVELS];
    float *wy[WESUP_MAX_LEVELS], *wx[WESUP_MAX_LEVELS];
    int ky[WESUP_MAX_LEVELS], kx[WESUP_MAX_LEVELS];
};

__device__ __forceinline__ void locate_level(const Levels &L, int l0, int l1, int c, int &l, int &cl) {
    l = l0;
    while (l + 1 < l1 && c >= L.C[l]) { c -= L.C[l]; ++l; }
    cl = c;
}

// ---------------------------------------------------------------------------
// forward: one block per superpixel
//
// Weight aggregation: every pixel adds its 2x2 tap weights wy*wx to the cells of the
// superpixel's low-res bounding boxes (one per resolution group, side by side in shared
// memory) as 32-bit FIXED-POINT sums -- native shared-memory integer atomics.  Integer
// adds commute, so the result does not depend on the order in which threads arrive:
// deterministic without a fixed summation order.  The number of fractional bits is chosen
// per group so that the largest possible cell sum (the cell's footprint area) cannot
// overflow: 27 bits at stride 2 ... 21 bits at stride 16 (fp32 itself keeps 24).
// The non-zero cells are then compacted into per-group lists (cell offset, weight) and
// the channel work streams over a list with independent 128-bit loads in flight.
// ---------------------------------------------------------------------------
struct CellW { int off; float w; };
struct GInfo { int i_lo, j_lo, gw, cells, goff, lbeg, lend; };

__global__ void __launch_bounds__(PF_THREADS) levels_pool_fwd_kernel(const Levels L, const Groups G,
                                                                     const int32_t *__restrict__ seg_offsets,
                                                                     const int32_t *__restrict__ seg_pixels,
                                                                     float *__restrict__ pooled) {
    __shared__ unsigned grid[PF_GRID_CAP];
    __shared__ CellW cellw[PF_GRID_CAP];
    __shared__ float4 part[PF_THREADS];
    __shared__ GInfo gi[WESUP_MAX_LEVELS];
    __shared__ int bbx[2];
    __shared__ int warp_cnt[PF_THREADS / 32];
    __shared__ int total_cells;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int k = blockIdx.x;
    const int beg = __ldg(seg_offsets + k), end = __ldg(seg_offsets + k + 1);
    const int n = end - beg;
    float *__restrict__ out = pooled + (long)k * L.Ctot;
    if (n <= 0) {
        for (int c = tid * 4; c < L.Ctot; c += PF_THREADS * 4) *reinterpret_cast<float4 *>(out + c) = make_float4(0.f, 0.f, 0.f, 0.f);
        return;
    }
    const int W = L.W;
    // bounding box: pixel ids ascend inside a row of the CSR, so rows come from the two ends
    if (tid == 0) { bbx[0] = INT_MAX; bbx[1] = -1; }
    __syncthreads();
    {
        int xmin = INT_MAX, xmax = -1;
        for (int i = beg + tid; i < end; i += PF_THREADS) {
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
        for (int g = 0; g < G.n; ++g) {
            GInfo q;
            q.i_lo = q.j_lo = 0; q.gw = 1; q.cells = 0; q.goff = off; q.lbeg = q.lend = 0;
            if (!G.ident[g] && G.fscale[g] > 0.f) {
                q.i_lo = bilinear_tap(ymin, G.sy[g], G.h[g]).i0;
                q.j_lo = bilinear_tap(xmin, G.sx[g], G.w[g]).i0;
                const long gh = bilinear_tap(ymax, G.sy[g], G.h[g]).i1 - q.i_lo + 1;
                q.gw = bilinear_tap(xmax, G.sx[g], G.w[g]).i1 - q.j_lo + 1;
                if (off + gh * q.gw <= PF_GRID_CAP) { q.cells = (int)gh * q.gw; off += q.cells; }
            }
            gi[g] = q;
        }
        total_cells = off;
    }
    __syncthreads();
    const int total = total_cells;
    for (int e = tid; e < total; e += PF_THREADS) grid[e] = 0u;
    __syncthreads();
    for (int it = beg + tid; it < end; it += PF_THREADS) {
        const int p = __ldg(seg_pixels + it);
        const int y = p / W, x = p - y * W;
        for (int g = 0; g < G.n; ++g) {
            const GInfo q = gi[g];
            if (q.cells == 0) continue;
            const Tap ty = bilinear_tap(y, G.sy[g], G.h[g]), tx = bilinear_tap(x, G.sx[g], G.w[g]);
            unsigned *gp = grid + q.goff;
            const int r0 = (ty.i0 - q.i_lo) * q.gw, r1 = (ty.i1 - q.i_lo) * q.gw, c0 = tx.i0 - q.j_lo, c1 = tx.i1 - q.j_lo;
            const float fs = G.fscale[g];
            const float a0 = ty.w0 * fs, a1 = ty.w1 * fs;            // exact: fs is a power of two
            atomicAdd(gp + r0 + c0, __float2uint_rn(a0 * tx.w0));
            atomicAdd(gp + r0 + c1, __float2uint_rn(a0 * tx.w1));
            atomicAdd(gp + r1 + c0, __float2uint_rn(a1 * tx.w0));
            atomicAdd(gp + r1 + c1, __float2uint_rn(a1 * tx.w1));
        }
    }
    __syncthreads();
    // compact the non-zero cells, group after group (order: cell index -- fixed)
    int n_list = 0;
    for (int e0 = 0; e0 < total; e0 += PF_THREADS) {
        const int e = e0 + tid;
        const unsigned v = e < total ? grid[e] : 0u;
        const unsigned m = __ballot_sync(0xffffffffu, v != 0u);
        if (lane == 0) warp_cnt[wid] = __popc(m);
        __syncthreads();
        int before = n_list, tot = n_list;
#pragma unroll
        for (int q = 0; q < PF_THREADS / 32; ++q) {
            const int c = warp_cnt[q];
            if (q < wid) before += c;
            tot += c;
        }
        if (e < total) {
            int g = 0;
            while (e >= gi[g].goff + gi[g].cells) ++g;
            const int pos = before + __popc(m & ((1u << lane) - 1u));
            const int local = e - gi[g].goff, gw = gi[g].gw;
            if (local == 0) gi[g].lbeg = pos;
            if (v != 0u) {
                const int i = __float2int_rd(((float)local + 0.5f) / (float)gw), j = local - i * gw;
                CellW cw;
                cw.off = i * G.w[g] + j;
                cw.w = (float)v * G.finv[g];
                cellw[pos] = cw;
            }
        }
        n_list = tot;
        __syncthreads();
    }
    if (tid == 0) {
        int nxt = n_list;
        for (int g = G.n - 1; g >= 0; --g)
            if (gi[g].cells > 0) { gi[g].lend = nxt; nxt = gi[g].lbeg; }
    }
    __syncthreads();
    const float inv = 1.0f / (float)n;

    for (int g = 0; g < G.n; ++g) {
        const int nch4 = G.Cg[g] >> 2;
        const int hl = G.h[g], wl = G.w[g];
        const bool ident = G.ident[g] != 0;
        const GInfo q = gi[g];
        const bool use_grid = q.cells > 0;
        // ---- channel work: thread = (slice of the items, 4-channel group) ----------------
        for (int cb = 0; cb < nch4; cb += PF_THREADS) {
            const int active = min(nch4 - cb, PF_THREADS);
            const int nsl = PF_THREADS / active;
            const int sl = tid / active;
            const int c4 = cb + (tid - sl * active);
            const bool live = sl < nsl;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            if (live) {
                int l, cl;
                locate_level(L, G.l0[g], G.l1[g], c4 << 2, l, cl);
                const int Cl = L.C[l];
                const float *__restrict__ src = L.src[l] + cl;
                if (ident) {
                    int it = beg + sl;
                    for (; it + 3 * nsl < end; it += 4 * nsl) {
                        const int p0 = __ldg(seg_pixels + it), p1 = __ldg(seg_pixels + it + nsl), p2 = __ldg(seg_pixels + it + 2 * nsl),
                                  p3 = __ldg(seg_pixels + it + 3 * nsl);
                        const float4 v0 = __ldg(reinterpret_cast<const float4 *>(src + (long)p0 * Cl));
                        const float4 v1 = __ldg(reinterpret_cast<const float4 *>(src + (long)p1 * Cl));
                        const float4 v2 = __ldg(reinterpret_cast<const float4 *>(src + (long)p2 * Cl));
                        const float4 v3 = __ldg(reinterpret_cast<const float4 *>(src + (long)p3 * Cl));
                        acc = acc + v0; acc = acc + v1; acc = acc + v2; acc = acc + v3;
                    }
                    for (; it < end; it += nsl) acc = acc + __ldg(reinterpret_cast<const float4 *>(src + (long)__ldg(seg_pixels + it) * Cl));
                } else if (use_grid) {
                    const float *__restrict__ base = src + ((long)q.i_lo * wl + q.j_lo) * Cl;
                    int e = q.lbeg + sl;
                    for (; e + 3 * nsl < q.lend; e += 4 * nsl) {
                        const CellW a0 = cellw[e], a1 = cellw[e + nsl], a2 = cellw[e + 2 * nsl], a3 = cellw[e + 3 * nsl];
                        const float4 v0 = __ldg(reinterpret_cast<const float4 *>(base + (long)a0.off * Cl));
                        const float4 v1 = __ldg(reinterpret_cast<const float4 *>(base + (long)a1.off * Cl));
                        const float4 v2 = __ldg(reinterpret_cast<const float4 *>(base + (long)a2.off * Cl));
                        const float4 v3 = __ldg(reinterpret_cast<const float4 *>(base + (long)a3.off * Cl));
                        fma4(acc, a0.w, v0); fma4(acc, a1.w, v1); fma4(acc, a2.w, v2); fma4(acc, a3.w, v3);
                    }
                    for (; e < q.lend; e += nsl) {
                        const CellW a0 = cellw[e];
                        fma4(acc, a0.w, __ldg(reinterpret_cast<const float4 *>(base + (long)a0.off * Cl)));
                    }
                } else {
                    // bounding box too large for the shared grid (huge / scattered superpixel): per-pixel taps
                    const float sy = G.sy[g], sx = G.sx[g];
                    for (int it = beg + sl; it < end; it += nsl) {
                        const int p = __ldg(seg_pixels + it);
                        const int y = p / W, x = p - y * W;
                        const Tap ty = bilinear_tap(y, sy, hl), tx = bilinear_tap(x, sx, wl);
                        const float *r0 = src + (long)ty.i0 * wl * Cl, *r1 = src + (long)ty.i1 * wl * Cl;
                        float4 a = ty.w0 * __ldg(reinterpret_cast<const float4 *>(r0 + (long)tx.i0 * Cl));
                        fma4(a, ty.w1, __ldg(reinterpret_cast<const float4 *>(r1 + (long)tx.i0 * Cl)));
                        float4 b = ty.w0 * __ldg(reinterpret_cast<const float4 *>(r0 + (long)tx.i1 * Cl));
                        fma4(b, ty.w1, __ldg(reinterpret_cast<const float4 *>(r1 + (long)tx.i1 * Cl)));
                        fma4(acc, tx.w0, a);
                        fma4(acc, tx.w1, b);
                    }
                }
            }
            if (nsl > 1) {                      // block-uniform
                part[tid] = acc;
                __syncthreads();
                if (sl == 0) {
                    for (int s_ = 1; s_ < nsl; ++s_) acc = acc + part[s_ * active + tid];
                }
            }
            if (live && sl == 0) *reinterpret_cast<float4 *>(out + G.coff[g] + (c4 << 2)) = inv * acc;
            if (nsl > 1) __syncthreads();
        }
    }
}

// ---------------------------------------------------------------------------
// backward: one warp per low-resolution cell q
//   grad_level_l[q, c] = sum_k G_k(q) / |S_k| * grad_pooled[k, coff_l + c]
// The warp reads the labels of the cell's high-resolution footprint ONCE and folds the
// tap weights into a 64-slot hash table keyed by superpixel (per-warp, shared memory):
// the key is claimed with an atomicCAS, the weight is added as a 32-bit fixed-point
// integer -- commutative, so the table's CONTENT does not depend on thread order.  The
// occupied slots are then ranked by key (the slot a key lands in does depend on the
// order), which gives a list sorted by superpixel id: the floating-point gather that
// follows has a fixed order and the result is bit-reproducible.  All loads of a list
// entry (count, pooled-gradient row) are independent and issued together.
// A footprint that meets more than 64 superpixels takes the slow path: distinct labels
// visited in ascending order, the footprint re-read for each.
// ---------------------------------------------------------------------------
struct Staged { int lab; float w; };
constexpr int PB_SLOTS = 64;

__device__ __forceinline__ float tap_weight(int dst, int cell, float scale, int in_size) {
    const Tap t = bilinear_tap(dst, scale, in_size);
    return (t.i0 == cell ? t.w0 : 0.f) + (t.i1 == cell ? t.w1 : 0.f);
}

// blockIdx.y = 2 * group + axis; one thread per low-resolution index: the exact support of the
// cell among the output indices (bracket from the inverse map, trimmed with the forward's own tap
// arithmetic) and the tap weights inside it
__global__ void axis_tables_kernel(const Groups G, int H, int W) {
    const int g = blockIdx.y >> 1, axis = blockIdx.y & 1;
    if (G.ident[g]) return;
    const int in_size = axis ? G.w[g] : G.h[g], out_size = axis ? W : H;
    const float scale = axis ? G.sx[g] : G.sy[g];
    const int K = axis ? G.kx[g] : G.ky[g];
    int32_t *lo_t = axis ? G.xlo[g] : G.ylo[g], *n_t = axis ? G.xn[g] : G.yn[g];
    float *w_t = axis ? G.wx[g] : G.wy[g];
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

template <int V>
__global__ void __launch_bounds__(PB_WARPS * 32, V <= 2 ? 5 : (V <= 6 ? 3 : 2)) levels_pool_bwd_kernel(const Levels L, const Groups G, int g_lo, int g_hi,
                                                                        const float *__restrict__ gp,
                                                                        const int32_t *__restrict__ row_labels,
                                                                        const int32_t *__restrict__ counts) {
    __shared__ int keys_all[PB_WARPS][PB_SLOTS];
    __shared__ unsigned wfix_all[PB_WARPS][PB_SLOTS];
    __shared__ Staged list_all[PB_WARPS][PB_SLOTS];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int g = g_hi - 1;                                   // coarse groups own the first blocks
    while (g > g_lo && (int)blockIdx.x >= G.blk1[g]) --g;
    const int blk = (int)blockIdx.x - (g == g_hi - 1 ? 0 : G.blk1[g + 1]);
    const int hg = G.h[g], wg = G.w[g];
    const long q = (long)blk * PB_WARPS + wid;
    if (q >= (long)hg * wg) return;
    const int nch4 = G.Cg[g] >> 2, Ctot = L.Ctot, W = L.W;
    const float *__restrict__ gpg = gp + G.coff[g];
    int *keys = keys_all[wid];
    unsigned *wfix = wfix_all[wid];
    Staged *list = list_all[wid];
    float4 acc[V];
#pragma unroll
    for (int v = 0; v < V; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int i = (int)(q / wg), j = (int)(q - (long)i * wg);
    const int ylo = __ldg(G.ylo[g] + i), xlo = __ldg(G.xlo[g] + j);
    const int nx = __ldg(G.xn[g] + j), nf = __ldg(G.yn[g] + i) * nx;
    const float *__restrict__ wyt = G.wy[g] + (long)i * G.ky[g];
    const float *__restrict__ wxt = G.wx[g] + (long)j * G.kx[g];
    const float inv_nx = 1.0f / (float)nx;
    auto fetch = [&](int t) {
        const int a = __float2int_rd(((float)t + 0.5f) * inv_nx), b = t - a * nx;
        Staged s;
        s.w = __ldg(wyt + a) * __ldg(wxt + b);
        s.lab = (s.w != 0.f) ? __ldg(row_labels + (long)(ylo + a) * W + xlo + b) : -1;
        return s;
    };
    // gather the rows of the listed superpixels: the count and the row of an entry depend only on its
    // id, so all of them are issued together (the weight meets the count at the FMA)
    auto flush = [&](int nl) {
        __syncwarp();
        constexpr int U = V >= 6 ? 1 : 2;
        int e = 0;
        for (; e + U <= nl; e += U) {
            float4 val[U][V];
            float wv[U];
            int cnt[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const Staged en = list[e + u];
                wv[u] = en.w;
                cnt[u] = __ldg(counts + en.lab);
                const float4 *__restrict__ row = reinterpret_cast<const float4 *>(gpg + (long)en.lab * Ctot);
#pragma unroll
                for (int v = 0; v < V; ++v) {
                    const int c4 = lane + 32 * v;
                    val[u][v] = c4 < nch4 ? __ldg(row + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const float wq = cnt[u] > 0 ? wv[u] / (float)cnt[u] : 0.f;
#pragma unroll
                for (int v = 0; v < V; ++v) fma4(acc[v], wq, val[u][v]);
            }
        }
        for (; e < nl; ++e) {
            const Staged en = list[e];
            const int cnt = __ldg(counts + en.lab);
            const float4 *__restrict__ row = reinterpret_cast<const float4 *>(gpg + (long)en.lab * Ctot);
            float4 val[V];
#pragma unroll
            for (int v = 0; v < V; ++v) {
                const int c4 = lane + 32 * v;
                val[v] = c4 < nch4 ? __ldg(row + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            const float wq = cnt > 0 ? en.w / (float)cnt : 0.f;
#pragma unroll
            for (int v = 0; v < V; ++v) fma4(acc[v], wq, val[v]);
        }
        __syncwarp();
    };
    // ---- fold the footprint into the hash table -------------------------------------------
    keys[lane] = -1; keys[lane + 32] = -1;
    wfix[lane] = 0u; wfix[lane + 32] = 0u;
    __syncwarp();
    const float fs = G.fscale[g];
    bool overflow = !(fs > 0.f);
    if (!overflow) {
        auto insert = [&](const Staged &s_) {
            if (s_.lab < 0) return;
            unsigned h = ((unsigned)s_.lab * 2654435761u) >> 26;
            int probes = 0;
            for (; probes < PB_SLOTS; ++probes) {
                const int old = atomicCAS(&keys[h], -1, s_.lab);
                if (old == -1 || old == s_.lab) { atomicAdd(&wfix[h], __float2uint_rn(s_.w * fs)); break; }
                h = (h + 1) & (PB_SLOTS - 1);
            }
            if (probes == PB_SLOTS) overflow = true;
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
    if (!overflow) {
        // rank the occupied slots by key -> list sorted by superpixel id
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
        const float finv = G.finv[g];
        if (k0 >= 0) { list[r0].lab = k0; list[r0].w = (float)wfix[lane] * finv; }
        if (k1 >= 0) { list[r1].lab = k1; list[r1].w = (float)wfix[lane + 32] * finv; }
        flush(__popc(occ0) + __popc(occ1));
    } else {
        int cur = INT_MAX;
        for (int t = lane; t < nf; t += 32) {
            const Staged s_ = fetch(t);
            if (s_.lab >= 0) cur = min(cur, s_.lab);
        }
        cur = __reduce_min_sync(0xffffffffu, cur);
        int nl = 0;
        while (cur != INT_MAX) {
            float ws = 0.f;
            int nxt = INT_MAX;
            for (int t = lane; t < nf; t += 32) {
                const Staged s_ = fetch(t);
                if (s_.lab == cur) ws += s_.w;
                else if (s_.lab > cur) nxt = min(nxt, s_.lab);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) ws += __shfl_xor_sync(0xffffffffu, ws, o);
            nxt = __reduce_min_sync(0xffffffffu, nxt);
            if (lane == 0) { list[nl].lab = cur; list[nl].w = ws; }
            if (++nl == PB_SLOTS) { flush(nl); nl = 0; }
            cur = nxt;
        }
        flush(nl);
    }
#pragma unroll
    for (int v = 0; v < V; ++v) {
        const int c4 = lane + 32 * v;
        if (c4 < nch4) {
            int l, cl;
            locate_level(L, G.l0[g], G.l1[g], c4 << 2, l, cl);
            *reinterpret_cast<float4 *>(L.dst[l] + q * L.C[l] + cl) = acc[v];
        }
    }
}

// full-resolution levels: grad[p, c] = grad_pooled[row(p), c] / |S_row(p)|; four pixels per thread so
// that the label -> count -> row chain of four pixels is in flight together
__global__ void __launch_bounds__(256) levels_pool_bwd_ident_kernel(const Levels L, const Groups G, int g,
                                                                    const float *__restrict__ gp,
                                                                    const int32_t *__restrict__ row_labels,
                                                                    const int32_t *__restrict__ counts, long HW) {
    const int nch4 = G.Cg[g] >> 2, Ctot = L.Ctot;
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long quad = idx / nch4;
    const int c4 = (int)(idx - quad * nch4);
    const long p0 = quad * 4;
    if (p0 >= HW) return;
    int lab[4];
    if (p0 + 3 < HW && (reinterpret_cast<uintptr_t>(row_labels) & 15u) == 0) {
        const int4 t = __ldg(reinterpret_cast<const int4 *>(row_labels + p0));
        lab[0] = t.x; lab[1] = t.y; lab[2] = t.z; lab[3] = t.w;
    } else {
#pragma unroll
        for (int u = 0; u < 4; ++u) lab[u] = p0 + u < HW ? __ldg(row_labels + p0 + u) : -1;
    }
    int cnt[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) cnt[u] = lab[u] >= 0 ? __ldg(counts + lab[u]) : 0;
    float4 val[4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
        val[u] = lab[u] >= 0 ? __ldg(reinterpret_cast<const float4 *>(gp + G.coff[g] + (long)lab[u] * Ctot) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
    int l, cl;
    locate_level(L, G.l0[g], G.l1[g], c4 << 2, l, cl);
    float *__restrict__ dst = L.dst[l] + cl;
    const int Cl = L.C[l];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        if (p0 + u < HW) {
            const float wv = cnt[u] > 0 ? 1.0f / (float)cnt[u] : 0.f;
            *reinterpret_cast<float4 *>(dst + (p0 + u) * Cl) = wv * val[u];
        }
    }
}

constexpr int PB_VMAX = 12;          // float4 accumulators per lane: a group carries at most 32*4*12 = 1536 channels

// group consecutive levels of equal resolution; a group never exceeds PB_VMAX*128 channels
static int build_groups(Groups &G, const Levels &L) {
    G.n = 0;
    for (int l = 0; l < L.n; ++l) {
        if (L.C[l] > PB_VMAX * 128) return -1;
        const int g = G.n - 1;
        if (g >= 0 && G.h[g] == L.h[l] && G.w[g] == L.w[l] && G.Cg[g] + L.C[l] <= PB_VMAX * 128) {
            G.l1[g] = l + 1;
            G.Cg[g] += L.C[l];
        } else {
            const int n = G.n++;
            G.l0[n] = l; G.l1[n] = l + 1;
            G.h[n] = L.h[l]; G.w[n] = L.w[l];
            G.sy[n] = L.sy[l]; G.sx[n] = L.sx[l];
            G.coff[n] = L.coff[l]; G.Cg[n] = L.C[l];
            G.ident[n] = (L.h[l] == L.H && L.w[l] == L.W) ? 1 : 0;
            // fixed-point format of the forward's weight sums: a cell's sum is at most its footprint area
            const double fy = L.sy[l] > 0.f ? 2.0 / L.sy[l] + 2.0 : (double)L.H;
            const double fx = L.sx[l] > 0.f ? 2.0 / L.sx[l] + 2.0 : (double)L.W;
            const int bits = 31 - (int)ceil(log2(fy * fx + 1.0));
            G.fscale[n] = bits >= 16 ? (float)ldexp(1.0, bits) : 0.f;
            G.finv[n] = bits >= 16 ? (float)ldexp(1.0, -bits) : 0.f;
        }
    }
    return 0;
}

static int fill_pool_levels(Levels &L, const char *who, const void *const *ptrs, bool is_dst, const int *C, const int *h, const int *w,
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

static inline size_t up256(size_t x) { return (x + 255) / 256 * 256; }
static inline int table_len(float scale, int out_size) {
    if (!(scale > 0.f)) return out_size;
    const int k = (int)(2.0f / scale) + 7;
    return k < out_size + 2 ? k : out_size + 2;
}
// carve the per-axis tables of every non-identity group out of `ws` (ws == nullptr: size only)
static size_t plan_tables(Groups &G, int H, int W, char *ws) {
    size_t off = 0;
    auto take = [&](size_t bytes) { char *p = ws ? ws + off : nullptr; off += up256(bytes); return p; };
    for (int g = 0; g < G.n; ++g) {
        G.ky[g] = G.kx[g] = 0;
        if (G.ident[g]) continue;
        G.ky[g] = table_len(G.sy[g], H);
        G.kx[g] = table_len(G.sx[g], W);
        G.ylo[g] = (int32_t *)take(sizeof(int32_t) * G.h[g]);
        G.yn[g] = (int32_t *)take(sizeof(int32_t) * G.h[g]);
        G.wy[g] = (float *)take(sizeof(float) * (size_t)G.h[g] * G.ky[g]);
        G.xlo[g] = (int32_t *)take(sizeof(int32_t) * G.w[g]);
        G.xn[g] = (int32_t *)take(sizeof(int32_t) * G.w[g]);
        G.wx[g] = (float *)take(sizeof(float) * (size_t)G.w[g] * G.kx[g]);
    }
    return off > 256 ? off : 256;
}

// The backward is several latency-bound launches of very different shapes (few heavy cells at the
// coarse resolutions, many light ones at the fine ones).  They are independent, so they are forked
// onto auxiliary streams and joined back into the caller's stream -- event dependencies only, which
// also capture into a CUDA graph as parallel branches.  The auxiliary streams belong to the library
// (created once per process; calls from several host threads merely share them).
constexpr int PB_AUX = 3;
struct AuxStreams {
    cudaStream_t s[PB_AUX];
    cudaEvent_t fork, join[PB_AUX];
    bool ok = false;
};
static AuxStreams *aux_streams() {
    static AuxStreams a;
    static bool tried = false;
    if (!tried) {
        tried = true;
        bool ok = cudaEventCreateWithFlags(&a.fork, cudaEventDisableTiming) == cudaSuccess;
        for (int i = 0; ok && i < PB_AUX; ++i)
            ok = cudaStreamCreateWithFlags(&a.s[i], cudaStreamNonBlocking) == cudaSuccess &&
                 cudaEventCreateWithFlags(&a.join[i], cudaEventDisableTiming) == cudaSuccess;
        a.ok = ok;
        if (!ok) cudaGetLastError();
    }
    return a.ok ? &a : nullptr;
}

template <int V>
static void launch_bwd(const Levels &L, Groups &G, int g_lo, int g_hi, const float *gp, const int32_t *row_labels,
                       const int32_t *counts, cudaStream_t stream) {
    int blocks = 0;
    for (int g = g_hi - 1; g >= g_lo; --g) {            // coarse groups (longest footprint walks) own the first blocks
        blocks += cdiv((long)G.h[g] * G.w[g], PB_WARPS);
        G.blk1[g] = blocks;
    }
    levels_pool_bwd_kernel<V><<<blocks, PB_WARPS * 32, 0, stream>>>(L, G, g_lo, g_hi, gp, row_labels, counts);
}

static inline int bwd_slots(int Cg) {                   // float4 accumulators per lane the group needs
    const int v = (Cg / 4 + 31) / 32;
    return v <= 1 ? 1 : v <= 2 ? 2 : v <= 4 ? 4 : v <= 6 ? 6 : PB_VMAX;
}

}  // namespace wesup

using namespace wesup;

extern "C" int wesup_levels_pool_fwd(const void *const *level, const int *C, const int *h, const int *w, int n_levels, int H,
                                     int W, const int32_t *seg_offsets, const int32_t *seg_pixels, int N, float *pooled,
                                     void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    WESUP_REQUIRE(level && C && h && w && seg_offsets && seg_pixels && pooled, WESUP_E_ARG, "wesup_levels_pool_fwd: null pointer");
    WESUP_REQUIRE(n_levels > 0 && n_levels <= WESUP_MAX_LEVELS, WESUP_E_ARG, "wesup_levels_pool_fwd: n_levels=%d out of range", n_levels);
    WESUP_REQUIRE(H > 0 && W > 0 && N > 0, WESUP_E_ARG, "wesup_levels_pool_fwd: bad size H=%d W=%d N=%d", H, W, N);
    WESUP_REQUIRE((long)H * W < (1L << 31), WESUP_E_UNSUPPORTED, "wesup_levels_pool_fwd: H*W must fit int32");
    WESUP_REQUIRE(H < 65536 && W < 65536, WESUP_E_UNSUPPORTED, "wesup_levels_pool_fwd: H and W must be below 65536");
    WESUP_REQUIRE(aligned16(pooled), WESUP_E_ALIGN, "wesup_levels_pool_fwd: pooled must be 16-byte aligned");
    Levels L;
    int rc = fill_pool_levels(L, "wesup_levels_pool_fwd", level, false, C, h, w, n_levels, H, W);
    if (rc) return rc;
    Groups G;
    WESUP_REQUIRE(build_groups(G, L) == 0, WESUP_E_UNSUPPORTED, "wesup_levels_pool_fwd: a level has more than %d channels", PB_VMAX * 128);
    levels_pool_fwd_kernel<<<N, PF_THREADS, 0, stream>>>(L, G, seg_offsets, seg_pixels, pooled);
    WESUP_CHECK_LAUNCH("wesup_levels_pool_fwd", 1);
    return 0;
}

extern "C" size_t wesup_levels_pool_bwd_workspace_bytes(const int *C, const int *h, const int *w, int n_levels, int H, int W) {
    if (!C || !h || !w || n_levels <= 0 || n_levels > WESUP_MAX_LEVELS || H <= 0 || W <= 0) return 0;
    Levels L;
    L.n = n_levels; L.H = H; L.W = W;
    int off = 0;
    for (int l = 0; l < n_levels; ++l) {
        if (C[l] <= 0 || h[l] <= 0 || w[l] <= 0 || C[l] > PB_VMAX * 128) return 0;
        L.C[l] = C[l]; L.h[l] = h[l]; L.w[l] = w[l]; L.coff[l] = off;
        L.sy[l] = bilinear_scale(h[l], H); L.sx[l] = bilinear_scale(w[l], W);
        off += C[l];
    }
    L.Ctot = off;
    Groups G;
    if (build_groups(G, L) != 0) return 0;
    return plan_tables(G, H, W, nullptr);
}

extern "C" int wesup_levels_pool_bwd(const float *grad_pooled, const int32_t *row_labels, const int32_t *counts, const int *C,
                                     const int *h, const int *w, int n_levels, int H, int W, int N, void *const *grad_level,
                                     void *ws, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    WESUP_REQUIRE(grad_pooled && row_labels && counts && C && h && w && grad_level && ws, WESUP_E_ARG, "wesup_levels_pool_bwd: null pointer");
    WESUP_REQUIRE(aligned16(ws), WESUP_E_ALIGN, "wesup_levels_pool_bwd: ws must be 16-byte aligned");
    WESUP_REQUIRE(n_levels > 0 && n_levels <= WESUP_MAX_LEVELS, WESUP_E_ARG, "wesup_levels_pool_bwd: n_levels=%d out of range", n_levels);
    WESUP_REQUIRE(H > 0 && W > 0 && N > 0, WESUP_E_ARG, "wesup_levels_pool_bwd: bad size H=%d W=%d N=%d", H, W, N);
    WESUP_REQUIRE((long)H * W < (1L << 31), WESUP_E_UNSUPPORTED, "wesup_levels_pool_bwd: H*W must fit int32");
    WESUP_REQUIRE(aligned16(grad_pooled), WESUP_E_ALIGN, "wesup_levels_pool_bwd: grad_pooled must be 16-byte aligned");
    Levels L;
    int rc = fill_pool_levels(L, "wesup_levels_pool_bwd", grad_level, true, C, h, w, n_levels, H, W);
    if (rc) return rc;
    Groups G;
    WESUP_REQUIRE(build_groups(G, L) == 0, WESUP_E_UNSUPPORTED, "wesup_levels_pool_bwd: a level has more than %d channels", PB_VMAX * 128);
    plan_tables(G, H, W, static_cast<char *>(ws));
    int in_max = 1, launched = 0;
    bool any_table = false;
    for (int g = 0; g < G.n; ++g)
        if (!G.ident[g]) { any_table = true; in_max = in_max > G.h[g] ? in_max : G.h[g]; in_max = in_max > G.w[g] ? in_max : G.w[g]; }
    if (any_table) {
        axis_tables_kernel<<<dim3(cdiv(in_max, 128), 2 * G.n), 128, 0, stream>>>(G, H, W);
        ++launched;
    }
    // non-identity groups: consecutive groups with the same accumulator width share one launch (coarse
    // first); launch k runs on auxiliary stream k-1 (the first one and anything beyond the pool on `stream`)
    AuxStreams *aux = aux_streams();
    if (aux && cudaEventRecord(aux->fork, stream) != cudaSuccess) aux = nullptr;
    int n_forked = 0, k_launch = 0;
    for (int g_hi = G.n; g_hi > 0;) {
        if (G.ident[g_hi - 1]) { --g_hi; continue; }
        const int v = bwd_slots(G.Cg[g_hi - 1]);
        int g_lo = g_hi - 1;
        while (g_lo > 0 && !G.ident[g_lo - 1] && bwd_slots(G.Cg[g_lo - 1]) == v) --g_lo;
        cudaStream_t s_launch = stream;
        if (aux && k_launch > 0 && n_forked < PB_AUX) {
            s_launch = aux->s[n_forked++];
            cudaStreamWaitEvent(s_launch, aux->fork, 0);
        }
        ++k_launch;
        if (v == 1) launch_bwd<1>(L, G, g_lo, g_hi, grad_pooled, row_labels, counts, s_launch);
        else if (v == 2) launch_bwd<2>(L, G, g_lo, g_hi, grad_pooled, row_labels, counts, s_launch);
        else if (v == 4) launch_bwd<4>(L, G, g_lo, g_hi, grad_pooled, row_labels, counts, s_launch);
        else if (v == 6) launch_bwd<6>(L, G, g_lo, g_hi, grad_pooled, row_labels, counts, s_launch);
        else launch_bwd<PB_VMAX>(L, G, g_lo, g_hi, grad_pooled, row_labels, counts, s_launch);
        ++launched;
        g_hi = g_lo;
    }
    for (int g = 0; g < G.n; ++g) {
        if (!G.ident[g]) continue;
        const long HW = (long)H * W;
        const long items = ((HW + 3) / 4) * (G.Cg[g] / 4);
        levels_pool_bwd_ident_kernel<<<cdiv(items, 256), 256, 0, stream>>>(L, G, g, grad_pooled, row_labels, counts, HW);
        ++launched;
    }
    for (int i = 0; i < n_forked; ++i) {
        cudaEventRecord(aux->join[i], aux->s[i]);
        cudaStreamWaitEvent(stream, aux->join[i], 0);
    }
    WESUP_CHECK_LAUNCH("wesup_levels_pool_bwd", launched);
    return 0;
}

// The historical entry points of the fused path keep their signatures and now run the
// footprint kernels; the per-pixel walk kernels stay exported as *_walk (cross-checks, benches).
extern "C" int wesup_hypercolumn_pool_fwd(const void *const *side, const int *C, const int *h, const int *w, int n_levels,
                                          int H, int W, const int32_t *seg_offsets, const int32_t *seg_pixels, int N,
                                          float *pooled, void *stream) {
    return wesup_levels_pool_fwd(side, C, h, w, n_levels, H, W, seg_offsets, seg_pixels, N, pooled, stream);
}

extern "C" int wesup_sp_pool_hypercolumn_bwd(const float *grad_pooled, const int32_t *row_labels, const int32_t *counts,
                                             const int *C, const int *h, const int *w, int n_levels, int H, int W, int N,
                                             void *const *grad_side, void *ws, void *stream) {
    return wesup_levels_pool_bwd(grad_pooled, row_labels, counts, C, h, w, n_levels, H, W, N, grad_side, ws, stream);
}
