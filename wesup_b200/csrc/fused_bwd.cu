// Fused backward of (b) o (a): the adjoint of "bilinear upsample + concat, then
// superpixel mean" evaluated straight from the pooled gradient, without ever
// materialising the (H*W, C) gradient of the hypercolumn.
//
// Reference: autograd of torch.mm(sp_maps, x.t()) (/root/reference/models/wesup.py:284-285)
// followed by autograd of the cat + F.interpolate chain (:254-261).  Because
//     d feat[p, c] = d pooled[row(p), c] / |S_row(p)|
// is constant over a superpixel, the side gradient is
//     d side_l[i, j, c] = sum_{(y,x) in footprint(i,j)} wy(y,i) wx(x,j) * gpn[row(y,x), coff_l + c]
// with gpn = d pooled / count (N x C, 9 MB at 464^2: L2-resident).  HBM traffic
// drops from 2 x 1.82 GB (write + re-read of the dense gradient) to the
// 124 MB of side gradients that must be written anyway.
//
// Deterministic: gather form, fixed summation order, no atomics.  The bilinear
// weights come from per-level tables built on the device with the same fp32
// tap arithmetic as the forward kernel (bilinear_tap), so fwd and bwd stay
// adjoint to rounding.
#include "common.cuh"

namespace wesup {

struct FusedLevels {
    float *dst[WESUP_MAX_LEVELS];
    int C[WESUP_MAX_LEVELS], h[WESUP_MAX_LEVELS], w[WESUP_MAX_LEVELS], coff[WESUP_MAX_LEVELS];
    float sy[WESUP_MAX_LEVELS], sx[WESUP_MAX_LEVELS];
    int ky[WESUP_MAX_LEVELS], kx[WESUP_MAX_LEVELS];          // table row length (max footprint)
    // per level tables inside the workspace
    int32_t *ylo[WESUP_MAX_LEVELS], *xlo[WESUP_MAX_LEVELS];  // first output index of the footprint
    int32_t *yn[WESUP_MAX_LEVELS], *xn[WESUP_MAX_LEVELS];    // footprint length
    float *wy[WESUP_MAX_LEVELS], *wx[WESUP_MAX_LEVELS];      // weights, row stride ky / kx
    int n, H, W, Ctot;
};

static inline size_t up256(size_t x) { return (x + 255) / 256 * 256; }
static inline int table_len(float scale, int out_size) {
    if (!(scale > 0.f)) return out_size;
    int k = (int)(2.0f / scale) + 7;
    return k < out_size + 2 ? k : out_size + 2;
}

__device__ __forceinline__ void footprint_bracket(int i, float scale, int out_size, int &lo, int &hi) {
    if (!(scale > 0.f)) { lo = 0; hi = out_size - 1; return; }
    float inv = 1.0f / scale;
    lo = (int)floorf((float)(i - 1) * inv) - 1;
    hi = (int)ceilf((float)(i + 1) * inv) + 1;
    lo = max(lo, 0);
    hi = min(hi, out_size - 1);
}

// blockIdx.y = 2*level + axis; one thread per low-resolution index
__global__ void fused_tables_kernel(const FusedLevels L) {
    const int l = blockIdx.y >> 1, axis = blockIdx.y & 1;
    const int in_size = axis ? L.w[l] : L.h[l], out_size = axis ? L.W : L.H;
    const float scale = axis ? L.sx[l] : L.sy[l];
    const int K = axis ? L.kx[l] : L.ky[l];
    int32_t *lo_t = axis ? L.xlo[l] : L.ylo[l], *n_t = axis ? L.xn[l] : L.yn[l];
    float *w_t = axis ? L.wx[l] : L.wy[l];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= in_size) return;
    int lo, hi;
    footprint_bracket(i, scale, out_size, lo, hi);
    // trim to the exact support so the hot loop has no zero-weight iterations at the ends
    int first = -1, last = -2;
    for (int d = lo; d <= hi; ++d) {
        Tap t = bilinear_tap(d, scale, in_size);
        if (t.i0 == i || t.i1 == i) { if (first < 0) first = d; last = d; }
    }
    int n = last - first + 1;
    if (first < 0) { first = 0; n = 0; }
    if (n > K) n = K;                                   // cannot happen (K bounds the bracket); keeps writes in range
    lo_t[i] = first;
    n_t[i] = n;
    for (int k = 0; k < n; ++k) {
        Tap t = bilinear_tap(first + k, scale, in_size);
        w_t[(long)i * K + k] = (t.i0 == i ? t.w0 : 0.f) + (t.i1 == i ? t.w1 : 0.f);
    }
}

// gpn[k, :] = grad_pooled[k, :] / counts[k]
__global__ void fused_prescale_kernel(const float *__restrict__ gp, const int32_t *__restrict__ counts, long n4, int C4,
                                      float *__restrict__ gpn) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    int k = (int)(i / C4);
    int cnt = __ldg(counts + k);
    float inv = cnt > 0 ? 1.0f / (float)cnt : 0.f;
    float4 v = __ldg(reinterpret_cast<const float4 *>(gp) + i);
    reinterpret_cast<float4 *>(gpn)[i] = inv * v;
}

// blockIdx.y = level; one thread per (low-res pixel, 4-channel group), channel group fastest
__global__ void __launch_bounds__(256) fused_pool_hyper_bwd_kernel(const FusedLevels L, const float *__restrict__ gpn,
                                                                  const int32_t *__restrict__ row_labels) {
    const int l = blockIdx.y;
    const int Cl = L.C[l], c4n = Cl >> 2, hl = L.h[l], wl = L.w[l];
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)hl * wl * c4n) return;
    const int c = ((int)(idx % c4n)) << 2;
    const long q = idx / c4n;
    const int j = (int)(q % wl), i = (int)(q / wl);
    const float *__restrict__ g = gpn + L.coff[l] + c;
    const int Ctot = L.Ctot, W = L.W;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (hl == L.H && wl == W) {
        int k = __ldg(row_labels + (long)i * W + j);
        if (k >= 0) acc = __ldg(reinterpret_cast<const float4 *>(g + (long)k * Ctot));
    } else {
        const int ylo = __ldg(L.ylo[l] + i), ny = __ldg(L.yn[l] + i);
        const int xlo = __ldg(L.xlo[l] + j), nx = __ldg(L.xn[l] + j);
        const float *__restrict__ wy = L.wy[l] + (long)i * L.ky[l];
        const float *__restrict__ wx = L.wx[l] + (long)j * L.kx[l];
        for (int a = 0; a < ny; ++a) {
            const int32_t *__restrict__ lab = row_labels + (long)(ylo + a) * W + xlo;
            float4 row = make_float4(0.f, 0.f, 0.f, 0.f);
            int cur = -1;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            float wrun = 0.f;                         // weights of a run of equal labels are summed first
            for (int b = 0; b < nx; ++b) {
                const int k = __ldg(lab + b);
                if (k != cur) {
                    fma4(row, wrun, v);
                    wrun = 0.f;
                    cur = k;
                    v = k >= 0 ? __ldg(reinterpret_cast<const float4 *>(g + (long)k * Ctot)) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
                wrun += __ldg(wx + b);
            }
            fma4(row, wrun, v);
            fma4(acc, __ldg(wy + a), row);
        }
    }
    *reinterpret_cast<float4 *>(L.dst[l] + ((long)i * wl + j) * Cl + c) = acc;
}

struct FusedPlan {
    size_t gpn_off, total;
    size_t ylo[WESUP_MAX_LEVELS], yn[WESUP_MAX_LEVELS], wy[WESUP_MAX_LEVELS];
    size_t xlo[WESUP_MAX_LEVELS], xn[WESUP_MAX_LEVELS], wx[WESUP_MAX_LEVELS];
    int ky[WESUP_MAX_LEVELS], kx[WESUP_MAX_LEVELS];
};

static void plan_fused(FusedPlan &P, const int *h, const int *w, int n_levels, int H, int W, long N, long Ctot) {
    size_t off = 0;
    P.gpn_off = off; off += up256(sizeof(float) * (size_t)N * (size_t)Ctot);
    for (int l = 0; l < n_levels; ++l) {
        P.ky[l] = table_len(bilinear_scale(h[l], H), H);
        P.kx[l] = table_len(bilinear_scale(w[l], W), W);
        P.ylo[l] = off; off += up256(sizeof(int32_t) * h[l]);
        P.yn[l] = off;  off += up256(sizeof(int32_t) * h[l]);
        P.wy[l] = off;  off += up256(sizeof(float) * (size_t)h[l] * P.ky[l]);
        P.xlo[l] = off; off += up256(sizeof(int32_t) * w[l]);
        P.xn[l] = off;  off += up256(sizeof(int32_t) * w[l]);
        P.wx[l] = off;  off += up256(sizeof(float) * (size_t)w[l] * P.kx[l]);
    }
    P.total = off;
}

}  // namespace wesup

using namespace wesup;

extern "C" size_t wesup_sp_pool_hypercolumn_bwd_workspace_bytes(const int *C, const int *h, const int *w, int n_levels,
                                                                int H, int W, int N) {
    if (!C || !h || !w || n_levels <= 0 || n_levels > WESUP_MAX_LEVELS || H <= 0 || W <= 0 || N <= 0) return 0;
    long Ctot = 0;
    for (int l = 0; l < n_levels; ++l) {
        if (C[l] <= 0 || h[l] <= 0 || w[l] <= 0) return 0;
        Ctot += C[l];
    }
    FusedPlan P;
    plan_fused(P, h, w, n_levels, H, W, N, Ctot);
    const size_t levels = wesup_levels_pool_bwd_workspace_bytes(C, h, w, n_levels, H, W);   // the footprint kernels' tables
    return P.total > levels ? P.total : levels;
}

extern "C" int wesup_sp_pool_hypercolumn_bwd_walk(const float *grad_pooled, const int32_t *row_labels, const int32_t *counts,
                                             const int *C, const int *h, const int *w, int n_levels, int H, int W, int N,
                                             void *const *grad_side, void *ws, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    WESUP_REQUIRE(grad_pooled && row_labels && counts && C && h && w && grad_side && ws, WESUP_E_ARG,
                  "wesup_sp_pool_hypercolumn_bwd_walk: null pointer");
    WESUP_REQUIRE(n_levels > 0 && n_levels <= WESUP_MAX_LEVELS, WESUP_E_ARG, "wesup_sp_pool_hypercolumn_bwd_walk: n_levels=%d out of range", n_levels);
    WESUP_REQUIRE(H > 0 && W > 0 && N > 0, WESUP_E_ARG, "wesup_sp_pool_hypercolumn_bwd_walk: bad size H=%d W=%d N=%d", H, W, N);
    WESUP_REQUIRE((long)H * W < (1L << 31), WESUP_E_UNSUPPORTED, "wesup_sp_pool_hypercolumn_bwd_walk: H*W must fit int32");
    WESUP_REQUIRE(aligned16(grad_pooled) && aligned16(ws), WESUP_E_ALIGN, "wesup_sp_pool_hypercolumn_bwd_walk: grad_pooled/ws must be 16-byte aligned");
    FusedLevels L;
    L.n = n_levels; L.H = H; L.W = W;
    int off = 0;
    long biggest = 0, in_max = 0;
    for (int l = 0; l < n_levels; ++l) {
        WESUP_REQUIRE(C[l] > 0 && h[l] > 0 && w[l] > 0, WESUP_E_ARG, "wesup_sp_pool_hypercolumn_bwd_walk: level %d has empty shape", l);
        WESUP_REQUIRE(C[l] % 4 == 0, WESUP_E_ALIGN, "wesup_sp_pool_hypercolumn_bwd_walk: C[%d]=%d must be a multiple of 4", l, C[l]);
        WESUP_REQUIRE(grad_side[l] != nullptr && aligned16(grad_side[l]), WESUP_E_ALIGN, "wesup_sp_pool_hypercolumn_bwd_walk: grad_side[%d] null or unaligned", l);
        L.C[l] = C[l]; L.h[l] = h[l]; L.w[l] = w[l]; L.coff[l] = off;
        L.sy[l] = bilinear_scale(h[l], H); L.sx[l] = bilinear_scale(w[l], W);
        L.dst[l] = static_cast<float *>(grad_side[l]);
        off += C[l];
        long n = (long)h[l] * w[l] * (C[l] / 4);
        biggest = biggest > n ? biggest : n;
        in_max = in_max > h[l] ? in_max : h[l];
        in_max = in_max > w[l] ? in_max : w[l];
    }
    L.Ctot = off;
    FusedPlan P;
    plan_fused(P, h, w, n_levels, H, W, N, off);
    char *base = static_cast<char *>(ws);
    float *gpn = reinterpret_cast<float *>(base + P.gpn_off);
    for (int l = 0; l < n_levels; ++l) {
        L.ky[l] = P.ky[l]; L.kx[l] = P.kx[l];
        L.ylo[l] = reinterpret_cast<int32_t *>(base + P.ylo[l]); L.yn[l] = reinterpret_cast<int32_t *>(base + P.yn[l]);
        L.wy[l] = reinterpret_cast<float *>(base + P.wy[l]);
        L.xlo[l] = reinterpret_cast<int32_t *>(base + P.xlo[l]); L.xn[l] = reinterpret_cast<int32_t *>(base + P.xn[l]);
        L.wx[l] = reinterpret_cast<float *>(base + P.wx[l]);
    }
    fused_tables_kernel<<<dim3(cdiv(in_max, 128), 2 * n_levels), 128, 0, stream>>>(L);
    long n4 = (long)N * (off / 4);
    fused_prescale_kernel<<<cdiv(n4, 256), 256, 0, stream>>>(grad_pooled, counts, n4, off / 4, gpn);
    fused_pool_hyper_bwd_kernel<<<dim3(cdiv(biggest, 256), n_levels), 256, 0, stream>>>(L, gpn, row_labels);
    WESUP_CHECK_LAUNCH("wesup_sp_pool_hypercolumn_bwd_walk", 3);
    return 0;
}
