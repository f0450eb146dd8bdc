// Fused forward of (b) o (a): superpixel means of the hypercolumn computed
// straight from the 13 low-resolution side outputs -- the (H*W, C) tensor is
// never written.  Mathematically identical to
// wesup_hypercolumn_fwd followed by wesup_sp_pool_fwd
// (/root/reference/models/wesup.py:254-261 then :284-285); the arithmetic is
// the same in the same order (vertical blend, horizontal blend, sequential sum
// over the superpixel's pixels in ascending order, one multiply by 1/|S|), so
// the result matches the two-kernel fp32 path to rounding of identical
// operations.
//
// Thread = (superpixel row k, 4-channel group).  It walks the row's CSR pixel
// list (raster order inside the superpixel => runs along x) and re-uses the two
// vertically blended source columns of its level while i0 does not change,
// exactly like the row-walk hypercolumn kernel.  Traffic: the side outputs are
// read from L2 (124 MB working set at 464^2), nothing else moves.
#include "common.cuh"

namespace wesup {

__global__ void __launch_bounds__(256) fused_hyper_pool_fwd_kernel(const Levels L, const int32_t *__restrict__ seg_offsets,
                                                                  const int32_t *__restrict__ seg_pixels, int G4, long n_items,
                                                                  float *__restrict__ pooled) {
    const long item = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= n_items) return;
    const int k = (int)(item / G4);
    const int c = ((int)(item - (long)k * G4)) << 2;
    int l = 0;
    while (l + 1 < L.n && c >= L.coff[l + 1]) ++l;
    const int Cl = L.C[l], hl = L.h[l], wl = L.w[l], W = L.W;
    const float *__restrict__ src = L.src[l] + (c - L.coff[l]);
    const bool ident = (hl == L.H && wl == W);
    const float sy = L.sy[l], sx = L.sx[l];
    const int beg = __ldg(seg_offsets + k), end = __ldg(seg_offsets + k + 1);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ident) {
        int i = beg;
        for (; i + 4 <= end; i += 4) {
            const int p0 = __ldg(seg_pixels + i), p1 = __ldg(seg_pixels + i + 1), p2 = __ldg(seg_pixels + i + 2), p3 = __ldg(seg_pixels + i + 3);
            const float4 v0 = __ldg(reinterpret_cast<const float4 *>(src + (long)p0 * Cl));
            const float4 v1 = __ldg(reinterpret_cast<const float4 *>(src + (long)p1 * Cl));
            const float4 v2 = __ldg(reinterpret_cast<const float4 *>(src + (long)p2 * Cl));
            const float4 v3 = __ldg(reinterpret_cast<const float4 *>(src + (long)p3 * Cl));
            acc = acc + v0; acc = acc + v1; acc = acc + v2; acc = acc + v3;
        }
        for (; i < end; ++i) acc = acc + __ldg(reinterpret_cast<const float4 *>(src + (long)__ldg(seg_pixels + i) * Cl));
    } else {
        int prev = -2, y = -1, x = 0, cur = -2;
        Tap ty; ty.i0 = ty.i1 = 0; ty.w0 = 1.f; ty.w1 = 0.f;
        const float *r0 = src, *r1 = src;
        float4 c0 = make_float4(0.f, 0.f, 0.f, 0.f), c1 = c0;
        auto column = [&](int j) {
            const float4 a = __ldg(reinterpret_cast<const float4 *>(r0 + (long)j * Cl));
            const float4 b = __ldg(reinterpret_cast<const float4 *>(r1 + (long)j * Cl));
            float4 r = ty.w0 * a;
            fma4(r, ty.w1, b);
            return r;
        };
        for (int i = beg; i < end; ++i) {
            const int p = __ldg(seg_pixels + i);
            if (p == prev + 1 && x + 1 < W) {
                ++x;
            } else {
                const int ny = p / W;
                x = p - ny * W;
                if (ny != y) {
                    y = ny;
                    ty = bilinear_tap(y, sy, hl);
                    r0 = src + (long)ty.i0 * wl * Cl;
                    r1 = src + (long)ty.i1 * wl * Cl;
                    cur = -2;
                }
            }
            prev = p;
            const Tap tx = bilinear_tap(x, sx, wl);
            if (tx.i0 != cur) {
                c0 = (tx.i0 == cur + 1) ? c1 : column(tx.i0);
                c1 = (tx.i1 != tx.i0) ? column(tx.i1) : c0;
                cur = tx.i0;
            }
            float4 v = tx.w0 * c0;
            fma4(v, tx.w1, c1);
            acc = acc + v;
        }
    }
    const int n = end - beg;
    const float inv = n > 0 ? 1.0f / (float)n : 0.0f;
    *reinterpret_cast<float4 *>(pooled + (long)k * L.Ctot + c) = inv * acc;
}

}  // namespace wesup

using namespace wesup;

extern "C" int wesup_hypercolumn_pool_fwd_walk(const void *const *side, const int *C, const int *h, const int *w, int n_levels,
                                          int H, int W, const int32_t *seg_offsets, const int32_t *seg_pixels, int N,
                                          float *pooled, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    WESUP_REQUIRE(side && C && h && w && seg_offsets && seg_pixels && pooled, WESUP_E_ARG, "wesup_hypercolumn_pool_fwd_walk: null pointer");
    WESUP_REQUIRE(n_levels > 0 && n_levels <= WESUP_MAX_LEVELS, WESUP_E_ARG, "wesup_hypercolumn_pool_fwd_walk: n_levels=%d out of range", n_levels);
    WESUP_REQUIRE(H > 0 && W > 0 && N > 0, WESUP_E_ARG, "wesup_hypercolumn_pool_fwd_walk: bad size H=%d W=%d N=%d", H, W, N);
    WESUP_REQUIRE((long)H * W < (1L << 31), WESUP_E_UNSUPPORTED, "wesup_hypercolumn_pool_fwd_walk: H*W must fit int32");
    WESUP_REQUIRE(aligned16(pooled), WESUP_E_ALIGN, "wesup_hypercolumn_pool_fwd_walk: pooled must be 16-byte aligned");
    Levels L;
    L.n = n_levels; L.H = H; L.W = W;
    int off = 0;
    for (int l = 0; l < n_levels; ++l) {
        WESUP_REQUIRE(C[l] > 0 && h[l] > 0 && w[l] > 0, WESUP_E_ARG, "wesup_hypercolumn_pool_fwd_walk: level %d has empty shape", l);
        WESUP_REQUIRE(C[l] % 4 == 0, WESUP_E_ALIGN, "wesup_hypercolumn_pool_fwd_walk: C[%d]=%d must be a multiple of 4", l, C[l]);
        WESUP_REQUIRE(side[l] != nullptr && aligned16(side[l]), WESUP_E_ALIGN, "wesup_hypercolumn_pool_fwd_walk: side[%d] null or unaligned", l);
        L.src[l] = static_cast<const float *>(side[l]);
        L.dst[l] = nullptr;
        L.C[l] = C[l]; L.h[l] = h[l]; L.w[l] = w[l]; L.coff[l] = off;
        L.sy[l] = bilinear_scale(h[l], H); L.sx[l] = bilinear_scale(w[l], W);
        off += C[l];
    }
    L.Ctot = off;
    const int G4 = off / 4;
    const long n_items = (long)N * G4;
    fused_hyper_pool_fwd_kernel<<<cdiv(n_items, 256), 256, 0, stream>>>(L, seg_offsets, seg_pixels, G4, n_items, pooled);
    WESUP_CHECK_LAUNCH("wesup_hypercolumn_pool_fwd_walk", 1);
    return 0;
}
