// (a) Hypercolumn construction: bilinear (align_corners=True) upsampling of the
// 13 VGG16 side outputs to the input size, concatenated along channels, written
// ONCE.  Replaces WESUP._hook_fn's interpolate + 12 incremental torch.cat calls
// (/root/reference/models/wesup.py:246-261), which re-copy the growing tensor
// (5x the final bytes).
//
// Layouts: pixel-major (H*W, C) is the primary one -- a warp writes 512
// contiguous bytes of one pixel (128-bit stores, full cache lines) and the
// low-resolution taps are contiguous channel vectors served from L1/L2.
// Channel-major (C,H,W) (the reference's layout) is also provided.
//
// Backward is the exact adjoint in gather form (each low-resolution element
// sums its footprint in a fixed order): deterministic, no atomics.
#include <stdlib.h>
#include <math.h>
#include "common.cuh"

namespace wesup {


constexpr int TILE_W = 16, TILE_H = 4, TILE_PX = TILE_W * TILE_H;

// ---------------------------------------------------------------------------
// forward, pixel-major.  Block = one 4x16 pixel tile; for every level the
// block's threads sweep (pixel, 4-channel group) items with the channel group
// fastest, so each warp store covers whole 128-byte lines of the output.
// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) hyper_fwd_hwc_kernel(const Levels L, T *__restrict__ out) {
    const int tx0 = blockIdx.x * TILE_W, ty0 = blockIdx.y * TILE_H;
    const int tid = threadIdx.x;
    for (int l = 0; l < L.n; ++l) {
        const int Cl = L.C[l], c4n = Cl >> 2, hl = L.h[l], wl = L.w[l];
        const float *__restrict__ src = L.src[l];
        const float sy = L.sy[l], sx = L.sx[l];
        const bool identity = (hl == L.H && wl == L.W);
        const int items = TILE_PX * c4n;
        for (int it = tid; it < items; it += 256) {
            int px = it / c4n;
            int c = (it - px * c4n) << 2;
            int y = ty0 + px / TILE_W, x = tx0 + (px % TILE_W);
            if (y >= L.H || x >= L.W) continue;
            float4 v;
            if (identity) {
                v = ldg_stream(reinterpret_cast<const float4 *>(src + ((long)y * wl + x) * Cl + c));
            } else {
                Tap ty = bilinear_tap(y, sy, hl), tx = bilinear_tap(x, sx, wl);
                const float *r0 = src + (long)ty.i0 * wl * Cl + c;
                const float *r1 = src + (long)ty.i1 * wl * Cl + c;
                float4 v00 = __ldg(reinterpret_cast<const float4 *>(r0 + (long)tx.i0 * Cl));
                float4 v01 = __ldg(reinterpret_cast<const float4 *>(r0 + (long)tx.i1 * Cl));
                float4 v10 = __ldg(reinterpret_cast<const float4 *>(r1 + (long)tx.i0 * Cl));
                float4 v11 = __ldg(reinterpret_cast<const float4 *>(r1 + (long)tx.i1 * Cl));
                float4 top = tx.w0 * v00; fma4(top, tx.w1, v01);
                float4 bot = tx.w0 * v10; fma4(bot, tx.w1, v11);
                v = ty.w0 * top; fma4(v, ty.w1, bot);
            }
            Vec4<T>::store(out + ((long)y * L.W + x) * L.Ctot + L.coff[l] + c, v);
        }
    }
}

// ---------------------------------------------------------------------------
// forward, pixel-major, "row walk" (the fast path).  Block = one output row
// segment; thread = one group of V consecutive channels (V*sizeof(T) = 16 bytes),
// so the block's threads cover the pixel's whole channel vector and every
// step of the walk stores Ctot*sizeof(T) contiguous bytes with 128-bit stores.
// The thread walks along x keeping the two vertically-blended source columns
// (i0, i1) of its level in registers: a new source column is fetched only when
// i0 advances (every 1/scale pixels), so L1/L2 read traffic is ~1/12 of the
// 4-taps-per-output form and the kernel is bound by the HBM write stream.
// ---------------------------------------------------------------------------
template <typename T, int V>
__global__ void __launch_bounds__(V == 4 ? 544 : 288, V == 4 ? 2 : 3) hyper_fwd_walk_kernel(const Levels L, T *__restrict__ out, int seg) {
    const int c = threadIdx.x * V;
    if (c >= L.Ctot) return;
    int l = 0;
    while (l + 1 < L.n && c >= L.coff[l + 1]) ++l;
    const int Cl = L.C[l], hl = L.h[l], wl = L.w[l];
    const float *__restrict__ src = L.src[l] + (c - L.coff[l]);
    const int y = blockIdx.y;
    const int x0 = blockIdx.x * seg, x1 = min(x0 + seg, L.W);
    const long ostride = L.Ctot;
    T *o = out + ((long)y * L.W + x0) * ostride + c;
    if (hl == L.H && wl == L.W) {                     // identity level: exact copy, 4 loads in flight
        const float *s = src + ((long)y * wl + x0) * Cl;
        int x = x0;
        for (; x + 4 <= x1; x += 4, s += 4 * (long)Cl, o += 4 * ostride) {
            FVec<V> v0 = ld_group<V>(s), v1 = ld_group<V>(s + Cl), v2 = ld_group<V>(s + 2 * (long)Cl), v3 = ld_group<V>(s + 3 * (long)Cl);
            st_group(o, v0); st_group(o + ostride, v1); st_group(o + 2 * ostride, v2); st_group(o + 3 * ostride, v3);
        }
        for (; x < x1; ++x, s += Cl, o += ostride) st_group(o, ld_group<V>(s));
        return;
    }
    const Tap ty = bilinear_tap(y, L.sy[l], hl);
    const float *__restrict__ r0 = src + (long)ty.i0 * wl * Cl;
    const float *__restrict__ r1 = src + (long)ty.i1 * wl * Cl;
    const float sx = L.sx[l];
    const int last = wl - 1;
    auto blend = [&](const FVec<V> &a, const FVec<V> &b) {
        FVec<V> r;
#pragma unroll
        for (int k = 0; k < V; ++k) r.v[k] = fmaf(ty.w1, b.v[k], ty.w0 * a.v[k]);
        return r;
    };
    // software pipeline: c0/c1 are the blended columns cur, cur+1; (na, nb) are the raw rows of column
    // cur+2, requested one advance ahead of their first use so the L2 round trip overlaps the walk
    int cur = bilinear_tap(x0, sx, wl).i0;
    FVec<V> c0 = blend(ld_group<V>(r0 + (long)cur * Cl), ld_group<V>(r1 + (long)cur * Cl));
    int i = min(cur + 1, last);
    FVec<V> c1 = blend(ld_group<V>(r0 + (long)i * Cl), ld_group<V>(r1 + (long)i * Cl));
    i = min(cur + 2, last);
    FVec<V> na = ld_group<V>(r0 + (long)i * Cl), nb = ld_group<V>(r1 + (long)i * Cl);
    for (int x = x0; x < x1; ++x, o += ostride) {
        const Tap tx = bilinear_tap(x, sx, wl);
        if (tx.i0 != cur) {
            if (tx.i0 == cur + 1) {
                c0 = c1;
                c1 = blend(na, nb);
            } else {                                   // scale > 1 never happens for an upsample; kept for safety
                c0 = blend(ld_group<V>(r0 + (long)tx.i0 * Cl), ld_group<V>(r1 + (long)tx.i0 * Cl));
                c1 = blend(ld_group<V>(r0 + (long)tx.i1 * Cl), ld_group<V>(r1 + (long)tx.i1 * Cl));
            }
            cur = tx.i0;
            i = min(cur + 2, last);
            na = ld_group<V>(r0 + (long)i * Cl);
            nb = ld_group<V>(r1 + (long)i * Cl);
        }
        FVec<V> r;
#pragma unroll
        for (int k = 0; k < V; ++k) r.v[k] = fmaf(tx.w1, c1.v[k], tx.w0 * c0.v[k]);
        st_group(o, r);
    }
}

// ---------------------------------------------------------------------------
// forward, pixel-major, bulk-staged row walk (the default fast path).
// Same decomposition as hyper_fwd_walk_kernel (block = one output row segment,
// thread = one 16-byte channel group, whole channel vector stored per step),
// but the source data never sits on the critical path: one thread per level
// issues cp.async.bulk (TMA 1-D bulk copies, SASS UBLKCP) of the two source
// rows x the few source columns the segment touches into shared memory,
// completion on an mbarrier; the walk then reads columns with conflict-free
// LDS.128.  With two such CTAs per SM one is always streaming stores while the
// other waits for its bulk loads, so global-load latency (which balloons while
// HBM is saturated with writes) no longer throttles the write stream.
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <typename T, int V>
__global__ void __launch_bounds__(V == 4 ? 544 : 288, 2) hyper_fwd_bulk_kernel(const Levels L, T *__restrict__ out, int seg) {
    extern __shared__ __align__(128) float stage[];
    __shared__ __align__(8) unsigned long long mbar;
    const int y = blockIdx.y;
    const int x0 = blockIdx.x * seg, x1 = min(x0 + seg, L.W);
    const uint32_t bar = smem_addr(&mbar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(L.n));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x < L.n) {                            // producer: one thread per level
        const int l = threadIdx.x;
        const int Cl = L.C[l], hl = L.h[l], wl = L.w[l];
        const bool ident = (hl == L.H && wl == L.W);
        int jlo, jhi, i0, i1;
        if (ident) { jlo = x0; jhi = x1 - 1; i0 = i1 = y; }
        else {
            jlo = bilinear_tap(x0, L.sx[l], wl).i0;
            jhi = bilinear_tap(x1 - 1, L.sx[l], wl).i1;
            const Tap ty = bilinear_tap(y, L.sy[l], hl);
            i0 = ty.i0; i1 = ty.i1;
        }
        const uint32_t bytes = (uint32_t)(jhi - jlo + 1) * Cl * 4u;
        const uint32_t dst0 = smem_addr(stage + L.soff[l]);
        const uint32_t dst1 = dst0 + (uint32_t)L.ncol[l] * Cl * 4u;
        const float *g0 = L.src[l] + ((long)i0 * wl + jlo) * Cl;
        const float *g1 = L.src[l] + ((long)i1 * wl + jlo) * Cl;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(ident ? bytes : 2u * bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst0), "l"(g0), "r"(bytes), "r"(bar) : "memory");
        if (!ident)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(dst1), "l"(g1), "r"(bytes), "r"(bar) : "memory");
    }
    const int c = threadIdx.x * V;
    const bool active = c < L.Ctot;
    int l = 0;
    while (l + 1 < L.n && c >= L.coff[l + 1]) ++l;
    const int Cl = L.C[l], hl = L.h[l], wl = L.w[l];
    const float sx = L.sx[l];
    const bool ident = (hl == L.H && wl == L.W);
    const Tap ty = bilinear_tap(y, L.sy[l], hl);
    const int jlo = ident ? x0 : bilinear_tap(x0, sx, wl).i0;
    const float *s0 = stage + L.soff[l] + (c - L.coff[l]) - (long)jlo * Cl;   // column j lives at s0 + j*Cl
    const float *s1 = s0 + (long)L.ncol[l] * Cl;
    T *o = out + ((long)y * L.W + x0) * L.Ctot + c;
    const long ostride = L.Ctot;
    {                                                   // wait for the bulk copies (phase 0)
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "HC_WAIT:\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t"
            "@p bra HC_DONE;\n\t"
            "bra HC_WAIT;\n\t"
            "HC_DONE:\n\t"
            "}\n" ::"r"(bar) : "memory");
    }
    if (!active) return;
    auto lds = [](const float *p) {
        FVec<V> r;
#pragma unroll
        for (int k = 0; k < V / 4; ++k) {
            float4 t = *reinterpret_cast<const float4 *>(p + 4 * k);
            r.v[4 * k] = t.x; r.v[4 * k + 1] = t.y; r.v[4 * k + 2] = t.z; r.v[4 * k + 3] = t.w;
        }
        return r;
    };
    if (ident) {
#pragma unroll 4
        for (int x = x0; x < x1; ++x, o += ostride) st_group(o, lds(s0 + (long)x * Cl));
        return;
    }
    auto column = [&](int i) {
        FVec<V> a = lds(s0 + (long)i * Cl), b = lds(s1 + (long)i * Cl), r;
#pragma unroll
        for (int k = 0; k < V; ++k) r.v[k] = fmaf(ty.w1, b.v[k], ty.w0 * a.v[k]);
        return r;
    };
    int cur = -2;
    FVec<V> c0, c1;
    for (int x = x0; x < x1; ++x, o += ostride) {
        const Tap tx = bilinear_tap(x, sx, wl);
        if (tx.i0 != cur) {
            c0 = (tx.i0 == cur + 1) ? c1 : column(tx.i0);
            c1 = (tx.i1 != tx.i0) ? column(tx.i1) : c0;
            cur = tx.i0;
        }
        FVec<V> r;
#pragma unroll
        for (int k = 0; k < V; ++k) r.v[k] = fmaf(tx.w1, c1.v[k], tx.w0 * c0.v[k]);
        st_group(o, r);
    }
}

// ---------------------------------------------------------------------------
// forward, pixel-major, PERSISTENT double-buffered variant of the bulk-staged
// walk: one CTA per SM loops over (row, segment) tiles; while the CTA streams
// the stores of tile t from stage t&1, the bulk copies of tile t+1 land in the
// other stage (one mbarrier per stage, phase = use count parity).  Loads are
// completely off the critical path: the kernel is a pure HBM write stream.
// ---------------------------------------------------------------------------
template <typename T, int V>
__global__ void __launch_bounds__(V == 4 ? 544 : 288, 1) hyper_fwd_pipe_kernel(const Levels L, T *__restrict__ out, int seg, int nseg,
                                                                              int n_tiles, int stage_floats) {
    extern __shared__ __align__(128) float stage[];
    __shared__ __align__(8) unsigned long long mbar[2];
    const uint32_t bar0 = smem_addr(&mbar[0]);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar0), "r"(L.n));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar0 + 8u), "r"(L.n));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // per-thread level constants
    const int c = threadIdx.x * V;
    const bool active = c < L.Ctot;
    int l = 0;
    while (l + 1 < L.n && c >= L.coff[l + 1]) ++l;
    const int Cl = L.C[l], hl = L.h[l], wl = L.w[l];
    const float sx = L.sx[l], sy = L.sy[l];
    const bool ident = (hl == L.H && wl == L.W);
    const int lane_off = L.soff[l] + (c - L.coff[l]);
    const long row_floats = (long)L.ncol[l] * Cl;
    const long ostride = L.Ctot;
    // producer constants (threads 0 .. n-1 each own one level)
    const bool producer = threadIdx.x < L.n;
    const int pl = producer ? threadIdx.x : 0;
    const int pC = L.C[pl], ph = L.h[pl], pw = L.w[pl];
    const bool pident = (ph == L.H && pw == L.W);

    auto issue = [&](int tile, int st) {               // called by producer threads only
        const int y = tile / nseg, x0 = (tile - y * nseg) * seg, x1 = min(x0 + seg, L.W);
        int jlo, jhi, i0, i1;
        if (pident) { jlo = x0; jhi = x1 - 1; i0 = i1 = y; }
        else {
            jlo = bilinear_tap(x0, L.sx[pl], pw).i0;
            jhi = bilinear_tap(x1 - 1, L.sx[pl], pw).i1;
            const Tap ty = bilinear_tap(y, L.sy[pl], ph);
            i0 = ty.i0; i1 = ty.i1;
        }
        const uint32_t bytes = (uint32_t)(jhi - jlo + 1) * pC * 4u;
        const uint32_t bar = bar0 + 8u * st;
        const uint32_t dst0 = smem_addr(stage + (long)st * stage_floats + L.soff[pl]);
        const uint32_t dst1 = dst0 + (uint32_t)L.ncol[pl] * pC * 4u;
        const float *g0 = L.src[pl] + ((long)i0 * pw + jlo) * pC;
        const float *g1 = L.src[pl] + ((long)i1 * pw + jlo) * pC;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(pident ? bytes : 2u * bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst0), "l"(g0), "r"(bytes), "r"(bar) : "memory");
        if (!pident)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(dst1), "l"(g1), "r"(bytes), "r"(bar) : "memory");
    };
    auto lds = [](const float *p) {
        FVec<V> r;
#pragma unroll
        for (int k = 0; k < V / 4; ++k) {
            float4 t = *reinterpret_cast<const float4 *>(p + 4 * k);
            r.v[4 * k] = t.x; r.v[4 * k + 1] = t.y; r.v[4 * k + 2] = t.z; r.v[4 * k + 3] = t.w;
        }
        return r;
    };

    __syncthreads();                                    // barrier init visible
    int tile = blockIdx.x;
    if (producer && tile < n_tiles) issue(tile, 0);
    for (int it = 0; tile < n_tiles; ++it, tile += gridDim.x) {
        const int st = it & 1;
        __syncthreads();                                // every warp finished tile it-1 => stage st^1 is free
        if (producer && tile + (int)gridDim.x < n_tiles) issue(tile + gridDim.x, st ^ 1);
        {                                               // wait for this tile's bulk copies
            const uint32_t bar = bar0 + 8u * st, parity = (uint32_t)(it >> 1) & 1u;
            asm volatile(
                "{\n\t"
                ".reg .pred p;\n\t"
                "HCP_WAIT:\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                "@p bra HCP_DONE;\n\t"
                "bra HCP_WAIT;\n\t"
                "HCP_DONE:\n\t"
                "}\n" ::"r"(bar), "r"(parity) : "memory");
        }
        if (!active) continue;
        const int y = tile / nseg, x0 = (tile - y * nseg) * seg, x1 = min(x0 + seg, L.W);
        T *o = out + ((long)y * L.W + x0) * ostride + c;
        const float *base = stage + (long)st * stage_floats + lane_off;
        if (ident) {
            const float *s0 = base - (long)x0 * Cl;
#pragma unroll 4
            for (int x = x0; x < x1; ++x, o += ostride) st_group(o, lds(s0 + (long)x * Cl));
            continue;
        }
        const Tap ty = bilinear_tap(y, sy, hl);
        const int jlo = bilinear_tap(x0, sx, wl).i0;
        const float *s0 = base - (long)jlo * Cl;        // column j lives at s0 + j*Cl
        const float *s1 = s0 + row_floats;
        auto column = [&](int i) {
            FVec<V> a = lds(s0 + (long)i * Cl), b = lds(s1 + (long)i * Cl), r;
#pragma unroll
            for (int k = 0; k < V; ++k) r.v[k] = fmaf(ty.w1, b.v[k], ty.w0 * a.v[k]);
            return r;
        };
        int cur = -2;
        FVec<V> c0, c1;
        for (int x = x0; x < x1; ++x, o += ostride) {
            const Tap tx = bilinear_tap(x, sx, wl);
            if (tx.i0 != cur) {
                c0 = (tx.i0 == cur + 1) ? c1 : column(tx.i0);
                c1 = (tx.i1 != tx.i0) ? column(tx.i1) : c0;
                cur = tx.i0;
            }
            FVec<V> r;
#pragma unroll
            for (int k = 0; k < V; ++k) r.v[k] = fmaf(tx.w1, c1.v[k], tx.w0 * c0.v[k]);
            st_group(o, r);
        }
    }
}

template <typename T, int V>
static cudaError_t launch_pipe(const Levels &L, size_t stage_bytes, T *out, int seg, int H, int W, cudaStream_t stream) {
    static size_t configured = 0;
    const size_t smem = 2 * stage_bytes;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(hyper_fwd_pipe_kernel<T, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    const int threads = (L.Ctot / V + 31) / 32 * 32;
    const int nseg = cdiv(W, seg);
    const long n_tiles = (long)nseg * H;
    const int grid = (int)(n_tiles < kNumSMs ? n_tiles : kNumSMs);
    hyper_fwd_pipe_kernel<T, V><<<grid, threads, smem, stream>>>(L, out, seg, nseg, (int)n_tiles, (int)(stage_bytes / 4));
    return cudaSuccess;
}

template <typename T, int V>
static cudaError_t launch_bulk(const Levels &L, size_t smem, T *out, int seg, int H, int W, cudaStream_t stream) {
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(hyper_fwd_bulk_kernel<T, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    const int threads = (L.Ctot / V + 31) / 32 * 32;
    hyper_fwd_bulk_kernel<T, V><<<dim3(cdiv(W, seg), H), threads, smem, stream>>>(L, out, seg);
    return cudaSuccess;
}

// forward, channel-major: one thread per (channel, y, 4 consecutive x)
template <typename T>
__global__ void __launch_bounds__(256) hyper_fwd_chw_kernel(const Levels L, T *__restrict__ out) {
    const int l = blockIdx.z;
    const int Cl = L.C[l], hl = L.h[l], wl = L.w[l];
    const int W4 = (L.W + 3) >> 2;
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    long total = (long)Cl * L.H * W4;
    if (idx >= total) return;
    int x4 = (int)(idx % W4);
    long t = idx / W4;
    int y = (int)(t % L.H), c = (int)(t / L.H);
    const float *__restrict__ plane = L.src[l] + (long)c * hl * wl;
    Tap ty = bilinear_tap(y, L.sy[l], hl);
    T *o = out + ((long)(L.coff[l] + c) * L.H + y) * L.W;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        int x = x4 * 4 + j;
        if (x >= L.W) break;
        Tap tx = bilinear_tap(x, L.sx[l], wl);
        float top = tx.w0 * __ldg(plane + (long)ty.i0 * wl + tx.i0);
        top = fmaf(tx.w1, __ldg(plane + (long)ty.i0 * wl + tx.i1), top);
        float bot = tx.w0 * __ldg(plane + (long)ty.i1 * wl + tx.i0);
        bot = fmaf(tx.w1, __ldg(plane + (long)ty.i1 * wl + tx.i1), bot);
        o[x] = static_cast<T>(fmaf(ty.w1, bot, ty.w0 * top));
    }
}

// ---------------------------------------------------------------------------
// backward (adjoint), gather form.  For low-res index i, the output indices d
// whose taps touch i form a contiguous range; it is bracketed conservatively
// and every candidate is tested with the *same* fp32 tap arithmetic as the
// forward, so fwd and bwd are exact adjoints of each other.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void footprint(int i, float scale, int out_size, int &lo, int &hi) {
    if (scale <= 0.f) { lo = 0; hi = out_size - 1; return; }
    float inv = 1.0f / scale;
    lo = (int)floorf((float)(i - 1) * inv) - 1;
    hi = (int)ceilf((float)(i + 1) * inv) + 1;
    lo = max(lo, 0);
    hi = min(hi, out_size - 1);
}
__device__ __forceinline__ float tap_weight(int d, int i, float scale, int in_size) {
    Tap t = bilinear_tap(d, scale, in_size);
    float wgt = 0.f;
    if (t.i0 == i) wgt += t.w0;
    if (t.i1 == i) wgt += t.w1;
    return wgt;
}

// pixel-major: one thread per (level, low-res pixel, 4-channel group)
template <typename T>
__global__ void __launch_bounds__(256) hyper_bwd_hwc_kernel(const Levels L, const T *__restrict__ grad_out) {
    const int l = blockIdx.y;
    const int Cl = L.C[l], c4n = Cl >> 2, hl = L.h[l], wl = L.w[l];
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    long total = (long)hl * wl * c4n;
    if (idx >= total) return;
    int c = ((int)(idx % c4n)) << 2;
    long q = idx / c4n;
    int j = (int)(q % wl), i = (int)(q / wl);
    const T *g = grad_out + L.coff[l] + c;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (hl == L.H && wl == L.W) {
        acc = Vec4<T>::load(g + ((long)i * L.W + j) * L.Ctot);
    } else {
        int ylo, yhi, xlo, xhi;
        footprint(i, L.sy[l], L.H, ylo, yhi);
        footprint(j, L.sx[l], L.W, xlo, xhi);
        for (int y = ylo; y <= yhi; ++y) {
            float wy = tap_weight(y, i, L.sy[l], hl);
            if (wy == 0.f) continue;
            float4 row = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int x = xlo; x <= xhi; ++x) {
                float wx = tap_weight(x, j, L.sx[l], wl);
                if (wx == 0.f) continue;
                fma4(row, wx, Vec4<T>::load_cached(g + ((long)y * L.W + x) * L.Ctot));
            }
            fma4(acc, wy, row);
        }
    }
    *reinterpret_cast<float4 *>(L.dst[l] + ((long)i * wl + j) * Cl + c) = acc;
}

// channel-major: one thread per (level, channel, low-res pixel)
template <typename T>
__global__ void __launch_bounds__(256) hyper_bwd_chw_kernel(const Levels L, const T *__restrict__ grad_out) {
    const int l = blockIdx.y;
    const int Cl = L.C[l], hl = L.h[l], wl = L.w[l];
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    long total = (long)Cl * hl * wl;
    if (idx >= total) return;
    int j = (int)(idx % wl);
    long t = idx / wl;
    int i = (int)(t % hl), c = (int)(t / hl);
    const T *plane = grad_out + (long)(L.coff[l] + c) * L.H * L.W;
    float acc = 0.f;
    int ylo, yhi, xlo, xhi;
    footprint(i, L.sy[l], L.H, ylo, yhi);
    footprint(j, L.sx[l], L.W, xlo, xhi);
    for (int y = ylo; y <= yhi; ++y) {
        float wy = tap_weight(y, i, L.sy[l], hl);
        if (wy == 0.f) continue;
        float row = 0.f;
        for (int x = xlo; x <= xhi; ++x) {
            float wx = tap_weight(x, j, L.sx[l], wl);
            if (wx == 0.f) continue;
            row = fmaf(wx, (float)plane[(long)y * L.W + x], row);
        }
        acc = fmaf(wy, row, acc);
    }
    L.dst[l][idx] = acc;
}

// ---------------------------------------------------------------------------
// backward, pixel-major, separable two-pass form (the fast path when the caller
// provides a workspace).  Pass 1 streams grad_out exactly once: block = one
// output row x a slice of the channel groups, thread = one 16-byte channel
// group walking along x with the two live low-res column accumulators in
// registers; a finished column is written once to R[l][y][j][c] (identity
// levels go straight to their gradient).  Pass 2 reduces R over the <= 2/scale+2
// output rows of each low-res row.  Deterministic (fixed order, no atomics);
// extra traffic = write + read of R (sum_l H*w_l*C_l*4 bytes, 0.32 GB at 464^2)
// on top of the 1.82 GB that must be read anyway.
// ---------------------------------------------------------------------------
// The per-level row partials R[l] (H, w_l, C_l) fp32 live in the workspace; their pointers
// travel in Levels::src (unused by the backward otherwise) so that the kernels keep a
// single parameter struct (a second dynamically indexed struct is copied to local memory).
template <int V> __device__ __forceinline__ FVec<V> ld_grad(const float *p) {
    FVec<V> r;
#pragma unroll
    for (int k = 0; k < V / 4; ++k) {
        float4 t = ldg_stream(reinterpret_cast<const float4 *>(p) + k);
        r.v[4 * k] = t.x; r.v[4 * k + 1] = t.y; r.v[4 * k + 2] = t.z; r.v[4 * k + 3] = t.w;
    }
    return r;
}
template <int V> __device__ __forceinline__ FVec<V> ld_grad(const __nv_bfloat16 *p) {
    FVec<V> r;
#pragma unroll
    for (int k = 0; k < V / 4; ++k) {
        float4 t = unpack_bf16x4(ldg_stream(reinterpret_cast<const uint2 *>(p) + k));
        r.v[4 * k] = t.x; r.v[4 * k + 1] = t.y; r.v[4 * k + 2] = t.z; r.v[4 * k + 3] = t.w;
    }
    return r;
}
__device__ __forceinline__ void st_f4(float *p, const FVec<4> &a) {
    *reinterpret_cast<float4 *>(p) = make_float4(a.v[0], a.v[1], a.v[2], a.v[3]);
}

template <typename T>
__global__ void __launch_bounds__(160, 4) hyper_bwd_rows_kernel(const Levels L, const T *__restrict__ grad_out, int groups_per_block) {
    constexpr int V = 4;
    const int g = blockIdx.y * groups_per_block + threadIdx.x;
    if (threadIdx.x >= groups_per_block || g * V >= L.Ctot) return;
    const int c = g * V;
    int l = 0;
    while (l + 1 < L.n && c >= L.coff[l + 1]) ++l;
    const int Cl = L.C[l], hl = L.h[l], wl = L.w[l], cl = c - L.coff[l];
    const int y = blockIdx.x, W = L.W;
    const T *__restrict__ gin = grad_out + (long)y * W * L.Ctot + c;
    const long gstride = L.Ctot;
    if (hl == L.H && wl == W) {                        // identity level: the gradient is a copy
        float *d = L.dst[l] + (long)y * wl * Cl + cl;
        int x = 0;
        for (; x + 4 <= W; x += 4) {
            FVec<V> v0 = ld_grad<V>(gin + (long)x * gstride), v1 = ld_grad<V>(gin + (long)(x + 1) * gstride),
                    v2 = ld_grad<V>(gin + (long)(x + 2) * gstride), v3 = ld_grad<V>(gin + (long)(x + 3) * gstride);
            st_f4(d + (long)x * Cl, v0); st_f4(d + (long)(x + 1) * Cl, v1); st_f4(d + (long)(x + 2) * Cl, v2); st_f4(d + (long)(x + 3) * Cl, v3);
        }
        for (; x < W; ++x) st_f4(d + (long)x * Cl, ld_grad<V>(gin + (long)x * gstride));
        return;
    }
    float *__restrict__ R = const_cast<float *>(L.src[l]) + (long)y * wl * Cl + cl;
    const float sx = L.sx[l];
    int cur = bilinear_tap(0, sx, wl).i0;
    FVec<V> a0, a1;
#pragma unroll
    for (int k = 0; k < V; ++k) { a0.v[k] = 0.f; a1.v[k] = 0.f; }
    auto step = [&](int x, const FVec<V> &gv) {
        const Tap tx = bilinear_tap(x, sx, wl);
        while (tx.i0 != cur) {                          // advance (by one for an upsample)
            st_f4(R + (long)cur * Cl, a0);
            a0 = a1;
#pragma unroll
            for (int k = 0; k < V; ++k) a1.v[k] = 0.f;
            ++cur;
        }
        if (tx.i1 != tx.i0) {
#pragma unroll
            for (int k = 0; k < V; ++k) { a0.v[k] = fmaf(tx.w0, gv.v[k], a0.v[k]); a1.v[k] = fmaf(tx.w1, gv.v[k], a1.v[k]); }
        } else {
            const float wsum = tx.w0 + tx.w1;
#pragma unroll
            for (int k = 0; k < V; ++k) a0.v[k] = fmaf(wsum, gv.v[k], a0.v[k]);
        }
    };
    int x = 0;
    for (; x + 8 <= W; x += 8) {                        // eight independent 16-byte loads in flight per thread
        FVec<V> v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = ld_grad<V>(gin + (long)(x + u) * gstride);
#pragma unroll
        for (int u = 0; u < 8; ++u) step(x + u, v[u]);
    }
    for (; x < W; ++x) step(x, ld_grad<V>(gin + (long)x * gstride));
    st_f4(R + (long)cur * Cl, a0);
    for (int j = cur + 1; j < wl; ++j) {               // at most one more column carries weight
        st_f4(R + (long)j * Cl, a1);
#pragma unroll
        for (int k = 0; k < V; ++k) a1.v[k] = 0.f;
    }
}

// y-footprint tables for pass 2 (same fp32 tap arithmetic as the forward): for low-res row i of
// level l, the first output row, the number of rows and their weights.  Pointers travel in
// Levels::dst-adjacent fields would need a second struct; the tables are therefore packed behind
// one base pointer with per-level offsets stored in Levels::ncol (row length K) / Levels::soff.
__global__ void hyper_bwd_ytable_kernel(const Levels L, int32_t *__restrict__ tab) {
    const int l = blockIdx.y;
    const int hl = L.h[l];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= hl || (hl == L.H && L.w[l] == L.W)) return;
    const int K = L.ncol[l];
    int32_t *base = tab + L.soff[l];                   // [hl] first, [hl] count, [hl*K] weights (float bits)
    int lo, hi;
    footprint(i, L.sy[l], L.H, lo, hi);
    int first = -1, last = -2;
    for (int d = lo; d <= hi; ++d) {
        Tap t = bilinear_tap(d, L.sy[l], hl);
        if (t.i0 == i || t.i1 == i) { if (first < 0) first = d; last = d; }
    }
    int n = first < 0 ? 0 : min(last - first + 1, K);
    base[i] = first < 0 ? 0 : first;
    base[hl + i] = n;
    for (int k = 0; k < n; ++k) base[2 * hl + (long)i * K + k] = __float_as_int(tap_weight(first + k, i, L.sy[l], hl));
}

// pass 2: one thread per (non-identity level, low-res pixel, 4-channel group)
__global__ void __launch_bounds__(256) hyper_bwd_cols_kernel(const Levels L, const int32_t *__restrict__ tab) {
    const int l = blockIdx.y;
    const int Cl = L.C[l], c4n = Cl >> 2, hl = L.h[l], wl = L.w[l];
    if (hl == L.H && wl == L.W) return;
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)hl * wl * c4n) return;
    const int c = ((int)(idx % c4n)) << 2;
    const long q = idx / c4n;
    const int j = (int)(q % wl), i = (int)(q / wl);
    const long rstride = (long)wl * Cl;
    const int32_t *__restrict__ base = tab + L.soff[l];
    const int y0 = __ldg(base + i), n = __ldg(base + hl + i);
    const int32_t *__restrict__ wy = base + 2 * hl + (long)i * L.ncol[l];
    const float *__restrict__ R = L.src[l] + (long)y0 * rstride + (long)j * Cl + c;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int k = 0;
    for (; k + 4 <= n; k += 4) {                        // four independent row loads in flight
        float4 r0 = __ldg(reinterpret_cast<const float4 *>(R + (long)k * rstride));
        float4 r1 = __ldg(reinterpret_cast<const float4 *>(R + (long)(k + 1) * rstride));
        float4 r2 = __ldg(reinterpret_cast<const float4 *>(R + (long)(k + 2) * rstride));
        float4 r3 = __ldg(reinterpret_cast<const float4 *>(R + (long)(k + 3) * rstride));
        fma4(acc, __int_as_float(__ldg(wy + k)), r0); fma4(acc, __int_as_float(__ldg(wy + k + 1)), r1);
        fma4(acc, __int_as_float(__ldg(wy + k + 2)), r2); fma4(acc, __int_as_float(__ldg(wy + k + 3)), r3);
    }
    for (; k < n; ++k) fma4(acc, __int_as_float(__ldg(wy + k)), __ldg(reinterpret_cast<const float4 *>(R + (long)k * rstride)));
    *reinterpret_cast<float4 *>(L.dst[l] + ((long)i * wl + j) * Cl + c) = acc;
}

static int fill_levels(Levels &L, const char *who, const int *C, const int *h, const int *w, int n_levels, int H, int W,
                       int layout) {
    WESUP_REQUIRE(C && h && w, WESUP_E_ARG, "%s: null geometry", who);
    WESUP_REQUIRE(n_levels > 0 && n_levels <= WESUP_MAX_LEVELS, WESUP_E_ARG, "%s: n_levels=%d out of range", who, n_levels);
    WESUP_REQUIRE(H > 0 && W > 0, WESUP_E_ARG, "%s: bad output size %dx%d", who, H, W);
    WESUP_REQUIRE(layout == WESUP_CHW || layout == WESUP_HWC, WESUP_E_ARG, "%s: bad layout %d", who, layout);
    L.n = n_levels; L.H = H; L.W = W;
    int off = 0;
    for (int l = 0; l < n_levels; ++l) {
        WESUP_REQUIRE(C[l] > 0 && h[l] > 0 && w[l] > 0, WESUP_E_ARG, "%s: level %d has empty shape", who, l);
        if (layout == WESUP_HWC)
            WESUP_REQUIRE(C[l] % 4 == 0, WESUP_E_ALIGN, "%s: HWC layout needs C[%d]=%d to be a multiple of 4", who, l, C[l]);
        L.C[l] = C[l]; L.h[l] = h[l]; L.w[l] = w[l]; L.coff[l] = off;
        L.sy[l] = bilinear_scale(h[l], H); L.sx[l] = bilinear_scale(w[l], W);
        off += C[l];
    }
    L.Ctot = off;
    return 0;
}

}  // namespace wesup

using namespace wesup;

extern "C" int wesup_hypercolumn_fwd(const void *const *side, const int *C, const int *h, const int *w, int n_levels,
                                     int H, int W, void *out, int out_dtype, int layout, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    WESUP_REQUIRE(side && out, WESUP_E_ARG, "wesup_hypercolumn_fwd: null pointer");
    WESUP_REQUIRE(out_dtype == WESUP_F32 || out_dtype == WESUP_BF16, WESUP_E_ARG, "wesup_hypercolumn_fwd: bad dtype %d", out_dtype);
    Levels L;
    int rc = fill_levels(L, "wesup_hypercolumn_fwd", C, h, w, n_levels, H, W, layout);
    if (rc) return rc;
    for (int l = 0; l < n_levels; ++l) {
        WESUP_REQUIRE(side[l] != nullptr, WESUP_E_ARG, "wesup_hypercolumn_fwd: side[%d] is null", l);
        WESUP_REQUIRE(layout == WESUP_CHW || aligned16(side[l]), WESUP_E_ALIGN, "wesup_hypercolumn_fwd: side[%d] not 16-byte aligned", l);
        L.src[l] = static_cast<const float *>(side[l]);
        L.dst[l] = nullptr;
    }
    if (layout == WESUP_HWC) {
        WESUP_REQUIRE(aligned16(out), WESUP_E_ALIGN, "wesup_hypercolumn_fwd: out not 16-byte aligned");
        // fast path: one thread per 16-byte channel group, whole channel vector per block
        const int V = out_dtype == WESUP_F32 ? 4 : 8;
        bool walk = (L.Ctot % V == 0) && (L.Ctot / V <= (V == 4 ? 544 : 288)) && H <= 65535;
        for (int l = 0; l < n_levels; ++l) walk = walk && (C[l] % V == 0);
        // 1 (default): bulk-staged walk, two CTAs per SM; 2: persistent double-buffered; 0: plain walk
        static const int variant = getenv("WESUP_HC_FWD") ? atoi(getenv("WESUP_HC_FWD")) : 1;
        static const int seg_env = getenv("WESUP_HC_SEG") ? atoi(getenv("WESUP_HC_SEG")) : 0;
        static const int bf16_v = getenv("WESUP_HC_BF16_V") ? atoi(getenv("WESUP_HC_BF16_V")) : 8;
        const size_t smem_cap = (variant == 2 ? 110 : 110) * 1024;      // two stages / two CTAs per SM
        auto plan = [&](int sg) {
            size_t bytes = 0;
            for (int l = 0; l < n_levels; ++l) {
                const bool ident = (h[l] == H && w[l] == W);
                int ncol = ident ? sg : (int)floorf((float)(sg - 1) * L.sx[l]) + 3;
                if (ncol > w[l]) ncol = w[l];
                L.ncol[l] = ncol;
                L.soff[l] = (int)(bytes / 4);
                bytes += (size_t)(ident ? 1 : 2) * ncol * C[l] * 4;
            }
            return bytes;
        };
        size_t smem = 0;
        int bseg = seg_env;
        if (walk && variant >= 1) {
            if (bseg > 0) smem = plan(bseg);
            else                                                       // longest segment whose staging fits
                for (bseg = 32; bseg >= 8; bseg -= 4)
                    if ((smem = plan(bseg)) <= smem_cap) break;
            if (bseg < 8) smem = smem_cap + 1;
        }
        const bool staged_ok = walk && variant >= 1 && smem <= smem_cap && n_levels <= 32;
        if (staged_ok && variant == 2 && (long)cdiv(W, bseg) * H < (1L << 31)) {
            cudaError_t e = out_dtype == WESUP_F32 ? launch_pipe<float, 4>(L, smem, (float *)out, bseg, H, W, stream)
                                                   : launch_pipe<__nv_bfloat16, 8>(L, smem, (__nv_bfloat16 *)out, bseg, H, W, stream);
            WESUP_REQUIRE(e == cudaSuccess, (int)e, "wesup_hypercolumn_fwd: %s", cudaGetErrorString(e));
        } else if (staged_ok) {
            cudaError_t e = out_dtype == WESUP_F32 ? launch_bulk<float, 4>(L, smem, (float *)out, bseg, H, W, stream)
                            : bf16_v == 4          ? launch_bulk<__nv_bfloat16, 4>(L, smem, (__nv_bfloat16 *)out, bseg, H, W, stream)
                                                   : launch_bulk<__nv_bfloat16, 8>(L, smem, (__nv_bfloat16 *)out, bseg, H, W, stream);
            WESUP_REQUIRE(e == cudaSuccess, (int)e, "wesup_hypercolumn_fwd: %s", cudaGetErrorString(e));
        } else if (walk) {
            const int seg = seg_env > 0 ? seg_env : 64;
            const int threads = (L.Ctot / V + 31) / 32 * 32;
            dim3 grid(cdiv(W, seg), H);
            if (out_dtype == WESUP_F32) hyper_fwd_walk_kernel<float, 4><<<grid, threads, 0, stream>>>(L, (float *)out, seg);
            else hyper_fwd_walk_kernel<__nv_bfloat16, 8><<<grid, threads, 0, stream>>>(L, (__nv_bfloat16 *)out, seg);
        } else {
            dim3 grid(cdiv(W, TILE_W), cdiv(H, TILE_H));
            if (out_dtype == WESUP_F32) hyper_fwd_hwc_kernel<float><<<grid, 256, 0, stream>>>(L, (float *)out);
            else hyper_fwd_hwc_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(L, (__nv_bfloat16 *)out);
        }
    } else {
        int cmax = 0;
        for (int l = 0; l < n_levels; ++l) cmax = cmax > C[l] ? cmax : C[l];
        long per_level = (long)cmax * H * ((W + 3) / 4);
        dim3 grid(cdiv(per_level, 256), 1, n_levels);
        if (out_dtype == WESUP_F32) hyper_fwd_chw_kernel<float><<<grid, 256, 0, stream>>>(L, (float *)out);
        else hyper_fwd_chw_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(L, (__nv_bfloat16 *)out);
    }
    WESUP_CHECK_LAUNCH("wesup_hypercolumn_fwd", 1);
    return 0;
}

static inline int ytable_len(int hl, int H) {
    float scale = bilinear_scale(hl, H);
    if (!(scale > 0.f)) return H;
    int k = (int)(2.0f / scale) + 7;
    return k < H + 2 ? k : H + 2;
}

extern "C" size_t wesup_hypercolumn_bwd_workspace_bytes(const int *C, const int *h, const int *w, int n_levels, int H, int W) {
    if (!C || !h || !w || n_levels <= 0 || n_levels > WESUP_MAX_LEVELS || H <= 0 || W <= 0) return 0;
    size_t total = 0;
    for (int l = 0; l < n_levels; ++l) {
        if (C[l] <= 0 || h[l] <= 0 || w[l] <= 0) return 0;
        if (!(h[l] == H && w[l] == W)) {
            total += ((size_t)H * w[l] * C[l] * sizeof(float) + 255) / 256 * 256;
            total += ((size_t)h[l] * (2 + ytable_len(h[l], H)) * sizeof(int32_t) + 255) / 256 * 256;
        }
    }
    return total > 256 ? total : 256;
}

extern "C" int wesup_hypercolumn_bwd(const void *grad_out, int grad_dtype, int layout, const int *C, const int *h,
                                     const int *w, int n_levels, int H, int W, void *const *grad_side, void *ws,
                                     void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    WESUP_REQUIRE(grad_out && grad_side, WESUP_E_ARG, "wesup_hypercolumn_bwd: null pointer");
    WESUP_REQUIRE(grad_dtype == WESUP_F32 || grad_dtype == WESUP_BF16, WESUP_E_ARG, "wesup_hypercolumn_bwd: bad dtype %d", grad_dtype);
    Levels L;
    int rc = fill_levels(L, "wesup_hypercolumn_bwd", C, h, w, n_levels, H, W, layout);
    if (rc) return rc;
    long biggest = 0;
    bool walk = layout == WESUP_HWC && ws != nullptr && aligned16(ws) && H <= 65535;
    for (int l = 0; l < n_levels; ++l) {
        WESUP_REQUIRE(grad_side[l] != nullptr, WESUP_E_ARG, "wesup_hypercolumn_bwd: grad_side[%d] is null", l);
        WESUP_REQUIRE(layout == WESUP_CHW || aligned16(grad_side[l]), WESUP_E_ALIGN, "wesup_hypercolumn_bwd: grad_side[%d] not 16-byte aligned", l);
        L.src[l] = nullptr;
        L.dst[l] = static_cast<float *>(grad_side[l]);
        long n = (long)h[l] * w[l] * (layout == WESUP_HWC ? C[l] / 4 : C[l]);
        biggest = biggest > n ? biggest : n;
    }
    if (layout == WESUP_HWC) {
        WESUP_REQUIRE(aligned16(grad_out), WESUP_E_ALIGN, "wesup_hypercolumn_bwd: grad_out not 16-byte aligned");
        if (walk) {
            char *p = static_cast<char *>(ws);
            for (int l = 0; l < n_levels; ++l) {
                L.src[l] = reinterpret_cast<const float *>(p);
                if (!(h[l] == H && w[l] == W)) p += ((size_t)H * w[l] * C[l] * sizeof(float) + 255) / 256 * 256;
            }
            int32_t *tab = reinterpret_cast<int32_t *>(p);
            size_t toff = 0;
            int hmax = 1;
            for (int l = 0; l < n_levels; ++l) {
                L.ncol[l] = ytable_len(h[l], H);
                L.soff[l] = (int)(toff / sizeof(int32_t));
                if (!(h[l] == H && w[l] == W)) toff += ((size_t)h[l] * (2 + L.ncol[l]) * sizeof(int32_t) + 255) / 256 * 256;
                hmax = hmax > h[l] ? hmax : h[l];
            }
            hyper_bwd_ytable_kernel<<<dim3(cdiv(hmax, 128), n_levels), 128, 0, stream>>>(L, tab);
            const int groups = L.Ctot / 4;
            const int slices = (groups + 131) / 132;                 // 132 groups (5 warps) per block
            const int gpb = (groups + slices - 1) / slices;
            dim3 grid1(H, slices);
            if (grad_dtype == WESUP_F32) hyper_bwd_rows_kernel<float><<<grid1, 160, 0, stream>>>(L, (const float *)grad_out, gpb);
            else hyper_bwd_rows_kernel<__nv_bfloat16><<<grid1, 160, 0, stream>>>(L, (const __nv_bfloat16 *)grad_out, gpb);
            hyper_bwd_cols_kernel<<<dim3(cdiv(biggest, 256), n_levels), 256, 0, stream>>>(L, tab);
            WESUP_CHECK_LAUNCH("wesup_hypercolumn_bwd", 3);
            return 0;
        }
        dim3 grid(cdiv(biggest, 256), n_levels);
        if (grad_dtype == WESUP_F32) hyper_bwd_hwc_kernel<float><<<grid, 256, 0, stream>>>(L, (const float *)grad_out);
        else hyper_bwd_hwc_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(L, (const __nv_bfloat16 *)grad_out);
    } else {
        dim3 grid(cdiv(biggest, 256), n_levels);
        if (grad_dtype == WESUP_F32) hyper_bwd_chw_kernel<float><<<grid, 256, 0, stream>>>(L, (const float *)grad_out);
        else hyper_bwd_chw_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(L, (const __nv_bfloat16 *)grad_out);
    }
    WESUP_CHECK_LAUNCH("wesup_hypercolumn_bwd", 1);
    return 0;
}
