// Pixel-wise inference, first MLP layer without the hypercolumn.
//
// WESUPPixelInference.forward (/root/reference/models/wesup.py:382-400) feeds the (H*W, 2112) hypercolumn --
// 13 x [1x1 side conv -> bilinear upsample] concatenated (:246-261) -- through Linear(2112, 1024) + ReLU.  All
// three steps before the ReLU are linear and the channel mixing commutes with the spatial interpolation:
//     Linear1(cat_l upsample(side_l(f_l))) = b' + sum_g upsample_g( cat_{l in g} f_l . W'_g^T ),
// one term per distinct level RESOLUTION g (5 for VGG16), with W'_g = cat_l (W1[:, slice_l] . Wside_l) and
// b' = b1 + sum_l W1[:, slice_l] . bside_l (bilinear weights sum to one).  The GEMMs then run at the levels' own
// resolution (88 GFLOP per 400-px tile instead of 692 on the hypercolumn, library GEMMs on N x C_l operands) and
// what is left at full resolution is this kernel:
//     out[p, :] = act( bias + sum_g bilinear_g(Z_g)[p, :] ),   Z_g: (h_g, w_g, C) pixel-major,
// a streaming pass that reads the full-resolution term once, re-reads the low-resolution terms from L2, and writes
// the layer's output once with 128-bit stores -- neither the (H*W, 2112) hypercolumn nor the 13 upsampled side
// outputs ever exist.  Same decomposition as the hypercolumn walk kernel (hypercolumn.cu): block = one output row
// segment, thread = one 16-byte channel group walking along x with the two vertically blended source columns of
// every low-resolution term in registers (a new column is fetched only when the source index advances).
#include "common.cuh"
#include <stdlib.h>

namespace wesup {

constexpr int UPS_MAX_GROUPS = 5;

// A thread's channel group as it sits in memory (Raw) and as fp32 values (FVec<4>): four channels per thread, i.e.
// 16-byte accesses for fp32 and 8-byte accesses for bf16 -- the narrower bf16 group keeps the register footprint of a
// thread (blended columns of four terms + one column of raw prefetch per term + eight pixels of the streamed term)
// near 120, so sixteen warps stay resident per SM; with eight channels per thread only eight did and the kernel sat
// at 0.2 of the HBM rate, stalled on its own L2 round trips.
template <typename T, int V> struct Raw;
template <> struct Raw<float, 4> {
    typedef uint4 type;
    static __device__ __forceinline__ uint4 zero() { return make_uint4(0u, 0u, 0u, 0u); }
    static __device__ __forceinline__ uint4 load(const float *p) { return __ldg(reinterpret_cast<const uint4 *>(p)); }
    static __device__ __forceinline__ uint4 load_stream(const float *p) {           // read once: keep it out of L1
        uint4 t;
        asm("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(t.x), "=r"(t.y), "=r"(t.z), "=r"(t.w) : "l"(p));
        return t;
    }
    static __device__ __forceinline__ void async_copy(void *smem_dst, const float *p) {     // 16 bytes, L2 -> shared, no L1
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(p) : "memory");
    }
    static __device__ __forceinline__ FVec<4> unpack(uint4 t) {
        FVec<4> r; r.v[0] = __uint_as_float(t.x); r.v[1] = __uint_as_float(t.y); r.v[2] = __uint_as_float(t.z); r.v[3] = __uint_as_float(t.w);
        return r;
    }
};
template <> struct Raw<__nv_bfloat16, 4> {
    typedef uint2 type;
    static __device__ __forceinline__ uint2 zero() { return make_uint2(0u, 0u); }
    static __device__ __forceinline__ uint2 load(const __nv_bfloat16 *p) { return __ldg(reinterpret_cast<const uint2 *>(p)); }
    static __device__ __forceinline__ uint2 load_stream(const __nv_bfloat16 *p) { return ldg_stream(reinterpret_cast<const uint2 *>(p)); }
    static __device__ __forceinline__ void async_copy(void *smem_dst, const __nv_bfloat16 *p) {   // 8 bytes
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(p) : "memory");
    }
    static __device__ __forceinline__ FVec<4> unpack(uint2 t) {
        const float4 a = unpack_bf16x4(t);
        FVec<4> r; r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
        return r;
    }
};

template <> struct Raw<__nv_bfloat16, 8> {       // eight channels = one 16-byte access
    typedef uint4 type;
    static __device__ __forceinline__ uint4 zero() { return make_uint4(0u, 0u, 0u, 0u); }
    static __device__ __forceinline__ uint4 load(const __nv_bfloat16 *p) { return __ldg(reinterpret_cast<const uint4 *>(p)); }
    static __device__ __forceinline__ void async_copy(void *smem_dst, const __nv_bfloat16 *p) {   // 16 bytes, L2 -> shared, no L1
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(p) : "memory");
    }
    static __device__ __forceinline__ FVec<8> unpack(uint4 t) {
        const float4 a = unpack_bf16x4(make_uint2(t.x, t.y)), b = unpack_bf16x4(make_uint2(t.z, t.w));
        FVec<8> r; r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
        return r;
    }
};

constexpr int UPS_LOW = 4;           // low-resolution terms (plus at most one full-resolution term)
constexpr int UPS_SEG = 64;          // output pixels per block (one row segment)
constexpr int UPS_PX = 4;            // pixels per batch of the streamed (full-resolution) term

struct UpsPlan {                     // low-resolution terms first (slot order), the full-resolution term apart
    const void *low[UPS_LOW];
    int h[UPS_LOW], w[UPS_LOW];
    float sy[UPS_LOW], sx[UPS_LOW];
    const void *full;
    int H, W, C;
};

// Instruction budget (r2 profile of the first version: 232 M warp instructions for a 400-px tile, 58 % issue-slot
// busy, 0.2 of the HBM rate -- issue-bound, not memory-bound).  What a pixel costs a thread here:
//   * one 128-bit + one 32-bit shared-memory load for the horizontal taps of ALL low-resolution terms (weights as a
//     float4, "source column advances here" as a bit mask) -- computed once per block;
//   * S0 = bias + sum_s L_s (the sum of the terms' current left columns) lives in registers, so a pixel is
//     acc = f + S0 + sum_s w_s * D_s: one add and NLOW fused multiply-adds per channel, D_s = R_s - L_s;
//   * an advance of term s (every 2nd / 4th / 8th / 16th pixel) costs S0 += D_s, one vertical blend of the prefetched
//     raw column, D_s = R_new - R_old; the number of terms is a template parameter, the mask test skips the block for
//     pixels where nothing advances.
// S0 accumulates one rounding per advance (<= 0.5 ulp each, ~800 along a 400-px row): ~2e-6 relative, far inside the
// 1e-4 fp32 tolerance of the tests.
// V = channels per thread: 4 (16-byte fp32 / 8-byte bf16 accesses, 256 threads) or, for bf16, 8 (16-byte accesses, 128
// threads, more registers): everything a pixel costs a thread apart from the multiply-adds is then paid once per eight
// channels instead of four.
template <typename T, int V, int NLOW>
__global__ void __launch_bounds__(1024 / V, V == 8 ? 3 : 2) upsample_sum_kernel(const UpsPlan G, const float *__restrict__ bias, int relu,
                                                                                T *__restrict__ out) {
    typedef Raw<T, V> RawT;
    typedef typename RawT::type raw_t;
    constexpr int TPB = 1024 / V;
    constexpr int NB = sizeof(T) == 4 ? 2 : 4;                 // batches in flight per thread (cp.async ring in shared memory)
    constexpr int PF = sizeof(T) == 4 || V == 8 ? 1 : 2;            // raw source columns requested ahead per slot (r2 ncu: half of
                                                               // the stalls were the unpack of a column requested ONE advance
                                                               // earlier -- under load the round trip is longer than that)
    // the streamed term never touches registers until it is used: every thread copies ITS OWN channel group of the next
    // NB batches into its own slots with cp.async (no block synchronisation: a thread only reads what it copied)
    __shared__ raw_t s_full[NB * UPS_PX][TPB];
    __shared__ float4 s_w[UPS_SEG];            // weight of column cur+1 for the four slots
    __shared__ unsigned s_adv[UPS_SEG];        // bit s: slot s moves to its next source column AT this pixel
    __shared__ int s_first[UPS_LOW];           // source column of the segment's first pixel
    const int y = blockIdx.y;
    const int x0 = blockIdx.x * UPS_SEG, x1 = min(x0 + UPS_SEG, G.W);
    for (int i = threadIdx.x; i < UPS_SEG; i += blockDim.x) {
        float wv[UPS_LOW] = {0.f, 0.f, 0.f, 0.f};
        unsigned adv = 0u;
#pragma unroll
        for (int s = 0; s < NLOW; ++s) {
            const Tap tx = bilinear_tap(min(x0 + i, G.W - 1), G.sx[s], G.w[s]);
            const int prev = i > 0 ? bilinear_tap(min(x0 + i - 1, G.W - 1), G.sx[s], G.w[s]).i0 : tx.i0;
            wv[s] = tx.w1;
            adv |= (tx.i0 != prev ? 1u : 0u) << s;
            if (i == 0) s_first[s] = tx.i0;
        }
        s_w[i] = make_float4(wv[0], wv[1], wv[2], wv[3]);
        s_adv[i] = adv;
    }
    __syncthreads();
    const int c = threadIdx.x * V;
    if (c >= G.C) return;
    const int C = G.C;
    FVec<V> S0;                                                  // bias + sum of the slots' left columns
#pragma unroll
    for (int k = 0; k < V; ++k) S0.v[k] = bias != nullptr ? __ldg(bias + c + k) : 0.f;
    const T *r0[NLOW > 0 ? NLOW : 1], *r1[NLOW > 0 ? NLOW : 1];
    float wy0[NLOW > 0 ? NLOW : 1], wy1[NLOW > 0 ? NLOW : 1];
    int nxt[NLOW > 0 ? NLOW : 1], wl[NLOW > 0 ? NLOW : 1];      // nxt: last source column requested (cur + 1 + PF, clamped)
    FVec<V> R[NLOW > 0 ? NLOW : 1], D[NLOW > 0 ? NLOW : 1];      // right column (cur+1) and right - left
    raw_t na[NLOW > 0 ? NLOW : 1][PF], nb[NLOW > 0 ? NLOW : 1][PF];
#pragma unroll
    for (int s = 0; s < NLOW; ++s) {
        const T *src = static_cast<const T *>(G.low[s]) + c;
        const Tap ty = bilinear_tap(y, G.sy[s], G.h[s]);
        r0[s] = src + (long)ty.i0 * G.w[s] * C;
        r1[s] = src + (long)ty.i1 * G.w[s] * C;
        wy0[s] = ty.w0; wy1[s] = ty.w1; wl[s] = G.w[s];
        const int i0 = s_first[s];
        const int i1 = min(i0 + 1, wl[s] - 1);
        nxt[s] = i1;
        const FVec<V> a0 = RawT::unpack(RawT::load(r0[s] + (long)i0 * C)), a1 = RawT::unpack(RawT::load(r1[s] + (long)i0 * C));
        const FVec<V> b0 = RawT::unpack(RawT::load(r0[s] + (long)i1 * C)), b1 = RawT::unpack(RawT::load(r1[s] + (long)i1 * C));
#pragma unroll
        for (int d = 0; d < PF; ++d) {
            nxt[s] = min(nxt[s] + 1, wl[s] - 1);
            na[s][d] = RawT::load(r0[s] + (long)nxt[s] * C);
            nb[s][d] = RawT::load(r1[s] + (long)nxt[s] * C);
        }
#pragma unroll
        for (int k = 0; k < V; ++k) {
            const float L = fmaf(ty.w1, a1.v[k], ty.w0 * a0.v[k]);
            R[s].v[k] = fmaf(ty.w1, b1.v[k], ty.w0 * b0.v[k]);
            D[s].v[k] = R[s].v[k] - L;
            S0.v[k] += L;
        }
    }
    const T *full = G.full != nullptr ? static_cast<const T *>(G.full) + ((long)y * G.W) * C + c : nullptr;
    T *o = out + ((long)y * G.W + x0) * C + c;
    auto issue = [&](int batch) {                                  // batch = index of a group of UPS_PX pixels of the segment
        const int xs = x0 + batch * UPS_PX;
        if (full != nullptr) {
#pragma unroll
            for (int j = 0; j < UPS_PX; ++j)
                if (xs + j < x1) RawT::async_copy(&s_full[(batch % NB) * UPS_PX + j][threadIdx.x], full + (long)(xs + j) * C);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
#pragma unroll
    for (int bt = 0; bt < NB; ++bt) issue(bt);
    int batch = 0;
    for (int xb = x0; xb < x1; xb += UPS_PX, ++batch) {
        asm volatile("cp.async.wait_group %0;" ::"n"(NB - 1) : "memory");      // the oldest batch has landed
        FVec<V> f[UPS_PX];
#pragma unroll
        for (int j = 0; j < UPS_PX; ++j)                           // all-zero bits decode to 0 in both formats
            f[j] = RawT::unpack((full != nullptr && xb + j < x1) ? s_full[(batch % NB) * UPS_PX + j][threadIdx.x] : RawT::zero());
        // refill the slots just read: the values above are in registers (unpacked), so the asynchronous writes cannot
        // overtake the reads
        issue(batch + NB);
#pragma unroll
        for (int j = 0; j < UPS_PX; ++j) {
            const int x = xb + j;
            if (x >= x1) break;
            const float4 w4 = s_w[x - x0];
            const unsigned adv = s_adv[x - x0];
            const float wv[UPS_LOW] = {w4.x, w4.y, w4.z, w4.w};
            if (adv != 0u) {                                        // block-uniform
#pragma unroll
                for (int s = 0; s < NLOW; ++s) {
                    if ((adv >> s) & 1u) {
                        // an upsample advances by at most one source column per output pixel: column cur+1 becomes the
                        // left one and the prefetched raw column (in flight since the previous advance) the right one
                        const FVec<V> p0 = RawT::unpack(na[s][0]), p1 = RawT::unpack(nb[s][0]);
#pragma unroll
                        for (int d = 0; d + 1 < PF; ++d) { na[s][d] = na[s][d + 1]; nb[s][d] = nb[s][d + 1]; }
                        nxt[s] = min(nxt[s] + 1, wl[s] - 1);
                        na[s][PF - 1] = RawT::load(r0[s] + (long)nxt[s] * C);
                        nb[s][PF - 1] = RawT::load(r1[s] + (long)nxt[s] * C);
#pragma unroll
                        for (int k = 0; k < V; ++k) {
                            const float rn = fmaf(wy1[s], p1.v[k], wy0[s] * p0.v[k]);
                            S0.v[k] += D[s].v[k];
                            D[s].v[k] = rn - R[s].v[k];
                            R[s].v[k] = rn;
                        }
                    }
                }
            }
            FVec<V> acc;
#pragma unroll
            for (int k = 0; k < V; ++k) acc.v[k] = f[j].v[k] + S0.v[k];
#pragma unroll
            for (int s = 0; s < NLOW; ++s)
#pragma unroll
                for (int k = 0; k < V; ++k) acc.v[k] = fmaf(wv[s], D[s].v[k], acc.v[k]);
            if (relu) {
#pragma unroll
                for (int k = 0; k < V; ++k) acc.v[k] = fmaxf(acc.v[k], 0.f);
            }
            st_group(o, acc);
            o += C;
        }
    }
}

template <typename T, int V>
static void launch_upsample_sum(int n_low, dim3 grid, int threads, cudaStream_t stream, const UpsPlan &G, const float *bias, int relu,
                                void *out) {
    T *o = static_cast<T *>(out);
    switch (n_low) {
    case 0: upsample_sum_kernel<T, V, 0><<<grid, threads, 0, stream>>>(G, bias, relu, o); break;
    case 1: upsample_sum_kernel<T, V, 1><<<grid, threads, 0, stream>>>(G, bias, relu, o); break;
    case 2: upsample_sum_kernel<T, V, 2><<<grid, threads, 0, stream>>>(G, bias, relu, o); break;
    case 3: upsample_sum_kernel<T, V, 3><<<grid, threads, 0, stream>>>(G, bias, relu, o); break;
    default: upsample_sum_kernel<T, V, 4><<<grid, threads, 0, stream>>>(G, bias, relu, o); break;
    }
}

}  // namespace wesup

using namespace wesup;

extern "C" int wesup_upsample_sum(const void *const *z, const int *h, const int *w, int n_terms, int H, int W, int C, int dtype,
                                  const float *bias, int relu, void *out, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    WESUP_REQUIRE(z && h && w && out, WESUP_E_ARG, "wesup_upsample_sum: null pointer");
    WESUP_REQUIRE(n_terms >= 1 && n_terms <= UPS_MAX_GROUPS, WESUP_E_UNSUPPORTED, "wesup_upsample_sum: 1..%d terms (got %d)", UPS_MAX_GROUPS, n_terms);
    WESUP_REQUIRE(H > 0 && W > 0 && C > 0, WESUP_E_ARG, "wesup_upsample_sum: bad size H=%d W=%d C=%d", H, W, C);
    WESUP_REQUIRE(dtype == WESUP_F32 || dtype == WESUP_BF16, WESUP_E_ARG, "wesup_upsample_sum: bad dtype %d", dtype);
    const int V = 4;
    WESUP_REQUIRE(C % V == 0 && C / V <= 256, WESUP_E_UNSUPPORTED, "wesup_upsample_sum: C must be a multiple of %d and at most %d (C=%d)", V, 256 * V, C);
    for (int g = 0; g < n_terms; ++g)
        WESUP_REQUIRE(h[g] == H && w[g] == W ? true : (h[g] <= H && w[g] <= W && (long)(W - 1) >= (long)(w[g] - 1)), WESUP_E_UNSUPPORTED,
                      "wesup_upsample_sum: term %d is larger than the output (only upsampling is supported)", g);
    WESUP_REQUIRE(aligned16(out) && (bias == nullptr || aligned16(bias)), WESUP_E_ALIGN, "wesup_upsample_sum: out/bias must be 16-byte aligned");
    int n_full = 0;
    for (int g = 0; g < n_terms; ++g) n_full += (h[g] == H && w[g] == W) ? 1 : 0;
    WESUP_REQUIRE(n_full <= 1 && n_terms - n_full <= UPS_LOW, WESUP_E_UNSUPPORTED,
                  "wesup_upsample_sum: at most one full-resolution and %d low-resolution terms", UPS_LOW);
    UpsPlan G;
    G.H = H; G.W = W; G.C = C; G.full = nullptr;
    int n_low = 0;
    for (int s = 0; s < UPS_LOW; ++s) { G.low[s] = nullptr; G.h[s] = 1; G.w[s] = 1; G.sy[s] = 0.f; G.sx[s] = 0.f; }
    for (int g = 0; g < n_terms; ++g) {
        WESUP_REQUIRE(z[g] && aligned16(z[g]), WESUP_E_ALIGN, "wesup_upsample_sum: term %d is null or not 16-byte aligned", g);
        WESUP_REQUIRE(h[g] > 0 && w[g] > 0 && h[g] <= H && w[g] <= W, WESUP_E_ARG, "wesup_upsample_sum: term %d has size %dx%d", g, h[g], w[g]);
        if (h[g] == H && w[g] == W) { G.full = z[g]; continue; }
        G.low[n_low] = z[g]; G.h[n_low] = h[g]; G.w[n_low] = w[g];
        G.sy[n_low] = bilinear_scale(h[g], H); G.sx[n_low] = bilinear_scale(w[g], W);
        ++n_low;
    }
    dim3 grid(cdiv(W, UPS_SEG), H);
    const char *v4 = getenv("WESUP_UPS_V4");                       // cross-check / A-B of the bf16 kernels
    if (dtype == WESUP_F32) launch_upsample_sum<float, 4>(n_low, grid, C / 4, stream, G, bias, relu, out);
    else if (C % 8 == 0 && !(v4 && v4[0] == '1')) launch_upsample_sum<__nv_bfloat16, 8>(n_low, grid, C / 8, stream, G, bias, relu, out);
    else launch_upsample_sum<__nv_bfloat16, 4>(n_low, grid, C / 4, stream, G, bias, relu, out);
    WESUP_CHECK_LAUNCH("wesup_upsample_sum", 1);
    return 0;
}
