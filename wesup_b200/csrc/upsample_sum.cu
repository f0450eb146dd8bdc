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

namespace wesup {

constexpr int UPS_MAX_GROUPS = 5;

struct UpsGroups {
    const void *src[UPS_MAX_GROUPS];
    int h[UPS_MAX_GROUPS], w[UPS_MAX_GROUPS];
    float sy[UPS_MAX_GROUPS], sx[UPS_MAX_GROUPS];
    int n, H, W, C;
};

// A thread's channel group as it sits in memory (Raw) and as fp32 values (FVec<4>): four channels per thread, i.e.
// 16-byte accesses for fp32 and 8-byte accesses for bf16 -- the narrower bf16 group keeps the register footprint of a
// thread (blended columns of four terms + one column of raw prefetch per term + eight pixels of the streamed term)
// near 120, so sixteen warps stay resident per SM; with eight channels per thread only eight did and the kernel sat
// at 0.2 of the HBM rate, stalled on its own L2 round trips.
template <typename T> struct Raw;
template <> struct Raw<float> {
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
template <> struct Raw<__nv_bfloat16> {
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

constexpr int UPS_V = 4;             // channels per thread
constexpr int UPS_LOW = 4;           // low-resolution terms (plus at most one full-resolution term)

constexpr int UPS_SEG = 64;          // output pixels per block (one row segment)

template <typename T>
__global__ void __launch_bounds__(256, 2) upsample_sum_kernel(const UpsGroups G, const float *__restrict__ bias, int relu,
                                                              T *__restrict__ out) {
    typedef typename Raw<T>::type raw_t;
    constexpr int V = UPS_V;
    constexpr int UPS_PX = 4;                                  // pixels per batch of the streamed (full-resolution) term
    constexpr int UPS_NB = sizeof(raw_t) == 16 ? 2 : 4;        // batches in flight per thread (cp.async ring in shared memory)
    // the streamed term never touches registers until it is used: every thread copies ITS OWN channel group of the next
    // UPS_NB batches into its own slots with cp.async (no block synchronisation: a thread only reads what it copied), so
    // 32 KB per block are in flight all the time instead of bursts of register loads
    __shared__ raw_t s_full[UPS_NB * UPS_PX][256];
    // horizontal taps of the segment, computed once per block and shared by all channel groups:
    // {weight of column cur+1, 1 when the source column advances at this pixel}
    __shared__ float2 s_tap[UPS_LOW][UPS_SEG];
    __shared__ int s_first[UPS_LOW];           // source column of the segment's first pixel
    const int y = blockIdx.y;
    const int x0 = blockIdx.x * UPS_SEG, x1 = min(x0 + UPS_SEG, G.W);
    {
        int slot = 0;
#pragma unroll
        for (int g = 0; g < UPS_MAX_GROUPS; ++g) {
            if (g >= G.n || (G.h[g] == G.H && G.w[g] == G.W)) continue;
            if (slot < UPS_LOW) {
                for (int i = threadIdx.x; i < UPS_SEG; i += blockDim.x) {
                    const Tap tx = bilinear_tap(min(x0 + i, G.W - 1), G.sx[g], G.w[g]);
                    const int prev = i > 0 ? bilinear_tap(min(x0 + i - 1, G.W - 1), G.sx[g], G.w[g]).i0 : tx.i0;
                    s_tap[slot][i] = make_float2(tx.w1, tx.i0 != prev ? 1.f : 0.f);
                    if (i == 0) s_first[slot] = tx.i0;
                }
            }
            ++slot;
        }
    }
    __syncthreads();
    const int c = threadIdx.x * V;
    if (c >= G.C) return;
    const int C = G.C;
    FVec<V> b;
#pragma unroll
    for (int k = 0; k < V; ++k) b.v[k] = bias != nullptr ? __ldg(bias + c + k) : 0.f;
    // the (at most one) full-resolution term is streamed, UPS_PX pixels in flight; every low-resolution term keeps its
    // two source rows, the vertically blended column cur and the difference to column cur+1 in registers, and the RAW
    // rows of column cur+2, requested one advance ahead of their first use so the L2 round trip overlaps the walk
    const T *full = nullptr;
    const T *r0[UPS_LOW], *r1[UPS_LOW];
    float wy0[UPS_LOW], wy1[UPS_LOW];
    int cur[UPS_LOW], wl[UPS_LOW];
    FVec<V> c0[UPS_LOW], dc[UPS_LOW];           // column cur, and (column cur+1) - (column cur)
    raw_t na[UPS_LOW], nb[UPS_LOW];
    int n_low = 0;
#pragma unroll
    for (int g = 0; g < UPS_MAX_GROUPS; ++g) {
        if (g >= G.n) continue;
        const T *src = static_cast<const T *>(G.src[g]) + c;
        if (G.h[g] == G.H && G.w[g] == G.W) {
            full = src + ((long)y * G.W) * C;
            continue;
        }
#pragma unroll
        for (int s = 0; s < UPS_LOW; ++s) {
            if (s != n_low) continue;                           // slot = running count of low-resolution terms
            const Tap ty = bilinear_tap(y, G.sy[g], G.h[g]);
            r0[s] = src + (long)ty.i0 * G.w[g] * C;
            r1[s] = src + (long)ty.i1 * G.w[g] * C;
            wy0[s] = ty.w0; wy1[s] = ty.w1; wl[s] = G.w[g];
            const int i0 = s_first[s];
            cur[s] = i0;
            const int i1 = min(i0 + 1, wl[s] - 1), i2 = min(i0 + 2, wl[s] - 1);
            const FVec<V> a0 = Raw<T>::unpack(Raw<T>::load(r0[s] + (long)i0 * C)), a1 = Raw<T>::unpack(Raw<T>::load(r1[s] + (long)i0 * C));
            const FVec<V> b0 = Raw<T>::unpack(Raw<T>::load(r0[s] + (long)i1 * C)), b1 = Raw<T>::unpack(Raw<T>::load(r1[s] + (long)i1 * C));
            na[s] = Raw<T>::load(r0[s] + (long)i2 * C);
            nb[s] = Raw<T>::load(r1[s] + (long)i2 * C);
#pragma unroll
            for (int k = 0; k < V; ++k) {
                c0[s].v[k] = fmaf(ty.w1, a1.v[k], ty.w0 * a0.v[k]);
                dc[s].v[k] = fmaf(ty.w1, b1.v[k], ty.w0 * b0.v[k]) - c0[s].v[k];
            }
        }
        ++n_low;
    }
    T *o = out + ((long)y * G.W + x0) * C + c;
    auto issue = [&](int batch) {                                  // batch = index of a group of UPS_PX pixels of the segment
        const int xs = x0 + batch * UPS_PX;
        if (full != nullptr) {
#pragma unroll
            for (int j = 0; j < UPS_PX; ++j)
                if (xs + j < x1) Raw<T>::async_copy(&s_full[(batch % UPS_NB) * UPS_PX + j][threadIdx.x], full + (long)(xs + j) * C);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
#pragma unroll
    for (int bt = 0; bt < UPS_NB; ++bt) issue(bt);
    int batch = 0;
    for (int xb = x0; xb < x1; xb += UPS_PX, ++batch) {
        asm volatile("cp.async.wait_group %0;" ::"n"(UPS_NB - 1) : "memory");      // the oldest batch has landed
        FVec<V> f[UPS_PX];
#pragma unroll
        for (int j = 0; j < UPS_PX; ++j)                           // all-zero bits decode to 0 in both formats
            f[j] = Raw<T>::unpack((full != nullptr && xb + j < x1) ? s_full[(batch % UPS_NB) * UPS_PX + j][threadIdx.x] : Raw<T>::zero());
        // refill the slots just read: the values above are in registers (unpacked), so the asynchronous writes cannot
        // overtake the reads
        issue(batch + UPS_NB);
#pragma unroll
        for (int j = 0; j < UPS_PX; ++j) {
            const int x = xb + j;
            if (x >= x1) break;
            FVec<V> acc = f[j];
#pragma unroll
            for (int k = 0; k < V; ++k) acc.v[k] += b.v[k];
#pragma unroll
            for (int s = 0; s < UPS_LOW; ++s) {
                if (s >= n_low) continue;
                const float2 tp = s_tap[s][x - x0];
                if (tp.y != 0.f) {
                    // an upsample advances by at most one source column per output pixel: column cur+1 becomes cur and
                    // column cur+2 (raw, in flight since the previous advance) becomes cur+1
                    const FVec<V> p0 = Raw<T>::unpack(na[s]), p1 = Raw<T>::unpack(nb[s]);
                    cur[s] += 1;
#pragma unroll
                    for (int k = 0; k < V; ++k) {
                        c0[s].v[k] += dc[s].v[k];
                        dc[s].v[k] = fmaf(wy1[s], p1.v[k], wy0[s] * p0.v[k]) - c0[s].v[k];
                    }
                    const int i2 = min(cur[s] + 2, wl[s] - 1);
                    na[s] = Raw<T>::load(r0[s] + (long)i2 * C);
                    nb[s] = Raw<T>::load(r1[s] + (long)i2 * C);
                }
#pragma unroll
                for (int k = 0; k < V; ++k) acc.v[k] += fmaf(tp.x, dc[s].v[k], c0[s].v[k]);
            }
            if (relu) {
#pragma unroll
                for (int k = 0; k < V; ++k) acc.v[k] = fmaxf(acc.v[k], 0.f);
            }
            st_group(o, acc);
            o += C;
        }
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
    UpsGroups G;
    G.n = n_terms; G.H = H; G.W = W; G.C = C;
    for (int g = 0; g < n_terms; ++g) {
        WESUP_REQUIRE(z[g] && aligned16(z[g]), WESUP_E_ALIGN, "wesup_upsample_sum: term %d is null or not 16-byte aligned", g);
        WESUP_REQUIRE(h[g] > 0 && w[g] > 0 && h[g] <= H && w[g] <= W, WESUP_E_ARG, "wesup_upsample_sum: term %d has size %dx%d", g, h[g], w[g]);
        G.src[g] = z[g]; G.h[g] = h[g]; G.w[g] = w[g];
        G.sy[g] = bilinear_scale(h[g], H); G.sx[g] = bilinear_scale(w[g], W);
    }
    int n_full = 0;
    for (int g = 0; g < n_terms; ++g) n_full += (h[g] == H && w[g] == W) ? 1 : 0;
    WESUP_REQUIRE(n_full <= 1 && n_terms - n_full <= UPS_LOW, WESUP_E_UNSUPPORTED,
                  "wesup_upsample_sum: at most one full-resolution and %d low-resolution terms", UPS_LOW);
    const int seg = UPS_SEG;
    const int threads = C / V;
    dim3 grid(cdiv(W, seg), H);
    if (dtype == WESUP_F32)
        upsample_sum_kernel<float><<<grid, threads, 0, stream>>>(G, bias, relu, static_cast<float *>(out));
    else
        upsample_sum_kernel<__nv_bfloat16><<<grid, threads, 0, stream>>>(G, bias, relu, static_cast<__nv_bfloat16 *>(out));
    WESUP_CHECK_LAUNCH("wesup_upsample_sum", 1);
    return 0;
}
