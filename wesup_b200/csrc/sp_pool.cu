// Superpixel statistics, mean pooling (fwd/bwd) and painting over an integer
// label map.  Replaces the reference's dense one-hot formulation:
//   _preprocess_superpixels   /root/reference/models/wesup.py:18-63
//   torch.mm(sp_maps, x.t())  /root/reference/models/wesup.py:284-285 (+ autograd adjoint)
//   argmax + paint loop       /root/reference/models/wesup.py:295-304
//
// Data layout in HBM: the label map is int32 (H*W); superpixels are rows in the
// reference's order (labeled ids ascending, then unlabeled ascending).  Pooling
// reads a CSR of pixel ids per row (built once per image, deterministic), so
// the hot loop is an atomic-free gather-sum of contiguous channel vectors in
// the pixel-major (H*W,C) layout: every feature byte is read exactly once.
#include "common.cuh"

namespace wesup {

// ---------------------------------------------------------------------------
// statistics
// ---------------------------------------------------------------------------
struct StatsWs {
    int32_t *cnt_raw;   // n_sp   pixels per original id
    int32_t *bbox;      // n_sp*4 ymin,ymax,xmin,xmax per original id
    int32_t *cls_cnt;   // n_sp*n_cls
    int32_t *rank;      // n_sp   original id -> row
};

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

static StatsWs carve_stats(void *ws, int n_sp, int n_cls) {
    char *p = static_cast<char *>(ws);
    StatsWs s;
    s.cnt_raw = reinterpret_cast<int32_t *>(p); p += align_up(sizeof(int32_t) * n_sp, 256);
    s.bbox = reinterpret_cast<int32_t *>(p);    p += align_up(sizeof(int32_t) * 4 * n_sp, 256);
    s.cls_cnt = reinterpret_cast<int32_t *>(p); p += align_up(sizeof(int32_t) * n_sp * (n_cls > 0 ? n_cls : 1), 256);
    s.rank = reinterpret_cast<int32_t *>(p);
    return s;
}

__global__ void stats_init_kernel(StatsWs s, int n_sp, int n_cls, int H, int W) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_sp) {
        s.cnt_raw[i] = 0;
        s.bbox[4 * i + 0] = H; s.bbox[4 * i + 1] = -1;
        s.bbox[4 * i + 2] = W; s.bbox[4 * i + 3] = -1;
    }
    if (i < n_sp * n_cls) s.cls_cnt[i] = 0;
}

// One thread per pixel.  Lanes of a warp that hold the same id (runs along a
// row) elect a leader that issues one count atomic for the group; bounding-box
// atomics are skipped when the racy pre-read shows the pixel is interior.
__global__ void stats_accumulate_kernel(const int32_t *__restrict__ labels, const int64_t *__restrict__ mask,
                                        StatsWs s, int H, int W, int n_cls, int n_sp) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    int HW = H * W;
    bool live = p < HW;
    int id = live ? labels[p] : -1;
    if (id < 0 || id >= n_sp) { live = false; id = -1; }
    unsigned peers = __match_any_sync(0xffffffffu, id);
    if (live) {
        int lane = threadIdx.x & 31;
        if ((peers & ((1u << lane) - 1)) == 0) atomicAdd(&s.cnt_raw[id], __popc(peers));
        int y = p / W, x = p - y * W;
        volatile int32_t *bb = s.bbox + 4 * id;
        if (y < bb[0]) atomicMin(&s.bbox[4 * id + 0], y);
        if (y > bb[1]) atomicMax(&s.bbox[4 * id + 1], y);
        if (x < bb[2]) atomicMin(&s.bbox[4 * id + 2], x);
        if (x > bb[3]) atomicMax(&s.bbox[4 * id + 3], x);
        if (mask != nullptr) {
            for (int c = 0; c < n_cls; ++c) {
                long m = mask[(long)c * HW + p];
                if (m != 0) atomicAdd(&s.cls_cnt[id * n_cls + c], (int)m);
            }
        }
    }
}

__device__ __forceinline__ int block_exclusive_scan(int v, int *smem /* >= 33 ints */, int &total) {
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    __syncthreads();
    if (lane == 31) smem[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = lane < nwarp ? smem[lane] : 0;
        int wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += t;
        }
        smem[lane] = wi - w;
        if (lane == 31) smem[32] = wi;
    }
    __syncthreads();
    total = smem[32];
    return smem[warp] + incl - v;
}

// Single block.  Row order, quantised labels, CSR offsets.
__global__ void __launch_bounds__(1024) stats_order_kernel(StatsWs s, int n_sp, int n_cls, bool has_mask,
                                                           int32_t *order, int32_t *counts, int32_t *seg_offsets,
                                                           float *sp_labels, int32_t *n_labeled_out) {
    __shared__ int scan_smem[33];
    __shared__ int s_n_labeled;
    const int tid = threadIdx.x, nt = blockDim.x;
    // pass 1: labeled ids first (ascending)
    int base = 0;
    for (int start = 0; start < n_sp; start += nt) {
        int id = start + tid;
        int flag = 0;
        if (id < n_sp && has_mask) {
            for (int c = 0; c < n_cls; ++c) flag |= (s.cls_cnt[id * n_cls + c] > 0);
        }
        int total;
        int pos = block_exclusive_scan(flag, scan_smem, total);
        if (flag) { s.rank[id] = base + pos; order[base + pos] = id; }
        base += total;
        __syncthreads();
    }
    if (tid == 0) { s_n_labeled = base; *n_labeled_out = base; }
    __syncthreads();
    const int n_l = s_n_labeled;
    // pass 2: unlabeled ids (ascending)
    base = n_l;
    for (int start = 0; start < n_sp; start += nt) {
        int id = start + tid;
        int flag = 0;
        if (id < n_sp) {
            flag = 1;
            if (has_mask)
                for (int c = 0; c < n_cls; ++c) flag &= (s.cls_cnt[id * n_cls + c] <= 0);
        }
        int total;
        int pos = block_exclusive_scan(flag, scan_smem, total);
        if (flag) { s.rank[id] = base + pos; order[base + pos] = id; }
        base += total;
        __syncthreads();
    }
    __threadfence_block();
    __syncthreads();
    // pass 3: counts in row order, CSR offsets, quantised labels
    base = 0;
    for (int start = 0; start < n_sp; start += nt) {
        int k = start + tid;
        int c = 0;
        if (k < n_sp) {
            int id = order[k];
            c = s.cnt_raw[id];
            counts[k] = c;
            if (sp_labels != nullptr) {
                int best = 0;
                for (int j = 0; j < n_cls; ++j) best = max(best, s.cls_cnt[id * n_cls + j]);
                for (int j = 0; j < n_cls; ++j)
                    sp_labels[k * n_cls + j] = (k < n_l && s.cls_cnt[id * n_cls + j] == best) ? 1.0f : 0.0f;
            }
        }
        int total;
        int pos = block_exclusive_scan(c, scan_smem, total);
        if (k < n_sp) seg_offsets[k] = base + pos;
        base += total;
        __syncthreads();
    }
    if (tid == 0) seg_offsets[n_sp] = base;
}

__global__ void relabel_kernel(const int32_t *__restrict__ labels, const int32_t *__restrict__ rank,
                               int32_t *__restrict__ row_labels, int HW, int n_sp) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < HW) {
        int id = labels[p];
        row_labels[p] = (id >= 0 && id < n_sp) ? rank[id] : -1;
    }
}

// One warp per row: scan the superpixel's bounding box in raster order and
// compact the matching pixel ids (ordered, no atomics => deterministic CSR).
__global__ void csr_fill_kernel(const int32_t *__restrict__ labels, StatsWs s, const int32_t *__restrict__ order,
                                const int32_t *__restrict__ seg_offsets, int32_t *__restrict__ seg_pixels,
                                int W, int n_sp) {
    int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (k >= n_sp) return;
    int id = order[k];
    int y0 = s.bbox[4 * id + 0], y1 = s.bbox[4 * id + 1];
    int x0 = s.bbox[4 * id + 2], x1 = s.bbox[4 * id + 3];
    if (y1 < y0) return;                       // id absent from the map
    int bw = x1 - x0 + 1;
    int total = bw * (y1 - y0 + 1);
    int out = seg_offsets[k];
    for (int t = 0; t < total; t += 32) {
        int i = t + lane;
        bool hit = false;
        int p = 0;
        if (i < total) {
            int yy = i / bw;
            p = (y0 + yy) * W + x0 + (i - yy * bw);
            hit = (__ldg(labels + p) == id);
        }
        unsigned m = __ballot_sync(0xffffffffu, hit);
        if (hit) seg_pixels[out + __popc(m & ((1u << lane) - 1))] = p;
        out += __popc(m);
    }
}

// ---------------------------------------------------------------------------
// pooling forward, pixel-major: one thread per (row, 4 channels)
// ---------------------------------------------------------------------------
template <typename T, int UNROLL>
__global__ void __launch_bounds__(256) pool_fwd_hwc_kernel(const T *__restrict__ feat, const int32_t *__restrict__ seg_offsets,
                                                          const int32_t *__restrict__ seg_pixels, int C, int C4, long n_items,
                                                          float *__restrict__ pooled) {
    long item = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= n_items) return;
    int k = (int)(item / C4);
    int c = ((int)(item - (long)k * C4)) << 2;
    int beg = __ldg(seg_offsets + k), end = __ldg(seg_offsets + k + 1);
    const T *base = feat + c;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int i = beg;
    for (; i + UNROLL <= end; i += UNROLL) {
        int px[UNROLL];
#pragma unroll
        for (int j = 0; j < UNROLL; ++j) px[j] = __ldg(seg_pixels + i + j);
        float4 v[UNROLL];
#pragma unroll
        for (int j = 0; j < UNROLL; ++j) v[j] = Vec4<T>::load(base + (long)px[j] * C);
#pragma unroll
        for (int j = 0; j < UNROLL; ++j) acc = acc + v[j];
    }
    for (; i < end; ++i) acc = acc + Vec4<T>::load(base + (long)__ldg(seg_pixels + i) * C);
    int n = end - beg;
    float inv = n > 0 ? 1.0f / (float)n : 0.0f;
    *reinterpret_cast<float4 *>(pooled + (long)k * C + c) = inv * acc;
}

// channel-major: one warp per (row, channel); fixed-shape shuffle tree
template <typename T>
__global__ void __launch_bounds__(256) pool_fwd_chw_kernel(const T *__restrict__ feat, const int32_t *__restrict__ seg_offsets,
                                                          const int32_t *__restrict__ seg_pixels, long HW, int C, int N,
                                                          float *__restrict__ pooled) {
    long warp = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (warp >= (long)N * C) return;
    int k = (int)(warp / C), c = (int)(warp - (long)k * C);
    int beg = __ldg(seg_offsets + k), end = __ldg(seg_offsets + k + 1);
    const T *plane = feat + (long)c * HW;
    float acc = 0.f;
    for (int i = beg + lane; i < end; i += 32) acc += (float)plane[__ldg(seg_pixels + i)];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
        int n = end - beg;
        pooled[(long)k * C + c] = n > 0 ? acc * (1.0f / (float)n) : 0.f;
    }
}

// ---------------------------------------------------------------------------
// pooling backward
// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) pool_bwd_hwc_kernel(const float *__restrict__ grad_pooled, const int32_t *__restrict__ row_labels,
                                                          const int32_t *__restrict__ counts, int C, int C4, long n_items,
                                                          T *__restrict__ grad_feat) {
    long item = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= n_items) return;
    long p = item / C4;
    int c = ((int)(item - p * C4)) << 2;
    int k = __ldg(row_labels + p);
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    if (k >= 0) {
        float inv = 1.0f / (float)__ldg(counts + k);
        g = inv * __ldg(reinterpret_cast<const float4 *>(grad_pooled + (long)k * C + c));
    }
    Vec4<T>::store(grad_feat + p * C + c, g);
}

// Fast path: block = a run of consecutive pixels, thread = one 16-byte channel
// group.  The (pre-scaled) pooled-gradient row is re-fetched only when the
// label changes along the walk (superpixels are ~14 px wide), so the kernel is
// a pure coalesced write stream of C*sizeof(T) bytes per pixel.
template <typename T, int V>
__global__ void __launch_bounds__(544) pool_bwd_walk_kernel(const float *__restrict__ grad_pooled, const int32_t *__restrict__ row_labels,
                                                           const int32_t *__restrict__ counts, int C, int HW, int seg,
                                                           T *__restrict__ grad_feat) {
    const int c = threadIdx.x * V;
    if (c >= C) return;
    const int p0 = blockIdx.x * seg, p1 = min(p0 + seg, HW);
    T *o = grad_feat + (long)p0 * C + c;
    int cur = -2;
    FVec<V> g;
    for (int p = p0; p < p1; ++p, o += C) {
        const int k = __ldg(row_labels + p);
        if (k != cur) {
            cur = k;
            if (k >= 0) {
                const float inv = 1.0f / (float)__ldg(counts + k);
                g = ld_group<V>(grad_pooled + (long)k * C + c);
#pragma unroll
                for (int j = 0; j < V; ++j) g.v[j] *= inv;
            } else {
#pragma unroll
                for (int j = 0; j < V; ++j) g.v[j] = 0.f;
            }
        }
        st_group(o, g);
    }
}

template <typename T>
__global__ void __launch_bounds__(256) pool_bwd_chw_kernel(const float *__restrict__ grad_pooled, const int32_t *__restrict__ row_labels,
                                                          const int32_t *__restrict__ counts, long HW, int C,
                                                          T *__restrict__ grad_feat) {
    long p = (long)blockIdx.x * blockDim.x + threadIdx.x;
    int c = blockIdx.y;
    if (p >= HW) return;
    int k = __ldg(row_labels + p);
    float g = 0.f;
    if (k >= 0) g = __ldg(grad_pooled + (long)k * C + c) * (1.0f / (float)__ldg(counts + k));
    grad_feat[(long)c * HW + p] = (T)g;
}

__global__ void paint_kernel(const int32_t *__restrict__ row_labels, const float *__restrict__ sp_pred, int HW,
                             int n_cls, int cls, float *__restrict__ out) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < HW) {
        int k = __ldg(row_labels + p);
        out[p] = k >= 0 ? __ldg(sp_pred + (long)k * n_cls + cls) : 0.f;
    }
}

}  // namespace wesup

using namespace wesup;

extern "C" size_t wesup_sp_stats_workspace_bytes(int H, int W, int n_sp, int n_cls) {
    (void)H; (void)W;
    if (n_sp <= 0) return 0;
    int nc = n_cls > 0 ? n_cls : 1;
    return align_up(sizeof(int32_t) * n_sp, 256) + align_up(sizeof(int32_t) * 4 * n_sp, 256) +
           align_up(sizeof(int32_t) * n_sp * nc, 256) + align_up(sizeof(int32_t) * n_sp, 256);
}

extern "C" int wesup_sp_stats(const int32_t *labels, const int64_t *mask, int H, int W, int n_cls, int n_sp,
                              int32_t *order, int32_t *row_labels, int32_t *counts, int32_t *seg_offsets,
                              int32_t *seg_pixels, float *sp_labels, int32_t *n_labeled, void *ws, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    WESUP_REQUIRE(labels && order && row_labels && counts && seg_offsets && seg_pixels && n_labeled && ws,
                  WESUP_E_ARG, "wesup_sp_stats: null pointer");
    WESUP_REQUIRE(H > 0 && W > 0 && n_sp > 0 && n_cls >= 0, WESUP_E_ARG, "wesup_sp_stats: bad size H=%d W=%d n_sp=%d", H, W, n_sp);
    WESUP_REQUIRE((long)H * W < (1L << 31), WESUP_E_UNSUPPORTED, "wesup_sp_stats: H*W must fit int32");
    WESUP_REQUIRE(mask == nullptr || (n_cls > 0 && sp_labels != nullptr), WESUP_E_ARG, "wesup_sp_stats: mask needs n_cls>0 and sp_labels");
    const int HW = H * W;
    StatsWs s = carve_stats(ws, n_sp, n_cls);
    int n_init = n_sp * (n_cls > 1 ? n_cls : 1);
    stats_init_kernel<<<cdiv(n_init, 256), 256, 0, stream>>>(s, n_sp, n_cls, H, W);
    stats_accumulate_kernel<<<cdiv(HW, 256), 256, 0, stream>>>(labels, mask, s, H, W, n_cls, n_sp);
    stats_order_kernel<<<1, 1024, 0, stream>>>(s, n_sp, n_cls, mask != nullptr, order, counts, seg_offsets,
                                              mask != nullptr ? sp_labels : nullptr, n_labeled);
    relabel_kernel<<<cdiv(HW, 256), 256, 0, stream>>>(labels, s.rank, row_labels, HW, n_sp);
    csr_fill_kernel<<<cdiv((long)n_sp * 32, 256), 256, 0, stream>>>(labels, s, order, seg_offsets, seg_pixels, W, n_sp);
    WESUP_CHECK_LAUNCH("wesup_sp_stats", 5);
    return 0;
}

extern "C" int wesup_sp_pool_fwd(const void *feat, int dtype, int layout, const int32_t *seg_offsets,
                                 const int32_t *seg_pixels, int HW, int C, int N, float *pooled, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    WESUP_REQUIRE(feat && seg_offsets && seg_pixels && pooled, WESUP_E_ARG, "wesup_sp_pool_fwd: null pointer");
    WESUP_REQUIRE(HW > 0 && C > 0 && N > 0, WESUP_E_ARG, "wesup_sp_pool_fwd: bad size HW=%d C=%d N=%d", HW, C, N);
    WESUP_REQUIRE(dtype == WESUP_F32 || dtype == WESUP_BF16, WESUP_E_ARG, "wesup_sp_pool_fwd: bad dtype %d", dtype);
    if (layout == WESUP_HWC) {
        WESUP_REQUIRE(C % 4 == 0, WESUP_E_ALIGN, "wesup_sp_pool_fwd: HWC layout needs C %% 4 == 0 (C=%d)", C);
        WESUP_REQUIRE(aligned16(feat) && aligned16(pooled), WESUP_E_ALIGN, "wesup_sp_pool_fwd: feat/pooled must be 16-byte aligned");
        int C4 = C / 4;
        long n_items = (long)N * C4;
        int grid = cdiv(n_items, 256);
        if (dtype == WESUP_F32)
            pool_fwd_hwc_kernel<float, 8><<<grid, 256, 0, stream>>>((const float *)feat, seg_offsets, seg_pixels, C, C4, n_items, pooled);
        else
            pool_fwd_hwc_kernel<__nv_bfloat16, 8><<<grid, 256, 0, stream>>>((const __nv_bfloat16 *)feat, seg_offsets, seg_pixels, C, C4, n_items, pooled);
    } else if (layout == WESUP_CHW) {
        long n_threads = (long)N * C * 32;
        int grid = cdiv(n_threads, 256);
        if (dtype == WESUP_F32)
            pool_fwd_chw_kernel<float><<<grid, 256, 0, stream>>>((const float *)feat, seg_offsets, seg_pixels, HW, C, N, pooled);
        else
            pool_fwd_chw_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>((const __nv_bfloat16 *)feat, seg_offsets, seg_pixels, HW, C, N, pooled);
    } else {
        WESUP_REQUIRE(false, WESUP_E_ARG, "wesup_sp_pool_fwd: bad layout %d", layout);
    }
    WESUP_CHECK_LAUNCH("wesup_sp_pool_fwd", 1);
    return 0;
}

extern "C" int wesup_sp_pool_bwd(const float *grad_pooled, const int32_t *row_labels, const int32_t *counts, int HW,
                                 int C, int N, void *grad_feat, int dtype, int layout, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    WESUP_REQUIRE(grad_pooled && row_labels && counts && grad_feat, WESUP_E_ARG, "wesup_sp_pool_bwd: null pointer");
    WESUP_REQUIRE(HW > 0 && C > 0 && N > 0, WESUP_E_ARG, "wesup_sp_pool_bwd: bad size HW=%d C=%d N=%d", HW, C, N);
    WESUP_REQUIRE(dtype == WESUP_F32 || dtype == WESUP_BF16, WESUP_E_ARG, "wesup_sp_pool_bwd: bad dtype %d", dtype);
    if (layout == WESUP_HWC) {
        WESUP_REQUIRE(C % 4 == 0, WESUP_E_ALIGN, "wesup_sp_pool_bwd: HWC layout needs C %% 4 == 0 (C=%d)", C);
        WESUP_REQUIRE(aligned16(grad_feat) && aligned16(grad_pooled), WESUP_E_ALIGN, "wesup_sp_pool_bwd: buffers must be 16-byte aligned");
        int C4 = C / 4;
        long n_items = (long)HW * C4;
        long grid = (n_items + 255) / 256;
        WESUP_REQUIRE(grid < (1L << 31), WESUP_E_UNSUPPORTED, "wesup_sp_pool_bwd: grid too large");
        const int V = dtype == WESUP_F32 ? 4 : 8;
        if (C % V == 0 && C / V <= 544) {
            const int seg = 64, threads = (C / V + 31) / 32 * 32;
            if (dtype == WESUP_F32)
                pool_bwd_walk_kernel<float, 4><<<cdiv(HW, seg), threads, 0, stream>>>(grad_pooled, row_labels, counts, C, HW, seg, (float *)grad_feat);
            else
                pool_bwd_walk_kernel<__nv_bfloat16, 8><<<cdiv(HW, seg), threads, 0, stream>>>(grad_pooled, row_labels, counts, C, HW, seg, (__nv_bfloat16 *)grad_feat);
        } else if (dtype == WESUP_F32)
            pool_bwd_hwc_kernel<float><<<(unsigned)grid, 256, 0, stream>>>(grad_pooled, row_labels, counts, C, C4, n_items, (float *)grad_feat);
        else
            pool_bwd_hwc_kernel<__nv_bfloat16><<<(unsigned)grid, 256, 0, stream>>>(grad_pooled, row_labels, counts, C, C4, n_items, (__nv_bfloat16 *)grad_feat);
    } else if (layout == WESUP_CHW) {
        WESUP_REQUIRE(C <= 65535, WESUP_E_UNSUPPORTED, "wesup_sp_pool_bwd: CHW layout supports C <= 65535");
        dim3 grid(cdiv(HW, 256), C);
        if (dtype == WESUP_F32)
            pool_bwd_chw_kernel<float><<<grid, 256, 0, stream>>>(grad_pooled, row_labels, counts, HW, C, (float *)grad_feat);
        else
            pool_bwd_chw_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(grad_pooled, row_labels, counts, HW, C, (__nv_bfloat16 *)grad_feat);
    } else {
        WESUP_REQUIRE(false, WESUP_E_ARG, "wesup_sp_pool_bwd: bad layout %d", layout);
    }
    WESUP_CHECK_LAUNCH("wesup_sp_pool_bwd", 1);
    return 0;
}

extern "C" int wesup_sp_paint(const int32_t *row_labels, const float *sp_pred, int HW, int n_cls, int cls, float *out,
                              void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    WESUP_REQUIRE(row_labels && sp_pred && out, WESUP_E_ARG, "wesup_sp_paint: null pointer");
    WESUP_REQUIRE(HW > 0 && n_cls > 0 && cls >= 0 && cls < n_cls, WESUP_E_ARG, "wesup_sp_paint: bad size/class");
    paint_kernel<<<cdiv(HW, 256), 256, 0, stream>>>(row_labels, sp_pred, HW, n_cls, cls, out);
    WESUP_CHECK_LAUNCH("wesup_sp_paint", 1);
    return 0;
}
