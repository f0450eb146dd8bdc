// (c) Label propagation.  Replaces _label_propagate
// (/root/reference/models/wesup.py:99-139): the reference builds the full
// (N,N,D) difference tensor twice and an (N,N) affinity; only the
// unlabeled x labeled block is ever used.  Here each unlabeled row scans the
// labeled rows once; distance -> similarity -> running arg-max -> threshold ->
// label copy are fused, nothing of size n_u x n_l is written.
//
// Exact path (this file): direct-difference fp32, d2 = sum_k (f_uk - f_jk)^2,
// sim = expf(-d2), first arg-max on sim (ties -> lowest labeled index, exactly
// like torch.max on the reference's W_ul), strict `sim > thr`.
#include <stdlib.h>
#include "common.cuh"

namespace wesup {

constexpr int LP_THREADS = 128;
constexpr int LP_TILE = 64;     // labeled rows staged per shared-memory tile

template <int D>
__global__ void __launch_bounds__(LP_THREADS) label_propagate_exact_kernel(
    const float *__restrict__ feats, int N, int n_l, const float *__restrict__ y_l, int n_cls, float thr,
    float *__restrict__ y_u, int32_t *__restrict__ src_idx, float *__restrict__ max_sim) {
    __shared__ float tile[LP_TILE * D];
    const int n_u = N - n_l;
    const int u = blockIdx.x * LP_THREADS + threadIdx.x;
    const bool live = u < n_u;
    float f[D];
    if (live) {
        const float4 *row = reinterpret_cast<const float4 *>(feats + (long)(n_l + u) * D);
#pragma unroll
        for (int k = 0; k < D / 4; ++k) {
            float4 v = __ldg(row + k);
            f[4 * k] = v.x; f[4 * k + 1] = v.y; f[4 * k + 2] = v.z; f[4 * k + 3] = v.w;
        }
    }
    float best = -1.0f;     // similarities are in [0,1]
    int best_j = 0;
    for (int j0 = 0; j0 < n_l; j0 += LP_TILE) {
        int rows = min(LP_TILE, n_l - j0);
        __syncthreads();
        for (int i = threadIdx.x; i < rows * (D / 4); i += LP_THREADS)
            reinterpret_cast<float4 *>(tile)[i] = __ldg(reinterpret_cast<const float4 *>(feats + (long)j0 * D) + i);
        __syncthreads();
        if (live) {
            for (int j = 0; j < rows; ++j) {
                const float *lrow = tile + j * D;
                float d2 = 0.f;
#pragma unroll
                for (int k = 0; k < D; ++k) {
                    float t = f[k] - lrow[k];
                    d2 = fmaf(t, t, d2);
                }
                float sim = expf(-d2);
                if (sim > best) { best = sim; best_j = j0 + j; }
            }
        }
    }
    if (live) {
        bool take = best > thr;
        for (int c = 0; c < n_cls; ++c) y_u[(long)u * n_cls + c] = take ? __ldg(y_l + (long)best_j * n_cls + c) : 0.f;
        if (src_idx) src_idx[u] = best_j;
        if (max_sim) max_sim[u] = best;
    }
}

// generic feature width (D not in the compiled set): rows re-read from L1/L2
__global__ void __launch_bounds__(LP_THREADS) label_propagate_generic_kernel(
    const float *__restrict__ feats, int N, int D, int n_l, const float *__restrict__ y_l, int n_cls, float thr,
    float *__restrict__ y_u, int32_t *__restrict__ src_idx, float *__restrict__ max_sim) {
    const int n_u = N - n_l;
    const int u = blockIdx.x * LP_THREADS + threadIdx.x;
    if (u >= n_u) return;
    const float *fu = feats + (long)(n_l + u) * D;
    float best = -1.0f;
    int best_j = 0;
    for (int j = 0; j < n_l; ++j) {
        const float *fj = feats + (long)j * D;
        float d2 = 0.f;
        for (int k = 0; k < D; ++k) {
            float t = __ldg(fu + k) - __ldg(fj + k);
            d2 = fmaf(t, t, d2);
        }
        float sim = expf(-d2);
        if (sim > best) { best = sim; best_j = j; }
    }
    bool take = best > thr;
    for (int c = 0; c < n_cls; ++c) y_u[(long)u * n_cls + c] = take ? __ldg(y_l + (long)best_j * n_cls + c) : 0.f;
    if (src_idx) src_idx[u] = best_j;
    if (max_sim) max_sim[u] = best;
}

// Row counts read from DEVICE memory (counts_dev = {n_rows, n_labeled}): the launch shape
// depends only on the capacity n_max, so the call can sit in a CUDA graph that is replayed
// for images with different superpixel counts.  Writes ALL n_max rows of y_full: rows
// [n_labeled, n_rows) get the propagated label (or zeros below the threshold), every other
// row zeros -- exactly the rows _cross_entropy ignores (/root/reference/models/wesup.py:84-90).
template <int D>
__global__ void __launch_bounds__(LP_THREADS) label_propagate_dev_kernel(
    const float *__restrict__ feats, int n_max, int D_rt, const int32_t *__restrict__ counts_dev, const float *__restrict__ y_l,
    int n_cls, float thr, float *__restrict__ y_full) {
    __shared__ float tile[LP_TILE * (D > 0 ? D : 1)];
    const int n_rows = min(__ldg(counts_dev), n_max), n_l = min(__ldg(counts_dev + 1), n_rows);
    const int r = blockIdx.x * LP_THREADS + threadIdx.x;
    const bool in_range = r < n_max;
    const bool live = in_range && r >= n_l && r < n_rows;
    float best = -1.0f;
    int best_j = 0;
    if (D > 0) {
        float f[D > 0 ? D : 1];
        if (live) {
            const float4 *row = reinterpret_cast<const float4 *>(feats + (long)r * D);
#pragma unroll
            for (int k = 0; k < D / 4; ++k) {
                float4 v = __ldg(row + k);
                f[4 * k] = v.x; f[4 * k + 1] = v.y; f[4 * k + 2] = v.z; f[4 * k + 3] = v.w;
            }
        }
        // blocks that hold no unlabeled row skip the scan (block-uniform: derived from blockIdx and the counts)
        const int r0 = blockIdx.x * LP_THREADS;
        const bool block_live = r0 < n_rows && r0 + LP_THREADS > n_l;
        for (int j0 = 0; block_live && j0 < n_l; j0 += LP_TILE) {
            int rows = min(LP_TILE, n_l - j0);
            __syncthreads();
            for (int i = threadIdx.x; i < rows * (D / 4); i += LP_THREADS)
                reinterpret_cast<float4 *>(tile)[i] = __ldg(reinterpret_cast<const float4 *>(feats + (long)j0 * D) + i);
            __syncthreads();
            if (live) {
                for (int j = 0; j < rows; ++j) {
                    const float *lrow = tile + j * D;
                    float d2 = 0.f;
#pragma unroll
                    for (int k = 0; k < D; ++k) {
                        float t = f[k] - lrow[k];
                        d2 = fmaf(t, t, d2);
                    }
                    float sim = expf(-d2);
                    if (sim > best) { best = sim; best_j = j0 + j; }
                }
            }
        }
    } else if (live) {
        const float *fu = feats + (long)r * D_rt;
        for (int j = 0; j < n_l; ++j) {
            const float *fj = feats + (long)j * D_rt;
            float d2 = 0.f;
            for (int k = 0; k < D_rt; ++k) {
                float t = __ldg(fu + k) - __ldg(fj + k);
                d2 = fmaf(t, t, d2);
            }
            float sim = expf(-d2);
            if (sim > best) { best = sim; best_j = j; }
        }
    }
    if (in_range) {
        const bool take = live && n_l > 0 && best > thr;
        for (int c = 0; c < n_cls; ++c) y_full[(long)r * n_cls + c] = take ? __ldg(y_l + (long)best_j * n_cls + c) : 0.f;
    }
}

}  // namespace wesup

using namespace wesup;

extern "C" int wesup_label_propagate_dev(const float *feats, int n_max, int D, const int32_t *counts_dev, const float *y_l,
                                         int n_cls, float thr, float *y_full, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    WESUP_REQUIRE(feats && counts_dev && y_l && y_full, WESUP_E_ARG, "wesup_label_propagate_dev: null pointer");
    WESUP_REQUIRE(n_max > 0 && D > 0 && n_cls > 0, WESUP_E_ARG, "wesup_label_propagate_dev: bad size n_max=%d D=%d n_cls=%d", n_max, D, n_cls);
    const int grid = cdiv(n_max, LP_THREADS);
    if (D == 32 && aligned16(feats))
        label_propagate_dev_kernel<32><<<grid, LP_THREADS, 0, stream>>>(feats, n_max, D, counts_dev, y_l, n_cls, thr, y_full);
    else
        label_propagate_dev_kernel<0><<<grid, LP_THREADS, 0, stream>>>(feats, n_max, D, counts_dev, y_l, n_cls, thr, y_full);
    WESUP_CHECK_LAUNCH("wesup_label_propagate_dev", 1);
    return 0;
}

extern "C" size_t wesup_label_propagate_tc_workspace_bytes(int N, int D, int n_l);
extern "C" int wesup_label_propagate_tc(const float *feats, int N, int D, int n_l, const float *y_l, int n_cls, float thr,
                                        float *y_u, int32_t *src_idx, float *max_sim, void *ws, void *stream_);

// Labeled rows from which the tensor-core path is taken.  Measured on B200 (profiles/r2d_bench_lp.txt): the CUDA-core
// kernel costs ~10 us + 0.14 us per labeled row (every unlabeled row walks all of them), the tcgen05 pipeline ~15 us
// flat up to a few thousand rows each way -- the crossover is near 40 labeled rows.
static long tc_min_labeled() {
    static long v = -1;
    if (v < 0) {
        const char *e = getenv("WESUP_LP_TC_MIN_LABELED");
        v = e ? atol(e) : 64;
        if (v < 1) v = 1;
    }
    return v;
}

extern "C" size_t wesup_label_propagate_workspace_bytes(int N, int D, int n_l) {
    size_t tc = D == 32 ? wesup_label_propagate_tc_workspace_bytes(N, D, n_l) : 0;
    return tc > 256 ? tc : 256;   // the exact path needs none; non-zero so callers always pass a valid pointer
}

extern "C" int wesup_label_propagate_exact(const float *feats, int N, int D, int n_l, const float *y_l, int n_cls, float thr,
                                           float *y_u, int32_t *src_idx, float *max_sim, void *ws, void *stream_);

extern "C" int wesup_label_propagate(const float *feats, int N, int D, int n_l, const float *y_l, int n_cls, float thr,
                                     float *y_u, int32_t *src_idx, float *max_sim, void *ws, void *stream_) {
    if (feats && ws && D == 32 && n_l > 0 && n_l < N && aligned16(feats) && aligned16(ws) &&
        n_l >= tc_min_labeled())
        return wesup_label_propagate_tc(feats, N, D, n_l, y_l, n_cls, thr, y_u, src_idx, max_sim, ws, stream_);
    return wesup_label_propagate_exact(feats, N, D, n_l, y_l, n_cls, thr, y_u, src_idx, max_sim, ws, stream_);
}

extern "C" int wesup_label_propagate_exact(const float *feats, int N, int D, int n_l, const float *y_l, int n_cls, float thr,
                                           float *y_u, int32_t *src_idx, float *max_sim, void *ws, void *stream_) {
    (void)ws;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    WESUP_REQUIRE(feats && y_l && y_u, WESUP_E_ARG, "wesup_label_propagate: null pointer");
    WESUP_REQUIRE(N > 0 && D > 0 && n_cls > 0, WESUP_E_ARG, "wesup_label_propagate: bad size N=%d D=%d n_cls=%d", N, D, n_cls);
    WESUP_REQUIRE(n_l > 0 && n_l <= N, WESUP_E_ARG, "wesup_label_propagate: n_l=%d must be in [1,N=%d]", n_l, N);
    const int n_u = N - n_l;
    if (n_u == 0) return 0;
    int grid = cdiv(n_u, LP_THREADS);
    if (D == 32 && aligned16(feats))
        label_propagate_exact_kernel<32><<<grid, LP_THREADS, 0, stream>>>(feats, N, n_l, y_l, n_cls, thr, y_u, src_idx, max_sim);
    else if (D == 64 && aligned16(feats))
        label_propagate_exact_kernel<64><<<grid, LP_THREADS, 0, stream>>>(feats, N, n_l, y_l, n_cls, thr, y_u, src_idx, max_sim);
    else
        label_propagate_generic_kernel<<<grid, LP_THREADS, 0, stream>>>(feats, N, D, n_l, y_l, n_cls, thr, y_u, src_idx, max_sim);
    WESUP_CHECK_LAUNCH("wesup_label_propagate", 1);
    return 0;
}
