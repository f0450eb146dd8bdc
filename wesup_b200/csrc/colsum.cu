// Column sums of a pixel-major (rows, C) fp32 matrix: the bias gradient of a channels_last convolution,
// grad_bias[c] = sum_p grad_out[p, c] -- what autograd evaluates as grad_out.sum((0, 2, 3)) for every nn.Conv2d of the
// reference (/root/reference/models/wesup.py:190-210: the 13 backbone convolutions and their 1x1 side convolutions).
// ATen's generic reduction reads the 232 MB of conv gradients of a 464^2 image at ~0.6 TB/s (375 us per image,
// 9 % of a training step); this is a plain two-stage streaming reduction at HBM speed.  Deterministic: fixed slab
// boundaries, fixed summation order in both stages, no floating-point atomics.
#include "common.cuh"

namespace wesup {

constexpr int CS_THREADS = 256;

// stage 1: block b sums rows [b * rpb, (b + 1) * rpb); thread = (row lane, 4-channel column), four loads in flight
__global__ void __launch_bounds__(CS_THREADS) colsum_partial_kernel(const float *__restrict__ x, long rows, int C4, int rpb,
                                                                    float *__restrict__ partial) {
    __shared__ float4 red[CS_THREADS];
    const int tid = threadIdx.x;
    const int nrl = CS_THREADS / C4;                        // row lanes (host: C4 <= 256)
    const int rl = tid / C4, c4 = tid - rl * C4;
    const long r0 = (long)blockIdx.x * rpb, r1 = min(r0 + (long)rpb, rows);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (rl < nrl) {
        const float4 *__restrict__ p = reinterpret_cast<const float4 *>(x) + c4;
        long r = r0 + rl;
        for (; r + 3L * nrl < r1; r += 4L * nrl) {
            const float4 a = ldg_stream(p + r * C4), b = ldg_stream(p + (r + nrl) * C4);
            const float4 c = ldg_stream(p + (r + 2L * nrl) * C4), d = ldg_stream(p + (r + 3L * nrl) * C4);
            acc = acc + a; acc = acc + b; acc = acc + c; acc = acc + d;
        }
        for (; r < r1; r += nrl) acc = acc + ldg_stream(p + r * C4);
    }
    red[tid] = acc;
    __syncthreads();
    if (rl == 0) {
        for (int s_ = 1; s_ < nrl; ++s_) acc = acc + red[s_ * C4 + c4];
        reinterpret_cast<float4 *>(partial)[(long)blockIdx.x * C4 + c4] = acc;
    }
}

// stage 2: block = 32 channels x 32 row lanes; lane j adds the block partials j, j + 32, ... (four independent chains in
// flight), the 32 lane sums meet in shared memory in lane order.  Fixed order => deterministic.  (r2 timeline: the
// first version, one thread per channel walking all ~600 partials in one dependent chain, took 12 us per call -- 13
// calls per training iteration, three times the streaming stage it finishes.)
constexpr int CF_LANES = 32;
__global__ void __launch_bounds__(32 * CF_LANES) colsum_final_kernel(const float *__restrict__ partial, int nblk, int C, float *__restrict__ out) {
    __shared__ float red[CF_LANES][32];
    const int cl = threadIdx.x & 31, j = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cl;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    if (c < C) {
        const float *__restrict__ p = partial + c;
        int b = j;
        for (; b + 3 * CF_LANES < nblk; b += 4 * CF_LANES) {
            a0 += p[(long)b * C]; a1 += p[(long)(b + CF_LANES) * C];
            a2 += p[(long)(b + 2 * CF_LANES) * C]; a3 += p[(long)(b + 3 * CF_LANES) * C];
        }
        for (; b < nblk; b += CF_LANES) a0 += p[(long)b * C];
    }
    red[j][cl] = (a0 + a1) + (a2 + a3);
    __syncthreads();
    if (j == 0 && c < C) {
        float acc = red[0][cl];
#pragma unroll
        for (int k = 1; k < CF_LANES; ++k) acc += red[k][cl];
        out[c] = acc;
    }
}

static inline int colsum_blocks(long rows, int C4, int *rpb) {
    const int nrl = CS_THREADS / C4;
    long per = (rows + 4L * kNumSMs - 1) / (4L * kNumSMs);  // ~4 blocks per SM
    const long quantum = 4L * nrl;
    per = (per + quantum - 1) / quantum * quantum;
    if (per < quantum) per = quantum;
    *rpb = (int)per;
    return (int)((rows + per - 1) / per);
}

}  // namespace wesup

using namespace wesup;

extern "C" size_t wesup_colsum_workspace_bytes(long rows, int C) {
    if (rows <= 0 || C <= 0 || C % 4 != 0 || C / 4 > CS_THREADS) return 0;
    int rpb;
    const int nblk = colsum_blocks(rows, C / 4, &rpb);
    return (size_t)nblk * (size_t)C * sizeof(float);
}

extern "C" int wesup_colsum(const float *x, long rows, int C, float *out, void *ws, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    WESUP_REQUIRE(x && out && ws, WESUP_E_ARG, "wesup_colsum: null pointer");
    WESUP_REQUIRE(rows > 0 && C > 0, WESUP_E_ARG, "wesup_colsum: bad size rows=%ld C=%d", rows, C);
    WESUP_REQUIRE(C % 4 == 0 && C / 4 <= CS_THREADS, WESUP_E_UNSUPPORTED, "wesup_colsum: C=%d must be a multiple of 4, at most %d", C, 4 * CS_THREADS);
    WESUP_REQUIRE(aligned16(x) && aligned16(ws), WESUP_E_ALIGN, "wesup_colsum: x and ws must be 16-byte aligned");
    int rpb;
    const int nblk = colsum_blocks(rows, C / 4, &rpb);
    colsum_partial_kernel<<<nblk, CS_THREADS, 0, stream>>>(x, rows, C / 4, rpb, static_cast<float *>(ws));
    colsum_final_kernel<<<cdiv(C, 32), 32 * CF_LANES, 0, stream>>>(static_cast<const float *>(ws), nblk, C, out);
    WESUP_CHECK_LAUNCH("wesup_colsum", 2);
    return 0;
}
