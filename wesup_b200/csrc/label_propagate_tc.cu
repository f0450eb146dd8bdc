// (c) Label propagation on the 5th-generation tensor cores (tcgen05, TMEM
// accumulators, fp32 accumulate), fused with the distance -> similarity ->
// arg-max -> threshold -> label-copy epilogue.  Replaces _label_propagate
// (/root/reference/models/wesup.py:99-139); only the unlabeled x labeled block
// of the affinity is ever formed, tile by tile, in tensor memory.
//
// Bit-exactness.  The reference computes d2 = sum_k (f_uk - f_jk)^2 directly in
// fp32; the GEMM form |a|^2 + |b|^2 - 2 a.b cancels exactly where the 0.8
// threshold lives, so the tensor-core result is used only as a FILTER:
//   * operands are split a = a_hi + a_lo into two TF32-representable parts and
//     the MMA runs over K' = 96 = [a_hi|a_hi|a_lo] . [b_hi|b_lo|b_hi], i.e.
//     a_hi.b_hi + a_hi.b_lo + a_lo.b_hi with fp32 accumulation in TMEM: the
//     dropped a_lo.b_lo term and the accumulation rounding bound the error of
//     the approximate d2 by KAPPA * (|a|^2 + |b|^2);
//   * every labeled row j whose approximate d2 is within the two error bounds (its
//     own and that of the column holding the minimum, + a slack for exp()
//     collapsing close distances to the same fp32 similarity) of the minimum seen
//     so far -- INCLUDING the current tile's own minimum, found in a first pass over
//     the accumulator -- is re-evaluated EXACTLY with the reference's arithmetic
//     (direct differences, same operation order as label_propagate.cu); the winner
//     is the packed maximum (similarity, then lowest index) = the first-arg-max rule.
// The true arg-max always survives the filter (its approximate d2 cannot exceed
// the approximate minimum by more than the two error bounds), so src / sim /
// y_u are bit-identical to the exact kernel; tests assert exactly that and
// check the measured error against KAPPA.
//
// Epilogue without divergence: a lane owns one row (its TMEM lane), but the columns
// that survive the filter differ from row to row, so evaluating them in place would
// make a warp pay for the union of its 32 rows' candidates.  Instead the surviving
// (row, column) pairs are compacted with ballots into a per-warp queue and evaluated
// LANE-PARALLEL, one pair per lane: the row's features come from the owning lane's
// registers through shuffles, the labeled row from global memory (L1/L2), the result
// goes to the row's packed maximum with a shared-memory atomicMax.
//
// Shape: one CTA = 128 unlabeled rows (UMMA M = 128, cta_group::1) x a chunk
// of the labeled rows walked in tiles of 128 (UMMA N = 128, K = 8 per
// instruction, 12 instructions per tile).  The labeled dimension is split
// across gridDim.y so small n_u still fills the 148 SMs; partial arg-maxes
// meet in a packed 64-bit atomicMax (similarity bits high, inverted index low
// => highest similarity, then lowest index) and the last CTA of a row block
// writes the outputs.
#include "common.cuh"

namespace wesup {

constexpr int TC_M = 128;            // unlabeled rows per CTA  (UMMA M)
constexpr int TC_N = 128;            // labeled rows per tile   (UMMA N)
constexpr int TC_D = 32;             // feature width
constexpr int TC_K = 3 * TC_D;       // split-TF32 contraction length
constexpr int TC_CHUNKS = TC_K / 4;  // 16-byte K chunks per row (24)
constexpr int TC_LBO = 128;                  // bytes between K-adjacent core matrices
constexpr int TC_SBO = TC_CHUNKS * 128;      // bytes between 8-row groups (3072)
constexpr int TC_THREADS = 256;
constexpr float TC_KAPPA = 1.52587890625e-05f;   // 2^-16, see header; measured error is ~30x smaller
constexpr float TC_TIE_SLACK = 2.0e-6f;

constexpr int TC_QCAP = 128;         // candidate pairs queued per warp and tile (the rest are evaluated in place)

struct TcShared {
    alignas(128) uint32_t a[TC_M * TC_K];       // 48 KB, canonical K-major no-swizzle core-matrix layout
    alignas(128) uint32_t b[TC_N * TC_K];       // 48 KB
    float nb[TC_N];                             // |b_j|^2
    alignas(8) unsigned long long row_key[TC_M];// best (similarity, lowest index) per row of this CTA, packed
    alignas(8) uint2 queue[TC_THREADS / 32][TC_QCAP];   // per warp: {row lane << 8 | column, approximate d2 bits}
    int qn[TC_THREADS / 32];                    // queue lengths
    alignas(8) unsigned long long mbar;
    uint32_t tmem_base;
    int last_flag;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint32_t to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}

// 64-bit shared-memory matrix descriptor, SWIZZLE_NONE, K-major (cute::UMMA::SmemDescriptor)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);              // start address   [0,14)
    d |= (uint64_t)(TC_LBO >> 4) << 16;                   // leading byte offset [16,30)
    d |= (uint64_t)(TC_SBO >> 4) << 32;                   // stride byte offset  [32,46)
    d |= (uint64_t)1 << 46;                               // descriptor version (Blackwell)
    return d;                                             // base_offset 0, lbo_mode 0, layout_type 0
}

// kind::tf32, fp32 accumulate, A and B K-major, M = 128, N = TC_N (cute::UMMA::InstrDescriptor)
constexpr uint32_t TC_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(TC_IDESC), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}\n" ::"r"(mbar), "r"(parity)
        : "memory");
}

// row r, 16-byte chunk q of a K-major no-swizzle operand tile (in uint32 units)
__device__ __forceinline__ int tile_word(int r, int q) { return ((r >> 3) * TC_SBO + q * TC_LBO + (r & 7) * 16) >> 2; }

// Split one fp32 row (32 values) into the K' = 96 operand row.  which = 0: A side
// [hi|hi|lo]; which = 1: B side [hi|lo|hi].
__device__ __forceinline__ void store_split_row(uint32_t *tile, int r, const float4 *v4, int which) {
#pragma unroll
    for (int q = 0; q < TC_D / 4; ++q) {
        float4 v = v4[q];
        uint4 hi = make_uint4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
        uint4 lo = make_uint4(to_tf32(v.x - __uint_as_float(hi.x)), to_tf32(v.y - __uint_as_float(hi.y)),
                              to_tf32(v.z - __uint_as_float(hi.z)), to_tf32(v.w - __uint_as_float(hi.w)));
        *reinterpret_cast<uint4 *>(tile + tile_word(r, q)) = hi;
        *reinterpret_cast<uint4 *>(tile + tile_word(r, q + 8)) = which ? lo : hi;
        *reinterpret_cast<uint4 *>(tile + tile_word(r, q + 16)) = which ? hi : lo;
    }
}

struct TcStats {                       // optional diagnostics in the workspace
    unsigned long long exact_evals;
    unsigned int max_err_ratio_bits;   // max |approx d2 - exact d2| / (|a|^2 + |b|^2), float bits
    unsigned int pad;
};

__device__ __forceinline__ unsigned long long pack_key(float sim, int j) {
    return ((unsigned long long)__float_as_uint(sim) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)j);
}

#define TC_TMEM_LD32(d, addr)                                                                                                   \
    asm volatile(                                                                                                              \
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                              \
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];" \
        : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6]), "=r"(d[7]), "=r"(d[8]), "=r"(d[9]),       \
          "=r"(d[10]), "=r"(d[11]), "=r"(d[12]), "=r"(d[13]), "=r"(d[14]), "=r"(d[15]), "=r"(d[16]), "=r"(d[17]), "=r"(d[18]),          \
          "=r"(d[19]), "=r"(d[20]), "=r"(d[21]), "=r"(d[22]), "=r"(d[23]), "=r"(d[24]), "=r"(d[25]), "=r"(d[26]), "=r"(d[27]),          \
          "=r"(d[28]), "=r"(d[29]), "=r"(d[30]), "=r"(d[31])                                                                         \
        : "r"(addr));                                                                                                          \
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory")

// 256 threads: warps 0-3 stage the operands (thread = row) and thread 0 issues the MMAs; in the epilogue ALL eight warps
// work -- warp w reads TMEM lanes 32*(w & 3) .. +31 (the lanes a warp may access) and the columns 64*(w >> 2) .. +63,
// so two warps share a row block and split its columns.
__global__ void __launch_bounds__(TC_THREADS, 2) label_propagate_tc_kernel(
    const float *__restrict__ feats, int N, int n_l, int rows_per_split, const float *__restrict__ y_l, int n_cls,
    float thr, float *__restrict__ y_u, int32_t *__restrict__ src_idx, float *__restrict__ max_sim,
    unsigned long long *__restrict__ keys, unsigned int *__restrict__ tickets, TcStats *__restrict__ stats) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    TcShared &S = *reinterpret_cast<TcShared *>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int rwarp = warp & 3, half = warp >> 2;
    const int row = rwarp * 32 + lane;                                  // row of the CTA = TMEM lane
    const int n_u = N - n_l;
    const int u = blockIdx.x * TC_M + row;
    const bool live = u < n_u;
    const int j_begin = blockIdx.y * rows_per_split;
    const int j_end = min(n_l, j_begin + rows_per_split);

    // ---- one-time setup: TMEM allocation, mbarrier, A operand ----------------
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&S.tmem_base)), "n"(TC_N));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&S.mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < TC_M) S.row_key[tid] = 0ull;
    if (lane == 0) S.qn[warp] = 0;
    float f[TC_D];
    float na = 0.f;
    {
        float4 rowv[TC_D / 4];
        const float4 *src = reinterpret_cast<const float4 *>(feats + (long)(n_l + (live ? u : 0)) * TC_D);
#pragma unroll
        for (int q = 0; q < TC_D / 4; ++q) {
            rowv[q] = live ? __ldg(src + q) : make_float4(0.f, 0.f, 0.f, 0.f);
            f[4 * q] = rowv[q].x; f[4 * q + 1] = rowv[q].y; f[4 * q + 2] = rowv[q].z; f[4 * q + 3] = rowv[q].w;
        }
#pragma unroll
        for (int k = 0; k < TC_D; ++k) na = fmaf(f[k], f[k], na);
        if (half == 0) store_split_row(S.a, row, rowv, 0);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = S.tmem_base;
    const uint32_t tmem_row = tmem + ((uint32_t)(rwarp * 32) << 16);    // this warp's 32 TMEM lanes
    const uint64_t adesc0 = make_smem_desc(smem_u32(S.a));
    const uint64_t bdesc0 = make_smem_desc(smem_u32(S.b));
    const uint32_t mbar = smem_u32(&S.mbar);

    float run_min = 3.0e38f, nb_max = 0.f;
    unsigned int n_exact = 0;
    float max_ratio = 0.f;
    uint32_t phase = 0;
    uint2 *queue = S.queue[warp];

    for (int j0 = j_begin; j0 < j_end; j0 += TC_N) {
        const int rows = min(TC_N, j_end - j0);
        if (j0 != j_begin) __syncthreads();        // every warp is done with the previous tile's accumulator and nb
        // ---- stage the labeled tile: split operand and norms (thread = labeled row) ----------
        if (tid < TC_N) {
            float4 rowv[TC_D / 4];
            const bool have = tid < rows;
            const float4 *src = reinterpret_cast<const float4 *>(feats + (long)(j0 + (have ? tid : 0)) * TC_D);
            float nrm = 0.f;
#pragma unroll
            for (int q = 0; q < TC_D / 4; ++q) {
                rowv[q] = have ? __ldg(src + q) : make_float4(0.f, 0.f, 0.f, 0.f);
                nrm = fmaf(rowv[q].x, rowv[q].x, nrm); nrm = fmaf(rowv[q].y, rowv[q].y, nrm);
                nrm = fmaf(rowv[q].z, rowv[q].z, nrm); nrm = fmaf(rowv[q].w, rowv[q].w, nrm);
            }
            S.nb[tid] = nrm;
            store_split_row(S.b, tid, rowv, 1);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> visible to the tensor core
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); // orders the previous tile's tcgen05.ld before the barrier
        __syncthreads();
        // ---- D[128 x 128] = A' . B'^T : 12 x (M128,N128,K8) kind::tf32 ---------
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int k = 0; k < TC_K / 8; ++k) {
                const uint64_t step = (uint64_t)((k * 2 * TC_LBO) >> 4);  // two 16-byte K chunks per instruction
                mma_tf32(tmem, adesc0 + step, bdesc0 + step, k > 0 ? 1u : 0u);
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory");
        }
        mbar_wait(mbar, phase);
        phase ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // ---- epilogue: approx d2 -> minimum of my columns -> filter -> queue -> lane-parallel exact evaluation ----
        const int c_lo = half * (TC_N / 2), c_hi = min(rows, c_lo + TC_N / 2);
        // pass 1: this row's minimum approximate distance over my columns of the tile, and the largest labeled norm
        float tmin = 3.0e38f;
#pragma unroll 1
        for (int cb = c_lo; cb < c_hi; cb += 32) {
            uint32_t d[32];
            TC_TMEM_LD32(d, tmem_row + (uint32_t)cb);
#pragma unroll
            for (int t = 0; t < 32; ++t) {
                if (cb + t < c_hi) {
                    const float nbj = S.nb[cb + t];
                    tmin = fminf(tmin, fmaf(-2.0f, __uint_as_float(d[t]), na + nbj));
                    nb_max = fmaxf(nb_max, nbj);
                }
            }
        }
        const float gate = fminf(run_min, tmin);
        const float e_gate = TC_KAPPA * (na + nb_max);                    // error bound of whichever column holds the minimum
        // pass 2: queue the columns within the error bounds of the minimum (rare: a shared-memory atomic per hit)
#pragma unroll 1
        for (int cb = c_lo; cb < c_hi; cb += 32) {
            uint32_t d[32];
            TC_TMEM_LD32(d, tmem_row + (uint32_t)cb);
#pragma unroll
            for (int t = 0; t < 32; ++t) {
                const int jj = cb + t;
                if (jj < c_hi) {
                    const float sj = na + S.nb[jj];
                    const float approx = fmaf(-2.0f, __uint_as_float(d[t]), sj);
                    if (live && approx <= gate + (fmaf(TC_KAPPA, sj, e_gate) + TC_TIE_SLACK)) {
                        const int pos = atomicAdd(&S.qn[warp], 1);
                        if (pos < TC_QCAP) {
                            queue[pos] = make_uint2(((unsigned)lane << 8) | (unsigned)jj, __float_as_uint(approx));
                        } else {
                            // queue full (collapsed features: every column is a candidate): evaluate in place
                            const float *lrow = feats + (long)(j0 + jj) * TC_D;
                            float d2 = 0.f;
#pragma unroll
                            for (int k = 0; k < TC_D; ++k) {
                                const float df = f[k] - __ldg(lrow + k);
                                d2 = fmaf(df, df, d2);
                            }
                            atomicMax(&S.row_key[row], pack_key(expf(-d2), j0 + jj));
                            ++n_exact;
                            if (stats != nullptr && sj > 0.f) max_ratio = fmaxf(max_ratio, fabsf(approx - d2) / sj);
                        }
                    }
                }
            }
        }
        __syncwarp();
        {
            const int qn = min(S.qn[warp], TC_QCAP);
            for (int i0 = 0; i0 < qn; i0 += 32) {
                const int i = i0 + lane;
                const bool have = i < qn;
                const uint2 e = have ? queue[i] : make_uint2(0u, 0u);
                const int r = (int)(e.x >> 8), jj = (int)(e.x & 0xffu);
                const float4 *brow = reinterpret_cast<const float4 *>(feats + (long)(j0 + jj) * TC_D);
                float d2 = 0.f;
#pragma unroll
                for (int q = 0; q < TC_D / 4; ++q) {
                    const float4 bv = __ldg(brow + q);
                    float df;
                    df = __shfl_sync(0xffffffffu, f[4 * q], r) - bv.x;     d2 = fmaf(df, df, d2);
                    df = __shfl_sync(0xffffffffu, f[4 * q + 1], r) - bv.y; d2 = fmaf(df, df, d2);
                    df = __shfl_sync(0xffffffffu, f[4 * q + 2], r) - bv.z; d2 = fmaf(df, df, d2);
                    df = __shfl_sync(0xffffffffu, f[4 * q + 3], r) - bv.w; d2 = fmaf(df, df, d2);
                }
                const float na_r = __shfl_sync(0xffffffffu, na, r);
                if (have) {
                    atomicMax(&S.row_key[rwarp * 32 + r], pack_key(expf(-d2), j0 + jj));
                    ++n_exact;
                    const float sr = na_r + S.nb[jj];
                    if (stats != nullptr && sr > 0.f) max_ratio = fmaxf(max_ratio, fabsf(__uint_as_float(e.y) - d2) / sr);
                }
            }
            __syncwarp();
            if (lane == 0) S.qn[warp] = 0;
            __syncwarp();
        }
        run_min = gate;
    }
    if (stats != nullptr) {
        for (int o = 16; o > 0; o >>= 1) {
            n_exact += __shfl_xor_sync(0xffffffffu, n_exact, o);
            max_ratio = fmaxf(max_ratio, __shfl_xor_sync(0xffffffffu, max_ratio, o));
        }
        if (lane == 0) {
            atomicAdd(&stats->exact_evals, (unsigned long long)n_exact);
            atomicMax(&stats->max_err_ratio_bits, __float_as_uint(max_ratio));
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();                                                     // both column halves have reported to row_key
    // ---- merge the labeled-dimension splits ------------------------------------
    if (half == 0 && live && j_end > j_begin) atomicMax(keys + u, S.row_key[row]);
    __threadfence();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TC_N));
    }
    __syncthreads();
    if (tid == 0) S.last_flag = (atomicAdd(tickets + blockIdx.x, 1u) == gridDim.y - 1) ? 1 : 0;
    __syncthreads();
    if (S.last_flag && half == 0 && live) {
        __threadfence();
        const unsigned long long key = *reinterpret_cast<volatile unsigned long long *>(keys + u);
        const float sim = __uint_as_float((unsigned)(key >> 32));
        const int j = (int)(0xFFFFFFFFu - (unsigned)(key & 0xFFFFFFFFull));
        const bool take = sim > thr;
        for (int c = 0; c < n_cls; ++c) y_u[(long)u * n_cls + c] = take ? __ldg(y_l + (long)j * n_cls + c) : 0.f;
        if (src_idx) src_idx[u] = j;
        if (max_sim) max_sim[u] = sim;
    }
}

static inline size_t up256(size_t x) { return (x + 255) / 256 * 256; }

}  // namespace wesup

using namespace wesup;

// workspace: [keys n_u x u64][tickets ceil(n_u/128) x u32][TcStats]
extern "C" size_t wesup_label_propagate_tc_workspace_bytes(int N, int D, int n_l) {
    (void)D;
    long n_u = (long)N - n_l;
    if (n_u <= 0) return 256;
    return up256(sizeof(unsigned long long) * n_u) + up256(sizeof(unsigned int) * ((n_u + TC_M - 1) / TC_M)) + up256(sizeof(TcStats));
}

extern "C" int wesup_label_propagate_tc(const float *feats, int N, int D, int n_l, const float *y_l, int n_cls, float thr,
                                        float *y_u, int32_t *src_idx, float *max_sim, void *ws, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    WESUP_REQUIRE(feats && y_l && y_u && ws, WESUP_E_ARG, "wesup_label_propagate_tc: null pointer");
    WESUP_REQUIRE(N > 0 && n_cls > 0, WESUP_E_ARG, "wesup_label_propagate_tc: bad size N=%d n_cls=%d", N, n_cls);
    WESUP_REQUIRE(D == TC_D, WESUP_E_UNSUPPORTED, "wesup_label_propagate_tc: the tensor-core path is built for D=%d (got %d)", TC_D, D);
    WESUP_REQUIRE(n_l > 0 && n_l <= N, WESUP_E_ARG, "wesup_label_propagate_tc: n_l=%d must be in [1,N=%d]", n_l, N);
    WESUP_REQUIRE(aligned16(feats) && aligned16(ws), WESUP_E_ALIGN, "wesup_label_propagate_tc: feats/ws must be 16-byte aligned");
    const int n_u = N - n_l;
    if (n_u == 0) return 0;
    const int row_blocks = (n_u + TC_M - 1) / TC_M;
    const int tiles = (n_l + TC_N - 1) / TC_N;
    int splits = (2 * kNumSMs + row_blocks - 1) / row_blocks;          // aim for >= 2 CTAs per SM
    if (splits > tiles) splits = tiles;
    if (splits < 1) splits = 1;
    const int tiles_per_split = (tiles + splits - 1) / splits;
    splits = (tiles + tiles_per_split - 1) / tiles_per_split;          // no empty splits
    char *base = static_cast<char *>(ws);
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(base);
    unsigned int *tickets = reinterpret_cast<unsigned int *>(base + up256(sizeof(unsigned long long) * n_u));
    TcStats *stats = reinterpret_cast<TcStats *>(base + up256(sizeof(unsigned long long) * n_u) + up256(sizeof(unsigned int) * row_blocks));
    size_t total = wesup_label_propagate_tc_workspace_bytes(N, D, n_l);
    cudaError_t e = cudaMemsetAsync(ws, 0, total, stream);
    WESUP_REQUIRE(e == cudaSuccess, (int)e, "wesup_label_propagate_tc: memset: %s", cudaGetErrorString(e));
    static bool configured = false;
    if (!configured) {
        e = cudaFuncSetAttribute(label_propagate_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TcShared));
        WESUP_REQUIRE(e == cudaSuccess, (int)e, "wesup_label_propagate_tc: smem attribute: %s", cudaGetErrorString(e));
        configured = true;
    }
    label_propagate_tc_kernel<<<dim3(row_blocks, splits), TC_THREADS, sizeof(TcShared), stream>>>(
        feats, N, n_l, tiles_per_split * TC_N, y_l, n_cls, thr, y_u, src_idx, max_sim, keys, tickets, stats);
    WESUP_CHECK_LAUNCH("wesup_label_propagate_tc", 1);
    return 0;
}

// diagnostics of the most recent call that used `ws` (read after synchronising the stream):
// out[0] = exact re-evaluations, out[1] = max |approx - exact| / (|a|^2+|b|^2) as float bits
extern "C" int wesup_label_propagate_tc_stats(const void *ws, int N, int n_l, unsigned long long *out_host) {
    WESUP_REQUIRE(ws && out_host, WESUP_E_ARG, "wesup_label_propagate_tc_stats: null pointer");
    const int n_u = N - n_l;
    WESUP_REQUIRE(n_u > 0, WESUP_E_ARG, "wesup_label_propagate_tc_stats: no unlabeled rows");
    const int row_blocks = (n_u + TC_M - 1) / TC_M;
    const char *base = static_cast<const char *>(ws);
    TcStats s;
    cudaError_t e = cudaMemcpy(&s, base + up256(sizeof(unsigned long long) * n_u) + up256(sizeof(unsigned int) * row_blocks),
                               sizeof(TcStats), cudaMemcpyDeviceToHost);
    WESUP_REQUIRE(e == cudaSuccess, (int)e, "wesup_label_propagate_tc_stats: %s", cudaGetErrorString(e));
    out_host[0] = s.exact_evals;
    out_host[1] = s.max_err_ratio_bits;
    return 0;
}
