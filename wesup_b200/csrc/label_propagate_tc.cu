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
//     the MMA runs over [a_hi|a_hi|a_lo|1 1] . [b_hi|b_lo|b_hi|n_hi n_lo], i.e.
//     a_hi.b_hi + a_hi.b_lo + a_lo.b_hi - |b|^2/2 (the labeled norm rides in the
//     contraction as two more TF32 parts) with fp32 accumulation in TMEM, so that
//     the approximate distance of row u to column j is |a_u|^2 - 2 acc_uj and the
//     nearest column of a tile is simply the LARGEST accumulator of the row; the
//     dropped a_lo.b_lo term and the roundings bound the error of the approximate
//     d2 by KAPPA * (|a|^2 + |b|^2);
//   * every labeled row j whose approximate d2 is within the error bounds of the
//     minimum seen so far -- INCLUDING the current tile's own minimum, found in a
//     first pass over the accumulator -- (+ a slack for exp() collapsing close
//     distances to the same fp32 similarity) is re-evaluated EXACTLY with the
//     reference's arithmetic (direct differences, same operation order as
//     label_propagate.cu); the winner is the packed maximum (similarity, then
//     lowest index) = the first-arg-max rule.
// The true arg-max always survives the filter, so src / sim / y_u are
// bit-identical to the exact kernel; tests assert exactly that and check the
// measured error against KAPPA.
//
// Pipeline (one CTA per SM, 192 threads, warp-specialised):
//   prep kernel    splits every feature row ONCE per call into the K-major core-matrix layout the MMA reads
//                  (a 128-row tile = one contiguous 36 KB block, [hi|lo|norm]) and records the largest labeled norm per tile;
//   warp 0         producer: one lane streams the CTA's A tile and then its B tiles into a 3-stage ring with
//                  cp.async.bulk (1-D TMA copies), completion on mbarriers (expect_tx);
//   warp 1         one lane issues 13 x tcgen05.mma (M128 N128 K8, kind::tf32) per tile into one of FOUR 128-column
//                  TMEM accumulators and commits to two mbarriers: "stage free" and "accumulator full";
//   warps 2-9      epilogue, thread = row = TMEM lane, two warps per lane quarter splitting a tile's columns: pass 1 reads the 128 columns (tcgen05.ld 32x32b.x32) and takes
//                  the row maximum -- one FMNMX per element; pass 2 re-reads only the 32-column groups whose maximum
//                  reaches the candidate threshold, pushes (row, column) pairs into a per-warp queue, and the queue is
//                  evaluated LANE-PARALLEL, one exact pair per lane (a lane owns a row, but the surviving columns
//                  differ from row to row: evaluated in place a warp would pay for the union of its rows' candidates);
//                  then the accumulator is handed back to the MMA warp ("accumulator empty", 128 arrivals).
// MMA of tile t+1..t+3 and the loads behind them overlap the epilogue of tile t.  The labeled dimension is split across
// gridDim.y so small n_u still fills the 148 SMs; partial arg-maxes meet in a packed 64-bit atomicMax (similarity
// bits high, inverted index low => highest similarity, then lowest index) and the last CTA of a row block writes
// the outputs.
#include "common.cuh"

namespace wesup {

constexpr int TC_M = 128;            // unlabeled rows per CTA  (UMMA M)
constexpr int TC_N = 128;            // labeled rows per tile   (UMMA N)
constexpr int TC_D = 32;             // feature width
constexpr int TC_CHUNKS = 18;        // 16-byte K chunks per operand row in memory: 8 hi + 8 lo, 1 norm chunk, 1 zero chunk
constexpr int TC_KSTEPS = 13;        // MMA instructions per tile (K = 8 each): hi.hi, hi.lo, lo.hi (4 each) + the norm step --
                                     // the hi halves are READ twice through their descriptors, not stored twice
constexpr int TC_LBO = 128;                  // bytes between K-adjacent core matrices
constexpr int TC_SBO = TC_CHUNKS * 128;      // bytes between 8-row groups (2304)
constexpr int TC_TILE_BYTES = (TC_M / 8) * TC_SBO;   // one 128-row operand tile, contiguous in the workspace (36 864)
constexpr int TC_TILE_WORDS = TC_TILE_BYTES / 4;
constexpr int TC_STAGES = 3;         // B tiles in flight (shared memory ring)
constexpr int TC_TBUFS = 4;          // TMEM accumulators of TC_N columns (all 512 columns: one CTA per SM)
constexpr int TC_EPI_WARPS = 8;      // two per TMEM lane quarter, each taking half of a tile's columns
constexpr int TC_THREADS = 64 + 32 * TC_EPI_WARPS;   // warp 0 producer, warp 1 MMA, warps 2-9 epilogue
constexpr float TC_KAPPA = 1.52587890625e-05f;   // 2^-16, see header; measured error is ~30x smaller
constexpr float TC_TIE_SLACK = 2.0e-6f;
constexpr float TC_PAD_NORM = 1.0e30f;       // |b|^2 of the padding rows of the last labeled tile: never a candidate

constexpr int TC_QDRAIN = 512;               // the queue is drained before a 16-column group is scanned if it holds more
constexpr int TC_QCAP = TC_QDRAIN + 32 * 16; // ... so a group (32 rows x 16 columns per warp) always fits
constexpr int TC_QJBITS = 27;                // queue entry: row lane << 27 | labeled row index

struct TcShared {
    alignas(128) uint32_t a[TC_TILE_WORDS];              // 36 KB, canonical K-major no-swizzle core-matrix layout
    alignas(128) uint32_t b[TC_STAGES][TC_TILE_WORDS];   // 3 x 36 KB
    alignas(8) unsigned long long row_key[TC_M];         // best (similarity, lowest index) per row of this CTA, packed
    alignas(8) uint2 queue[TC_EPI_WARPS][TC_QCAP];       // per warp: {row lane << 27 | labeled row, approximate d2 bits} (64 KB)
    int qn[TC_EPI_WARPS];                                // pairs left in each queue after the CTA's last tile
    float pair_max[2][2][TC_M];                          // [tile parity][column half][row]: largest accumulator of the half tile
    alignas(8) unsigned long long bar_a, bar_full[TC_STAGES], bar_empty[TC_STAGES], bar_tfull[TC_TBUFS], bar_tempty[TC_TBUFS];
    uint32_t tmem_base;
    int last_flag;
};
static_assert(sizeof(TcShared) <= 227 * 1024, "TcShared exceeds the shared memory of an SM");

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
__device__ __forceinline__ void mma_commit(uint32_t mbar) {      // arrives on mbar when every MMA issued so far has completed
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory");
}

__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}\n"
            : "=r"(done) : "r"(mbar), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void mbar_arrive(uint32_t mbar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mbar) : "memory");
}
// one contiguous tile, global -> shared, completion counted in bytes on mbar
__device__ __forceinline__ void bulk_load_tile(uint32_t smem_dst, const void *gsrc, uint32_t mbar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"((uint32_t)TC_TILE_BYTES) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_dst), "l"(gsrc),
                 "r"((uint32_t)TC_TILE_BYTES), "r"(mbar)
                 : "memory");
}

// row r, 16-byte chunk q of a K-major no-swizzle operand tile (in uint32 units)
__device__ __forceinline__ int tile_word(int r, int q) { return ((r >> 3) * TC_SBO + q * TC_LBO + (r & 7) * 16) >> 2; }

struct TcStats {                       // diagnostics in the workspace
    unsigned long long exact_evals;
    unsigned int max_err_ratio_bits;   // max |approx d2 - exact d2| / (|a|^2 + |b|^2), float bits
    unsigned int pad;
};

// workspace: [keys n_u x u64][tickets row_blocks x u32][TcStats][tile_nbmax tiles_l x f32][A' row_blocks tiles][B' tiles_l tiles]
struct TcLayout {
    int n_u, row_blocks, tiles_l;
    size_t off_tickets, off_stats, off_nbmax, off_a, off_b, head_bytes, total;
};
static inline size_t tc_up256(size_t x) { return (x + 255) / 256 * 256; }
static inline TcLayout tc_layout(int N, int n_l) {
    TcLayout Y;
    Y.n_u = N - n_l > 0 ? N - n_l : 0;
    Y.row_blocks = (Y.n_u + TC_M - 1) / TC_M;
    Y.tiles_l = (n_l + TC_N - 1) / TC_N;
    Y.off_tickets = tc_up256(sizeof(unsigned long long) * (size_t)Y.n_u);
    Y.off_stats = Y.off_tickets + tc_up256(sizeof(unsigned int) * (size_t)Y.row_blocks);
    Y.head_bytes = Y.off_stats + tc_up256(sizeof(TcStats));        // zeroed by the prep kernel on every call
    Y.off_nbmax = Y.head_bytes;
    Y.off_a = Y.off_nbmax + tc_up256(sizeof(float) * (size_t)Y.tiles_l);
    Y.off_b = Y.off_a + (size_t)Y.row_blocks * TC_TILE_BYTES;
    Y.total = Y.off_b + (size_t)Y.tiles_l * TC_TILE_BYTES;
    return Y;
}

#ifdef WESUP_TC_TRACE
__device__ long long g_tc_trace[512];
#define TC_STAMP(slot) do { if (blockIdx.x == 0 && blockIdx.y == 0) g_tc_trace[(slot)] = clock64(); } while (0)
#else
#define TC_STAMP(slot) do { } while (0)
#endif

__device__ __forceinline__ unsigned long long pack_key(float sim, int j) {
    return ((unsigned long long)__float_as_uint(sim) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)j);
}

// ---- prep: split operands, once per call -------------------------------------------------------------------------------
// block = one 128-row tile (labeled tiles first, then the unlabeled row blocks), thread = row.  A side [hi|lo|1 1 0 0|0],
// B side [hi|lo|n_hi n_lo 0 0|0] with n = -|b|^2/2.  Also zeroes the head of the workspace (keys, tickets, stats).
__global__ void __launch_bounds__(TC_M) label_propagate_tc_prep_kernel(const float *__restrict__ feats, int n_l, int n_u, int tiles_l,
                                                                        uint32_t *__restrict__ a_tiles, uint32_t *__restrict__ b_tiles,
                                                                        float *__restrict__ tile_nbmax, uint32_t *__restrict__ head,
                                                                        int head_words) {
    __shared__ float s_red[TC_M / 32];
    const int tile = blockIdx.x, r = threadIdx.x;
    asm volatile("griddepcontrol.launch_dependents;");   // the main kernel may start its prologue (it waits before reading)
    for (int i = tile * TC_M + r; i < head_words; i += gridDim.x * TC_M) head[i] = 0u;
    const bool is_b = tile < tiles_l;
    const int local = (is_b ? tile : tile - tiles_l) * TC_M + r;
    const bool have = local < (is_b ? n_l : n_u);
    uint32_t *dst = (is_b ? b_tiles : a_tiles) + (size_t)(is_b ? tile : tile - tiles_l) * TC_TILE_WORDS;
    const float4 *src = reinterpret_cast<const float4 *>(feats + (long)(have ? (is_b ? local : n_l + local) : 0) * TC_D);
    float nrm = 0.f;
#pragma unroll
    for (int q = 0; q < TC_D / 4; ++q) {
        const float4 v = have ? __ldg(src + q) : make_float4(0.f, 0.f, 0.f, 0.f);
        nrm = fmaf(v.x, v.x, nrm); nrm = fmaf(v.y, v.y, nrm); nrm = fmaf(v.z, v.z, nrm); nrm = fmaf(v.w, v.w, nrm);
        const uint4 hi = make_uint4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
        const uint4 lo = make_uint4(to_tf32(v.x - __uint_as_float(hi.x)), to_tf32(v.y - __uint_as_float(hi.y)),
                                    to_tf32(v.z - __uint_as_float(hi.z)), to_tf32(v.w - __uint_as_float(hi.w)));
        *reinterpret_cast<uint4 *>(dst + tile_word(r, q)) = hi;
        *reinterpret_cast<uint4 *>(dst + tile_word(r, q + 8)) = lo;
    }
    uint4 extra = make_uint4(__float_as_uint(1.0f), __float_as_uint(1.0f), 0u, 0u);
    if (is_b) {
        const float n = -0.5f * (have ? nrm : TC_PAD_NORM);
        const uint32_t n_hi = to_tf32(n);
        extra = make_uint4(n_hi, to_tf32(n - __uint_as_float(n_hi)), 0u, 0u);
    }
    *reinterpret_cast<uint4 *>(dst + tile_word(r, 16)) = extra;
    *reinterpret_cast<uint4 *>(dst + tile_word(r, 17)) = make_uint4(0u, 0u, 0u, 0u);
    if (is_b) {                                            // largest real norm of the tile (block reduction)
        float m = have ? nrm : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if ((r & 31) == 0) s_red[r >> 5] = m;
        __syncthreads();
        if (r == 0) tile_nbmax[tile] = fmaxf(fmaxf(s_red[0], s_red[1]), fmaxf(s_red[2], s_red[3]));
    }
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

#define TC_TMEM_LD16(d, addr)                                                                                                   \
    asm volatile(                                                                                                              \
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"               \
        : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6]), "=r"(d[7]), "=r"(d[8]), "=r"(d[9]),       \
          "=r"(d[10]), "=r"(d[11]), "=r"(d[12]), "=r"(d[13]), "=r"(d[14]), "=r"(d[15])                                               \
        : "r"(addr));                                                                                                          \
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory")

// largest of 16 accumulators, as a balanced tree (four independent chains: a single warp per scheduler lives on ILP)
__device__ __forceinline__ float tc_max16(const uint32_t *d) {
    float m[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) m[i] = fmaxf(__uint_as_float(d[2 * i]), __uint_as_float(d[2 * i + 1]));
    return fmaxf(fmaxf(fmaxf(m[0], m[1]), fmaxf(m[2], m[3])), fmaxf(fmaxf(m[4], m[5]), fmaxf(m[6], m[7])));
}

// The queued (row, labeled row) pairs of one epilogue warp, exact pairs two per lane and round (32 x 128-bit loads in flight
// per lane): both rows from global memory (L2), the reference's arithmetic in the reference's order, result to the row's
// packed maximum in shared memory.  Candidates are rare (about ten per row at the stress shape): whole rounds of 64 are
// evaluated at the end of a tile (beside the MMAs of the next tiles), the remainder after the CTA's last tile.
__device__ __forceinline__ void tc_eval_pair(TcShared &S, const uint2 e, const float *__restrict__ feats, long a_row0, int key_row0,
                                             float &max_ratio) {
    const int r = (int)(e.x >> TC_QJBITS), j = (int)(e.x & ((1u << TC_QJBITS) - 1u));
    const float4 *arow = reinterpret_cast<const float4 *>(feats + (a_row0 + r) * TC_D);
    const float4 *brow = reinterpret_cast<const float4 *>(feats + (long)j * TC_D);
    float4 av[TC_D / 4], bv[TC_D / 4];
#pragma unroll
    for (int q = 0; q < TC_D / 4; ++q) { av[q] = __ldg(arow + q); bv[q] = __ldg(brow + q); }
    float d2 = 0.f, sa = 0.f, sb = 0.f;
#pragma unroll
    for (int q = 0; q < TC_D / 4; ++q) {
        float df;
        df = av[q].x - bv[q].x; d2 = fmaf(df, df, d2);
        df = av[q].y - bv[q].y; d2 = fmaf(df, df, d2);
        df = av[q].z - bv[q].z; d2 = fmaf(df, df, d2);
        df = av[q].w - bv[q].w; d2 = fmaf(df, df, d2);
        sa = fmaf(av[q].x, av[q].x, sa); sa = fmaf(av[q].y, av[q].y, sa); sa = fmaf(av[q].z, av[q].z, sa); sa = fmaf(av[q].w, av[q].w, sa);
        sb = fmaf(bv[q].x, bv[q].x, sb); sb = fmaf(bv[q].y, bv[q].y, sb); sb = fmaf(bv[q].z, bv[q].z, sb); sb = fmaf(bv[q].w, bv[q].w, sb);
    }
    atomicMax(&S.row_key[key_row0 + r], pack_key(expf(-d2), j));
    if (sa + sb > 0.f) max_ratio = fmaxf(max_ratio, fabsf(__uint_as_float(e.y) - d2) / (sa + sb));
}
__device__ __noinline__ void tc_drain_queue(TcShared &S, int ew, int lane, const float *__restrict__ feats, long a_row0, int key_row0,
                                            int begin, int end, unsigned int &n_exact, float &max_ratio) {
    __syncwarp();                                           // the queue writes of the other lanes
    const uint2 *queue = S.queue[ew];
    for (int i0 = begin; i0 < end; i0 += 64) {
        const int i = i0 + lane;
        if (i < end) { tc_eval_pair(S, queue[i], feats, a_row0, key_row0, max_ratio); ++n_exact; }
        if (i + 32 < end) { tc_eval_pair(S, queue[i + 32], feats, a_row0, key_row0, max_ratio); ++n_exact; }
    }
    __syncwarp();
}

__global__ void __launch_bounds__(TC_THREADS, 1) label_propagate_tc_kernel(
    const float *__restrict__ feats, int N, int n_l, int tiles_per_split, const uint32_t *__restrict__ a_tiles,
    const uint32_t *__restrict__ b_tiles, const float *__restrict__ tile_nbmax, const float *__restrict__ y_l, int n_cls, float thr,
    float *__restrict__ y_u, int32_t *__restrict__ src_idx, float *__restrict__ max_sim, unsigned long long *__restrict__ keys,
    unsigned int *__restrict__ tickets, TcStats *__restrict__ stats) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    TcShared &S = *reinterpret_cast<TcShared *>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_u = N - n_l;
    const int tiles_l = (n_l + TC_N - 1) / TC_N;
    const int t_begin = blockIdx.y * tiles_per_split;
    const int n_tiles = max(0, min(tiles_l, t_begin + tiles_per_split) - t_begin);
    unsigned int n_exact = 0;                                                // diagnostics (every thread evaluates pairs at the end)
    float max_ratio = 0.f;

    // ---- one-time setup: TMEM allocation (all 512 columns), mbarriers -----------
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&S.tmem_base)), "n"(TC_TBUFS * TC_N));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 32) {
        mbar_init(smem_u32(&S.bar_a), 1);
        for (int s = 0; s < TC_STAGES; ++s) { mbar_init(smem_u32(&S.bar_full[s]), 1); mbar_init(smem_u32(&S.bar_empty[s]), 1); }
        for (int b = 0; b < TC_TBUFS; ++b) { mbar_init(smem_u32(&S.bar_tfull[b]), 1); mbar_init(smem_u32(&S.bar_tempty[b]), 32 * TC_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < TC_M) S.row_key[tid] = 0ull;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = S.tmem_base;
    if (tid == 0) TC_STAMP(0);

    if (warp == 0) {
        // ---- producer: A tile, then the B tiles through the ring ------------------
        asm volatile("griddepcontrol.wait;" ::: "memory");           // the prep kernel's tiles / zeroed tickets are complete
        if (lane == 0) TC_STAMP(1);
        if (lane == 0 && n_tiles > 0) {
            bulk_load_tile(smem_u32(S.a), a_tiles + (size_t)blockIdx.x * TC_TILE_WORDS, smem_u32(&S.bar_a));
            for (int t = 0; t < n_tiles; ++t) {
                const int s = t % TC_STAGES;
                if (t >= TC_STAGES) mbar_wait(smem_u32(&S.bar_empty[s]), (uint32_t)((t / TC_STAGES - 1) & 1));
                TC_STAMP(32 + t);
                bulk_load_tile(smem_u32(S.b[s]), b_tiles + (size_t)(t_begin + t) * TC_TILE_WORDS, smem_u32(&S.bar_full[s]));
            }
        }
    } else if (warp == 1) {
        // ---- MMA issuer: D[128 x 128] = A' . B'^T, 13 x (M128, N128, K8) kind::tf32 per tile ----
        if (lane == 0 && n_tiles > 0) {
            const uint64_t adesc0 = make_smem_desc(smem_u32(S.a));
            mbar_wait(smem_u32(&S.bar_a), 0u);
            for (int t = 0; t < n_tiles; ++t) {
                const int s = t % TC_STAGES, buf = t % TC_TBUFS;
                mbar_wait(smem_u32(&S.bar_full[s]), (uint32_t)((t / TC_STAGES) & 1));
                TC_STAMP(64 + t);
                if (t >= TC_TBUFS) mbar_wait(smem_u32(&S.bar_tempty[buf]), (uint32_t)((t / TC_TBUFS - 1) & 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint64_t bdesc0 = make_smem_desc(smem_u32(S.b[s]));
                const uint32_t d = tmem + (uint32_t)(buf * TC_N);
#pragma unroll
                for (int k = 0; k < TC_KSTEPS; ++k) {
                    // K chunk (16 bytes = 4 values) each operand starts at, two chunks per instruction:
                    // k 0-3 a_hi.b_hi, 4-7 a_hi.b_lo, 8-11 a_lo.b_hi, 12 [1 1].[n_hi n_lo]
                    const int ca = k < 4 ? 2 * k : k < 8 ? 2 * (k - 4) : k < 12 ? 8 + 2 * (k - 8) : 16;
                    const int cb = k < 4 ? 2 * k : k < 8 ? 8 + 2 * (k - 4) : k < 12 ? 2 * (k - 8) : 16;
                    mma_tf32(d, adesc0 + (uint64_t)((ca * TC_LBO) >> 4), bdesc0 + (uint64_t)((cb * TC_LBO) >> 4), k > 0 ? 1u : 0u);
                }
                mma_commit(smem_u32(&S.bar_empty[s]));      // the stage may be refilled once these MMAs have read it
                mma_commit(smem_u32(&S.bar_tfull[buf]));    // ... and the accumulator is complete
                TC_STAMP(96 + t);
            }
        }
    } else {
        // ---- epilogue warps: thread = row = TMEM lane; warps w and w + 4 share the rows and split a tile's columns ----
        const int ew = warp - 2;
        const int quarter = warp & 3;                                        // the 32 TMEM lanes this warp may access
        const int half = ew >> 2;                                            // columns [64 half, 64 half + 64) of every tile
        const int row = quarter * 32 + lane;
        const int u = blockIdx.x * TC_M + row;
        const bool live = u < n_u;
        const long a_row0 = (long)n_l + (long)blockIdx.x * TC_M + quarter * 32;
        float na = 0.f;
        if (live) {
            const float4 *src = reinterpret_cast<const float4 *>(feats + (long)(n_l + u) * TC_D);
#pragma unroll
            for (int q = 0; q < TC_D / 4; ++q) {
                const float4 v = __ldg(src + q);
                na = fmaf(v.x, v.x, na); na = fmaf(v.y, v.y, na); na = fmaf(v.z, v.z, na); na = fmaf(v.w, v.w, na);
            }
        }
        asm volatile("griddepcontrol.wait;" ::: "memory");           // tile_nbmax and the zeroed keys come from the prep kernel
        const uint32_t tmem_row = tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(half * (TC_N / 2));
        float run_min = 3.0e38f, nb_max = 0.f;
        uint2 *queue = S.queue[ew];
        int qn = 0;                                                          // queued pairs (warp-uniform)
        const unsigned lanes_below = (1u << lane) - 1u, lane_tag = (unsigned)lane << TC_QJBITS;
        for (int t = 0; t < n_tiles; ++t) {
            const int buf = t % TC_TBUFS;
            const int j0 = (t_begin + t) * TC_N + half * (TC_N / 2);
            nb_max = fmaxf(nb_max, __ldg(tile_nbmax + t_begin + t));
            mbar_wait(smem_u32(&S.bar_tfull[buf]), (uint32_t)((t / TC_TBUFS) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (ew == 0 && lane == 0) TC_STAMP(128 + t);
            const uint32_t tacc = tmem_row + (uint32_t)(buf * TC_N);
            // pass 1: this warp's 64 accumulators of the tile stay in registers; largest of every 16-column group = nearest
            // column (approximate d2 = na - 2 acc).  The TMEM buffer goes back to the MMA warp right away.
            uint32_t d[TC_N / 2];
            TC_TMEM_LD32(d, tacc);
            TC_TMEM_LD32((d + 32), tacc + 32u);
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(smem_u32(&S.bar_tempty[buf]));
            float gmax[TC_N / 32];
#pragma unroll
            for (int g = 0; g < TC_N / 32; ++g) gmax[g] = tc_max16(d + 16 * g);
            const float hmax = fmaxf(fmaxf(gmax[0], gmax[1]), fmaxf(gmax[2], gmax[3]));
            S.pair_max[t & 1][half][row] = hmax;
            asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");  // the two warps of this lane quarter
            const float tmax = fmaxf(hmax, S.pair_max[t & 1][half ^ 1][row]);
            if (ew == 0 && lane == 0) TC_STAMP(160 + t);
            const float gate = fminf(run_min, fmaf(-2.0f, tmax, na));
            // candidate: approximate d2 <= gate + (error bound of this column + of the column holding the minimum) + slack,
            // both bounds taken with the largest labeled norm seen so far (a superset of the per-column test)
            const float acc_thr = 0.5f * (na - (gate + (2.0f * TC_KAPPA * (na + nb_max) + TC_TIE_SLACK)));
            // pass 2: queue the candidates of the groups that have any.  One ballot per column; the queue position is
            // the running count + the rank among the voting lanes -- no atomics, no divergence on the common path
#pragma unroll
            for (int g = 0; g < TC_N / 32; ++g) {
                const bool mine = live && gmax[g] >= acc_thr;
                if (__any_sync(0xffffffffu, mine)) {
                    if (qn > TC_QDRAIN) {
                        tc_drain_queue(S, ew, lane, feats, a_row0, quarter * 32, 0, qn, n_exact, max_ratio);
                        qn = 0;
                    }
#pragma unroll
                    for (int i = 0; i < 16; ++i) {                          // branch-free: ballot, rank, predicated store
                        const float acc = __uint_as_float(d[16 * g + i]);
                        const bool hit = mine && acc >= acc_thr;
                        const unsigned m = __ballot_sync(0xffffffffu, hit);
                        const uint2 e = make_uint2(lane_tag | (unsigned)(j0 + g * 16 + i), __float_as_uint(fmaf(-2.0f, acc, na)));
                        if (hit) queue[qn + __popc(m & lanes_below)] = e;
                        qn += __popc(m);
                    }
                }
            }
            if (ew == 0 && lane == 0) TC_STAMP(192 + t);
            run_min = gate;
            if (qn >= 64) {                                                  // full rounds now, beside the next tiles' MMAs
                const int cnt = qn & ~63;
                tc_drain_queue(S, ew, lane, feats, a_row0, quarter * 32, qn - cnt, qn, n_exact, max_ratio);
                qn -= cnt;
            }
        }
        if (lane == 0) S.qn[ew] = qn;                                        // what is left goes to the CTA-wide round below
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid == 0) TC_STAMP(2);
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TC_TBUFS * TC_N));
    }
    // ---- the pairs still queued, over ALL threads of the CTA (the producer and MMA warps are idle by now): one pair per
    // thread instead of two rounds per epilogue warp at the very end of the kernel
    {
        int off[TC_EPI_WARPS + 1];
        off[0] = 0;
#pragma unroll
        for (int q = 0; q < TC_EPI_WARPS; ++q) off[q + 1] = off[q] + S.qn[q];
        for (int i = tid; i < off[TC_EPI_WARPS]; i += TC_THREADS) {
            int q = 0;
#pragma unroll
            for (int k = 1; k < TC_EPI_WARPS; ++k) q += (i >= off[k]) ? 1 : 0;
            int base = 0;
#pragma unroll
            for (int k = 1; k < TC_EPI_WARPS; ++k) base = (q == k) ? off[k] : base;
            const int quarter_q = (q + 2) & 3;                               // TMEM lane quarter of epilogue warp q
            tc_eval_pair(S, S.queue[q][i - base], feats, (long)n_l + (long)blockIdx.x * TC_M + quarter_q * 32, quarter_q * 32, max_ratio);
            ++n_exact;
        }
    }
    if (stats != nullptr) {
        for (int o = 16; o > 0; o >>= 1) {
            n_exact += __shfl_xor_sync(0xffffffffu, n_exact, o);
            max_ratio = fmaxf(max_ratio, __shfl_xor_sync(0xffffffffu, max_ratio, o));
        }
        if (lane == 0 && n_exact > 0) {
            atomicAdd(&stats->exact_evals, (unsigned long long)n_exact);
            atomicMax(&stats->max_err_ratio_bits, __float_as_uint(max_ratio));
        }
    }
    __syncthreads();
    if (tid == 0) TC_STAMP(3);
    // ---- merge the labeled-dimension splits ------------------------------------
    if (tid < TC_M && blockIdx.x * TC_M + tid < n_u && n_tiles > 0) atomicMax(keys + blockIdx.x * TC_M + tid, S.row_key[tid]);
    __threadfence();
    __syncthreads();
    if (tid == 0) TC_STAMP(4);
    if (tid == 0) S.last_flag = (atomicAdd(tickets + blockIdx.x, 1u) == gridDim.y - 1) ? 1 : 0;
    __syncthreads();
    if (S.last_flag && tid < TC_M) {
        const int u = blockIdx.x * TC_M + tid;
        if (u < n_u) {
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
}

}  // namespace wesup

using namespace wesup;

extern "C" size_t wesup_label_propagate_tc_workspace_bytes(int N, int D, int n_l) {
    (void)D;
    if ((long)N - n_l <= 0 || n_l <= 0) return 256;
    return tc_layout(N, n_l).total;
}

extern "C" int wesup_label_propagate_tc(const float *feats, int N, int D, int n_l, const float *y_l, int n_cls, float thr,
                                        float *y_u, int32_t *src_idx, float *max_sim, void *ws, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    WESUP_REQUIRE(feats && y_l && y_u && ws, WESUP_E_ARG, "wesup_label_propagate_tc: null pointer");
    WESUP_REQUIRE(N > 0 && n_cls > 0, WESUP_E_ARG, "wesup_label_propagate_tc: bad size N=%d n_cls=%d", N, n_cls);
    WESUP_REQUIRE(D == TC_D, WESUP_E_UNSUPPORTED, "wesup_label_propagate_tc: the tensor-core path is built for D=%d (got %d)", TC_D, D);
    WESUP_REQUIRE(n_l > 0 && n_l <= N, WESUP_E_ARG, "wesup_label_propagate_tc: n_l=%d must be in [1,N=%d]", n_l, N);
    WESUP_REQUIRE(n_l < (1 << TC_QJBITS), WESUP_E_UNSUPPORTED, "wesup_label_propagate_tc: n_l=%d exceeds %d", n_l, (1 << TC_QJBITS) - 1);
    WESUP_REQUIRE(aligned16(feats) && aligned16(ws), WESUP_E_ALIGN,
                  "wesup_label_propagate_tc: feats/ws must be 16-byte aligned");
    const TcLayout Y = tc_layout(N, n_l);
    if (Y.n_u == 0) return 0;
    // one CTA per SM (its shared memory holds three operand tiles): split the labeled tiles so that one wave covers the chip
    int splits = kNumSMs / Y.row_blocks;
    if (splits > Y.tiles_l) splits = Y.tiles_l;
    if (splits < 1) splits = 1;
    const int tiles_per_split = (Y.tiles_l + splits - 1) / splits;
    splits = (Y.tiles_l + tiles_per_split - 1) / tiles_per_split;      // no empty splits
    char *base = static_cast<char *>(ws);
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(base);
    unsigned int *tickets = reinterpret_cast<unsigned int *>(base + Y.off_tickets);
    TcStats *stats = reinterpret_cast<TcStats *>(base + Y.off_stats);
    float *tile_nbmax = reinterpret_cast<float *>(base + Y.off_nbmax);
    uint32_t *a_tiles = reinterpret_cast<uint32_t *>(base + Y.off_a);
    uint32_t *b_tiles = reinterpret_cast<uint32_t *>(base + Y.off_b);
    {   // per call, like the other kernels: the attribute belongs to the current device, a process may drive several
        cudaError_t e = cudaFuncSetAttribute(label_propagate_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TcShared));
        WESUP_REQUIRE(e == cudaSuccess, (int)e, "wesup_label_propagate_tc: smem attribute: %s", cudaGetErrorString(e));
    }
    label_propagate_tc_prep_kernel<<<Y.tiles_l + Y.row_blocks, TC_M, 0, stream>>>(feats, n_l, Y.n_u, Y.tiles_l, a_tiles, b_tiles, tile_nbmax,
                                                                                   reinterpret_cast<uint32_t *>(base), (int)(Y.head_bytes / 4));
    // programmatic dependent launch: the main kernel's prologue (TMEM allocation, barriers, row norms) overlaps the prep
    // kernel; it executes griddepcontrol.wait before touching anything the prep kernel writes
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(Y.row_blocks, splits);
    cfg.blockDim = dim3(TC_THREADS);
    cfg.dynamicSmemBytes = sizeof(TcShared);
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const uint32_t *a_c = a_tiles, *b_c = b_tiles;
    const float *nb_c = tile_nbmax;
    cudaError_t le = cudaLaunchKernelEx(&cfg, label_propagate_tc_kernel, feats, N, n_l, tiles_per_split, a_c, b_c, nb_c, y_l, n_cls, thr, y_u,
                                        src_idx, max_sim, keys, tickets, stats);
    WESUP_REQUIRE(le == cudaSuccess, (int)le, "wesup_label_propagate_tc: launch: %s", cudaGetErrorString(le));
    WESUP_CHECK_LAUNCH("wesup_label_propagate_tc", 2);
    return 0;
}

#ifdef WESUP_TC_TRACE
extern "C" int wesup_debug_tc_trace(long long *out_host) {
    return (int)cudaMemcpyFromSymbol(out_host, g_tc_trace, sizeof(long long) * 512);
}
#endif

// diagnostics of the most recent call that used `ws` (read after synchronising the stream):
// out[0] = exact re-evaluations, out[1] = max |approx - exact| / (|a|^2+|b|^2) as float bits
extern "C" int wesup_label_propagate_tc_stats(const void *ws, int N, int n_l, unsigned long long *out_host) {
    WESUP_REQUIRE(ws && out_host, WESUP_E_ARG, "wesup_label_propagate_tc_stats: null pointer");
    WESUP_REQUIRE(N - n_l > 0 && n_l > 0, WESUP_E_ARG, "wesup_label_propagate_tc_stats: no unlabeled rows");
    const TcLayout Y = tc_layout(N, n_l);
    TcStats s;
    cudaError_t e = cudaMemcpy(&s, static_cast<const char *>(ws) + Y.off_stats, sizeof(TcStats), cudaMemcpyDeviceToHost);
    WESUP_REQUIRE(e == cudaSuccess, (int)e, "wesup_label_propagate_tc_stats: %s", cudaGetErrorString(e));
    out_host[0] = s.exact_evals;
    out_host[1] = s.max_err_ratio_bits;
    return 0;
}
