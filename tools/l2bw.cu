// L2 -> SM read bandwidth of one B200: a grid-stride float4 read (ld.global.cg: L1 bypassed) of a buffer of
// S bytes repeated R times in ONE launch.  S below the L2 capacity measures the L2 -> SM fabric, S well above
// it measures HBM.  The pooling gathers over precomputed footprints re-read shared boundary cells from L2,
// so their bound is min(HBM bytes / HBM rate, L2->SM bytes / this rate) -- see DESIGN.md section 3.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/l2bw.cu -o tools/_l2bw && tools/_l2bw
#include <cuda_runtime.h>
#include <stdio.h>

__global__ void __launch_bounds__(512) read_kernel(const float4 *__restrict__ p, size_t n4, int reps, float4 *out) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int r = 0; r < reps; ++r) {
        size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
        for (; i + 3 * stride < n4; i += 4 * stride) {
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w) : "l"(p + i + u * stride));
#pragma unroll
            for (int u = 0; u < 4; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
        }
        for (; i < n4; i += stride) {
            float4 v;
            asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p + i));
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
    }
    if (acc.x == 1234.5f) out[0] = acc;
}

// HBM gather: every warp reads whole chunks of `chunk` bytes at pseudo-random chunk-aligned offsets of a 1 GB buffer
// (four 512-byte warp loads in flight).  chunk = 128 KB behaves like a stream; a few KB is what the footprint gathers
// issue (one bounding-box row of cells of one level: 3.5 - 8 KB).
__global__ void __launch_bounds__(256) gather_kernel(const float4 *__restrict__ p, size_t n_chunks_buf, int chunk_f4, size_t n_chunks_read,
                                                     float4 *out) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const int lane = threadIdx.x & 31;
    const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((size_t)gridDim.x * blockDim.x) >> 5;
    for (size_t c = warp; c < n_chunks_read; c += n_warps) {
        const size_t h = (c * 0x9E3779B97F4A7C15ull) >> 20;
        const float4 *src = p + (h % n_chunks_buf) * (size_t)chunk_f4;
        int i = lane;
        for (; i + 96 < chunk_f4; i += 128) {
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w) : "l"(src + i + 32 * u));
#pragma unroll
            for (int u = 0; u < 4; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
        }
        for (; i < chunk_f4; i += 32) {
            float4 v;
            asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(src + i));
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
    }
    if (acc.x == 1234.5f) out[0] = acc;
}

int main() {
    const size_t max_bytes = (size_t)1 << 30;
    float4 *buf, *out;
    if (cudaMalloc(&buf, max_bytes) != cudaSuccess || cudaMalloc(&out, 64) != cudaSuccess) { printf("{\"error\": \"alloc\"}\n"); return 1; }
    cudaMemset(buf, 0, max_bytes);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    const int mbs[] = {8, 16, 32, 48, 64, 96, 128, 256, 1024};
    printf("{");
    for (int t = 0; t < 9; ++t) {
        const size_t bytes = (size_t)mbs[t] << 20, n4 = bytes / 16;
        const int reps = mbs[t] <= 128 ? 40 : 4;
        read_kernel<<<148 * 4, 512>>>(buf, n4, 2, out);            // warm the L2
        cudaEventRecord(a);
        read_kernel<<<148 * 4, 512>>>(buf, n4, reps, out);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, a, b);
        printf("%s\"read_%dMB_gbs\": %.1f", t ? ", " : "", mbs[t], (double)bytes * reps / ms / 1e6);
    }
    const int chunks[] = {512, 1024, 2048, 4096, 8192, 32768, 131072};
    for (int t = 0; t < 7; ++t) {
        const int chunk_f4 = chunks[t] / 16;
        const size_t n_buf = max_bytes / chunks[t], n_read = ((size_t)512 << 20) / chunks[t];
        gather_kernel<<<148 * 8, 256>>>(buf, n_buf, chunk_f4, n_read / 8, out);
        cudaEventRecord(a);
        gather_kernel<<<148 * 8, 256>>>(buf, n_buf, chunk_f4, n_read, out);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, a, b);
        printf(", \"gather_%dB_gbs\": %.1f", chunks[t], (double)n_read * chunks[t] / ms / 1e6);
    }
    printf("}\n");
    return cudaGetLastError() != cudaSuccess;
}
