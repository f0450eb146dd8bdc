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
    printf("}\n");
    return cudaGetLastError() != cudaSuccess;
}
