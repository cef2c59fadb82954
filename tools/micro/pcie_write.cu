// Microbenchmark: bandwidth of SM-issued stores into pinned host memory vs a DMA copy (nvcc -O3 -arch=sm_100a).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int W> __global__ void k_write(uint4 *dst, size_t n16, int chunk16) {
    // each CTA takes contiguous chunks of chunk16 uint4 (like one replicate's curves), grid-stride over chunks
    for (size_t c = blockIdx.x; c * chunk16 < n16; c += gridDim.x) {
        uint4 *p = dst + c * chunk16;
        for (int k = threadIdx.x; k < chunk16; k += blockDim.x) {
            uint4 v = make_uint4((uint32_t)k, (uint32_t)c, 3u, 4u);
            if (W == 0) p[k] = v;
            else if (W == 1) __stcs(p + k, v);
            else asm volatile("st.global.wt.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p + k), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
        }
    }
}
int main() {
    const size_t bytes = 144000000, n16 = bytes / 16;
    uint4 *h, *d; cudaHostAlloc(&h, bytes, cudaHostAllocDefault); cudaMalloc(&d, bytes); cudaMemset(d, 1, bytes);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b); float ms;
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(a); cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost); cudaEventRecord(b); cudaEventSynchronize(b);
        cudaEventElapsedTime(&ms, a, b); printf("DMA D2H: %.3f ms  %.1f GB/s\n", ms, bytes / ms * 1e-6);
    }
    const int grids[] = {8, 16, 32, 74, 148, 296}; const int chunk16 = 144000 / 16;
    for (int g : grids) for (int w = 0; w < 3; ++w) for (int nt : {256, 1024}) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(a);
            if (w == 0) k_write<0><<<g, nt>>>(h, n16, chunk16); else if (w == 1) k_write<1><<<g, nt>>>(h, n16, chunk16); else k_write<2><<<g, nt>>>(h, n16, chunk16);
            cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
        }
        printf("SM stores grid %3d x %4d mode %d: %.3f ms  %.1f GB/s\n", g, nt, w, ms, bytes / ms * 1e-6);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
