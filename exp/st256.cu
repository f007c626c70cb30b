#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void st256(void* p, unsigned a, unsigned b, unsigned c, unsigned d, unsigned e, unsigned f, unsigned g, unsigned h) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d), "r"(e), "r"(f), "r"(g), "r"(h) : "memory");
}
__device__ __forceinline__ void ld256(const void* p, unsigned* w) {
    asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "l"(p));
}
__global__ void fill32v8(unsigned* p, size_t n_cells, int cells_per_thread) {
    size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t lane = tid % 32, warp = tid / 32;
    for (int c = 0; c < cells_per_thread; c++) {
        size_t cell = (warp * cells_per_thread + c) * 32 + lane;
        if (cell < n_cells) st256(p + 8 * cell, tid, 1, 2, 3, 4, 5, 6, 7);
    }
}
__global__ void copy_check(const unsigned* p, unsigned* out) {
    unsigned w[8]; ld256(p + 8 * threadIdx.x, w);
    out[threadIdx.x] = w[0] + w[7];
}
int main() {
    size_t bytes = 4ull << 30;
    unsigned* p; cudaMalloc(&p, bytes);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int cpt : {8, 32, 133, 213}) {
        size_t n_cells = bytes / 32;
        size_t threads = (n_cells + cpt - 1) / cpt;
        threads = (threads + 31) / 32 * 32;
        for (int it = 0; it < 3; it++) fill32v8<<<(threads + 127) / 128, 128>>>(p, n_cells, cpt);
        cudaEventRecord(e0);
        for (int it = 0; it < 10; it++) fill32v8<<<(threads + 127) / 128, 128>>>(p, n_cells, cpt);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("fill32 v8 cells/thread=%d: %.1f GB/s (%s)\n", cpt, bytes / (ms / 10) / 1e6, cudaGetErrorString(cudaGetLastError()));
    }
    unsigned* out; cudaMalloc(&out, 128);
    copy_check<<<1, 32>>>(p, out);
    unsigned h[32]; cudaMemcpy(h, out, 128, cudaMemcpyDeviceToHost);
    printf("check %u %u\n", h[0], h[5]);
    return 0;
}
