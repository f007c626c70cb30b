// pure-store bandwidth probe: each thread writes 32 B cells, warp writes 1 KiB contiguous (same pattern as the VM)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void fill32(uint4* p, size_t n_cells, int cells_per_thread) {
    size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t lane = tid % 32, warp = tid / 32;
    uint4 v = make_uint4(tid, 1, 2, 3);
    for (int c = 0; c < cells_per_thread; c++) {
        size_t cell = (warp * cells_per_thread + c) * 32 + lane;
        if (cell < n_cells) { p[2 * cell] = v; p[2 * cell + 1] = v; }
    }
}
__global__ void fill_stream(uint4* p, size_t n16) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    uint4 v = make_uint4(i, 1, 2, 3);
    for (; i < n16; i += stride) p[i] = v;
}
int main() {
    size_t bytes = 4ull << 30;
    uint4* p; cudaMalloc(&p, bytes);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int cpt : {8, 32, 133, 213}) {
        size_t n_cells = bytes / 32;
        size_t threads = (n_cells + cpt - 1) / cpt;
        threads = (threads + 31) / 32 * 32;
        for (int it = 0; it < 3; it++) fill32<<<(threads + 127) / 128, 128>>>(p, n_cells, cpt);
        cudaEventRecord(e0);
        for (int it = 0; it < 10; it++) fill32<<<(threads + 127) / 128, 128>>>(p, n_cells, cpt);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("fill32 cells/thread=%d: %.1f GB/s\n", cpt, bytes / (ms / 10) / 1e6);
    }
    for (int blocks : {148 * 8, 148 * 16, 148 * 32}) {
        for (int it = 0; it < 3; it++) fill_stream<<<blocks, 256>>>(p, bytes / 16);
        cudaEventRecord(e0);
        for (int it = 0; it < 10; it++) fill_stream<<<blocks, 256>>>(p, bytes / 16);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("fill_stream blocks=%d: %.1f GB/s\n", blocks, bytes / (ms / 10) / 1e6);
    }
    cudaMemset(p, 0, bytes);
    cudaEventRecord(e0);
    for (int it = 0; it < 10; it++) cudaMemsetAsync(p, 0, bytes);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("cudaMemset: %.1f GB/s\n", bytes / (ms / 10) / 1e6);
    return 0;
}
