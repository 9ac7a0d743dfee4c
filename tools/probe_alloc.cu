// Development probe (not product): what the one-off allocations of a first frame cost on this box.
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
static double now() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main() {
    double t0 = now();
    cudaFree(0);
    printf("context_ms %.2f\n", now() - t0);
    for (size_t gb : { 1, 4, 9 }) {
        void *p = nullptr;
        t0 = now();
        cudaMalloc(&p, gb << 30);
        double t1 = now();
        cudaFree(p);
        printf("cudaMalloc_%zuGB_ms %.2f free_ms %.2f\n", gb, t1 - t0, now() - t1);
    }
    size_t bytes = 800 * 800 * 32;
    void *h = malloc(bytes), *d = nullptr;
    memset(h, 1, bytes);
    cudaMalloc(&d, bytes);
    cudaMemset(d, 0, bytes);
    cudaDeviceSynchronize();
    for (int rep = 0; rep < 3; ++rep) {
        t0 = now();
        cudaMemcpy(h, d, bytes, cudaMemcpyDeviceToHost);
        printf("d2h_pageable_20MB_ms %.2f\n", now() - t0);
    }
    t0 = now();
    cudaHostRegister(h, bytes, cudaHostRegisterPortable);
    printf("hostRegister_20MB_ms %.2f\n", now() - t0);
    for (int rep = 0; rep < 3; ++rep) {
        t0 = now();
        cudaMemcpy(h, d, bytes, cudaMemcpyDeviceToHost);
        printf("d2h_pinned_20MB_ms %.2f\n", now() - t0);
    }
    t0 = now();
    cudaHostUnregister(h);
    printf("hostUnregister_ms %.2f\n", now() - t0);
    void *hp = nullptr;
    t0 = now();
    cudaHostAlloc(&hp, bytes, cudaHostAllocPortable);
    printf("hostAlloc_20MB_ms %.2f\n", now() - t0);
    return 0;
}
