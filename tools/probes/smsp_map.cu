// Probe: which hardware warp slots share a scheduler (SM sub-core)?  Warp 0 runs a latency-bound dependent chain; one other
// warp k runs an issue-bound loop; the chain slows down only when both sit on the same sub-core.
//   nvcc -arch=sm_100a -O3 -o smsp_map smsp_map.cu && ./smsp_map
#include <cstdio>
#include <cuda_runtime.h>
__global__ void probe(int k, long long *out, unsigned *slots, volatile int *stop) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned slot; asm volatile("mov.u32 %0, %%warpid;" : "=r"(slot));
    if (lane == 0) slots[warp] = slot;
    __shared__ unsigned tab[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) tab[i] = (i * 7 + 3) & 1023;
    __syncthreads();
    if (warp == 0) {
        unsigned x = lane;
        long long t0 = clock64();
        for (int i = 0; i < 200000; i++) { x = tab[x]; x = (x * 5 + 1) & 1023; }
        long long t1 = clock64();
        if (lane == 0) { out[0] = t1 - t0; out[1] = x; *stop = 1; }
    } else if (warp == k) {
        unsigned a = lane, b = lane + 1, c = lane + 2, d = lane + 3;
        while (!*stop) {
#pragma unroll
            for (int j = 0; j < 64; j++) { a = a * 3 + 1; b = b * 5 + 1; c = c * 7 + 1; d = d * 9 + 1; }
        }
        if (lane == 0) out[2] = a + b + c + d;
    }
}
int main() {
    long long *out; unsigned *slots; int *stop;
    cudaMallocManaged(&out, 64); cudaMallocManaged(&slots, 32 * 4); cudaMallocManaged(&stop, 4);
    for (int k = 0; k < 32; k++) {
        *stop = 0; out[0] = 0;
        probe<<<1, 1024>>>(k == 0 ? 99 : k, out, slots, stop);
        cudaDeviceSynchronize();
        printf("aggressor warp %2d (slot %2u)  victim warp 0 (slot %2u): %lld cycles\n", k, k ? slots[k] : 0, slots[0], out[0]);
    }
    return 0;
}
