// Probe: what does one step of a shared-memory dependent chain cost on this GPU, for the access pattern of the FSE producer warp
// (k_seq: one warp, 28 active lanes, tables interleaved across the lanes so that bank = lane)?
//   A  one dependent LDS + 2 ALU per step                         -> LDS latency as the chain sees it
//   B  three LDS whose results are summed, then 4 dependent ALU    -> the register-window design (cells -> sum -> shifts -> address)
//   C  B + a second dependent LDS at a data-dependent address      -> the round-1 design (cells -> position -> ring words -> bits)
// each alone and beside 16 busy warps (shuffles + shared-memory loads, like phase 2).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o chain_lat chain_lat.cu && ./chain_lat
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned lds(unsigned a) { unsigned v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }

template <int MODE>
__global__ void __launch_bounds__(544, 1) probe(int steps, int noise, long long *out, volatile int *stop) {
    extern __shared__ __align__(1024) unsigned sm[];          // [3][512][32] tables + ring [32][128]
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (unsigned i = threadIdx.x; i < 3 * 512 * 32; i += blockDim.x) { unsigned r = i * 2654435761u; sm[i] = r ^ (r >> 13); }
    unsigned *ring = sm + 3 * 512 * 32;
    for (unsigned i = threadIdx.x; i < 32 * 128; i += blockDim.x) ring[i] = i * 40503u + 7u;
    __syncthreads();
    const unsigned t0a = (unsigned)__cvta_generic_to_shared(sm) + lane * 4, t1a = t0a + 512 * 128, t2a = t1a + 512 * 128;
    const unsigned ra = (unsigned)__cvta_generic_to_shared(ring) + lane * 512;
    if (warp == 0) {
        if (lane < 28) {
            unsigned a = t0a + (lane * 37 & 511) * 128, b = t1a + (lane * 11 & 511) * 128, c = t2a + (lane * 5 & 511) * 128, top = 4000 * 8;
            long long c0 = clock64();
            if (MODE == 0) {
                for (int i = 0; i < steps; i++) { const unsigned v = lds(a); a = t0a + ((v >> 7) & 511) * 128; }
            } else if (MODE == 1) {
                for (int i = 0; i < steps; i++) {
                    const unsigned x = lds(a), y = lds(b), z = lds(c);
                    const unsigned s = x + y + z;
                    const unsigned t = __funnelshift_l(y, x, s);
                    const unsigned u = (s & 32) ? t : z;
                    a = t0a + (__funnelshift_l(u, 0, x) & 511) * 128; b = t1a + (__funnelshift_l(u, 0, y) & 511) * 128; c = t2a + (__funnelshift_l(u, 0, z) & 511) * 128;
                }
            } else {
                for (int i = 0; i < steps; i++) {
                    const unsigned x = lds(a), y = lds(b), z = lds(c);
                    const unsigned s = x + y + z;
                    const unsigned e = (unsigned)__dp4a((int)s, 0x0000FF00, (int)top - 32);
                    const unsigned w0 = lds(((e >> 3) & 0x1FC) | ra), w1 = lds((((e >> 3) + 4) & 0x1FC) | ra);
                    const unsigned u = __funnelshift_r(w0, w1, e);
                    a = t0a + (__funnelshift_l(u, 0, x) & 511) * 128; b = t1a + (__funnelshift_l(u, 0, y) & 511) * 128; c = t2a + (__funnelshift_l(u, 0, z) & 511) * 128;
                    top = (unsigned)__dp4a((int)s, 0x00000101, (int)top) & 0x7FFF;
                }
            }
            long long c1 = clock64();
            if (lane == 0) { out[0] = c1 - c0; out[1] = a + b + c + top; }
        }
        __syncwarp();
        if (lane == 0) *stop = 1;
    } else if (noise) {
        unsigned x = lane, acc = 0;
        while (!*stop) {
#pragma unroll 8
            for (int j = 0; j < 32; j++) { x = __shfl_up_sync(0xFFFFFFFFu, x, 1) + sm[(x * 33 + lane) & 16383]; acc += x; }
        }
        if (acc == 12345) out[3] = acc;
    }
}

int main() {
    long long *out; int *stop;
    cudaMallocManaged(&out, 64); cudaMallocManaged(&stop, 4);
    const size_t smem = (3 * 512 * 32 + 32 * 128) * 4;
    cudaFuncSetAttribute(probe<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(probe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(probe<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int steps = 100000;
    for (int noise = 0; noise < 2; noise++)
        for (int mode = 0; mode < 3; mode++) {
            *stop = 0; out[0] = 0;
            if (mode == 0) probe<0><<<1, 544, smem>>>(steps, noise, out, stop);
            if (mode == 1) probe<1><<<1, 544, smem>>>(steps, noise, out, stop);
            if (mode == 2) probe<2><<<1, 544, smem>>>(steps, noise, out, stop);
            cudaError_t e = cudaDeviceSynchronize();
            printf("mode %c noise %d: %.1f cycles/step (%s)\n", 'A' + mode, noise, (double)out[0] / steps, cudaGetErrorString(e));
        }
    return 0;
}
