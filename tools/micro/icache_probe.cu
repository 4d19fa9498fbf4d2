// Micro-benchmark: cycles per instruction of ONE warp per SM running the same number of independent FFMAs as
// (a) a long straight-line body (code far larger than the instruction caches), (b) a compact loop.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o icache_probe icache_probe.cu ; run: ./icache_probe
#include <cstdio>
#include <cuda_runtime.h>

template <int BODY, int OUTER>
__global__ void probe(float* out, long long* cyc, float a, float b) {
    float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    long long best = 1ll << 60;
    for (int rep = 0; rep < 6; ++rep) {
        const long long t0 = clock64();
#pragma unroll 1
        for (int o = 0; o < OUTER; ++o) {
#pragma unroll
            for (int i = 0; i < BODY; ++i) {
                x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
                x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
            }
        }
        const long long t1 = clock64();
        if (t1 - t0 < best) best = t1 - t0;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
    if (threadIdx.x == 0) cyc[blockIdx.x] = best;
}

template <int BODY, int OUTER>
void run(const char* name, int grid, int threads) {
    float* out; long long* cyc;
    cudaMalloc(&out, grid * threads * 4); cudaMalloc(&cyc, grid * 8);
    probe<BODY, OUTER><<<grid, threads>>>(out, cyc, 1.0001f, 0.5f);
    cudaDeviceSynchronize();
    long long h[256]; cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost);
    double m = 0; for (int i = 0; i < grid; ++i) m += h[i]; m /= grid;
    const double n = 8.0 * BODY * OUTER;
    printf("%-34s grid %3d threads %3d  static %6d instr (%4d KB)  dynamic %6.0f  cycles %8.0f  cyc/instr %.2f\n", name, grid, threads, 8 * BODY,
           8 * BODY * 16 / 1024, n, m, m / n);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int threads : {32, 128, 384}) {
        run<8, 512>("loop body 64 instr (1 KB)", 16, threads);
        run<64, 64>("loop body 512 instr (8 KB)", 16, threads);
        run<192, 21>("loop body 1536 instr (24 KB)", 16, threads);
        run<256, 16>("loop body 2048 instr (32 KB)", 16, threads);
        run<512, 8>("loop body 4096 instr (64 KB)", 16, threads);
        run<1024, 4>("loop body 8192 instr (128 KB)", 16, threads);
        run<512, 8>("64 KB body, grid 148", 148, threads);
    }
    return 0;
}
