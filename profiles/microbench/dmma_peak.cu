// Microbenchmark: fp64 tensor (mma.sync m8n8k4 f64) vs fp64 FMA throughput on B200, alone and mixed.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_peak dmma_peak.cu && ./dmma_peak
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int MODE>   // 0: DFMA only, 1: DMMA only, 2: both interleaved
__global__ void __launch_bounds__(256) k(double* out, int iters) {
    double a = 1.0 + threadIdx.x * 1e-9, b = 0.999;
    double c[8][2], f[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i][0] = i; c[i][1] = -i; f[i] = i + 0.5; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE != 0) dmma(c[i][0], c[i][1], a, b);
            if (MODE != 1) {
                f[i] = fma(f[i], b, a);
                if (MODE == 2) { f[i] = fma(f[i], b, a); f[i] = fma(f[i], b, a); f[i] = fma(f[i], b, a); }
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1] + f[i];
    if (s == 123.456) out[0] = s;
}

template <int MODE>
void run(const char* name, int nsm, int ctas, double flop_per_iter_thread) {
    double* d; cudaMalloc(&d, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0);
        k<MODE><<<nsm * ctas, 256>>>(d, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (r) best = ms < best ? ms : best;
    }
    double flops = flop_per_iter_thread * iters * 256.0 * nsm * ctas;
    printf("%-28s ctas/SM=%d  %.2f ms  %.2f TFLOP/s\n", name, ctas, best, flops / best / 1e9);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int nsm = p.multiProcessorCount;
    printf("%s, %d SMs\n", p.name, nsm);
    // per thread per iter: DFMA: 8 fma = 16 flop ; DMMA: 8 mma x (8*8*4*2 flop / 32 lanes = 16 flop/lane) = 128 flop
    for (int ctas : {1, 2, 4, 8}) {
        run<0>("DFMA only", nsm, ctas, 16.0);
        run<1>("DMMA m8n8k4 only", nsm, ctas, 128.0);
        run<2>("DMMA + 4x DFMA interleaved", nsm, ctas, 128.0 + 64.0);
    }
    return 0;
}
