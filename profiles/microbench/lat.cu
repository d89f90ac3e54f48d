// Latency microbenchmarks (single warp / single CTA per SM) for the ops on the critical path of the
// fused Gauss-Jordan step.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lat lat.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ double rcp2(double x) {
    double r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0); r = fma(r, e, r); e = fma(-x, r, 1.0); return fma(r, e, r);
}
template <int MODE>
__global__ void k(double* out, long long* cyc, int iters, double seed) {
    __shared__ double sh[256];
    double x = seed + threadIdx.x * 1e-12, y = 0.999999;
    sh[threadIdx.x] = x;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        if (MODE == 0) x = fma(x, y, 1e-9);                                   // dependent DFMA
        if (MODE == 1) { x = x * y; x = fma(x, y, 1e-9); }                    // DMUL -> DFMA
        if (MODE == 2) { asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(x)); }   // MUFU.RCP64H chain
        if (MODE == 3) x = rcp2(x);                                            // full reciprocal chain
        if (MODE == 4) x = 1.0 / x;                                            // compiler reciprocal
        if (MODE == 5) { sh[threadIdx.x] = x; x = sh[(threadIdx.x + 1) & 31] + 1e-9; }     // STS -> LDS -> DADD
        if (MODE == 6) { x = fma(x, y, 1e-9); __syncthreads(); }               // DFMA + bar.sync (CTA of blockDim)
        if (MODE == 7) { float f = (float)x; f = f * 0.999f; x = (double)f; }  // F2F round trip
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
    if (x == 123.456) out[0] = x;
}
template <int MODE> void run(const char* name, int threads, int opsPerIter) {
    double* d; long long* c; cudaMalloc(&d, 8); cudaMalloc(&c, 8);
    const int iters = 4096;
    k<MODE><<<1, threads>>>(d, c, iters, 1.0000001);
    k<MODE><<<1, threads>>>(d, c, iters, 1.0000001);
    long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    printf("%-34s threads=%3d  %7.1f cycles/iter (%d dependent ops)\n", name, threads, (double)h / iters, opsPerIter);
}
int main() {
    run<0>("DFMA dependent chain", 32, 1);
    run<1>("DMUL->DFMA", 32, 2);
    run<2>("rcp.approx.ftz.f64 chain", 32, 1);
    run<3>("rcp.approx + 2 Newton", 32, 5);
    run<4>("1.0/x (compiler)", 32, 1);
    run<5>("STS->LDS->DADD", 32, 3);
    run<6>("DFMA + bar.sync", 32, 1);
    run<6>("DFMA + bar.sync", 128, 1);
    run<6>("DFMA + bar.sync", 256, 1);
    run<7>("F2F f64->f32->f64 (+FMUL)", 32, 3);
    return 0;
}
