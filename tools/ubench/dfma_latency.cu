// Microbenchmark (tuning aid): cycles per DFMA for C independent dependent-chains per thread, W warps per SM sub-partition.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dfma_latency dfma_latency.cu && ./dfma_latency
#include <cstdio>
#include <cuda_runtime.h>

template <int C>
__global__ void chain(double* out, long long* cyc, int iters, double b, double c) {
    double a[C];
#pragma unroll
    for (int i = 0; i < C; ++i) a[i] = threadIdx.x * 1e-3 + i;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < C; ++i) a[i] = fma(a[i], b, c);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < C; ++i) s += a[i];
    if (s == 12345.678) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int C>
void run(int warps_per_block) {
    double* out;
    long long* cyc;
    cudaMalloc(&out, 8);
    cudaMalloc(&cyc, 8);
    const int iters = 2000;
    chain<C><<<148, warps_per_block * 32>>>(out, cyc, iters, 0.999999, 1e-7);
    chain<C><<<148, warps_per_block * 32>>>(out, cyc, iters, 0.999999, 1e-7);
    long long h;
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double per = (double)h / (iters * 8.0 * C);
    printf("chains/thread %d  warps/SM %2d (per scheduler %.1f): %.2f cycles per DFMA per warp, %.2f DFMA/clk/SM-subpartition\n", C,
           warps_per_block, warps_per_block / 4.0, per, (warps_per_block / 4.0) / per);
    cudaFree(out);
    cudaFree(cyc);
}

int main() {
    for (int w : {4, 8, 16}) {
        run<1>(w); run<2>(w); run<3>(w); run<4>(w); run<6>(w); run<8>(w);
    }
    return 0;
}
