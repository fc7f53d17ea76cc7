#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void chain(double *out, long long *cyc, double a, double b, int iters) {
    double x = a + threadIdx.x * 1e-9, y = b;
    float xf = (float)x, yf = (float)b;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
        #pragma unroll
        for (int k = 0; k < 16; k++) {
            if (OP == 0) x = __fma_rn(x, y, y);
            if (OP == 1) x = __dmul_rn(x, y);
            if (OP == 2) x = __dadd_rn(x, y);
            if (OP == 3) x = __ddiv_rn(y, x);
            if (OP == 4) xf = __fmaf_rn(xf, yf, yf);
            if (OP == 5) x = floor(x) + y;
            if (OP == 6) x = (double)(int)(x) + y;
        }
    }
    long long t1 = clock64();
    if (OP == 4) x = xf;
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int OP> void run(const char *name, double *out, long long *cyc) {
    int iters = 256;
    for (int threads : {32, 128, 256, 512, 1024}) {
        chain<OP><<<1, threads>>>(out, cyc, 1.0000001, 0.9999999, iters);
        cudaDeviceSynchronize();
        chain<OP><<<1, threads>>>(out, cyc, 1.0000001, 0.9999999, iters);
        cudaDeviceSynchronize();
        long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        printf("%-10s threads/SM %4d: %.2f cycles per op per warp-chain; %.2f warp-ops/cycle/SM\n", name, threads, (double)c / (iters * 16), (double)(threads / 32) * iters * 16 / c);
    }
}
int main() {
    double *out; long long *cyc;
    cudaMalloc(&out, 8 * 4096); cudaMalloc(&cyc, 8);
    run<0>("DFMA", out, cyc); run<1>("DMUL", out, cyc); run<2>("DADD", out, cyc); run<3>("DDIV", out, cyc); run<4>("FFMA", out, cyc);
    run<5>("floor+add", out, cyc); run<6>("d2i2d+add", out, cyc);
    return 0;
}
