// Dependent-issue latency of fp64 operations on the device (one warp, clock64 around a chain):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false tools/micro/fp64_latency.cu -o /tmp/fp64_latency
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void chain(double *out, double x0, double a, double b, long long *cycles) {
    double x = x0 + threadIdx.x;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            if (OP == 0)
                x = fma(x, a, b);
            else if (OP == 1)
                x = x * a;
            else if (OP == 2)
                x = x + b;
            else if (OP == 3)
                x = b - x * a; // DMUL + DADD (no contraction: -fmad=false)
            else if (OP == 4)
                x = 1.0 / x + b; // reciprocal + DADD
            else if (OP == 5) {
                float f = (float)x;
                f = fmaf(f, 1.0001f, 0.5f);
                x = f;
            }
        }
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0)
        *cycles = t1 - t0;
}

int main() {
    double *out;
    long long *cyc, h;
    cudaMalloc(&out, 32 * 8);
    cudaMalloc(&cyc, 8);
    const char *names[] = {"DFMA", "DMUL", "DADD", "DMUL+DADD", "1/x + DADD", "F2F+FFMA+F2F"};
    for (int op = 0; op < 5; ++op) {
        for (int rep = 0; rep < 2; ++rep) {
            switch (op) {
            case 0: chain<0><<<1, 32>>>(out, 1.0, 0.999, 0.001, cyc); break;
            case 1: chain<1><<<1, 32>>>(out, 1.0, 1.0000001, 0.001, cyc); break;
            case 2: chain<2><<<1, 32>>>(out, 1.0, 0.999, 0.001, cyc); break;
            case 3: chain<3><<<1, 32>>>(out, 1.0, 0.5, 2.0, cyc); break;
            case 4: chain<4><<<1, 32>>>(out, 1.5, 0.5, 0.25, cyc); break;
            }
            cudaDeviceSynchronize();
        }
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("%-14s %7.1f cycles per dependent step (one warp)\n", names[op], h / 4096.0);
    }
    return 0;
}
