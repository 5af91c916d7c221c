// fp64_peak.cu -- microbenchmark: sustained DFMA issue rate of the B200 FP64 pipe (the co-limit of
// the fused ERK kernel; MEASURED_PEAKS.json has no FP64 figure).  Prints warp-instructions/s/SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_peak tools/fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void dfma_kernel(double *out, int iters, double a, double b)
{
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 16; ++r)
#pragma unroll
            for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    double *out;
    cudaMalloc(&out, sizeof(double) * sms * 16 * 1024);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    printf("%s: %d SMs, clock %d kHz\n", p.name, sms, p.clockRate);
    for (int warps_per_sm : {4, 8, 16, 32, 64}) {
        int threads = 128, blocks = sms * warps_per_sm * 32 / threads;
        const int iters = 20000, ILP = 4;
        dfma_kernel<ILP><<<blocks, threads>>>(out, 100, 1.0000001, 1e-9);
        cudaDeviceSynchronize();
        float best = 1e30f;
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            dfma_kernel<ILP><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (ms < best) best = ms;
        }
        double fma = (double)blocks * threads * iters * 16.0 * ILP;
        printf("warps/SM %2d: %.2f ms  %.2f TFLOP/s (fp64 FMA=2)  %.1f DFMA lanes/clk/SM @1.965GHz-equivalent\n",
               warps_per_sm, best, 2 * fma / best / 1e9, fma / (best * 1e-3) / sms / 1.965e9);
    }
    return 0;
}
