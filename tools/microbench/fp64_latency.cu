// fp64_latency.cu — dependent-issue latency and throughput of the fp64 pipe on B200 (DADD, DMUL, DFMA, IEEE division,
// sqrt), the numbers behind the "fp64-bound" statements of DESIGN.md §4 (face kernels, LU-SGS dependent path).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false fp64_latency.cu -o fp64_latency
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void chainLatency(double* out, double a0, double b, int n, long long* cycles)
{
    double a = a0 + threadIdx.x;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; i++) {
        if (OP == 0) a = a + b;
        else if (OP == 1) a = a * b;
        else if (OP == 2) a = __fma_rn(a, b, b);
        else if (OP == 3) a = b / a + 1.0;
        else a = sqrt(a) + b;
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = a;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

// 8 independent chains per thread, many warps: throughput
template <int OP>
__global__ void throughput(double* out, double a0, double b, int n)
{
    double a[8];
#pragma unroll
    for (int k = 0; k < 8; k++) a[k] = a0 + threadIdx.x + k;
    for (int i = 0; i < n; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if (OP == 0) a[k] = a[k] + b;
            else if (OP == 1) a[k] = a[k] * b;
            else if (OP == 2) a[k] = __fma_rn(a[k], b, b);
            else if (OP == 3) a[k] = b / a[k] + 1.0;
            else a[k] = sqrt(a[k]) + b;
        }
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s += a[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP>
void run(const char* name, double* d_out, long long* d_cyc, int nSM, double clockGHz)
{
    const int n = 4096;
    chainLatency<OP><<<1, 32>>>(d_out, 1.5, 1.0000001, n, d_cyc);
    cudaDeviceSynchronize();
    long long cyc = 0;
    cudaMemcpy(&cyc, d_cyc, sizeof(cyc), cudaMemcpyDeviceToHost);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = nSM * 8, threads = 256, iters = 2000;
    throughput<OP><<<blocks, threads>>>(d_out, 1.5, 1.0000001, 10);
    cudaEventRecord(e0);
    throughput<OP><<<blocks, threads>>>(d_out, 1.5, 1.0000001, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double ops = (double)blocks * threads * 8.0 * iters;
    printf("%-28s dependent latency %6.1f cycles/op   throughput %8.2f Gop/s  (%.1f lanes/clk/SM at %.2f GHz)\n", name, (double)cyc / n, ops / (ms * 1e6),
           ops / (ms * 1e-3) / (nSM * clockGHz * 1e9), clockGHz);
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double ghz = clk * 1e-6;
    printf("%s, %d SMs, clock %.3f GHz (nominal; ops per result below: add/mul/fma = 1, div = div+add, sqrt = sqrt+add)\n", p.name, p.multiProcessorCount, ghz);
    double* d_out; long long* d_cyc;
    cudaMalloc(&d_out, sizeof(double) * p.multiProcessorCount * 8 * 256);
    cudaMalloc(&d_cyc, sizeof(long long));
    run<0>("DADD", d_out, d_cyc, p.multiProcessorCount, ghz);
    run<1>("DMUL", d_out, d_cyc, p.multiProcessorCount, ghz);
    run<2>("DFMA", d_out, d_cyc, p.multiProcessorCount, ghz);
    run<3>("IEEE division (+ add)", d_out, d_cyc, p.multiProcessorCount, ghz);
    run<4>("IEEE sqrt (+ add)", d_out, d_cyc, p.multiProcessorCount, ghz);
    return 0;
}
