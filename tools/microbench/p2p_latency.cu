// p2p_latency.cu — measures the producer->consumer hop latency through L2 between warps on different SMs on B200,
// for the store/load flavours the LU-SGS level pipeline can use.  Chain of NB blocks: block b waits for block b-1's
// value (32 lanes each their own double + optional hint flag), adds 1, publishes.  Time per hop = total / NB.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ double ldRelaxed(const double* p) { double v; asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ double ldVolatile(const double* p) { double v; asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ int ldRelaxedI(const int* p) { int v; asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ int ldAcquireI(const int* p) { int v; asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void stReleaseI(int* p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void stRelaxed(double* p, double v) { asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory"); }

constexpr unsigned long long SENT = 0xFFF8DEADBEEF0B1DULL;

// mode 0: sentinel data only (poll data)   1: hint flag (relaxed) + sentinel data   2: release/acquire flag + plain data
// mode 3: sentinel data, volatile loads    4: like 0 but producer uses st.relaxed.gpu
template <int MODE>
__global__ void chain(double* data, int* flag, int nHops, int extraLoads, const double* junk, double* sink, long long* cycles)
{
    const int b = blockIdx.x, lane = threadIdx.x;
    long long t0 = clock64();
    double acc = 0;
    for (int h = b; h < nHops; h += gridDim.x) {
        double v = 0.0;
        if (h > 0) {
            const double* src = data + (size_t)(h - 1) * 32 + lane;
            if (MODE == 0 || MODE == 4) { do { v = ldRelaxed(src); } while ((unsigned long long)__double_as_longlong(v) == SENT); }
            else if (MODE == 3) { do { v = ldVolatile(src); } while ((unsigned long long)__double_as_longlong(v) == SENT); }
            else if (MODE == 1) {
                while (true) {
                    if (ldRelaxedI(flag + h - 1) == 1) { v = ldRelaxed(src); if ((unsigned long long)__double_as_longlong(v) != SENT) break; }
                }
            } else {
                while (ldAcquireI(flag + h - 1) != 1) { }
                v = __ldcg(src);
            }
        }
        // optional bulk traffic in flight on the same warp (like the 5x5 block prefetch)
        for (int e = 0; e < extraLoads; e++) acc += __ldcs(junk + ((size_t)h * extraLoads + e) * 32 + lane);
        v = v + 1.0;
        double* dst = data + (size_t)h * 32 + lane;
        if (MODE == 4) stRelaxed(dst, v); else __stcg(dst, v);
        if (MODE == 1) { __syncwarp(); if (lane == 0) __stcg(flag + h, 1); }
        if (MODE == 2) { __syncwarp(); if (lane == 0) stReleaseI(flag + h, 1); }
    }
    if (acc == 12345.678) sink[0] = acc;
    if (lane == 0 && b == (nHops - 1) % gridDim.x) cycles[0] = clock64() - t0;
}

template <int MODE>
void run(const char* name, int grid, int nHops, int extra, double* junk)
{
    double *data, *sink; int* flag; long long* cyc;
    cudaMalloc(&data, sizeof(double) * 32 * nHops); cudaMalloc(&flag, sizeof(int) * nHops); cudaMalloc(&sink, 8); cudaMalloc(&cyc, 8);
    for (int rep = 0; rep < 3; rep++) {
        unsigned long long* hs = new unsigned long long[32 * (size_t)nHops];
        for (size_t i = 0; i < 32 * (size_t)nHops; i++) hs[i] = SENT;
        cudaMemcpy(data, hs, sizeof(double) * 32 * nHops, cudaMemcpyHostToDevice);
        delete[] hs;
        cudaMemset(flag, 0, sizeof(int) * nHops);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        void* args[] = {&data, &flag, &nHops, &extra, &junk, &sink, &cyc};
        cudaLaunchCooperativeKernel((void*)chain<MODE>, dim3(grid), dim3(32), args, 0, 0);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double last; cudaMemcpy(&last, data + (size_t)(nHops - 1) * 32, 8, cudaMemcpyDeviceToHost);
        if (rep == 2) printf("%-34s grid=%4d extra=%3d hops=%d  %.3f us/hop  (check %.0f) err=%s\n", name, grid, extra, nHops, 1e3 * ms / nHops, last, cudaGetErrorString(cudaGetLastError()));
    }
    cudaFree(data); cudaFree(flag); cudaFree(sink); cudaFree(cyc);
}

int main()
{
    double* junk; cudaMalloc(&junk, sizeof(double) * 32 * 20000ull * 80); cudaMemset(junk, 0, sizeof(double) * 32 * 20000ull * 80);
    const int H = 20000;
    for (int grid : {2, 148, 592}) {
        run<0>("sentinel ld.relaxed/st.cg", grid, H, 0, junk);
        run<4>("sentinel ld.relaxed/st.relaxed", grid, H, 0, junk);
        run<3>("sentinel ld.volatile/st.cg", grid, H, 0, junk);
        run<1>("hint flag + sentinel", grid, H, 0, junk);
        run<2>("release/acquire flag", grid, H, 0, junk);
    }
    for (int extra : {25, 75}) {
        run<0>("sentinel + bulk loads", 148, H, extra, junk);
        run<1>("hint+sentinel + bulk loads", 148, H, extra, junk);
        run<2>("rel/acq + bulk loads", 148, H, extra, junk);
    }
    return 0;
}
