// Micro-benchmarks behind DESIGN.md's latency notes: dependent FP64 add chain, FP64 add
// throughput, L2 / L1 load-to-use latency (pointer chase).   nvcc -arch=sm_100a lat.cu -o lat
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dadd_chain(double *out, double x, int n, long long *cyc)
{
    double s = out[threadIdx.x];
    long long t0 = clock64();
    for (int i = 0; i < n; i++) s = __dadd_rn(s, x);
    long long t1 = clock64();
    out[threadIdx.x] = s;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void fadd_chain(float *out, float x, int n, long long *cyc)
{
    float s = out[threadIdx.x];
    long long t0 = clock64();
    for (int i = 0; i < n; i++) s = __fadd_rn(s, x);
    long long t1 = clock64();
    out[threadIdx.x] = s;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void dadd_tput(double *out, double x, int n, long long *cyc)
{
    double s[8];
    for (int k = 0; k < 8; k++) s[k] = out[threadIdx.x + k];
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < n; i++)
#pragma unroll
        for (int k = 0; k < 8; k++) s[k] = __dadd_rn(s[k], x);
    __syncthreads();
    long long t1 = clock64();
    double r = 0;
    for (int k = 0; k < 8; k++) r += s[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void chase(const int *next, int n, int start, int *out, long long *cyc, int cached)
{
    int p = start;
    long long t0 = clock64();
    for (int i = 0; i < n; i++) p = cached ? __ldg(next + p) : __ldcg(next + p);
    long long t1 = clock64();
    *out = p;
    *cyc = t1 - t0;
}
int main()
{
    double *d; float *f; long long *c; int *nx, *o;
    cudaMalloc(&d, 1 << 20); cudaMalloc(&f, 1 << 20); cudaMalloc(&c, 8); cudaMemset(d, 0, 1 << 20); cudaMemset(f, 0, 1 << 20);
    long long h;
    const int n = 4096;
    for (int rep = 0; rep < 2; rep++) {
        dadd_chain<<<1, 32>>>(d, 1e-3, n, c); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
        if (rep) printf("dependent DADD: %.1f cycles / op\n", (double)h / n);
        fadd_chain<<<1, 32>>>(f, 1e-3f, n, c); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
        if (rep) printf("dependent FADD: %.1f cycles / op\n", (double)h / n);
        for (int warps = 4; warps <= 32; warps *= 2) {
            dadd_tput<<<148, warps * 32>>>(d, 1e-3, 512, c); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
            if (rep) printf("DADD throughput, %2d warps/SM x 8 independent: %.2f lanes / clk / SM\n", warps,
                            512.0 * 8 * warps * 32 / h);
        }
    }
    // pointer chase over 32 MiB (L2 resident, > L1) and over 16 KiB (L1 resident)
    for (int kb : {16, 32 * 1024}) {
        const int m = kb * 1024 / 4, stride = 4099;
        int *hn = new int[m];
        for (int i = 0; i < m; i++) hn[i] = (int)(((long long)i + stride * 32LL) % m);
        cudaMalloc(&nx, (size_t)m * 4); cudaMalloc(&o, 4);
        cudaMemcpy(nx, hn, (size_t)m * 4, cudaMemcpyHostToDevice);
        for (int cached = 0; cached < 2; cached++)
            for (int rep = 0; rep < 2; rep++) {
                chase<<<1, 1>>>(nx, 2000, 0, o, c, cached); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
                if (rep) printf("pointer chase, %5d KiB, %s: %.0f cycles / load\n", kb, cached ? "ld.global.nc (L1)" : "ld.global.cg (L2)", (double)h / 2000);
            }
        cudaFree(nx); cudaFree(o); delete[] hn;
    }
    return 0;
}
