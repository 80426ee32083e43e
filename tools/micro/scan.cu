// How long does the update_error path scan take with 1024 CTAs resident (7 per SM)?
// Variants: FP64 as in step.cuh, FP64 without the coverage store, FP32 prefilter.
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>
#include <cmath>
template <int VARIANT>
__global__ void __launch_bounds__(160, 7) scan(const double2 *path, const float2 *pathf, int n_path, const double *pos,
                                               unsigned char *cover, double *out, double thr2, int reps)
{
    const int b = blockIdx.x, tid = threadIdx.x;
    __shared__ double s_red[8];
    double acc = 0;
    for (int r = 0; r < reps; r++) {
        const double x = pos[2 * b] + r, y = pos[2 * b + 1];
        double m = 1e300;
        if (VARIANT < 2) {
            if (tid < 128)
#pragma unroll 4
            for (int n = tid; n < n_path; n += 128) {
                const double2 pt = __ldg(path + n);
                const double dx = __dsub_rn(pt.x, x), dy = __dsub_rn(pt.y, y);
                const double d2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
                m = fmin(m, d2);
                if (VARIANT == 0 && d2 <= thr2) cover[(size_t)b * n_path + n] = 1;
            }
        } else {
            const float xf = (float)x, yf = (float)y, thr2f = (float)thr2 * 1.001f;
            float mf = 1e30f;
            if (tid < 128)
#pragma unroll 4
            for (int n = tid; n < n_path; n += 128) {
                const float2 pt = __ldg(pathf + n);
                const float dx = pt.x - xf, dy = pt.y - yf;
                const float d2 = fmaf(dx, dx, dy * dy);
                mf = fminf(mf, d2);
                if (d2 <= thr2f) {   // rare: exact decision
                    const double2 pd = __ldg(path + n);
                    const double ex = __dsub_rn(pd.x, x), ey = __dsub_rn(pd.y, y);
                    if (__dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey)) <= thr2) cover[(size_t)b * n_path + n] = 1;
                }
            }
            m = mf;
        }
        for (int o = 16; o > 0; o >>= 1) m = fmin(m, __shfl_xor_sync(0xFFFFFFFFu, m, o));
        if ((tid & 31) == 0) s_red[tid >> 5] = m;
        __syncthreads();
        m = s_red[0];
        for (int w = 1; w < 5; w++) m = fmin(m, s_red[w]);
        acc += m;
        __syncthreads();
    }
    if (tid == 0) out[b] = acc;
}
int main()
{
    const int n_path = 1414, B = 1024, reps = 50;
    std::vector<double2> hp(n_path); std::vector<float2> hf(n_path); std::vector<double> pos(2 * B);
    for (int i = 0; i < n_path; i++) { hp[i] = make_double2(300 + i, 1000 + 400 * sin(i * 0.005)); hf[i] = make_float2((float)hp[i].x, (float)hp[i].y); }
    for (int b = 0; b < B; b++) { pos[2 * b] = 300 + (b % 700); pos[2 * b + 1] = 1000 + (b % 37); }
    double2 *dp; float2 *df; double *dpos, *dout; unsigned char *cov;
    cudaMalloc(&dp, n_path * 16); cudaMalloc(&df, n_path * 8); cudaMalloc(&dpos, B * 16); cudaMalloc(&dout, B * 8); cudaMalloc(&cov, (size_t)B * n_path);
    cudaMemcpy(dp, hp.data(), n_path * 16, cudaMemcpyHostToDevice); cudaMemcpy(df, hf.data(), n_path * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(dpos, pos.data(), B * 16, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const char *names[3] = {"FP64 scan + coverage store", "FP64 scan only", "FP32 prefilter + exact near threshold"};
    for (int v = 0; v < 3; v++)
        for (int rep = 0; rep < 2; rep++) {
            cudaEventRecord(e0);
            if (v == 0) scan<0><<<B, 160>>>(dp, df, n_path, dpos, cov, dout, 196.0, reps);
            if (v == 1) scan<1><<<B, 160>>>(dp, df, n_path, dpos, cov, dout, 196.0, reps);
            if (v == 2) scan<2><<<B, 160>>>(dp, df, n_path, dpos, cov, dout, 196.0, reps);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep) printf("%-40s %.2f us per scan of %d points (1024 CTAs x 128 scanning threads)\n", names[v], ms * 1e3 / reps, n_path);
        }
    return 0;
}
