// stream_probe.cu -- what does reading an 80 MB library ONCE cost on this GPU, by access path?
// (roofline denominator of the C4 one-agent configuration: MEASURED_PEAKS.json's HBM figure is
// a long-running copy; a 12-us pass also pays ramp-up and tail)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -o tools/micro/stream_probe tools/micro/stream_probe.cu
//   tools/micro/stream_probe [MB]
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../navigation-by-deja-vu_b200/csrc/distance.cuh"

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s (%d)\n", #call, cudaGetErrorString(e_), __LINE__); exit(2); } } while (0)

// plain coalesced 16-byte loads, UNR in flight per thread
template <int UNR>
__global__ void k_read_ldg(const uint4 *p, long long n16, unsigned *sink)
{
    unsigned acc = 0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (UNR - 1) * stride < n16; i += UNR * stride) {
        uint4 v[UNR];
#pragma unroll
        for (int u = 0; u < UNR; u++) v[u] = __ldg(p + i + u * stride);
#pragma unroll
        for (int u = 0; u < UNR; u++) acc ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
    }
    for (; i < n16; i += stride) { const uint4 v = __ldg(p + i); acc ^= v.x ^ v.y ^ v.z ^ v.w; }
    if (acc == 0x12345678u) *sink = acc;
}

// per-warp bulk copies (the k2_stream load path) with almost no compute: one LDS.128 per lane and chunk
template <int ST, int CHUNK>
__global__ void k_read_bulk(const uint8_t *p, long long bytes, unsigned *sink)
{
    extern __shared__ __align__(128) uint8_t sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, NW = blockDim.x >> 5;
    uint8_t *ws = sm + (size_t)warp * ST * CHUNK;
    uint64_t *bar = reinterpret_cast<uint64_t *>(sm + (size_t)NW * ST * CHUNK) + warp * ST;
    const long long n_chunks_all = bytes / CHUNK, n_warps = (long long)gridDim.x * NW, gw = (long long)blockIdx.x * NW + warp;
    const long long c0 = n_chunks_all * gw / n_warps, c1 = n_chunks_all * (gw + 1) / n_warps;
    const int n = (int)(c1 - c0);
    auto issue = [&](int ch) {
        nvb_mbar_expect_tx(bar + ch % ST, CHUNK);
        nvb_bulk_load_1d(ws + (size_t)(ch % ST) * CHUNK, p + (size_t)(c0 + ch) * CHUNK, CHUNK, bar + ch % ST);
    };
    if (lane == 0) {
        for (int s = 0; s < ST; s++) nvb_mbar_init(bar + s, 1);
        nvb_fence_barrier_init();
        for (int s = 0; s < ST && s < n; s++) issue(s);
    }
    __syncwarp();
    unsigned acc = 0;
    for (int ch = 0; ch < n; ch++) {
        nvb_mbar_wait(bar + ch % ST, (unsigned)((ch / ST) & 1));
        const uint4 v = *reinterpret_cast<const uint4 *>(ws + (size_t)(ch % ST) * CHUNK + lane * 16);
        acc ^= v.x ^ v.y ^ v.z ^ v.w;
        asm volatile("" ::"r"(acc) : "memory");
        __syncwarp();
        if (lane == 0 && ch + ST < n) issue(ch + ST);
    }
    if (acc == 0x12345678u) *sink = acc;
}

static float time_it(void (*launch)(void *), void *ctx, int reps, void *flush, size_t flush_bytes, bool cold)
{
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float total = 0;
    for (int r = 0; r < reps + 2; r++) {
        if (cold) CK(cudaMemsetAsync(flush, r, flush_bytes));
        CK(cudaEventRecord(e0));
        launch(ctx);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        CK(cudaGetLastError());
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (r >= 2) total += ms;
    }
    return total / reps * 1e3f;
}

struct Ctx { const uint8_t *p; long long bytes; unsigned *sink; int sms; DistArgs da; };

int main(int argc, char **argv)
{
    const long long MB = argc > 1 ? atoll(argv[1]) : 80;
    const long long bytes = MB * 1000000LL / 5120 * 5120;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    Ctx c;
    c.bytes = bytes; c.sms = prop.multiProcessorCount;
    uint8_t *d;
    CK(cudaMalloc(&d, bytes));
    CK(cudaMemset(d, 3, bytes));
    c.p = d;
    CK(cudaMalloc(&c.sink, 4));
    void *flush;
    const size_t flush_bytes = 512u << 20;
    CK(cudaMalloc(&flush, flush_bytes));
    const int reps = 20;
    auto report = [&](const char *name, float us_cold, float us_warm) {
        printf("{\"path\": \"%s\", \"MB\": %.1f, \"us_cold_l2\": %.2f, \"GBs_cold\": %.0f, \"us_back_to_back\": %.2f, \"GBs_back_to_back\": %.0f}\n",
               name, bytes / 1e6, us_cold, bytes / (us_cold * 1e-6) / 1e9, us_warm, bytes / (us_warm * 1e-6) / 1e9);
        fflush(stdout);
    };
#define RUN(name, body)                                                                        \
    {                                                                                          \
        auto fn = [](void *vc) { Ctx &c = *(Ctx *)vc; body; };                                 \
        report(name, time_it(fn, &c, reps, flush, flush_bytes, true), time_it(fn, &c, reps, flush, flush_bytes, false)); \
    }
    RUN("ldg x4, 148x4 CTAs of 256", (k_read_ldg<4><<<c.sms * 4, 256>>>((const uint4 *)c.p, c.bytes / 16, c.sink)));
    RUN("ldg x8, 148x4 CTAs of 256", (k_read_ldg<8><<<c.sms * 4, 256>>>((const uint4 *)c.p, c.bytes / 16, c.sink)));
    RUN("ldg x8, 148x8 CTAs of 256", (k_read_ldg<8><<<c.sms * 8, 256>>>((const uint4 *)c.p, c.bytes / 16, c.sink)));
    {
        auto k = k_read_bulk<3, 5120>;
        CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 3 * 5120 + 256));
        RUN("bulk 5 KB x 3 stages per warp, 148x3 CTAs of 128", (k_read_bulk<3, 5120><<<c.sms * 3, 128, 4 * 3 * 5120 + 256>>>(c.p, c.bytes, c.sink)));
    }
    {
        auto k = k_read_bulk<4, 10240>;
        CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 4 * 10240 + 256));
        RUN("bulk 10 KB x 4 stages per warp, 148 CTAs of 128", (k_read_bulk<4, 10240><<<c.sms, 128, 4 * 4 * 10240 + 256>>>(c.p, c.bytes, c.sink)));
    }
    {
        auto k = k_read_bulk<2, 20480>;
        CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 2 * 20480 + 256));
        RUN("bulk 20 KB x 2 stages per warp, 148 CTAs of 128", (k_read_bulk<2, 20480><<<c.sms, 128, 4 * 2 * 20480 + 256>>>(c.p, c.bytes, c.sink)));
    }
    // the real kernel: G = 10 and G = 1 glimpses (the difference is its arithmetic)
    {
        unsigned long long *keys;
        uint8_t *g;
        CK(cudaMalloc(&keys, 16 * 8));
        CK(cudaMemset(keys, 0x7F, 16 * 8));
        CK(cudaMalloc(&g, 16 * 80));
        CK(cudaMemset(g, 7, 16 * 80));
        DistArgs a{};
        a.gv = g; a.lv = d; a.N = (int)(bytes / 80); a.Ppad = 80; a.nk = 1; a.keys = keys; a.idx_bits = 32;
        c.da = a;
        const int smem = nvb_stream_smem(80, 10);
        CK(cudaFuncSetAttribute(k2_stream<5, 10, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        CK(cudaFuncSetAttribute(k2_stream<5, 10, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        c.da.G = 10;
#define EX true
        RUN("k2_stream G=10 exact", (k2_stream<5, 10, EX><<<c.sms * 2, NVB_STREAM_THREADS, nvb_stream_smem(80, 10)>>>(c.da)));
#undef EX
#define EX false
        RUN("k2_stream G=10 predicated", (k2_stream<5, 10, EX><<<c.sms * 2, NVB_STREAM_THREADS, nvb_stream_smem(80, 10)>>>(c.da)));
        c.da.G = 1;
        RUN("k2_stream G=1", (k2_stream<5, 10, EX><<<c.sms * 2, NVB_STREAM_THREADS, nvb_stream_smem(80, 10)>>>(c.da)));
    }
    return 0;
}
