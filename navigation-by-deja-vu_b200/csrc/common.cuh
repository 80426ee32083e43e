// common.cuh -- shared definitions for the navsim B200 engine (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#define NVB_MAX_BLOCK_PX 64
// "no candidate" markers.  Both stay positive when read as int64 so that a
// MIN all-reduce over ranks (view-sharded library) treats them as +infinity.
#define NVB_KEY_NONE 0x7FFFFFFFFFFFFFFFull
#define NVB_EXACT_NONE 0x7FF0000000000000ull  /* +inf as FP64 bits */  // landscape pixels per sensor pixel handled by the hue vote

// Device-side view of one world (landscape + sensor + library).  Passed by
// value to kernels.
struct NvbWorld {
    // landscape: 3 planes (H, S, V), each rows x pitch bytes
    const uint8_t *land;
    int rows, cols, pitch;
    long long plane_stride;
    // sensor
    int W, H, pw, ph;   // sensor pixels, landscape pixels per sensor pixel
    int P, Ppad;        // W*H and the row pitch of glimpse / library planes
    int Wpx, Hpx;
    int mask_lo, mask_hi;  // centre-column mask [lo, hi)
    double r;           // bounds-test radius, NavBySceneFamiliarity.py:94
    // staged window (TMA box) geometry; R = 0 when the window does not fit
    int R, BW, BH;
    const uint8_t *lut;  // [3][256]
};

__host__ __device__ inline int nvb_round_up(int a, int b) { return (a + b - 1) / b * b; }

// Python/NumPy float modulo with a positive divisor
// (NavBySceneFamiliarity.py:291,317): fmod, then fold negatives up.
__device__ __forceinline__ double nvb_pymod_pos(double a, double b)
{
    double m = fmod(a, b);
    if (m != 0.0) {
        if (m < 0.0) m = __dadd_rn(m, b);
    } else {
        m = 0.0;
    }
    return m;
}

// Four unsigned-byte absolute differences, summed and accumulated in ONE
// instruction (SASS: VABSDIFF4.U8.ACC).  The __vsadu4 intrinsic passes 0 as the
// accumulator and adds separately; the PTX form below fuses the accumulate.
__device__ __forceinline__ uint32_t nvb_sad4(uint32_t a, uint32_t b, uint32_t acc)
{
    uint32_t d;
    asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(acc));
    return d;
}

#define NVB_TWO_PI 6.283185307179586476925286766559
#define NVB_PI 3.14159265358979323846

// Programmatic dependent launch: blocks until the preceding kernel of the stream has
// completed and its writes are visible (a no-op for a normally launched kernel).  The
// kernels of the step sequence are launched with programmatic stream serialization,
// so their launch latency and prologue overlap the tail of the previous kernel.
__device__ __forceinline__ void nvb_grid_dep_wait()
{
    asm volatile("griddepcontrol.wait;" ::: "memory");
}
// Lets the NEXT kernel of the stream become resident as soon as every CTA of this one has
// got here (instead of when this grid has drained): its CTAs fill the SMs this grid leaves
// free, run their prologue -- which may only READ data no kernel of the step sequence
// writes (library, path, tables) -- and then block in nvb_grid_dep_wait().
__device__ __forceinline__ void nvb_grid_dep_launch()
{
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// Tuning aid (nvb_debug_timeline): per-CTA wall-clock stamps of the step kernels,
// tl[kernel][min(cta, NVB_TL_CTAS - 1)][3] = {resident, dependency met, done} in ns.
#define NVB_TL_CTAS 2048
__device__ __forceinline__ void nvb_tl_stamp(long long *tl, int kernel, int what)
{
    if (tl != nullptr && threadIdx.x == 0) {
        long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        const int cta = blockIdx.x < NVB_TL_CTAS ? blockIdx.x : NVB_TL_CTAS - 1;
        tl[((size_t)kernel * NVB_TL_CTAS + cta) * 3 + what] = t;
    }
}

// ---- mbarrier / TMA helpers (inline PTX; sm_100a) -------------------------
__device__ __forceinline__ uint32_t nvb_smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void nvb_mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(nvb_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void nvb_fence_barrier_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void nvb_mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(nvb_smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void nvb_mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(nvb_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void nvb_mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(nvb_smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
// 3-D tiled TMA load: box at (x, y, z) of the tensor map -> shared memory.
__device__ __forceinline__ void nvb_tma_load_3d(void *dst, const CUtensorMap *tmap, int x, int y,
                                                int z, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(nvb_smem_u32(dst)),
        "l"(tmap), "r"(x), "r"(y), "r"(z), "r"(nvb_smem_u32(bar))
        : "memory");
}
// 1-D bulk copy global -> shared (bytes multiple of 16, both 16-B aligned).
__device__ __forceinline__ void nvb_bulk_load_1d(void *dst, const void *src, uint32_t bytes,
                                                 uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(nvb_smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(nvb_smem_u32(bar))
        : "memory");
}
