// distance_tc.cuh -- K2 on the 5th-generation tensor cores (tcgen05, sm_100a): the
// glimpse-vs-library sum of absolute differences as an EXACT int8 contraction.
//
// Replaces the same reference lines as distance.cuh (sads_hsv_metric,
// navsim/util.pyx:28-73, with chem_weight == 0, plus the per-heading max of
// navsim/NavBySceneFamiliarity.py:313) for sensors quantised to few levels
// (n_sensor_levels, NavBySceneFamiliarity.py:100-104,175-186: default 5).
//
// After quantisation a V pixel takes one of n levels l_0 < l_1 < ... < l_{n-1}
// (the distinct values of the quantisation table).  With the thermometer code
// T_k(v) = [v > l_k], k = 0..n-2, and weights w_k = l_{k+1} - l_k,
//     |a - b| = sum_k w_k [T_k(a) != T_k(b)] = sum_k w_k (1 - s_k(a) s_k(b)) / 2,
// s_k = 2 T_k - 1 in {-1, +1}.  Summed over the P sensor pixels:
//     SAD(a, b) = (C - dot(A, B)) / 2,   C = P * sum_k w_k = P * (l_{n-1} - l_0),
// A[k, p] = w_k s_k(a_p) (int8, |w_k| <= 127: larger weights are split into several
// planes), B[k, p] = s_k(b_p) (int8, +-1).  dot is an int8 x int8 -> int32 GEMM of
// shape [G x K] . [K x N], K = planes * P, exact in 32-bit integers: the smallest SAD is
// the largest dot, and the integer SAD this kernel reports is bit-identical to the
// byte-SIMD kernel's (k2_sad_v).
//
// Structure (one persistent CTA per SM, 192 threads, warp-specialised):
//   warp 0      TMA producer: 128 x KCH-byte glimpse chunk + NT x KCH-byte view chunk per
//               pipeline stage (cp.async.bulk.tensor.2d, 64- or 128-byte swizzle)
//   warp 1      MMA issuer: tcgen05.mma.cta_group::1.kind::i8, M = 128, N = NT, K = 32 per
//               instruction, accumulator in TMEM (2 buffers of 256 columns)
//   warps 2..5  epilogue: tcgen05.ld of the finished accumulator (one glimpse row per
//               thread), key = -256 * dot + column folded with a 3-input minimum, one
//               64-bit atomicMin per glimpse row and work item
// Work items (glimpse tile x view tile) are cut into one contiguous span per CTA exactly
// like k2_sad_v's units.
#pragma once
#include "common.cuh"

#define NVB_TC_THREADS 192
#define NVB_TC_TM 128          /* glimpse rows per tile (UMMA M) */
#define NVB_TC_MAX_PLANES 8
#define NVB_TC_NT 256          /* views per tile (UMMA N) */

struct TcPlanes {
    int n_planes;                          // thermometer planes (after splitting heavy weights)
    int8_t weight[NVB_TC_MAX_PLANES];      // w_k, 1..127
    uint8_t thr_level[NVB_TC_MAX_PLANES];  // plane k is set where level index > thr_level[k]
};

struct TcArgs {
    int G, N;                    // glimpses, views (local shard)
    int n_vt;                    // view tiles
    int n_gt;                    // glimpse tiles
    int vt_major;                // item u = vt * n_gt + gt (consecutive items share a view tile) instead of gt * n_vt + vt
    int kchunks;                 // K-chunks of KCH bytes per row
    const int *spans;            // [gridDim.x + 1] item boundaries per CTA
    unsigned long long *keys;    // [G], pre-set to NVB_KEY_NONE
    long long view_offset;       // global index of local view 0
    int sad_const;               // C = P * sum of plane weights
    int *step_counter;           // resident loop: bumped once per launch, else nullptr
    int *tie_count;
    int2 *tmin;                  // TILEMIN kernels: [G][n_vt] the two smallest tile-local keys (-256 * dot + column;
                                 // 0x7FFFFFFF = none) of every glimpse and view tile, for the single-launch step
                                 // (step_tm.cuh, k3_step_tm); the packed keys are then not written at all
    unsigned long long *epoch;   // view shards over NVLink: launches so far (step.cuh), else nullptr
    int pdl_early;
    long long *tl;
};


// Tuning aids: per-item stamps of the view-tile-stationary kernel.  NVB_TC_EXP_STAMPS (tools/micro/tc_sad.cu):
// CTA 0, SM cycles; NVB_TC_SITU_STAMPS (tools/k2_situ.py builds a second library with it): every CTA, global
// timer, into the unused decide rows of the engine's timeline buffer, to line up with the step kernels' stamps.
#if defined(NVB_TC_SITU_STAMPS)
#define NVB_TC_STAMP(it_, x_)                                                                          \
    do {                                                                                               \
        if (a.tl != nullptr && (it_) < 4) {                                                            \
            long long t_;                                                                              \
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                                     \
            a.tl[6144 + blockIdx.x * 32 + (it_) * 8 + (x_)] = t_;                                      \
        }                                                                                              \
    } while (0)
#elif defined(NVB_TC_EXP_STAMPS)
#define NVB_TC_STAMP(it_, x_)                                                                          \
    do {                                                                                               \
        if (a.tl != nullptr && blockIdx.x == 0 && (it_) < 60) a.tl[20000 + (it_) * 8 + (x_)] = clock64(); \
    } while (0)
#else
#define NVB_TC_STAMP(it_, x_) do { } while (0)
#endif

// ---- PTX wrappers -----------------------------------------------------------------
__device__ __forceinline__ void nvb_tma_load_2d(void *dst, const CUtensorMap *tmap, int x, int y, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::
            "r"(nvb_smem_u32(dst)),
        "l"(tmap), "r"(x), "r"(y), "r"(nvb_smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void nvb_prefetch_tmap(const CUtensorMap *tmap)
{
    asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
__device__ __forceinline__ void nvb_tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void nvb_tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void nvb_tmem_alloc(uint32_t *slot, uint32_t ncols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(nvb_smem_u32(slot)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void nvb_tmem_dealloc(uint32_t taddr, uint32_t ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem] . B[smem], int8 x int8 -> int32, issued by ONE thread
__device__ __forceinline__ void nvb_umma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// the MMAs issued so far by this thread arrive on `bar` when they have completed
__device__ __forceinline__ void nvb_umma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(nvb_smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void nvb_tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 32 lanes x 32 consecutive columns (one accumulator row slice per thread)
__device__ __forceinline__ void nvb_tmem_ld32(uint32_t taddr, uint32_t (&r)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void nvb_tmem_ld16(uint32_t taddr, uint32_t (&r)[16])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}

// Shared-memory matrix descriptor of a K-major operand tile whose rows are KCH bytes
// (KCH = 64: 64-byte swizzle, KCH = 128: 128-byte swizzle), rows packed, groups of 8 rows
// 8 * KCH bytes apart.  Bits: [0,14) start address >> 4, [16,30) leading byte offset >> 4
// (unused for swizzled K-major, 1), [32,46) stride byte offset >> 4, [46,48) version = 1
// (sm_100), [61,64) layout: 2 = 128-byte swizzle, 4 = 64-byte swizzle.
template <int KCH>
__device__ __forceinline__ uint64_t nvb_umma_desc(const void *tile)
{
    const uint32_t addr = nvb_smem_u32(tile);
    uint64_t d = (uint64_t)((addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)((8 * KCH) >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(KCH == 128 ? 2 : 4) << 61;
    return d;
}

// Instruction descriptor, kind::i8: bits [4,6) accumulator format 2 = S32, [7,10) A format
// 1 = signed 8 bit, [10,13) B format 1 = signed 8 bit, bit 15 / 16 A / B major 0 = K,
// [17,23) N >> 3, [24,29) M >> 4.
__host__ __device__ constexpr uint32_t nvb_umma_idesc_i8(int M, int N)
{
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// Folds W (32 or 16) accumulator columns [c0, c0 + W) of this thread's row into the running
// minimum of -256 * dot + column (largest dot first, then lowest column); columns >= nvalid
// (past the end of the library) are skipped.
template <int W>
__device__ __forceinline__ int nvb_tc_fold(uint32_t taddr, int c0, int nvalid, int best)
{
    uint32_t r[W];
    if (W == 32) nvb_tmem_ld32(taddr + (uint32_t)c0, reinterpret_cast<uint32_t (&)[32]>(r));
    else nvb_tmem_ld16(taddr + (uint32_t)c0, reinterpret_cast<uint32_t (&)[16]>(r));
    nvb_tmem_wait_ld();
    if (c0 + W <= nvalid) {
#pragma unroll
        for (int j = 0; j + 1 < W; j += 2)
            best = __vimin3_s32(best, (int)r[j] * -256 + (c0 + j), (int)r[j + 1] * -256 + (c0 + j + 1));
    } else {
#pragma unroll
        for (int j = 0; j < W; j++)
            if (c0 + j < nvalid) best = min(best, (int)r[j] * -256 + (c0 + j));
    }
    return best;
}

// 32 lanes x 64 consecutive columns
__device__ __forceinline__ void nvb_tmem_ld64(uint32_t taddr, uint32_t (&r)[64])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]), "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]), "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
        : "r"(taddr)
        : "memory");
}

// minimum of -256 * dot + column over 64 accumulator columns [c0, c0 + 64) already in registers.
// TOP2: also the second smallest key (best <= second): the step kernel resolves headings tied
// at the integer minimum from these two instead of rescanning the tile (step_tm.cuh).  Second
// smallest of {best, second, lo, hi} with best <= second, lo <= hi is min3(max(best, lo), second, hi).
template <bool TOP2>
__device__ __forceinline__ void nvb_tc_min64(const uint32_t (&r)[64], int c0, int nvalid, int &best, int &second)
{
#if defined(NVB_TC_EXP_NO_FOLD)   /* tools/micro experiments only */
    best ^= (int)r[0] ^ (int)r[63];
    return;
#endif
    if (c0 + 64 <= nvalid) {
#pragma unroll
        for (int j = 0; j < 64; j += 2) {
            const int k1 = (int)r[j] * -256 + (c0 + j), k2 = (int)r[j + 1] * -256 + (c0 + j + 1);
            if (TOP2) {
                const int lo = min(k1, k2), hi = max(k1, k2);
                second = __vimin3_s32(max(best, lo), second, hi);
                best = min(best, lo);
            } else {
                best = __vimin3_s32(best, k1, k2);
            }
        }
    } else {
#pragma unroll
        for (int j = 0; j < 64; j++)
            if (c0 + j < nvalid) {
                const int k1 = (int)r[j] * -256 + (c0 + j);
                if (TOP2) second = min(second, max(best, k1));
                best = min(best, k1);
            }
    }
}

// One finished 128 x 256 accumulator (this thread's row): the row's minimum of -256 * dot +
// column.  The TMEM loads are software-pipelined -- 64 columns are folded while the next 64 are
// in flight (a load waited for right after its issue exposes the whole TMEM latency, eight
// times per item: measured 1.6 us per item against 0.7 us of MMA) -- and `release` runs as soon
// as the last load has landed, before the last fold, so the MMA warp gets the buffer back early.
template <bool TOP2, typename Release>
__device__ __forceinline__ int2 nvb_tc_fold_item(uint32_t taddr, int nvalid, Release release)
{
    static_assert(NVB_TC_NT == 256, "four chunks of 64 columns");
    uint32_t ra[64], rb[64];
    int best = 0x7FFFFFFF, second = 0x7FFFFFFF;
#if defined(NVB_TC_EXP_NO_EPI_LD)   /* tools/micro experiments only */
    release();
    return make_int2((int)taddr, 0);
#endif
    nvb_tmem_ld64(taddr, ra);
    nvb_tmem_wait_ld();
    nvb_tmem_ld64(taddr + 64u, rb);
    nvb_tc_min64<TOP2>(ra, 0, nvalid, best, second);
    nvb_tmem_wait_ld();
    nvb_tmem_ld64(taddr + 128u, ra);
    nvb_tc_min64<TOP2>(rb, 64, nvalid, best, second);
    nvb_tmem_wait_ld();
    nvb_tmem_ld64(taddr + 192u, rb);
    nvb_tc_min64<TOP2>(ra, 128, nvalid, best, second);
    nvb_tmem_wait_ld();
    release();
    nvb_tc_min64<TOP2>(rb, 192, nvalid, best, second);
    return make_int2(best, second);
}

// nvb_tc_min64 for W columns
template <bool TOP2, int W>
__device__ __forceinline__ void nvb_tc_minw(const uint32_t (&r)[W], int c0, int nvalid, int &best, int &second)
{
    if (c0 + W <= nvalid) {
#pragma unroll
        for (int j = 0; j < W; j += 2) {
            const int k1 = (int)r[j] * -256 + (c0 + j), k2 = (int)r[j + 1] * -256 + (c0 + j + 1);
            if (TOP2) {
                const int lo = min(k1, k2), hi = max(k1, k2);
                second = __vimin3_s32(max(best, lo), second, hi);
                best = min(best, lo);
            } else {
                best = __vimin3_s32(best, k1, k2);
            }
        }
    } else {
#pragma unroll
        for (int j = 0; j < W; j++)
            if (c0 + j < nvalid) {
                const int k1 = (int)r[j] * -256 + (c0 + j);
                if (TOP2) second = min(second, max(best, k1));
                best = min(best, k1);
            }
    }
}

// The same for ONE HALF of the accumulator, columns [c0, c0 + 128) (taddr points at column 0):
// two warps per TMEM lane quarter share an item (k2_tc_bs).  Four loads of 32 columns through two
// register buffers keep the kernel at ~104 registers per thread: with 352 threads that leaves room
// on every SM sub-partition (register files are per sub-partition) for two CTAs of the step
// kernel to become resident beside this one.  What that buys is launch work taken off the gap
// between the kernels: the dependent grid's remaining CTAs are launched when this grid's CTAs
// retire, and its dependency is released only after that burst (measured: gap = ~0.8 us + ~2.5 ns
// per CTA still to launch; a 168-register version with both 64-column loads in flight finished
// 0.5 us sooner and handed over 0.7 us later).  The accumulator goes back to the MMA warps as
// soon as the last load has landed, before the last two folds.
template <bool TOP2, typename Release>
__device__ __forceinline__ int2 nvb_tc_fold_half(uint32_t taddr, int c0, int nvalid, Release release)
{
    uint32_t ra[32], rb[32];
    int best = 0x7FFFFFFF, second = 0x7FFFFFFF;
#if defined(NVB_TC_EXP_NO_EPI_LD)   /* tools/micro experiments only */
    release();
    return make_int2((int)taddr, 0);
#endif
    nvb_tmem_ld32(taddr + (uint32_t)c0, ra);
    nvb_tmem_ld32(taddr + (uint32_t)c0 + 32u, rb);
    nvb_tmem_wait_ld();
    nvb_tc_minw<TOP2, 32>(ra, c0, nvalid, best, second);
    nvb_tmem_ld32(taddr + (uint32_t)c0 + 64u, ra);
    nvb_tc_minw<TOP2, 32>(rb, c0 + 32, nvalid, best, second);
    nvb_tmem_ld32(taddr + (uint32_t)c0 + 96u, rb);
    nvb_tmem_wait_ld();
    release();
    nvb_tc_minw<TOP2, 32>(ra, c0 + 64, nvalid, best, second);
    nvb_tc_minw<TOP2, 32>(rb, c0 + 96, nvalid, best, second);
    return make_int2(best, second);
}

template <int KCH, int NT, int STAGES>
struct TcCfg {
    static constexpr int TM = NVB_TC_TM;
    static constexpr int A_BYTES = TM * KCH;
    static constexpr int B_BYTES = NT * KCH;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int SMEM = STAGE_BYTES * STAGES + 1024 /* alignment slack */ + 256 /* barriers */;
    static_assert(KCH == 64 || KCH == 128, "K chunk is one 64- or 128-byte swizzle atom");
    static_assert(NT % 16 == 0 && NT >= 16 && NT <= 256, "UMMA N for M = 128");
    static_assert(A_BYTES % 1024 == 0 && B_BYTES % 1024 == 0, "operand tiles start on swizzle-pattern boundaries");
};

// Work item u -> (glimpse tile, view tile).  gt-major keeps a CTA on one glimpse tile; vt-major
// makes consecutive items share the VIEW tile: with a library far larger than L2 and a few
// glimpse tiles, every view tile then comes from HBM once and is reused from L2 by the other
// glimpse tiles (10^6 views x 640 glimpses: 0.4 GB of operand traffic instead of 1.9 GB).
__device__ __forceinline__ void nvb_tc_item(const TcArgs &a, int u, int &gt, int &vt)
{
    if (a.vt_major) { vt = u / a.n_gt; gt = u - vt * a.n_gt; }
    else { gt = u / a.n_vt; vt = u - gt * a.n_vt; }
}
__device__ __forceinline__ void nvb_tc_next(const TcArgs &a, int &gt, int &vt)
{
    if (a.vt_major) { if (++gt == a.n_gt) { gt = 0; vt++; } }
    else { if (++vt == a.n_vt) { vt = 0; gt++; } }
}

template <int KCH, int NT, int STAGES, bool TILEMIN = false>
__global__ void __launch_bounds__(NVB_TC_THREADS, 1)
k2_tc(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, TcArgs a)
{
    using C = TcCfg<KCH, NT, STAGES>;
    extern __shared__ uint8_t smem_tc_raw[];
    // 1024-byte alignment in the shared window (the swizzle pattern is a function of the address)
    uint8_t *smem = smem_tc_raw + ((1024u - (nvb_smem_u32(smem_tc_raw) & 1023u)) & 1023u);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + C::STAGE_BYTES * STAGES);
    uint64_t *empty = full + STAGES;
    uint64_t *tfull = empty + STAGES;    // accumulator buffer ready for the epilogue
    uint64_t *tempty = tfull + 2;        // accumulator buffer drained
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tempty + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    nvb_tl_stamp(a.tl, 0, 0);
    if (a.pdl_early) nvb_grid_dep_launch();
    if (warp == 0 && lane == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; s++) {
            nvb_mbar_init(full + s, 1);
            nvb_mbar_init(empty + s, 1);
        }
        nvb_mbar_init(tfull + 0, 1); nvb_mbar_init(tfull + 1, 1);
        nvb_mbar_init(tempty + 0, 4); nvb_mbar_init(tempty + 1, 4);
        nvb_fence_barrier_init();
        nvb_prefetch_tmap(&tm_a);
        nvb_prefetch_tmap(&tm_b);
    }
    if (warp == 2) nvb_tmem_alloc(tmem_slot, 512);
    nvb_tc_fence_before();
    __syncthreads();
    nvb_tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int u0 = a.spans[blockIdx.x], u1 = a.spans[blockIdx.x + 1];

    nvb_grid_dep_wait();   // the glimpses are written by the previous kernel of the step sequence
    nvb_tl_stamp(a.tl, 0, 1);
    if (a.step_counter != nullptr && blockIdx.x == 0 && tid == 0) {
        *a.step_counter += 1;
        a.tie_count[0] = 0;
        a.tie_count[1] = 0;
        if (a.epoch != nullptr) *a.epoch += 1;
    }

    if (warp == 0) {
        // ---- TMA producer
        int s = 0;
        uint32_t ph = 0;
        int gt, vt;
        nvb_tc_item(a, u0, gt, vt);
        for (int u = u0; u < u1; u++) {
            for (int kc = 0; kc < a.kchunks; kc++) {
                if (lane == 0) {
                    nvb_mbar_wait(empty + s, ph ^ 1u);
                    uint8_t *st = smem + s * C::STAGE_BYTES;
                    nvb_mbar_expect_tx(full + s, (uint32_t)C::STAGE_BYTES);
                    nvb_tma_load_2d(st, &tm_a, kc * KCH, gt * C::TM, full + s);
                    nvb_tma_load_2d(st + C::A_BYTES, &tm_b, kc * KCH, vt * NT, full + s);
                }
                __syncwarp();
                if (++s == STAGES) { s = 0; ph ^= 1u; }
            }
            nvb_tc_next(a, gt, vt);
        }
    } else if (warp == 1) {
        // ---- MMA issuer
        constexpr uint32_t idesc = nvb_umma_idesc_i8(C::TM, NT);
        int s = 0;
        uint32_t ph = 0;
        int it = 0;
        for (int u = u0; u < u1; u++, it++) {
            const int buf = it & 1;
            const uint32_t tph = (uint32_t)(it >> 1) & 1u;
            if (lane == 0) {
                nvb_mbar_wait(tempty + buf, tph ^ 1u);   // the epilogue has drained this buffer
                nvb_tc_fence_after();
            }
            __syncwarp();
            for (int kc = 0; kc < a.kchunks; kc++) {
                if (lane == 0) {
                    nvb_mbar_wait(full + s, ph);
                    nvb_tc_fence_after();
                    const uint8_t *st = smem + s * C::STAGE_BYTES;
                    const uint64_t da = nvb_umma_desc<KCH>(st), db = nvb_umma_desc<KCH>(st + C::A_BYTES);
#pragma unroll
                    for (int j = 0; j < KCH / 32; j++)   // 32 bytes of K per instruction: +2 in 16-byte units
                        nvb_umma_i8(tmem_base + (uint32_t)(buf * 256), da + (uint64_t)(2 * j), db + (uint64_t)(2 * j),
                                    idesc, (uint32_t)((kc | j) != 0));
                    nvb_umma_commit(empty + s);                       // stage free once these MMAs have read it
                    if (kc == a.kchunks - 1) nvb_umma_commit(tfull + buf);   // accumulator complete
                }
                __syncwarp();
                if (++s == STAGES) { s = 0; ph ^= 1u; }
            }
        }
    } else {
        // ---- epilogue: warp w may touch TMEM lanes 32 * (w % 4) .. + 31
        const int ew = warp & 3;
        const int row = ew * 32 + lane;
        int gt, vt;
        nvb_tc_item(a, u0, gt, vt);
        int it = 0;
        for (int u = u0; u < u1; u++, it++) {
            const int buf = it & 1;
            const uint32_t tph = (uint32_t)(it >> 1) & 1u;
            nvb_mbar_wait(tfull + buf, tph);
            nvb_tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(buf * 256);
            const int nvalid = min(NT, a.N - vt * NT);
            int best, second = 0x7FFFFFFF;   // min over columns of -256 * dot + column (and the runner-up)
            if (NT == 256) {
                const int2 b2 = nvb_tc_fold_item<TILEMIN>(taddr, nvalid, [&]() {
                    nvb_tc_fence_before();
                    __syncwarp();
                    if (lane == 0) nvb_mbar_arrive(tempty + buf);
                });
                best = b2.x; second = b2.y;
            } else {
                best = 0x7FFFFFFF;
#pragma unroll
                for (int c0 = 0; c0 < NT; c0 += 32) {
                    if (NT - c0 >= 32) best = nvb_tc_fold<32>(taddr, c0, nvalid, best);
                    else best = nvb_tc_fold<16>(taddr, c0, nvalid, best);
                }
                nvb_tc_fence_before();
                __syncwarp();
                if (lane == 0) nvb_mbar_arrive(tempty + buf);
            }
            const int g = gt * C::TM + row;
            if (TILEMIN) {
                if (g < a.G) a.tmin[(size_t)g * a.n_vt + vt] = make_int2(best, second);
            } else if (g < a.G && best != 0x7FFFFFFF) {
                const int col = best & 255;
                const int dot = -(best >> 8);
                const unsigned long long sad = (unsigned long long)((a.sad_const - dot) >> 1);
                const unsigned long long v = (unsigned long long)(a.view_offset + (long long)vt * NT + col);
                atomicMin(a.keys + g, (sad << 32) | v);
            }
            nvb_tc_next(a, gt, vt);
        }
    }
    nvb_tc_fence_before();
    __syncthreads();
    if (warp == 2) nvb_tmem_dealloc(tmem_base, 512);
    nvb_tl_stamp(a.tl, 0, 2);
}

// ---- view-tile-stationary variant ----------------------------------------------------------
// Sensors of few pixels (the reference's default 40 x 2 sensor at 5 levels: K = 320 bytes per
// row) leave the streaming kernel above bound by its own bookkeeping, not by the tensor pipe.
// Measured with in-kernel cycle stamps (tools/micro/tc_sad.cu, NVB_TC_EXP_STAMPS): the thread
// that issues the MMAs spends ~450 cycles in every tcgen05.commit, and a kernel that releases
// each K-chunk stage with its own commit pays that five or six times per 128 x 256 item --
// 3200 cycles per item against 1340 cycles of MMA.  Here:
//   * the VIEW tile (256 rows x all K) stays resident in shared memory while the CTA walks the
//     glimpse tiles of its span (items are ordered view-tile-major): an item only loads its 128
//     glimpse rows, and a view tile of a 10^6-view library comes from HBM exactly once;
//   * the glimpse rows of an item land as ONE group of K-chunks on one barrier, the MMA thread
//     issues the whole item and commits ONCE (accumulator complete); the epilogue, woken by
//     that commit, hands the glimpse slot back to the producer with a plain mbarrier arrive;
//   * TWO warps issue MMAs, one the even items of the span (accumulator buffer 0), one the odd
//     items (buffer 1): while one sits out its commit the other keeps the tensor pipe fed
//     (one issuer alone: 1840 cycles per item, 1340 of them MMA).
// K is cut into KCH-byte chunks (64: 64-byte swizzle, no padding beyond a multiple of 64).
// Shared memory: kchunks x NT x KCH (view tile) + a_slots x kchunks x TM x KCH (glimpse ring).
#define NVB_TCBS_KCH 64
#define NVB_TCBS_MAX_SLOTS 4
#ifndef NVB_TCBS_MAXNREG
#define NVB_TCBS_MAXNREG 112   /* 352 threads x 112 registers leave room for two step-kernel CTAs beside this one (see nvb_tc_fold_half) */
#endif
#define NVB_TCBS_THREADS 352   /* producer, MMA issuer (even items), 4 epilogue warps (columns 0..127), MMA issuer (odd items), 4 epilogue warps (columns 128..255) */

__host__ __device__ inline int nvb_tcbs_smem(int kchunks, int a_slots, int kch = NVB_TCBS_KCH)
{
    return kchunks * NVB_TC_NT * kch + a_slots * kchunks * NVB_TC_TM * kch + 1024 /* alignment slack */ + 256 /* barriers */;
}
// glimpse slots that fit beside the view tile (at least 2 for the kernel to apply)
__host__ __device__ inline int nvb_tcbs_slots(int kchunks, int kch = NVB_TCBS_KCH)
{
    // (200 KB of the 227: measured on C2, a third glimpse slot -- 206 KB -- does not pay: the item
    // period is set by the accumulator hand-over, not by the glimpse loads, tools/k2_situ.py)
    const int s = (200 * 1024 - kchunks * NVB_TC_NT * kch) / (kchunks * NVB_TC_TM * kch);
    return s > NVB_TCBS_MAX_SLOTS ? NVB_TCBS_MAX_SLOTS : s;
}

// LEAN: few items per CTA -- 352 threads, eight register-lean epilogue warps (see nvb_tc_fold_half); else 224
// threads, four epilogue warps
template <bool TILEMIN, int KCH = NVB_TCBS_KCH, bool LEAN = true>
__global__ void __maxnreg__(LEAN ? NVB_TCBS_MAXNREG : 168)
k2_tc_bs(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, TcArgs a, int a_slots)
{
    constexpr int NT = NVB_TC_NT, TM = NVB_TC_TM;
    constexpr int A_BYTES = TM * KCH, B_BYTES = NT * KCH;
    extern __shared__ uint8_t smem_tc_raw[];
    uint8_t *smem = smem_tc_raw + ((1024u - (nvb_smem_u32(smem_tc_raw) & 1023u)) & 1023u);
    uint8_t *smem_b = smem;                                   // [kchunks][NT][KCH]
    uint8_t *smem_a = smem + (size_t)a.kchunks * B_BYTES;     // [a_slots][kchunks][TM][KCH]
    const size_t slot_bytes = (size_t)a.kchunks * A_BYTES;
    uint64_t *afull = reinterpret_cast<uint64_t *>(smem_a + (size_t)a_slots * slot_bytes);
    uint64_t *aempty = afull + NVB_TCBS_MAX_SLOTS;
    uint64_t *bfull = aempty + NVB_TCBS_MAX_SLOTS;   // view tile landed
    uint64_t *bempty = bfull + 1;                    // every MMA that reads the view tile has completed
    uint64_t *tfull = bempty + 1;                    // accumulator buffer ready for the epilogue
    uint64_t *tempty = tfull + 2;                    // accumulator buffer drained
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tempty + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    nvb_tl_stamp(a.tl, 0, 0);
    if (a.pdl_early) nvb_grid_dep_launch();
    if (warp == 0 && lane == 0) {
        for (int s = 0; s < a_slots; s++) {
            nvb_mbar_init(afull + s, 1);
            nvb_mbar_init(aempty + s, 1);
        }
        nvb_mbar_init(bfull, 1);
        nvb_mbar_init(bempty, 3);   // both issuers have seen the tile + the epilogue has seen its last item complete
        nvb_mbar_init(tfull + 0, 1); nvb_mbar_init(tfull + 1, 1);
        nvb_mbar_init(tempty + 0, LEAN ? 8 : 4); nvb_mbar_init(tempty + 1, LEAN ? 8 : 4);
        nvb_fence_barrier_init();
        nvb_prefetch_tmap(&tm_a);
        nvb_prefetch_tmap(&tm_b);
    }
    if (warp == 2) nvb_tmem_alloc(tmem_slot, 512);
    nvb_tc_fence_before();
    __syncthreads();
    nvb_tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int u0 = a.spans[blockIdx.x], u1 = a.spans[blockIdx.x + 1];

    if (warp == 0) {
        // ---- TMA producer.  The library is not written by any kernel of the step sequence: the
        // first view tile is requested before the dependency wait, the glimpse rows after it.
        int s = 0;
        uint32_t ph = 0, bph = 0;
        int cur_vt = -1;
        bool waited = false;
        for (int u = u0; u < u1; u++) {
            const int vt = u / a.n_gt, gt = u - vt * a.n_gt;
            if (lane == 0 && vt != cur_vt) {
                if (cur_vt >= 0) { nvb_mbar_wait(bempty, bph); bph ^= 1u; }
                nvb_mbar_expect_tx(bfull, (uint32_t)(a.kchunks * B_BYTES));
                for (int kc = 0; kc < a.kchunks; kc++)
                    nvb_tma_load_2d(smem_b + (size_t)kc * B_BYTES, &tm_b, kc * KCH, vt * NT, bfull);
            }
            cur_vt = vt;
            if (!waited) {
                nvb_grid_dep_wait();   // the glimpses are written by the previous kernel of the step sequence
                nvb_tl_stamp(a.tl, 0, 1);
                if (lane == 0) NVB_TC_STAMP(0, 3);
                waited = true;
            }
            if (lane == 0) {
                nvb_mbar_wait(aempty + s, ph ^ 1u);
                nvb_mbar_expect_tx(afull + s, (uint32_t)slot_bytes);
                for (int kc = 0; kc < a.kchunks; kc++)
                    nvb_tma_load_2d(smem_a + (size_t)s * slot_bytes + (size_t)kc * A_BYTES, &tm_a, kc * KCH, gt * TM, afull + s);
            }
            __syncwarp();
            if (++s == a_slots) { s = 0; ph ^= 1u; }
        }
        if (!waited) nvb_grid_dep_wait();
    } else {
        nvb_grid_dep_wait();
    }
    if (a.step_counter != nullptr && blockIdx.x == 0 && tid == 32) {
        *a.step_counter += 1;
        a.tie_count[0] = 0;
        a.tie_count[1] = 0;
        if (a.epoch != nullptr) *a.epoch += 1;
    }

    if (warp == 1 || warp == 6) {
        // ---- MMA issuers: warp 1 the even items of the span, warp 6 the odd ones.  Per item one
        // wait for its inputs, all its MMAs, one commit.  Both walk every item to keep track of
        // the view tile: each acknowledges a tile once it has seen it land (third of the three
        // arrivals that let the producer replace it is the epilogue's, after the tile's last
        // item), so neither can fall a barrier phase behind.
        constexpr uint32_t idesc = nvb_umma_idesc_i8(TM, NT);
        const int mine = (warp == 1) ? 0 : 1;
        uint32_t bph = 0;
        int cur_vt = -1;
        for (int u = u0; u < u1; u++) {
            const int it = u - u0;
            const int vt = u / a.n_gt;
            if (lane == 0 && vt != cur_vt) {
                nvb_mbar_wait(bfull, bph);
                bph ^= 1u;
                if (vt != (u1 - 1) / a.n_gt) nvb_mbar_arrive(bempty);   // (the last tile of the span is never replaced)
            }
            cur_vt = vt;
            if ((it & 1) != mine) continue;
            const int buf = mine, s = it % a_slots;
            const uint32_t tph = (uint32_t)(it >> 1) & 1u, ph = (uint32_t)(it / a_slots) & 1u;
            if (lane == 0) {
                nvb_mbar_wait(tempty + buf, tph ^ 1u);   // the epilogue has drained this buffer
                nvb_mbar_wait(afull + s, ph);
                nvb_tc_fence_after();
                NVB_TC_STAMP(it, 0);
                const uint8_t *slot = smem_a + (size_t)s * slot_bytes;
                for (int kc = 0; kc < a.kchunks; kc++) {
                    const uint64_t da = nvb_umma_desc<KCH>(slot + (size_t)kc * A_BYTES);
                    const uint64_t db = nvb_umma_desc<KCH>(smem_b + (size_t)kc * B_BYTES);
#pragma unroll
                    for (int j = 0; j < KCH / 32; j++)   // 32 bytes of K per instruction: +2 in 16-byte units
                        nvb_umma_i8(tmem_base + (uint32_t)(buf * 256), da + (uint64_t)(2 * j), db + (uint64_t)(2 * j),
                                    idesc, (uint32_t)((kc | j) != 0));
                }
                NVB_TC_STAMP(it, 1);
                nvb_umma_commit(tfull + buf);   // accumulator complete
                NVB_TC_STAMP(it, 2);
            }
            __syncwarp();
        }
    } else if (warp >= 2) {
        // ---- epilogue: warp w may touch TMEM lanes 32 * (w % 4) .. + 31.  TWO warps per lane quarter:
        // warps 2..5 fold columns 0..127 of an item, warps 7..10 columns 128..255 (measured inside
        // the step graph, tools/k2_situ.py: four warps folding 256 columns each took 0.9 - 1.5 us
        // per item against 0.68 us of MMA -- the epilogue, not the tensor pipe, paced the kernel).
        // Tile minima: the upper half hands its (best, runner-up) to the lower half through shared
        // memory (double buffered like the accumulators), which merges and stores.
        // (not LEAN -- many items per CTA, launched with 224 threads: four epilogue warps fold the
        // whole accumulator each, four pipelined 64-column loads, no merge: 0.83 us per item at
        // 2000 items per CTA against 0.88 us for the eight-warp form)
        __shared__ int2 s_part[2][NVB_TC_TM];
        const int ew = warp & 3, half = (warp >= 7) ? 1 : 0;
        const int row = ew * 32 + lane;
        int it = 0, s = 0;
        for (int u = u0; u < u1; u++, it++) {
            const int vt = u / a.n_gt, gt = u - vt * a.n_gt;
            const int buf = it & 1;
            const uint32_t tph = (uint32_t)(it >> 1) & 1u;
            nvb_mbar_wait(tfull + buf, tph);
            nvb_tc_fence_after();
            // the item's MMAs have completed (and, items being waited for in order, those of every
            // earlier item): its glimpse slot goes back to the producer, and so does the view tile
            // if this was the last item on it
            if (warp == 2 && lane == 0) {
                nvb_mbar_arrive(aempty + s);
                if (u + 1 < u1 && (u + 1) / a.n_gt != vt) nvb_mbar_arrive(bempty);
            }
            if (++s == a_slots) s = 0;
            if (tid == 64) NVB_TC_STAMP(it, 4);
            const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(buf * 256);
            const int nvalid = min(NT, a.N - vt * NT);
            // min over this half's columns of -256 * dot + column
            auto release = [&]() {
                nvb_tc_fence_before();
                __syncwarp();
                if (lane == 0) nvb_mbar_arrive(tempty + buf);
                if (tid == 64) NVB_TC_STAMP(it, 5);
            };
            int2 b2;
            if (LEAN) b2 = nvb_tc_fold_half<TILEMIN>(taddr, 128 * half, nvalid, release);
            else b2 = nvb_tc_fold_item<TILEMIN>(taddr, nvalid, release);
            if (tid == 64) NVB_TC_STAMP(it, 6);
            const int g = gt * TM + row;
            if (TILEMIN && !LEAN) {
                if (g < a.G) a.tmin[(size_t)g * a.n_vt + vt] = b2;
            } else if (TILEMIN) {
                if (half) s_part[buf][row] = b2;
                asm volatile("bar.sync 1, 256;" ::: "memory");   // the eight epilogue warps, once per item
                if (!half) {
                    const int2 o = s_part[buf][row];
                    // the two smallest of {b2.x <= b2.y} and {o.x <= o.y}
                    b2 = make_int2(min(b2.x, o.x), __vimin3_s32(max(b2.x, o.x), b2.y, o.y));
                    if (g < a.G) a.tmin[(size_t)g * a.n_vt + vt] = b2;
                }
            } else if (g < a.G && b2.x != 0x7FFFFFFF) {
                const int best = b2.x;
                const int col = best & 255;
                const int dot = -(best >> 8);
                const unsigned long long sad = (unsigned long long)((a.sad_const - dot) >> 1);
                const unsigned long long v = (unsigned long long)(a.view_offset + (long long)vt * NT + col);
                atomicMin(a.keys + g, (sad << 32) | v);
            }
        }
    }
    nvb_tc_fence_before();
    __syncthreads();
    if (warp == 2) nvb_tmem_dealloc(tmem_base, 512);
    nvb_tl_stamp(a.tl, 0, 2);
}

// ---- operand encoding ---------------------------------------------------------------
// src [rows][Ppad] uint8 (quantised V plane) -> dst [rows][Kpad] int8, pixel-major:
// k = p * n_planes + plane; glimpse side (IS_A): +-w_k, library side: +-1.  Bytes
// k >= n_planes * P stay zero (the buffer is zeroed once when it is allocated).
// level_of[v] = index of the quantised value v among the levels, level_of[256 + v] = 1 if v
// is one of the levels; a value that is not (a library or a query uploaded from the host that
// was not produced by this sensor's quantisation) sets *bad: the caller then falls back to the
// byte-SIMD kernel, which takes any bytes.
// (In the stepping loop the sampler writes the glimpse side itself, sampler.cuh.)
template <bool IS_A>
__global__ void k_tc_encode(const uint8_t *src, long long rows, int P, int Ppad, int Kpad, TcPlanes pl,
                            const uint8_t *level_of, int8_t *dst, int *bad)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * P) return;
    const long long r = i / P;
    const int p = (int)(i - r * P);
    const int v = src[(size_t)r * Ppad + p];
    const int lvl = level_of[v];
    if (!level_of[256 + v]) *bad = 1;
    int8_t *o = dst + (size_t)r * Kpad + (size_t)p * pl.n_planes;
#pragma unroll
    for (int k = 0; k < NVB_TC_MAX_PLANES; k++)
        if (k < pl.n_planes) {
            const int sgn = (lvl > (int)pl.thr_level[k]) ? 1 : -1;
            o[k] = (int8_t)(IS_A ? sgn * (int)pl.weight[k] : sgn);
        }
}

// int8 tensor-core issue-rate probe (bench.py's roofline denominator for k2_tc, next to the
// measured bf16 figure of MEASURED_PEAKS.json): every CTA (one per SM) issues `iters` x 4
// back-to-back 128 x 256 x 32 MMAs on operands that stay in shared memory.
#define NVB_PROBE_UMMA_SMEM (16384 + 32768 + 1024)
__global__ void __launch_bounds__(128, 1) k_probe_umma(int iters)
{
    extern __shared__ uint8_t smem_probe_raw[];
    uint8_t *smem = smem_probe_raw + ((1024u - (nvb_smem_u32(smem_probe_raw) & 1023u)) & 1023u);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x;
    for (int i = tid; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x01FF01FFu * (uint32_t)(i & 3);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the MMA's reads
    if (tid == 0) { nvb_mbar_init(&bar, 1); nvb_fence_barrier_init(); }
    if (tid < 32) nvb_tmem_alloc(&slot, 512);
    nvb_tc_fence_before();
    __syncthreads();
    nvb_tc_fence_after();
    const uint32_t tmem_base = slot;
    if (tid == 0) {
        constexpr uint32_t idesc = nvb_umma_idesc_i8(128, 256);
        const uint64_t da = nvb_umma_desc<128>(smem), db = nvb_umma_desc<128>(smem + 16384);
        for (int it = 0; it < iters; it++)
#pragma unroll
            for (int j = 0; j < 4; j++)
                nvb_umma_i8(tmem_base + (uint32_t)((it & 1) * 256), da + (uint64_t)(2 * j), db + (uint64_t)(2 * j), idesc, 1u);
        nvb_umma_commit(&bar);
        nvb_mbar_wait(&bar, 0);
    }
    nvb_tc_fence_before();
    __syncthreads();
    if (tid < 32) nvb_tmem_dealloc(tmem_base, 512);
}
