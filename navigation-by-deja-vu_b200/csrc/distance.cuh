// distance.cuh -- K2: batched glimpse-vs-library image difference with a fused
// per-glimpse min/argmin.
//
// Replaces sads_hsv_metric (navsim/util.pyx:28-73) plus the per-heading max of
// step_forward (navsim/NavBySceneFamiliarity.py:313): for every glimpse
// g = (agent, heading) it finds the library view with the smallest image
// difference.  With chem_weight == 0 (util.pyx:10 default) the metric is
// sum |dV| / 255 on one uint8 plane, so the kernel works on the integer sum
// (order-equivalent to the reference's FP64 sum whenever two sums differ;
// equal sums are resolved in exact FP64 by the tie pass in step.cuh).
//
// Layout: glimpses gv [G][Ppad], library lv [N][Ppad], uint8, rows 16-B aligned,
// pad bytes zero.  A CTA owns a TG x TN (glimpse x view) tile; both operand
// tiles are streamed into shared memory in K-chunks of KC = 16*CPR bytes with
// a cp.async multi-stage ring (XOR-swizzled 16-B chunks when CPR is even so the
// row-strided LDS.128 reads are bank-conflict free); every thread accumulates
// an MG x MV register tile with byte-SIMD VABSDIFF4+accumulate (nvb_sad4: four
// pixel differences per instruction, SASS VABSDIFF4.U8.ACC).  The epilogue folds (sum, view) into one
// packed key and reduces min over views with warp shuffles and one 64-bit
// atomicMin per glimpse row.
//
// key = (score << idx_bits) | view_index     (lower is more familiar)
#pragma once
#include <type_traits>

#include "common.cuh"

struct DistArgs {
    const uint8_t *gv, *gh, *gs;  // glimpses [G][Ppad]
    const uint8_t *lv, *lh, *ls;  // library  [N][Ppad]
    int G, N, Ppad;
    int nk;                       // K-chunks per row
    int n_vt, vt_per_split;       // view tiles (and, for the HSV kernel, tiles per blockIdx.y)
    const int *spans;             // [gridDim.x + 1] unit boundaries per CTA (k2_sad_v)
    int *step_counter;            // resident loop: device step index, bumped once per launch; else nullptr
    int *tie_count;               // resident loop: [0] tie work list length, [1] units done; reset per launch
    unsigned long long *epoch;    // view shards over NVLink: launches so far (sequence base of the exchanges), else nullptr
    long long view_offset;        // global index of local view 0 (library shards)
    unsigned long long *keys;     // [G], pre-set to ~0
    double cw;
    int idx_bits;
    int pdl_early;                // trigger the dependent launch at the top of the kernel
    long long *tl;                // tuning aid: timeline stamps (nvb_tl_stamp) or nullptr
};

#define NVB_DIST_THREADS 256
#ifndef NVB_K2_UNROLL_MAX
#define NVB_K2_UNROLL_MAX 5   /* rows of up to this many 16-byte chunks get a fully unrolled SAD loop */
#endif

__device__ __forceinline__ void nvb_cp_async16(void *dst, const void *src, int src_bytes)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(nvb_smem_u32(dst)), "l"(src),
                 "r"(src_bytes)
                 : "memory");
}
__device__ __forceinline__ void nvb_cp_async_commit()
{
    asm volatile("cp.async.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void nvb_cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// swizzle of the 16-B chunk index for a row of CPR chunks
template <int CPR>
__device__ __forceinline__ int nvb_swz(int row)
{
    if (CPR == 8) return row & 7;          // 128-B rows: chunk ^= row[0:3]
    if (CPR == 4) return (row >> 1) & 3;   // 64-B rows
    if (CPR == 2) return (row >> 2) & 1;   // 32-B rows
    return 0;                              // odd chunk counts are conflict free as they are
}

template <int TY, int MG, int MV, int CPR, int STAGES>
struct DistCfg {
    static constexpr int TX = NVB_DIST_THREADS / TY;
    static constexpr int TG = TY * MG;
    static constexpr int TN = TX * MV;
    static constexpr int KC = 16 * CPR;
    static constexpr int STAGE_BYTES = (TG + TN) * KC;
    static constexpr int SMEM = STAGE_BYTES * STAGES + 128;  // + full / empty mbarriers
};

// Work decomposition: a "unit" is one TG x TN (glimpse tile x view tile) block,
// linearised glimpse-tile-major (u = gt * n_vt + vt).  The host cuts the unit list
// into one contiguous, cost-balanced span per CTA (spans[blockIdx.x] ..
// spans[blockIdx.x + 1]); a CTA therefore stays on one glimpse tile for nearly its
// whole span, keeps a per-thread running minimum across its units, and only
// reduces across lanes / touches global memory when the glimpse tile changes.
//
// BULK = true (rows are exactly KC bytes: nk == 1, odd CPR): both operand tiles
// are contiguous in global memory and are staged by ONE TMA bulk copy each
// (cp.async.bulk, SASS UBLKCP) completing on a per-stage mbarrier.
// BULK = false: 16-B cp.async (LDGSTS) with an XOR swizzle, any row length.
template <int TY, int MG, int MV, int CPR, int STAGES, bool BULK>
__global__ void __launch_bounds__(NVB_DIST_THREADS, (MG * MV <= 32) ? 3 : 2)
k2_sad_v(DistArgs a)
{
    using C = DistCfg<TY, MG, MV, CPR, STAGES>;
    constexpr int TX = C::TX, TG = C::TG, TN = C::TN, KC = C::KC;
    static_assert(MV >= 2 && MV <= 16, "MV (views per thread) must be in 2..16");
    extern __shared__ __align__(1024) uint8_t smem_k2[];
    uint8_t *smem = smem_k2;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + C::STAGE_BYTES * STAGES);
    uint64_t *empty = full + STAGES;   // BULK: one arrival per warp when a stage has been consumed

    const int tid = threadIdx.x;
    const int tx = tid % TX, ty = tid / TX;
    if (BULK) {
        if (tid == 0) {
#pragma unroll
            for (int s = 0; s < STAGES; s++) {
                nvb_mbar_init(full + s, 1);
                nvb_mbar_init(empty + s, NVB_DIST_THREADS / 32);
            }
            nvb_fence_barrier_init();
        }
        __syncthreads();
    }
    nvb_tl_stamp(a.tl, 0, 0);
    if (a.pdl_early) nvb_grid_dep_launch();
    // spans, the library and the tile geometry are not written by any kernel of the step
    // sequence: a CTA that became resident early starts fetching its first view tiles here
    const int u0 = a.spans[blockIdx.x], u1 = a.spans[blockIdx.x + 1];
    const int nk = BULK ? 1 : a.nk;
    const int total = (u1 - u0) * nk;

    // (gt, vt, kc) of the next job to LOAD, advanced incrementally (no divisions in the loop)
    int l_gt = u0 / a.n_vt, l_vt = u0 - l_gt * a.n_vt, l_kc = 0, l_it = 0;
    // BULK, thread 0: the two bulk copies of job (g_t, v_t) into `slot`.  The view part arms
    // the stage's barrier for both parts, so it is always issued first.
    auto issue_bulk = [&](int g_t, int v_t, int slot, bool part_v, bool part_g) {
        uint8_t *st = smem + slot * C::STAGE_BYTES;
        const int rows_g = min(TG, a.G - g_t * TG), rows_v = min(TN, a.N - v_t * TN);
        if (part_v) {
            nvb_mbar_expect_tx(full + slot, (uint32_t)((rows_g + rows_v) * KC));
            nvb_bulk_load_1d(st + TG * KC, a.lv + (size_t)v_t * TN * KC, (uint32_t)(rows_v * KC), full + slot);
        }
        if (part_g)
            nvb_bulk_load_1d(st, a.gv + (size_t)g_t * TG * KC, (uint32_t)(rows_g * KC), full + slot);
    };
    if (BULK && tid == 0) {
        int g_t = l_gt, v_t = l_vt;
        for (int s = 0; s < STAGES - 1 && s < total; s++) {
            issue_bulk(g_t, v_t, s, true, false);
            if (++v_t == a.n_vt) { v_t = 0; g_t++; }
        }
    }
    nvb_grid_dep_wait();   // everything above overlaps the previous kernel's tail
    nvb_tl_stamp(a.tl, 0, 1);
    if (a.step_counter != nullptr && blockIdx.x == 0 && tid == 0) {
        // one step-batch = one launch of this kernel: next log slot, empty tie list
        // (the kernels that read them run after this one)
        *a.step_counter += 1;
        a.tie_count[0] = 0;   // list length
        a.tie_count[1] = 0;   // tie units completed (tie pass folded into move+sample)
        if (a.epoch != nullptr) *a.epoch += 1;
    }
    if (total <= 0) return;

    auto load_next = [&]() {
        uint8_t *st = smem + (l_it % STAGES) * C::STAGE_BYTES;
        if (BULK) {
            if (tid == 0) issue_bulk(l_gt, l_vt, l_it % STAGES, true, true);
        } else {
            const int kbyte = l_kc * KC;
            for (int q = tid; q < (TG + TN) * CPR; q += NVB_DIST_THREADS) {
                const int row = q / CPR, c = q - row * CPR;
                const uint8_t *src;
                int ok;
                if (row < TG) {
                    const int g = l_gt * TG + row;
                    ok = (g < a.G) && (kbyte + 16 * c < a.Ppad);
                    src = a.gv + (size_t)(ok ? g : 0) * a.Ppad + (ok ? kbyte + 16 * c : 0);
                } else {
                    const int v = l_vt * TN + (row - TG);
                    ok = (v < a.N) && (kbyte + 16 * c < a.Ppad);
                    src = a.lv + (size_t)(ok ? v : 0) * a.Ppad + (ok ? kbyte + 16 * c : 0);
                }
                const int r = (row < TG) ? row : row - TG;
                uint8_t *dst = st + (row < TG ? 0 : TG * KC) + r * KC + 16 * (c ^ nvb_swz<CPR>(r));
                nvb_cp_async16(dst, src, ok ? 16 : 0);
            }
        }
        l_it++;
        if (++l_kc == nk) {
            l_kc = 0;
            if (++l_vt == a.n_vt) { l_vt = 0; l_gt++; }
        }
    };

    uint32_t acc[MG][MV];
#pragma unroll
    for (int i = 0; i < MG; i++)
#pragma unroll
        for (int j = 0; j < MV; j++) acc[i][j] = 0;
    // per-thread running minimum over the units of the current glimpse tile
    uint32_t best[MG];     // smallest sum so far
    uint32_t best_at[MG];  // (view tile << 4) | j of the first view that reached it
#pragma unroll
    for (int i = 0; i < MG; i++) { best[i] = 0x0EFFFFFFu; best_at[i] = 0; }

    int goff[MG], gsw[MG], voff[MV], vsw[MV];
#pragma unroll
    for (int i = 0; i < MG; i++) { int r = ty + TY * i; goff[i] = r * KC; gsw[i] = nvb_swz<CPR>(r) << 4; }
#pragma unroll
    for (int j = 0; j < MV; j++) { int r = tx + TX * j; voff[j] = TG * KC + r * KC; vsw[j] = nvb_swz<CPR>(r) << 4; }

    if (BULK) {
        // the view tiles of these jobs are already in flight: add the glimpse tiles
        for (int s = 0; s < STAGES - 1 && s < total; s++) {
            if (tid == 0) issue_bulk(l_gt, l_vt, s, false, true);
            l_it++;
            if (++l_vt == a.n_vt) { l_vt = 0; l_gt++; }
        }
    } else {
#pragma unroll
        for (int s = 0; s < STAGES - 1; s++) {
            if (s < total) load_next();
            nvb_cp_async_commit();
        }
    }

    constexpr int RW = (TX < 32) ? TX : 32;
    int gt = u0 / a.n_vt, vt = u0 - gt * a.n_vt, kc = 0;   // job being computed

    for (int it = 0; it < total; it++) {
        if (BULK) {
            // no CTA-wide barrier: the loading thread alone waits until every warp has released
            // the slot of job it-1, refills it, and each warp runs ahead as far as data has landed
            if (tid == 0 && l_it < total) {
                if (it > 0) nvb_mbar_wait(empty + ((it - 1) % STAGES), (uint32_t)(((it - 1) / STAGES) & 1));
                load_next();
            }
            nvb_mbar_wait(full + (it % STAGES), (uint32_t)((it / STAGES) & 1));
        } else {
            nvb_cp_async_wait<STAGES - 2>();
            __syncthreads();   // everyone is done with job it-1: its slot may be refilled
            if (l_it < total) load_next();
            nvb_cp_async_commit();
        }

        const uint8_t *st = smem + (it % STAGES) * C::STAGE_BYTES;
        constexpr int UNR = (CPR <= NVB_K2_UNROLL_MAX) ? CPR : 1;
#pragma unroll(UNR)
        for (int c = 0; c < CPR; c++) {
            uint4 av[MG];
#pragma unroll
            for (int i = 0; i < MG; i++)
                av[i] = *reinterpret_cast<const uint4 *>(st + goff[i] + ((c << 4) ^ gsw[i]));
#pragma unroll
            for (int j = 0; j < MV; j++) {
                const uint4 b = *reinterpret_cast<const uint4 *>(st + voff[j] + ((c << 4) ^ vsw[j]));
#pragma unroll
                for (int i = 0; i < MG; i++) {
                    uint32_t s = acc[i][j];
                    s = nvb_sad4(av[i].x, b.x, s);
                    s = nvb_sad4(av[i].y, b.y, s);
                    s = nvb_sad4(av[i].z, b.z, s);
                    s = nvb_sad4(av[i].w, b.w, s);
                    acc[i][j] = s;
                }
            }
        }

        if (BULK) {
            __syncwarp();
            if ((tid & 31) == 0) nvb_mbar_arrive(empty + (it % STAGES));   // this warp is done with the slot
        }
        if (++kc == nk) {
            kc = 0;
            // unit finished: fold its MG x MV sums into the per-thread running minimum.
            // Views past the end of the library (last tile) are excluded first.
            const int nvalid = a.N - vt * TN;
            if (nvalid < TN) {
#pragma unroll
                for (int j = 0; j < MV; j++)
                    if (tx + TX * j >= nvalid) {
#pragma unroll
                        for (int i = 0; i < MG; i++) acc[i][j] = 0x0EFFFFFFu;   // > any real sum, * 17 + 16 still fits in 32 bits
                    }
            }
#pragma unroll
            for (int i = 0; i < MG; i++) {
                // key = sum * 17 + j (j < 17): a true multiply-add, which runs on the FMA pipe
                // (a power-of-two scale would become a shift-add on the busy ALU pipe); the
                // 3-input minimum is the only ALU-pipe work per pair.  The smallest key is the
                // smallest sum and, among equal sums, the lowest view of this thread.
                uint32_t m = acc[i][0] * 17u;
#pragma unroll
                for (int j = 1; j + 1 < MV; j += 2)
                    m = __vimin3_u32(m, acc[i][j] * 17u + (uint32_t)j, acc[i][j + 1] * 17u + (uint32_t)(j + 1));
                if ((MV & 1) == 0) m = min(m, acc[i][MV - 1] * 17u + (uint32_t)(MV - 1));
                // strict < on the sum alone: an equal sum in a later tile has a higher view index
                if (m < best[i] * 17u) {
                    const uint32_t sum = m / 17u;
                    best[i] = sum;
                    best_at[i] = ((uint32_t)vt << 4) | (m - sum * 17u);
                }
#pragma unroll
                for (int j = 0; j < MV; j++) acc[i][j] = 0;
            }
            // glimpse tile ends (or span ends): reduce across the TX lanes of each row
            if (it == total - 1 || vt == a.n_vt - 1) {
#pragma unroll
                for (int i = 0; i < MG; i++) {
                    unsigned long long key = NVB_KEY_NONE;
                    if (best[i] != 0x0EFFFFFFu) {
                        const unsigned long long v = (unsigned long long)(
                            a.view_offset + (long long)(best_at[i] >> 4) * TN + tx + TX * (int)(best_at[i] & 15u));
                        key = ((unsigned long long)best[i] << a.idx_bits) | v;
                    }
#pragma unroll
                    for (int o = RW / 2; o > 0; o >>= 1) {
                        unsigned long long other = __shfl_xor_sync(0xFFFFFFFFu, key, o);
                        key = (other < key) ? other : key;
                    }
                    const int g = gt * TG + ty + TY * i;
                    if ((tx % RW) == 0 && g < a.G && key != NVB_KEY_NONE) atomicMin(a.keys + g, key);
                    best[i] = 0x0EFFFFFFu;
                    best_at[i] = 0;
                }
            }
            if (++vt == a.n_vt) { vt = 0; gt++; }
        }
    }
    nvb_tl_stamp(a.tl, 0, 2);
}

// ---- few glimpses against a large library: the library streams through once --------------
// One agent's heading sweep (G <= 16 glimpses) against 10^5 .. 10^7 views (BASELINE C4, and every
// rank's slice of a view-sharded library): every library byte is read once and meets G glimpse
// bytes -- 12.4 us of memory traffic per 10^6 views of 80 bytes and, for G = 10, 10.7 us of
// VABSDIFF4 issue: both roofs at once.  k2_sad_v's CTA-wide tiles leave the SMs waiting there,
// and its 10 x 256 tile spends more shared-memory bandwidth on broadcasting glimpse chunks than
// the ALU pipe spends on SADs.  Here every WARP streams on its own: the library is cut into one
// contiguous range of views per warp, a warp's lane 0 keeps NVB_STREAM_STAGES bulk copies
// (cp.async.bulk, 128 views = 128 * 16 * CPR contiguous bytes each) in flight on per-warp
// mbarriers -- no CTA-wide barrier, no producer warp -- and each lane scores FOUR views (rows
// in registers, read from the stage with conflict-free LDS.128: odd CPR) against all glimpses
// (rows in shared memory; one broadcast LDS.128 of a glimpse chunk now feeds 16 SADs: with two
// views per lane the kernel was bound by shared-memory return bandwidth, 128 B/clk/SM, measured
// in tools/micro/stream_probe.cu).  Per-thread running minima in registers; one shuffle
// reduction and G atomicMin per CTA at the end.
// Same keys as k2_sad_v: (sum << idx_bits) | global view index, lowest index among equal sums.
#define NVB_STREAM_THREADS 128
#define NVB_STREAM_STAGES 2
#define NVB_STREAM_VPL 4     /* views per lane and chunk: one broadcast read of a glimpse chunk serves them all */
#define NVB_STREAM_VPC (32 * NVB_STREAM_VPL)   /* views per chunk */
#define NVB_STREAM_VBITS 17 /* a warp's range of views: index bits in the packed per-thread key */

__host__ __device__ inline int nvb_stream_smem(int Ppad, int gmax)
{
    return (NVB_STREAM_THREADS / 32) * NVB_STREAM_STAGES * NVB_STREAM_VPC * Ppad + gmax * Ppad +
           (NVB_STREAM_THREADS / 32) * gmax * 8 + (NVB_STREAM_THREADS / 32) * NVB_STREAM_STAGES * 8 + 128;
}

// EXACT: a.G == GMAX, no per-glimpse branch -- the SAD chains of different glimpses interleave (two
// chains per glimpse alone leave the ALU pipe a third idle: measured, tools/micro/stream_probe.cu)
template <int CPR, int GMAX, bool EXACT>
__global__ void __launch_bounds__(NVB_STREAM_THREADS, 2)
k2_stream(DistArgs a)
{
    constexpr int KC = 16 * CPR, NW = NVB_STREAM_THREADS / 32, ST = NVB_STREAM_STAGES, VPC = NVB_STREAM_VPC, VPL = NVB_STREAM_VPL;
    constexpr int CHUNK = VPC * KC;
    extern __shared__ __align__(128) uint8_t smem_st[];
    uint8_t *stage0 = smem_st;                                              // [NW][ST][VPC][KC]
    uint4 *s_q = reinterpret_cast<uint4 *>(smem_st + NW * ST * CHUNK);      // [GMAX][CPR]
    unsigned long long *s_keys = reinterpret_cast<unsigned long long *>(s_q + GMAX * CPR);   // [NW][GMAX]
    uint64_t *bars = reinterpret_cast<uint64_t *>(s_keys + NW * GMAX);      // [NW][ST]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint8_t *wstage = stage0 + (size_t)warp * ST * CHUNK;
    uint64_t *wbar = bars + warp * ST;

    nvb_tl_stamp(a.tl, 0, 0);
    if (a.pdl_early) nvb_grid_dep_launch();
    // this warp's views [v0, v1): equal shares of the library, to the view
    const long long n_warps = (long long)gridDim.x * NW, gw = (long long)blockIdx.x * NW + warp;
    const long long v0 = a.N * gw / n_warps, v1 = a.N * (gw + 1) / n_warps;
    const int n_chunks = (int)((v1 - v0 + VPC - 1) / VPC);
    auto issue = [&](int ch) {   // lane 0: chunk ch of this warp into its stage
        const long long c0 = v0 + (long long)ch * VPC;
        const uint32_t bytes = (uint32_t)(min((long long)VPC, v1 - c0) * KC);
        uint64_t *bar = wbar + ch % ST;
        nvb_mbar_expect_tx(bar, bytes);
        nvb_bulk_load_1d(wstage + (size_t)(ch % ST) * CHUNK, a.lv + (size_t)c0 * KC, bytes, bar);
    };
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < ST; s++) nvb_mbar_init(wbar + s, 1);
        nvb_fence_barrier_init();
        // the library is not written by any kernel of the step sequence: first chunks now
        for (int s = 0; s < ST && s < n_chunks; s++) issue(s);
    }
    nvb_grid_dep_wait();   // the glimpses are written by the previous kernel
    nvb_tl_stamp(a.tl, 0, 1);
    if (a.step_counter != nullptr && blockIdx.x == 0 && tid == 0) {
        *a.step_counter += 1;
        a.tie_count[0] = 0;
        a.tie_count[1] = 0;
        if (a.epoch != nullptr) *a.epoch += 1;
    }
    for (int q = tid; q < a.G * CPR; q += NVB_STREAM_THREADS) s_q[q] = reinterpret_cast<const uint4 *>(a.gv)[q];
    __syncthreads();

    // per-thread running minimum of (sum << NVB_STREAM_VBITS) | view index within this warp's range:
    // sums stay below 2^15 (rows of up to 128 bytes), a warp's range below 2^17 views (the host checks)
    uint32_t best[GMAX];
#pragma unroll
    for (int i = 0; i < GMAX; i++) best[i] = 0xFFFFFFFFu;

    // FULL: all 64 rows of the chunk are views of this warp's range (every chunk but a ragged last one)
    auto score = [&](int ch, auto full_tag) {
        constexpr bool FULL = decltype(full_tag)::value;
        nvb_mbar_wait(wbar + ch % ST, (uint32_t)((ch / ST) & 1));
        const uint8_t *st = wstage + (size_t)(ch % ST) * CHUNK;
        uint4 r[VPL][CPR];   // this lane's views: rows lane, lane + 32, ...
#pragma unroll
        for (int j = 0; j < VPL; j++)
#pragma unroll
            for (int c = 0; c < CPR; c++) r[j][c] = *reinterpret_cast<const uint4 *>(st + (size_t)(lane + 32 * j) * KC + 16 * c);
        // every lane must HOLD its rows (not merely have requested them) before the stage is
        // refilled through the async proxy: an instruction that consumes one register of every
        // 16-byte load waits for them
        uint32_t landed = 0;
#pragma unroll
        for (int j = 0; j < VPL; j++)
#pragma unroll
            for (int c = 0; c < CPR; c++) landed ^= r[j][c].x;
        asm volatile("" ::"r"(landed) : "memory");
        __syncwarp();
        if (lane == 0 && ch + ST < n_chunks) issue(ch + ST);
        const uint32_t lv0 = (uint32_t)(ch * VPC + lane);
        const int rows = FULL ? VPC : (int)(v1 - (v0 + (long long)ch * VPC));
#pragma unroll
        for (int i = 0; i < GMAX; i++) {
            if (EXACT || i < a.G) {
                uint32_t sum[VPL];
#pragma unroll
                for (int j = 0; j < VPL; j++) sum[j] = 0;
#pragma unroll
                for (int c = 0; c < CPR; c++) {
                    const uint4 q = s_q[i * CPR + c];   // one broadcast read serves VPL views
#pragma unroll
                    for (int j = 0; j < VPL; j++) sum[j] = nvb_sad4(q.x, r[j][c].x, sum[j]);
#pragma unroll
                    for (int j = 0; j < VPL; j++) sum[j] = nvb_sad4(q.y, r[j][c].y, sum[j]);
#pragma unroll
                    for (int j = 0; j < VPL; j++) sum[j] = nvb_sad4(q.z, r[j][c].z, sum[j]);
#pragma unroll
                    for (int j = 0; j < VPL; j++) sum[j] = nvb_sad4(q.w, r[j][c].w, sum[j]);
                }
                // keys on the FMA pipe (a true multiply-add), minima on the ALU pipe, which the SADs own
                uint32_t k[VPL];
#pragma unroll
                for (int j = 0; j < VPL; j++) {
                    k[j] = sum[j] * (1u << NVB_STREAM_VBITS) + (lv0 + 32u * j);
                    if (!FULL) k[j] = (lane + 32 * j < rows) ? k[j] : 0xFFFFFFFFu;
                }
                uint32_t m = best[i];
#pragma unroll
                for (int j = 0; j + 1 < VPL; j += 2) m = __vimin3_u32(m, k[j], k[j + 1]);
                if (VPL & 1) m = min(m, k[VPL - 1]);
                best[i] = m;
            }
        }
    };
    const int n_full = (int)((v1 - v0) / VPC);
    for (int ch = 0; ch < n_full; ch++) score(ch, std::true_type{});
    if (n_full < n_chunks) score(n_full, std::false_type{});

    // minimum over the lanes, then over the warps of the CTA, then one atomic per glimpse
#pragma unroll
    for (int i = 0; i < GMAX; i++) {
        if (EXACT || i < a.G) {
            uint32_t k = best[i];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) k = min(k, __shfl_xor_sync(0xFFFFFFFFu, k, o));
            if (lane == 0) {
                unsigned long long key = NVB_KEY_NONE;
                if (k != 0xFFFFFFFFu)
                    key = ((unsigned long long)(k >> NVB_STREAM_VBITS) << a.idx_bits) |
                          (unsigned long long)(a.view_offset + v0 + (long long)(k & ((1u << NVB_STREAM_VBITS) - 1u)));
                s_keys[warp * GMAX + i] = key;
            }
        }
    }
    __syncthreads();
    if (tid < a.G) {
        unsigned long long key = s_keys[tid];
#pragma unroll
        for (int w = 1; w < NW; w++) { const unsigned long long other = s_keys[w * GMAX + tid]; key = (other < key) ? other : key; }
        if (key != NVB_KEY_NONE) atomicMin(a.keys + tid, key);
    }
    nvb_tl_stamp(a.tl, 0, 2);
}

// ---- chem_weight > 0: hue/saturation branch + V, three planes ----------------
// X = (Hq == Hn) ? |Sq - Sn| : Sq + Sn   (util.pyx:48-56), V = |Vq - Vn| (:69);
// score = floor(4096 * (cw * 0.5 * sum X + (1 - cw) * sum V)): a fixed-point
// surrogate of 255 * diff; candidates within one quantum of the minimum are
// re-evaluated in exact FP64 by the tie pass.
__device__ __forceinline__ unsigned long long nvb_hsv_score(uint32_t xs, uint32_t vs, double cw)
{
    double f = __dadd_rn(__dmul_rn(__dmul_rn((double)xs, 0.5), cw),
                         __dmul_rn(__dsub_rn(1.0, cw), (double)vs));
    return (unsigned long long)(f * 4096.0);
}

__device__ __forceinline__ void nvb_hsv_word(uint32_t qh, uint32_t qs, uint32_t qv, uint32_t fh,
                                             uint32_t fs, uint32_t fv, uint32_t &xs, uint32_t &vs)
{
    const uint32_t eq = __vcmpeq4(qh, fh);          // 0xFF where hues match
    xs = nvb_sad4(qs & eq, fs & eq, xs);            // |Sq - Sn| where equal
    xs = nvb_sad4(qs & ~eq, 0u, xs);                // Sq + Sn where different
    xs = nvb_sad4(fs & ~eq, 0u, xs);
    vs = nvb_sad4(qv, fv, vs);
}

#define NVB_HSV_TG 8
#define NVB_HSV_THREADS 128

// one thread per view, NVB_HSV_TG glimpses per CTA held in shared memory
__global__ void __launch_bounds__(NVB_HSV_THREADS)
k2_sad_hsv(DistArgs a)
{
    extern __shared__ __align__(16) uint8_t smem_hsv[];   // [3][TG][Ppad]
    uint8_t *smem = smem_hsv;
    const int g0 = blockIdx.x * NVB_HSV_TG;
    const int tid = threadIdx.x;
    nvb_grid_dep_wait();
    if (a.step_counter != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && tid == 0) {
        *a.step_counter += 1;
        a.tie_count[0] = 0;   // list length
        a.tie_count[1] = 0;   // tie units completed (tie pass folded into move+sample)
        if (a.epoch != nullptr) *a.epoch += 1;
    }
    const int words = a.Ppad / 4;
    uint32_t *sm = reinterpret_cast<uint32_t *>(smem);
    for (int q = tid; q < 3 * NVB_HSV_TG * words; q += NVB_HSV_THREADS) {
        const int pl = q / (NVB_HSV_TG * words), r = (q / words) % NVB_HSV_TG, wd = q % words;
        const uint8_t *src = (pl == 0) ? a.gh : (pl == 1) ? a.gs : a.gv;
        const int g = g0 + r;
        sm[q] = (g < a.G) ? reinterpret_cast<const uint32_t *>(src + (size_t)g * a.Ppad)[wd] : 0u;
    }
    __syncthreads();
    const uint32_t *qh = sm, *qs = sm + NVB_HSV_TG * words, *qv = sm + 2 * NVB_HSV_TG * words;

    const int v0 = blockIdx.y * a.vt_per_split * NVB_HSV_THREADS;
    const int v1 = min(v0 + a.vt_per_split * NVB_HSV_THREADS, a.N);
    unsigned long long best[NVB_HSV_TG];
#pragma unroll
    for (int i = 0; i < NVB_HSV_TG; i++) best[i] = NVB_KEY_NONE;

    for (int v = v0 + tid; v < v1; v += NVB_HSV_THREADS) {
        uint32_t xs[NVB_HSV_TG], vs[NVB_HSV_TG];
#pragma unroll
        for (int i = 0; i < NVB_HSV_TG; i++) { xs[i] = 0; vs[i] = 0; }
        const uint32_t *fh = reinterpret_cast<const uint32_t *>(a.lh + (size_t)v * a.Ppad);
        const uint32_t *fs = reinterpret_cast<const uint32_t *>(a.ls + (size_t)v * a.Ppad);
        const uint32_t *fv = reinterpret_cast<const uint32_t *>(a.lv + (size_t)v * a.Ppad);
        for (int wd = 0; wd < words; wd++) {
            const uint32_t h = __ldg(fh + wd), s = __ldg(fs + wd), vv = __ldg(fv + wd);
#pragma unroll
            for (int i = 0; i < NVB_HSV_TG; i++)
                nvb_hsv_word(qh[i * words + wd], qs[i * words + wd], qv[i * words + wd], h, s, vv,
                             xs[i], vs[i]);
        }
#pragma unroll
        for (int i = 0; i < NVB_HSV_TG; i++) {
            const unsigned long long key = (nvb_hsv_score(xs[i], vs[i], a.cw) << a.idx_bits) |
                                           (unsigned long long)(a.view_offset + v);
            best[i] = (key < best[i]) ? key : best[i];
        }
    }
#pragma unroll
    for (int i = 0; i < NVB_HSV_TG; i++) {
        unsigned long long key = best[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            unsigned long long other = __shfl_xor_sync(0xFFFFFFFFu, key, o);
            key = (other < key) ? other : key;
        }
        const int g = g0 + i;
        if ((tid & 31) == 0 && g < a.G && key != NVB_KEY_NONE) atomicMin(a.keys + g, key);
    }
}

// ---- chem_weight > 0, tiled ------------------------------------------------------------------
// The same metric on register tiles, like k2_sad_v: a CTA owns a TG x TN (glimpse x view) tile,
// the three planes of both operands are staged per K-chunk by a cp.async ring (swizzled 16-byte
// chunks), every thread accumulates MG x MV pairs with two sums each (X = hue/saturation part,
// V part; nvb_hsv_word: one VCMP4, two LOP3 pairs, four VABSDIFF4 per word = 4 pixels), the
// fixed-point score and the packed key are formed once per unit.  Work is cut into contiguous
// unit spans per CTA exactly as in k2_sad_v.
template <int TY, int MG, int MV, int CPR, int STAGES>
struct HsvCfg {
    static constexpr int TX = NVB_DIST_THREADS / TY;
    static constexpr int TG = TY * MG;
    static constexpr int TN = TX * MV;
    static constexpr int KC = 16 * CPR;
    static constexpr int PLANE_BYTES = (TG + TN) * KC;
    static constexpr int STAGE_BYTES = 3 * PLANE_BYTES;
    static constexpr int SMEM = STAGE_BYTES * STAGES;
};

template <int TY, int MG, int MV, int CPR, int STAGES>
__global__ void __launch_bounds__(NVB_DIST_THREADS, 2)
k2_sad_hsv_t(DistArgs a)
{
    using C = HsvCfg<TY, MG, MV, CPR, STAGES>;
    constexpr int TX = C::TX, TG = C::TG, TN = C::TN, KC = C::KC;
    extern __shared__ __align__(128) uint8_t smem_hsv_t[];
    uint8_t *smem = smem_hsv_t;
    const int tid = threadIdx.x;
    const int tx = tid % TX, ty = tid / TX;
    if (a.pdl_early) nvb_grid_dep_launch();
    const int u0 = a.spans[blockIdx.x], u1 = a.spans[blockIdx.x + 1];
    const int nk = a.nk;
    const int total = (u1 - u0) * nk;
    nvb_grid_dep_wait();
    if (a.step_counter != nullptr && blockIdx.x == 0 && tid == 0) {
        *a.step_counter += 1;
        a.tie_count[0] = 0;
        a.tie_count[1] = 0;
        if (a.epoch != nullptr) *a.epoch += 1;
    }
    if (total <= 0) return;

    int l_gt = u0 / a.n_vt, l_vt = u0 - l_gt * a.n_vt, l_kc = 0, l_it = 0;
    auto load_next = [&]() {
        uint8_t *st = smem + (l_it % STAGES) * C::STAGE_BYTES;
        const int kbyte = l_kc * KC;
        for (int q = tid; q < 3 * (TG + TN) * CPR; q += NVB_DIST_THREADS) {
            const int pl = q / ((TG + TN) * CPR), qq = q - pl * (TG + TN) * CPR;
            const int row = qq / CPR, c = qq - row * CPR;
            const uint8_t *src;
            int ok;
            if (row < TG) {
                const int g = l_gt * TG + row;
                ok = (g < a.G) && (kbyte + 16 * c < a.Ppad);
                const uint8_t *base = (pl == 0) ? a.gh : (pl == 1) ? a.gs : a.gv;
                src = base + (size_t)(ok ? g : 0) * a.Ppad + (ok ? kbyte + 16 * c : 0);
            } else {
                const int v = l_vt * TN + (row - TG);
                ok = (v < a.N) && (kbyte + 16 * c < a.Ppad);
                const uint8_t *base = (pl == 0) ? a.lh : (pl == 1) ? a.ls : a.lv;
                src = base + (size_t)(ok ? v : 0) * a.Ppad + (ok ? kbyte + 16 * c : 0);
            }
            const int r = (row < TG) ? row : row - TG;
            uint8_t *dst = st + pl * C::PLANE_BYTES + (row < TG ? 0 : TG * KC) + r * KC + 16 * (c ^ nvb_swz<CPR>(r));
            nvb_cp_async16(dst, src, ok ? 16 : 0);
        }
        l_it++;
        if (++l_kc == nk) {
            l_kc = 0;
            if (++l_vt == a.n_vt) { l_vt = 0; l_gt++; }
        }
    };

    uint32_t xs[MG][MV], vs[MG][MV];
#pragma unroll
    for (int i = 0; i < MG; i++)
#pragma unroll
        for (int j = 0; j < MV; j++) { xs[i][j] = 0; vs[i][j] = 0; }
    unsigned long long best[MG];
#pragma unroll
    for (int i = 0; i < MG; i++) best[i] = NVB_KEY_NONE;

    int goff[MG], gsw[MG], voff[MV], vsw[MV];
#pragma unroll
    for (int i = 0; i < MG; i++) { int r = ty + TY * i; goff[i] = r * KC; gsw[i] = nvb_swz<CPR>(r) << 4; }
#pragma unroll
    for (int j = 0; j < MV; j++) { int r = tx + TX * j; voff[j] = TG * KC + r * KC; vsw[j] = nvb_swz<CPR>(r) << 4; }

#pragma unroll
    for (int s = 0; s < STAGES - 1; s++) {
        if (s < total) load_next();
        nvb_cp_async_commit();
    }
    constexpr int RW = (TX < 32) ? TX : 32;
    int gt = u0 / a.n_vt, vt = u0 - gt * a.n_vt, kc = 0;
    for (int it = 0; it < total; it++) {
        nvb_cp_async_wait<STAGES - 2>();
        __syncthreads();
        if (l_it < total) load_next();
        nvb_cp_async_commit();
        const uint8_t *st = smem + (it % STAGES) * C::STAGE_BYTES;
        for (int c = 0; c < CPR; c++) {
            uint4 qh[MG], qs[MG], qv[MG];
#pragma unroll
            for (int i = 0; i < MG; i++) {
                const int o = goff[i] + ((c << 4) ^ gsw[i]);
                qh[i] = *reinterpret_cast<const uint4 *>(st + o);
                qs[i] = *reinterpret_cast<const uint4 *>(st + C::PLANE_BYTES + o);
                qv[i] = *reinterpret_cast<const uint4 *>(st + 2 * C::PLANE_BYTES + o);
            }
#pragma unroll
            for (int j = 0; j < MV; j++) {
                const int o = voff[j] + ((c << 4) ^ vsw[j]);
                const uint4 fh = *reinterpret_cast<const uint4 *>(st + o);
                const uint4 fs = *reinterpret_cast<const uint4 *>(st + C::PLANE_BYTES + o);
                const uint4 fv = *reinterpret_cast<const uint4 *>(st + 2 * C::PLANE_BYTES + o);
#pragma unroll
                for (int i = 0; i < MG; i++) {
                    nvb_hsv_word(qh[i].x, qs[i].x, qv[i].x, fh.x, fs.x, fv.x, xs[i][j], vs[i][j]);
                    nvb_hsv_word(qh[i].y, qs[i].y, qv[i].y, fh.y, fs.y, fv.y, xs[i][j], vs[i][j]);
                    nvb_hsv_word(qh[i].z, qs[i].z, qv[i].z, fh.z, fs.z, fv.z, xs[i][j], vs[i][j]);
                    nvb_hsv_word(qh[i].w, qs[i].w, qv[i].w, fh.w, fs.w, fv.w, xs[i][j], vs[i][j]);
                }
            }
        }
        if (++kc == nk) {
            kc = 0;
            const int nvalid = a.N - vt * TN;
#pragma unroll
            for (int i = 0; i < MG; i++)
#pragma unroll
                for (int j = 0; j < MV; j++) {
                    const int vloc = tx + TX * j;
                    if (vloc < nvalid) {
                        const unsigned long long key = (nvb_hsv_score(xs[i][j], vs[i][j], a.cw) << a.idx_bits) |
                                                       (unsigned long long)(a.view_offset + (long long)vt * TN + vloc);
                        best[i] = key < best[i] ? key : best[i];
                    }
                    xs[i][j] = 0;
                    vs[i][j] = 0;
                }
            if (it == total - 1 || vt == a.n_vt - 1) {
#pragma unroll
                for (int i = 0; i < MG; i++) {
                    unsigned long long key = best[i];
#pragma unroll
                    for (int o = RW / 2; o > 0; o >>= 1) {
                        const unsigned long long other = __shfl_xor_sync(0xFFFFFFFFu, key, o);
                        key = other < key ? other : key;
                    }
                    const int g = gt * TG + ty + TY * i;
                    if ((tx % RW) == 0 && g < a.G && key != NVB_KEY_NONE) atomicMin(a.keys + g, key);
                    best[i] = NVB_KEY_NONE;
                }
            }
            if (++vt == a.n_vt) { vt = 0; gt++; }
        }
    }
}

// ---- exact FP64 difference of one (glimpse, view) pair -----------------------
// The literal operation sequence of util.pyx:48-72 (no FMA contraction), so
// that fam = H*W - diff is bit-identical to the reference.  q*/f* are rows of
// the planar glimpse / library arrays; div255 = {k / 255.} for k in 0..255.
__device__ __forceinline__ double nvb_exact_diff(const uint8_t *qh, const uint8_t *qs,
                                                 const uint8_t *qv, const uint8_t *fh,
                                                 const uint8_t *fs, const uint8_t *fv, int P,
                                                 double cw, const double *div255)
{
    double diff = 0.0;
    if (cw == 0.0) {
        // thispx = X*0.5*0 (= +0) + 1*|dV|, then / 255.  (util.pyx:59-72)
        for (int p = 0; p < P; p++) {
            int d = (int)qv[p] - (int)fv[p];
            d = d < 0 ? -d : d;
            diff = __dadd_rn(diff, div255[d]);
        }
    } else {
        const double omc = __dsub_rn(1.0, cw);
        for (int p = 0; p < P; p++) {
            int sq = qs[p], sn = fs[p];
            int x = (qh[p] == fh[p]) ? (sq > sn ? sq - sn : sn - sq) : (sq + sn);
            int d = (int)qv[p] - (int)fv[p];
            d = d < 0 ? -d : d;
            double t = (double)x;
            t = __dmul_rn(t, 0.5);
            t = __dmul_rn(t, cw);
            t = __dadd_rn(t, __dmul_rn(omc, (double)d));
            t = __ddiv_rn(t, 255.0);
            diff = __dadd_rn(diff, t);
        }
    }
    return diff;
}

// A5 for the single-call API: fam [G][N] = H*W - diff, one thread per pair.
__global__ void k_familiarity_exact(DistArgs a, int P, double maxfam, const double *div255,
                                    double *fam)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)a.G * a.N) return;
    const int g = (int)(i / a.N), n = (int)(i - (long long)g * a.N);
    const size_t qo = (size_t)g * a.Ppad, fo = (size_t)n * a.Ppad;
    const double d = nvb_exact_diff(a.gh + qo, a.gs + qo, a.gv + qo, a.lh + fo, a.ls + fo,
                                    a.lv + fo, P, a.cw, div255);
    fam[i] = __dsub_rn(maxfam, d);
}

// Register-resident VABSDIFF4+accumulate issue-rate probe (bench.py's ALU
// roofline denominator).  Each thread runs `iters` x 32 dependent-free SADs.
__global__ void k_probe_sad(int iters, uint32_t seed, uint32_t *sink)
{
    uint32_t acc[8], a[8], b[4];
#pragma unroll
    for (int i = 0; i < 8; i++) { acc[i] = i; a[i] = seed * (i + 1) + threadIdx.x; }
#pragma unroll
    for (int r = 0; r < 4; r++) b[r] = seed * 3u + blockIdx.x + r * 0x01010101u;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int i = 0; i < 8; i++) acc[i] = nvb_sad4(a[i], b[r], acc[i]);
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += acc[i];
    if (s == 0xDEADBEEFu) sink[0] = s;
}
