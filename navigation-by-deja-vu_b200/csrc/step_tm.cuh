// step_tm.cuh -- the whole step after the tensor-core distance kernel in ONE launch.
//
// Replaces, like step.cuh, the tail of step_forward (navsim/NavBySceneFamiliarity.py:313-329)
// and update_error (:252-276), and samples the glimpses of the NEXT step (sampler.cuh) in the
// same launch: K2 (k2_tc, TILEMIN) | k3_step_tm -- two launches per step-batch instead of
// K2 | decide | ties | move+sample.
//
// What makes one launch enough: k2_tc leaves, per glimpse and view tile, the tile's smallest
// key (-256 * dot + column).  For an agent whose headings do NOT tie at the step's integer
// minimum -- 99.2 % of agent-steps on the bench workload -- the heading with the smallest
// integer difference IS the reference's argmax over doubles (two integer sums that differ by
// at least 1 give FP64 sums that differ by 1/255, far above their rounding), so the move
// needs no FP64 value at all: warp 0 picks the heading and moves while the LAST warp computes
// the exact FP64 differences that only the log wants (angle_familiarity, step_familiarity),
// off the critical path.  For an agent WITH ties every view that attains the
// minimum must be compared in FP64: those views can only sit in tiles whose tile minimum equals
// the step's minimum, behind the tile's best column, so the CTA rescans just those tiles (a
// few hundred views instead of the library) and then takes the reference's first maximum.
//
// update_error: the bounding-circle prefilter of step.cuh evaluated in FP32 on a float copy of
// the circles with outward-rounded bounds (a filter only has to be conservative); the
// surviving blocks are scanned with exactly the FP64 arithmetic of the full scan, so the
// minimum and the coverage marks are unchanged.
//
// Applies when: chem_weight == 0, tensor-core distance kernel, sweeps of up to 32 headings, at
// most NVB_TM_MAX_VT view tiles, sensors of up to NVB_PTAB_MAX pixels, staged window (R > 0),
// single-level path blocks, coverage threshold <= max distance, no view shards.  Everything
// else runs the launches of step.cuh.
#pragma once
#include "step.cuh"

#define NVB_TM_MAX_A 32
#define NVB_TM_MAX_VT 64
#define NVB_TM_LIVE_MAX 192     /* path blocks the prefilter can list (more: full scan) */
#define NVB_TM_NONE 0x7FFFFFFFFFFFFFFFll
#define NVB_TM_MAX_JOBS 64      /* (tied heading, tied tile) pairs rescanned by the whole CTA at once */

// All samples of one sensor-pixel block with the reference's FP64 expression (the lean gather
// calls this for the rare block that has a sample within the guard band of a rounding tie).
// The agent is `safe` (sampler.cuh): every index lies inside the landscape and the window.
__device__ __noinline__ int nvb_block_exact_sum(const uint8_t *win_v, int BW, int ox, int oy, double x, double y,
                                                double c, double s, double half_w, double half_h, int col0,
                                                int row0, int ph, int pw)
{
    int sum = 0;
    for (int i = 0; i < ph; i++)
        for (int j = 0; j < pw; j++) {
            const double px = (double)(col0 + j) - half_w;   // util.pyx:159
            const double py = (double)(row0 + i) - half_h;   // util.pyx:160
            const double rx = __dsub_rn(__dmul_rn(px, c), __dmul_rn(py, s));   // :161
            const double ry = __dadd_rn(__dmul_rn(px, s), __dmul_rn(py, c));   // :162
            const int iy = (int)round(__dadd_rn(ry, y));                       // :166
            const int ix = (int)round(__dadd_rn(rx, x));                       // :167
            sum += win_v[(iy - oy) * BW + (ix - ox)];
        }
    return sum;
}

// Lean gather of a `safe` agent, V plane only: every sensor pixel of every heading gathers
// its PH x PW block from the staged window (coordinates in packed FP32x2 relative to
// (floor x, floor y), add-magic-number rounding, see sampler.cuh), block mean, quantise, mask,
// V plane + thermometer bytes out.  Same results as nvb_sample_gather<false, PH, PW>.
template <int PH, int PW>
__device__ __forceinline__ void nvb_gather_lean(const SamplerArgs &a, int b, double x, double y, const SamplerSmem &L)
{
    const NvbWorld &w = a.w;
    const int tid = threadIdx.x, T = (int)blockDim.x;
    const int ph = PH ? PH : w.ph, pw = PW ? PW : w.pw, nblk = ph * pw;
    const int xi = __double2int_rd(x), yi = __double2int_rd(y);
    const int ox = (xi - w.R) & ~15, oy = yi - w.R;
    const float xf = (float)(x - (double)xi), yf = (float)(y - (double)yi);
    const float tie = 0.5f - a.band;
    // shared-window address of sample (ux, uy) = sbase + uy * BW + ux, the magic-number bias of
    // both float bit patterns folded in (32-bit wrap-around arithmetic is exact)
    const uint32_t BW = (uint32_t)w.BW;
    const uint32_t sbase = nvb_smem_u32(L.win_v) + (uint32_t)(((yi - oy) - NVB_RND_MAGIC_BITS) * w.BW + ((xi - ox) - NVB_RND_MAGIC_BITS));
    const float half_wf = 0.5f * (float)w.Wpx, half_hf = 0.5f * (float)w.Hpx;
    const float mask_lo = (float)(w.mask_lo * pw) - half_wf, mask_hi = (float)(w.mask_hi * pw) - half_wf;
    const int P = w.P, n_items = a.A * P;
    const int dk = T / P, dp = T - dk * P;
    int k = tid / P, p = tid - k * P;
    const bool linear_out = (w.Ppad == P);
    uint8_t *gv = a.gv + (size_t)b * a.A * w.Ppad;
    int8_t *genc = a.genc + (size_t)b * a.A * a.Kpad;
    const float2 magic = make_float2(NVB_RND_MAGIC, NVB_RND_MAGIC);
    const float2 neg_magic = make_float2(-NVB_RND_MAGIC, -NVB_RND_MAGIC), neg_one = make_float2(-1.0f, -1.0f);

    for (int it = tid; it < n_items; it += T) {
        const float2 o = L.ptab[p];      // (px0, py0) of the block's first sample
        const float2 cs = L.csf[k];      // (cos, sin) of heading k
        const float2 t0 = make_float2(fmaf(o.x, cs.x, fmaf(-o.y, cs.y, xf)), fmaf(o.x, cs.y, fmaf(o.y, cs.x, yf)));
        const float2 step_j = cs, step_i = make_float2(-cs.y, cs.x);
        float2 t_row = t0;
        float worst = 0.0f;
        int sum = 0;
#pragma unroll
        for (int i = 0; i < ph; i++) {
            float2 t = t_row;
#pragma unroll
            for (int j = 0; j < pw; j++) {
                const float2 u = __fadd2_rn(t, magic);                                  // low mantissa bits = round(t)
                const float2 e = __ffma2_rn(__fadd2_rn(u, neg_magic), neg_one, t);     // t - round(t)
                worst = fmaxf(worst, fmaxf(fabsf(e.x), fabsf(e.y)));
                uint32_t v;
                asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"((uint32_t)__float_as_int(u.y) * BW + (uint32_t)__float_as_int(u.x) + sbase));
                sum += (int)v;
                if (j + 1 < pw) t = __fadd2_rn(t, step_j);
            }
            if (i + 1 < ph) t_row = __fadd2_rn(t_row, step_i);
        }
        if (worst >= tie)   // about one block in a thousand
            sum = nvb_block_exact_sum(L.win_v, w.BW, ox, oy, x, y, L.cs[2 * k], L.cs[2 * k + 1], 0.5 * (double)w.Wpx,
                                      0.5 * (double)w.Hpx, (int)(o.x + half_wf), (int)(o.y + half_hf), ph, pw);
        // util.pyx:121-123: V = (uint8) round(sum / (fr*fc)), half away from zero, in integers
        const int v = (2 * sum + nblk) / (2 * nblk);
        const bool masked = (o.x >= mask_lo && o.x < mask_hi);   // NavBySceneFamiliarity.py:189-190
        gv[linear_out ? it : k * w.Ppad + p] = masked ? (uint8_t)0 : L.lut[512 + v];
        const int lvl = masked ? 0 : (int)L.tc[v];
        const uint8_t *enc = L.tc + 256 + lvl * 8;
        int8_t *eo = genc + (size_t)k * a.Kpad + (size_t)p * a.n_planes;
        if (a.n_planes == 4) {
            *reinterpret_cast<uint32_t *>(eo) = *reinterpret_cast<const uint32_t *>(enc);
        } else {
            for (int q = 0; q < a.n_planes; q++) eo[q] = (int8_t)enc[q];
        }
        k += dk;
        p += dp;
        if (p >= P) { p -= P; k++; }
    }
}

// nvb_pymod_pos(a, 2 pi) without the iterative fmod for the arguments the loop produces: the
// heading stays in [0, 2 pi) and an offset is at most pi, so a lies in (-2 pi, 4 pi), where
// fmod(a, 2 pi) is a itself or a - 2 pi -- exact by Sterbenz's lemma, as fmod is by definition.
__device__ __forceinline__ double nvb_pymod_2pi(double a)
{
    const double b = NVB_TWO_PI;
    if (a >= 0.0 && a < b) return a == 0.0 ? 0.0 : a;      // fmod(a, b) == a (a zero comes back as +0.0)
    if (a >= b && a < 2.0 * b) return __dsub_rn(a, b);    // exact
    if (a < 0.0 && a > -b) return __dadd_rn(a, b);        // fmod keeps a, Python folds it up
    return nvb_pymod_pos(a, b);
}

// Exact FP64 difference (bit pattern) of a glimpse row and a view row, V plane, in the
// reference's pixel order (util.pyx:59-73) -- as nvb_exact_bits (step.cuh), but with up to five
// 16-byte chunks of both rows requested before the dependent add chain starts: one round trip
// per row instead of one per chunk.
__device__ __forceinline__ unsigned long long nvb_exact_bits_batched(const uint4 *q4, const uint4 *f4, int nc, const double *div255)
{
    double diff = 0.0;
    for (int c0 = 0; c0 < nc; c0 += 5) {
        uint32_t d[5][4];
#pragma unroll
        for (int u = 0; u < 5; u++)
            if (c0 + u < nc) {
                const uint4 qq = q4[c0 + u], ff = __ldg(f4 + c0 + u);
                d[u][0] = __vabsdiffu4(qq.x, ff.x); d[u][1] = __vabsdiffu4(qq.y, ff.y);
                d[u][2] = __vabsdiffu4(qq.z, ff.z); d[u][3] = __vabsdiffu4(qq.w, ff.w);
            }
#pragma unroll
        for (int u = 0; u < 5; u++)
            if (c0 + u < nc) {
#pragma unroll
                for (int w = 0; w < 4; w++)
#pragma unroll
                    for (int k = 0; k < 4; k++) diff = __dadd_rn(diff, div255[(d[u][w] >> (8 * k)) & 0xFFu]);
            }
    }
    return (unsigned long long)__double_as_longlong(diff);
}

// CTA-wide OR of `pred` over the first `n` threads (named barrier 3; the last warp is elsewhere)
__device__ __forceinline__ bool nvb_or_front(bool pred, int n, int *flag)
{
    if (pred) *flag = 1;
    asm volatile("bar.sync 3, %0;" ::"r"(n) : "memory");
    return *flag != 0;
}

__device__ __forceinline__ double nvb_warp_fmin(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xFFFFFFFFu, v, o));
    return v;
}

template <int PH, int PW>
__global__ void __launch_bounds__(NVB_MS_MAX_THREADS, 7)   // 1024 agents = 7 CTAs per SM: one wave
k3_step_tm(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ StepArgs a,
           const __grid_constant__ SamplerArgs sa)
{
    nvb_tl_stamp(a.tl, 3, 0);
    if (a.pdl_early) nvb_grid_dep_launch();
    __shared__ unsigned long long s_exact[NVB_TM_MAX_A]; // per heading: exact FP64 difference (bits)
    __shared__ double s_div[256];
    __shared__ double s_pose[3], s_red[8];
    __shared__ float s_ub[8];
    __shared__ int s_live[NVB_TM_LIVE_MAX];
    __shared__ int s_nlive, s_active, s_more, s_njobs, s_any;
    __shared__ int s_jobs[NVB_TM_MAX_JOBS];                // tie path: indices of (heading, tile) pairs to rescan
    extern __shared__ __align__(128) uint8_t smem_tm[];
    const SamplerSmem L = nvb_sampler_layout<false>(sa.w, sa.A, smem_tm);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nw = (int)(blockDim.x >> 5);
    const int b = blockIdx.x, A = a.A;
    const int n_blk = (a.n_path + NVB_PATH_BLOCK - 1) / NVB_PATH_BLOCK;

    // ---- prologue (constant data only; overlaps the distance kernel's tail): every load is
    // requested before the first store, so the whole prologue costs one round trip to L2
    {
        const uint32_t *lut_g = reinterpret_cast<const uint32_t *>(sa.w.lut + 512);   // V quantisation table
        const uint32_t *tc_g = reinterpret_cast<const uint32_t *>(sa.tc_tab);
        uint32_t r_lut = 0, r_tc = 0;
        double r_off = 0.0, r_div0 = 0.0, r_div1 = 0.0;
        if (tid < 64) r_lut = __ldg(lut_g + tid);
        if (tid < NVB_TC_TAB_BYTES / 4) r_tc = __ldg(tc_g + tid);
        if (tid < A) r_off = __ldg(a.offsets + tid);
        if (tid < 128) { r_div0 = __ldg(a.div255 + tid); r_div1 = __ldg(a.div255 + 128 + tid); }
        // into L1, where the dependent loads of the move find them: the block circles and (short
        // paths) the path itself for update_error, glibc's sin / cos table for the rotations
        for (int o = tid * 128; o < n_blk * 16; o += blockDim.x * 128)
            asm volatile("prefetch.global.L1 [%0];" ::"l"(reinterpret_cast<const char *>(a.pblk_f) + o));
        if (a.n_path <= 4096)
            for (int o = tid * 128; o < a.n_path * 16; o += blockDim.x * 128)
                asm volatile("prefetch.global.L1 [%0];" ::"l"(reinterpret_cast<const char *>(a.path) + o));
        for (int o = tid * 128; o < (int)sizeof(nvb_sincos_tab_dev); o += blockDim.x * 128)
            asm volatile("prefetch.global.L1 [%0];" ::"l"(reinterpret_cast<const char *>(nvb_sincos_tab_dev) + o));
        nvb_sampler_stage_ptab(sa.w, L.ptab);
        if (tid < NVB_TM_MAX_A) s_exact[tid] = NVB_EXACT_NONE;
        if (tid == 0) { s_nlive = 0; s_njobs = 0; s_any = 0; }
        if (tid < 64) reinterpret_cast<uint32_t *>(L.lut + 512)[tid] = r_lut;
        if (tid < NVB_TC_TAB_BYTES / 4) reinterpret_cast<uint32_t *>(L.tc)[tid] = r_tc;
        if (tid < A) L.offs[tid] = r_off;
        if (tid < 128) { s_div[tid] = r_div0; s_div[128 + tid] = r_div1; }
    }
    __syncthreads();
    nvb_grid_dep_wait();
    nvb_tl_stamp(a.tl, 3, 1);

    // ---- one round trip: the tile minima of this agent's headings, its state, the step index
    const int n_keys = A * a.n_vt;
    const int2 *tm = a.tmin + (size_t)b * n_keys;
    int2 *s_tk = reinterpret_cast<int2 *>(smem_tm + nvb_round_up((int)nvb_sampler_smem(sa.w.BW, sa.w.BH, 1, A), 16));   // [A][n_vt]
    const int t = *a.step_counter;
    int completed = 0, budget = 0, nav_frames = 0, err_n = 0;
    double err_sum = 0.0, px = 0.0, py = 0.0, ang0 = 0.0;
    double2 path_end = make_double2(0.0, 0.0);
    if (tid == 0) {
        path_end = __ldg(reinterpret_cast<const double2 *>(a.path) + (a.n_path - 1));
        const int pf = (a.pending_fail != nullptr) ? a.pending_fail[b] : 0;
        const int status = a.ag.status[b];
        completed = a.ag.completed[b];
        budget = a.ag.budget[b];
        nav_frames = a.ag.nav_frames[b];
        err_n = a.ag.err_n[b];
        err_sum = a.ag.err_sum[b];
        px = a.ag.poses[3 * b]; py = a.ag.poses[3 * b + 1]; ang0 = a.ag.poses[3 * b + 2];
        // a failure the sampler found for THIS step while it ran at the end of the previous
        // launch becomes the agent's status now (as k3_decide)
        if (pf != 0) {
            a.ag.status[b] = pf;
            a.pending_fail[b] = 0;
        }
        s_active = (pf == 0) && status == 0 && completed < budget;
    }
    for (int i = tid; i < n_keys; i += blockDim.x) {
        s_tk[i] = __ldcg(tm + i);
    }
    __syncthreads();
    nvb_tl_stamp(a.tl, 4, 0);   // (tuning aid: timeline slots 4 and 5 hold this kernel's checkpoints)
    if (!s_active) {
        nvb_log_idle(a, b);
        return;
    }
    const bool logging = (t >= 0 && t < a.log_cap);

    // ---- every warp: per heading (lane) the minimum over the view tiles and its first view (tiles
    // are in view order and a tile's smallest key has the lowest column: the lowest view index, as
    // the packed-key minimum gives it), then the step's integer minimum and the headings that attain it
    long long hk = NVB_TM_NONE;   // (-dot << 32) | best view
    if (lane < A)
        for (int tt = 0; tt < a.n_vt; tt++) {
            const int key = s_tk[lane * a.n_vt + tt].x;
            if (key != 0x7FFFFFFF) hk = min(hk, ((long long)(key >> 8) << 32) | (long long)(unsigned)(tt * NVB_TC_NT + (key & 255)));
        }
    const int negdot = (int)(hk >> 32);
    int M = negdot;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) M = min(M, __shfl_xor_sync(0xFFFFFFFFu, M, o));
    const unsigned tied = __ballot_sync(0xFFFFFFFFu, lane < A && negdot == M);
    int best = __ffs(tied) - 1;
    const bool ex_warp = (warp == nw - 1);
    const int nc = a.Ppad / 16;
    unsigned long long eb = NVB_EXACT_NONE;
    if (ex_warp) {
        // exact FP64 difference of every heading's best view (util.pyx:59-73 operation order)
        if (lane < A && hk != NVB_TM_NONE)
            eb = nvb_exact_bits_batched(reinterpret_cast<const uint4 *>(a.gv + ((size_t)b * A + lane) * a.Ppad),
                                        reinterpret_cast<const uint4 *>(a.lv + (size_t)(unsigned)(hk & 0xFFFFFFFFll) * a.Ppad), nc, s_div);
        if (eb != NVB_EXACT_NONE) atomicMin(&s_exact[lane], eb);   // (the tie path lowers these concurrently)
    }
    if (__popc(tied) > 1) {
        // ---- ties: FP64 over EVERY view that attains the minimum, then the
        // reference's first maximum (:313-315).  CTA-uniform branch.
        // The last warp goes on with the exact difference of every heading's best view (above); the
        // other warps, meanwhile:
        if (!ex_warp) {
            const int n_front = (nw - 1) * 32;
            // the (tied heading, tile) pairs whose tile minimum is the step's minimum -- only they can
            // hold more views at the minimum -- and this agent's glimpse rows into shared memory (the
            // window region is free until the pose is known)
            const bool q_sm = A * a.Ppad <= nvb_round_up(sa.w.BW * sa.w.BH, 128);
            for (int i = tid; i < n_keys; i += n_front) {
                const int key = s_tk[i].x;
                if (key != 0x7FFFFFFF && (key >> 8) == M && ((tied >> (i / a.n_vt)) & 1u)) {
                    const int slot = atomicAdd(&s_njobs, 1);
                    if (slot < NVB_TM_MAX_JOBS) s_jobs[slot] = i;
                }
            }
            if (q_sm)
                for (int c = tid; c < A * nc; c += n_front)
                    reinterpret_cast<uint4 *>(L.win_v)[c] = reinterpret_cast<const uint4 *>(a.gv + (size_t)b * A * a.Ppad)[c];
            asm volatile("bar.sync 3, %0;" ::"r"(n_front) : "memory");
            const unsigned sad_m = (unsigned)((a.sad_const + M) >> 1);
            const int n_jobs = min(s_njobs, NVB_TM_MAX_JOBS);
            const bool overflow = s_njobs > NVB_TM_MAX_JOBS;
            auto q_row = [&](int k) { return q_sm ? L.win_v + (size_t)k * a.Ppad : a.gv + ((size_t)b * A + k) * a.Ppad; };
            // the distance kernel kept the TWO best views of every tile: two threads per pair, one
            // per candidate, evaluate them in FP64.  Only a tile whose runner-up ties as well may
            // hold a third view at the minimum; such tiles (rare) are rescanned behind the runner-up.
            bool dense = false;
            for (int j = tid; j < 2 * n_jobs; j += n_front) {
                const int i = s_jobs[j >> 1], k = i / a.n_vt, tt = i - k * a.n_vt;
                const int2 kk = s_tk[i];
                const int key = (j & 1) ? kk.y : kk.x;
                if (key != 0x7FFFFFFF && (key >> 8) == M) {
                    const uint8_t *frow = a.lv + (size_t)(tt * NVB_TC_NT + (key & 255)) * a.Ppad;
                    atomicMin(&s_exact[k], nvb_exact_bits_batched(reinterpret_cast<const uint4 *>(q_row(k)),
                                                                  reinterpret_cast<const uint4 *>(frow), nc, s_div));
                    if (j & 1) dense = true;
                }
            }
            if (nvb_or_front(dense, n_front, &s_any) || overflow) {
                // one (pair, column) at a time per thread, behind the runner-up's column: the view row
                // in one round trip (up to five 16-byte chunks in flight), the glimpse row from shared memory
                auto rescan = [&](int i, int col) {
                    const int k = i / a.n_vt, tt = i - k * a.n_vt;
                    const int v = tt * NVB_TC_NT + col;
                    const int2 kk = s_tk[i];
                    if (kk.y == 0x7FFFFFFF || (kk.y >> 8) != M || col <= (kk.y & 255) || v >= a.N) return;
                    const uint4 *q4 = reinterpret_cast<const uint4 *>(q_row(k));
                    const uint4 *f4 = reinterpret_cast<const uint4 *>(a.lv + (size_t)v * a.Ppad);
                    uint32_t sum = 0;
                    for (int c0 = 0; c0 < nc; c0 += 5) {
                        uint4 ff[5];
#pragma unroll
                        for (int u = 0; u < 5; u++)
                            if (c0 + u < nc) ff[u] = __ldg(f4 + c0 + u);
#pragma unroll
                        for (int u = 0; u < 5; u++)
                            if (c0 + u < nc) {
                                const uint4 qq = q4[c0 + u];
                                sum = nvb_sad4(qq.x, ff[u].x, sum); sum = nvb_sad4(qq.y, ff[u].y, sum);
                                sum = nvb_sad4(qq.z, ff[u].z, sum); sum = nvb_sad4(qq.w, ff[u].w, sum);
                            }
                    }
                    if (sum == sad_m) atomicMin(&s_exact[k], nvb_exact_bits_batched(q4, f4, nc, s_div));
                };
                for (int idx = tid; idx < n_jobs * NVB_TC_NT; idx += n_front) rescan(s_jobs[idx / NVB_TC_NT], idx % NVB_TC_NT);
                if (overflow) {   // more pairs than the list holds: every candidate and column of all of them
                    for (int i = 0; i < n_keys; i++) {
                        const int2 kk = s_tk[i];
                        const int k = i / a.n_vt, tt = i - k * a.n_vt;
                        if (kk.x == 0x7FFFFFFF || (kk.x >> 8) != M || !((tied >> k) & 1u)) continue;
                        for (int col = tid; col < NVB_TC_NT && tt * NVB_TC_NT + col < a.N; col += n_front) {
                            const uint8_t *frow = a.lv + (size_t)(tt * NVB_TC_NT + col) * a.Ppad;
                            if (nvb_pair_score_v(q_row(k), frow, nc) == sad_m)
                                atomicMin(&s_exact[k], nvb_exact_bits(q_row(k), frow, nc, s_div));
                        }
                    }
                }
            }
        }
        __syncthreads();
        eb = s_exact[lane];
        const double fam = __dsub_rn(a.maxfam, __longlong_as_double((long long)eb));
        double best_fam = 0.0;
        for (int k = 0; k < A; k++) {
            const double fk = __shfl_sync(0xFFFFFFFFu, fam, k);
            if (k == 0 || fk > best_fam) { best = k; best_fam = fk; }
        }
    }
    if (ex_warp) {
        // the log's familiarities (util.pyx:73, NavBySceneFamiliarity.py:313), off the critical path
        const double fam = __dsub_rn(a.maxfam, __longlong_as_double((long long)eb));
        if (logging && a.log_afam != nullptr && lane < A) a.log_afam[((size_t)t * a.B + b) * A + lane] = fam;
        if (lane == best) {
            if (logging) a.log_sfam[(size_t)t * a.B + b] = fam;
            if (a.out_sfam != nullptr) a.out_sfam[b] = fam;
        }
    } else {
        // ---- move (:317-323)
        // (np.cos / np.sin = glibc's, :319-320; lane 1 of warp 0 computes the sine beside lane 0's cosine)
        double ang = 0.0, cs = 0.0, sn = 0.0;
        if (warp == 0) {
            ang = nvb_pymod_2pi(__dadd_rn(__shfl_sync(0xFFFFFFFFu, ang0, 0), L.offs[best]));
            double trig = 0.0;
            if (lane < 2) {
                if (nvb_trig::hi_abs(ang) >= NVB_TRIG_MAX_K) {   // (never: the angle is reduced mod 2 pi)
                    double s2, c2;
                    sincos(ang, &s2, &c2);
                    trig = lane ? s2 : c2;
                } else {
                    trig = lane ? nvb_trig::sin_(ang) : nvb_trig::cos_(ang);
                }
            }
            cs = trig;
            sn = __shfl_sync(0xFFFFFFFFu, trig, 1);
        }
        if (tid == 0) {
            const double x = __dadd_rn(px, __dmul_rn(a.step_size, cs));
            const double y = __dadd_rn(py, __dmul_rn(a.step_size, sn));
            s_pose[0] = x; s_pose[1] = y; s_pose[2] = ang;
            a.ag.poses[3 * b] = x; a.ag.poses[3 * b + 1] = y; a.ag.poses[3 * b + 2] = ang;
            if (a.out_best != nullptr) a.out_best[b] = (int16_t)best;
            if (a.out_poses != nullptr) { a.out_poses[3 * b] = x; a.out_poses[3 * b + 1] = y; a.out_poses[3 * b + 2] = ang; }
            if (logging) {
                a.log_best[(size_t)t * a.B + b] = (int16_t)best;
                a.log_pose[((size_t)t * a.B + b) * 3] = x;
                a.log_pose[((size_t)t * a.B + b) * 3 + 1] = y;
                a.log_pose[((size_t)t * a.B + b) * 3 + 2] = ang;
            }
            if (a.fake) {
                a.ag.completed[b] = completed + 1;
                s_more = (completed + 1 < budget);
            }
        }
        nvb_tl_stamp(a.tl, 4, 1);
        const int n_front = (nw - 1) * 32;   // every warp but the last
        asm volatile("bar.sync 2, %0;" ::"r"(n_front) : "memory");
        const double x = s_pose[0], y = s_pose[1];
        // the window of the next glimpses depends only on the new pose: its TMA loads fly during
        // the path scan; warp 1 computes the rotations meanwhile
        const bool oob = a.no_sample ? true : nvb_sample_window<false>(&tmap, sa, b, x, y, L, 0, A);
        if (warp == 1) {
            if (!oob) nvb_sample_rotations(sa, b, s_pose[2], L, L.offs, 32, 32, 0, A);
        } else if (!a.fake) {
            // ---- update_error (:252-276) by warps 0, 2 .. nw-2
            const int n_scan = (nw - 2) * 32, sid = (warp == 0) ? lane : tid - 32;
            const double thr = __dmul_rn(a.coverage_factor, a.step_size);   // :271
            const double thr2 = a.cover_thr2;
            const double2 *path = reinterpret_cast<const double2 *>(a.path);
            const float xf = (float)x, yf = (float)y;
            const float inf_f = __int_as_float(0x7F800000);
            // FP32 bounds of the distances of a block's points, rounded outwards: slack covers the
            // rounding of the position, of the arithmetic below and of the comparison
            auto bounds = [&](int j, float &lo, float &hi) {
                const float4 c = __ldg(reinterpret_cast<const float4 *>(a.pblk_f) + j);
                const float ex = c.x - xf, ey = c.y - yf;
                const float dc = __fsqrt_rn(fmaf(ex, ex, ey * ey));
                const float sl = fmaf(1e-6f, fabsf(xf) + fabsf(yf) + dc, 1e-5f);
                hi = dc + c.z + sl;
                lo = dc - c.z - sl;
            };
            float ub = inf_f, lo, hi;
            for (int j = sid; j < n_blk; j += n_scan) { bounds(j, lo, hi); ub = fminf(ub, hi); }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) ub = fminf(ub, __shfl_xor_sync(0xFFFFFFFFu, ub, o));
            if (lane == 0) s_ub[warp] = ub;
            asm volatile("bar.sync 1, %0;" ::"r"(n_scan) : "memory");
            float U = s_ub[0];
            for (int wq = 2; wq < nw - 1; wq++) U = fminf(U, s_ub[wq]);
            U = fmaxf(U, fmaf((float)thr, 1.0f + 1e-6f, 1e-5f));
            for (int j = sid; j < n_blk; j += n_scan) {
                bounds(j, lo, hi);
                if (lo <= U) { const int slot = atomicAdd(&s_nlive, 1); if (slot < NVB_TM_LIVE_MAX) s_live[slot] = j; }
            }
            asm volatile("bar.sync 1, %0;" ::"r"(n_scan) : "memory");
            const int n_live = s_nlive;
            double m = __longlong_as_double(0x7FF0000000000000ll);
            if (n_live <= NVB_TM_LIVE_MAX) {
                for (int idx = sid; idx < n_live * NVB_PATH_BLOCK; idx += n_scan) {
                    const int n = s_live[idx / NVB_PATH_BLOCK] * NVB_PATH_BLOCK + (idx % NVB_PATH_BLOCK);
                    if (n < a.n_path) {
                        const double2 pt = __ldg(path + n);
                        const double dx = __dsub_rn(pt.x, x), dy = __dsub_rn(pt.y, y);
                        const double d2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
                        m = fmin(m, d2);
                        if (d2 <= thr2) a.ag.coverage[(size_t)b * a.n_path + n] = 1;
                    }
                }
            } else {
                for (int n = sid; n < a.n_path; n += n_scan) {
                    const double2 pt = __ldg(path + n);
                    const double dx = __dsub_rn(pt.x, x), dy = __dsub_rn(pt.y, y);
                    const double d2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
                    m = fmin(m, d2);
                    if (d2 <= thr2) a.ag.coverage[(size_t)b * a.n_path + n] = 1;
                }
            }
            m = nvb_warp_fmin(m);
            if (lane == 0) s_red[warp] = m;
            asm volatile("bar.sync 1, %0;" ::"r"(n_scan) : "memory");
            if (tid == 0) {
                m = s_red[0];
                for (int wq = 2; wq < nw - 1; wq++) m = fmin(m, s_red[wq]);
                const double dmin = __dsqrt_rn(m);
                int more = 0;
                a.ag.nav_frames[b] = nav_frames + 1;                    // :253
                if (dmin > a.max_dist) {                                // :263-264
                    a.ag.status[b] = -1;
                } else {
                    a.ag.err_sum[b] = __dadd_rn(err_sum, __dmul_rn(dmin, dmin));   // :267
                    a.ag.err_n[b] = err_n + 1;                                     // :268
                    const double ex = __dsub_rn(path_end.x, x), ey = __dsub_rn(path_end.y, y);   // :328 end-of-path test
                    const double de = __dsqrt_rn(__dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey)));
                    if (de <= __dmul_rn(a.threshold_factor, a.step_size)) {
                        a.ag.status[b] = 1;
                    } else {
                        a.ag.completed[b] = completed + 1;
                        more = (completed + 1 < budget);
                    }
                }
                s_more = more;
                nvb_tl_stamp(a.tl, 4, 2);
            }
        }
    }
    __syncthreads();   // rotations, s_more, the log's familiarities read the old glimpse rows
    nvb_tl_stamp(a.tl, 5, 0);
    if (a.no_sample) {   // (no window in flight)
        nvb_tl_stamp(a.tl, 3, 2);
        return;
    }
    const double x = s_pose[0], y = s_pose[1];
    const NvbWorld &w = sa.w;
    const bool oob = (x <= w.r || y <= w.r || x >= (double)w.cols - w.r || y >= (double)w.rows - w.r);
    if (!s_more) {
        // the agent stops here; a window in flight must land before the CTA's shared memory goes
        if (!oob && tid == 0) nvb_mbar_wait(L.mbar, 0);
        return;
    }
    if (oob) {   // NavBySceneFamiliarity.py:156-158, reported at the step it belongs to
        if (tid == 0) a.pending_fail[b] = -2;
        return;
    }
    const int xi = __double2int_rd(x), yi = __double2int_rd(y);
    const bool safe = (xi - w.R >= 0) && (xi + w.R + 1 < w.cols) && (yi - w.R >= 0) && (yi + w.R + 1 < w.rows);
    if (safe) {
        nvb_mbar_wait(L.mbar, 0);
        nvb_tl_stamp(a.tl, 5, 1);
        nvb_gather_lean<PH, PW>(sa, b, x, y, L);
    } else {
        nvb_sample_gather<false, PH, PW>(sa, b, x, y, L, a.pending_fail + b, 0, A);
    }
    nvb_tl_stamp(a.tl, 3, 2);
}
