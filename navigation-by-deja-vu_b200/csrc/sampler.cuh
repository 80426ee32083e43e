// sampler.cuh -- K1: glimpse extraction for all agents x headings.
//
// Replaces, fused into one launch per step-batch:
//   fill_sensor_from   navsim/util.pyx:137-168          (nearest-neighbour rotated gather)
//   downscale_chem     navsim/util.pyx:91-134           (block mean / hue vote / S quirk)
//   quantise + mask    navsim/NavBySceneFamiliarity.py:175-192
//   bounds test        navsim/NavBySceneFamiliarity.py:153-158
//
// One CTA per agent (or per pose).  The landscape window the rotated sensor can
// reach is staged in shared memory by TMA (one 3-D tensor-map box per plane);
// all A headings of the agent gather from it.  Samples that leave the window
// (negative-index wrap-around, util.pyx:165-168 with wraparound on) fall back
// to global memory.  Output: planar glimpses gv/gh/gs [G][Ppad].
#pragma once
#include "common.cuh"
#include "glibc_trig.cuh"

struct SamplerArgs {
    NvbWorld w;
    const double *poses;     // [B][3] x, y, angle
    const double *offsets;   // [A] heading offsets (agent mode) or nullptr
    const double *cs;        // [B][2] host cos/sin of -(pi/2 - angle) (pose mode) or nullptr
    int A;
    int agent_mode;          // 1: resident stepping loop, 0: explicit pose list (A == 1)
    int need_hs;             // also produce the H and S planes
    int32_t *status;         // agent mode: agent status (in/out); pose mode: [B] out
    const int32_t *completed;  // agent mode: frames completed
    const int32_t *budget;     // agent mode: frame budget
    uint8_t *gv, *gh, *gs;   // [B*A][Ppad]
    unsigned long long *keys;  // [B*A] reset to "none" (agent mode) or nullptr
    float band;              // FP32 fast path: distance from a rounding tie below which FP64 decides
    long long *dbg;          // optional [B][8] clock64 checkpoints (tuning aid), else nullptr
    // host-driven form without copy operations: the poses are read straight from the caller's
    // page-locked buffer (device-mapped), written through to `poses_dst`, and the parked
    // failure of the agent is cleared; all nullptr otherwise
    const double *poses_src;
    double *poses_dst;
    int32_t *pending_clear;
    // tensor-core distance kernel (distance_tc.cuh): the sampler also writes the thermometer
    // encoding of every glimpse, genc [B*A][Kpad] int8, pixel-major (k = p * n_planes + plane);
    // tc_tab = {[256] raw block mean -> level index, [NVB_TC_TAB_LEVELS][8] level -> the
    // glimpse-side bytes +-w_k}.  genc == nullptr: not wanted.
    int8_t *genc;
    const uint8_t *tc_tab;
    int Kpad, n_planes;
};

#define NVB_TC_TAB_LEVELS 16
#define NVB_TC_TAB_BYTES (256 + NVB_TC_TAB_LEVELS * 8)

#define NVB_SAMPLER_THREADS 256
#define NVB_PTAB_MAX 256   /* sensors of up to this many pixels keep a per-pixel origin table in shared memory */

// Host + device: dynamic shared memory of k1_sample.
__host__ __device__ inline size_t nvb_sampler_smem(int BW, int BH, int nplanes, int A)
{
    size_t win = (size_t)nvb_round_up(BW * BH, 128) * (size_t)nplanes;
    // + rotations (FP64 and FP32), heading offsets, mbarrier, quantisation tables, pixel table,
    // thermometer tables of the tensor-core distance kernel (256 + 16 * 8 bytes)
    return win + (size_t)A * 32 + 64 + 768 + NVB_PTAB_MAX * 8 + 384;
}

// Sample coordinates (util.pyx:159-168).  The reference evaluates
//   ix = (ssize_t) round(px*c - py*s + x),  iy = (ssize_t) round(px*s + py*c + y)
// in FP64.  Here the coordinate relative to (floor x, floor y) is tracked in FP32
// (one FADD per sample along a sensor-pixel block) and rounded with the
// add-magic-number trick (no conversion instructions); whenever it lands within
// `band` of a rounding tie -- band is several times the worst FP32 error -- the
// reference's exact FP64 expression decides instead.  Every index is therefore
// exactly the reference's.
#define NVB_RND_MAGIC 12582912.0f        /* 1.5 * 2^23: (t + M) - M = t rounded to an integer */
#define NVB_RND_MAGIC_BITS 0x4B400000

struct SampleCtx {
    double c, s, x, y, half_w, half_h;
    int xi, yi, rows, cols;
    float band;
};

// exact FP64 path + the reference's wrap-around / bounds semantics
__device__ __noinline__ bool nvb_sample_exact(const SampleCtx &q, int col, int row, int &ix, int &iy)
{
    const double px = (double)col - q.half_w;   // util.pyx:159
    const double py = (double)row - q.half_h;   // util.pyx:160
    const double rx = __dsub_rn(__dmul_rn(px, q.c), __dmul_rn(py, q.s));   // :161
    const double ry = __dadd_rn(__dmul_rn(px, q.s), __dmul_rn(py, q.c));   // :162
    long long ly = (long long)round(__dadd_rn(ry, q.y));                  // :166
    long long lx = (long long)round(__dadd_rn(rx, q.x));                  // :167
    // boundscheck on, wraparound on (no decorators at util.pyx:137): a negative
    // index wraps once; anything still outside raises IndexError
    if (lx < 0) lx += q.cols;
    if (ly < 0) ly += q.rows;
    if (lx < 0 || lx >= q.cols || ly < 0 || ly >= q.rows) return false;
    ix = (int)lx;
    iy = (int)ly;
    return true;
}

#undef NVB_SAMPLER_THREADS
#define NVB_SAMPLER_THREADS 128

// Everything one sensor-pixel block needs besides the per-heading rotation.
struct BlockCtx {
    SampleCtx q;
    const uint8_t *win_v, *win_h, *win_s;        // staged window planes (shared memory)
    const uint8_t *land_v, *land_h, *land_s;     // landscape planes (global memory)
    int use_win, ox, oy, BW, BH, pitch, ph, pw;
    float xf, yf, tie;
};

// Careful path: per sample FP32 coordinate, exact FP64 on near-ties, wrap-around,
// bounds (-> *err) and window tests.  Returns the V sum of the block; fills hh/ss
// (hue / saturation samples) when NEED_HS.
template <bool NEED_HS>
__device__ __noinline__ int nvb_block_careful(const BlockCtx &c, float cf, float sf, int col0, int row0,
                                              uint8_t *hh, uint8_t *ss, int *n_out, int *err)
{
    const SampleCtx &q = c.q;
    const float px0 = (float)col0 - (float)q.half_w, py0 = (float)row0 - (float)q.half_h;
    float tx_row = fmaf(px0, cf, fmaf(-py0, sf, c.xf));
    float ty_row = fmaf(px0, sf, fmaf(py0, cf, c.yf));
    int sum_v = 0, n = 0;
    for (int i = 0; i < c.ph; i++) {
        float tx = tx_row, ty = ty_row;
        for (int j = 0; j < c.pw; j++) {
            const float ux = tx + NVB_RND_MAGIC, uy = ty + NVB_RND_MAGIC;
            const float dx = fabsf(tx - (ux - NVB_RND_MAGIC)), dy = fabsf(ty - (uy - NVB_RND_MAGIC));
            tx += cf;
            ty += sf;
            int ix, iy;
            if (fmaxf(dx, dy) < c.tie) {
                ix = q.xi + (__float_as_int(ux) - NVB_RND_MAGIC_BITS);
                iy = q.yi + (__float_as_int(uy) - NVB_RND_MAGIC_BITS);
                if (ix < 0) ix += q.cols;
                if (iy < 0) iy += q.rows;
                if ((unsigned)ix >= (unsigned)q.cols || (unsigned)iy >= (unsigned)q.rows) {
                    *err = 1;
                    continue;
                }
            } else if (!nvb_sample_exact(q, col0 + j, row0 + i, ix, iy)) {
                *err = 1;
                continue;
            }
            const int lx = ix - c.ox, ly = iy - c.oy;
            if (c.use_win && (unsigned)lx < (unsigned)c.BW && (unsigned)ly < (unsigned)c.BH) {
                const int o = ly * c.BW + lx;
                sum_v += c.win_v[o];
                if (NEED_HS) { hh[n] = c.win_h[o]; ss[n] = c.win_s[o]; }
            } else {
                const size_t o = (size_t)iy * c.pitch + (size_t)ix;
                sum_v += __ldg(c.land_v + o);
                if (NEED_HS) { hh[n] = __ldg(c.land_h + o); ss[n] = __ldg(c.land_s + o); }
            }
            n++;
        }
        tx_row -= sf;
        ty_row += cf;
    }
    *n_out = n;
    return sum_v;
}

// Glimpses of one agent / pose at (x, y, ang): bounds test, TMA window, rotation
// per heading, sampling, block reduce, quantise, mask.  Called by the whole CTA
// (NVB_SAMPLER_THREADS threads).  A failing pose writes -2 (out of bounds,
// NavBySceneFamiliarity.py:156-158) or -3 (IndexError, util.pyx:165-168) to
// *fail_out and produces no (or partial) glimpses.
// PH, PW: landscape pixels per sensor pixel known at compile time (0 = runtime).
// Copy of the quantisation tables [3][256] from global memory to their shared-memory home.
__device__ __forceinline__ void nvb_sampler_stage_lut(const NvbWorld &w, uint8_t *lut_sm)
{
    for (int k = threadIdx.x; k < 768 / 4; k += blockDim.x)
        reinterpret_cast<uint32_t *>(lut_sm)[k] = __ldg(reinterpret_cast<const uint32_t *>(w.lut) + k);
}
// Sensor pixel p = bi * W + bj -> sensor-frame coordinates of its block's first sample
// (util.pyx:159-160: col - Wpx/2, row - Hpx/2; halves are exact in FP32).  Constant per world.
__device__ __forceinline__ void nvb_sampler_stage_ptab(const NvbWorld &w, float2 *ptab)
{
    if (w.P > NVB_PTAB_MAX) return;
    const float half_wf = 0.5f * (float)w.Wpx, half_hf = 0.5f * (float)w.Hpx;
    for (int p = threadIdx.x; p < w.P; p += blockDim.x) {
        const int bi = p / w.W, bj = p - bi * w.W;
        ptab[p] = make_float2((float)(bj * w.pw) - half_wf, (float)(bi * w.ph) - half_hf);
    }
}

// Thermometer tables of the tensor-core distance kernel -> shared memory (constant per world).
__device__ __forceinline__ void nvb_sampler_stage_tc(const uint8_t *tc_tab, uint8_t *tc_sm)
{
    if (tc_tab == nullptr) return;
    for (int k = threadIdx.x; k < NVB_TC_TAB_BYTES / 4; k += blockDim.x)
        reinterpret_cast<uint32_t *>(tc_sm)[k] = __ldg(reinterpret_cast<const uint32_t *>(tc_tab) + k);
}

// Shared-memory layout of one sampling CTA (dynamic shared memory, nvb_sampler_smem bytes).
struct SamplerSmem {
    uint8_t *win_v, *win_h, *win_s;   // staged window planes
    double *cs;                       // [A][2] cos, sin of every heading's rotation (FP64)
    uint64_t *mbar;                   // completion barrier of the window's TMA loads
    uint8_t *lut;                     // [3][256] quantisation tables
    float2 *csf;                      // [A] the rotations rounded to FP32 (fast path)
    double *offs;                     // [A] heading offsets (staged by kernels that want them close)
    float2 *ptab;                     // [P <= NVB_PTAB_MAX] sensor pixel -> (px0, py0), its block's origin
    int *err;                         // IndexError flag of this CTA
    uint8_t *tc;                      // [NVB_TC_TAB_BYTES] thermometer tables (SamplerArgs::tc_tab), 8-byte aligned
};

template <bool NEED_HS>
__device__ __forceinline__ SamplerSmem nvb_sampler_layout(const NvbWorld &w, int A, uint8_t *smem)
{
    constexpr int nplanes = NEED_HS ? 3 : 1;
    const size_t plane_sz = (w.R > 0) ? (size_t)nvb_round_up(w.BW * w.BH, 128) : 0;
    SamplerSmem L;
    L.win_v = smem;                   // plane order in smem: V, H, S
    L.win_h = smem + plane_sz;
    L.win_s = smem + 2 * plane_sz;
    L.cs = (double *)(smem + plane_sz * nplanes);
    L.mbar = (uint64_t *)(L.cs + 2 * A);
    L.lut = (uint8_t *)(L.mbar + 1);
    L.csf = (float2 *)(L.lut + 768);
    L.offs = (double *)(L.csf + A);
    L.ptab = (float2 *)(L.offs + A);
    L.err = (int *)(L.ptab + NVB_PTAB_MAX);
    L.tc = (uint8_t *)(L.err + 2);
    return L;
}

// Step 1 (every thread): bounds test of the pose (NavBySceneFamiliarity.py:156-158; true =
// out of bounds, nothing touched), then thread 0 starts the TMA loads of the window the
// rotated sensor can reach; the keys of this agent's headings are reset for the next K2.
template <bool NEED_HS>
__device__ __forceinline__ bool nvb_sample_window(const CUtensorMap *tmap, const SamplerArgs &a, int b,
                                                  double x, double y, const SamplerSmem &L, int k0, int k1)
{
    const NvbWorld &w = a.w;
    const int tid = threadIdx.x;
    constexpr int nplanes = NEED_HS ? 3 : 1;
    if (x <= w.r || y <= w.r || x >= (double)w.cols - w.r || y >= (double)w.rows - w.r) return true;
    // TMA needs the box's innermost start coordinate on a 16-byte boundary (an
    // unaligned start traps with an illegal-instruction error on sm_100); BW has
    // 15 spare columns for that.
    const int xi = (int)floor(x), yi = (int)floor(y);
    const int ox = (xi - w.R) & ~15, oy = yi - w.R;
    if (tid == 0) {
        *L.err = 0;
        if (w.R > 0) {
            nvb_mbar_init(L.mbar, 1);
            nvb_fence_barrier_init();
            nvb_mbar_expect_tx(L.mbar, (uint32_t)(w.BW * w.BH * nplanes));
            nvb_tma_load_3d(L.win_v, tmap, ox, oy, 2, L.mbar);
            if (NEED_HS) {
                nvb_tma_load_3d(L.win_h, tmap, ox, oy, 0, L.mbar);
                nvb_tma_load_3d(L.win_s, tmap, ox, oy, 1, L.mbar);
            }
        }
    }
    if (a.keys != nullptr)
        for (int k = k0 + tid; k < k1; k += blockDim.x) a.keys[(size_t)b * a.A + k] = NVB_KEY_NONE;
    return false;
}

// Step 2 (threads first_thread .. first_thread + n_threads - 1): per-heading rotation
// (util.pyx:143-145).  offsets: a.offsets or a copy of them in shared memory.
__device__ __forceinline__ void nvb_sample_rotations(const SamplerArgs &a, int b, double ang,
                                                     const SamplerSmem &L, const double *offsets,
                                                     int first_thread, int n_threads, int k0, int k1)
{
    const int j = (int)threadIdx.x - first_thread;
    if (j < 0 || j >= n_threads) return;
    for (int k = k0 + j; k < k1; k += n_threads) {
        double c, s;
        if (a.cs != nullptr) {
            c = a.cs[2 * b];
            s = a.cs[2 * b + 1];
        } else {
            double angle = ang;
            if (a.agent_mode) angle = nvb_pymod_pos(__dadd_rn(ang, offsets[k]), NVB_TWO_PI);
            double rot = -__dsub_rn(0.5 * NVB_PI, angle);
            nvb_glibc_sincos(rot, &s, &c);   // the host libm's bits (util.pyx:144-145 call libc cos / sin)
        }
        L.cs[2 * k] = c;
        L.cs[2 * k + 1] = s;
        L.csf[k] = make_float2((float)c, (float)s);
    }
}

template <bool NEED_HS, int PH, int PW>
__device__ __forceinline__ void nvb_sample_gather(const SamplerArgs &a, int b, double x, double y,
                                                  const SamplerSmem &L, int32_t *fail_out, int k0, int k1);

// The three steps in one go.  LUT_STAGED: the caller has already copied the quantisation
// tables to L.lut (they are constant; nvb_sampler_stage_lut).
// k0, k1: the headings [k0, k1) this CTA samples (all of them, or one slice of a wide sweep).
template <bool NEED_HS, int PH, int PW, bool LUT_STAGED = false>
__device__ __forceinline__ void nvb_sample_body(const CUtensorMap *tmap, const SamplerArgs &a, int b,
                                                double x, double y, double ang, uint8_t *smem,
                                                int32_t *fail_out, int k0, int k1)
{
    const SamplerSmem L = nvb_sampler_layout<NEED_HS>(a.w, a.A, smem);
    if (nvb_sample_window<NEED_HS>(tmap, a, b, x, y, L, k0, k1)) {
        if (threadIdx.x == 0) *fail_out = -2;
        return;
    }
    // while the window is in flight: quantisation tables -> shared memory, rotations
    if (!LUT_STAGED) {
        nvb_sampler_stage_lut(a.w, L.lut);
        nvb_sampler_stage_ptab(a.w, L.ptab);
        nvb_sampler_stage_tc(a.genc != nullptr ? a.tc_tab : nullptr, L.tc);
    }
    nvb_sample_rotations(a, b, ang, L, a.offsets, 0, (int)blockDim.x, k0, k1);
    nvb_sample_gather<NEED_HS, PH, PW>(a, b, x, y, L, fail_out, k0, k1);
}

// Step 3 (every thread): wait for the window, gather, block mean, quantise, mask.
template <bool NEED_HS, int PH, int PW>
__device__ __forceinline__ void nvb_sample_gather(const SamplerArgs &a, int b, double x, double y,
                                                  const SamplerSmem &L, int32_t *fail_out, int k0, int k1)
{
    const NvbWorld &w = a.w;
    const int tid = threadIdx.x;
    const int use_win = (w.R > 0);
    const uint8_t *win_v = L.win_v, *win_h = L.win_h, *win_s = L.win_s;
    const double *cs_sm = L.cs;
    const uint8_t *lut_sm = L.lut;
    const float2 *csf_sm = L.csf;
    const double fx = floor(x), fy = floor(y);
    const int xi = (int)fx, yi = (int)fy;
    const int ox = (xi - w.R) & ~15, oy = yi - w.R;
    if (a.dbg && tid == 0) a.dbg[b * 8 + 4] = clock64();
    __syncthreads();   // rotations, tables, barrier initialisation
    if (use_win) nvb_mbar_wait(L.mbar, 0);
    if (a.dbg && tid == 0) a.dbg[b * 8 + 5] = clock64();

    const int ph = PH ? PH : w.ph, pw = PW ? PW : w.pw;
    const int nblk = pw * ph;
    BlockCtx c;
    c.q.x = x; c.q.y = y;
    c.q.half_w = 0.5 * (double)w.Wpx; c.q.half_h = 0.5 * (double)w.Hpx;
    c.q.xi = xi; c.q.yi = yi;
    c.q.band = a.band;
    c.q.rows = w.rows; c.q.cols = w.cols;
    c.win_v = win_v; c.win_h = win_h; c.win_s = win_s;
    c.land_h = w.land; c.land_s = w.land + w.plane_stride; c.land_v = w.land + 2 * w.plane_stride;
    c.use_win = use_win; c.ox = ox; c.oy = oy; c.BW = w.BW; c.BH = w.BH; c.pitch = w.pitch;
    c.ph = ph; c.pw = pw;
    c.xf = (float)(x - fx); c.yf = (float)(y - fy);
    c.tie = 0.5f - a.band;
    // `safe`: every sample of this agent lies inside the landscape and inside the
    // staged window (R = reach + 1), so no wrap / bounds / window test is needed
    const bool safe = !NEED_HS && use_win && (xi - w.R >= 0) && (xi + w.R + 1 < w.cols) &&
                      (yi - w.R >= 0) && (yi + w.R + 1 < w.rows);
    // window index of sample (lx, ly) relative to (floor x, floor y), with the magic-number
    // bias of both float bit patterns folded in (32-bit wrap-around arithmetic is exact)
    const int kbase = ((yi - oy) - NVB_RND_MAGIC_BITS) * w.BW + ((xi - ox) - NVB_RND_MAGIC_BITS);
    const float inv_w = 1.0f / (float)w.W;
    const float half_wf = (float)c.q.half_w, half_hf = (float)c.q.half_h;
    int err = 0;

    // it = k * P + p, p = bi * W + bj: k and p advance incrementally (no divisions in the loop)
    const int T = (int)blockDim.x;
    const int dk = T / w.P, dp = T - dk * w.P;
    int k = k0 + tid / w.P, p = tid - (tid / w.P) * w.P;
    const bool have_tab = w.P <= NVB_PTAB_MAX;
    const bool linear_out = (w.Ppad == w.P);   // then the output offset of item `it` is just `it`
    // NavBySceneFamiliarity.py:189-190 in units of px0 = bj * pw - Wpx / 2 (exact in FP32)
    const float mask_lo = (float)(w.mask_lo * pw) - half_wf, mask_hi = (float)(w.mask_hi * pw) - half_wf;
    const size_t out0 = (size_t)b * a.A * w.Ppad;

    for (int it = k0 * w.P + tid; it < k1 * w.P; it += T) {
        float px0, py0;
        if (have_tab) {
            const float2 e = L.ptab[p];
            px0 = e.x; py0 = e.y;
        } else {
            // small integers: the float quotient is exact
            const int bi = (int)(((float)p + 0.5f) * inv_w), bj = p - bi * w.W;
            px0 = (float)(bj * pw) - half_wf; py0 = (float)(bi * ph) - half_hf;
        }
        const float2 csf = csf_sm[k];
        const float cf = csf.x, sf = csf.y;
        int sum_v = 0, n = 0;
        uint8_t hh[NEED_HS ? NVB_MAX_BLOCK_PX : 1], ss[NEED_HS ? NVB_MAX_BLOCK_PX : 1];
        bool done = false;
        if (safe) {
            // lean path: no branches; the largest distance from a rounding tie seen in
            // the block decides afterwards whether it must be redone
            // (sm_100 packed FP32x2: x and y coordinate advance, round and compare together)
            const float2 t_row0 = make_float2(fmaf(px0, cf, fmaf(-py0, sf, c.xf)), fmaf(px0, sf, fmaf(py0, cf, c.yf)));
            float2 t_row = t_row0;
            const float2 step_j = make_float2(cf, sf), step_i = make_float2(-sf, cf);
            const float2 magic = make_float2(NVB_RND_MAGIC, NVB_RND_MAGIC);
            const float2 neg_magic = make_float2(-NVB_RND_MAGIC, -NVB_RND_MAGIC), neg_one = make_float2(-1.0f, -1.0f);
            float worst = 0.0f;
#pragma unroll
            for (int i = 0; i < ph; i++) {
                float2 t = t_row;
#pragma unroll
                for (int j = 0; j < pw; j++) {
                    const float2 u = __fadd2_rn(t, magic);                     // low mantissa bits = round(t)
                    const float2 e = __ffma2_rn(__fadd2_rn(u, neg_magic), neg_one, t);   // t - round(t)
                    worst = fmaxf(worst, fmaxf(fabsf(e.x), fabsf(e.y)));
                    sum_v += win_v[__float_as_int(u.y) * w.BW + __float_as_int(u.x) + kbase];
                    t = __fadd2_rn(t, step_j);
                }
                t_row = __fadd2_rn(t_row, step_i);
            }
            if (worst >= c.tie) {
                // rare (about one block in a thousand): some sample sits within the band of a
                // rounding tie.  Same coordinates again; the samples inside the band are
                // re-addressed with the reference's FP64 expression and the sum corrected.
                if (a.dbg) atomicAdd((unsigned long long *)a.dbg + b * 8 + 7, 1ull);
                c.q.c = cs_sm[2 * k]; c.q.s = cs_sm[2 * k + 1];
                const int col0 = (int)(px0 + half_wf), row0 = (int)(py0 + half_hf);
                t_row = t_row0;
#pragma unroll
                for (int i = 0; i < ph; i++) {
                    float2 t = t_row;
#pragma unroll
                    for (int j = 0; j < pw; j++) {
                        const float2 u = __fadd2_rn(t, magic);
                        const float2 e = __ffma2_rn(__fadd2_rn(u, neg_magic), neg_one, t);
                        if (!(fmaxf(fabsf(e.x), fabsf(e.y)) < c.tie)) {
                            int ix, iy;
                            if (nvb_sample_exact(c.q, col0 + j, row0 + i, ix, iy))
                                sum_v += (int)win_v[(iy - oy) * w.BW + (ix - ox)] -
                                         (int)win_v[__float_as_int(u.y) * w.BW + __float_as_int(u.x) + kbase];
                            else
                                err = 1;   // cannot happen for a `safe` agent
                        }
                        t = __fadd2_rn(t, step_j);
                    }
                    t_row = __fadd2_rn(t_row, step_i);
                }
            }
            done = true;
        }
        if (!done) {
            c.q.c = cs_sm[2 * k]; c.q.s = cs_sm[2 * k + 1];
            const int col0 = (int)(px0 + half_wf), row0 = (int)(py0 + half_hf);
            sum_v = nvb_block_careful<NEED_HS>(c, cf, sf, col0, row0, hh, ss, &n, &err);
        }
        // util.pyx:121-123: V = (uint8) round(sum / (fr*fc)), half away from zero.  In
        // integers: the quotient is either an exact tie or at least 1/(2*fr*fc) away
        // from one, far more than the FP64 division's rounding.
        uint8_t v = (uint8_t)((2 * sum_v + nblk) / (2 * nblk));
        const bool masked = (px0 >= mask_lo && px0 < mask_hi);
        const size_t o = out0 + (linear_out ? (size_t)it : (size_t)k * w.Ppad + p);
        a.gv[o] = masked ? 0 : lut_sm[512 + v];
        if (a.genc != nullptr) {
            // the same pixel as +-w_k thermometer bytes for the tensor-core distance kernel
            const int lvl = masked ? 0 : (int)L.tc[v];
            const uint8_t *enc = L.tc + 256 + lvl * 8;
            int8_t *eo = a.genc + ((size_t)b * a.A + k) * a.Kpad + (size_t)p * a.n_planes;
            if (a.n_planes == 4) {
                *reinterpret_cast<uint32_t *>(eo) = *reinterpret_cast<const uint32_t *>(enc);
            } else {
                for (int q = 0; q < a.n_planes; q++) eo[q] = (int8_t)enc[q];
            }
        }
        if (NEED_HS) {
            // util.pyx:126-132: hue with the largest summed S (strict >, so the
            // lowest hue wins ties and hue 0 wins when every sum is 0);
            // S = (uint8) round((conc / fr) * fc) with C integer division.
            int best_sum = 0, best_h = 0;
            for (int t = 0; t < n; t++)
                if (hh[t] == 0) best_sum += ss[t];
            for (int t = 0; t < n; t++) {
                int sq = 0;
                for (int u = 0; u < n; u++)
                    if (hh[u] == hh[t]) sq += ss[u];
                if (sq > best_sum || (sq == best_sum && (int)hh[t] < best_h)) {
                    best_sum = sq;
                    best_h = hh[t];
                }
            }
            const uint8_t sat = (uint8_t)((best_sum / ph) * pw);
            a.gh[o] = masked ? 0 : lut_sm[best_h];
            a.gs[o] = masked ? 0 : lut_sm[256 + sat];
        }
        k += dk;
        p += dp;
        if (p >= w.P) { p -= w.P; k++; }
    }
    if (err) *L.err = 1;
    __syncthreads();
    if (tid == 0 && *L.err) *fail_out = -3;   // IndexError, util.pyx:165-168
    if (a.dbg && tid == 0) a.dbg[b * 8 + 6] = clock64();
}

// K1 as its own launch: one CTA per agent (resident loop) or per pose (A == 1).
template <bool NEED_HS, int PH, int PW>
__global__ void __launch_bounds__(NVB_SAMPLER_THREADS)
k1_sample(const __grid_constant__ CUtensorMap tmap, SamplerArgs a)
{
    nvb_grid_dep_wait();
    extern __shared__ __align__(128) uint8_t smem_k1[];
    const int b = blockIdx.x;
    double px, py, pa;
    if (a.poses_src != nullptr) {
        // zero-copy input: one thread fetches the pose over the host link, everybody uses it;
        // the first slice also refreshes the device copy (stopped agents included, like the
        // host-to-device copy this replaces)
        __shared__ double s_pose_in[3];
        if (threadIdx.x == 0) {
            for (int q = 0; q < 3; q++) {
                const double v = a.poses_src[3 * b + q];
                s_pose_in[q] = v;
                if (blockIdx.y == 0) a.poses_dst[3 * b + q] = v;
            }
            if (blockIdx.y == 0 && a.pending_clear != nullptr) a.pending_clear[b] = 0;
        }
        __syncthreads();
        px = s_pose_in[0]; py = s_pose_in[1]; pa = s_pose_in[2];
    } else {
        px = a.poses[3 * b]; py = a.poses[3 * b + 1]; pa = a.poses[3 * b + 2];
    }
    if (a.agent_mode) {
        if (!(a.status[b] == 0 && a.completed[b] < a.budget[b])) return;
    } else if (threadIdx.x == 0) {
        a.status[b] = 0;
    }
    // wide sweeps of few agents: gridDim.y CTAs share an agent, each a slice of its headings
    const int k0 = (int)((long long)a.A * blockIdx.y / gridDim.y), k1 = (int)((long long)a.A * (blockIdx.y + 1) / gridDim.y);
    nvb_sample_body<NEED_HS, PH, PW>(&tmap, a, b, px, py, pa, smem_k1, a.status + b, k0, k1);
}

// planar [G][Ppad] x3 -> interleaved [G][P][3] (familiar_scenes / get_sensor_mat layout)
__global__ void k_planar_to_hsv(const uint8_t *gh, const uint8_t *gs, const uint8_t *gv, int Ppad,
                                int P, long long G, uint8_t *out)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= G * P) return;
    long long g = i / P;
    int p = (int)(i - g * P);
    size_t o = (size_t)g * Ppad + p;
    out[3 * i] = gh[o];
    out[3 * i + 1] = gs[o];
    out[3 * i + 2] = gv[o];
}

// interleaved [G][P][3] -> planar [G][Ppad] x3 (pad bytes are zeroed by the caller)
__global__ void k_hsv_to_planar(const uint8_t *in, int Ppad, int P, long long G, uint8_t *gh,
                                uint8_t *gs, uint8_t *gv)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= G * P) return;
    long long g = i / P;
    int p = (int)(i - g * P);
    size_t o = (size_t)g * Ppad + p;
    gh[o] = in[3 * i];
    gs[o] = in[3 * i + 1];
    gv[o] = in[3 * i + 2];
}

// Stand-alone A1 (util.pyx:137-168) for the single-call API: one thread per
// sensor sample, straight from global memory.  cs = host cos/sin.
__global__ void k_fill_sensor(NvbWorld w, int Hpx, int Wpx, double x, double y, double c, double s,
                              uint8_t *out, int *err)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Hpx * Wpx) return;
    int r = i / Wpx, q = i - r * Wpx;
    double px = (double)q - 0.5 * (double)Wpx, py = (double)r - 0.5 * (double)Hpx;
    double rx = __dsub_rn(__dmul_rn(px, c), __dmul_rn(py, s));
    double ry = __dadd_rn(__dmul_rn(px, s), __dmul_rn(py, c));
    long long iy = (long long)round(__dadd_rn(ry, y));
    long long ix = (long long)round(__dadd_rn(rx, x));
    if (iy < 0) iy += w.rows;
    if (ix < 0) ix += w.cols;
    if (iy < 0 || iy >= w.rows || ix < 0 || ix >= w.cols) {
        *err = 1;
        return;
    }
    size_t o = (size_t)iy * w.pitch + (size_t)ix;
    out[3 * i] = w.land[o];
    out[3 * i + 1] = w.land[w.plane_stride + o];
    out[3 * i + 2] = w.land[2 * w.plane_stride + o];
}

// Stand-alone A2 (util.pyx:91-134): one thread per output pixel.
__global__ void k_downscale_chem(const uint8_t *img, int R, int C, int fr, int fc, uint8_t *out)
{
    int nrb = R / fr, ncb = C / fc;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nrb * ncb) return;
    int bi = i / ncb, bj = i - bi * ncb;
    long long sum_v = 0;
    // hue vote without a 256-bin histogram: for every sample, total S of its hue
    long long best_sum = 0;
    int best_h = 0;
    for (int r = 0; r < fr; r++)
        for (int c = 0; c < fc; c++) {
            const uint8_t *p = img + ((size_t)(bi * fr + r) * C + (bj * fc + c)) * 3;
            sum_v += p[2];
            if (p[0] == 0) best_sum += p[1];
        }
    for (int r = 0; r < fr; r++)
        for (int c = 0; c < fc; c++) {
            const uint8_t *p = img + ((size_t)(bi * fr + r) * C + (bj * fc + c)) * 3;
            long long sq = 0;
            for (int r2 = 0; r2 < fr; r2++)
                for (int c2 = 0; c2 < fc; c2++) {
                    const uint8_t *q = img + ((size_t)(bi * fr + r2) * C + (bj * fc + c2)) * 3;
                    if (q[0] == p[0]) sq += q[1];
                }
            if (sq > best_sum || (sq == best_sum && (int)p[0] < best_h)) {
                best_sum = sq;
                best_h = p[0];
            }
        }
    out[3 * i + 2] = (uint8_t)(int)round(__ddiv_rn((double)sum_v, (double)((long long)fr * fc)));
    out[3 * i] = (uint8_t)best_h;
    out[3 * i + 1] = (uint8_t)((best_sum / fr) * fc);
}
