// step.cuh -- K3: the device-resident agent-stepping loop around K1/K2.
//
// Replaces the tail of step_forward (navsim/NavBySceneFamiliarity.py:313-329)
// and update_error (:252-276) for a whole batch of agents, one CTA per agent:
//   decide  exact FP64 difference of the best view of every heading
//           (util.pyx:59-73 operation order); headings tied at the step's integer
//           minimum are marked
//   ties    for tied headings: exact FP64 difference of EVERY view that attains
//           the minimum (the reference's argmax over doubles breaks integer ties
//           by rounding noise)
//   move    argmax heading (first maximum, :315), move (:317-323),
//           update_error (:252-276), end-of-path test (:328), step log
//
// Two forms.  k3_step fuses all three in one launch (each agent's CTA scans the
// library itself for its tied headings): used when the library is small enough
// for that scan to be cheap.  k3_decide / k3_ties / k3_move are the same code as
// three launches with a grid-wide tie pass in between: used for large or
// view-sharded libraries, where the two reduction buffers (keys, exact) are
// MIN-all-reduced across ranks between the phases.
#pragma once
#include "common.cuh"
#include "distance.cuh"

struct AgentState {
    double *poses;          // [B][3]
    int32_t *status;        // [B] 0 running, else stop code
    int32_t *completed;     // [B] steps that returned normally
    int32_t *budget;        // [B] frame budget
    int32_t *nav_frames;    // [B] navigated_for_frames
    double *err_sum;        // [B]
    int32_t *err_n;         // [B]
    uint8_t *coverage;      // [B][N]
    int32_t *stepped;       // [B] scratch: 1 if the agent takes part in the current step
};

struct StepArgs {
    AgentState ag;
    int B, A, N, P, Ppad;
    const double *offsets;       // [A]
    const uint8_t *gv, *gh, *gs; // glimpses [B*A][Ppad]
    const uint8_t *lv, *lh, *ls; // library  [N][Ppad] (local shard)
    const double *path;          // [n_path][2] training path (whole path on every rank)
    int n_path;
    long long view_offset;       // first global view index of the local shard
    unsigned long long *keys;    // [B*A]
    unsigned long long *exact;   // [B*A] FP64 bit patterns of the exact min difference
    int idx_bits;
    unsigned long long band;     // score band treated as tied (0 for chem_weight == 0)
    double cw;
    const double *div255;
    double maxfam;               // H*W
    double step_size, max_dist, threshold_factor, coverage_factor;
    int fake;
    // tie work list (three-launch form)
    int *tie_count;
    int2 *tie_items;             // (glimpse index, unused)
    unsigned long long *tie_thr; // per item: largest score still treated as tied
    // log
    const int *step_counter;     // device step index
    int log_cap;
    int16_t *log_best;           // [cap][B]
    double *log_pose;            // [cap][B][3]
    double *log_sfam;            // [cap][B]
    double *log_afam;            // [cap][B][A] or nullptr
};

#define NVB_STEP_THREADS 128

__device__ __forceinline__ bool nvb_agent_active(const AgentState &ag, int b)
{
    return ag.status[b] == 0 && ag.completed[b] < ag.budget[b];
}

// Exact FP64 difference with 16-byte row loads (rows are 16-B aligned and zero
// padded; a zero pad pixel adds +0.0, which leaves the sum unchanged).
__device__ __forceinline__ double nvb_exact_diff_rows(const StepArgs &a, size_t qo, size_t fo,
                                                      const double *div255)
{
    if (a.cw != 0.0)
        return nvb_exact_diff(a.gh + qo, a.gs + qo, a.gv + qo, a.lh + fo, a.ls + fo, a.lv + fo, a.P,
                              a.cw, div255);
    const uint4 *q = reinterpret_cast<const uint4 *>(a.gv + qo);
    const uint4 *f = reinterpret_cast<const uint4 *>(a.lv + fo);
    double diff = 0.0;
    for (int c = 0; c < a.Ppad / 16; c++) {
        const uint4 qq = q[c], ff = __ldg(f + c);
        const uint32_t d[4] = {__vabsdiffu4(qq.x, ff.x), __vabsdiffu4(qq.y, ff.y),
                               __vabsdiffu4(qq.z, ff.z), __vabsdiffu4(qq.w, ff.w)};
#pragma unroll
        for (int w = 0; w < 4; w++)
#pragma unroll
            for (int k = 0; k < 4; k++) diff = __dadd_rn(diff, div255[(d[w] >> (8 * k)) & 0xFFu]);
    }
    return diff;
}

// integer score of one (glimpse row, view row) pair, as K2 computes it
__device__ __forceinline__ unsigned long long nvb_pair_score(const StepArgs &a, size_t qo, size_t fo)
{
    const int words = a.Ppad / 4;
    const uint32_t *qv = reinterpret_cast<const uint32_t *>(a.gv + qo);
    const uint32_t *fv = reinterpret_cast<const uint32_t *>(a.lv + fo);
    if (a.cw == 0.0) {
        uint32_t s = 0;
        for (int wd = 0; wd < words; wd++) s = nvb_sad4(qv[wd], __ldg(fv + wd), s);
        return s;
    }
    const uint32_t *qh = reinterpret_cast<const uint32_t *>(a.gh + qo);
    const uint32_t *qs = reinterpret_cast<const uint32_t *>(a.gs + qo);
    const uint32_t *fh = reinterpret_cast<const uint32_t *>(a.lh + fo);
    const uint32_t *fs = reinterpret_cast<const uint32_t *>(a.ls + fo);
    uint32_t xs = 0, vs = 0;
    for (int wd = 0; wd < words; wd++)
        nvb_hsv_word(qh[wd], qs[wd], qv[wd], __ldg(fh + wd), __ldg(fs + wd), __ldg(fv + wd), xs, vs);
    return nvb_hsv_score(xs, vs, a.cw);
}

// decide: exact difference of every heading's best view and detection of headings
// tied at the step's minimum.  FUSED: the CTA scans the library for its tied
// headings right away; otherwise they go to the work list of k3_ties.
template <bool FUSED>
__device__ __forceinline__ void nvb_decide(const StepArgs &a, int b, const double *div255)
{
    const int tid = threadIdx.x;
    __shared__ unsigned long long s_min;
    __shared__ int s_ntied;
    if (tid == 0) { s_min = ~0ull; s_ntied = 0; }
    __syncthreads();
    const unsigned long long idx_mask = (1ull << a.idx_bits) - 1ull;
    unsigned long long local = ~0ull;
    for (int k = tid; k < a.A; k += blockDim.x)
        local = min(local, a.keys[(size_t)b * a.A + k] >> a.idx_bits);
    if (local != ~0ull) atomicMin(&s_min, local);
    __syncthreads();
    const unsigned long long thr = s_min + a.band;
    for (int k = tid; k < a.A; k += blockDim.x)
        if ((a.keys[(size_t)b * a.A + k] >> a.idx_bits) <= thr) atomicAdd(&s_ntied, 1);
    __syncthreads();
    const bool have_ties = s_ntied > 1;
    for (int k = tid; k < a.A; k += blockDim.x) {
        const size_t g = (size_t)b * a.A + k;
        const unsigned long long key = a.keys[g];
        const bool tied = have_ties && (key >> a.idx_bits) <= thr;
        unsigned long long ebits = NVB_EXACT_NONE;
        if (!(FUSED && tied)) {   // a fused tie scan revisits the best view anyway
            const long long v = (long long)(key & idx_mask) - a.view_offset;
            if (key != NVB_KEY_NONE && v >= 0 && v < a.N) {
                const double d = nvb_exact_diff_rows(a, g * a.Ppad, (size_t)v * a.Ppad, div255);
                ebits = (unsigned long long)__double_as_longlong(d);
            }
        }
        a.exact[g] = ebits;
        if (!FUSED && tied) {
            const int slot = atomicAdd(a.tie_count, 1);
            a.tie_items[slot] = make_int2((int)g, 0);
            a.tie_thr[slot] = thr;
        }
    }
    if (FUSED && have_ties) {
        __syncthreads();   // exact[] initialised
        for (int k = 0; k < a.A; k++) {
            const size_t g = (size_t)b * a.A + k;
            if ((a.keys[g] >> a.idx_bits) > thr) continue;   // CTA-uniform
            for (int v = tid; v < a.N; v += blockDim.x) {
                if (nvb_pair_score(a, g * a.Ppad, (size_t)v * a.Ppad) <= thr) {
                    const double d = nvb_exact_diff_rows(a, g * a.Ppad, (size_t)v * a.Ppad, div255);
                    atomicMin(a.exact + g, (unsigned long long)__double_as_longlong(d));
                }
            }
        }
    }
}

// move: argmax, pose update, update_error, end test, log.  Requires a.exact final.
__device__ __forceinline__ void nvb_move(const StepArgs &a, int b)
{
    const int tid = threadIdx.x;
    const int t = *a.step_counter;
    const bool logging = (t >= 0 && t < a.log_cap);
    __shared__ int s_go;
    __shared__ double s_x, s_y;
    __shared__ double s_red[NVB_STEP_THREADS / 32];

    if (tid == 0) {
        // angle_familiarity[k] = maxfam - diff (util.pyx:73, NavBySceneFamiliarity.py:313);
        // first maximum wins (:315)
        int best = 0;
        double best_fam = 0.0;
        for (int k = 0; k < a.A; k++) {
            const double d = __longlong_as_double((long long)a.exact[(size_t)b * a.A + k]);
            const double fam = __dsub_rn(a.maxfam, d);
            if (logging && a.log_afam) a.log_afam[((size_t)t * a.B + b) * a.A + k] = fam;
            if (k == 0 || fam > best_fam) { best = k; best_fam = fam; }
        }
        const double ang0 = a.ag.poses[3 * b + 2];
        const double ang = nvb_pymod_pos(__dadd_rn(ang0, a.offsets[best]), NVB_TWO_PI);   // :317
        double sn, cs;
        sincos(ang, &sn, &cs);
        const double x = __dadd_rn(a.ag.poses[3 * b], __dmul_rn(a.step_size, cs));       // :319
        const double y = __dadd_rn(a.ag.poses[3 * b + 1], __dmul_rn(a.step_size, sn));   // :320
        a.ag.poses[3 * b] = x;
        a.ag.poses[3 * b + 1] = y;
        a.ag.poses[3 * b + 2] = ang;
        s_x = x;
        s_y = y;
        if (logging) {
            a.log_best[(size_t)t * a.B + b] = (int16_t)best;
            a.log_pose[((size_t)t * a.B + b) * 3] = x;
            a.log_pose[((size_t)t * a.B + b) * 3 + 1] = y;
            a.log_pose[((size_t)t * a.B + b) * 3 + 2] = ang;
            a.log_sfam[(size_t)t * a.B + b] = best_fam;
        }
        if (a.fake) a.ag.completed[b] += 1;
    }
    __syncthreads();
    if (a.fake) return;

    // update_error, :252-276.  min over sqrt(d2) == sqrt(min d2) (sqrt is monotone
    // and correctly rounded), so reduce d2 and take one sqrt.
    const double x = s_x, y = s_y;
    const double2 *path = reinterpret_cast<const double2 *>(a.path);
    double m = __longlong_as_double(0x7FF0000000000000ll);
    for (int n = tid; n < a.n_path; n += blockDim.x) {
        const double2 pt = __ldg(path + n);
        const double dx = __dsub_rn(pt.x, x), dy = __dsub_rn(pt.y, y);
        m = fmin(m, __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmin(m, __shfl_xor_sync(0xFFFFFFFFu, m, o));
    if ((tid & 31) == 0) s_red[tid >> 5] = m;
    __syncthreads();
    m = s_red[0];
#pragma unroll
    for (int wq = 1; wq < NVB_STEP_THREADS / 32; wq++) m = fmin(m, s_red[wq]);
    const double dmin = __dsqrt_rn(m);
    const double thr = __dmul_rn(a.coverage_factor, a.step_size);   // :271
    if (tid == 0) {
        int go = 1;
        a.ag.nav_frames[b] += 1;                                    // :253
        if (dmin > a.max_dist) {                                    // :263-264
            a.ag.status[b] = -1;
            go = 0;
        } else {
            a.ag.err_sum[b] = __dadd_rn(a.ag.err_sum[b], __dmul_rn(dmin, dmin));   // :267
            a.ag.err_n[b] += 1;                                                    // :268
        }
        s_go = go;
    }
    __syncthreads();
    if (!s_go) return;
    if (dmin <= thr) {                                              // :272-276
        for (int n = tid; n < a.n_path; n += blockDim.x) {
            const double2 pt = __ldg(path + n);
            const double dx = __dsub_rn(pt.x, x), dy = __dsub_rn(pt.y, y);
            const double d = __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
            if (d <= thr) a.ag.coverage[(size_t)b * a.n_path + n] = 1;
        }
    }
    if (tid == 0) {
        // :328 end-of-path test
        const double2 pe = __ldg(path + (a.n_path - 1));
        const double ex = __dsub_rn(pe.x, x), ey = __dsub_rn(pe.y, y);
        const double de = __dsqrt_rn(__dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey)));
        if (de <= __dmul_rn(a.threshold_factor, a.step_size))
            a.ag.status[b] = 1;
        else
            a.ag.completed[b] += 1;
    }
}

// log entry of an agent that does not take part in this step
__device__ __forceinline__ void nvb_log_idle(const StepArgs &a, int b)
{
    const int tid = threadIdx.x;
    const int t = *a.step_counter;
    if (!(t >= 0 && t < a.log_cap)) return;
    const double nan = __longlong_as_double(0x7FF8000000000000ll);
    if (tid == 0) {
        a.log_best[(size_t)t * a.B + b] = -1;
        for (int q = 0; q < 3; q++) a.log_pose[((size_t)t * a.B + b) * 3 + q] = a.ag.poses[3 * b + q];
        a.log_sfam[(size_t)t * a.B + b] = nan;
    }
    if (a.log_afam)
        for (int k = tid; k < a.A; k += blockDim.x) a.log_afam[((size_t)t * a.B + b) * a.A + k] = nan;
}

__device__ __forceinline__ void nvb_load_div255(const StepArgs &a, double *s_div)
{
    for (int k = threadIdx.x; k < 256; k += blockDim.x) s_div[k] = a.div255[k];
    __syncthreads();
}

// ---- one launch: decide + ties + move ------------------------------------------
__global__ void __launch_bounds__(NVB_STEP_THREADS)
k3_step(StepArgs a)
{
    __shared__ double s_div[256];
    const int b = blockIdx.x;
    // an agent K1 stopped in this step (out of bounds / index error) is no longer active
    if (!nvb_agent_active(a.ag, b)) {
        nvb_log_idle(a, b);
        return;
    }
    nvb_load_div255(a, s_div);
    nvb_decide<true>(a, b, s_div);
    __syncthreads();
    nvb_move(a, b);
}

// ---- three launches (large / view-sharded libraries) -----------------------------
__global__ void __launch_bounds__(NVB_STEP_THREADS)
k3_decide(StepArgs a)
{
    __shared__ double s_div[256];
    const int b = blockIdx.x;
    const bool active = nvb_agent_active(a.ag, b);
    if (threadIdx.x == 0) a.ag.stepped[b] = active ? 1 : 0;
    if (!active) return;
    nvb_load_div255(a, s_div);
    nvb_decide<false>(a, b, s_div);
}

// Tie pass: every (tied glimpse, local view) pair whose score is within the
// band gets its exact FP64 difference; min per glimpse.
#define NVB_TIE_THREADS 256
__global__ void __launch_bounds__(NVB_TIE_THREADS)
k3_ties(StepArgs a)
{
    const int n_items = *a.tie_count;
    if (n_items == 0) return;
    const int chunks = (a.N + NVB_TIE_THREADS - 1) / NVB_TIE_THREADS;
    const long long units = (long long)n_items * chunks;
    for (long long u = blockIdx.x; u < units; u += gridDim.x) {
        const int item = (int)(u / chunks), ch = (int)(u - (long long)item * chunks);
        const int g = a.tie_items[item].x;
        const unsigned long long thr = a.tie_thr[item];
        const int v = ch * NVB_TIE_THREADS + threadIdx.x;
        if (v >= a.N) continue;
        const size_t qo = (size_t)g * a.Ppad, fo = (size_t)v * a.Ppad;
        if (nvb_pair_score(a, qo, fo) <= thr) {
            const double d = nvb_exact_diff_rows(a, qo, fo, a.div255);
            atomicMin(a.exact + g, (unsigned long long)__double_as_longlong(d));
        }
    }
}

__global__ void __launch_bounds__(NVB_STEP_THREADS)
k3_move(StepArgs a)
{
    const int b = blockIdx.x;
    if (!a.ag.stepped[b]) {
        nvb_log_idle(a, b);
        return;
    }
    nvb_move(a, b);
}
