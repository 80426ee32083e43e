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
#include "sampler.cuh"

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

// ---- MIN exchange over NVLink peer memory (view-sharded library, one process per GPU) ----
// Every rank owns an exchange area that its peers have mapped through CUDA IPC:
//   flags[2][world][B]      sequence number of the last push of (kind, source rank, agent)
//   data [2][world][cap]    the pushed values, cap >= B * A
// kind 0 = packed keys (exchanged in the prologue of decide), kind 1 = exact differences (in
// the prologue of the move).  The exchange is PUSH based and per agent: the CTA of agent b
// stores its A values and then a flag into EVERY PEER's area (P2P stores over NVLink), polls
// its OWN area until the same agent's flags of all peers have arrived (local loads, bounded
// spin: a missing peer becomes an error flag, not a hang) and takes the minimum of the values
// that were pushed to it -- one NVLink store latency per exchange, no separate launch, no
// round trip.  One buffer per kind suffices: a rank pushes the keys of step t+1 only after its
// move of step t, which waited for every peer's exact differences of step t, which a peer
// pushes after its decide of step t has finished reading the keys of step t (and likewise for
// the other kind).  The sequence number is 2 * epoch + kind + 1 with a device-resident epoch
// that the distance kernel bumps on every launch and nothing ever resets, so a whole run of
// sharded steps replays as a CUDA graph without host round trips.
// All ranks must queue the same sequence of step calls (they hold identical agent states).
#define NVB_P2P_MAX_RANKS 8

struct P2PArgs {
    unsigned long long *self;                       // this rank's area
    unsigned long long *peer[NVB_P2P_MAX_RANKS];    // every rank's area as mapped here (peer[rank] == self)
    int rank, world, B;
    long long cap;
    const unsigned long long *epoch;                // device: distance-kernel launches so far
    int *error;                                     // device: set to 1 if a peer did not show up in time
    long long spin_limit;                           // clock64 ticks
};

__device__ __forceinline__ unsigned long long *nvb_p2p_flags(const P2PArgs &x, unsigned long long *area, int kind, int src)
{
    return area + ((size_t)kind * x.world + src) * x.B;
}
__device__ __forceinline__ unsigned long long *nvb_p2p_data(const P2PArgs &x, unsigned long long *area, int kind, int src)
{
    return area + (size_t)2 * x.world * x.B + ((size_t)kind * x.world + src) * x.cap;
}

// values[b * A .. b * A + A) := MIN over ranks, for the agent of this CTA.  Called by every
// thread of the CTA; values are this rank's global buffer (keys or exact).  world <= 1: no-op.
__device__ __forceinline__ void nvb_p2p_min_agent(const P2PArgs &x, unsigned long long *values, int b, int A, int kind)
{
    if (x.world <= 1) return;
    const int tid = threadIdx.x;
    const unsigned long long seq = 2ull * (*x.epoch) + (unsigned long long)kind + 1ull;
    // push: my values of this agent into every peer's area, then (after a system-wide fence,
    // by one thread per peer) the flag
    for (int k = tid; k < A; k += blockDim.x) {
        const unsigned long long v = values[(size_t)b * A + k];
        for (int p = 0; p < x.world; p++)
            if (p != x.rank) nvb_p2p_data(x, x.peer[p], kind, x.rank)[(size_t)b * A + k] = v;
    }
    __threadfence_system();
    __syncthreads();
    if (tid < x.world && tid != x.rank) {
        *(volatile unsigned long long *)(nvb_p2p_flags(x, x.peer[tid], kind, x.rank) + b) = seq;
        // wait for the same agent's push of peer `tid` to land in MY area
        volatile unsigned long long *mine = nvb_p2p_flags(x, x.self, kind, tid) + b;
        const long long t0 = clock64();
        while (*mine < seq) {
            if (clock64() - t0 > x.spin_limit) { *x.error = 1; break; }
        }
        __threadfence_system();
    }
    __syncthreads();
    for (int k = tid; k < A; k += blockDim.x) {
        unsigned long long m = values[(size_t)b * A + k];
        for (int p = 0; p < x.world; p++) {
            if (p == x.rank) continue;
            const unsigned long long v = __ldcv(nvb_p2p_data(x, x.self, kind, p) + (size_t)b * A + k);
            m = v < m ? v : m;
        }
        values[(size_t)b * A + k] = m;
    }
    __syncthreads();
}

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
    const double *div255;        // [256] k / 255.
    double maxfam;               // H*W
    double step_size, max_dist, threshold_factor, coverage_factor;
    int fake;
    // tie work list (three-launch form)
    int *tie_count;
    int2 *tie_items;             // (glimpse index, unused)
    unsigned long long *tie_thr; // per item: largest score still treated as tied
    int *tie_next;               // per item: next unclaimed view chunk (k3_decide_help)
    int *tie_ready;              // per item: step index + 1 once the item is published
    // log
    const int *step_counter;     // device step index
    int log_cap;
    int16_t *log_best;           // [cap][B]
    double *log_pose;            // [cap][B][3]
    double *log_sfam;            // [cap][B]
    double *log_afam;            // [cap][B][A] or nullptr
    int32_t *pending_fail;       // [B] failure the sampler found for the NEXT step, or nullptr
    long long *dbg;              // tuning aid (see SamplerArgs::dbg), else nullptr
    unsigned long long *dmin2;   // [B] squared distance to the path (FP64 bits), long-path form
    int pdl_early;               // trigger the dependent launch at the top of every kernel
    int no_sample;               // k3_step_tm: do not sample the next glimpses (host-driven calls: the host may move the agents first)
    long long *tl;               // tuning aid: timeline stamps (nvb_tl_stamp) or nullptr
    const double *pblk;          // [ceil(n_path / NVB_PATH_BLOCK)][4] bounding circles of path blocks, or nullptr
    const double *pblk2;         // [ceil(blocks / NVB_PATH_GROUP)][4] bounding circles of groups of blocks (long paths), or nullptr
    // host-driven form without copy operations: results of the step written straight into the
    // caller's page-locked buffers (device-mapped); each nullptr when not wanted
    int16_t *out_best;           // [B]
    double *out_poses;           // [B][3]
    double *out_sfam;            // [B]
    double cover_thr2;           // largest double whose sqrt is <= coverage_factor * step_size (host)
    P2PArgs p2p;                 // exchange descriptor (view shards over NVLink); world == 0: none
    // single-launch step (step_tm.cuh; tensor-core distance kernel, TILEMIN): [B*A][n_vt] the
    // smallest tile-local key (-256 * dot + column; 0x7FFFFFFF = none) of every glimpse and view tile
    const int2 *tmin;
    int n_vt, sad_const;         // view tiles of NVB_TC_NT views; C of SAD = (C - dot) / 2
    const float *pblk_f;         // [blocks][4] FP32 copy of pblk, radius rounded up (step_tm.cuh prefilter)
};

#define NVB_PATH_BLOCK 16      /* path points per bounding circle (update_error prefilter) */
#define NVB_PATH_LIVE_MAX 1024  /* blocks the prefilter can list */
#define NVB_PATH_GROUP 64       /* blocks per second-level group (long paths) */
#define NVB_PATH_GRP_MAX 16     /* second-level groups the prefilter can list */
#define NVB_STEP_THREADS 128   /* == NVB_SAMPLER_THREADS: k31_step_sample runs both bodies */
#define NVB_STEP_MAX_A_SMEM 512 /* headings whose exact differences are kept in shared memory */
#define NVB_TIE_Q_CHUNKS 64     /* fused tie scan: glimpse rows up to 1 KB are staged in shared memory */
#define NVB_TIE_VPT 6           /* fused tie scan: views per thread in flight */

// {k / 255.} for k in 0..255 (util.pyx:71): a.div255 in global memory, staged per CTA in
// shared memory (the lookups are divergent: constant memory would serialise them)

__device__ __forceinline__ bool nvb_agent_active(const AgentState &ag, int b)
{
    return ag.status[b] == 0 && ag.completed[b] < ag.budget[b];
}

// Exact FP64 difference with 16-byte row loads (rows are 16-B aligned and zero
// padded; a zero pad pixel adds +0.0, which leaves the sum unchanged).
__device__ __forceinline__ double nvb_exact_diff_rows(const StepArgs &a, size_t qo, size_t fo,
                                                      const double *c_div255)
{
    if (a.cw != 0.0)
        return nvb_exact_diff(a.gh + qo, a.gs + qo, a.gv + qo, a.lh + fo, a.ls + fo, a.lv + fo, a.P,
                              a.cw, c_div255);
    const uint4 *q = reinterpret_cast<const uint4 *>(a.gv + qo);
    const uint4 *f = reinterpret_cast<const uint4 *>(a.lv + fo);
    double diff = 0.0;
    const int nc = a.Ppad / 16;
    for (int c0 = 0; c0 < nc; c0 += 4) {
        // up to four 16-B chunks of both rows in flight before the dependent FP64 chain
        uint4 qq[4], ff[4];
#pragma unroll
        for (int u = 0; u < 4; u++)
            if (c0 + u < nc) { qq[u] = q[c0 + u]; ff[u] = __ldg(f + c0 + u); }
#pragma unroll
        for (int u = 0; u < 4; u++)
            if (c0 + u < nc) {
                const uint32_t d[4] = {__vabsdiffu4(qq[u].x, ff[u].x), __vabsdiffu4(qq[u].y, ff[u].y),
                                       __vabsdiffu4(qq[u].z, ff[u].z), __vabsdiffu4(qq[u].w, ff[u].w)};
#pragma unroll
                for (int w = 0; w < 4; w++)
#pragma unroll
                    for (int k = 0; k < 4; k++) diff = __dadd_rn(diff, c_div255[(d[w] >> (8 * k)) & 0xFFu]);
            }
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
        // rows are 16-byte aligned and zero padded: 16-byte loads, up to five chunks of both
        // rows (an 80-pixel sensor) in flight at once
        const uint4 *q4 = reinterpret_cast<const uint4 *>(qv);
        const uint4 *f4 = reinterpret_cast<const uint4 *>(fv);
        const int nc = a.Ppad / 16;
        uint32_t s = 0;
        for (int c0 = 0; c0 < nc; c0 += 5) {
            uint4 qq[5], ff[5];
#pragma unroll
            for (int u = 0; u < 5; u++)
                if (c0 + u < nc) { qq[u] = q4[c0 + u]; ff[u] = __ldg(f4 + c0 + u); }
#pragma unroll
            for (int u = 0; u < 5; u++)
                if (c0 + u < nc) {
                    s = nvb_sad4(qq[u].x, ff[u].x, s); s = nvb_sad4(qq[u].y, ff[u].y, s);
                    s = nvb_sad4(qq[u].z, ff[u].z, s); s = nvb_sad4(qq[u].w, ff[u].w, s);
                }
        }
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
// headings right away; otherwise they go to the work list of k3_ties.  The exact
// differences end up in a.exact (global) and, for A <= NVB_STEP_MAX_A_SMEM, in
// s_exact (shared) for the move that follows in the same CTA.
template <bool FUSED>
__device__ __forceinline__ void nvb_decide(const StepArgs &a, int b, unsigned long long *s_exact,
                                           const double *div255)
{
    const int tid = threadIdx.x;
    __shared__ unsigned long long s_min;
    __shared__ int s_ntied;
    const unsigned long long idx_mask = (1ull << a.idx_bits) - 1ull;
    const bool small = a.A <= 32;
    unsigned long long thr;
    bool have_ties;
    if (small) {
        // one warp holds every key: minimum and tie count by shuffles, broadcast through
        // shared memory with a single barrier
        if (tid < 32) {
            const unsigned long long sc =
                (tid < a.A) ? (a.keys[(size_t)b * a.A + tid] >> a.idx_bits) : ~0ull;
            unsigned long long m = sc;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) m = min(m, __shfl_xor_sync(0xFFFFFFFFu, m, o));
            const unsigned nt = __popc(__ballot_sync(0xFFFFFFFFu, sc <= m + a.band));
            if (tid == 0) { s_min = m; s_ntied = (int)nt; }
        }
        __syncthreads();
    } else {
        if (tid == 0) { s_min = ~0ull; s_ntied = 0; }
        __syncthreads();
        unsigned long long local = ~0ull;
        for (int k = tid; k < a.A; k += blockDim.x)
            local = min(local, a.keys[(size_t)b * a.A + k] >> a.idx_bits);
        if (local != ~0ull) atomicMin(&s_min, local);
        __syncthreads();
        const unsigned long long t = s_min + a.band;
        for (int k = tid; k < a.A; k += blockDim.x)
            if ((a.keys[(size_t)b * a.A + k] >> a.idx_bits) <= t) atomicAdd(&s_ntied, 1);
        __syncthreads();
    }
    thr = s_min + a.band;
    have_ties = s_ntied > 1;
    if (!FUSED && have_ties && tid == 0) a.ag.stepped[b] = 2;   // takes part AND waits for the tie pass
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
        if (FUSED && k < NVB_STEP_MAX_A_SMEM) s_exact[k] = ebits;
        if (!FUSED || k >= NVB_STEP_MAX_A_SMEM || have_ties) a.exact[g] = ebits;
        if (!FUSED && tied) {
            const int slot = atomicAdd(a.tie_count, 1);
            a.tie_items[slot] = make_int2((int)g, 0);
            a.tie_thr[slot] = thr;
        }
    }
    if (FUSED && have_ties) {
        __syncthreads();   // exact[] initialised
        __shared__ uint4 s_q[NVB_TIE_Q_CHUNKS];
        const int nc = a.Ppad / 16;
        const bool fast = (a.cw == 0.0) && nc <= NVB_TIE_Q_CHUNKS;
        for (int k = 0; k < a.A; k++) {
            const size_t g = (size_t)b * a.A + k;
            if ((a.keys[g] >> a.idx_bits) > thr) continue;   // CTA-uniform
            if (fast) {
                // glimpse row in shared memory (broadcast reads); two views per thread in
                // flight, 16-byte loads: the scan is bound by L2 latency, not arithmetic
                __syncthreads();
                for (int c = tid; c < nc; c += blockDim.x) s_q[c] = reinterpret_cast<const uint4 *>(a.gv + g * a.Ppad)[c];
                __syncthreads();
                // NVB_TIE_VPT views per thread at a time, chunk by chunk: every load of a
                // chunk is in flight before the first sum needs it
                for (int v0 = 0; v0 < a.N; v0 += NVB_TIE_VPT * NVB_STEP_THREADS) {
                    uint32_t sum[NVB_TIE_VPT];
                    const uint4 *fp[NVB_TIE_VPT];
#pragma unroll
                    for (int u = 0; u < NVB_TIE_VPT; u++) {
                        const int v = v0 + u * NVB_STEP_THREADS + tid;
                        fp[u] = reinterpret_cast<const uint4 *>(a.lv + (size_t)(v < a.N ? v : 0) * a.Ppad);
                        sum[u] = 0;
                    }
                    for (int c = 0; c < nc; c++) {
                        uint4 x[NVB_TIE_VPT];
#pragma unroll
                        for (int u = 0; u < NVB_TIE_VPT; u++) x[u] = __ldg(fp[u] + c);
                        const uint4 q = s_q[c];
#pragma unroll
                        for (int u = 0; u < NVB_TIE_VPT; u++) {
                            sum[u] = nvb_sad4(q.x, x[u].x, sum[u]); sum[u] = nvb_sad4(q.y, x[u].y, sum[u]);
                            sum[u] = nvb_sad4(q.z, x[u].z, sum[u]); sum[u] = nvb_sad4(q.w, x[u].w, sum[u]);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < NVB_TIE_VPT; u++) {
                        const int v = v0 + u * NVB_STEP_THREADS + tid;
                        if (v < a.N && sum[u] <= thr) {
                            if (a.dbg) atomicAdd((unsigned long long *)a.dbg + b * 8 + 7, 1ull);
                            const double d = nvb_exact_diff_rows(a, g * a.Ppad, (size_t)v * a.Ppad, div255);
                            atomicMin(a.exact + g, (unsigned long long)__double_as_longlong(d));
                        }
                    }
                }
            } else {
                for (int v = tid; v < a.N; v += blockDim.x) {
                    if (nvb_pair_score(a, g * a.Ppad, (size_t)v * a.Ppad) <= thr) {
                        const double d = nvb_exact_diff_rows(a, g * a.Ppad, (size_t)v * a.Ppad, div255);
                        atomicMin(a.exact + g, (unsigned long long)__double_as_longlong(d));
                    }
                }
            }
        }
        __syncthreads();
        for (int k = tid; k < a.A && k < NVB_STEP_MAX_A_SMEM; k += blockDim.x)
            if ((a.keys[(size_t)b * a.A + k] >> a.idx_bits) <= thr) s_exact[k] = a.exact[(size_t)b * a.A + k];
    }
}

// Row helpers of the tie passes (V plane only: chem_weight 0).
// (not inlined, handed plain pointers: they sit in front of register-tight kernels)
__device__ __noinline__ unsigned long long nvb_exact_bits(const uint8_t *qrow, const uint8_t *frow, int nc,
                                                          const double *div255)
{
    const uint4 *q = reinterpret_cast<const uint4 *>(qrow);
    const uint4 *f = reinterpret_cast<const uint4 *>(frow);
    double diff = 0.0;
    for (int c = 0; c < nc; c++) {
        const uint4 qq = q[c], ff = __ldg(f + c);
        const uint32_t d[4] = {__vabsdiffu4(qq.x, ff.x), __vabsdiffu4(qq.y, ff.y), __vabsdiffu4(qq.z, ff.z),
                               __vabsdiffu4(qq.w, ff.w)};
#pragma unroll
        for (int w = 0; w < 4; w++)
#pragma unroll
            for (int k = 0; k < 4; k++) diff = __dadd_rn(diff, div255[(d[w] >> (8 * k)) & 0xFFu]);
    }
    return (unsigned long long)__double_as_longlong(diff);
}
__device__ __noinline__ unsigned nvb_pair_score_v(const uint8_t *qrow, const uint8_t *frow, int nc)
{
    const uint4 *q = reinterpret_cast<const uint4 *>(qrow);
    const uint4 *f = reinterpret_cast<const uint4 *>(frow);
    uint32_t sum = 0;
    for (int c = 0; c < nc; c++) {
        const uint4 qq = q[c], ff = __ldg(f + c);
        sum = nvb_sad4(qq.x, ff.x, sum); sum = nvb_sad4(qq.y, ff.y, sum);
        sum = nvb_sad4(qq.z, ff.z, sum); sum = nvb_sad4(qq.w, ff.w, sum);
    }
    return sum;
}
// move: argmax, pose update, update_error, end test, log.  s_exact (shared, may be
// nullptr) / a.exact hold the exact differences.  Returns through *pose_out the
// new pose and whether the agent will take another step (status still 0 and
// budget left); every thread gets the same answer.
// MODE 0: everything.  Long training paths (n_path > NVB_PATH_SPLIT) are scanned grid-wide
// instead of by the agent's own CTA: MODE 1 stops after the pose update (k3_move_pose),
// k3_path_scan reduces the squared distances into a.dmin2, MODE 2 (k3_move_finish) resumes
// with the bookkeeping.
struct NvbNoHook {
    __device__ __forceinline__ void operator()(const double *) const {}
};

// This agent's inputs, fetched in ONE round trip to L2 (nvb_move_preload) before anything
// branches on them: thread 0 holds the counters and the pose, lane k of warp 0 the exact
// difference of heading k (sweeps of up to 32 headings).
struct MovePre {
    int nav_frames, err_n, completed, budget, t;
    double err_sum, px, py, ang0;
    unsigned long long eb;
};

__device__ __forceinline__ MovePre nvb_move_preload(const StepArgs &a, int b, const unsigned long long *s_exact)
{
    MovePre p;
    const int tid = threadIdx.x;
    p.t = *a.step_counter;
    p.nav_frames = p.err_n = p.completed = p.budget = 0;
    p.err_sum = p.px = p.py = p.ang0 = 0.0;
    p.eb = NVB_EXACT_NONE;
    if (tid == 0) {
        p.nav_frames = a.ag.nav_frames[b];
        p.err_n = a.ag.err_n[b];
        p.err_sum = a.ag.err_sum[b];
        p.completed = a.ag.completed[b];
        p.budget = a.ag.budget[b];
        p.px = a.ag.poses[3 * b]; p.py = a.ag.poses[3 * b + 1]; p.ang0 = a.ag.poses[3 * b + 2];
    }
    if (a.A <= 32 && tid < a.A) p.eb = (s_exact != nullptr) ? s_exact[tid] : a.exact[(size_t)b * a.A + tid];
    return p;
}

// AfterPose(pose): called by every thread once the new pose is published (shared memory).
// BesideScan(pose): called by warp 1 instead of its share of the path scan when W1_HOOK
// (the other warps scan the whole path meanwhile).
// offsets: the heading offsets (a.offsets, or a copy of them in shared memory).
template <int MODE, bool W1_HOOK = false, typename AfterPose = NvbNoHook, typename BesideScan = NvbNoHook>
__device__ __forceinline__ bool nvb_move(const StepArgs &a, int b, const unsigned long long *s_exact,
                                         const MovePre &pre, const double *offsets, double *pose_out,
                                         AfterPose after_pose = AfterPose(), BesideScan beside_scan = BesideScan())
{
    const int tid = threadIdx.x;
    const int t = pre.t;
    const bool logging = (t >= 0 && t < a.log_cap);
    __shared__ int s_go, s_more;
    __shared__ double s_pose[3];
    __shared__ double s_red[8];
    const int completed = pre.completed, budget = pre.budget;

    if (MODE == 2 && tid == 0) {
        s_pose[0] = pre.px; s_pose[1] = pre.py; s_pose[2] = pre.ang0;
    }
    if (MODE != 2 && tid < 32) {
        // angle_familiarity[k] = maxfam - diff (util.pyx:73, NavBySceneFamiliarity.py:313);
        // first maximum wins (:315)
        int best = 0x7FFFFFFF;
        double best_fam = __longlong_as_double(0xFFF0000000000000ll);   // -inf
        if (a.A <= 32) {
            // a short sweep: lane k holds heading k; every lane runs the same short compare chain
            const double fam = __dsub_rn(a.maxfam, __longlong_as_double((long long)pre.eb));
            if (logging && a.log_afam && tid < a.A) a.log_afam[((size_t)t * a.B + b) * a.A + tid] = fam;
            for (int k = 0; k < a.A; k++) {
                const double fk = __shfl_sync(0xFFFFFFFFu, fam, k);
                if (k == 0 || fk > best_fam) { best = k; best_fam = fk; }
            }
        } else {
            // warp 0: lane l looks at headings l, l + 32, ...; then a shuffle reduction
            for (int k = tid; k < a.A; k += 32) {
                // (ld.cg: other SMs may have lowered these with atomicMin during this kernel)
                const unsigned long long eb = (s_exact != nullptr && k < NVB_STEP_MAX_A_SMEM)
                                                  ? s_exact[k] : __ldcg(a.exact + (size_t)b * a.A + k);
                const double fam = __dsub_rn(a.maxfam, __longlong_as_double((long long)eb));
                if (logging && a.log_afam) a.log_afam[((size_t)t * a.B + b) * a.A + k] = fam;
                if (best == 0x7FFFFFFF || fam > best_fam) { best = k; best_fam = fam; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double of = __shfl_xor_sync(0xFFFFFFFFu, best_fam, o);
                const int ok = __shfl_xor_sync(0xFFFFFFFFu, best, o);
                if (ok != 0x7FFFFFFF && (best == 0x7FFFFFFF || of > best_fam || (of == best_fam && ok < best))) {
                    best = ok;
                    best_fam = of;
                }
            }
        }
        if (tid == 0) {
            const double ang = nvb_pymod_pos(__dadd_rn(pre.ang0, offsets[best]), NVB_TWO_PI);   // :317
            double sn, cs;
            nvb_glibc_sincos(ang, &sn, &cs);   // np.cos / np.sin = glibc's (:319-320)
            const double x = __dadd_rn(pre.px, __dmul_rn(a.step_size, cs));   // :319
            const double y = __dadd_rn(pre.py, __dmul_rn(a.step_size, sn));   // :320
            s_pose[0] = x; s_pose[1] = y; s_pose[2] = ang;
            a.ag.poses[3 * b] = x;
            a.ag.poses[3 * b + 1] = y;
            a.ag.poses[3 * b + 2] = ang;
            if (a.out_best != nullptr) a.out_best[b] = (int16_t)best;
            if (a.out_sfam != nullptr) a.out_sfam[b] = best_fam;
            if (a.out_poses != nullptr) { a.out_poses[3 * b] = x; a.out_poses[3 * b + 1] = y; a.out_poses[3 * b + 2] = ang; }
            if (logging) {
                a.log_best[(size_t)t * a.B + b] = (int16_t)best;
                a.log_pose[((size_t)t * a.B + b) * 3] = x;
                a.log_pose[((size_t)t * a.B + b) * 3 + 1] = y;
                a.log_pose[((size_t)t * a.B + b) * 3 + 2] = ang;
                a.log_sfam[(size_t)t * a.B + b] = best_fam;
            }
            if (a.fake) {
                a.ag.completed[b] = completed + 1;
                s_more = (completed + 1 < budget);
            }
        }
    }
    __syncthreads();
    pose_out[0] = s_pose[0]; pose_out[1] = s_pose[1]; pose_out[2] = s_pose[2];
    after_pose(pose_out);
    long long *dbg2 = a.dbg ? a.dbg + (size_t)8 * a.B + (size_t)8 * b : nullptr;
    if (dbg2 && tid == 0) dbg2[0] = clock64();    // window requested
    const bool warp1 = W1_HOOK && (tid >> 5) == 1;
    if (warp1) beside_scan(pose_out);
    if (dbg2 && tid == 32) dbg2[1] = clock64();   // warp 1 hook done
    if (a.fake) return s_more != 0;
    if (MODE == 1) {
        if (tid == 0) a.dmin2[b] = 0x7FF0000000000000ull;   // +inf: k3_path_scan takes the minimum
        return false;
    }

    // update_error, :252-276.  min over sqrt(d2) == sqrt(min d2) (sqrt is monotone and
    // correctly rounded), so reduce d2 and take one sqrt.  Coverage (:272-276) marks
    // every point with d <= thr when dmin <= thr; a point with d <= thr implies
    // dmin <= thr, so when thr <= max_dist (no TooFar possible then) it is marked in
    // the same pass, comparing d2 with thr2 = the largest double whose sqrt is <= thr.
    const double x = s_pose[0], y = s_pose[1];
    const double thr = __dmul_rn(a.coverage_factor, a.step_size);   // :271
    const double thr2 = a.cover_thr2;
    const bool one_pass = thr <= a.max_dist;
    const double2 *path = reinterpret_cast<const double2 *>(a.path);
    // the threads that scan: all of them, or all but warp 1 (busy with the hook)
    const int scan_n = W1_HOOK ? (int)blockDim.x - 32 : (int)blockDim.x;
    const int scan_id = !W1_HOOK ? tid : warp1 ? a.n_path : (tid < 32 ? tid : tid - 32);
    double m = __longlong_as_double(0x7FF0000000000000ll);
    if (MODE == 2) {
        m = __longlong_as_double((long long)a.dmin2[b]);
    } else if (a.pblk != nullptr && one_pass) {
        // Prefilter: a block of 16 consecutive path points lies inside a circle (c, r).  Its
        // points are no nearer than |p - c| - r and one of them is no farther than |p - c| + r.
        // With U = the smallest such upper bound, the nearest point and every point within
        // thr (coverage) sit in blocks whose lower bound is <= max(U, thr); only those blocks
        // are scanned, with exactly the arithmetic of the full scan, so the minimum and the
        // coverage marks are the same.  Long paths (more than NVB_PATH_LIVE_MAX blocks) get a
        // second level first: groups of NVB_PATH_GROUP blocks with their own circles
        // (a.pblk2), the same argument applied twice -- update_error over 10^6 path points then
        // costs this CTA what it costs over 10^3.  Lists that overflow (a degenerate path with
        // thousands of points on one spot near the agent) fall back to scanning everything.
        // (The scanning warps meet at a named barrier: warp 1 may be busy with the rotations.)
        __shared__ double s_ub[8];
        __shared__ int s_live[NVB_PATH_LIVE_MAX];
        __shared__ int s_grp[NVB_PATH_GRP_MAX];
        __shared__ int s_nlive, s_ngrp;
        const int n_blk = (a.n_path + NVB_PATH_BLOCK - 1) / NVB_PATH_BLOCK;
        const bool two_level = n_blk > NVB_PATH_LIVE_MAX && a.pblk2 != nullptr;
        const int n_grp = (n_blk + NVB_PATH_GROUP - 1) / NVB_PATH_GROUP;
        const bool scanning = !warp1;
        const int sid = (!W1_HOOK || tid < 32) ? tid : tid - 32;   // index among the scanning threads
        const int n_scan_warps = (int)(blockDim.x >> 5);
        bool full_scan = n_blk > NVB_PATH_LIVE_MAX && !two_level;
        // (cx, cy, r) of circle j of a level; lower / upper bound of the distances of its points
        auto circle = [&](const double *tab, int j, double &lo, double &hi) {
            const double2 cc = __ldg(reinterpret_cast<const double2 *>(tab) + 2 * j);
            const double2 cr = __ldg(reinterpret_cast<const double2 *>(tab) + 2 * j + 1);
            const double ex = __dsub_rn(cc.x, x), ey = __dsub_rn(cc.y, y);
            const double dc = __dsqrt_rn(__dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey)));
            hi = __dadd_rn(dc, cr.x);
            lo = __dsub_rn(__dsub_rn(dc, cr.x), 1e-9 * (1.0 + dc));   // slack far above the rounding of dc and of the comparison
        };
        // block-wide minimum of `v` over the scanning warps
        auto scan_min = [&](double v) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xFFFFFFFFu, v, o));
            asm volatile("bar.sync 1, %0;" ::"r"(scan_n) : "memory");   // s_ub free again
            if ((tid & 31) == 0) s_ub[tid >> 5] = v;
            asm volatile("bar.sync 1, %0;" ::"r"(scan_n) : "memory");
            double r = __longlong_as_double(0x7FF0000000000000ll);
            for (int wq = 0; wq < n_scan_warps; wq++)
                if (!(W1_HOOK && wq == 1)) r = fmin(r, s_ub[wq]);
            return r;
        };
        if (scanning && !full_scan) {
            if (sid == 0) { s_nlive = 0; s_ngrp = 0; }
            const double inf = __longlong_as_double(0x7FF0000000000000ll);
            double lo, hi;
            int n_cand = n_blk;      // candidate blocks: all of them, or those of the live groups
            if (two_level) {
                double ub = inf;
                for (int j = sid; j < n_grp; j += scan_n) { circle(a.pblk2, j, lo, hi); ub = fmin(ub, hi); }
                const double U2 = fmax(thr, scan_min(ub));
                for (int j = sid; j < n_grp; j += scan_n) {
                    circle(a.pblk2, j, lo, hi);
                    if (lo <= U2) { const int slot = atomicAdd(&s_ngrp, 1); if (slot < NVB_PATH_GRP_MAX) s_grp[slot] = j; }
                }
                asm volatile("bar.sync 1, %0;" ::"r"(scan_n) : "memory");
                if (s_ngrp > NVB_PATH_GRP_MAX) full_scan = true;
                n_cand = s_ngrp * NVB_PATH_GROUP;
            }
            if (!full_scan) {
                auto cand_block = [&](int idx) {
                    if (!two_level) return idx;
                    const int b1 = s_grp[idx / NVB_PATH_GROUP] * NVB_PATH_GROUP + (idx % NVB_PATH_GROUP);
                    return b1 < n_blk ? b1 : -1;
                };
                double ub = inf;
                for (int idx = sid; idx < n_cand; idx += scan_n) {
                    const int j = cand_block(idx);
                    if (j >= 0) { circle(a.pblk, j, lo, hi); ub = fmin(ub, hi); }
                }
                const double U = fmax(thr, scan_min(ub));
                for (int idx = sid; idx < n_cand; idx += scan_n) {
                    const int j = cand_block(idx);
                    if (j < 0) continue;
                    circle(a.pblk, j, lo, hi);
                    if (lo <= U) { const int slot = atomicAdd(&s_nlive, 1); if (slot < NVB_PATH_LIVE_MAX) s_live[slot] = j; }
                }
                asm volatile("bar.sync 1, %0;" ::"r"(scan_n) : "memory");
                if (s_nlive > NVB_PATH_LIVE_MAX) full_scan = true;
            }
            if (!full_scan) {
                const int n_live = s_nlive;
                for (int idx = sid; idx < n_live * NVB_PATH_BLOCK; idx += scan_n) {
                    const int n = s_live[idx / NVB_PATH_BLOCK] * NVB_PATH_BLOCK + (idx % NVB_PATH_BLOCK);
                    if (n < a.n_path) {
                        const double2 pt = __ldg(path + n);
                        const double dx = __dsub_rn(pt.x, x), dy = __dsub_rn(pt.y, y);
                        const double d2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
                        m = fmin(m, d2);
                        if (d2 <= thr2) a.ag.coverage[(size_t)b * a.n_path + n] = 1;
                    }
                }
            }
        }
        if (scanning && full_scan) {   // (CTA-uniform among the scanning threads)
            for (int n = sid; n < a.n_path; n += scan_n) {
                const double2 pt = __ldg(path + n);
                const double dx = __dsub_rn(pt.x, x), dy = __dsub_rn(pt.y, y);
                const double d2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
                m = fmin(m, d2);
                if (d2 <= thr2) a.ag.coverage[(size_t)b * a.n_path + n] = 1;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmin(m, __shfl_xor_sync(0xFFFFFFFFu, m, o));
        if ((tid & 31) == 0) s_red[tid >> 5] = m;
        __syncthreads();
        m = s_red[0];
        for (int wq = 1; wq < (int)(blockDim.x >> 5); wq++) m = fmin(m, s_red[wq]);
    } else {
#pragma unroll 4
        for (int n = scan_id; n < a.n_path; n += scan_n) {
            const double2 pt = __ldg(path + n);
            const double dx = __dsub_rn(pt.x, x), dy = __dsub_rn(pt.y, y);
            const double d2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
            m = fmin(m, d2);
            if (one_pass && d2 <= thr2) a.ag.coverage[(size_t)b * a.n_path + n] = 1;
        }
        if (dbg2 && tid == 0) dbg2[2] = clock64();    // warp 0 scanned its share
        if (dbg2 && tid == 64) dbg2[3] = clock64();   // warp 2 scanned its share
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmin(m, __shfl_xor_sync(0xFFFFFFFFu, m, o));
        if ((tid & 31) == 0) s_red[tid >> 5] = m;
        __syncthreads();
        if (dbg2 && tid == 0) dbg2[4] = clock64();    // reduced
        m = s_red[0];
        for (int wq = 1; wq < (int)(blockDim.x >> 5); wq++) m = fmin(m, s_red[wq]);
    }
    const double dmin = __dsqrt_rn(m);
    if (tid == 0) {
        int go = 1, more = 0;
        a.ag.nav_frames[b] = pre.nav_frames + 1;                    // :253
        if (dmin > a.max_dist) {                                    // :263-264
            a.ag.status[b] = -1;
            go = 0;
        } else {
            a.ag.err_sum[b] = __dadd_rn(pre.err_sum, __dmul_rn(dmin, dmin));   // :267
            a.ag.err_n[b] = pre.err_n + 1;                                     // :268
            // :328 end-of-path test
            const double2 pe = __ldg(path + (a.n_path - 1));
            const double ex = __dsub_rn(pe.x, x), ey = __dsub_rn(pe.y, y);
            const double de = __dsqrt_rn(__dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey)));
            if (de <= __dmul_rn(a.threshold_factor, a.step_size)) {
                a.ag.status[b] = 1;
            } else {
                a.ag.completed[b] = completed + 1;
                more = (completed + 1 < budget);
            }
        }
        s_go = go;
        s_more = more;
        if (dbg2) dbg2[5] = clock64();                // bookkeeping done
    }
    __syncthreads();
    if (s_go && !one_pass && dmin <= thr) {                         // :272-276, two-pass form
        for (int n = tid; n < a.n_path; n += blockDim.x) {
            const double2 pt = __ldg(path + n);
            const double dx = __dsub_rn(pt.x, x), dy = __dsub_rn(pt.y, y);
            if (__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)) <= thr2)
                a.ag.coverage[(size_t)b * a.n_path + n] = 1;
        }
    }
    return s_more != 0;
}

// log entry of an agent that does not take part in this step
__device__ __forceinline__ void nvb_log_idle(const StepArgs &a, int b)
{
    const int tid = threadIdx.x;
    const int t = *a.step_counter;
    const double nan = __longlong_as_double(0x7FF8000000000000ll);
    if (tid == 0) {
        if (a.out_best != nullptr) a.out_best[b] = -1;
        if (a.out_sfam != nullptr) a.out_sfam[b] = nan;
        if (a.out_poses != nullptr)
            for (int q = 0; q < 3; q++) a.out_poses[3 * b + q] = a.ag.poses[3 * b + q];
    }
    if (!(t >= 0 && t < a.log_cap)) return;
    if (tid == 0) {
        a.log_best[(size_t)t * a.B + b] = -1;
        for (int q = 0; q < 3; q++) a.log_pose[((size_t)t * a.B + b) * 3 + q] = a.ag.poses[3 * b + q];
        a.log_sfam[(size_t)t * a.B + b] = nan;
    }
    if (a.log_afam)
        for (int k = tid; k < a.A; k += blockDim.x) a.log_afam[((size_t)t * a.B + b) * a.A + k] = nan;
}

// A failure the sampler found for THIS step while it ran at the end of the previous
// launch (k31_step_sample) becomes the agent's status now, at the step it belongs to.
__device__ __forceinline__ void nvb_commit_pending(const StepArgs &a, int b)
{
    if (a.pending_fail != nullptr && threadIdx.x == 0) {
        const int pf = a.pending_fail[b];
        if (pf != 0) {
            a.ag.status[b] = pf;
            a.pending_fail[b] = 0;
        }
    }
    __syncthreads();
}

// ---- one launch: decide + ties + move ------------------------------------------
__global__ void __launch_bounds__(NVB_STEP_THREADS, 8)
k3_step(StepArgs a)
{
    if (a.pdl_early) nvb_grid_dep_launch();
    nvb_grid_dep_wait();
    __shared__ unsigned long long s_exact[NVB_STEP_MAX_A_SMEM];
    __shared__ double s_div[256];
    const int b = blockIdx.x;
    for (int k = threadIdx.x; k < 256; k += blockDim.x) s_div[k] = a.div255[k];
    nvb_commit_pending(a, b);
    // an agent the sampler stopped in this step (out of bounds / index error) is no longer active
    if (!nvb_agent_active(a.ag, b)) {
        nvb_log_idle(a, b);
        return;
    }
    nvb_decide<true>(a, b, s_exact, s_div);
    __syncthreads();
    double pose[3];
    const MovePre pre = nvb_move_preload(a, b, s_exact);
    nvb_move<0>(a, b, s_exact, pre, a.offsets, pose);
}

// ---- one launch: decide + ties + move of step t, then the glimpses of step t+1 ----
// (the sampler's failures for step t+1 are parked in a.pending_fail until then)
template <bool NEED_HS, int PH, int PW>
__global__ void __launch_bounds__(NVB_STEP_THREADS, 8)   // 1024 agents = 7 CTAs per SM: one wave
k31_step_sample(const __grid_constant__ CUtensorMap tmap, StepArgs a, SamplerArgs sa)
{
    if (a.pdl_early) nvb_grid_dep_launch();
    nvb_grid_dep_wait();
    extern __shared__ __align__(128) uint8_t smem_k31[];
    __shared__ unsigned long long s_exact[NVB_STEP_MAX_A_SMEM];
    __shared__ double s_div[256];
    const int b = blockIdx.x;
    if (sa.dbg && threadIdx.x == 0) sa.dbg[b * 8 + 0] = clock64();
    for (int k = threadIdx.x; k < 256; k += blockDim.x) s_div[k] = a.div255[k];
    nvb_commit_pending(a, b);
    if (!nvb_agent_active(a.ag, b)) {
        nvb_log_idle(a, b);
        return;
    }
    if (sa.dbg && threadIdx.x == 0) sa.dbg[b * 8 + 1] = clock64();
    nvb_decide<true>(a, b, s_exact, s_div);
    __syncthreads();
    if (sa.dbg && threadIdx.x == 0) sa.dbg[b * 8 + 2] = clock64();
    double pose[3];
    const MovePre pre = nvb_move_preload(a, b, s_exact);
    const bool more = nvb_move<0>(a, b, s_exact, pre, a.offsets, pose);
    if (sa.dbg && threadIdx.x == 0) sa.dbg[b * 8 + 3] = clock64();
    if (!more) return;
    __syncthreads();
    nvb_sample_body<NEED_HS, PH, PW>(&tmap, sa, b, pose[0], pose[1], pose[2], smem_k31, a.pending_fail + b, 0, sa.A);
}

// ---- two launches: [decide + cooperative tie scan] then [move + sample] ------------
// Ties hit a few agents per step, but each needs a rescan of the whole library for
// its tied headings: far too much for one CTA while a thousand others sit idle.
// So every CTA first decides for its own agent, publishes its tied headings to a
// queue, and then helps: view chunks of queued items are claimed with an atomic
// counter by whichever CTA gets there.  The owner always sweeps its own items to
// the end, so every chunk is scanned by the time the kernel completes; the kernel
// boundary is the barrier before k3_move_sample reads the exact differences.
#define NVB_HELP_CHUNK (2 * NVB_STEP_THREADS)   /* views per claimed chunk */
#define NVB_HELP_BUDGET 2                       /* chunks a helping CTA scans at most */

__device__ __forceinline__ void nvb_tie_scan_chunk(const StepArgs &a, int g, unsigned long long thr,
                                                   int chunk, const double *div255)
{
    const int nc = a.Ppad / 16;
    const size_t qo = (size_t)g * a.Ppad;
    const int v0 = chunk * NVB_HELP_CHUNK + threadIdx.x, v1 = v0 + NVB_STEP_THREADS;
    if (a.cw == 0.0) {
        const uint4 *q = reinterpret_cast<const uint4 *>(a.gv + qo);
        const uint4 *f0 = reinterpret_cast<const uint4 *>(a.lv + (size_t)(v0 < a.N ? v0 : 0) * a.Ppad);
        const uint4 *f1 = reinterpret_cast<const uint4 *>(a.lv + (size_t)(v1 < a.N ? v1 : 0) * a.Ppad);
        uint32_t s0 = 0, s1 = 0;
        for (int c = 0; c < nc; c++) {
            const uint4 qq = __ldg(q + c), x0 = __ldg(f0 + c), x1 = __ldg(f1 + c);
            s0 = nvb_sad4(qq.x, x0.x, s0); s0 = nvb_sad4(qq.y, x0.y, s0);
            s0 = nvb_sad4(qq.z, x0.z, s0); s0 = nvb_sad4(qq.w, x0.w, s0);
            s1 = nvb_sad4(qq.x, x1.x, s1); s1 = nvb_sad4(qq.y, x1.y, s1);
            s1 = nvb_sad4(qq.z, x1.z, s1); s1 = nvb_sad4(qq.w, x1.w, s1);
        }
        if (v0 < a.N && s0 <= thr)
            atomicMin(a.exact + g, (unsigned long long)__double_as_longlong(
                                       nvb_exact_diff_rows(a, qo, (size_t)v0 * a.Ppad, div255)));
        if (v1 < a.N && s1 <= thr)
            atomicMin(a.exact + g, (unsigned long long)__double_as_longlong(
                                       nvb_exact_diff_rows(a, qo, (size_t)v1 * a.Ppad, div255)));
    } else {
        for (int v = v0; v < a.N && v <= v1; v += NVB_STEP_THREADS)
            if (nvb_pair_score(a, qo, (size_t)v * a.Ppad) <= thr)
                atomicMin(a.exact + g, (unsigned long long)__double_as_longlong(
                                           nvb_exact_diff_rows(a, qo, (size_t)v * a.Ppad, div255)));
    }
}

// claims and scans chunks of item `slot` until none is left or `budget` chunks are done;
// returns the number of chunks scanned
__device__ __forceinline__ int nvb_tie_sweep(const StepArgs &a, int slot, int n_chunks, int budget,
                                             const double *div255)
{
    __shared__ int s_chunk;
    const int g = a.tie_items[slot].x;
    const unsigned long long thr = a.tie_thr[slot];
    int done = 0;
    while (done < budget) {
        __syncthreads();
        if (threadIdx.x == 0) s_chunk = atomicAdd(a.tie_next + slot, 1);
        __syncthreads();
        const int c = s_chunk;
        if (c >= n_chunks) break;
        nvb_tie_scan_chunk(a, g, thr, c, div255);
        done++;
    }
    return done;
}

__global__ void __launch_bounds__(NVB_STEP_THREADS, 8)
k3_decide_help(StepArgs a)
{
    if (a.pdl_early) nvb_grid_dep_launch();
    nvb_grid_dep_wait();
    __shared__ double s_div[256];
    __shared__ int s_slot0, s_nslots, s_count, s_go;
    const int b = blockIdx.x, tid = threadIdx.x;
    const int epoch = *a.step_counter + 1;
    const int n_chunks = (a.N + NVB_HELP_CHUNK - 1) / NVB_HELP_CHUNK;
    for (int k = tid; k < 256; k += blockDim.x) s_div[k] = a.div255[k];
    if (tid == 0) { s_slot0 = 0; s_nslots = 0; }
    nvb_commit_pending(a, b);
    const bool active = nvb_agent_active(a.ag, b);
    if (tid == 0) a.ag.stepped[b] = active ? 1 : 0;
    if (active) {
        // --- decide (as nvb_decide<false>, with the tied headings of this agent in one
        //     contiguous run of queue slots)
        __shared__ unsigned long long s_min;
        __shared__ int s_ntied;
        const unsigned long long idx_mask = (1ull << a.idx_bits) - 1ull;
        if (tid == 0) { s_min = ~0ull; s_ntied = 0; }
        __syncthreads();
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
        __syncthreads();   // everyone has read s_ntied before it is reused
        if (have_ties && tid == 0) {
            s_nslots = s_ntied;
            s_slot0 = atomicAdd(a.tie_count, s_ntied);
            s_ntied = 0;   // reused as the running slot offset below
        }
        __syncthreads();
        for (int k = tid; k < a.A; k += blockDim.x) {
            const size_t g = (size_t)b * a.A + k;
            const unsigned long long key = a.keys[g];
            const bool tied = have_ties && (key >> a.idx_bits) <= thr;
            unsigned long long ebits = NVB_EXACT_NONE;
            if (!tied) {   // the tie scan revisits the best view of a tied heading anyway
                const long long v = (long long)(key & idx_mask) - a.view_offset;
                if (key != NVB_KEY_NONE && v >= 0 && v < a.N)
                    ebits = (unsigned long long)__double_as_longlong(
                        nvb_exact_diff_rows(a, g * a.Ppad, (size_t)v * a.Ppad, s_div));
            }
            a.exact[g] = ebits;
            if (tied) {
                const int slot = s_slot0 + atomicAdd(&s_ntied, 1);
                a.tie_items[slot] = make_int2((int)g, 0);
                a.tie_thr[slot] = thr;
                a.tie_next[slot] = 0;
                __threadfence();
                *(volatile int *)(a.tie_ready + slot) = epoch;   // published
            }
        }
        __syncthreads();
        // --- the owner sweeps its own items to the end
        for (int i = 0; i < s_nslots; i++) nvb_tie_sweep(a, s_slot0 + i, n_chunks, 1 << 30, s_div);
    }
    // --- everybody helps once with what is published by now: one parallel look at the
    //     queue (a thread per slot), then at most NVB_HELP_BUDGET chunks, starting at a
    //     CTA-dependent item so that the helpers spread over the queue
    __shared__ int s_list[NVB_STEP_THREADS];
    __syncthreads();
    if (tid == 0) { s_count = *(volatile int *)a.tie_count; s_go = 0; }
    __syncthreads();
    const int count = min(s_count, NVB_STEP_THREADS);
    if (count == 0) return;
    if (tid < count && !(tid >= s_slot0 && tid < s_slot0 + s_nslots) &&
        *(volatile int *)(a.tie_ready + tid) == epoch && *(volatile int *)(a.tie_next + tid) < n_chunks)
        s_list[atomicAdd(&s_go, 1)] = tid;
    __syncthreads();
    const int n_list = s_go;
    if (n_list == 0) return;
    __threadfence();
    // at most NVB_HELP_BUDGET chunks and NVB_HELP_BUDGET + 1 claim attempts: an item that
    // looked unfinished may have been drained by others in the meantime
    int budget = NVB_HELP_BUDGET;
    for (int i = 0; i < n_list && i <= NVB_HELP_BUDGET && budget > 0; i++)
        budget -= nvb_tie_sweep(a, s_list[(i + b) % n_list], n_chunks, budget, s_div);
}

#define NVB_MS_MAX_THREADS 160   /* k3_move_sample runs with 128 or 160 threads (see launch_k3ms_t) */

// ---- the tie pass folded into move+sample (TIES) ---------------------------------------
// k3_decide has listed every (tied heading) item.  Every CTA of move+sample first scores its
// share of the (item, view chunk) units -- unit u belongs to CTA u mod gridDim -- and adds
// the number it did to tie_count[1]; only the agents that HAVE ties wait until every unit is
// done.  Nobody waits before finishing its own share, and the engine uses this form only
// when the whole grid is co-resident, so the wait is short; should it ever last too long
// (another engine's kernels holding SM slots) the waiting CTA scans its own items alone:
// atomicMin is idempotent, the result is the same.
__device__ __forceinline__ void nvb_tie_unit(const StepArgs &a, int g, unsigned long long thr, int v)
{
    if (v >= a.N) return;
    const size_t qo = (size_t)g * a.Ppad, fo = (size_t)v * a.Ppad;
    unsigned long long score;
    if (a.cw == 0.0) {
        // as nvb_pair_score, three chunks of both rows in flight (this runs inside a kernel
        // compiled for 56 registers)
        const uint4 *q4 = reinterpret_cast<const uint4 *>(a.gv + qo);
        const uint4 *f4 = reinterpret_cast<const uint4 *>(a.lv + fo);
        const int nc = a.Ppad / 16;
        uint32_t s = 0;
        for (int c0 = 0; c0 < nc; c0 += 3) {
            uint4 qq[3], ff[3];
#pragma unroll
            for (int u = 0; u < 3; u++)
                if (c0 + u < nc) { qq[u] = q4[c0 + u]; ff[u] = __ldg(f4 + c0 + u); }
#pragma unroll
            for (int u = 0; u < 3; u++)
                if (c0 + u < nc) {
                    s = nvb_sad4(qq[u].x, ff[u].x, s); s = nvb_sad4(qq[u].y, ff[u].y, s);
                    s = nvb_sad4(qq[u].z, ff[u].z, s); s = nvb_sad4(qq[u].w, ff[u].w, s);
                }
        }
        score = s;
    } else {
        score = nvb_pair_score(a, qo, fo);
    }
    if (score <= thr) {
        const double d = nvb_exact_diff_rows(a, qo, fo, a.div255);
        atomicMin(a.exact + g, (unsigned long long)__double_as_longlong(d));
    }
}

__device__ __forceinline__ void nvb_tie_help(const StepArgs &a, int n_items)
{
    const int T = (int)blockDim.x, chunks = (a.N + T - 1) / T;
    const long long units = (long long)n_items * chunks;
    int mine = 0;
    for (long long u = blockIdx.x; u < units; u += gridDim.x) {
        const int item = (int)(u / chunks), ch = (int)(u - (long long)item * chunks);
        nvb_tie_unit(a, a.tie_items[item].x, a.tie_thr[item], ch * T + (int)threadIdx.x);
        mine++;
    }
    if (mine) {   // CTA-uniform
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) atomicAdd(a.tie_count + 1, mine);
    }
}

__device__ __forceinline__ void nvb_tie_wait(const StepArgs &a, int b, int n_items)
{
    __shared__ int s_tie_ok;
    const int T = (int)blockDim.x, chunks = (a.N + T - 1) / T;
    const long long units = (long long)n_items * chunks;
    if (threadIdx.x == 0) {
        // patience: ~50 us plus ~5 us per round of units (clock64 ticks at <= 2 GHz)
        const long long patience = 100000ll + 10000ll * (units / gridDim.x);
        const long long t0 = clock64();
        int ok = 1;
        while (*(volatile int *)(a.tie_count + 1) < units) {
            if (clock64() - t0 > patience) { ok = 0; break; }
            __nanosleep(64);
        }
        s_tie_ok = ok;
    }
    __syncthreads();
    if (!s_tie_ok) {
        for (int item = 0; item < n_items; item++) {
            const int g = a.tie_items[item].x;
            if (g / a.A != b) continue;   // CTA-uniform
            const unsigned long long thr = a.tie_thr[item];
            for (int v0 = 0; v0 < a.N; v0 += T) nvb_tie_unit(a, g, thr, v0 + (int)threadIdx.x);
        }
        __threadfence();
        __syncthreads();
    }
    __threadfence();
}

template <bool NEED_HS, int PH, int PW, bool TIES>
__global__ void __launch_bounds__(NVB_MS_MAX_THREADS, 7)   // 1024 agents = 7 CTAs per SM: one wave
k3_move_sample(const __grid_constant__ CUtensorMap tmap, StepArgs a, SamplerArgs sa)
{
    nvb_tl_stamp(a.tl, 3, 0);
    if (a.pdl_early) nvb_grid_dep_launch();
    extern __shared__ __align__(128) uint8_t smem_k3ms[];
    const SamplerSmem L = nvb_sampler_layout<NEED_HS>(sa.w, sa.A, smem_k3ms);
    // constant data while the tie pass drains: quantisation tables and heading offsets ->
    // shared memory, the training path (every CTA of the SM scans all of it) -> L1
    nvb_sampler_stage_lut(sa.w, L.lut);
    nvb_sampler_stage_ptab(sa.w, L.ptab);
    nvb_sampler_stage_tc(sa.genc != nullptr ? sa.tc_tab : nullptr, L.tc);
    for (int k = threadIdx.x; k < a.A; k += blockDim.x) L.offs[k] = a.offsets[k];
    if (a.pblk != nullptr) {   // update_error reads the block bounds of the whole path, then a few blocks
        const int bytes = min(((a.n_path + NVB_PATH_BLOCK - 1) / NVB_PATH_BLOCK) * 32, 64 * 1024);
        for (int o = threadIdx.x * 128; o < bytes; o += blockDim.x * 128)
            asm volatile("prefetch.global.L1 [%0];" ::"l"(reinterpret_cast<const char *>(a.pblk) + o));
    } else if (a.n_path <= 4096) {
        for (int o = threadIdx.x * 128; o < a.n_path * 16; o += blockDim.x * 128)
            asm volatile("prefetch.global.L1 [%0];" ::"l"(reinterpret_cast<const char *>(a.path) + o));
    }
    if (sa.dbg && threadIdx.x == 0) sa.dbg[blockIdx.x * 8 + 0] = clock64();
    __syncthreads();
    nvb_grid_dep_wait();
    nvb_tl_stamp(a.tl, 3, 1);
    const int b = blockIdx.x;
    if (sa.dbg && threadIdx.x == 0) sa.dbg[b * 8 + 1] = clock64();
    int stepped, n_items = 0;
    MovePre pre;
    {
        // everything this agent's move needs, requested before the first branch on any of it
        stepped = a.ag.stepped[b];
        n_items = TIES ? a.tie_count[0] : 0;
        pre = nvb_move_preload(a, b, nullptr);
        if (TIES && n_items > 0) nvb_tie_help(a, n_items);
        if (!stepped) {
            nvb_log_idle(a, b);
            return;
        }
        if (a.p2p.world > 1) {   // view shards: exact differences, MIN over ranks, then fetched again
            nvb_p2p_min_agent(a.p2p, a.exact, b, a.A, 1);
            pre = nvb_move_preload(a, b, nullptr);
        }
    }
    const unsigned long long *s_ex = nullptr;
    if (TIES && stepped == 2) {
        nvb_tie_wait(a, b, n_items);
        if (a.A <= 32 && threadIdx.x < a.A) pre.eb = __ldcg(a.exact + (size_t)b * a.A + threadIdx.x);
    }
    // The window of the NEXT glimpses depends only on the new pose: its TMA loads start as
    // soon as the pose is known and fly during the path scan of update_error; for sweeps of
    // up to 32 headings warp 1 computes the rotations meanwhile and the other warps scan.
    double pose[3];
    bool oob = false;
    const bool short_sweep = a.A <= 32;
    auto window = [&](const double *p) {
        if (sa.dbg && threadIdx.x == 0) sa.dbg[b * 8 + 2] = clock64();
        oob = nvb_sample_window<NEED_HS>(&tmap, sa, b, p[0], p[1], L, 0, sa.A);
    };
    bool more;
    if (short_sweep) {
        more = nvb_move<0, true>(a, b, s_ex, pre, L.offs, pose, window,
                                 [&](const double *p) { nvb_sample_rotations(sa, b, p[2], L, L.offs, 32, 32, 0, sa.A); });
    } else {
        more = nvb_move<0, false>(a, b, s_ex, pre, L.offs, pose, window);
    }
    if (sa.dbg && threadIdx.x == 0) sa.dbg[b * 8 + 3] = clock64();
    if (!more) {
        // the agent stops here; a window in flight must land before the CTA's shared memory goes
        if (!oob && sa.w.R > 0) nvb_mbar_wait(L.mbar, 0);
        return;
    }
    if (oob) {   // NavBySceneFamiliarity.py:156-158, reported at the step it belongs to
        if (threadIdx.x == 0) a.pending_fail[b] = -2;
        return;
    }
    if (!short_sweep) nvb_sample_rotations(sa, b, pose[2], L, L.offs, 0, (int)blockDim.x, 0, sa.A);
    nvb_sample_gather<NEED_HS, PH, PW>(sa, b, pose[0], pose[1], L, a.pending_fail + b, 0, sa.A);
    nvb_tl_stamp(a.tl, 3, 2);
}

// ---- three launches (large / view-sharded libraries) -----------------------------
__global__ void __launch_bounds__(NVB_STEP_THREADS, 7)   // 1024 agents = 7 CTAs per SM: one wave
k3_decide(StepArgs a)
{
    nvb_tl_stamp(a.tl, 1, 0);
    if (a.pdl_early) nvb_grid_dep_launch();
    __shared__ double s_div[256];
    for (int k = threadIdx.x; k < 256; k += blockDim.x) s_div[k] = a.div255[k];   // constant table
    nvb_grid_dep_wait();
    nvb_tl_stamp(a.tl, 1, 1);
    const int b = blockIdx.x;
    // everything the first decisions need, requested in one round trip; thread 0 alone
    // decides whether the agent takes part (the others must not race its status update)
    __shared__ int s_active;
    if (threadIdx.x < a.A && threadIdx.x < 32)
        asm volatile("prefetch.global.L1 [%0];" ::"l"(a.keys + (size_t)b * a.A + threadIdx.x));
    if (threadIdx.x == 0) {
        const int pf = (a.pending_fail != nullptr) ? a.pending_fail[b] : 0;
        const int status = a.ag.status[b], completed = a.ag.completed[b], budget = a.ag.budget[b];
        // a failure the sampler found for THIS step while it ran at the end of the previous
        // one becomes the agent's status now (as nvb_commit_pending)
        if (pf != 0) {
            a.ag.status[b] = pf;
            a.pending_fail[b] = 0;
        }
        const int active = (pf == 0) && status == 0 && completed < budget;
        a.ag.stepped[b] = active;
        s_active = active;
    }
    __syncthreads();   // s_active, s_div
    const bool active = s_active != 0;
    // view shards: the keys of this agent become the MIN over ranks before anything reads them.
    // Every rank holds the same agent states, so `active` is the same everywhere and the
    // exchange is skipped (by all of them) for an agent that takes no part in the step.
    if (!active) return;
    nvb_p2p_min_agent(a.p2p, a.keys, b, a.A, 0);
    nvb_decide<false>(a, b, nullptr, s_div);
    nvb_tl_stamp(a.tl, 1, 2);
}

// Tie pass: every (tied glimpse, local view) pair whose score is within the
// band gets its exact FP64 difference; min per glimpse.
#define NVB_TIE_THREADS 256
__global__ void __launch_bounds__(NVB_TIE_THREADS)
k3_ties(StepArgs a)
{
    nvb_tl_stamp(a.tl, 2, 0);
    if (a.pdl_early) nvb_grid_dep_launch();
    nvb_grid_dep_wait();
    nvb_tl_stamp(a.tl, 2, 1);
    const int chunks = (a.N + NVB_TIE_THREADS - 1) / NVB_TIE_THREADS;
    // the list entry this CTA would start with is requested together with the list length
    // (one round trip instead of two; the list arrays hold at least B * A entries)
    const int item0 = min((int)(blockIdx.x / chunks), a.B * a.A - 1);
    const int g0 = a.tie_items[item0].x;
    const unsigned long long thr0 = a.tie_thr[item0];
    const int n_items = *a.tie_count;
    nvb_tl_stamp(a.tl, 2, 2);   // (overwritten below when there is work)
    if (n_items == 0) return;
    const long long units = (long long)n_items * chunks;
    for (long long u = blockIdx.x; u < units; u += gridDim.x) {
        const int item = (int)(u / chunks), ch = (int)(u - (long long)item * chunks);
        const int g = (u == blockIdx.x) ? g0 : a.tie_items[item].x;
        const unsigned long long thr = (u == blockIdx.x) ? thr0 : a.tie_thr[item];
        const int v = ch * NVB_TIE_THREADS + threadIdx.x;
        if (v >= a.N) continue;
        const size_t qo = (size_t)g * a.Ppad, fo = (size_t)v * a.Ppad;
        if (nvb_pair_score(a, qo, fo) <= thr) {
            const double d = nvb_exact_diff_rows(a, qo, fo, a.div255);
            atomicMin(a.exact + g, (unsigned long long)__double_as_longlong(d));
        }
    }
    nvb_tl_stamp(a.tl, 2, 2);
}

// Tie pass for LARGE libraries (chem_weight 0, rows of up to 128 bytes): view-major.  k3_ties
// gives every (tied glimpse, 256-view chunk) pair to a CTA, so the library is read once PER tied
// glimpse -- 80 MB each at 10^6 views, 350 us per step-batch with a dozen tied glimpses.  Here a
// thread keeps ONE view row in registers and scores it against every tied glimpse (their rows
// staged in shared memory, NVB_TIEV_ITEMS at a time): the library is read once per launch.
#define NVB_TIEV_ITEMS 32
#define NVB_TIEV_MAX_CHUNKS 8
__global__ void __launch_bounds__(NVB_TIE_THREADS)
k3_ties_v(StepArgs a)
{
    nvb_tl_stamp(a.tl, 2, 0);
    if (a.pdl_early) nvb_grid_dep_launch();
    nvb_grid_dep_wait();
    nvb_tl_stamp(a.tl, 2, 1);
    __shared__ uint4 s_q[NVB_TIEV_ITEMS][NVB_TIEV_MAX_CHUNKS];
    __shared__ unsigned s_thr[NVB_TIEV_ITEMS];
    __shared__ int s_g[NVB_TIEV_ITEMS];
    const int n_items = *a.tie_count;
    nvb_tl_stamp(a.tl, 2, 2);   // (overwritten below when there is work)
    if (n_items == 0) return;
    const int nc = a.Ppad / 16, tid = threadIdx.x;
    for (int i0 = 0; i0 < n_items; i0 += NVB_TIEV_ITEMS) {
        const int ni = min(NVB_TIEV_ITEMS, n_items - i0);
        __syncthreads();   // the previous batch has been read
        for (int q = tid; q < ni * nc; q += blockDim.x) {
            const int i = q / nc, c = q - i * nc;
            s_q[i][c] = reinterpret_cast<const uint4 *>(a.gv + (size_t)a.tie_items[i0 + i].x * a.Ppad)[c];
        }
        if (tid < ni) {
            s_g[tid] = a.tie_items[i0 + tid].x;
            const unsigned long long th = a.tie_thr[i0 + tid];
            s_thr[tid] = th > 0xFFFFFFFFull ? 0xFFFFFFFFu : (unsigned)th;
        }
        __syncthreads();
        for (long long v = (long long)blockIdx.x * blockDim.x + tid; v < a.N; v += (long long)gridDim.x * blockDim.x) {
            const uint4 *f4 = reinterpret_cast<const uint4 *>(a.lv + (size_t)v * a.Ppad);
            uint4 row[NVB_TIEV_MAX_CHUNKS];
#pragma unroll
            for (int c = 0; c < NVB_TIEV_MAX_CHUNKS; c++)
                if (c < nc) row[c] = __ldg(f4 + c);
            for (int i = 0; i < ni; i++) {
                uint32_t sum = 0;
#pragma unroll
                for (int c = 0; c < NVB_TIEV_MAX_CHUNKS; c++)
                    if (c < nc) {
                        const uint4 qq = s_q[i][c];
                        sum = nvb_sad4(qq.x, row[c].x, sum); sum = nvb_sad4(qq.y, row[c].y, sum);
                        sum = nvb_sad4(qq.z, row[c].z, sum); sum = nvb_sad4(qq.w, row[c].w, sum);
                    }
                if (sum <= s_thr[i]) {
                    const double d = nvb_exact_diff_rows(a, (size_t)s_g[i] * a.Ppad, (size_t)v * a.Ppad, a.div255);
                    atomicMin(a.exact + s_g[i], (unsigned long long)__double_as_longlong(d));
                }
            }
        }
    }
    nvb_tl_stamp(a.tl, 2, 2);
}

__global__ void __launch_bounds__(NVB_STEP_THREADS)
k3_move(StepArgs a)
{
    if (a.pdl_early) nvb_grid_dep_launch();
    nvb_grid_dep_wait();
    const int b = blockIdx.x;
    if (!a.ag.stepped[b]) {
        nvb_log_idle(a, b);
        return;
    }
    nvb_p2p_min_agent(a.p2p, a.exact, b, a.A, 1);   // view shards: exact differences, MIN over ranks
    double pose[3];
    const MovePre pre = nvb_move_preload(a, b, nullptr);
    nvb_move<0>(a, b, nullptr, pre, a.offsets, pose);
}


// ---- long training paths: the distance scan of update_error spread over the grid ----
#define NVB_PATH_SPLIT 16384      /* paths longer than this use the three-kernel move */
#define NVB_PATH_CHUNK 8192       /* path points per CTA of k3_path_scan */

__global__ void __launch_bounds__(NVB_STEP_THREADS)
k3_move_pose(StepArgs a)
{
    if (a.pdl_early) nvb_grid_dep_launch();
    nvb_grid_dep_wait();
    const int b = blockIdx.x;
    if (!a.ag.stepped[b]) {
        nvb_log_idle(a, b);
        return;
    }
    nvb_p2p_min_agent(a.p2p, a.exact, b, a.A, 1);   // view shards: exact differences, MIN over ranks
    double pose[3];
    const MovePre pre = nvb_move_preload(a, b, nullptr);
    nvb_move<1>(a, b, nullptr, pre, a.offsets, pose);
}

__global__ void __launch_bounds__(256)
k3_path_scan(StepArgs a)
{
    if (a.pdl_early) nvb_grid_dep_launch();
    nvb_grid_dep_wait();
    const int b = blockIdx.y;
    if (!a.ag.stepped[b] || a.fake) return;
    const double x = a.ag.poses[3 * b], y = a.ag.poses[3 * b + 1];
    const double thr = __dmul_rn(a.coverage_factor, a.step_size);
    const double thr2 = a.cover_thr2;
    const bool one_pass = thr <= a.max_dist;
    const double2 *path = reinterpret_cast<const double2 *>(a.path);
    const int n0 = blockIdx.x * NVB_PATH_CHUNK, n1 = min(n0 + NVB_PATH_CHUNK, a.n_path);
    double m = __longlong_as_double(0x7FF0000000000000ll);
#pragma unroll 4
    for (int n = n0 + threadIdx.x; n < n1; n += 256) {
        const double2 pt = __ldg(path + n);
        const double dx = __dsub_rn(pt.x, x), dy = __dsub_rn(pt.y, y);
        const double d2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
        m = fmin(m, d2);
        if (one_pass && d2 <= thr2) a.ag.coverage[(size_t)b * a.n_path + n] = 1;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmin(m, __shfl_xor_sync(0xFFFFFFFFu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMin(a.dmin2 + b, (unsigned long long)__double_as_longlong(m));
}

__global__ void __launch_bounds__(NVB_STEP_THREADS)
k3_move_finish(StepArgs a)
{
    if (a.pdl_early) nvb_grid_dep_launch();
    nvb_grid_dep_wait();
    const int b = blockIdx.x;
    if (!a.ag.stepped[b]) return;
    double pose[3];
    const MovePre pre = nvb_move_preload(a, b, nullptr);
    nvb_move<2>(a, b, nullptr, pre, a.offsets, pose);
}

