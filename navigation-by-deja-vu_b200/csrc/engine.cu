// engine.cu -- host side of the C ABI declared in include/navsim_b200.h.
//
// One engine = one world (landscape + sensor + heading sweep + library) on one
// B200 and one stream.  All device memory is owned here; Python (ctypes) and
// torch only hand in host pointers, a stream handle and, for the view-sharded
// mode, read the two reduction buffers through nvb_device_ptr().
// There is no CPU fallback: without an sm_100 device nvb_engine_create fails.
#include <math.h>
#include <cmath>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/navsim_b200.h"
#include "common.cuh"
#include "distance.cuh"
#include "distance_tc.cuh"
#include "landscape.cuh"
#include "sampler.cuh"
#include "step.cuh"
#include "step_tm.cuh"

#define KEY_NONE NVB_KEY_NONE

static thread_local std::string g_err;

static int fail(int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CK(call)                                                                           \
    do {                                                                                   \
        cudaError_t _e = (call);                                                           \
        if (_e != cudaSuccess)                                                             \
            return fail(NVB_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), \
                        __FILE__, __LINE__);                                               \
    } while (0)

typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                        const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                        const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct nvb_engine {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int sm_count = 148;
    int64_t launches = 0;
    // landscape
    uint8_t *d_land = nullptr;
    int rows = 0, cols = 0, pitch = 0;
    CUtensorMap tmap;
    // sensor
    bool have_sensor = false;
    int W = 0, H = 0, pw = 0, ph = 0, P = 0, Ppad = 0, cpr = 0, nk = 0;
    int mask_lo = 0, mask_hi = 0;
    int R = 0, BW = 0, BH = 0;
    uint8_t *d_lut = nullptr;
    double *d_div255 = nullptr;
    // saccade
    int A = 0;
    double *d_offsets = nullptr;
    // nav params
    double step_size = 1.0, max_dist = INFINITY, thf = 2.0, cvf = 0.8, cw = 0.0;
    // library
    int N = 0;
    uint8_t *d_lh = nullptr, *d_ls = nullptr, *d_lv = nullptr;
    double *d_path = nullptr;
    double *d_pblk = nullptr;       // [ceil(n_path / 16)][4]: bounding circle (cx, cy, r, -) of 16 consecutive path points
    double *d_pblk2 = nullptr;      // [ceil(blocks / 64)][4]: the same for groups of 64 blocks (long paths)
    int n_path = 0;
    long long view_offset = 0, n_total = 0;
    // glimpse buffers
    long long Gcap = 0;
    uint8_t *d_gh = nullptr, *d_gs = nullptr, *d_gv = nullptr;
    unsigned long long *d_keys = nullptr, *d_exact = nullptr;
    int *d_tie_count = nullptr;
    int2 *d_tie_items = nullptr;
    unsigned long long *d_tie_thr = nullptr;
    int *d_tie_next = nullptr, *d_tie_ready = nullptr;
    // agents
    int B = 0;
    AgentState ag{};
    int *d_step = nullptr;
    int steps_done = 0;
    int *d_spans = nullptr;         // per-CTA unit spans of the distance kernel
    int span_key[4] = {-1, -1, -1, -1};
    // tensor-core distance kernel (distance_tc.cuh): thermometer planes of the V quantisation,
    // encoded glimpses / library and their tensor maps
    bool tc_ok = false;             // the V quantisation has few enough levels (<= NVB_TC_MAX_PLANES planes)
    bool tc_off = false;            // nvb_set_distance_kernel(e, 1): byte-SIMD kernel everywhere
    TcPlanes tc_planes{};
    int tc_K = 0, tc_Kpad = 0, tc_kch = 0, tc_sad_const = 0;
    bool tc_bs = false;             // view-tile-stationary kernel (k2_tc_bs): rows short enough for the tile + 2 glimpse slots
    uint8_t *d_tc_tab = nullptr;    // sampler tables (SamplerArgs::tc_tab)
    uint8_t *d_tc_level_of = nullptr;   // [256] quantised value -> level index, [256] is-a-level flags (k_tc_encode)
    int *d_tc_bad = nullptr;            // set by k_tc_encode when it meets a value that is not a level
    bool tc_lib_ok = false;             // the library holds level values only (checked when it is encoded)
    int8_t *d_genc = nullptr, *d_lenc = nullptr;
    bool lenc_valid = false;
    CUtensorMap tm_genc, tm_lenc;
    int *d_spans_tc = nullptr;
    int span_tc_key[3] = {-1, -1, -1};
    int2 *d_tmin = nullptr;         // [Gcap][n_vt] two best keys per tile of the tensor-core kernel (single-launch step, step_tm.cuh)
    long long tmin_cap = 0;
    bool want_tmin = false;         // the step-batch being queued ends in k3_step_tm (set by one_step)
    float *d_pblk_f = nullptr;      // [ceil(n_path / 16)][4] FP32 copy of d_pblk, radius rounded up
    // landscape preparation (landscape.cuh): grain labels of the current landscape
    long long *d_labels = nullptr;
    int *d_grain_area = nullptr;
    long long n_grains = -1;        // -1: not labelled
    // view-sharded library over NVLink peer memory
    unsigned long long *d_xarea = nullptr;
    P2PArgs p2p{};
    bool p2p_on = false;
    unsigned long long *d_p2p_seq = nullptr;
    int *d_p2p_err = nullptr;
    void *p2p_opened[NVB_P2P_MAX_RANKS] = {nullptr};
    unsigned long long *d_dmin2 = nullptr;   // [B] long-path form of update_error
    long long *d_dbg = nullptr;     // tuning aid: per-agent clock64 checkpoints of one step
    long long *d_tl = nullptr;      // tuning aid: per-CTA wall-clock stamps of the step kernels
    int ms_capacity = -1;           // CTAs of k3_move_sample the device holds at once (form 4 needs B <= this)
    int32_t *d_pending = nullptr;   // [B] sampler failure parked for the next step (fused loop)
    bool glimpses_pending = false;  // the glimpses of the next step are already sampled
    double *d_poses0 = nullptr;     // start poses / budgets kept for nvb_agents_rewind
    int32_t *d_budget0 = nullptr;
    // CUDA graph of one step-batch (phase1+2+3), keyed on (fake, log_afam, log buffers)
    cudaGraphExec_t graph_exec = nullptr;
    cudaGraphExec_t graph_multi = nullptr;   // NVB_GRAPH_UNROLL step-batches in one graph (valid with graph_exec)
    cudaGraphExec_t graph_io = nullptr;   // one step-batch from fresh poses, nothing sampled ahead
    // the same without copy operations, for a caller that keeps passing the same page-locked
    // buffers: the sampler reads the poses from the host buffer and the move writes the results
    // into the host buffers (zero-copy); keyed on the four pointers
    cudaGraphExec_t graph_zc = nullptr;
    const void *zc_key[4] = {nullptr, nullptr, nullptr, nullptr}, *seen_key[4] = {nullptr, nullptr, nullptr, nullptr};
    void *zc_dev[4] = {nullptr, nullptr, nullptr, nullptr};   // device-mapped views the graph was captured with
    const double *zc_in = nullptr;     // device-mapped views of the caller's buffers, set while capturing
    int16_t *zc_best = nullptr;
    double *zc_pose = nullptr, *zc_fam = nullptr;
    const void *graph_io_log_ptr = nullptr;
    int graph_io_log_cap = -1;
    int graph_fake = -1, graph_afam = -1, graph_log_cap = -1, graph_B = -1;
    const void *graph_log_ptr = nullptr;
    bool graph_dirty = true;   // set by every call that changes what the captured kernels were given
    bool use_graph = true;
    // per-kernel timing (bench.py roofline): events around K2 inside the step sequence
    bool timing = false;
    std::vector<cudaEvent_t> ev_pool;
    size_t ev_used = 0;
    double k2_ms = 0.0;
    long long k2_count = 0;
    std::vector<cudaEvent_t> ev3;   // tuning aid: decide / ties / move+sample
    size_t ev3_used = 0;
    // log
    int log_cap = 0, log_A = 0;
    int16_t *log_best = nullptr;
    double *log_pose = nullptr, *log_sfam = nullptr, *log_afam = nullptr;
};

// Step form for small un-sharded libraries, NAVSIM_B200_STEP_FORM (C2, warm, us per step-batch):
//   5 (default)  K2 (tensor-core kernel, leaving the minimum of every view tile) | k3_step_tm: decide
//                from the tile minima + move + sample in ONE launch (step_tm.cuh); two launches per
//                step.  Where it does not apply (tm_form()) form 3 runs.
//   3            K2 | decide | grid-wide tie pass | move + sample
//   4            K2 | decide | move + sample with the tie pass folded into its front
//                (needs the whole move+sample grid co-resident, else form 3 is used; the agents
//                with ties are the critical path either way, so only a launch boundary is saved
//                and the extra code in the big kernel costs more)
//   1            K2 | decide + ties + move + sample in one launch      (tie agents set the tail)
//   2            K2 | decide + cooperative ties | move + sample
static int step_form()
{
    static const int v = getenv("NAVSIM_B200_STEP_FORM") ? atoi(getenv("NAVSIM_B200_STEP_FORM")) : 5;
    return (v >= 1 && v <= 5) ? v : 5;
}

static bool split_step() { return step_form() != 1; }

// Launch with programmatic stream serialization (PDL): the kernel may start while the
// previous kernel of the stream drains; every kernel of the step sequence begins
// with griddepcontrol.wait before it touches global memory.
static bool use_pdl()
{
    static const bool v = getenv("NAVSIM_B200_NO_PDL") == nullptr;
    return v;
}

// Every kernel of the step sequence releases its dependent at its top
// (griddepcontrol.launch_dependents), so the next kernel's CTAs fill SMs as they free up.
static bool early_trigger()
{
    static const bool v = use_pdl() && getenv("NAVSIM_B200_NO_EARLY_TRIGGER") == nullptr;
    return v;
}

template <typename... KArgs, typename... Args>
static cudaError_t launch_seq(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args... args)
{
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = use_pdl() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

// ---------------------------------------------------------------------------
static void free_dev(void *p)
{
    if (p) cudaFree(p);
}

template <typename T>
static int alloc_dev(T **p, size_t n)
{
    free_dev(*p);
    *p = nullptr;
    if (n == 0) n = 1;
    CK(cudaMalloc((void **)p, n * sizeof(T)));
    return NVB_OK;
}

static NvbWorld make_world(const nvb_engine *e)
{
    NvbWorld w;
    w.land = e->d_land;
    w.rows = e->rows;
    w.cols = e->cols;
    w.pitch = e->pitch;
    w.plane_stride = (long long)e->pitch * e->rows;
    w.W = e->W; w.H = e->H; w.pw = e->pw; w.ph = e->ph;
    w.P = e->P; w.Ppad = e->Ppad;
    w.Wpx = e->W * e->pw; w.Hpx = e->H * e->ph;
    w.mask_lo = e->mask_lo; w.mask_hi = e->mask_hi;
    w.r = (double)(w.Wpx > w.Hpx ? w.Wpx : w.Hpx) / 2.0;
    w.R = e->R; w.BW = e->BW; w.BH = e->BH;
    w.lut = e->d_lut;
    return w;
}

static int rebuild_tmap(nvb_engine *e)
{
    memset(&e->tmap, 0, sizeof e->tmap);
    e->R = 0;
    if (!e->d_land || !e->have_sensor) return NVB_OK;
    if (getenv("NAVSIM_B200_NO_TMA")) return NVB_OK;   // debugging knob: gather from global memory
    const double hw = 0.5 * e->W * e->pw, hh = 0.5 * e->H * e->ph;
    const int R = (int)ceil(sqrt(hw * hw + hh * hh)) + 1;
    const int BH = 2 * R + 2, BW = nvb_round_up(2 * R + 2 + 15, 16);   // +15: 16-B aligned box origin
    if (BW > 256 || BH > 256) return NVB_OK;   // window does not fit a TMA box: gather from global
    if (nvb_sampler_smem(BW, BH, 3, e->A > 0 ? e->A : 1) > 200 * 1024) return NVB_OK;
    static PFN_tmapEncodeTiled encode = nullptr;
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (!fn || qres != cudaDriverEntryPointSuccess)
            return fail(NVB_E_CUDA, "cuTensorMapEncodeTiled not available from the driver");
        encode = (PFN_tmapEncodeTiled)fn;
    }
    cuuint64_t gdim[3] = {(cuuint64_t)e->cols, (cuuint64_t)e->rows, 3};
    cuuint64_t gstride[2] = {(cuuint64_t)e->pitch, (cuuint64_t)e->pitch * (cuuint64_t)e->rows};
    cuuint32_t box[3] = {(cuuint32_t)BW, (cuuint32_t)BH, 1};
    cuuint32_t estride[3] = {1, 1, 1};
    CUresult r = encode(&e->tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, e->d_land, gdim, gstride, box,
                        estride, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(NVB_E_CUDA, "cuTensorMapEncodeTiled failed: %d", (int)r);
    e->R = R; e->BW = BW; e->BH = BH;
    return NVB_OK;
}

static PFN_tmapEncodeTiled tmap_encoder()
{
    static PFN_tmapEncodeTiled encode = nullptr;
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
            fn && qres == cudaDriverEntryPointSuccess)
            encode = (PFN_tmapEncodeTiled)fn;
    }
    return encode;
}

// ---- tensor-core distance kernel: operand planes, buffers, tensor maps ---------------------
// NAVSIM_B200_NO_TC=1 keeps every configuration on the byte-SIMD kernel (k2_sad_v).
#define NVB_TC_MIN_G 96   /* fewer glimpses than this leave most of a 128-row MMA tile empty */

static bool use_tc(const nvb_engine *e, long long G)
{
    static const bool off = getenv("NAVSIM_B200_NO_TC") != nullptr;
    return e->tc_ok && !e->tc_off && e->cw == 0.0 && G >= NVB_TC_MIN_G && !off;
}

// The step after the tensor-core distance kernel as ONE launch (step_tm.cuh) instead of
// decide | grid-wide tie pass | move + sample.  Step form 5 (the default); where it does not
// apply form 3 runs.
static bool fused_step(const nvb_engine *e);
static bool long_path_split(const nvb_engine *e);
static bool tm_form(const nvb_engine *e)
{
    const int n_blk = (e->n_path + NVB_PATH_BLOCK - 1) / NVB_PATH_BLOCK;
    return step_form() == 5 && fused_step(e) && use_tc(e, (long long)e->B * e->A) && e->lenc_valid && e->tc_lib_ok &&
           !e->p2p_on && e->view_offset == 0 && e->n_total == e->N && e->R > 0 && e->P <= NVB_PTAB_MAX &&
           e->A <= NVB_TM_MAX_A && (e->N + NVB_TC_NT - 1) / NVB_TC_NT <= NVB_TM_MAX_VT &&
           e->d_pblk_f != nullptr && n_blk <= 4096 && !long_path_split(e) && e->cvf * e->step_size <= e->max_dist &&
           getenv("NAVSIM_B200_NO_PATH_BLOCKS") == nullptr;
}

// Thermometer planes of the V-channel quantisation table (distance_tc.cuh): levels = the
// distinct table values plus 0 (masked pixels are 0 whatever the table says), one plane per
// gap between consecutive levels, gaps above 127 split so that every weight fits an int8.
static int build_tc_planes(nvb_engine *e, const uint8_t *lut_v)
{
    e->tc_ok = false;
    e->lenc_valid = false;
    bool seen[256] = {false};
    seen[0] = true;
    for (int x = 0; x < 256; x++) seen[lut_v[x]] = true;
    std::vector<int> levels;
    for (int v = 0; v < 256; v++)
        if (seen[v]) levels.push_back(v);
    if ((int)levels.size() > NVB_TC_TAB_LEVELS) return NVB_OK;
    TcPlanes pl{};
    for (size_t k = 0; k + 1 < levels.size(); k++) {
        int w = levels[k + 1] - levels[k];
        const int parts = (w + 126) / 127;
        for (int q = 0; q < parts; q++) {
            if (pl.n_planes == NVB_TC_MAX_PLANES) return NVB_OK;   // too many planes: byte-SIMD kernel
            const int wq = w / (parts - q);
            pl.weight[pl.n_planes] = (int8_t)wq;
            pl.thr_level[pl.n_planes] = (uint8_t)k;
            pl.n_planes++;
            w -= wq;
        }
    }
    if (pl.n_planes == 0) return NVB_OK;
    uint8_t level_of[512] = {0}, tab[NVB_TC_TAB_BYTES] = {0};
    for (size_t i = 0; i < levels.size(); i++) { level_of[levels[i]] = (uint8_t)i; level_of[256 + levels[i]] = 1; }
    for (int v = 0; v < 256; v++) tab[v] = level_of[lut_v[v]];
    int sum_w = 0;
    for (int k = 0; k < pl.n_planes; k++) sum_w += pl.weight[k];
    for (size_t l = 0; l < levels.size(); l++)
        for (int k = 0; k < pl.n_planes; k++)
            tab[256 + l * 8 + k] = (uint8_t)(int8_t)(((int)l > (int)pl.thr_level[k]) ? pl.weight[k] : -pl.weight[k]);
    int rc;
    if ((rc = alloc_dev(&e->d_tc_tab, (size_t)NVB_TC_TAB_BYTES))) return rc;
    if ((rc = alloc_dev(&e->d_tc_level_of, (size_t)512))) return rc;
    if ((rc = alloc_dev(&e->d_tc_bad, (size_t)1))) return rc;
    CK(cudaMemset(e->d_tc_bad, 0, sizeof(int)));
    CK(cudaMemcpy(e->d_tc_tab, tab, sizeof tab, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(e->d_tc_level_of, level_of, sizeof level_of, cudaMemcpyHostToDevice));
    e->tc_planes = pl;
    e->tc_K = pl.n_planes * e->P;
    const int k64 = nvb_round_up(e->tc_K, 64), k128 = nvb_round_up(e->tc_K, 128);
    // short rows: the view tile stays resident in shared memory (k2_tc_bs, 64-byte chunks);
    // otherwise 128-byte K chunks (fewer, larger TMA transactions) unless the zero padding gets heavy
    e->tc_bs = nvb_tcbs_slots(k64 / NVB_TCBS_KCH) >= 2 && getenv("NAVSIM_B200_TC_STREAM") == nullptr;
    e->tc_kch = e->tc_bs ? 64 : (k128 * 4 <= e->tc_K * 5) ? 128 : 64;
    e->tc_Kpad = e->tc_kch == 128 ? k128 : k64;
    e->tc_sad_const = e->P * sum_w;
    // 32-bit keys of the epilogue: 256 * |dot| + column must stay below 2^31
    e->tc_ok = (long long)e->tc_sad_const * 256 + 256 < (1ll << 31) && tmap_encoder() != nullptr;
    return NVB_OK;
}

static int make_tc_map(CUtensorMap *m, void *base, long long rows, int Kpad, int kch, int box_rows)
{
    cuuint64_t gdim[2] = {(cuuint64_t)Kpad, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)Kpad};
    cuuint32_t box[2] = {(cuuint32_t)kch, (cuuint32_t)box_rows};
    cuuint32_t estride[2] = {1, 1};
    CUresult r = tmap_encoder()(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, base, gdim, gstride, box, estride,
                                CU_TENSOR_MAP_INTERLEAVE_NONE,
                                kch == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(NVB_E_CUDA, "cuTensorMapEncodeTiled (operand planes) failed: %d", (int)r);
    return NVB_OK;
}


// Encoded glimpse buffer for Gcap glimpses (zeroed: the K padding must be 0) + its tensor map.
static int ensure_tc_glimpses(nvb_engine *e, long long Gcap)
{
    if (!e->tc_ok) return NVB_OK;
    int rc;
    const size_t bytes = (size_t)Gcap * e->tc_Kpad;
    if ((rc = alloc_dev(&e->d_genc, bytes))) return rc;
    CK(cudaMemsetAsync(e->d_genc, 0, bytes, e->stream));
    return make_tc_map(&e->tm_genc, e->d_genc, Gcap, e->tc_Kpad, e->tc_kch, NVB_TC_TM);
}

// Encoded library (built on first use: a 10^6-view library scored by a handful of glimpses
// never needs it).
static int ensure_tc_library(nvb_engine *e)
{
    if (e->lenc_valid) return NVB_OK;
    int rc;
    const size_t bytes = (size_t)e->N * e->tc_Kpad;
    if ((rc = alloc_dev(&e->d_lenc, bytes))) return rc;
    CK(cudaMemsetAsync(e->d_lenc, 0, bytes, e->stream));
    CK(cudaMemsetAsync(e->d_tc_bad, 0, sizeof(int), e->stream));
    const long long n = (long long)e->N * e->P;
    k_tc_encode<false><<<(unsigned)((n + 255) / 256), 256, 0, e->stream>>>(e->d_lv, e->N, e->P, e->Ppad, e->tc_Kpad,
                                                                            e->tc_planes, e->d_tc_level_of, e->d_lenc,
                                                                            e->d_tc_bad);
    e->launches++;
    CK(cudaGetLastError());
    int bad = 0;
    CK(cudaMemcpyAsync(&bad, e->d_tc_bad, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    e->tc_lib_ok = (bad == 0);   // a library uploaded from the host with other values: byte-SIMD kernel
    if (!e->tc_lib_ok) { free_dev(e->d_lenc); e->d_lenc = nullptr; }
    else if ((rc = make_tc_map(&e->tm_lenc, e->d_lenc, e->N, e->tc_Kpad, e->tc_kch, NVB_TC_NT))) return rc;
    e->lenc_valid = true;
    e->graph_dirty = true;
    return NVB_OK;
}

static int ensure_glimpse_cap(nvb_engine *e, long long G)
{
    if (G <= e->Gcap) return NVB_OK;
    e->graph_dirty = true;
    e->glimpses_pending = false;   // the buffers that held the pre-sampled glimpses are replaced
    int rc;
    const size_t bytes = (size_t)G * e->Ppad;
    if ((rc = alloc_dev(&e->d_gh, bytes))) return rc;
    if ((rc = alloc_dev(&e->d_gs, bytes))) return rc;
    if ((rc = alloc_dev(&e->d_gv, bytes))) return rc;
    CK(cudaMemsetAsync(e->d_gh, 0, bytes, e->stream));
    CK(cudaMemsetAsync(e->d_gs, 0, bytes, e->stream));
    CK(cudaMemsetAsync(e->d_gv, 0, bytes, e->stream));
    if ((rc = alloc_dev(&e->d_keys, (size_t)G))) return rc;
    if ((rc = alloc_dev(&e->d_exact, (size_t)G))) return rc;
    if ((rc = alloc_dev(&e->d_tie_items, (size_t)G))) return rc;
    if ((rc = alloc_dev(&e->d_tie_thr, (size_t)G))) return rc;
    if ((rc = alloc_dev(&e->d_tie_next, (size_t)G))) return rc;
    if ((rc = alloc_dev(&e->d_tie_ready, (size_t)G))) return rc;
    CK(cudaMemsetAsync(e->d_tie_ready, 0, sizeof(int) * (size_t)G, e->stream));
    if ((rc = ensure_tc_glimpses(e, G))) return rc;
    e->Gcap = G;
    return NVB_OK;
}

// ---------------------------------------------------------------------------
extern "C" const char *nvb_last_error(void) { return g_err.c_str(); }
extern "C" const char *nvb_version(void) { return "navsim_b200 0.1 (sm_100a)"; }

extern "C" int nvb_engine_create(int device, void *stream, nvb_engine **out)
{
    if (!out) return fail(NVB_E_INVALID, "out is NULL");
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        cudaGetLastError();
        return fail(NVB_E_NO_DEVICE, "no CUDA device visible; the engine has no CPU fallback");
    }
    if (device < 0 || device >= count) return fail(NVB_E_INVALID, "device %d out of range", device);
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(NVB_E_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only",
                    device, prop.major, prop.minor);
    CK(cudaSetDevice(device));
    nvb_engine *e = new nvb_engine();
    e->device = device;
    e->sm_count = prop.multiProcessorCount;
    memset(&e->tmap, 0, sizeof e->tmap);
    if (stream) {
        e->stream = (cudaStream_t)stream;
    } else {
        if (cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) != cudaSuccess) {
            delete e;
            return fail(NVB_E_CUDA, "cudaStreamCreate failed");
        }
        e->own_stream = true;
    }
    double div[256];
    for (int k = 0; k < 256; k++) div[k] = (double)k / 255.;
    int rc = alloc_dev(&e->d_div255, 256);
    if (rc) { delete e; return rc; }
    cudaMemcpy(e->d_div255, div, sizeof div, cudaMemcpyHostToDevice);
    alloc_dev(&e->d_step, 1);
    alloc_dev(&e->d_tie_count, 2);   // [0] list length, [1] tie units done (step form 4)
    alloc_dev(&e->d_lut, 768);
    *out = e;
    return NVB_OK;
}

extern "C" void nvb_engine_destroy(nvb_engine *e)
{
    if (!e) return;
    cudaSetDevice(e->device);
    cudaStreamSynchronize(e->stream);
    void *ptrs[] = {e->d_land, e->d_lut, e->d_div255, e->d_offsets, e->d_lh, e->d_ls, e->d_lv,
                    e->d_path, e->d_gh, e->d_gs, e->d_gv, e->d_keys, e->d_exact, e->d_tie_count,
                    e->d_tie_items, e->d_tie_thr, e->d_tie_next, e->d_tie_ready, e->ag.poses, e->ag.status, e->ag.completed,
                    e->ag.budget, e->ag.nav_frames, e->ag.err_sum, e->ag.err_n, e->ag.coverage,
                    e->ag.stepped, e->d_step, e->log_best, e->log_pose, e->log_sfam, e->log_afam,
                    e->d_poses0, e->d_budget0, e->d_spans, e->d_pending, e->d_dmin2, e->d_pblk, e->d_pblk2,
                    e->d_tc_tab, e->d_tc_level_of, e->d_genc, e->d_lenc, e->d_spans_tc, e->d_tc_bad, e->d_tmin, e->d_pblk_f, e->d_labels, e->d_grain_area};
    for (int i = 0; i < NVB_P2P_MAX_RANKS; i++)
        if (e->p2p_opened[i]) cudaIpcCloseMemHandle(e->p2p_opened[i]);
    free_dev(e->d_xarea); free_dev(e->d_p2p_seq); free_dev(e->d_p2p_err);
    if (e->graph_exec) cudaGraphExecDestroy(e->graph_exec);
    if (e->graph_multi) cudaGraphExecDestroy(e->graph_multi);
    if (e->graph_io) cudaGraphExecDestroy(e->graph_io);
    if (e->graph_zc) cudaGraphExecDestroy(e->graph_zc);
    for (cudaEvent_t ev : e->ev_pool) cudaEventDestroy(ev);
    for (void *p : ptrs) free_dev(p);
    if (e->own_stream) cudaStreamDestroy(e->stream);
    delete e;
}

extern "C" int nvb_sync(nvb_engine *e)
{
    CK(cudaSetDevice(e->device));
    CK(cudaStreamSynchronize(e->stream));
    return NVB_OK;
}

extern "C" int64_t nvb_launch_count(nvb_engine *e) { return e->launches; }

extern "C" void *nvb_stream_handle(nvb_engine *e) { return (void *)e->stream; }


extern "C" int nvb_set_landscape(nvb_engine *e, const uint8_t *hsv, int rows, int cols,
                                 ptrdiff_t s_row, ptrdiff_t s_col, ptrdiff_t s_chan)
{
    e->graph_dirty = true;
    if (!hsv || rows <= 0 || cols <= 0) return fail(NVB_E_INVALID, "bad landscape");
    CK(cudaSetDevice(e->device));
    CK(cudaStreamSynchronize(e->stream));
    const int pitch = nvb_round_up(cols, 16);
    std::vector<uint8_t> planar((size_t)3 * rows * pitch, 0);
    for (int ch = 0; ch < 3; ch++)
        for (int y = 0; y < rows; y++) {
            uint8_t *dst = planar.data() + ((size_t)ch * rows + y) * pitch;
            const uint8_t *src = hsv + (ptrdiff_t)y * s_row + (ptrdiff_t)ch * s_chan;
            for (int x = 0; x < cols; x++) dst[x] = src[(ptrdiff_t)x * s_col];
        }
    int rc = alloc_dev(&e->d_land, planar.size());
    if (rc) return rc;
    CK(cudaMemcpy(e->d_land, planar.data(), planar.size(), cudaMemcpyHostToDevice));
    e->rows = rows; e->cols = cols; e->pitch = pitch;
    e->n_grains = -1;
    return rebuild_tmap(e);
}

// ---- landscape preparation on the device (landscape.cuh) ------------------
extern "C" int nvb_landscape_label_grains(nvb_engine *e, int threshold, int modal_w, int64_t *n_grains)
{
    if (!e->d_land) return fail(NVB_E_INVALID, "no landscape");
    if (modal_w < 0 || (modal_w > 0 && modal_w % 2 == 0)) return fail(NVB_E_INVALID, "modal footprint must be odd (or 0 for none)");
    CK(cudaSetDevice(e->device));
    const int rows = e->rows, cols = e->cols;
    const long long n = (long long)rows * cols;
    if (n >= (1ll << 31)) return fail(NVB_E_INVALID, "landscape too large to label");
    uint8_t *d_m0 = nullptr, *d_m1 = nullptr;
    int *d_parent = nullptr, *d_newid = nullptr, *d_rows = nullptr;
    CK(cudaMalloc(&d_m0, (size_t)n));
    CK(cudaMalloc(&d_m1, (size_t)n));
    CK(cudaMalloc(&d_parent, sizeof(int) * (size_t)n));
    CK(cudaMalloc(&d_newid, sizeof(int) * (size_t)n));
    CK(cudaMalloc(&d_rows, sizeof(int) * (size_t)rows));
    const dim3 g2((cols + 127) / 128, rows), b2(128);
    const unsigned g1 = (unsigned)((n + 255) / 256);
    k_ls_threshold<<<g2, b2, 0, e->stream>>>(e->d_land + 2 * (size_t)e->pitch * rows, rows, cols, e->pitch, threshold, d_m0);
    const uint8_t *mask = d_m0;
    if (modal_w > 1) {
        k_ls_modal<<<g2, b2, 0, e->stream>>>(d_m0, rows, cols, modal_w, d_m1);
        mask = d_m1;
    }
    k_ls_init<<<g1, 256, 0, e->stream>>>(mask, n, d_parent);
    k_ls_link<<<g2, b2, 0, e->stream>>>(mask, rows, cols, d_parent);
    k_ls_flatten<<<g1, 256, 0, e->stream>>>(n, d_parent);
    k_ls_row_roots<<<rows, 128, 0, e->stream>>>(d_parent, rows, cols, d_rows);
    e->launches += 6;
    std::vector<int> rc(rows), off(rows);
    CK(cudaMemcpyAsync(rc.data(), d_rows, sizeof(int) * rows, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    long long total = 0;
    for (int y = 0; y < rows; y++) { off[y] = (int)total; total += rc[y]; }
    CK(cudaMemcpyAsync(d_rows, off.data(), sizeof(int) * rows, cudaMemcpyHostToDevice, e->stream));
    k_ls_number<<<rows, 32, 0, e->stream>>>(d_parent, rows, cols, d_rows, d_newid);
    int rcode;
    if ((rcode = alloc_dev(&e->d_labels, (size_t)n))) return rcode;
    if ((rcode = alloc_dev(&e->d_grain_area, (size_t)(total > 0 ? total : 1)))) return rcode;
    CK(cudaMemsetAsync(e->d_grain_area, 0, sizeof(int) * (size_t)(total > 0 ? total : 1), e->stream));
    k_ls_relabel<<<g1, 256, 0, e->stream>>>(d_parent, d_newid, n, e->d_labels, e->d_grain_area);
    e->launches += 2;
    CK(cudaStreamSynchronize(e->stream));
    CK(cudaGetLastError());
    cudaFree(d_m0); cudaFree(d_m1); cudaFree(d_parent); cudaFree(d_newid); cudaFree(d_rows);
    e->n_grains = total;
    if (n_grains) *n_grains = total;
    return NVB_OK;
}

extern "C" int nvb_landscape_grains_get(nvb_engine *e, int32_t *areas, int64_t *labels)
{
    if (e->n_grains < 0) return fail(NVB_E_INVALID, "grains not labelled");
    CK(cudaSetDevice(e->device));
    if (areas && e->n_grains > 0)
        CK(cudaMemcpyAsync(areas, e->d_grain_area, sizeof(int32_t) * (size_t)e->n_grains, cudaMemcpyDeviceToHost, e->stream));
    if (labels)
        CK(cudaMemcpyAsync(labels, e->d_labels, sizeof(int64_t) * (size_t)e->rows * e->cols, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return NVB_OK;
}

extern "C" int nvb_landscape_paint(nvb_engine *e, const uint8_t *H, const uint8_t *S, int64_t n_grains)
{
    if (e->n_grains < 0) return fail(NVB_E_INVALID, "grains not labelled");
    if (n_grains != e->n_grains) return fail(NVB_E_INVALID, "%lld table entries for %lld grains", (long long)n_grains, e->n_grains);
    if (n_grains == 0) return NVB_OK;
    CK(cudaSetDevice(e->device));
    e->graph_dirty = true;
    uint8_t *d_t = nullptr;
    CK(cudaMalloc(&d_t, (size_t)2 * n_grains));
    CK(cudaMemcpyAsync(d_t, H, (size_t)n_grains, cudaMemcpyHostToDevice, e->stream));
    CK(cudaMemcpyAsync(d_t + n_grains, S, (size_t)n_grains, cudaMemcpyHostToDevice, e->stream));
    k_ls_paint<<<dim3((e->cols + 127) / 128, e->rows), 128, 0, e->stream>>>(e->d_labels, e->rows, e->cols, e->pitch,
                                                                             (long long)e->pitch * e->rows, d_t, d_t + n_grains, e->d_land);
    e->launches++;
    CK(cudaStreamSynchronize(e->stream));
    cudaFree(d_t);
    e->N = 0; e->B = 0;     // the library was sampled from the old pixels
    return NVB_OK;
}

extern "C" int nvb_landscape_flip(nvb_engine *e, int flip_v, int flip_h)
{
    if (!e->d_land) return fail(NVB_E_INVALID, "no landscape");
    if (!flip_v && !flip_h) return NVB_OK;
    CK(cudaSetDevice(e->device));
    CK(cudaStreamSynchronize(e->stream));
    e->graph_dirty = true;
    uint8_t *d_new = nullptr;
    const size_t bytes = (size_t)3 * e->pitch * e->rows;
    CK(cudaMalloc(&d_new, bytes));
    CK(cudaMemsetAsync(d_new, 0, bytes, e->stream));
    k_ls_flip_planes<<<dim3((e->cols + 127) / 128, e->rows, 3), 128, 0, e->stream>>>(e->d_land, e->rows, e->cols, e->pitch,
                                                                                     (long long)e->pitch * e->rows, flip_v, flip_h, d_new);
    e->launches++;
    CK(cudaStreamSynchronize(e->stream));
    cudaFree(e->d_land);
    e->d_land = d_new;
    e->n_grains = -1;       // labels refer to the unflipped pixels
    e->N = 0; e->B = 0;
    return rebuild_tmap(e);
}

extern "C" int nvb_diffuse(nvb_engine *e, const double *initial, int64_t side, int64_t nstep, double multiplier, double *out)
{
    if (!initial || !out || side <= 0 || side > 46340 || nstep < 0) return fail(NVB_E_INVALID, "diffuse: bad arguments");
    CK(cudaSetDevice(e->device));
    const size_t bytes = sizeof(double) * (size_t)side * side;
    double *d_a = nullptr, *d_b = nullptr;
    CK(cudaMalloc(&d_a, bytes));
    if (cudaMalloc(&d_b, bytes) != cudaSuccess) { cudaFree(d_a); return fail(NVB_E_CUDA, "diffuse: out of device memory"); }
    cudaError_t ce = cudaMemcpyAsync(d_a, initial, bytes, cudaMemcpyHostToDevice, e->stream);
    const dim3 grid((unsigned)((side + 255) / 256), (unsigned)side);
    for (int64_t t = 0; t < nstep && ce == cudaSuccess; t++) {
        k_diffuse_step<<<grid, 256, 0, e->stream>>>(d_a, d_b, (int)side, multiplier);
        e->launches++;
        std::swap(d_a, d_b);
        if ((t & 63) == 63) ce = cudaGetLastError();
    }
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(out, d_a, bytes, cudaMemcpyDeviceToHost, e->stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
    cudaFree(d_a);
    cudaFree(d_b);
    if (ce != cudaSuccess) return fail(NVB_E_CUDA, "diffuse: %s", cudaGetErrorString(ce));
    return NVB_OK;
}

extern "C" int nvb_landscape_download(nvb_engine *e, uint8_t *hsv)
{
    if (!e->d_land) return fail(NVB_E_INVALID, "no landscape");
    CK(cudaSetDevice(e->device));
    std::vector<uint8_t> planar((size_t)3 * e->pitch * e->rows);
    CK(cudaMemcpyAsync(planar.data(), e->d_land, planar.size(), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    for (int ch = 0; ch < 3; ch++)
        for (int y = 0; y < e->rows; y++) {
            const uint8_t *src = planar.data() + ((size_t)ch * e->rows + y) * e->pitch;
            for (int x = 0; x < e->cols; x++) hsv[((size_t)y * e->cols + x) * 3 + ch] = src[x];
        }
    return NVB_OK;
}

extern "C" int nvb_set_sensor(nvb_engine *e, int W, int H, int pw, int ph, const uint8_t *lut,
                              int mask_middle_n)
{
    e->graph_dirty = true;
    if (W <= 0 || H <= 0 || pw <= 0 || ph <= 0 || !lut) return fail(NVB_E_INVALID, "bad sensor");
    if ((W * pw) % 2 || (H * ph) % 2)
        return fail(NVB_E_INVALID, "sensor footprint must be even (NavBySceneFamiliarity.py:93)");
    if (pw * ph > NVB_MAX_BLOCK_PX)
        return fail(NVB_E_INVALID, "sensor pixel of %dx%d landscape pixels exceeds %d", pw, ph,
                    NVB_MAX_BLOCK_PX);
    CK(cudaSetDevice(e->device));
    CK(cudaStreamSynchronize(e->stream));
    e->W = W; e->H = H; e->pw = pw; e->ph = ph;
    e->P = W * H;
    int chunks = nvb_round_up(e->P, 16) / 16;
    if (chunks >= 8) {
        e->cpr = 8;
        e->nk = (chunks + 7) / 8;
    } else {
        if (chunks % 2 == 0 && chunks != 2 && chunks != 4) chunks += 1;   // 6 -> 7
        e->cpr = chunks;
        e->nk = 1;
    }
    e->Ppad = chunks * 16;
    // out[:, r1-m : r1+m] = 0 with Python slice semantics (NavBySceneFamiliarity.py:189-190)
    long lo = W / 2 - mask_middle_n, hi = W / 2 + mask_middle_n;
    if (lo < 0) { lo += W; if (lo < 0) lo = 0; }
    if (hi < 0) { hi += W; if (hi < 0) hi = 0; }
    if (hi > W) hi = W;
    if (lo > W) lo = W;
    e->mask_lo = (int)lo; e->mask_hi = (int)hi;
    CK(cudaMemcpy(e->d_lut, lut, 768, cudaMemcpyHostToDevice));
    e->have_sensor = true;
    // sensor change invalidates library and glimpse buffers
    e->N = 0; e->n_path = 0; e->Gcap = 0; e->B = 0;
    int rc = build_tc_planes(e, lut + 512);
    if (rc) return rc;
    return rebuild_tmap(e);
}

extern "C" int nvb_set_saccade(nvb_engine *e, int A, const double *offs)
{
    e->graph_dirty = true;
    if (A <= 0 || A > 32767 || !offs) return fail(NVB_E_INVALID, "bad saccade");
    CK(cudaSetDevice(e->device));
    CK(cudaStreamSynchronize(e->stream));
    int rc = alloc_dev(&e->d_offsets, (size_t)A);
    if (rc) return rc;
    CK(cudaMemcpy(e->d_offsets, offs, sizeof(double) * A, cudaMemcpyHostToDevice));
    e->A = A;
    e->B = 0;
    e->ms_capacity = -1;
    return rebuild_tmap(e);
}

extern "C" int nvb_set_nav_params(nvb_engine *e, double step_size, double max_dist, double thf,
                                  double cvf, double cw)
{
    if (!(cw >= 0.0 && cw <= 1.0)) return fail(NVB_E_INVALID, "chem_weight must be in [0, 1]");
    if (e->step_size != step_size || e->max_dist != max_dist || e->thf != thf || e->cvf != cvf || e->cw != cw)
        e->graph_dirty = true;
    e->step_size = step_size; e->max_dist = max_dist; e->thf = thf; e->cvf = cvf; e->cw = cw;
    return NVB_OK;
}

// ---------------------------------------------------------------------------
// FP32 fast-path guard band of K1 (see sampler.cuh): comfortably above the worst
// FP32 error of px*c - py*s + frac(x) for |px|, |py| <= half the sensor footprint.
static float sampler_band(const nvb_engine *e)
{
    // Worst FP32 error of a sample coordinate t = px*c - py*s + frac(x) as sampler.cuh
    // computes it, |t| < M:  c and s rounded to FP32 (2^-25 each, times |px| + |py|), frac(x)
    // rounded (2^-25), two FMAs and at most (ph - 1) + (pw - 1) running adds of half an ulp
    // of M each.  The band is 1.5 x that bound.
    const double hx = 0.5 * e->W * e->pw, hy = 0.5 * e->H * e->ph;
    const double M = hx + hy + 2.0;
    const double half_ulp = ldexp(1.0, (int)floor(log2(M)) - 24);
    const double err = (hx + hy + 1.0) * ldexp(1.0, -25) + (double)(e->ph + e->pw) * half_ulp;
    return (float)(1.5 * err + 1e-6);
}

template <bool HS, int PH, int PW>
static int launch_sampler_t(nvb_engine *e, const SamplerArgs &sa, int nblocks, size_t smem, int slices)
{
    static size_t attr_set[64] = {0};
    size_t &cur = attr_set[e->device & 63];
    if (smem > cur) {
        CK(cudaFuncSetAttribute(k1_sample<HS, PH, PW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cur = smem;
    }
    k1_sample<HS, PH, PW><<<dim3(nblocks, slices), NVB_SAMPLER_THREADS, smem, e->stream>>>(e->tmap, sa);
    e->launches++;
    CK(cudaGetLastError());
    return NVB_OK;
}

// Wide sweeps of few agents (BASELINE configs[2]: 360 headings x 4096 sensor pixels, 1..64
// agents): one CTA per agent would leave most SMs idle while a handful gather millions of
// samples, so the headings of an agent are sliced over several CTAs -- which needs the
// sampler as its own launch (the un-fused step sequence K1, K2, decide, ties, move).
static int sampler_slices(const nvb_engine *e)
{
    static const bool off = getenv("NAVSIM_B200_NO_SLICES") != nullptr;
    if (off) return 1;
    if ((long long)e->A * e->P < 32768 || e->B >= 2 * e->sm_count) return 1;
    int s = (2 * e->sm_count + e->B - 1) / e->B;
    return s < e->A ? s : e->A;
}

static int launch_sampler(nvb_engine *e, const SamplerArgs &sa, int nblocks, int slices = 1)
{
    const int nplanes = sa.need_hs ? 3 : 1;
    const size_t smem = e->R > 0 ? nvb_sampler_smem(e->BW, e->BH, nplanes, sa.A)
                                 : nvb_sampler_smem(0, 0, 0, sa.A);
#ifndef NVB_DEV_MIN
    if (sa.need_hs) return launch_sampler_t<true, 0, 0>(e, sa, nblocks, smem, slices);
#endif
    // sensor-pixel footprints the reference's drivers use get an unrolled sampling loop
    if (e->ph == 4 && e->pw == 2) return launch_sampler_t<false, 4, 2>(e, sa, nblocks, smem, slices);
#ifndef NVB_DEV_MIN
    if (e->ph == 2 && e->pw == 2) return launch_sampler_t<false, 2, 2>(e, sa, nblocks, smem, slices);
    if (e->ph == 1 && e->pw == 1) return launch_sampler_t<false, 1, 1>(e, sa, nblocks, smem, slices);
#endif
    return launch_sampler_t<false, 0, 0>(e, sa, nblocks, smem, slices);
}

// Cuts the glimpse-tile-major unit list into one contiguous span per CTA.  Every
// unit costs the same, so spans hold floor(units / n_cta) units and the first
// (units % n_cta) CTAs one more.  CTA b and CTA b + sm_count share an SM
// (classic launch order), so putting the longer spans first spreads them over
// distinct SMs instead of stacking two of them on one.
static int build_spans(nvb_engine *e, int n_gt, int n_vt, int n_cta)
{
    const long long units = (long long)n_gt * n_vt;
    if (e->span_key[0] == n_gt && e->span_key[1] == n_vt && e->span_key[2] == n_cta) return NVB_OK;
    std::vector<int> spans(n_cta + 1);
    const long long base = units / n_cta, rem = units % n_cta;
    long long u = 0;
    for (int c = 0; c < n_cta; c++) {
        spans[c] = (int)u;
        u += base + (c < rem ? 1 : 0);
    }
    spans[n_cta] = (int)units;
    int rc = alloc_dev(&e->d_spans, (size_t)n_cta + 1);
    if (rc) return rc;
    CK(cudaMemcpyAsync(e->d_spans, spans.data(), sizeof(int) * (n_cta + 1), cudaMemcpyHostToDevice, e->stream));
    CK(cudaStreamSynchronize(e->stream));   // spans is a stack vector
    e->span_key[0] = n_gt; e->span_key[1] = n_vt; e->span_key[2] = n_cta;
    e->graph_dirty = true;
    return NVB_OK;
}

template <int TY, int MG, int MV, int CPR, bool BULK>
static int launch_dist_cfg2(nvb_engine *e, DistArgs da)
{
    // library streaming (one thread column per view, TY == 1): two stages so that two CTAs
    // fit per SM; the big square tiles keep three
    constexpr int STAGES = (TY == 1) ? 2 : (TY == 2) ? 4 : 3;
    using C = DistCfg<TY, MG, MV, CPR, STAGES>;
    auto kern = k2_sad_v<TY, MG, MV, CPR, STAGES, BULK>;
    static int occ_dev[64] = {0};
    int &occ = occ_dev[e->device & 63];
    if (occ == 0) {
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NVB_DIST_THREADS, C::SMEM));
        if (occ < 1) occ = 1;
    }
    const int n_gt = (da.G + C::TG - 1) / C::TG;
    const int n_vt = (da.N + C::TN - 1) / C::TN;
    const long long units = (long long)n_gt * n_vt;
    long long n_cta = (long long)e->sm_count * occ;
    if (n_cta > units) n_cta = units;
    int rc = build_spans(e, n_gt, n_vt, (int)n_cta);
    if (rc) return rc;
    da.n_vt = n_vt;
    da.vt_per_split = 0;
    da.spans = e->d_spans;
    CK(launch_seq(kern, dim3((unsigned)n_cta), dim3(NVB_DIST_THREADS), C::SMEM, e->stream, da));
    e->launches++;
    CK(cudaGetLastError());
    return NVB_OK;
}

template <int TY, int MG, int MV, int CPR>
static int launch_dist_cfg(nvb_engine *e, const DistArgs &da)
{
    // odd chunk counts below 8 mean rows of exactly 16*CPR bytes (nk == 1): contiguous
    // tiles, staged by TMA bulk copies; everything else goes through swizzled cp.async
    constexpr bool BULK = (CPR % 2) == 1 && CPR < 8;
    return launch_dist_cfg2<TY, MG, MV, CPR, BULK>(e, da);
}

// Few glimpses against a large library: every warp streams its own range of views (k2_stream).
template <int CPR, int GMAX, bool EXACT>
static int launch_stream(nvb_engine *e, const DistArgs &da)
{
    auto kern = k2_stream<CPR, GMAX, EXACT>;
    const int smem = nvb_stream_smem(16 * CPR, GMAX);
    static int occ_dev[64] = {0};
    int &occ = occ_dev[e->device & 63];
    if (occ == 0) {
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NVB_STREAM_THREADS, smem));
        if (occ < 1) occ = 1;
    }
    // one wave; a warp wants at least a few chunks of 64 views
    long long n_cta = (long long)e->sm_count * occ;
    const long long want = (da.N + 4LL * NVB_STREAM_VPC * (NVB_STREAM_THREADS / 32) - 1) / (4LL * NVB_STREAM_VPC * (NVB_STREAM_THREADS / 32));
    if (n_cta > want) n_cta = want;
    if (n_cta < 1) n_cta = 1;
    CK(launch_seq(kern, dim3((unsigned)n_cta), dim3(NVB_STREAM_THREADS), (size_t)smem, e->stream, da));
    e->launches++;
    CK(cudaGetLastError());
    return NVB_OK;
}

static bool use_stream(const nvb_engine *e, const DistArgs &da)
{
    static const bool off = getenv("NAVSIM_B200_NO_STREAM") != nullptr;
    // (a warp's range of views must fit the index bits of the kernel's packed per-thread key)
    return !off && da.G <= 16 && e->nk == 1 && (e->cpr & 1) && da.N >= 16384 &&
           da.N / ((long long)e->sm_count * (NVB_STREAM_THREADS / 32)) < (1 << NVB_STREAM_VBITS) - NVB_STREAM_VPC;
}

template <int CPR>
static int launch_dist_cpr(nvb_engine *e, const DistArgs &da)
{
    if constexpr ((CPR & 1) != 0) if (use_stream(e, da)) {
        if (da.G == 10) return launch_stream<CPR, 10, true>(e, da);   // the reference's default sweep
        if (da.G < 10) return launch_stream<CPR, 10, false>(e, da);
        return launch_stream<CPR, 16, false>(e, da);
    }
    // few glimpses (one agent's heading sweep): G x 512 tiles, the library streams through
    // once; the row count is matched to the reference's default sweep of 10 headings
    if (da.G <= 10) return launch_dist_cfg<2, 5, 2, CPR>(e, da);       // 10 x 256 tiles, 4 stages in flight
    if (da.G <= 16) return launch_dist_cfg<1, 16, 2, CPR>(e, da);
    if (da.G <= 64) return launch_dist_cfg<4, 8, 4, CPR>(e, da);       // 32 x 256
    // many glimpses: 64-glimpse x 240- or 256-view tiles (4 x 15|16 sums per thread);
    // the width that wastes fewer padded views wins
    const long long pad15 = ((da.N + 239) / 240) * 240LL, pad16 = ((da.N + 255) / 256) * 256LL;
    // 64- or 48-glimpse tiles: whichever leaves the busiest SM with less to do (tiles are
    // dealt out whole; the smaller tile pays ~2 % in shared-memory traffic per comparison).
    // C2 (10240 glimpses x 6 view tiles on 148 SMs): 7 x 64 vs 9 x 48 rows -> 48.
    static const int forced_mg = getenv("NAVSIM_B200_K2_MG") ? atoi(getenv("NAVSIM_B200_K2_MG")) : 0;
    const long long n_vt = (pad15 < pad16 ? pad15 / 240 : pad16 / 256), sms = e->sm_count;
    const long long busiest4 = ((((da.G + 63) / 64) * n_vt + sms - 1) / sms) * 64;
    const long long busiest3 = ((((da.G + 47) / 48) * n_vt + sms - 1) / sms) * 48;
    const int mg = forced_mg ? forced_mg : (busiest3 * 102 < busiest4 * 100 ? 3 : 4);
    if (mg == 3) {
        if (pad15 < pad16) return launch_dist_cfg<16, 3, 15, CPR>(e, da);
        return launch_dist_cfg<16, 3, 16, CPR>(e, da);
    }
    if (pad15 < pad16) return launch_dist_cfg<16, 4, 15, CPR>(e, da);
    return launch_dist_cfg<16, 4, 16, CPR>(e, da);
}

template <int KCH, int STAGES, bool TILEMIN>
static int launch_tc_cfg(nvb_engine *e, TcArgs ta)
{
    using C = TcCfg<KCH, NVB_TC_NT, STAGES>;
    auto kern = k2_tc<KCH, NVB_TC_NT, STAGES, TILEMIN>;
    static bool attr_set[64] = {false};
    if (!attr_set[e->device & 63]) {
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
        attr_set[e->device & 63] = true;
    }
    const int n_gt = (ta.G + C::TM - 1) / C::TM, n_vt = (ta.N + NVB_TC_NT - 1) / NVB_TC_NT;
    const long long items = (long long)n_gt * n_vt;
    const int n_cta = (int)(items < e->sm_count ? items : e->sm_count);
    if (!(e->span_tc_key[0] == n_gt && e->span_tc_key[1] == n_vt && e->span_tc_key[2] == n_cta)) {
        std::vector<int> spans(n_cta + 1);
        const long long base = items / n_cta, rem = items % n_cta;
        long long u = 0;
        for (int c = 0; c < n_cta; c++) { spans[c] = (int)u; u += base + (c < rem ? 1 : 0); }
        spans[n_cta] = (int)items;
        int rc = alloc_dev(&e->d_spans_tc, (size_t)n_cta + 1);
        if (rc) return rc;
        CK(cudaMemcpyAsync(e->d_spans_tc, spans.data(), sizeof(int) * (n_cta + 1), cudaMemcpyHostToDevice, e->stream));
        CK(cudaStreamSynchronize(e->stream));   // spans is a stack vector
        e->span_tc_key[0] = n_gt; e->span_tc_key[1] = n_vt; e->span_tc_key[2] = n_cta;
        e->graph_dirty = true;
    }
    ta.n_vt = n_vt;
    ta.n_gt = n_gt;
    ta.vt_major = (n_gt > 1 && (long long)ta.N * e->tc_Kpad > (32ll << 20)) ? 1 : 0;   // library beyond L2 reach: reuse view tiles
    ta.kchunks = e->tc_Kpad / KCH;
    ta.spans = e->d_spans_tc;
    CK(launch_seq(kern, dim3((unsigned)n_cta), dim3(NVB_TC_THREADS), (size_t)C::SMEM, e->stream, e->tm_genc, e->tm_lenc, ta));
    e->launches++;
    CK(cudaGetLastError());
    return NVB_OK;
}

template <bool TILEMIN, bool LEAN>
static int launch_tc_bs_v(nvb_engine *e, TcArgs ta)
{
    auto kern = k2_tc_bs<TILEMIN, NVB_TCBS_KCH, LEAN>;
    const int kchunks = e->tc_Kpad / NVB_TCBS_KCH;
    int a_stages = nvb_tcbs_slots(kchunks);   // glimpse slots (whole items) beside the resident view tile
    static const int max_slots = getenv("NAVSIM_B200_TCBS_SLOTS") ? atoi(getenv("NAVSIM_B200_TCBS_SLOTS")) : 0;   // tuning knob
    if (max_slots >= 2 && a_stages > max_slots) a_stages = max_slots;
    const int smem = nvb_tcbs_smem(kchunks, a_stages);
    static int attr_set[64] = {0};
    if (attr_set[e->device & 63] < smem) {
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_set[e->device & 63] = smem;
    }
    const int n_gt = (ta.G + NVB_TC_TM - 1) / NVB_TC_TM, n_vt = (ta.N + NVB_TC_NT - 1) / NVB_TC_NT;
    const long long items = (long long)n_gt * n_vt;
    const int n_cta = (int)(items < e->sm_count ? items : e->sm_count);
    if (!(e->span_tc_key[0] == n_gt && e->span_tc_key[1] == n_vt && e->span_tc_key[2] == n_cta)) {
        std::vector<int> spans(n_cta + 1);
        const long long base = items / n_cta, rem = items % n_cta;
        long long u = 0;
        for (int c = 0; c < n_cta; c++) { spans[c] = (int)u; u += base + (c < rem ? 1 : 0); }
        spans[n_cta] = (int)items;
        int rc = alloc_dev(&e->d_spans_tc, (size_t)n_cta + 1);
        if (rc) return rc;
        CK(cudaMemcpyAsync(e->d_spans_tc, spans.data(), sizeof(int) * (n_cta + 1), cudaMemcpyHostToDevice, e->stream));
        CK(cudaStreamSynchronize(e->stream));   // spans is a stack vector
        e->span_tc_key[0] = n_gt; e->span_tc_key[1] = n_vt; e->span_tc_key[2] = n_cta;
        e->graph_dirty = true;
    }
    ta.n_vt = n_vt;
    ta.n_gt = n_gt;
    ta.vt_major = 1;
    ta.kchunks = kchunks;
    ta.spans = e->d_spans_tc;
    CK(launch_seq(kern, dim3((unsigned)n_cta), dim3(LEAN ? NVB_TCBS_THREADS : 224), (size_t)smem, e->stream, e->tm_genc, e->tm_lenc, ta, a_stages));
    e->launches++;
    CK(cudaGetLastError());
    return NVB_OK;
}

// few items per CTA: the register-lean epilogue (step-kernel CTAs fit beside the kernel); many: the
// epilogue with both halves of its columns in flight
template <bool TILEMIN>
static int launch_tc_bs(nvb_engine *e, TcArgs ta)
{
    const long long items = (long long)((ta.G + NVB_TC_TM - 1) / NVB_TC_TM) * ((ta.N + NVB_TC_NT - 1) / NVB_TC_NT);
    static const int forced = getenv("NAVSIM_B200_TCBS_LEAN") ? atoi(getenv("NAVSIM_B200_TCBS_LEAN")) : -1;   // tuning knob
    const bool lean = forced >= 0 ? forced != 0 : items <= 8LL * e->sm_count;
    return lean ? launch_tc_bs_v<TILEMIN, true>(e, ta) : launch_tc_bs_v<TILEMIN, false>(e, ta);
}

// K2 on the tensor cores.  encode_glimpses: the glimpse planes were not written by the sampler
// (queries uploaded from the host): encode them from the V plane first.
static int launch_distance_tc(nvb_engine *e, int G, bool bump_step, bool encode_glimpses)
{
    if (encode_glimpses) {
        const long long n = (long long)G * e->P;
        k_tc_encode<true><<<(unsigned)((n + 255) / 256), 256, 0, e->stream>>>(e->d_gv, G, e->P, e->Ppad, e->tc_Kpad,
                                                                             e->tc_planes, e->d_tc_level_of, e->d_genc,
                                                                             e->d_tc_bad);
        e->launches++;
        CK(cudaGetLastError());
    }
    TcArgs ta{};
    ta.G = G; ta.N = e->N;
    ta.keys = e->d_keys;
    ta.view_offset = e->view_offset;
    ta.sad_const = e->tc_sad_const;
    ta.step_counter = bump_step ? e->d_step : nullptr;
    ta.tie_count = bump_step ? e->d_tie_count : nullptr;
    ta.epoch = (bump_step && e->p2p_on) ? e->d_p2p_seq : nullptr;
    ta.pdl_early = (bump_step && early_trigger()) ? 1 : 0;
    ta.tl = bump_step ? e->d_tl : nullptr;
    ta.tmin = nullptr;
    if (bump_step && e->want_tmin) {
        // the single-launch step: the minimum of every view tile instead of the packed keys (the
        // array is allocated by prepare_step_buffers before the step arguments are built)
        ta.tmin = e->d_tmin;
        if (e->tc_bs) return launch_tc_bs<true>(e, ta);
        if (e->tc_kch == 128) return launch_tc_cfg<128, 4, true>(e, ta);
        return launch_tc_cfg<64, 8, true>(e, ta);
    }
    if (e->tc_bs) return launch_tc_bs<false>(e, ta);
    if (e->tc_kch == 128) return launch_tc_cfg<128, 4, false>(e, ta);
    return launch_tc_cfg<64, 8, false>(e, ta);
}

// chem_weight > 0 on register tiles (16 glimpses x 128 views per CTA, 2 x 4 pairs per thread)
template <int CPR>
static int launch_hsv_tiled(nvb_engine *e, DistArgs da)
{
    constexpr int TY = 8, MG = 2, MV = 4, STAGES = 3;
    using C = HsvCfg<TY, MG, MV, CPR, STAGES>;
    auto kern = k2_sad_hsv_t<TY, MG, MV, CPR, STAGES>;
    static int occ_dev[64] = {0};
    int &occ = occ_dev[e->device & 63];
    if (occ == 0) {
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NVB_DIST_THREADS, C::SMEM));
        if (occ < 1) occ = 1;
    }
    const int n_gt = (da.G + C::TG - 1) / C::TG, n_vt = (da.N + C::TN - 1) / C::TN;
    const long long units = (long long)n_gt * n_vt;
    long long n_cta = (long long)e->sm_count * occ;
    if (n_cta > units) n_cta = units;
    int rc = build_spans(e, n_gt, n_vt, (int)n_cta);
    if (rc) return rc;
    da.n_vt = n_vt;
    da.vt_per_split = 0;
    da.spans = e->d_spans;
    CK(launch_seq(kern, dim3((unsigned)n_cta), dim3(NVB_DIST_THREADS), (size_t)C::SMEM, e->stream, da));
    e->launches++;
    CK(cudaGetLastError());
    return NVB_OK;
}

// K2 over glimpses [0, G) of the engine's glimpse buffers; keys must be reset.
// glimpses_encoded: the sampler of the stepping loop wrote the thermometer planes as well.
static int launch_distance(nvb_engine *e, int G, bool bump_step = false, bool glimpses_encoded = false)
{
    if (use_tc(e, G)) {
        // (first use: encodes the library, outside any stream capture -- every graph is captured
        // after one step with plain launches)
        int rc = ensure_tc_library(e);
        if (rc) return rc;
        if (e->tc_lib_ok) return launch_distance_tc(e, G, bump_step, !glimpses_encoded);
    }
    DistArgs da;
    da.gv = e->d_gv; da.gh = e->d_gh; da.gs = e->d_gs;
    da.lv = e->d_lv; da.lh = e->d_lh; da.ls = e->d_ls;
    da.G = G; da.N = e->N; da.Ppad = e->Ppad; da.nk = e->nk;
    da.n_vt = 0; da.vt_per_split = 0; da.spans = nullptr;
    da.step_counter = bump_step ? e->d_step : nullptr;
    da.tie_count = bump_step ? e->d_tie_count : nullptr;
    da.epoch = (bump_step && e->p2p_on) ? e->d_p2p_seq : nullptr;
    da.view_offset = e->view_offset;
    da.keys = e->d_keys;
    da.cw = e->cw;
    da.idx_bits = (e->cw == 0.0) ? 32 : 28;
    da.pdl_early = (bump_step && early_trigger()) ? 1 : 0;
    da.tl = bump_step ? e->d_tl : nullptr;
    static const bool hsv_untiled = getenv("NAVSIM_B200_HSV_UNTILED") != nullptr;
    if (e->cw != 0.0 && G >= 16 && !hsv_untiled) {
        switch (e->cpr) {
    #ifndef NVB_DEV_MIN
    case 1: return launch_hsv_tiled<1>(e, da);
#endif
    #ifndef NVB_DEV_MIN
    case 2: return launch_hsv_tiled<2>(e, da);
#endif
    #ifndef NVB_DEV_MIN
    case 3: return launch_hsv_tiled<3>(e, da);
#endif
    #ifndef NVB_DEV_MIN
    case 4: return launch_hsv_tiled<4>(e, da);
#endif
        case 5: return launch_hsv_tiled<5>(e, da);
    #ifndef NVB_DEV_MIN
    case 7: return launch_hsv_tiled<7>(e, da);
#endif
    #ifndef NVB_DEV_MIN
    case 8: return launch_hsv_tiled<8>(e, da);
#endif
        }
        return fail(NVB_E_INVALID, "unsupported chunk count %d", e->cpr);
    }
    if (e->cw != 0.0) {
        const size_t smem = (size_t)3 * NVB_HSV_TG * e->Ppad;
        static size_t attr_set[64] = {0};
        if (smem > 48 * 1024 && smem > attr_set[e->device & 63]) {
            CK(cudaFuncSetAttribute(k2_sad_hsv, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr_set[e->device & 63] = smem;
        }
        const int n_gt = (G + NVB_HSV_TG - 1) / NVB_HSV_TG;
        const int n_vt = (e->N + NVB_HSV_THREADS - 1) / NVB_HSV_THREADS;
        int splits = (int)((4LL * e->sm_count + n_gt - 1) / n_gt);
        if (splits < 1) splits = 1;
        if (splits > n_vt) splits = n_vt;
        da.n_vt = n_vt;
        da.vt_per_split = (n_vt + splits - 1) / splits;
        dim3 grid(n_gt, (n_vt + da.vt_per_split - 1) / da.vt_per_split);
        k2_sad_hsv<<<grid, NVB_HSV_THREADS, smem, e->stream>>>(da);
        e->launches++;
        CK(cudaGetLastError());
        return NVB_OK;
    }
    switch (e->cpr) {
#ifndef NVB_DEV_MIN
    case 1: return launch_dist_cpr<1>(e, da);
#endif
#ifndef NVB_DEV_MIN
    case 2: return launch_dist_cpr<2>(e, da);
#endif
#ifndef NVB_DEV_MIN
    case 3: return launch_dist_cpr<3>(e, da);
#endif
#ifndef NVB_DEV_MIN
    case 4: return launch_dist_cpr<4>(e, da);
#endif
    case 5: return launch_dist_cpr<5>(e, da);
#ifndef NVB_DEV_MIN
    case 7: return launch_dist_cpr<7>(e, da);
#endif
#ifndef NVB_DEV_MIN
    case 8: return launch_dist_cpr<8>(e, da);
#endif
    }
    return fail(NVB_E_INVALID, "unsupported chunk count %d", e->cpr);
}

extern "C" int nvb_set_distance_kernel(nvb_engine *e, int mode)
{
    if (mode != 0 && mode != 1) return fail(NVB_E_INVALID, "mode must be 0 (automatic) or 1 (byte SIMD)");
    if (e->tc_off != (mode == 1)) e->graph_dirty = true;
    e->tc_off = (mode == 1);
    e->glimpses_pending = false;   // glimpses sampled ahead may lack the operand planes the other kernel reads
    return NVB_OK;
}

extern "C" int nvb_distance_kernel(nvb_engine *e)
{
    if (e->B <= 0 || e->N <= 0) return 0;
    if (!use_tc(e, (long long)e->B * e->A)) {
        DistArgs da{};
        da.G = e->B * e->A; da.N = e->N;
        return (e->cw == 0.0 && use_stream(e, da)) ? 2 : 0;
    }
    if (cudaSetDevice(e->device) != cudaSuccess) return 0;
    if (ensure_tc_library(e) != NVB_OK) return 0;
    return e->tc_lib_ok ? 1 : 0;
}

__global__ void k_fill_u64(unsigned long long *p, long long n, unsigned long long v)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

static int fill_u64(nvb_engine *e, unsigned long long *p, long long n, unsigned long long v)
{
    if (n <= 0) return NVB_OK;
    k_fill_u64<<<(unsigned)((n + 255) / 256), 256, 0, e->stream>>>(p, n, v);
    e->launches++;
    CK(cudaGetLastError());
    return NVB_OK;
}

// ---------------------------------------------------------------------------
extern "C" int nvb_fill_sensor(nvb_engine *e, uint8_t *sensor, int Hpx, int Wpx, double x, double y,
                               double c, double s)
{
    if (!e->d_land) return fail(NVB_E_INVALID, "no landscape");
    CK(cudaSetDevice(e->device));
    const size_t n = (size_t)Hpx * Wpx;
    uint8_t *d_out = nullptr;
    int *d_err = nullptr;
    CK(cudaMalloc(&d_out, n * 3 + 16));
    CK(cudaMalloc(&d_err, sizeof(int)));
    CK(cudaMemsetAsync(d_err, 0, sizeof(int), e->stream));
    NvbWorld w = make_world(e);
    if (n) {
        k_fill_sensor<<<(unsigned)((n + 127) / 128), 128, 0, e->stream>>>(w, Hpx, Wpx, x, y, c, s, d_out, d_err);
        e->launches++;
    }
    int err = 0;
    CK(cudaMemcpyAsync(&err, d_err, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaMemcpyAsync(sensor, d_out, n * 3, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    cudaFree(d_out);
    cudaFree(d_err);
    return err ? NVB_INDEX_ERROR : NVB_OK;
}

extern "C" int nvb_downscale_chem(nvb_engine *e, const uint8_t *image, int R, int C, int fr, int fc,
                                  uint8_t *out)
{
    if (fr <= 0 || fc <= 0 || R < 0 || C < 0) return fail(NVB_E_INVALID, "bad downscale arguments");
    CK(cudaSetDevice(e->device));
    const size_t nin = (size_t)R * C * 3, nout = (size_t)(R / fr) * (C / fc);
    if (nout == 0) return NVB_OK;
    uint8_t *d_in = nullptr, *d_out = nullptr;
    CK(cudaMalloc(&d_in, nin));
    CK(cudaMalloc(&d_out, nout * 3));
    CK(cudaMemcpyAsync(d_in, image, nin, cudaMemcpyHostToDevice, e->stream));
    k_downscale_chem<<<(unsigned)((nout + 127) / 128), 128, 0, e->stream>>>(d_in, R, C, fr, fc, d_out);
    e->launches++;
    CK(cudaMemcpyAsync(out, d_out, nout * 3, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    cudaFree(d_in);
    cudaFree(d_out);
    return NVB_OK;
}

// pose-mode K1 into the given planes; status to d_status
static int sample_poses(nvb_engine *e, const double *poses, const double *cs, int G, uint8_t *ph,
                        uint8_t *ps, uint8_t *pv, int32_t *d_status)
{
    double *d_poses = nullptr, *d_cs = nullptr;
    CK(cudaMalloc(&d_poses, sizeof(double) * 3 * G));
    CK(cudaMemcpyAsync(d_poses, poses, sizeof(double) * 3 * G, cudaMemcpyHostToDevice, e->stream));
    if (cs) {
        CK(cudaMalloc(&d_cs, sizeof(double) * 2 * G));
        CK(cudaMemcpyAsync(d_cs, cs, sizeof(double) * 2 * G, cudaMemcpyHostToDevice, e->stream));
    }
    SamplerArgs sa;
    sa.w = make_world(e);
    sa.poses = d_poses; sa.offsets = nullptr; sa.cs = d_cs;
    sa.A = 1; sa.agent_mode = 0; sa.need_hs = 1;
    sa.status = d_status; sa.completed = nullptr; sa.budget = nullptr;
    sa.gv = pv; sa.gh = ph; sa.gs = ps;
    sa.keys = nullptr;
    sa.band = sampler_band(e);
    sa.dbg = nullptr;
    sa.poses_src = nullptr; sa.poses_dst = nullptr; sa.pending_clear = nullptr;
    sa.genc = nullptr; sa.tc_tab = nullptr; sa.Kpad = 0; sa.n_planes = 0;
    int rc = launch_sampler(e, sa, G);
    CK(cudaStreamSynchronize(e->stream));
    cudaFree(d_poses);
    if (d_cs) cudaFree(d_cs);
    return rc;
}

extern "C" int nvb_glimpse_batch(nvb_engine *e, const double *poses, const double *cs, int G,
                                 uint8_t *out, int32_t *status)
{
    if (!e->d_land || !e->have_sensor) return fail(NVB_E_INVALID, "landscape and sensor must be set");
    if (G <= 0) return NVB_OK;
    CK(cudaSetDevice(e->device));
    int rc = ensure_glimpse_cap(e, G);
    if (rc) return rc;
    // the resident glimpse buffers are reused as scratch: glimpses sampled ahead for the next
    // step of the stepping loop are overwritten, so that step samples again (phase1 / run_steps)
    e->glimpses_pending = false;
    int32_t *d_status = nullptr;
    uint8_t *d_out = nullptr;
    CK(cudaMalloc(&d_status, sizeof(int32_t) * G));
    CK(cudaMalloc(&d_out, (size_t)G * e->P * 3));
    rc = sample_poses(e, poses, cs, G, e->d_gh, e->d_gs, e->d_gv, d_status);
    if (!rc) {
        const long long n = (long long)G * e->P;
        k_planar_to_hsv<<<(unsigned)((n + 255) / 256), 256, 0, e->stream>>>(e->d_gh, e->d_gs, e->d_gv,
                                                                           e->Ppad, e->P, G, d_out);
        e->launches++;
        CK(cudaMemcpyAsync(out, d_out, (size_t)n * 3, cudaMemcpyDeviceToHost, e->stream));
        if (status)
            CK(cudaMemcpyAsync(status, d_status, sizeof(int32_t) * G, cudaMemcpyDeviceToHost, e->stream));
        CK(cudaStreamSynchronize(e->stream));
    }
    cudaFree(d_status);
    cudaFree(d_out);
    return rc;
}

static int alloc_library(nvb_engine *e, int N)
{
    int rc;
    const size_t bytes = (size_t)N * e->Ppad;
    if ((rc = alloc_dev(&e->d_lh, bytes))) return rc;
    if ((rc = alloc_dev(&e->d_ls, bytes))) return rc;
    if ((rc = alloc_dev(&e->d_lv, bytes))) return rc;
    CK(cudaMemsetAsync(e->d_lh, 0, bytes, e->stream));
    CK(cudaMemsetAsync(e->d_ls, 0, bytes, e->stream));
    CK(cudaMemsetAsync(e->d_lv, 0, bytes, e->stream));
    e->N = N;
    e->view_offset = 0;
    e->n_total = N;
    e->lenc_valid = false;
    return NVB_OK;
}

static int set_path(nvb_engine *e, const double *path, int n)
{
    int rc = alloc_dev(&e->d_path, (size_t)2 * n);
    if (rc) return rc;
    CK(cudaMemcpy(e->d_path, path, sizeof(double) * 2 * n, cudaMemcpyHostToDevice));
    e->n_path = n;
    // Bounding circles of blocks of NVB_PATH_BLOCK consecutive points: update_error first
    // discards whole blocks that cannot hold the nearest point or a covered point (step.cuh).
    // The radius is inflated so that it bounds the true distances whatever the rounding.
    const int nb = (n + NVB_PATH_BLOCK - 1) / NVB_PATH_BLOCK;
    std::vector<double> blk((size_t)4 * nb);
    for (int j = 0; j < nb; j++) {
        const int n0 = j * NVB_PATH_BLOCK, n1 = n0 + NVB_PATH_BLOCK < n ? n0 + NVB_PATH_BLOCK : n;
        double cx = 0, cy = 0;
        for (int i = n0; i < n1; i++) { cx += path[2 * i]; cy += path[2 * i + 1]; }
        cx /= (n1 - n0); cy /= (n1 - n0);
        double r = 0;
        for (int i = n0; i < n1; i++) {
            const double d = hypot(path[2 * i] - cx, path[2 * i + 1] - cy);
            if (d > r) r = d;
        }
        blk[4 * j] = cx; blk[4 * j + 1] = cy; blk[4 * j + 2] = r * (1.0 + 1e-12) + 1e-9; blk[4 * j + 3] = 0.0;
    }
    if ((rc = alloc_dev(&e->d_pblk, (size_t)4 * nb))) return rc;
    CK(cudaMemcpy(e->d_pblk, blk.data(), sizeof(double) * 4 * nb, cudaMemcpyHostToDevice));
    // FP32 copy for the single-launch step's prefilter (step_tm.cuh): the radius also covers the
    // rounding of the centre, and is rounded up
    std::vector<float> blkf((size_t)4 * nb);
    for (int j = 0; j < nb; j++) {
        const float cxf = (float)blk[4 * j], cyf = (float)blk[4 * j + 1];
        const double rr = blk[4 * j + 2] + fabs(blk[4 * j] - (double)cxf) + fabs(blk[4 * j + 1] - (double)cyf);
        float rf = (float)rr;
        while ((double)rf < rr) rf = nextafterf(rf, INFINITY);
        blkf[4 * j] = cxf; blkf[4 * j + 1] = cyf; blkf[4 * j + 2] = nextafterf(rf, INFINITY); blkf[4 * j + 3] = 0.0f;
    }
    if ((rc = alloc_dev(&e->d_pblk_f, (size_t)4 * nb))) return rc;
    CK(cudaMemcpy(e->d_pblk_f, blkf.data(), sizeof(float) * 4 * nb, cudaMemcpyHostToDevice));
    // second level for long paths: one circle per NVB_PATH_GROUP blocks
    const int ng = (nb + NVB_PATH_GROUP - 1) / NVB_PATH_GROUP, gpts = NVB_PATH_GROUP * NVB_PATH_BLOCK;
    std::vector<double> grp((size_t)4 * ng);
    for (int j = 0; j < ng; j++) {
        const int n0 = j * gpts, n1 = n0 + gpts < n ? n0 + gpts : n;
        double cx = 0, cy = 0;
        for (int i = n0; i < n1; i++) { cx += path[2 * i]; cy += path[2 * i + 1]; }
        cx /= (n1 - n0); cy /= (n1 - n0);
        double r = 0;
        for (int i = n0; i < n1; i++) {
            const double d = hypot(path[2 * i] - cx, path[2 * i + 1] - cy);
            if (d > r) r = d;
        }
        grp[4 * j] = cx; grp[4 * j + 1] = cy; grp[4 * j + 2] = r * (1.0 + 1e-12) + 1e-9; grp[4 * j + 3] = 0.0;
    }
    if ((rc = alloc_dev(&e->d_pblk2, (size_t)4 * ng))) return rc;
    CK(cudaMemcpy(e->d_pblk2, grp.data(), sizeof(double) * 4 * ng, cudaMemcpyHostToDevice));
    return NVB_OK;
}

extern "C" int nvb_library_build(nvb_engine *e, const double *path, const double *angles,
                                 const double *cs, int N, int *bad_index)
{
    e->graph_dirty = true;
    if (!e->d_land || !e->have_sensor) return fail(NVB_E_INVALID, "landscape and sensor must be set");
    if (N <= 0 || !path || !angles) return fail(NVB_E_INVALID, "bad path");
    CK(cudaSetDevice(e->device));
    CK(cudaStreamSynchronize(e->stream));
    int rc = alloc_library(e, N);
    if (rc) return rc;
    std::vector<double> poses((size_t)3 * N);
    for (int i = 0; i < N; i++) {
        poses[3 * i] = path[2 * i];
        poses[3 * i + 1] = path[2 * i + 1];
        poses[3 * i + 2] = angles[i];
    }
    int32_t *d_status = nullptr;
    CK(cudaMalloc(&d_status, sizeof(int32_t) * N));
    rc = sample_poses(e, poses.data(), cs, N, e->d_lh, e->d_ls, e->d_lv, d_status);
    std::vector<int32_t> st(N);
    if (!rc) CK(cudaMemcpy(st.data(), d_status, sizeof(int32_t) * N, cudaMemcpyDeviceToHost));
    cudaFree(d_status);
    if (rc) return rc;
    for (int i = 0; i < N; i++)
        if (st[i] != 0) {
            if (bad_index) *bad_index = i;
            e->N = 0;
            return st[i];
        }
    e->B = 0;
    e->ms_capacity = -1;
    return set_path(e, path, N);
}

extern "C" int nvb_library_upload(nvb_engine *e, const uint8_t *scenes, const double *path, int N)
{
    e->graph_dirty = true;
    if (!e->have_sensor) return fail(NVB_E_INVALID, "sensor must be set");
    if (N <= 0 || !scenes) return fail(NVB_E_INVALID, "bad library");
    CK(cudaSetDevice(e->device));
    CK(cudaStreamSynchronize(e->stream));
    int rc = alloc_library(e, N);
    if (rc) return rc;
    uint8_t *d_in = nullptr;
    const long long n = (long long)N * e->P;
    CK(cudaMalloc(&d_in, (size_t)n * 3));
    CK(cudaMemcpyAsync(d_in, scenes, (size_t)n * 3, cudaMemcpyHostToDevice, e->stream));
    k_hsv_to_planar<<<(unsigned)((n + 255) / 256), 256, 0, e->stream>>>(d_in, e->Ppad, e->P, N, e->d_lh,
                                                                       e->d_ls, e->d_lv);
    e->launches++;
    CK(cudaStreamSynchronize(e->stream));
    cudaFree(d_in);
    e->B = 0;
    e->ms_capacity = -1;
    if (path) return set_path(e, path, N);
    e->n_path = 0;
    return NVB_OK;
}

extern "C" int nvb_set_training_path(nvb_engine *e, const double *path, int n)
{
    e->graph_dirty = true;
    if (!path || n <= 0) return fail(NVB_E_INVALID, "bad path");
    CK(cudaSetDevice(e->device));
    CK(cudaStreamSynchronize(e->stream));
    e->B = 0;   // coverage is sized by the path
    e->ms_capacity = -1;
    return set_path(e, path, n);
}

extern "C" int nvb_library_download(nvb_engine *e, uint8_t *scenes)
{
    if (e->N <= 0) return fail(NVB_E_INVALID, "no library");
    CK(cudaSetDevice(e->device));
    uint8_t *d_out = nullptr;
    const long long n = (long long)e->N * e->P;
    CK(cudaMalloc(&d_out, (size_t)n * 3));
    k_planar_to_hsv<<<(unsigned)((n + 255) / 256), 256, 0, e->stream>>>(e->d_lh, e->d_ls, e->d_lv, e->Ppad,
                                                                       e->P, e->N, d_out);
    e->launches++;
    CK(cudaMemcpyAsync(scenes, d_out, (size_t)n * 3, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    cudaFree(d_out);
    return NVB_OK;
}

extern "C" int nvb_library_set_shard(nvb_engine *e, int64_t view_offset, int64_t n_total)
{
    e->graph_dirty = true;
    if (e->N <= 0) return fail(NVB_E_INVALID, "no library");
    if (view_offset < 0 || view_offset + e->N > n_total) return fail(NVB_E_INVALID, "bad shard");
    if (n_total >= (1ll << 28) && e->cw != 0.0) return fail(NVB_E_INVALID, "library too large");
    e->view_offset = view_offset;
    e->n_total = n_total;
    return NVB_OK;
}

// upload G query scenes [G][H][W][3] into the planar glimpse buffers
static int upload_queries(nvb_engine *e, const uint8_t *scenes_q, int G)
{
    int rc = ensure_glimpse_cap(e, G);
    if (rc) return rc;
    e->glimpses_pending = false;   // as nvb_glimpse_batch: the pre-sampled glimpses (and keys) are overwritten
    uint8_t *d_in = nullptr;
    const long long n = (long long)G * e->P;
    CK(cudaMalloc(&d_in, (size_t)n * 3));
    CK(cudaMemcpyAsync(d_in, scenes_q, (size_t)n * 3, cudaMemcpyHostToDevice, e->stream));
    k_hsv_to_planar<<<(unsigned)((n + 255) / 256), 256, 0, e->stream>>>(d_in, e->Ppad, e->P, G, e->d_gh,
                                                                       e->d_gs, e->d_gv);
    e->launches++;
    CK(cudaStreamSynchronize(e->stream));
    cudaFree(d_in);
    return NVB_OK;
}

extern "C" int nvb_familiarity(nvb_engine *e, const uint8_t *scenes_q, int G, double *fam)
{
    if (e->N <= 0) return fail(NVB_E_INVALID, "no library");
    if (G <= 0) return NVB_OK;
    CK(cudaSetDevice(e->device));
    int rc = upload_queries(e, scenes_q, G);
    if (rc) return rc;
    double *d_fam = nullptr;
    const long long n = (long long)G * e->N;
    CK(cudaMalloc(&d_fam, sizeof(double) * n));
    DistArgs da{};
    da.gv = e->d_gv; da.gh = e->d_gh; da.gs = e->d_gs;
    da.lv = e->d_lv; da.lh = e->d_lh; da.ls = e->d_ls;
    da.G = G; da.N = e->N; da.Ppad = e->Ppad; da.cw = e->cw;
    k_familiarity_exact<<<(unsigned)((n + 127) / 128), 128, 0, e->stream>>>(da, e->P, (double)e->P,
                                                                           e->d_div255, d_fam);
    e->launches++;
    CK(cudaMemcpyAsync(fam, d_fam, sizeof(double) * n, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    cudaFree(d_fam);
    return NVB_OK;
}

extern "C" int nvb_familiarity_min(nvb_engine *e, const uint8_t *scenes_q, int G, double *min_diff,
                                   int64_t *view_idx)
{
    if (e->N <= 0) return fail(NVB_E_INVALID, "no library");
    if (G <= 0) return NVB_OK;
    CK(cudaSetDevice(e->device));
    int rc = upload_queries(e, scenes_q, G);
    if (rc) return rc;
    if ((rc = fill_u64(e, e->d_keys, G, KEY_NONE))) return rc;
    std::vector<unsigned long long> keys(G);
    for (int attempt = 0; attempt < 2; attempt++) {
        const bool tc = attempt == 0 && use_tc(e, G);
        if (tc) CK(cudaMemsetAsync(e->d_tc_bad, 0, sizeof(int), e->stream));
        if (tc) {
            if ((rc = launch_distance(e, G))) return rc;
        } else {
            // byte-SIMD kernel: any byte values
            const bool ok = e->tc_ok;
            e->tc_ok = false;
            rc = launch_distance(e, G);
            e->tc_ok = ok;
            if (rc) return rc;
        }
        int bad = 0;
        if (tc) CK(cudaMemcpyAsync(&bad, e->d_tc_bad, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
        CK(cudaMemcpyAsync(keys.data(), e->d_keys, sizeof(unsigned long long) * G, cudaMemcpyDeviceToHost,
                           e->stream));
        CK(cudaStreamSynchronize(e->stream));
        if (!bad) break;
        // a query held a value this sensor's quantisation cannot produce: score again without the planes
        if ((rc = fill_u64(e, e->d_keys, G, KEY_NONE))) return rc;
    }
    const int ib = (e->cw == 0.0) ? 32 : 28;
    for (int g = 0; g < G; g++) {
        const unsigned long long score = keys[g] >> ib;
        if (min_diff) min_diff[g] = (e->cw == 0.0) ? (double)score : (double)score / 4096.0;
        if (view_idx) view_idx[g] = (int64_t)(keys[g] & ((1ull << ib) - 1ull));
    }
    return NVB_OK;
}

// ---------------------------------------------------------------------------
static int ensure_log(nvb_engine *e, int cap, bool want_afam)
{
    const bool need_afam = want_afam && (e->log_afam == nullptr || e->log_A != e->A);
    if (cap <= e->log_cap && !need_afam) return NVB_OK;
    int new_cap = e->log_cap > 0 ? e->log_cap : 64;
    while (new_cap < cap) new_cap *= 2;
    CK(cudaStreamSynchronize(e->stream));
    const size_t B = e->B;
    int16_t *nb = nullptr;
    double *np = nullptr, *ns = nullptr, *na = nullptr;
    CK(cudaMalloc(&nb, sizeof(int16_t) * B * new_cap));
    CK(cudaMalloc(&np, sizeof(double) * 3 * B * new_cap));
    CK(cudaMalloc(&ns, sizeof(double) * B * new_cap));
    const bool keep_afam = want_afam || e->log_afam != nullptr;
    if (keep_afam) {
        CK(cudaMalloc(&na, sizeof(double) * B * e->A * new_cap));
        CK(cudaMemset(na, 0xFF, sizeof(double) * B * e->A * new_cap));
    }
    if (e->steps_done > 0 && e->log_best) {
        const size_t k = e->steps_done;
        CK(cudaMemcpy(nb, e->log_best, sizeof(int16_t) * B * k, cudaMemcpyDeviceToDevice));
        CK(cudaMemcpy(np, e->log_pose, sizeof(double) * 3 * B * k, cudaMemcpyDeviceToDevice));
        CK(cudaMemcpy(ns, e->log_sfam, sizeof(double) * B * k, cudaMemcpyDeviceToDevice));
        if (na && e->log_afam && e->log_A == e->A)
            CK(cudaMemcpy(na, e->log_afam, sizeof(double) * B * e->A * k, cudaMemcpyDeviceToDevice));
    }
    free_dev(e->log_best); free_dev(e->log_pose); free_dev(e->log_sfam); free_dev(e->log_afam);
    e->log_best = nb; e->log_pose = np; e->log_sfam = ns; e->log_afam = na;
    e->log_cap = new_cap;
    e->log_A = e->A;
    return NVB_OK;
}

extern "C" int nvb_agents_set(nvb_engine *e, const double *poses, const int32_t *budget, int B)
{
    if (!e->d_land || !e->have_sensor || e->A <= 0) return fail(NVB_E_INVALID, "world not configured");
    if (e->N <= 0) return fail(NVB_E_INVALID, "no library");
    if (B <= 0 || !poses) return fail(NVB_E_INVALID, "bad agents");
    CK(cudaSetDevice(e->device));
    CK(cudaStreamSynchronize(e->stream));
    int rc;
    if (e->p2p_on && (B != e->p2p.B || (long long)B * e->A > e->p2p.cap)) {
        // the exchange area exported to the peers is too small for this batch: the peers hold
        // mappings of the old area, so the exchange is switched off until a new export + attach
        e->p2p_on = false;
        e->graph_dirty = true;
    }
    if (B != e->B) {
        e->graph_dirty = true;
        if ((rc = alloc_dev(&e->ag.poses, (size_t)3 * B))) return rc;
        if ((rc = alloc_dev(&e->ag.status, (size_t)B))) return rc;
        if ((rc = alloc_dev(&e->ag.completed, (size_t)B))) return rc;
        if ((rc = alloc_dev(&e->ag.budget, (size_t)B))) return rc;
        if ((rc = alloc_dev(&e->ag.nav_frames, (size_t)B))) return rc;
        if ((rc = alloc_dev(&e->ag.err_sum, (size_t)B))) return rc;
        if ((rc = alloc_dev(&e->ag.err_n, (size_t)B))) return rc;
        if ((rc = alloc_dev(&e->ag.stepped, (size_t)B))) return rc;
        if ((rc = alloc_dev(&e->ag.coverage, (size_t)B * (e->n_path > 0 ? e->n_path : 1)))) return rc;
        if ((rc = alloc_dev(&e->d_poses0, (size_t)3 * B))) return rc;
        if ((rc = alloc_dev(&e->d_budget0, (size_t)B))) return rc;
        if ((rc = alloc_dev(&e->d_pending, (size_t)B))) return rc;
        if ((rc = alloc_dev(&e->d_dmin2, (size_t)B))) return rc;
        free_dev(e->log_best); free_dev(e->log_pose); free_dev(e->log_sfam); free_dev(e->log_afam);
        e->log_best = nullptr; e->log_pose = e->log_sfam = e->log_afam = nullptr;
        e->log_cap = 0;
        e->B = B;
    }
    if ((rc = ensure_glimpse_cap(e, (long long)B * e->A))) return rc;
    CK(cudaMemcpyAsync(e->ag.poses, poses, sizeof(double) * 3 * B, cudaMemcpyHostToDevice, e->stream));
    if (budget) {
        CK(cudaMemcpyAsync(e->ag.budget, budget, sizeof(int32_t) * B, cudaMemcpyHostToDevice, e->stream));
    } else {
        std::vector<int32_t> big(B, 0x7FFFFFFF);
        CK(cudaMemcpyAsync(e->ag.budget, big.data(), sizeof(int32_t) * B, cudaMemcpyHostToDevice, e->stream));
        CK(cudaStreamSynchronize(e->stream));
    }
    CK(cudaMemcpyAsync(e->d_poses0, e->ag.poses, sizeof(double) * 3 * B, cudaMemcpyDeviceToDevice, e->stream));
    CK(cudaMemcpyAsync(e->d_budget0, e->ag.budget, sizeof(int32_t) * B, cudaMemcpyDeviceToDevice, e->stream));
    CK(cudaMemsetAsync(e->ag.status, 0, sizeof(int32_t) * B, e->stream));
    CK(cudaMemsetAsync(e->ag.completed, 0, sizeof(int32_t) * B, e->stream));
    CK(cudaMemsetAsync(e->ag.nav_frames, 0, sizeof(int32_t) * B, e->stream));
    CK(cudaMemsetAsync(e->ag.err_sum, 0, sizeof(double) * B, e->stream));
    CK(cudaMemsetAsync(e->ag.err_n, 0, sizeof(int32_t) * B, e->stream));
    CK(cudaMemsetAsync(e->ag.stepped, 0, sizeof(int32_t) * B, e->stream));
    CK(cudaMemsetAsync(e->ag.coverage, 0, (size_t)B * (e->n_path > 0 ? e->n_path : 1), e->stream));
    CK(cudaMemsetAsync(e->d_step, 0xFF, sizeof(int), e->stream));   // -1: first step bumps it to 0
    CK(cudaMemsetAsync(e->d_pending, 0, sizeof(int32_t) * B, e->stream));
    CK(cudaMemsetAsync(e->d_tie_ready, 0, sizeof(int) * (size_t)B * e->A, e->stream));   // epochs restart
    e->glimpses_pending = false;
    CK(cudaMemsetAsync(e->d_tie_count, 0, 2 * sizeof(int), e->stream));
    e->steps_done = 0;
    CK(cudaStreamSynchronize(e->stream));
    return NVB_OK;
}

static int ensure_tc_library(nvb_engine *e);

// Buffers the step kernels are handed by value must exist before the arguments are built:
// the encoded library (decides whether the tensor-core kernel runs) and, for the
// single-launch step, the tile-minimum array.
static void prepare_step_buffers(nvb_engine *e)
{
    if (!use_tc(e, (long long)e->B * e->A) || e->N <= 0) return;
    if (ensure_tc_library(e) != NVB_OK) return;
    if (!tm_form(e)) return;
    const long long n_vt = (e->N + NVB_TC_NT - 1) / NVB_TC_NT, need = (long long)e->Gcap * n_vt;
    if (need > e->tmin_cap) {
        if (alloc_dev(&e->d_tmin, (size_t)need) != NVB_OK) { e->d_tmin = nullptr; e->tmin_cap = 0; return; }
        e->tmin_cap = need;
        e->graph_dirty = true;
    }
}

static StepArgs make_step_args(nvb_engine *e, int fake, int log_afam)
{
    prepare_step_buffers(e);
    StepArgs s;
    s.ag = e->ag;
    s.B = e->B; s.A = e->A; s.N = e->N; s.P = e->P; s.Ppad = e->Ppad;
    s.offsets = e->d_offsets;
    s.gv = e->d_gv; s.gh = e->d_gh; s.gs = e->d_gs;
    s.lv = e->d_lv; s.lh = e->d_lh; s.ls = e->d_ls;
    s.path = e->d_path; s.n_path = e->n_path;
    s.pblk = getenv("NAVSIM_B200_NO_PATH_BLOCKS") ? nullptr : e->d_pblk;
    s.pblk2 = s.pblk ? e->d_pblk2 : nullptr;
    s.view_offset = e->view_offset;
    s.no_sample = 0;
    s.keys = e->d_keys; s.exact = e->d_exact;
    s.idx_bits = (e->cw == 0.0) ? 32 : 28;
    s.band = (e->cw == 0.0) ? 0ull : 1ull;
    s.cw = e->cw;
    s.div255 = e->d_div255;
    s.maxfam = (double)e->P;
    s.step_size = e->step_size; s.max_dist = e->max_dist;
    s.threshold_factor = e->thf; s.coverage_factor = e->cvf;
    s.fake = fake;
    s.tie_count = e->d_tie_count; s.tie_items = e->d_tie_items; s.tie_thr = e->d_tie_thr;
    s.tie_next = e->d_tie_next; s.tie_ready = e->d_tie_ready;
    s.step_counter = e->d_step;
    s.log_cap = e->log_cap;
    s.log_best = e->log_best; s.log_pose = e->log_pose; s.log_sfam = e->log_sfam;
    s.log_afam = log_afam ? e->log_afam : nullptr;
    s.pending_fail = e->d_pending;
    s.dbg = e->d_dbg;
    s.dmin2 = e->d_dmin2;
    s.pdl_early = early_trigger() ? 1 : 0;
    s.tl = e->d_tl;
    s.out_best = e->zc_best; s.out_poses = e->zc_pose; s.out_sfam = e->zc_fam;
    s.tmin = e->d_tmin;
    s.pblk_f = e->d_pblk_f;
    s.n_vt = (e->N + NVB_TC_NT - 1) / NVB_TC_NT;
    s.sad_const = e->tc_sad_const;
    s.p2p = e->p2p;
    if (!e->p2p_on) s.p2p.world = 0;
    s.p2p.B = e->B;
    {   // thr2 = the largest double whose (correctly rounded) square root is <= thr, so that
        // d2 <= thr2  <=>  sqrt(d2) <= thr  (NavBySceneFamiliarity.py:271-276)
        const double thr = e->cvf * e->step_size;
        double thr2 = thr * thr;
        if (thr2 > 0.0 && std::isfinite(thr2)) {
            while (std::sqrt(thr2) > thr) thr2 = std::nextafter(thr2, 0.0);
            while (std::sqrt(std::nextafter(thr2, INFINITY)) <= thr) thr2 = std::nextafter(thr2, INFINITY);
        }
        s.cover_thr2 = thr2;
    }
    return s;
}

static int launch_distance_timed(nvb_engine *e, int G);

static SamplerArgs agent_sampler_args(nvb_engine *e)
{
    SamplerArgs sa;
    sa.w = make_world(e);
    sa.poses = e->ag.poses; sa.offsets = e->d_offsets; sa.cs = nullptr;
    sa.A = e->A; sa.agent_mode = 1; sa.need_hs = (e->cw != 0.0);
    sa.status = e->ag.status; sa.completed = e->ag.completed; sa.budget = e->ag.budget;
    sa.gv = e->d_gv; sa.gh = e->d_gh; sa.gs = e->d_gs;
    sa.keys = e->d_keys;
    sa.band = sampler_band(e);
    sa.dbg = e->d_dbg;
    sa.poses_src = e->zc_in; sa.poses_dst = e->ag.poses; sa.pending_clear = e->d_pending;
    const bool tc = use_tc(e, (long long)e->B * e->A) && (!e->lenc_valid || e->tc_lib_ok);
    sa.genc = tc ? e->d_genc : nullptr;
    sa.tc_tab = tc ? e->d_tc_tab : nullptr;
    sa.Kpad = e->tc_Kpad; sa.n_planes = e->tc_planes.n_planes;
    return sa;
}

static int phase1(nvb_engine *e)
{
    if (!e->glimpses_pending) {
        int rc = launch_sampler(e, agent_sampler_args(e), e->B, sampler_slices(e));
        if (rc) return rc;
    }
    e->glimpses_pending = false;
    return launch_distance_timed(e, e->B * e->A);
}

// decide + ties + move of this step and, in the same launch, the glimpses of the next
template <bool HS, int PH, int PW>
static int launch_k31_t(nvb_engine *e, const StepArgs &s, const SamplerArgs &sa, size_t smem)
{
    static size_t attr_set[64] = {0};
    size_t &cur = attr_set[e->device & 63];
    if (smem > cur) {
        CK(cudaFuncSetAttribute(k31_step_sample<HS, PH, PW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cur = smem;
    }
    CK(launch_seq(k31_step_sample<HS, PH, PW>, dim3(e->B), dim3(NVB_STEP_THREADS), smem, e->stream, e->tmap, s, sa));
    e->launches++;
    CK(cudaGetLastError());
    return NVB_OK;
}

// The form actually run for this engine: form 4 only while every CTA of move+sample (one per
// agent) is resident at once -- its tie agents wait for work done by the other CTAs.
static int effective_form(const nvb_engine *e)
{
    int f = step_form();
    if (f == 5) f = 3;   // (the single-launch form is chosen by tm_form(); everything else runs as form 3)
    if (f == 4 && e->ms_capacity >= 0 && e->B > e->ms_capacity) return 3;
    return f;
}

// Grid-wide tie pass: one CTA per (tied glimpse, view chunk) for libraries that sit in L2, the
// view-major form (library read once per launch) for large ones.
static cudaError_t launch_ties(nvb_engine *e, const StepArgs &s)
{
    static const bool no_v = getenv("NAVSIM_B200_NO_TIES_V") != nullptr;
    if (e->cw == 0.0 && e->Ppad <= 16 * NVB_TIEV_MAX_CHUNKS && (long long)e->N * e->Ppad > (8ll << 20) && !no_v)
        return launch_seq(k3_ties_v, dim3(e->sm_count * 2), dim3(NVB_TIE_THREADS), 0, e->stream, s);
    return launch_seq(k3_ties, dim3(e->sm_count * 2), dim3(NVB_TIE_THREADS), 0, e->stream, s);
}

template <bool HS, int PH, int PW>
static int launch_k3ms_t(nvb_engine *e, const StepArgs &s, const SamplerArgs &sa, size_t smem)
{
    static size_t attr_set[64] = {0};
    size_t &cur = attr_set[e->device & 63];
    if (smem > cur) {
        CK(cudaFuncSetAttribute(k3_move_sample<HS, PH, PW, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CK(cudaFuncSetAttribute(k3_move_sample<HS, PH, PW, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cur = smem;
    }
    // 128 or 160 threads, whichever leaves fewer idle lanes in the gather loop over the
    // A * P sensor pixels (the default 10 x 80 = 800 is exactly five passes of 160)
    const long long items = (long long)e->A * e->P;
    const long long w128 = ((items + 127) / 128) * 128, w160 = ((items + 159) / 160) * 160;
    static const int forced = getenv("NAVSIM_B200_MS_THREADS") ? atoi(getenv("NAVSIM_B200_MS_THREADS")) : 0;
    const int threads = forced ? forced : (w160 < w128 ? NVB_MS_MAX_THREADS : NVB_STEP_THREADS);
    if (e->ms_capacity < 0) {
        int occ = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k3_move_sample<HS, PH, PW, true>, threads, smem));
        e->ms_capacity = occ * e->sm_count;
    }
    const int form = effective_form(e);
    cudaEvent_t *tev = nullptr;
    if (e->timing && getenv("NAVSIM_B200_TIME_ALL")) {   // tuning aid: events around every kernel
        if (e->ev3.size() < e->ev3_used + 4) {
            for (int i = 0; i < 256; i++) { cudaEvent_t ev; cudaEventCreate(&ev); e->ev3.push_back(ev); }
        }
        tev = &e->ev3[e->ev3_used];
        e->ev3_used += 4;
        cudaEventRecord(tev[0], e->stream);
    }
    if (form >= 3) {   // decide | [grid-wide tie pass |] move + sample
        CK(launch_seq(k3_decide, dim3(e->B), dim3(NVB_STEP_THREADS), 0, e->stream, s));
        if (tev) cudaEventRecord(tev[1], e->stream);
        if (form == 3) {
            CK(launch_ties(e, s));
            e->launches += 1;
        }
        if (tev) cudaEventRecord(tev[2], e->stream);
    } else {
        k3_decide_help<<<e->B, NVB_STEP_THREADS, 0, e->stream>>>(s);
    }
    if (form == 4)
        CK(launch_seq(k3_move_sample<HS, PH, PW, true>, dim3(e->B), dim3(threads), smem, e->stream, e->tmap, s, sa));
    else
        CK(launch_seq(k3_move_sample<HS, PH, PW, false>, dim3(e->B), dim3(threads), smem, e->stream, e->tmap, s, sa));
    if (tev) cudaEventRecord(tev[3], e->stream);
    e->launches += 2;
    CK(cudaGetLastError());
    return NVB_OK;
}

template <int PH, int PW>
static int launch_k3tm_t(nvb_engine *e, const StepArgs &s, SamplerArgs sa, size_t smem)
{
    smem = (size_t)nvb_round_up((int)smem, 16) + sizeof(int2) * (size_t)s.A * s.n_vt;   // + the agent's tile candidates
    static size_t attr_set[64] = {0};
    size_t &cur = attr_set[e->device & 63];
    auto kern = k3_step_tm<PH, PW>;
    if (smem > cur) {
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cur = smem;
    }
    const long long items = (long long)e->A * e->P;
    const long long w128 = ((items + 127) / 128) * 128, w160 = ((items + 159) / 160) * 160;
    const int threads = w160 < w128 ? NVB_MS_MAX_THREADS : NVB_STEP_THREADS;
    sa.keys = nullptr;   // the packed keys are not used in this form: nothing to reset
    CK(launch_seq(kern, dim3(e->B), dim3(threads), smem, e->stream, e->tmap, s, sa));
    e->launches += 1;
    CK(cudaGetLastError());
    return NVB_OK;
}

// [decide + cooperative tie scan] then [move + sample next]
static int launch_k3_split(nvb_engine *e, const StepArgs &s, bool tm)
{
    const SamplerArgs sa = agent_sampler_args(e);
    const int nplanes = sa.need_hs ? 3 : 1;
    const size_t smem = e->R > 0 ? nvb_sampler_smem(e->BW, e->BH, nplanes, sa.A)
                                 : nvb_sampler_smem(0, 0, 0, sa.A);
    if (tm) {   // K2 left tile minima (chem_weight 0: V plane only)
        if (e->ph == 4 && e->pw == 2) return launch_k3tm_t<4, 2>(e, s, sa, smem);
#ifndef NVB_DEV_MIN
        if (e->ph == 2 && e->pw == 2) return launch_k3tm_t<2, 2>(e, s, sa, smem);
        if (e->ph == 1 && e->pw == 1) return launch_k3tm_t<1, 1>(e, s, sa, smem);
#endif
        return launch_k3tm_t<0, 0>(e, s, sa, smem);
    }
#ifndef NVB_DEV_MIN
    if (sa.need_hs) return launch_k3ms_t<true, 0, 0>(e, s, sa, smem);
#endif
    if (e->ph == 4 && e->pw == 2) return launch_k3ms_t<false, 4, 2>(e, s, sa, smem);
#ifndef NVB_DEV_MIN
    if (e->ph == 2 && e->pw == 2) return launch_k3ms_t<false, 2, 2>(e, s, sa, smem);
    if (e->ph == 1 && e->pw == 1) return launch_k3ms_t<false, 1, 1>(e, s, sa, smem);
#endif
    return launch_k3ms_t<false, 0, 0>(e, s, sa, smem);
}

static int launch_k31(nvb_engine *e, const StepArgs &s)
{
    const SamplerArgs sa = agent_sampler_args(e);
    const int nplanes = sa.need_hs ? 3 : 1;
    const size_t smem = e->R > 0 ? nvb_sampler_smem(e->BW, e->BH, nplanes, sa.A)
                                 : nvb_sampler_smem(0, 0, 0, sa.A);
#ifndef NVB_DEV_MIN
    if (sa.need_hs) return launch_k31_t<true, 0, 0>(e, s, sa, smem);
#endif
    if (e->ph == 4 && e->pw == 2) return launch_k31_t<false, 4, 2>(e, s, sa, smem);
#ifndef NVB_DEV_MIN
    if (e->ph == 2 && e->pw == 2) return launch_k31_t<false, 2, 2>(e, s, sa, smem);
    if (e->ph == 1 && e->pw == 1) return launch_k31_t<false, 1, 1>(e, s, sa, smem);
#endif
    return launch_k31_t<false, 0, 0>(e, s, sa, smem);
}

static int phase2(nvb_engine *e, const StepArgs &s)
{
    // (view shards over NVLink: decide's prologue MINs the keys over the ranks, nvb_p2p_min_agent)
    k3_decide<<<e->B, NVB_STEP_THREADS, 0, e->stream>>>(s);
    k3_ties<<<e->sm_count * 4, NVB_TIE_THREADS, 0, e->stream>>>(s);
    e->launches += 2;
    CK(cudaGetLastError());
    return NVB_OK;
}

// update_error over a long training path: with the two-level bounding circles the agent's own
// CTA prunes it (step.cuh); the grid-wide three-kernel scan remains for the cases the prefilter
// does not cover (circles switched off, or a coverage threshold above max_distance).
static bool long_path_split(const nvb_engine *e)
{
    static const bool no_blocks = getenv("NAVSIM_B200_NO_PATH_BLOCKS") != nullptr;
    const bool one_pass = e->cvf * e->step_size <= e->max_dist;
    return e->n_path > NVB_PATH_SPLIT && (no_blocks || !one_pass);
}

static int phase3(nvb_engine *e, const StepArgs &s)
{
    // (view shards over NVLink: the move's prologue MINs the exact differences over the ranks)
    if (long_path_split(e) && !s.fake) {
        // long training path: pose update | grid-wide distance scan | bookkeeping
        const int chunks = (e->n_path + NVB_PATH_CHUNK - 1) / NVB_PATH_CHUNK;
        CK(launch_seq(k3_move_pose, dim3(e->B), dim3(NVB_STEP_THREADS), 0, e->stream, s));
        CK(launch_seq(k3_path_scan, dim3(chunks, e->B), dim3(256), 0, e->stream, s));
        CK(launch_seq(k3_move_finish, dim3(e->B), dim3(NVB_STEP_THREADS), 0, e->stream, s));
        e->launches += 3;
        return NVB_OK;
    }
    k3_move<<<e->B, NVB_STEP_THREADS, 0, e->stream>>>(s);
    e->launches++;
    CK(cudaGetLastError());
    return NVB_OK;
}

static int check_step_ready(nvb_engine *e, int fake)
{
    if (e->B <= 0) return fail(NVB_E_INVALID, "no agents set");
    if (!fake && e->n_path <= 0) return fail(NVB_E_INVALID, "no training path bound");
    return NVB_OK;
}

// K2 with optional event timing around it
static int launch_distance_timed(nvb_engine *e, int G)
{
    if (!e->timing) return launch_distance(e, G, true, true);
    if (e->ev_used + 2 > e->ev_pool.size()) {
        for (int i = 0; i < 64; i++) {
            cudaEvent_t ev;
            CK(cudaEventCreate(&ev));
            e->ev_pool.push_back(ev);
        }
    }
    CK(cudaEventRecord(e->ev_pool[e->ev_used], e->stream));
    int rc = launch_distance(e, G, true, true);
    CK(cudaEventRecord(e->ev_pool[e->ev_used + 1], e->stream));
    e->ev_used += 2;
    return rc;
}

// Small un-sharded libraries: decide + ties + move in ONE launch (every agent's CTA
// scans the library for its own tied headings); otherwise the three-launch form
// with the grid-wide tie pass.
#define NVB_FUSED_STEP_MAX_VIEWS 65536

static bool fused_step(const nvb_engine *e)
{
    static const bool off = getenv("NAVSIM_B200_NO_FUSED_STEP") != nullptr;
    // forms 1 and 2 rescan the library inside the agent's own CTA for ties: small un-sharded
    // libraries only.  Forms 3+ (grid-wide tie pass) take any library, sharded ones included
    // (the exchanges sit in the prologues of decide and of move+sample), unless update_error
    // needs the grid-wide path scan.
    const bool small = e->view_offset == 0 && e->n_total == e->N && e->N <= NVB_FUSED_STEP_MAX_VIEWS;
    return (small || (step_form() >= 3 && !long_path_split(e))) && sampler_slices(e) == 1 && !off;
}

// One step-batch.  Small un-sharded libraries (fused form): K2, then ONE launch that
// decides, moves and already samples the next step's glimpses (sample_next) -- two
// launches per step in steady state.  Otherwise: K1, K2, decide, ties, move.
static int one_step(nvb_engine *e, const StepArgs &s, bool sample_next = true)
{
    int rc;
    // K2 leaves tile minima instead of packed keys only when k3_step_tm is what reads them
    e->want_tmin = fused_step(e) && split_step() && tm_form(e) && e->d_tmin != nullptr;
    rc = phase1(e);
    const bool tm = e->want_tmin;
    e->want_tmin = false;
    if (rc) return rc;
    if (fused_step(e)) {
        if (sample_next) {
            // measured on the 1024-agent workload: one fused launch (69 us/step) beats
            // [decide + cooperative ties] + [move + sample] (79 us/step); the split form
            // stays available as a tuning knob
            if ((rc = split_step() ? launch_k3_split(e, s, tm) : launch_k31(e, s))) return rc;
            e->glimpses_pending = true;
            return NVB_OK;
        }
        if (tm) {   // the single-launch step without its gather: nothing sampled ahead
            StepArgs s2 = s;
            s2.no_sample = 1;
            return launch_k3_split(e, s2, true);
        }
        if (step_form() >= 3) {   // decide | grid-wide tie pass | move, nothing sampled ahead
            CK(launch_seq(k3_decide, dim3(e->B), dim3(NVB_STEP_THREADS), 0, e->stream, s));
            CK(launch_ties(e, s));
            CK(launch_seq(k3_move, dim3(e->B), dim3(NVB_STEP_THREADS), 0, e->stream, s));
            e->launches += 3;
            return NVB_OK;
        }
        k3_step<<<e->B, NVB_STEP_THREADS, 0, e->stream>>>(s);
        e->launches++;
        CK(cudaGetLastError());
        return NVB_OK;
    }
    if ((rc = phase2(e, s))) return rc;
    return phase3(e, s);
}

// kernels in one host-driven step-batch (sample | distance | step kernels, nothing sampled ahead)
static int io_step_launches(const nvb_engine *e)
{
    if (fused_step(e) && split_step() && tm_form(e) && e->d_tmin != nullptr) return 3;   // k1_sample | k2_tc* | k3_step_tm
    return fused_step(e) && effective_form(e) < 3 ? 3 : 5;
}

#define NVB_GRAPH_UNROLL 8

static bool graph_valid(const nvb_engine *e, int fake, int log_afam)
{
    return e->graph_exec && !e->graph_dirty && e->graph_fake == fake && e->graph_afam == log_afam &&
           e->graph_log_cap == e->log_cap && e->graph_B == e->B && e->graph_log_ptr == e->log_best;
}

// Captures one step-batch into a CUDA graph (the launch-bound inner loop) and
// replays it; the device-side step counter makes every replay log to its own slot.
static int ensure_graph(nvb_engine *e, int fake, int log_afam)
{
    if (graph_valid(e, fake, log_afam)) return NVB_OK;
    if (e->graph_exec) { cudaGraphExecDestroy(e->graph_exec); e->graph_exec = nullptr; }
    if (e->graph_multi) { cudaGraphExecDestroy(e->graph_multi); e->graph_multi = nullptr; }
    if (e->graph_io) { cudaGraphExecDestroy(e->graph_io); e->graph_io = nullptr; }   // shares graph_dirty
    if (e->graph_zc) { cudaGraphExecDestroy(e->graph_zc); e->graph_zc = nullptr; }
    const StepArgs s = make_step_args(e, fake, log_afam);
    // warm every kernel's lazy attribute setup outside the capture
    const int64_t before = e->launches;
    cudaGraph_t graph = nullptr;
    CK(cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal));
    int rc = one_step(e, s);
    cudaError_t ce = cudaStreamEndCapture(e->stream, &graph);
    e->launches = before;
    if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (ce != cudaSuccess) return fail(NVB_E_CUDA, "graph capture failed: %s", cudaGetErrorString(ce));
    ce = cudaGraphInstantiate(&e->graph_exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ce != cudaSuccess) return fail(NVB_E_CUDA, "graph instantiate failed: %s", cudaGetErrorString(ce));
    e->graph_fake = fake; e->graph_afam = log_afam; e->graph_log_cap = e->log_cap; e->graph_B = e->B;
    e->graph_log_ptr = e->log_best;
    e->graph_dirty = false;
    // Several step-batches in one graph: between graph launches the next kernel cannot become
    // resident before the previous graph has drained, inside a graph the programmatic edges let
    // the distance kernel's prologue (barriers, TMEM, first library tiles) overlap the step
    // kernel's tail.  Steady-state steps are identical (device-side step counter), so the same
    // capture repeated NVB_GRAPH_UNROLL times is NVB_GRAPH_UNROLL steps.  Not fatal if it fails.
    static const bool no_multi = getenv("NAVSIM_B200_NO_MULTI_GRAPH") != nullptr;
    if (fused_step(e) && !no_multi) {
        graph = nullptr;
        const bool pending = e->glimpses_pending;
        if (cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
            for (int i = 0; i < NVB_GRAPH_UNROLL && rc == NVB_OK; i++) rc = one_step(e, s);
            ce = cudaStreamEndCapture(e->stream, &graph);
            e->launches = before;
            e->glimpses_pending = pending;
            if (rc == NVB_OK && ce == cudaSuccess && cudaGraphInstantiate(&e->graph_multi, graph, 0) != cudaSuccess)
                e->graph_multi = nullptr;
            if (graph) cudaGraphDestroy(graph);
        }
        cudaGetLastError();
    }
    return NVB_OK;
}


// eager = plain launches, last step without sampling ahead (host-driven per-call use)
static int run_steps(nvb_engine *e, int nsteps, int fake, int log_afam, bool eager = false)
{
    int rc;
    if ((rc = ensure_log(e, e->steps_done + nsteps, log_afam != 0))) return rc;
    if (e->use_graph && !e->timing && !eager) {
        int done = 0;
        if (!graph_valid(e, fake, log_afam)) {
            // first use of a configuration: one step with plain launches, so that every
            // cudaFuncSetAttribute / occupancy query happens outside stream capture; in
            // the fused form it also leaves the next step's glimpses sampled, which is the
            // steady state the graph ({K2, step+sample}) is captured in
            const StepArgs s = make_step_args(e, fake, log_afam);
            if ((rc = one_step(e, s))) return rc;
            done = 1;
            if ((rc = ensure_graph(e, fake, log_afam))) return rc;
        }
        if (done < nsteps && fused_step(e) && !e->glimpses_pending) {
            if ((rc = launch_sampler(e, agent_sampler_args(e), e->B))) return rc;
            e->glimpses_pending = true;
        }
        // launches per step-batch: K2, decide, ties, move+sample | K1, K2, decide, ties, move
        // (+2 for the long-path move, +2 for the NVLink exchanges)
        const int ef = effective_form(e);
        const int per_step = fused_step(e) ? ((tm_form(e) && e->d_tmin && split_step()) ? 2 : ef == 3 ? 4 : (ef == 2 || ef == 4) ? 3 : 2)
                                           : 5 + (long_path_split(e) && !fake ? 2 : 0);
        for (int i = done; i < nsteps;) {
            if (e->graph_multi && nsteps - i >= NVB_GRAPH_UNROLL) {
                CK(cudaGraphLaunch(e->graph_multi, e->stream));
                e->launches += per_step * NVB_GRAPH_UNROLL;
                i += NVB_GRAPH_UNROLL;
            } else {
                CK(cudaGraphLaunch(e->graph_exec, e->stream));
                e->launches += per_step;
                i += 1;
            }
        }
    } else {
        const StepArgs s = make_step_args(e, fake, log_afam);
        for (int i = 0; i < nsteps; i++)
            if ((rc = one_step(e, s, !(eager && i == nsteps - 1)))) return rc;
    }
    e->steps_done += nsteps;
    return NVB_OK;
}

extern "C" int nvb_agents_step(nvb_engine *e, int nsteps, int fake, int log_afam)
{
    int rc = check_step_ready(e, fake);
    if (rc) return rc;
    if (nsteps <= 0) return NVB_OK;
    CK(cudaSetDevice(e->device));
    return run_steps(e, nsteps, fake, log_afam);
}

extern "C" int nvb_agents_rewind(nvb_engine *e)
{
    if (e->B <= 0) return fail(NVB_E_INVALID, "no agents set");
    CK(cudaSetDevice(e->device));
    const size_t B = e->B;
    CK(cudaMemcpyAsync(e->ag.poses, e->d_poses0, sizeof(double) * 3 * B, cudaMemcpyDeviceToDevice, e->stream));
    CK(cudaMemcpyAsync(e->ag.budget, e->d_budget0, sizeof(int32_t) * B, cudaMemcpyDeviceToDevice, e->stream));
    CK(cudaMemsetAsync(e->ag.status, 0, sizeof(int32_t) * B, e->stream));
    CK(cudaMemsetAsync(e->ag.completed, 0, sizeof(int32_t) * B, e->stream));
    CK(cudaMemsetAsync(e->ag.nav_frames, 0, sizeof(int32_t) * B, e->stream));
    CK(cudaMemsetAsync(e->ag.err_sum, 0, sizeof(double) * B, e->stream));
    CK(cudaMemsetAsync(e->ag.err_n, 0, sizeof(int32_t) * B, e->stream));
    CK(cudaMemsetAsync(e->ag.coverage, 0, B * (e->n_path > 0 ? e->n_path : 1), e->stream));
    CK(cudaMemsetAsync(e->d_step, 0xFF, sizeof(int), e->stream));
    CK(cudaMemsetAsync(e->d_pending, 0, sizeof(int32_t) * B, e->stream));
    CK(cudaMemsetAsync(e->d_tie_ready, 0, sizeof(int) * (size_t)B * e->A, e->stream));   // epochs restart
    e->glimpses_pending = false;
    e->steps_done = 0;
    return NVB_OK;
}

// Device-mapped view of a page-locked host buffer, or nullptr (pageable memory, no mapping).
static void *mapped_view(const void *host)
{
    if (!host) return nullptr;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, host) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (at.type != cudaMemoryTypeHost) return nullptr;
    void *dev = nullptr;
    if (cudaHostGetDevicePointer(&dev, const_cast<void *>(host), 0) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return dev;
}

// The buffers behind the same addresses may have been freed and re-allocated as pageable memory
// since the graph was captured: the mapping is verified on every call (a sub-microsecond query).
static bool still_mapped(const void *const key[4], void *const dev[4])
{
    for (int i = 0; i < 4; i++)
        if (key[i] && mapped_view(key[i]) != dev[i]) return false;
    return true;
}

// Captures the per-call step-batch with the caller's buffers bound into the kernels (see graph_zc).
// Failure is not fatal: the copy-based form keeps working.
static int capture_zero_copy_graph(nvb_engine *e, const void *const key[4])
{
    if (e->graph_zc) { cudaGraphExecDestroy(e->graph_zc); e->graph_zc = nullptr; }
    void *v[4] = {mapped_view(key[0]), mapped_view(key[1]), mapped_view(key[2]), mapped_view(key[3])};
    for (int i = 0; i < 4; i++)
        if (key[i] && !v[i]) return NVB_OK;
    e->zc_in = (const double *)v[0]; e->zc_best = (int16_t *)v[1]; e->zc_pose = (double *)v[2]; e->zc_fam = (double *)v[3];
    const StepArgs s = make_step_args(e, 0, 0);
    const int64_t before = e->launches;
    const bool pending = e->glimpses_pending;
    cudaGraph_t graph = nullptr;
    cudaError_t ce = cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal);
    int rc = NVB_OK;
    if (ce == cudaSuccess) {
        e->glimpses_pending = false;
        rc = one_step(e, s, false);
        ce = cudaStreamEndCapture(e->stream, &graph);
    }
    e->launches = before;
    e->glimpses_pending = pending;
    e->zc_in = nullptr; e->zc_best = nullptr; e->zc_pose = nullptr; e->zc_fam = nullptr;
    if (rc || ce != cudaSuccess) {
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        return NVB_OK;
    }
    ce = cudaGraphInstantiate(&e->graph_zc, graph, 0);
    cudaGraphDestroy(graph);
    if (ce != cudaSuccess) { e->graph_zc = nullptr; cudaGetLastError(); return NVB_OK; }
    memcpy(e->zc_key, key, sizeof e->zc_key);
    memcpy(e->zc_dev, v, sizeof e->zc_dev);
    return NVB_OK;
}

extern "C" int nvb_agents_step_io(nvb_engine *e, const double *poses_in, int nsteps, int16_t *best_idx,
                                  double *poses_out, double *step_fam)
{
    int rc = check_step_ready(e, 0);
    if (rc) return rc;
    if (nsteps <= 0) return fail(NVB_E_INVALID, "nsteps must be positive");
    CK(cudaSetDevice(e->device));
    const size_t B = e->B;
    const void *key[4] = {poses_in, best_idx, poses_out, step_fam};
    const bool per_call = poses_in != nullptr && nsteps == 1 && e->use_graph && !e->timing;
    if (per_call && e->graph_zc && !e->graph_dirty && e->graph_io_log_ptr == e->log_best &&
        e->graph_io_log_cap == e->log_cap && e->steps_done + 1 <= e->log_cap &&
        memcmp(key, e->zc_key, sizeof key) == 0 && still_mapped(key, e->zc_dev)) {
        // one graph launch, no copy operations: K1 reads the poses from the caller's buffer,
        // the move writes the results into the caller's buffers
        CK(cudaGraphLaunch(e->graph_zc, e->stream));
        e->glimpses_pending = false;
        e->launches += io_step_launches(e);
        e->steps_done += 1;
        CK(cudaStreamSynchronize(e->stream));
        return NVB_OK;
    }
    if (poses_in) {
        CK(cudaMemcpyAsync(e->ag.poses, poses_in, sizeof(double) * 3 * B, cudaMemcpyHostToDevice, e->stream));
        CK(cudaMemsetAsync(e->d_pending, 0, sizeof(int32_t) * B, e->stream));
        e->glimpses_pending = false;   // sampled for the old poses
    }
    if (poses_in != nullptr && nsteps == 1 && e->use_graph && !e->timing) {
        // host-driven per-call form: the five launches of one step-batch replayed as a graph
        if ((rc = ensure_log(e, e->steps_done + 1, false))) return rc;
        const bool valid = e->graph_io && !e->graph_dirty && e->graph_io_log_ptr == e->log_best &&
                           e->graph_io_log_cap == e->log_cap;
        if (!valid) {
            const StepArgs s = make_step_args(e, 0, 0);
            if ((rc = one_step(e, s, false))) return rc;   // plain launches first: lazy attribute setup
            e->steps_done += 1;
            if (e->graph_io) { cudaGraphExecDestroy(e->graph_io); e->graph_io = nullptr; }
            if (e->graph_zc) { cudaGraphExecDestroy(e->graph_zc); e->graph_zc = nullptr; }   // same baked pointers
            if (e->graph_exec) { cudaGraphExecDestroy(e->graph_exec); e->graph_exec = nullptr; }
            if (e->graph_multi) { cudaGraphExecDestroy(e->graph_multi); e->graph_multi = nullptr; }
            const int64_t before = e->launches;
            cudaGraph_t graph = nullptr;
            CK(cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal));
            e->glimpses_pending = false;
            rc = one_step(e, s, false);
            cudaError_t ce = cudaStreamEndCapture(e->stream, &graph);
            e->launches = before;
            if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
            if (ce != cudaSuccess) return fail(NVB_E_CUDA, "graph capture failed: %s", cudaGetErrorString(ce));
            ce = cudaGraphInstantiate(&e->graph_io, graph, 0);
            cudaGraphDestroy(graph);
            if (ce != cudaSuccess) return fail(NVB_E_CUDA, "graph instantiate failed: %s", cudaGetErrorString(ce));
            e->graph_io_log_ptr = e->log_best;
            e->graph_io_log_cap = e->log_cap;
            e->graph_dirty = false;
        } else {
            CK(cudaGraphLaunch(e->graph_io, e->stream));
            e->launches += io_step_launches(e);
            e->steps_done += 1;
        }
    } else if ((rc = run_steps(e, nsteps, 0, 0, poses_in != nullptr))) {
        return rc;
    }
    const size_t t = (size_t)e->steps_done - 1;
    if (best_idx)
        CK(cudaMemcpyAsync(best_idx, e->log_best + t * B, sizeof(int16_t) * B, cudaMemcpyDeviceToHost, e->stream));
    if (poses_out)
        CK(cudaMemcpyAsync(poses_out, e->ag.poses, sizeof(double) * 3 * B, cudaMemcpyDeviceToHost, e->stream));
    if (step_fam)
        CK(cudaMemcpyAsync(step_fam, e->log_sfam + t * B, sizeof(double) * B, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    // a caller that hands in the same buffers twice in a row gets the zero-copy graph next time
    if (per_call && e->graph_io && !e->graph_dirty) {
        static const bool off = getenv("NAVSIM_B200_NO_ZERO_COPY") != nullptr;
        const bool again = memcmp(key, e->seen_key, sizeof key) == 0;
        memcpy(e->seen_key, key, sizeof key);
        const bool have = e->graph_zc && memcmp(key, e->zc_key, sizeof key) == 0;
        if (again && !have && !off) return capture_zero_copy_graph(e, key);
    }
    return NVB_OK;
}

extern "C" int nvb_set_options(nvb_engine *e, int use_graph, int kernel_timing)
{
    CK(cudaSetDevice(e->device));
    CK(cudaStreamSynchronize(e->stream));
    e->use_graph = use_graph != 0;
    e->timing = kernel_timing != 0;
    e->ev_used = 0;
    e->k2_ms = 0.0;
    e->k2_count = 0;
    return NVB_OK;
}

extern "C" double nvb_kernel_time_ms(nvb_engine *e, int64_t *count)
{
    if (cudaSetDevice(e->device) != cudaSuccess) return -1.0;
    cudaStreamSynchronize(e->stream);
    for (size_t i = 0; i + 1 < e->ev_used; i += 2) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, e->ev_pool[i], e->ev_pool[i + 1]) == cudaSuccess) {
            e->k2_ms += ms;
            e->k2_count += 1;
        }
    }
    e->ev_used = 0;
    if (e->ev3_used) {
        double t[3] = {0, 0, 0};
        for (size_t i = 0; i + 3 < e->ev3_used + 1 && i + 3 < e->ev3.size(); i += 4)
            for (int k = 0; k < 3; k++) {
                float ms = 0;
                if (cudaEventElapsedTime(&ms, e->ev3[i + k], e->ev3[i + k + 1]) == cudaSuccess) t[k] += ms;
            }
        const double n = (double)(e->ev3_used / 4);
        fprintf(stderr, "[navsim_b200] per step: K2 %.2f us, decide %.2f us, ties %.2f us, move+sample %.2f us\n",
                e->k2_count ? e->k2_ms / e->k2_count * 1e3 : 0.0, t[0] / n * 1e3, t[1] / n * 1e3, t[2] / n * 1e3);
        e->ev3_used = 0;
    }
    if (count) *count = e->k2_count;
    return e->k2_ms;
}

extern "C" int nvb_agents_phase(nvb_engine *e, int phase, int fake, int log_afam)
{
    int rc = check_step_ready(e, fake);
    if (rc) return rc;
    CK(cudaSetDevice(e->device));
    if (phase == 1) {
        if ((rc = ensure_log(e, e->steps_done + 1, log_afam != 0))) return rc;
        return phase1(e);
    }
    const StepArgs s = make_step_args(e, fake, log_afam);
    if (phase == 2) return phase2(e, s);
    if (phase == 3) {
        rc = phase3(e, s);
        if (!rc) e->steps_done += 1;
        return rc;
    }
    return fail(NVB_E_INVALID, "phase must be 1, 2 or 3");
}

extern "C" int nvb_agents_steps_done(nvb_engine *e) { return e->steps_done; }

extern "C" int nvb_agents_get(nvb_engine *e, double *poses, int32_t *status, int32_t *completed,
                              int32_t *nav_frames, double *err_sum, int32_t *err_n, uint8_t *coverage)
{
    if (e->B <= 0) return fail(NVB_E_INVALID, "no agents set");
    CK(cudaSetDevice(e->device));
    const size_t B = e->B;
    if (poses) CK(cudaMemcpyAsync(poses, e->ag.poses, sizeof(double) * 3 * B, cudaMemcpyDeviceToHost, e->stream));
    if (status) CK(cudaMemcpyAsync(status, e->ag.status, sizeof(int32_t) * B, cudaMemcpyDeviceToHost, e->stream));
    if (completed) CK(cudaMemcpyAsync(completed, e->ag.completed, sizeof(int32_t) * B, cudaMemcpyDeviceToHost, e->stream));
    if (nav_frames) CK(cudaMemcpyAsync(nav_frames, e->ag.nav_frames, sizeof(int32_t) * B, cudaMemcpyDeviceToHost, e->stream));
    if (err_sum) CK(cudaMemcpyAsync(err_sum, e->ag.err_sum, sizeof(double) * B, cudaMemcpyDeviceToHost, e->stream));
    if (err_n) CK(cudaMemcpyAsync(err_n, e->ag.err_n, sizeof(int32_t) * B, cudaMemcpyDeviceToHost, e->stream));
    if (coverage && e->n_path > 0)
        CK(cudaMemcpyAsync(coverage, e->ag.coverage, B * e->n_path, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return NVB_OK;
}

extern "C" int nvb_agents_log(nvb_engine *e, int step0, int nsteps, int16_t *best_idx, double *poses,
                              double *step_fam, double *afam)
{
    if (e->B <= 0) return fail(NVB_E_INVALID, "no agents set");
    if (step0 < 0 || nsteps < 0 || step0 + nsteps > e->steps_done)
        return fail(NVB_E_INVALID, "log range [%d, %d) outside the %d logged steps", step0,
                    step0 + nsteps, e->steps_done);
    if (afam && !e->log_afam) return fail(NVB_E_INVALID, "angle_familiarity was not logged");
    CK(cudaSetDevice(e->device));
    const size_t B = e->B, o = (size_t)step0, k = (size_t)nsteps;
    if (k == 0) return NVB_OK;
    if (best_idx)
        CK(cudaMemcpyAsync(best_idx, e->log_best + o * B, sizeof(int16_t) * B * k, cudaMemcpyDeviceToHost, e->stream));
    if (poses)
        CK(cudaMemcpyAsync(poses, e->log_pose + o * B * 3, sizeof(double) * 3 * B * k, cudaMemcpyDeviceToHost, e->stream));
    if (step_fam)
        CK(cudaMemcpyAsync(step_fam, e->log_sfam + o * B, sizeof(double) * B * k, cudaMemcpyDeviceToHost, e->stream));
    if (afam)
        CK(cudaMemcpyAsync(afam, e->log_afam + o * B * e->A, sizeof(double) * B * e->A * k, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return NVB_OK;
}

extern "C" void *nvb_device_ptr(nvb_engine *e, int which)
{
    switch (which) {
    case NVB_PTR_KEYS: return e->d_keys;
    case NVB_PTR_TIE: return e->d_exact;
    case NVB_PTR_POSES: return e->ag.poses;
    }
    return nullptr;
}

// ---------------------------------------------------------------------------
// Tuning aid: runs ONE step-batch with plain launches and returns, per agent, the
// clock64 checkpoints of the fused step+sample kernel (start, active, decided,
// moved, window requested, window landed, sampled; SM cycles).
extern "C" int nvb_debug_step_clocks(nvb_engine *e, long long *out)
{
    int rc = check_step_ready(e, 0);
    if (rc) return rc;
    CK(cudaSetDevice(e->device));
    CK(cudaMalloc(&e->d_dbg, sizeof(long long) * 16 * e->B));
    CK(cudaMemsetAsync(e->d_dbg, 0, sizeof(long long) * 16 * e->B, e->stream));
    rc = run_steps(e, 1, 0, 0, false);
    if (!rc) {
        const bool g = e->use_graph;
        e->use_graph = false;
        rc = run_steps(e, 1, 0, 0, false);
        e->use_graph = g;
    }
    if (!rc) CK(cudaMemcpyAsync(out, e->d_dbg, sizeof(long long) * 16 * e->B, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    cudaFree(e->d_dbg);
    e->d_dbg = nullptr;
    e->graph_dirty = true;
    return rc;
}

// Tuning aid: `nsteps` step-batches the way nvb_agents_step runs them (graph replay), with
// every CTA of the four step kernels stamping %globaltimer when it becomes resident, when
// its grid dependency is met and when it is done; out [4][NVB_TL_CTAS][3] holds the stamps
// of the LAST step-batch (ns; 0 = CTA not present).
extern "C" int nvb_debug_timeline(nvb_engine *e, int nsteps, long long *out)
{
    int rc = check_step_ready(e, 0);
    if (rc) return rc;
    CK(cudaSetDevice(e->device));
    const size_t n = (size_t)6 * NVB_TL_CTAS * 3;   // kernels 0..3 + two slots of k3_step_tm checkpoints
    CK(cudaMalloc(&e->d_tl, sizeof(long long) * n));
    CK(cudaMemsetAsync(e->d_tl, 0, sizeof(long long) * n, e->stream));
    e->graph_dirty = true;
    rc = run_steps(e, nsteps, 0, 0, false);
    if (!rc) CK(cudaMemcpyAsync(out, e->d_tl, sizeof(long long) * n, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    cudaFree(e->d_tl);
    e->d_tl = nullptr;
    e->graph_dirty = true;
    return rc;
}

// ---- view shards over NVLink peer memory -----------------------------------------
extern "C" int nvb_p2p_export(nvb_engine *e, void *handle64)
{
    if (e->B <= 0) return fail(NVB_E_INVALID, "set the agents before exporting the exchange area");
    CK(cudaSetDevice(e->device));
    CK(cudaStreamSynchronize(e->stream));
    const long long cap = (long long)e->B * e->A;
    // flags [2][ranks][B] + data [2][ranks][cap], sized for the largest world (step.cuh)
    const size_t words = (size_t)2 * NVB_P2P_MAX_RANKS * ((size_t)e->B + (size_t)cap);
    e->p2p_on = false;
    e->graph_dirty = true;
    for (int i = 0; i < NVB_P2P_MAX_RANKS; i++)
        if (e->p2p_opened[i]) { cudaIpcCloseMemHandle(e->p2p_opened[i]); e->p2p_opened[i] = nullptr; }
    free_dev(e->d_xarea);
    e->d_xarea = nullptr;
    CK(cudaMalloc((void **)&e->d_xarea, words * sizeof(unsigned long long)));
    CK(cudaMemset(e->d_xarea, 0, words * sizeof(unsigned long long)));
    int rc;
    if ((rc = alloc_dev(&e->d_p2p_seq, 1))) return rc;
    if ((rc = alloc_dev(&e->d_p2p_err, 1))) return rc;
    CK(cudaMemset(e->d_p2p_seq, 0, sizeof(unsigned long long)));
    CK(cudaMemset(e->d_p2p_err, 0, sizeof(int)));
    e->p2p.cap = cap;
    e->p2p.B = e->B;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    CK(cudaIpcGetMemHandle((cudaIpcMemHandle_t *)handle64, e->d_xarea));
    return NVB_OK;
}

extern "C" int nvb_p2p_attach(nvb_engine *e, int rank, int world, const void *handles64)
{
    if (!e->d_xarea) return fail(NVB_E_INVALID, "call nvb_p2p_export first");
    if (world < 2 || world > NVB_P2P_MAX_RANKS || rank < 0 || rank >= world)
        return fail(NVB_E_INVALID, "bad rank/world");
    CK(cudaSetDevice(e->device));
    e->p2p.self = e->d_xarea;
    e->p2p.rank = rank;
    e->p2p.world = world;
    e->p2p.epoch = e->d_p2p_seq;
    e->p2p.error = e->d_p2p_err;
    e->p2p.spin_limit = 4000000000ll;   // ~2 s of SM clock: a peer that never shows up is an error, not a hang
    if (getenv("NAVSIM_B200_P2P_SPIN")) e->p2p.spin_limit = atoll(getenv("NAVSIM_B200_P2P_SPIN"));
    for (int p = 0; p < world; p++) {
        if (p == rank) { e->p2p.peer[p] = e->d_xarea; continue; }
        void *ptr = nullptr;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char *)handles64 + 64 * p, 64);
        CK(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        e->p2p_opened[p] = ptr;
        e->p2p.peer[p] = (unsigned long long *)ptr;
    }
    e->p2p_on = true;
    e->graph_dirty = true;
    return NVB_OK;
}

extern "C" int nvb_p2p_error(nvb_engine *e)
{
    if (!e->d_p2p_err) return 0;
    int v = 0;
    cudaSetDevice(e->device);
    cudaStreamSynchronize(e->stream);
    cudaMemcpy(&v, e->d_p2p_err, sizeof(int), cudaMemcpyDeviceToHost);
    return v;
}

extern "C" double nvb_probe_sad_peak(nvb_engine *e, int iters)
{
    if (cudaSetDevice(e->device) != cudaSuccess) return -1.0;
    uint32_t *d_sink = nullptr;
    if (cudaMalloc(&d_sink, 64) != cudaSuccess) return -1.0;
    const int grid = e->sm_count * 8, threads = 256;
    cudaEvent_t t0, t1;
    cudaEventCreate(&t0);
    cudaEventCreate(&t1);
    k_probe_sad<<<grid, threads, 0, e->stream>>>(iters, 12345u, d_sink);   // warm-up
    cudaEventRecord(t0, e->stream);
    k_probe_sad<<<grid, threads, 0, e->stream>>>(iters, 12345u, d_sink);
    cudaEventRecord(t1, e->stream);
    e->launches += 2;
    cudaStreamSynchronize(e->stream);
    float ms = 0;
    cudaEventElapsedTime(&ms, t0, t1);
    cudaEventDestroy(t0);
    cudaEventDestroy(t1);
    cudaFree(d_sink);
    if (ms <= 0) return -1.0;
    return (double)grid * threads * (double)iters * 32.0 * 4.0 / (ms * 1e-3);
}

__global__ void k_debug_sincos(const double *x, long long n, double *s, double *c)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) nvb_glibc_sincos(x[i], s + i, c + i);
}

// Test hook: the device's sin / cos (csrc/glibc_trig.cuh) of n host doubles.
extern "C" int nvb_debug_sincos(nvb_engine *e, const double *x, int64_t n, double *s, double *c)
{
    if (n <= 0) return NVB_OK;
    CK(cudaSetDevice(e->device));
    double *d = nullptr;
    CK(cudaMalloc(&d, sizeof(double) * 3 * (size_t)n));
    CK(cudaMemcpyAsync(d, x, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, e->stream));
    k_debug_sincos<<<(unsigned)((n + 255) / 256), 256, 0, e->stream>>>(d, n, d + n, d + 2 * n);
    e->launches++;
    CK(cudaMemcpyAsync(s, d + n, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaMemcpyAsync(c, d + 2 * n, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    cudaFree(d);
    return NVB_OK;
}

extern "C" double nvb_probe_mma_peak(nvb_engine *e, int iters)
{
    if (cudaSetDevice(e->device) != cudaSuccess || iters <= 0) return -1.0;
    if (cudaFuncSetAttribute(k_probe_umma, cudaFuncAttributeMaxDynamicSharedMemorySize, NVB_PROBE_UMMA_SMEM) != cudaSuccess)
        return -1.0;
    cudaEvent_t t0, t1;
    cudaEventCreate(&t0);
    cudaEventCreate(&t1);
    k_probe_umma<<<e->sm_count, 128, NVB_PROBE_UMMA_SMEM, e->stream>>>(iters);   // warm-up
    cudaEventRecord(t0, e->stream);
    k_probe_umma<<<e->sm_count, 128, NVB_PROBE_UMMA_SMEM, e->stream>>>(iters);
    cudaEventRecord(t1, e->stream);
    e->launches += 2;
    cudaStreamSynchronize(e->stream);
    float ms = 0;
    cudaEventElapsedTime(&ms, t0, t1);
    cudaEventDestroy(t0);
    cudaEventDestroy(t1);
    if (ms <= 0 || cudaGetLastError() != cudaSuccess) return -1.0;
    return (double)e->sm_count * (double)iters * 4.0 * (2.0 * 128.0 * 256.0 * 32.0) / (ms * 1e-3);
}

extern "C" int nvb_tc_planes(nvb_engine *e) { return e->tc_ok ? e->tc_planes.n_planes : 0; }

extern "C" double nvb_time_distance_kernel(nvb_engine *e, int reps)
{
    if (e->B <= 0 || e->N <= 0 || reps <= 0) return -1.0;
    if (cudaSetDevice(e->device) != cudaSuccess) return -1.0;
    const int G = e->B * e->A;
    cudaEvent_t t0, t1;
    cudaEventCreate(&t0);
    cudaEventCreate(&t1);
    fill_u64(e, e->d_keys, G, KEY_NONE);
    launch_distance(e, G, false, false);   // (re-)encodes the glimpse planes from the V plane when the tensor-core kernel runs
    cudaStreamSynchronize(e->stream);
    float total = 0;
    for (int r = 0; r < reps; r++) {
        fill_u64(e, e->d_keys, G, KEY_NONE);
        cudaEventRecord(t0, e->stream);
        launch_distance(e, G, false, true);
        cudaEventRecord(t1, e->stream);
        cudaStreamSynchronize(e->stream);
        float ms = 0;
        cudaEventElapsedTime(&ms, t0, t1);
        total += ms;
    }
    cudaEventDestroy(t0);
    cudaEventDestroy(t1);
    return total / reps;
}
