/*
 * navsim_b200.h -- C ABI of the B200-native scene-familiarity engine.
 *
 * This is the drop-in boundary for navsim's hot path: the entry points a
 * binding for the reference's `navsim.util` seam
 * (navsim/NavBySceneFamiliarity.py:20,
 *  `from navsim.util import sads_familiarity, downscale_chem, fill_sensor_from`)
 * and for the stepping loop around it (NavBySceneFamiliarity.py:279-329)
 * calls.  Plain C types only: host pointers + sizes in, status code out.  No
 * exceptions cross this boundary; nvb_last_error() returns the message of
 * the last failing call on the calling thread.
 *
 * Built by __graft_entry__.build() into
 *   navigation-by-deja-vu_b200/lib/libnavsim_b200.so   (sm_100a only)
 * and bound from Python with ctypes in
 *   navigation-by-deja-vu_b200/navsim/_cabi.py
 *
 * Unless a function says "device", every pointer is a HOST pointer and the
 * call is synchronous with respect to the host buffers it names.  Calls on one
 * engine are not thread-safe (the reference object is not re-entrant either:
 * it reuses _roundbuf/_landscape_glimpse_buf, NavBySceneFamiliarity.py:95-96).
 */
#ifndef NAVSIM_B200_H
#define NAVSIM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes ------------------------------------------------------ */
/* Per-agent stop codes mirror StopNavigationException.get_code()
 * (NavBySceneFamiliarity.py:30-49); 0 = still running / ran out of frames
 * (scripts/run_experiment.py:249). */
#define NVB_OK 0
#define NVB_REACHED_END 1
#define NVB_TOO_FAR (-1)
#define NVB_OUT_OF_BOUNDS (-2)
/* IndexError from fill_sensor_from's bounds-checked indexing
 * (util.pyx:165-168; not a StopNavigationException in the reference). */
#define NVB_INDEX_ERROR (-3)
/* API-level failures (never a per-agent status). */
#define NVB_E_INVALID (-100)  /* bad argument / call order */
#define NVB_E_CUDA (-101)     /* CUDA runtime or driver error */
#define NVB_E_NO_DEVICE (-102) /* no sm_100 device: there is no CPU fallback */

typedef struct nvb_engine nvb_engine;

/* ---- engine ------------------------------------------------------------ */
/* Creates an engine on CUDA device `device`.  `stream` is a cudaStream_t the
 * engine launches on (NULL: the engine creates its own non-blocking stream).
 * Fails with NVB_E_NO_DEVICE when no sm_100 GPU is present. */
int nvb_engine_create(int device, void *stream, nvb_engine **out);
void nvb_engine_destroy(nvb_engine *e);
const char *nvb_last_error(void);
const char *nvb_version(void);
/* Blocks until everything launched on the engine's stream has finished. */
int nvb_sync(nvb_engine *e);
/* The cudaStream_t the engine launches on (its own, or the one given at creation): a caller
 * that touches the engine's device buffers from another stream (the NCCL MIN all-reduce of
 * the view-sharded mode, navsim/sharded.py) orders the two streams with events on it. */
void *nvb_stream_handle(nvb_engine *e);

/* ---- world: landscape, sensor, heading sweep, navigation parameters ----- */
/* Landscape: uint8 HSV, shape (rows, cols, 3), arbitrary byte strides
 * (negative strides = flipped views, scripts/run_experiment.py:196-199).
 * Copied into a planar, pitch-aligned device layout.
 * Replaces the `landscape` memoryview argument of util.pyx:141. */
int nvb_set_landscape(nvb_engine *e, const uint8_t *hsv, int rows, int cols,
                      ptrdiff_t stride_row, ptrdiff_t stride_col, ptrdiff_t stride_chan);

/* Sensor: W x H sensor pixels of pw x ph landscape pixels
 * (NavBySceneFamiliarity.py:90-96), three 256-entry quantisation tables
 * (the float32 level rounding of :178-186 as a lookup, built by the host
 * binding with the same NumPy ops) and mask_middle_n (:189-190). */
int nvb_set_sensor(nvb_engine *e, int W, int H, int pw, int ph,
                   const uint8_t *lut3x256, int mask_middle_n);

/* Heading sweep: A offsets in radians (np.linspace of :86-88). */
int nvb_set_saccade(nvb_engine *e, int A, const double *angle_offsets);

/* step_size (:106), max_distance_to_training_path (:109), threshold_factor
 * (:81), coverage_threshold_factor (:82), chem_weight (util.pyx:10). */
int nvb_set_nav_params(nvb_engine *e, double step_size, double max_dist,
                       double threshold_factor, double coverage_factor, double chem_weight);

/* ---- glimpses: util.pyx fill_sensor_from / downscale_chem, get_sensor_mat - */
/* A1, util.pyx:137-168: nearest-neighbour rotated gather of an
 * (Hpx, Wpx, 3) patch.  cos_rot/sin_rot are cos/sin of -(pi/2 - angle),
 * evaluated by the caller with the host libm so the result is bit-identical
 * to the reference.  Returns NVB_OK or NVB_INDEX_ERROR. */
int nvb_fill_sensor(nvb_engine *e, uint8_t *sensor, int Hpx, int Wpx, double xpos, double ypos,
                    double cos_rot, double sin_rot);

/* A2, util.pyx:91-134: image (R, C, 3) contiguous -> out (R/fr, C/fc, 3). */
int nvb_downscale_chem(nvb_engine *e, const uint8_t *image, int R, int C, int fr, int fc,
                       uint8_t *out);

/* A4, NavBySceneFamiliarity.py:151-192 for G poses at once.
 * poses [G][3] = x, y, angle.  cs [G][2] = cos/sin of -(pi/2 - angle) from the
 * host libm, or NULL to evaluate on the device.  out [G][H][W][3];
 * status [G] = NVB_OK / NVB_OUT_OF_BOUNDS / NVB_INDEX_ERROR. */
int nvb_glimpse_batch(nvb_engine *e, const double *poses, const double *cs, int G,
                      uint8_t *out, int32_t *status);

/* ---- library: train_from_path, familiar_scenes -------------------------- */
/* A8, NavBySceneFamiliarity.py:118-140: one glimpse per path point at the
 * tangent angle.  path [N][2]; angles [N] (np.arctan2 of :124-126, last one
 * repeated :132) and cs [N][2] come from the host binding.  On a failing
 * point returns its status and stores the index in *bad_index.  Also binds
 * `path` as the training path of update_error (:252-276). */
int nvb_library_build(nvb_engine *e, const double *path, const double *angles, const double *cs,
                      int N, int *bad_index);
/* Binds an existing library: scenes [N][H][W][3] (what util.pyx:11-20 captures
 * at closure creation) and, optionally, its training path [N][2] (or NULL). */
int nvb_library_upload(nvb_engine *e, const uint8_t *scenes, const double *path, int N);
/* Binds the training path of update_error (NavBySceneFamiliarity.py:252-276)
 * on its own: path [n][2].  In the view-sharded mode every rank keeps the
 * whole path while holding only its slice of the views. */
int nvb_set_training_path(nvb_engine *e, const double *path, int n);
/* familiar_scenes back on the host: scenes [N][H][W][3]. */
int nvb_library_download(nvb_engine *e, uint8_t *scenes);
/* View-sharded library (one process per GPU): this engine holds views
 * [view_offset, view_offset + N) of a library of n_total views; keys carry
 * global view indices.  Call after the library is bound. */
int nvb_library_set_shard(nvb_engine *e, int64_t view_offset, int64_t n_total);

/* ---- distance: util.pyx sads_hsv_metric ---------------------------------- */
/* A5, util.pyx:28-73, exact FP64 operation order: G query scenes
 * [G][H][W][3] against every library view -> fam [G][N]
 * (fam = H*W - diff).  This is what the closure returned by
 * sads_familiarity(cw)(scenes) fills into `fambuf`. */
int nvb_familiarity(nvb_engine *e, const uint8_t *scenes_q, int G, double *fam);
/* The hot kernel on the same inputs: per query the minimum integer
 * difference over the library and the lowest view index attaining it
 * (chem_weight == 0: sum |dV|; otherwise the FP64 surrogate, view = -1). */
int nvb_familiarity_min(nvb_engine *e, const uint8_t *scenes_q, int G, double *min_diff,
                        int64_t *view_idx);

/* ---- resident stepping loop: step_forward + update_error ------------------ */
/* B agents: poses [B][3] = x, y, angle; frame_budget [B] or NULL (no limit).
 * Resets error accumulators, coverage and the step log (reset_error,
 * NavBySceneFamiliarity.py:195-207). */
int nvb_agents_set(nvb_engine *e, const double *poses, const int32_t *frame_budget, int B);
/* Advances every running agent by up to nsteps steps (A6 + A7,
 * NavBySceneFamiliarity.py:279-329) without host interaction.  Asynchronous:
 * returns after the launches are queued.  log_afam != 0 also logs
 * angle_familiarity per step. */
int nvb_agents_step(nvb_engine *e, int nsteps, int fake, int log_afam);
/* Restores the start poses, budgets, counters, coverage and step log of the
 * last nvb_agents_set() from a device-side snapshot (no host traffic). */
int nvb_agents_rewind(nvb_engine *e);
/* One call = host poses in (pinned or pageable; NULL keeps the device poses),
 * nsteps step-batches, the last step's results back on the host, one
 * synchronisation: the per-call form a host-driven loop uses (the reference's
 * step_forward() contract, NavBySceneFamiliarity.py:279, for B agents).
 * best_idx [B], poses_out [B][3], step_fam [B]; any may be NULL.
 * When the same page-locked buffers are passed call after call (nsteps == 1) the
 * engine binds them into the kernels of the step: the poses are read from, and the
 * results written to, the host buffers directly, with no copy operations. */
int nvb_agents_step_io(nvb_engine *e, const double *poses_in, int nsteps, int16_t *best_idx,
                       double *poses_out, double *step_fam);
/* Current state (synchronises).  Any pointer may be NULL.  poses [B][3],
 * status [B] (stop codes above), completed [B] (steps that returned normally,
 * scripts/run_experiment.py:243-245), nav_frames [B] (navigated_for_frames),
 * err_sum [B] (_navigation_error), err_n [B] (_n_navigation_error),
 * coverage [B][N] (_coverage_array). */
int nvb_agents_get(nvb_engine *e, double *poses, int32_t *status, int32_t *completed,
                   int32_t *nav_frames, double *err_sum, int32_t *err_n, uint8_t *coverage);
/* Step log since nvb_agents_set (synchronises): steps [step0, step0+nsteps).
 * best_idx [nsteps][B] (-1 where the agent did not step), poses
 * [nsteps][B][3] (after the step), step_fam [nsteps][B], afam
 * [nsteps][B][A] (NaN where not evaluated).  Any pointer may be NULL. */
int nvb_agents_log(nvb_engine *e, int step0, int nsteps, int16_t *best_idx, double *poses,
                   double *step_fam, double *afam);
/* Number of steps logged so far. */
int nvb_agents_steps_done(nvb_engine *e);

/* Phased stepping for a view-sharded library: phase 1 samples glimpses and
 * scores them against the local shard, leaving one packed 64-bit key per
 * (agent, heading) in the device buffer nvb_device_ptr(e, NVB_PTR_KEYS); the
 * caller MIN-all-reduces it across ranks; phase 2 picks headings and finds
 * ties, leaving exact tie differences in NVB_PTR_TIE (MIN-all-reduce again);
 * phase 3 moves the agents.  nvb_agents_step() runs the three back to back. */
int nvb_agents_phase(nvb_engine *e, int phase, int fake, int log_afam);

/* View shards over NVLink peer memory (no NCCL, no host round trip per step): with
 * the exchange attached, nvb_agents_step() runs the sharded sequence itself --
 * K1, K2, MIN exchange of the keys, decide, ties, MIN exchange of the exact
 * differences, move -- where the exchange is a kernel that publishes into the
 * peers' mapped memory, waits for their flags (bounded) and reduces with P2P loads.
 * Order: nvb_library_set_shard, nvb_agents_set, nvb_p2p_export (64-byte CUDA IPC
 * handle of this rank's exchange area), all-gather the handles out of band,
 * nvb_p2p_attach(rank, world, handles [world][64]).  nvb_p2p_error() != 0 after a
 * synchronisation means a peer did not arrive within ~2 s. */
int nvb_p2p_export(nvb_engine *e, void *handle64);
int nvb_p2p_attach(nvb_engine *e, int rank, int world, const void *handles64);
int nvb_p2p_error(nvb_engine *e);

#define NVB_PTR_KEYS 0 /* uint64 [B*A]  */
#define NVB_PTR_TIE 1  /* uint64 [B*A]  (FP64 bit patterns) */
#define NVB_PTR_POSES 2 /* double [B][3] */
void *nvb_device_ptr(nvb_engine *e, int which);

/* use_graph: replay one captured CUDA graph per step-batch (default on).
 * kernel_timing: record CUDA events around the distance kernel inside the step
 * sequence (disables graph replay); read back with nvb_kernel_time_ms(), which
 * returns the summed milliseconds and stores the launch count. */
int nvb_set_options(nvb_engine *e, int use_graph, int kernel_timing);
double nvb_kernel_time_ms(nvb_engine *e, int64_t *count);

/* Tuning aid: one step-batch with per-agent clock64 checkpoints of the fused
 * step+sample kernel; out [2][B][8] (SM cycles; 0 = not reached): block 0 the main
 * checkpoints, block 1 finer ones inside the move. */
int nvb_debug_step_clocks(nvb_engine *e, long long *out);

/* Tuning aid: `nsteps` step-batches as nvb_agents_step runs them, every CTA of the four
 * step kernels (distance, decide, ties, move+sample) stamping the global timer when it
 * becomes resident, when its grid dependency is met and when it is done.
 * out [6][2048][3], ns, of the last step-batch; 0 = CTA not present.  Slots 4 and 5 hold
 * per-CTA checkpoints of the single-launch step kernel (k3_step_tm): inputs in, pose known,
 * update_error done | all warps at the gather, window landed, -. */
int nvb_debug_timeline(nvb_engine *e, int nsteps, long long *out);

/* ---- landscape preparation on the device ------------------------------------------------
 * What make_nsf does per landscape / trial on the host (scripts/run_experiment.py:160-199,
 * navsim/util.pyx:76-88 set_HS_where_equal), on the landscape the engine holds:
 * label_grains: threshold V >= `threshold`, modal filter with a modal_w x modal_w footprint
 *   (odd; 0 or 1: none; skimage.filters.rank.modal on the 0/1 image), 8-connected components
 *   numbered in raster order of their first pixel (skimage.measure.label); returns the count.
 * grains_get:   per-grain pixel counts (regionprops area; equivalent_diameter = sqrt(4 area / pi))
 *   and / or the label image [rows][cols] int64 (0 = background); either pointer may be NULL.
 * paint:        H and S of every labelled pixel from per-grain tables (set_HS_where_equal).
 * flip:         landscape[::-1] / [:, ::-1] (run_experiment.py:196-199); labels are dropped.
 * download:     the landscape as [rows][cols][3] uint8 (tests).
 * Painting and flipping invalidate the library and the agents. */
int nvb_landscape_label_grains(nvb_engine *e, int threshold, int modal_w, int64_t *n_grains);
int nvb_landscape_grains_get(nvb_engine *e, int32_t *areas, int64_t *labels);
int nvb_landscape_paint(nvb_engine *e, const uint8_t *H, const uint8_t *S, int64_t n_grains);
int nvb_landscape_flip(nvb_engine *e, int flip_v, int flip_h);
int nvb_landscape_download(nvb_engine *e, uint8_t *hsv);

/* Offline landscape generation, navsim.util.diffuse (navsim/util.pyx:186-235): nstep explicit
 * time steps of the 2-D heat equation with periodic boundaries on a side x side float64 field,
 *   new = m + multiplier * (m[i+1] + m[i-1] - 4 m + m[j+1] + m[j-1]),
 * the reference's operation order without fused multiply-add (bit-identical).  The caller
 * computes `multiplier` exactly as util.pyx:203-206 does.  Host pointers; out may alias initial. */
int nvb_diffuse(nvb_engine *e, const double *initial, int64_t side, int64_t nstep, double multiplier, double *out);

/* Which distance kernel scores the glimpses.  mode 0 (default): the tensor-core kernel
 * (tcgen05 int8, exact thermometer form of the sum of absolute differences) whenever the V
 * quantisation has at most 9 levels, chem_weight is 0 and the batch has at least 96
 * glimpses, else the byte-SIMD kernel; mode 1: the byte-SIMD kernel everywhere.  Both give
 * the same integer minima and view indices. */
int nvb_set_distance_kernel(nvb_engine *e, int mode);
/* Which kernel scores the current agent batch: 1 = tensor cores (k2_tc / k2_tc_bs), 2 = the
 * library-streaming byte-SIMD kernel for a handful of glimpses (k2_stream), 0 = the tiled
 * byte-SIMD kernels (k2_sad_v, k2_sad_hsv_t). */
int nvb_distance_kernel(nvb_engine *e);

/* Test hook: sin and cos of n host doubles as the device-resident stepping loop computes
 * them (csrc/glibc_trig.cuh: the GNU C Library's algorithm, so that positions match the
 * reference's host trigonometry bit for bit, NavBySceneFamiliarity.py:319-320). */
int nvb_debug_sincos(nvb_engine *e, const double *x, int64_t n, double *s, double *c);

/* Counters for bench.py: kernels launched by this engine so far. */
int64_t nvb_launch_count(nvb_engine *e);
/* Integer byte-SIMD issue-rate probe (register-resident VABSDIFF4+accumulate
 * loop on every SM); returns pixel-compares per second, the denominator of
 * the distance kernel's ALU roofline. */
double nvb_probe_sad_peak(nvb_engine *e, int iters);
/* int8 tensor-core issue-rate probe (tcgen05.mma kind::i8, 128 x 256 x 32 tiles on operands
 * resident in shared memory, one CTA per SM); returns integer operations (2 per multiply-add)
 * per second: the denominator of the tensor-core distance kernel's roofline. */
double nvb_probe_mma_peak(nvb_engine *e, int iters);
/* Thermometer planes the V quantisation needs (0: the tensor-core kernel does not apply). */
int nvb_tc_planes(nvb_engine *e);
/* Device time of the distance kernel alone on the current agent batch
 * (CUDA events on the engine stream, `reps` launches); milliseconds/launch. */
double nvb_time_distance_kernel(nvb_engine *e, int reps);

#ifdef __cplusplus
}
#endif
#endif /* NAVSIM_B200_H */
