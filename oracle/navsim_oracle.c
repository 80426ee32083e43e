/*
 * navsim_oracle.c -- CPU restatement of navsim's scene-familiarity hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product
 * path: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library, and only as the checker (or as
 * the timed CPU baseline).  The product (navigation-by-deja-vu_b200/) never
 * links, imports or calls it.
 *
 * Parity status: the reference ships no tests and no golden vectors
 * (SURVEY.md section 4), so this restatement is pinned against the
 * reference's own code run in the build container: oracle/build_ref.py
 * compiles navsim/util.pyx and navsim/NavBySceneFamiliarity.py (where they
 * lie under /root/reference) into oracle/_ref/, tests/test_oracle_vs_ref.py
 * compares every function below with it, and tests/golden/ holds vectors
 * generated from that reference build (tests/golden/make_golden.py).
 *
 * Every function cites the reference file:line it follows (paths relative to
 * the reference root).  Plain C, IEEE double semantics, no FMA contraction:
 * build with  gcc -O2 -ffp-contract=off  (see oracle/build.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define NVO_OK 0
#define NVO_REACHED_END 1   /* NavBySceneFamiliarity.py:30-34  */
#define NVO_TOO_FAR (-1)    /* NavBySceneFamiliarity.py:39-43  */
#define NVO_OUT_OF_BOUNDS (-2) /* NavBySceneFamiliarity.py:45-49 */
#define NVO_INDEX_ERROR (-3)   /* IndexError out of util.pyx:165-168 (bounds-check on) */

typedef struct {
    /* landscape, uint8 HSV, element strides in bytes (may be negative: flips,
       scripts/run_experiment.py:196-199) */
    const uint8_t *land;
    long rows, cols;
    long s_row, s_col, s_chan;
    /* sensor: W x H sensor pixels, each pw x ph landscape pixels */
    long W, H, pw, ph;
    uint8_t lut[3][256];
    long mask_middle_n;
    /* heading sweep */
    long A;
    const double *offsets;
    /* navigation */
    double step_size, max_dist, threshold_factor, coverage_factor, chem_weight;
    /* library: scenes [N][H][W][3] contiguous; path [N][2] */
    long N;
    const uint8_t *scenes;
    const double *path;
} nvo_world;

typedef struct {
    double x, y, angle;
    long navigated_for_frames;
    double nav_err;
    long n_nav_err;
    uint8_t *coverage; /* [N] */
} nvo_agent;

/* Python / NumPy float modulo: fmod, then fold into the divisor's sign.
 * Used at NavBySceneFamiliarity.py:291,317 with a positive divisor. */
static double py_mod(double a, double b)
{
    double m = fmod(a, b);
    if (m != 0.0) {
        if ((b < 0) != (m < 0)) m += b;
    } else {
        m = copysign(0.0, b);
    }
    return m;
}

/* Cython indexing with boundscheck on, wraparound on (no decorators at
 * util.pyx:137): a negative index gets the dimension added once. */
static int wrap_index(long *i, long dim)
{
    if (*i < 0) *i += dim;
    return (*i < 0 || *i >= dim) ? -1 : 0;
}

/* A1: nearest-neighbour rotated gather.  navsim/util.pyx:137-168. */
int nvo_fill_sensor(uint8_t *sensor, long Hpx, long Wpx, double xpos, double ypos,
                    double angle, const uint8_t *land, long rows, long cols,
                    long s_row, long s_col, long s_chan)
{
    double rot = -(0.5 * M_PI - angle);      /* :143 */
    double c = cos(rot), s = sin(rot);       /* :144-145 */
    double half_w = 0.5 * (double)Wpx;       /* :150 */
    double half_h = 0.5 * (double)Hpx;       /* :151 */
    for (long i = 0; i < Hpx; i++) {
        for (long j = 0; j < Wpx; j++) {
            double px = (double)j - half_w;  /* :159 */
            double py = (double)i - half_h;  /* :160 */
            double rx = px * c - py * s;     /* :161 */
            double ry = px * s + py * c;     /* :162 */
            long iy = (long)round(ry + ypos); /* :166 */
            long ix = (long)round(rx + xpos); /* :167 */
            if (wrap_index(&iy, rows) || wrap_index(&ix, cols)) return NVO_INDEX_ERROR;
            const uint8_t *src = land + iy * s_row + ix * s_col;
            uint8_t *dst = sensor + (i * Wpx + j) * 3;
            dst[0] = src[0];
            dst[1] = src[s_chan];
            dst[2] = src[2 * s_chan];
        }
    }
    return NVO_OK;
}

/* A2: block "chemical" downscale.  navsim/util.pyx:91-134.
 * image [R][C][3] contiguous -> out [R/fr][C/fc][3]. */
void nvo_downscale_chem(const uint8_t *image, long R, long C, long fr, long fc, uint8_t *out)
{
    long conc[256];
    long nrb = R / fr, ncb = C / fc;         /* :102 */
    for (long bi = 0; bi < nrb; bi++) {
        for (long bj = 0; bj < ncb; bj++) {
            for (int k = 0; k < 256; k++) conc[k] = 0;   /* :112-113 */
            double avg = 0;
            for (long i = 0; i < fr; i++) {
                for (long j = 0; j < fc; j++) {
                    const uint8_t *p = image + ((bi * fr + i) * C + (bj * fc + j)) * 3;
                    conc[p[0]] += p[1];      /* :117-119 */
                    avg += p[2];             /* :120 */
                }
            }
            avg /= (double)(fr * fc);        /* :121 */
            avg = round(avg);                /* :122 */
            uint8_t *o = out + (bi * ncb + bj) * 3;
            o[2] = (uint8_t)avg;             /* :123 */
            uint8_t most = 0;                /* :126-129, strict > : lowest hue wins ties */
            for (int k = 0; k < 256; k++)
                if (conc[k] > conc[most]) most = (uint8_t)k;
            o[0] = most;
            /* :131 -- cdivision(True) makes '/' C integer division, evaluated
             * left to right: (conc / fr) * fc.  The double -> uint8 cast of a
             * value above 255 wraps modulo 256 on x86. */
            double sat = round((double)((conc[most] / fr) * fc));
            o[1] = (uint8_t)(long long)sat;  /* :132 */
        }
    }
}

/* A3: per-channel level quantisation as a 256-entry table.
 * navsim/NavBySceneFamiliarity.py:178-186: float32 x/255*(n-1) -> rint ->
 * /(n-1)*255 -> truncating store into uint8. */
void nvo_quant_lut(int nlevels, uint8_t *lut)
{
    for (int x = 0; x < 256; x++) {
        volatile float v = (float)x;
        v = v / 255.0f;
        v = v * (float)(nlevels - 1);
        v = rintf(v);
        v = v / (float)(nlevels - 1);
        v = v * 255.0f;
        lut[x] = (uint8_t)v;
    }
}

/* A4: bounds test + A1 + A2 + A3 + centre mask.
 * navsim/NavBySceneFamiliarity.py:151-192.  out [H][W][3]; scratch holds the
 * full-resolution buffer [H*ph][W*pw][3]. */
int nvo_get_sensor_mat(const nvo_world *w, double x, double y, double angle,
                       uint8_t *out, uint8_t *scratch)
{
    long Wpx = w->W * w->pw, Hpx = w->H * w->ph;
    double r = (double)(Wpx > Hpx ? Wpx : Hpx) / 2.0;     /* :94 */
    if (x <= r || y <= r || x >= (double)w->cols - r || y >= (double)w->rows - r)
        return NVO_OUT_OF_BOUNDS;                          /* :156-158 */
    int rc = nvo_fill_sensor(scratch, Hpx, Wpx, x, y, angle, w->land, w->rows, w->cols,
                             w->s_row, w->s_col, w->s_chan); /* :161-166 */
    if (rc) return rc;
    nvo_downscale_chem(scratch, Hpx, Wpx, w->ph, w->pw, out); /* :169-173 */
    long P = w->W * w->H;
    for (long p = 0; p < P; p++)                            /* :178-186 */
        for (int ch = 0; ch < 3; ch++) out[p * 3 + ch] = w->lut[ch][out[p * 3 + ch]];
    long r1 = w->W / 2;                                     /* :189-190 */
    long lo = r1 - w->mask_middle_n, hi = r1 + w->mask_middle_n;
    /* Python slice semantics: negative bounds wrap once, then clamp. */
    if (lo < 0) { lo += w->W; if (lo < 0) lo = 0; }
    if (hi < 0) { hi += w->W; if (hi < 0) hi = 0; }
    if (hi > w->W) hi = w->W;
    for (long i = 0; i < w->H; i++)
        for (long j = lo; j < hi; j++) {
            uint8_t *o = out + (i * w->W + j) * 3;
            o[0] = o[1] = o[2] = 0;
        }
    return NVO_OK;
}

/* A5: HSV sum-of-absolute-differences familiarity, exact FP64 operation
 * order.  navsim/util.pyx:28-73. */
void nvo_sads_hsv(const uint8_t *scenes, long N, long H, long W, const uint8_t *scene,
                  double *fambuf, double cw)
{
    long P = H * W;
    double maxfam = (double)(H * W);                       /* :42 */
    for (long n = 0; n < N; n++) {
        const uint8_t *f = scenes + n * P * 3;
        double diff = 0.0;
        for (long p = 0; p < P; p++) {                      /* row-major i,j :46-47 */
            double t;
            if (scene[p * 3] == f[p * 3])                   /* :48 */
                t = (double)abs((int)scene[p * 3 + 1] - (int)f[p * 3 + 1]); /* :50 */
            else
                t = (double)((int)scene[p * 3 + 1] + (int)f[p * 3 + 1]);    /* :56 */
            t *= 0.5;                                       /* :59 */
            t *= cw;                                        /* :68 */
            t += (1 - cw) * (double)abs((int)scene[p * 3 + 2] - (int)f[p * 3 + 2]); /* :69 */
            t /= 255.;                                      /* :71 */
            diff += t;                                      /* :72 */
        }
        fambuf[n] = maxfam - diff;                          /* :73 */
    }
}

/* Integer surrogates of A5 (not in the reference): per view, the sum of the
 * hue/saturation term X and the sum of |dV|.  The CUDA distance kernel works
 * on these; the tests check them against nvo_sads_hsv. */
void nvo_sad_int(const uint8_t *scenes, long N, long P, const uint8_t *scene,
                 uint32_t *x_tot, uint32_t *v_tot)
{
    for (long n = 0; n < N; n++) {
        const uint8_t *f = scenes + n * P * 3;
        uint32_t xs = 0, vs = 0;
        for (long p = 0; p < P; p++) {
            int sq = scene[p * 3 + 1], sn = f[p * 3 + 1];
            xs += (scene[p * 3] == f[p * 3]) ? (uint32_t)abs(sq - sn) : (uint32_t)(sq + sn);
            vs += (uint32_t)abs((int)scene[p * 3 + 2] - (int)f[p * 3 + 2]);
        }
        if (x_tot) x_tot[n] = xs;
        v_tot[n] = vs;
    }
}

/* A8: library build.  navsim/NavBySceneFamiliarity.py:118-140.
 * Returns 0, or the failing status with *bad_index set. */
int nvo_train_from_path(const nvo_world *w, const double *pts, long N, uint8_t *scenes,
                        uint8_t *scratch, long *bad_index)
{
    long P3 = w->W * w->H * 3;
    double ang = 0.0;
    for (long i = 0; i < N; i++) {
        if (i < N - 1) {                                    /* :124-126 */
            double dx = pts[2 * (i + 1)] - pts[2 * i];
            double dy = pts[2 * (i + 1) + 1] - pts[2 * i + 1];
            ang = atan2(dy, dx);
        }                                                   /* :132 last point reuses ang */
        int rc = nvo_get_sensor_mat(w, pts[2 * i], pts[2 * i + 1], ang, scenes + i * P3, scratch);
        if (rc) { if (bad_index) *bad_index = i; return rc; }
    }
    return NVO_OK;
}

/* update_error.  navsim/NavBySceneFamiliarity.py:252-276. */
static int update_error(const nvo_world *w, nvo_agent *a)
{
    a->navigated_for_frames += 1;                           /* :253 */
    double dmin = INFINITY;
    for (long n = 0; n < w->N; n++) {                       /* :255-258 */
        double dx = w->path[2 * n] - a->x, dy = w->path[2 * n + 1] - a->y;
        double d = sqrt(dx * dx + dy * dy);
        if (d < dmin) dmin = d;
    }
    if (dmin > w->max_dist) return NVO_TOO_FAR;             /* :263-264 */
    a->nav_err += dmin * dmin;                              /* :267 */
    a->n_nav_err += 1;                                      /* :268 */
    double thr = w->coverage_factor * w->step_size;         /* :271 */
    if (dmin <= thr) {                                      /* :272-276 */
        for (long n = 0; n < w->N; n++) {
            double dx = w->path[2 * n] - a->x, dy = w->path[2 * n + 1] - a->y;
            double d = sqrt(dx * dx + dy * dy);
            if (d <= thr) a->coverage[n] = 1;
        }
    }
    return NVO_OK;
}

/* A6 + A7: one agent step.  navsim/NavBySceneFamiliarity.py:279-329.
 * angle_fam [A] (NaN-initialised like :286), scene_fam [N] or NULL (:287,
 * :301-303), best_out / step_fam_out optional.  sad_min_out [A] (optional)
 * receives the integer |dV| sum of the most familiar view per heading.
 * scratch: [Hpx*Wpx*3 + H*W*3] bytes; fam_tmp: [N] doubles. */
int nvo_step_forward(const nvo_world *w, nvo_agent *a, int fake, double *angle_fam,
                     double *scene_fam, long *best_out, double *step_fam_out,
                     uint8_t *scratch, double *fam_tmp)
{
    long Wpx = w->W * w->pw, Hpx = w->H * w->ph;
    uint8_t *smat = scratch + Hpx * Wpx * 3;
    for (long k = 0; k < w->A; k++) angle_fam[k] = NAN;     /* :286 */
    if (scene_fam)
        for (long n = 0; n < w->N; n++) scene_fam[n] = INFINITY; /* :287 */
    for (long k = 0; k < w->A; k++) {                        /* :289 */
        double angle = py_mod(a->angle + w->offsets[k], 2 * M_PI); /* :291 */
        int rc = nvo_get_sensor_mat(w, a->x, a->y, angle, smat, scratch); /* :293 */
        if (rc) return rc;
        nvo_sads_hsv(w->scenes, w->N, w->H, w->W, smat, fam_tmp, w->chem_weight); /* :299 */
        double best = -INFINITY;
        for (long n = 0; n < w->N; n++) {
            if (scene_fam && fam_tmp[n] < scene_fam[n]) scene_fam[n] = fam_tmp[n]; /* :301-303 */
            if (fam_tmp[n] > best) best = fam_tmp[n];
        }
        angle_fam[k] = best;                                /* :313 */
    }
    long bi = 0;                                            /* :315 first maximum */
    for (long k = 1; k < w->A; k++)
        if (angle_fam[k] > angle_fam[bi]) bi = k;
    if (best_out) *best_out = bi;
    if (step_fam_out) *step_fam_out = angle_fam[bi];        /* :316 */
    double angle = py_mod(a->angle + w->offsets[bi], 2 * M_PI); /* :317 */
    a->x = a->x + w->step_size * cos(angle);                /* :319 */
    a->y = a->y + w->step_size * sin(angle);                /* :320 */
    a->angle = angle;                                       /* :323 */
    if (!fake) {
        int rc = update_error(w, a);                        /* :326 */
        if (rc) return rc;
        double ex = w->path[2 * (w->N - 1)] - a->x, ey = w->path[2 * (w->N - 1) + 1] - a->y;
        if (sqrt(ex * ex + ey * ey) <= w->threshold_factor * w->step_size) /* :328 */
            return NVO_REACHED_END;
    }
    return NVO_OK;
}

/* run_experiment's frame loop (scripts/run_experiment.py:235-249) for one
 * agent, logging every step.  Returns the stop status (0 = ran out of
 * frames); *completed counts steps that returned normally (:243-245).
 * Logs (each optional): best_idx [frames], pos [frames][3] (x, y, angle after
 * the step), afam [frames][A]. */
int nvo_run(const nvo_world *w, nvo_agent *a, long frames, long *completed,
            int32_t *best_idx, double *pos, double *afam)
{
    long Wpx = w->W * w->pw, Hpx = w->H * w->ph;
    uint8_t *scratch = (uint8_t *)malloc((size_t)(Hpx * Wpx * 3 + w->W * w->H * 3));
    double *fam_tmp = (double *)malloc(sizeof(double) * (size_t)w->N);
    double *af = (double *)malloc(sizeof(double) * (size_t)w->A);
    int status = NVO_OK;
    long done = 0;
    for (long f = 0; f < frames; f++) {
        long bi = -1;
        int rc = nvo_step_forward(w, a, 0, af, NULL, &bi, NULL, scratch, fam_tmp);
        if (afam) memcpy(afam + f * w->A, af, sizeof(double) * (size_t)w->A);
        if (best_idx) best_idx[f] = (int32_t)bi;
        if (pos) { pos[3 * f] = a->x; pos[3 * f + 1] = a->y; pos[3 * f + 2] = a->angle; }
        if (rc) { status = rc; break; }
        done++;
    }
    *completed = done;
    free(scratch); free(fam_tmp); free(af);
    return status;
}

/* Many independent agents on one world, one after another (the caller
 * parallelises over processes, like mpirun over trials,
 * scripts/run_experiment.py:327-328).  poses [B][3] in/out; status,
 * completed [B]; nav_err, n_nav_err [B]; coverage [B][N]. */
void nvo_run_batch(const nvo_world *w, long B, double *poses, long frames, int32_t *status,
                   int64_t *completed, double *nav_err, int64_t *n_nav_err,
                   uint8_t *coverage, int32_t *best_idx /*[B][frames] or NULL*/)
{
    for (long b = 0; b < B; b++) {
        nvo_agent a;
        a.x = poses[3 * b]; a.y = poses[3 * b + 1]; a.angle = poses[3 * b + 2];
        a.navigated_for_frames = 0; a.nav_err = 0.0; a.n_nav_err = 0;
        a.coverage = coverage + b * w->N;
        memset(a.coverage, 0, (size_t)w->N);
        long done = 0;
        if (best_idx)
            for (long f = 0; f < frames; f++) best_idx[b * frames + f] = -1;
        status[b] = nvo_run(w, &a, frames, &done, best_idx ? best_idx + b * frames : NULL, NULL, NULL);
        completed[b] = done;
        poses[3 * b] = a.x; poses[3 * b + 1] = a.y; poses[3 * b + 2] = a.angle;
        nav_err[b] = a.nav_err; n_nav_err[b] = a.n_nav_err;
    }
}
