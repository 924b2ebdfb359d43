/*
 * oracle/rangelib_oracle.c -- CPU restatement of the batched lidar scan path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under pyracecarsimulator_b200/ may import,
 * link or execute this file; it is the checker for tests/, for
 * __graft_entry__.smoke() and for bench.py's cpu_baseline / --impl reference
 * legs.  The product path is the CUDA library behind include/rangelib_b200.h.
 *
 * PARITY UNPINNED.  The arithmetic of this path lives in the third-party
 * Python extension `range_libc` (github.com/felrock/range_libc, a fork of
 * github.com/kctess5/range_libc; no commit pinned anywhere in the reference:
 * README.md:22, README.md:86, .gitignore:6, package.xml:51-63).  Its source is
 * not under /root/reference and the reference ships no tests, golden vectors
 * or fixtures for it, so this file restates the published algorithm
 * (RangeLib.h: OMap, DistanceTransform, RangeMethod::numpy_calc_range,
 * numpy_calc_range_angles, RayMarching::calc_range; vendor/distance_transform.h:
 * Felzenszwalb & Huttenlocher 1-D lower-envelope transform) as specified in
 * SURVEY.md Appendix A, anchored on the reference's own call sites:
 *   scripts/scan_simulator.py:72-76    PyRayMarching / PyRayMarchingGPU ctor
 *   scripts/scan_simulator.py:103-106  4-arg calc_range_many, single pose
 *   scripts/scan_simulator.py:130-133  4-arg calc_range_many, batch
 *   scripts/two_player/scan.py:56-70   2-arg calc_range_many + beam convention
 *   scripts/ros_interface.py:80-86     binarisation
 *   scripts/ros_interface.py:210       PyOMap(map_msg)
 *   scripts/racecar_simulator_v2.py:196 max_range_px = int(max_range / res)
 * Cross-checks that do not depend on the recall being right live in
 * tests/test_oracle_properties.py (brute-force EDT, scipy EDT, analytic box
 * room, DDA caster, equivariance).
 *
 * Floating-point conventions fixed here (SURVEY.md A.4/A.5/A.7 level 3) and
 * mirrored instruction for instruction by the CUDA kernels:
 *   - all march arithmetic is fp32; build with -ffp-contract=off so that only
 *     the fmaf() calls written below are fused;
 *   - sample position  p = fmaf(dir, t, p0); hit distance sqrtf(fmaf(xd,xd,yd*yd));
 *   - rotation  x' = fmaf(c, x, -(s*y)),  y' = fmaf(s, x, c*y);
 *   - beam fan  a_j = fmaf((float)j, fov/(float)num_rays, -0.5f*fov), theta_j = theta + a_j.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

#define ORC_EXPORT __attribute__((visibility("default")))

/* ------------------------------------------------------------------------- */
/* A.1  PGM pixel -> map_server OccupancyGrid cell, y-flipped.                 */
/* maps/map.yaml:1-6 supplies negate / occupied_thresh / free_thresh; every   */
/* shipped yaml uses the default trinary mode (mode 0).  Modes 1 (scale) and  */
/* 2 (raw) restate ROS1 map_server's image_loader for 8-bit grey images       */
/* (not exercised by the reference's own maps).          */
/* ------------------------------------------------------------------------- */
static int8_t mapserver_cell(int p, int negate, double occupied_thresh, double free_thresh, int mode)
{
    if (mode == 2) return (int8_t)(unsigned char)(negate ? 255 - p : p);   /* raw: the (negated) pixel value itself */
    double shade = negate ? p / 255.0 : (255 - p) / 255.0;
    if (shade > occupied_thresh) return 100;
    if (shade < free_thresh) return 0;
    if (mode == 0) return -1;                                   /* trinary: unknown */
    double ratio = (shade - free_thresh) / (occupied_thresh - free_thresh);
    return (int8_t)(unsigned char)(1 + 98 * ratio);             /* scale */
}

ORC_EXPORT void orc_mapserver_occupancy_mode(const uint8_t *img, int img_w, int img_h, int negate,
                                             double occupied_thresh, double free_thresh, int mode,
                                             int8_t *grid)
{
    for (int j = 0; j < img_h; ++j) {
        int8_t *dst = grid + (size_t)(img_h - 1 - j) * img_w;
        const uint8_t *src = img + (size_t)j * img_w;
        for (int i = 0; i < img_w; ++i) dst[i] = mapserver_cell(src[i], negate, occupied_thresh, free_thresh, mode);
    }
}

/* Colour / alpha images: ROS1 map_server image_loader.cpp (third-party, not in the reference checkout;   */
/* launch/simulate.launch:8-9 hands it whatever image the yaml names).  Per pixel of `channels` bytes:    */
/*   avg_channels = (mode == trinary || !has_alpha) ? channels : channels - 1;                            */
/*   color_avg = sum(first avg_channels bytes) / (double)avg_channels;  alpha = channels == 1 ? 1 : last  */
/*   byte;  negate -> 255 - color_avg;  raw: value = color_avg;  else occ = (255 - color_avg) / 255 and   */
/*   > occupied_thresh -> 100, < free_thresh -> 0, trinary or alpha < 1 -> -1, else 1 + 98 * ratio.       */
ORC_EXPORT void orc_mapserver_occupancy_channels(const uint8_t *img, int img_w, int img_h, int channels,
                                                 int has_alpha, int negate, double occupied_thresh,
                                                 double free_thresh, int mode, int8_t *grid)
{
    const int avg = (mode == 0 || !has_alpha) ? channels : channels - 1;
    for (int j = 0; j < img_h; ++j) {
        int8_t *dst = grid + (size_t)(img_h - 1 - j) * img_w;
        for (int i = 0; i < img_w; ++i) {
            const uint8_t *p = img + ((size_t)j * img_w + i) * channels;
            int sum = 0;
            for (int k = 0; k < avg; ++k) sum += p[k];
            double color_avg = sum / (double)avg;
            const double alpha = channels == 1 ? 1.0 : (double)p[channels - 1];
            if (negate) color_avg = 255 - color_avg;
            if (mode == 2) { dst[i] = (int8_t)(unsigned char)color_avg; continue; }
            const double occ = (255 - color_avg) / 255.0;
            if (occ > occupied_thresh) dst[i] = 100;
            else if (occ < free_thresh) dst[i] = 0;
            else if (mode == 0 || alpha < 1.0) dst[i] = -1;
            else dst[i] = (int8_t)(unsigned char)(1 + 98 * ((occ - free_thresh) / (occupied_thresh - free_thresh)));
        }
    }
}

ORC_EXPORT void orc_mapserver_occupancy(const uint8_t *img, int img_w, int img_h,
                                        int negate, double occupied_thresh,
                                        double free_thresh, int8_t *grid)
{
    for (int j = 0; j < img_h; ++j) {
        int8_t *dst = grid + (size_t)(img_h - 1 - j) * img_w; /* row 0 = bottom image row */
        const uint8_t *src = img + (size_t)j * img_w;
        for (int i = 0; i < img_w; ++i) {
            double shade = negate ? src[i] / 255.0 : (255 - src[i]) / 255.0;
            int8_t v = -1;
            if (shade > occupied_thresh) v = 100;
            else if (shade < free_thresh) v = 0;
            dst[i] = v;
        }
    }
}

/* ------------------------------------------------------------------------- */
/* A.2  binarise (scripts/ros_interface.py:80-86: >0 -> 255 else 0) followed  */
/* by PyOMap's `> 10` occupancy cut.  binarise=0 skips the first half (a      */
/* caller that hands PyOMap a raw OccupancyGrid, scripts/two_player/scan.py:45)*/
/* ------------------------------------------------------------------------- */
ORC_EXPORT void orc_omap_from_grid(const int8_t *grid, int64_t n, int binarise,
                                   uint8_t *occupied)
{
    for (int64_t i = 0; i < n; ++i) {
        int v = grid[i];
        if (binarise) v = (v > 0) ? 255 : 0;
        occupied[i] = (v > 10) ? 1 : 0;
    }
}

/* ------------------------------------------------------------------------- */
/* A.3  DistanceTransform: float lower-envelope transform, INF = 1e20f,       */
/* one axis then the other, then sqrtf.  `rows` = OMap.width (msg rows),      */
/* `cols` = OMap.height (msg columns); storage dist[row*cols + col].          */
/* ------------------------------------------------------------------------- */
#define ORC_INF 1e20f

static void envelope_1d_float(const float *f, int n, float *d, int *v, float *z)
{
    int k = 0;
    v[0] = 0;
    z[0] = -ORC_INF;
    z[1] = ORC_INF;
    for (int q = 1; q < n; ++q) {
        float fq = f[q] + (float)(q * q);
        float s;
        for (;;) {
            int p = v[k];
            s = (fq - (f[p] + (float)(p * p))) / (float)(2 * q - 2 * p);
            if (s <= z[k]) --k; else break;
        }
        ++k;
        v[k] = q;
        z[k] = s;
        z[k + 1] = ORC_INF;
    }
    k = 0;
    for (int q = 0; q < n; ++q) {
        while (z[k + 1] < (float)q) ++k;
        int p = v[k];
        d[q] = (float)((q - p) * (q - p)) + f[p];
    }
}

ORC_EXPORT void orc_edt_float(const uint8_t *occupied, int rows, int cols,
                              float *dist, float *dist2 /* nullable */)
{
    int n = rows > cols ? rows : cols;
    float *f = (float *)malloc(sizeof(float) * n);
    float *d = (float *)malloc(sizeof(float) * n);
    int *v = (int *)malloc(sizeof(int) * n);
    float *z = (float *)malloc(sizeof(float) * (n + 1));
    float *g = (float *)malloc(sizeof(float) * (size_t)rows * cols);
    for (size_t i = 0; i < (size_t)rows * cols; ++i) g[i] = occupied[i] ? 0.0f : ORC_INF;
    /* along the second index (msg columns) for every row */
    for (int r = 0; r < rows; ++r) {
        memcpy(f, g + (size_t)r * cols, sizeof(float) * cols);
        envelope_1d_float(f, cols, d, v, z);
        memcpy(g + (size_t)r * cols, d, sizeof(float) * cols);
    }
    /* along the first index (msg rows) for every column */
    for (int c = 0; c < cols; ++c) {
        for (int r = 0; r < rows; ++r) f[r] = g[(size_t)r * cols + c];
        envelope_1d_float(f, rows, d, v, z);
        for (int r = 0; r < rows; ++r) g[(size_t)r * cols + c] = d[r];
    }
    for (size_t i = 0; i < (size_t)rows * cols; ++i) {
        if (dist2) dist2[i] = g[i];
        dist[i] = sqrtf(g[i]);
    }
    free(f); free(d); free(v); free(z); free(g);
}

/* Exact integer squared EDT (Meijster, Roerdink & Hesselink 2000), the       */
/* definition of "reference DT" above 2896 px/side where the float envelope   */
/* is no longer provably exact.  Unreachable cells (no        */
/* occupied cell anywhere) get ORC_D2_INF.                                    */
#define ORC_D2_INF 0x3fffffff
#define ORC_G_INF ((int64_t)1 << 28)

ORC_EXPORT void orc_edt_exact(const uint8_t *occupied, int rows, int cols, int32_t *dist2)
{
    int64_t *g = (int64_t *)malloc(sizeof(int64_t) * (size_t)rows * cols);
    /* phase 1: nearest occupied cell along the first index, per column */
    for (int c = 0; c < cols; ++c) {
        int64_t run = ORC_G_INF;
        for (int r = 0; r < rows; ++r) {
            run = occupied[(size_t)r * cols + c] ? 0 : (run >= ORC_G_INF ? ORC_G_INF : run + 1);
            g[(size_t)r * cols + c] = run;
        }
        run = ORC_G_INF;
        for (int r = rows - 1; r >= 0; --r) {
            run = occupied[(size_t)r * cols + c] ? 0 : (run >= ORC_G_INF ? ORC_G_INF : run + 1);
            if (run < g[(size_t)r * cols + c]) g[(size_t)r * cols + c] = run;
        }
    }
    /* phase 2: integer lower envelope along the second index, per row */
    int *s = (int *)malloc(sizeof(int) * cols);
    int64_t *t = (int64_t *)malloc(sizeof(int64_t) * cols);
    for (int r = 0; r < rows; ++r) {
        const int64_t *gr = g + (size_t)r * cols;
        int32_t *out = dist2 + (size_t)r * cols;
        int q = 0;
        s[0] = 0; t[0] = 0;
#define F_(x, i) (((int64_t)(x) - (i)) * ((int64_t)(x) - (i)) + gr[i] * gr[i])
        for (int u = 1; u < cols; ++u) {
            while (q >= 0 && F_(t[q], s[q]) > F_(t[q], u)) --q;
            if (q < 0) { q = 0; s[0] = u; t[0] = 0; }
            else {
                int64_t i = s[q];
                int64_t num = (int64_t)u * u - i * i + gr[u] * gr[u] - gr[i] * gr[i];
                int64_t den = 2 * ((int64_t)u - i);
                /* floor division (num may be negative) */
                int64_t sep = num / den;
                if ((num % den != 0) && ((num < 0) != (den < 0))) --sep;
                int64_t w = 1 + sep;
                if (w < cols) { ++q; s[q] = u; t[q] = w; }
            }
        }
        for (int u = cols - 1; u >= 0; --u) {
            int64_t val = F_(u, s[q]);
            out[u] = (val >= ORC_D2_INF) ? ORC_D2_INF : (int32_t)val;
            if (u == t[q]) --q;
        }
#undef F_
    }
    free(s); free(t); free(g);
}

/* dist = sqrt_rn((float)d2); unreachable -> sqrtf(1e20f), what orc_edt_float gives */
ORC_EXPORT void orc_sqrt_dist2(const int32_t *dist2, int64_t n, float *dist)
{
    for (int64_t i = 0; i < n; ++i)
        dist[i] = (dist2[i] >= ORC_D2_INF) ? sqrtf(ORC_INF) : sqrtf((float)dist2[i]);
}

/* ------------------------------------------------------------------------- */
/* A.4  marcher                                                               */
/* ------------------------------------------------------------------------- */
typedef struct {
    int width;            /* OMap.width  = msg.info.height (rows)    */
    int height;           /* OMap.height = msg.info.width  (columns) */
    const float *dist;    /* dist[x*height + y], borrowed            */
    float max_range;      /* pixels */
    float world_scale, world_angle, world_origin_x, world_origin_y;
    float world_sin_angle, world_cos_angle;
    float inv_world_scale, rotation_const;
} orc_marcher;

ORC_EXPORT orc_marcher *orc_marcher_create(const float *dist, int rows, int cols,
                                           float max_range_px, double resolution,
                                           double origin_x, double origin_y, double yaw)
{
    orc_marcher *m = (orc_marcher *)calloc(1, sizeof(orc_marcher));
    m->width = rows;
    m->height = cols;
    m->dist = dist;
    m->max_range = max_range_px;
    double angle = -1.0 * yaw;                 /* PyOMap: world_angle = -yaw */
    m->world_scale = (float)resolution;
    m->world_angle = (float)angle;
    m->world_origin_x = (float)origin_x;
    m->world_origin_y = (float)origin_y;
    m->world_sin_angle = (float)sin(angle);
    m->world_cos_angle = (float)cos(angle);
    m->inv_world_scale = (float)(1.0 / (double)m->world_scale);
    m->rotation_const = (float)(-1.0 * (double)m->world_angle - 3.0 * M_PI / 2.0);
    return m;
}

ORC_EXPORT void orc_marcher_destroy(orc_marcher *m) { free(m); }

/* RayMarching::calc_range in grid coordinates; *steps counts DT loads. */
static inline float march_grid(const orc_marcher *m, float x0, float y0, float heading,
                               int32_t *steps)
{
    const float dx = cosf(heading);
    const float dy = sinf(heading);
    const float mr = m->max_range;
    const float fw = (float)m->width, fh = (float)m->height;
    float t = 0.0f;
    int32_t n = 0;
    while (t < mr) {
        float fx = fmaf(dx, t, x0);
        float fy = fmaf(dy, t, y0);
        /* (int) truncation toward zero: (-1,0) -> cell 0 is in bounds; NaN and  */
        /* out-of-int-range behave like x86 cvttss2si (INT_MIN) -> out of map.   */
        if (!(fx > -1.0f && fx < fw && fy > -1.0f && fy < fh)) { if (steps) *steps = n; return mr; }
        int px = (int)fx, py = (int)fy;
        float d = m->dist[(size_t)px * m->height + py];
        ++n;
        if (d <= 0.0f) {
            float xd = (float)px - x0;
            float yd = (float)py - y0;
            if (steps) *steps = n;
            return sqrtf(fmaf(xd, xd, yd * yd));
        }
        t += fmaxf(d * 0.999f, 1.0f);
    }
    if (steps) *steps = n;
    return mr;
}

typedef struct { float x, y, theta; } grid_pose;

static inline grid_pose world_to_grid(const orc_marcher *m, float xw, float yw, float thw)
{
    float x = (xw - m->world_origin_x) * m->inv_world_scale;
    float y = (yw - m->world_origin_y) * m->inv_world_scale;
    grid_pose g;
    g.x = fmaf(m->world_cos_angle, x, -(m->world_sin_angle * y));
    g.y = fmaf(m->world_sin_angle, x, m->world_cos_angle * y);
    g.theta = -thw + m->rotation_const;
    return g;
}

/* calc_range(x, y, theta): world pose in, metres out (PyRayMarching.calc_range). */
ORC_EXPORT float orc_calc_range(const orc_marcher *m, float xw, float yw, float thw)
{
    grid_pose g = world_to_grid(m, xw, yw, thw);
    return march_grid(m, g.y, g.x, g.theta, NULL) * m->world_scale;
}

/* ---- minimal pthread parallel-for (this image's default $CC has no libgomp) ---- */
typedef void (*orc_body)(void *ctx, int64_t begin, int64_t end);
typedef struct { orc_body body; void *ctx; int64_t n, chunk; int64_t *next; } orc_job;

static void *orc_worker(void *arg)
{
    orc_job *j = (orc_job *)arg;
    for (;;) {
        int64_t b = __atomic_fetch_add(j->next, j->chunk, __ATOMIC_RELAXED);
        if (b >= j->n) break;
        int64_t e = b + j->chunk < j->n ? b + j->chunk : j->n;
        j->body(j->ctx, b, e);
    }
    return NULL;
}

ORC_EXPORT int orc_max_threads(void)
{
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

/* Persistent worker pool: threads are created once and parked on a condition variable, so a  */
/* short call does not pay thread start-up and CPU migration every time.                       */
static struct {
    pthread_mutex_t mu;
    pthread_cond_t start, done;
    pthread_t tid[256];
    int n;                 /* workers created */
    unsigned long gen;     /* job generation */
    orc_job *job;
    int want, pending;     /* workers that should take part / still running */
} pool = { PTHREAD_MUTEX_INITIALIZER, PTHREAD_COND_INITIALIZER, PTHREAD_COND_INITIALIZER, {0}, 0, 0, NULL, 0, 0 };

static void *pool_main(void *arg)
{
    const int me = (int)(intptr_t)arg;
    unsigned long seen = 0;
    pthread_mutex_lock(&pool.mu);
    for (;;) {
        while (pool.gen == seen) pthread_cond_wait(&pool.start, &pool.mu);
        seen = pool.gen;
        if (me >= pool.want) continue;
        orc_job *j = pool.job;
        pthread_mutex_unlock(&pool.mu);
        orc_worker(j);
        pthread_mutex_lock(&pool.mu);
        if (--pool.pending == 0) pthread_cond_signal(&pool.done);
    }
    return NULL;
}

/* threads: 1 = caller's thread only (upstream's loop is single-threaded); 0 = all cores */
static void orc_parallel_for(int64_t n, int64_t chunk, int threads, orc_body body, void *ctx)
{
    if (threads <= 0) threads = orc_max_threads();
    if (threads > 256) threads = 256;
    if (threads == 1 || n <= chunk) { body(ctx, 0, n); return; }
    int64_t next = 0;
    orc_job job = { body, ctx, n, chunk, &next };
    pthread_mutex_lock(&pool.mu);
    while (pool.n < threads - 1) {
        if (pthread_create(&pool.tid[pool.n], NULL, pool_main, (void *)(intptr_t)pool.n) != 0) break;
        pthread_detach(pool.tid[pool.n]);
        ++pool.n;
    }
    const int helpers = pool.n < threads - 1 ? pool.n : threads - 1;
    pool.job = &job;
    pool.want = helpers;
    pool.pending = helpers;
    ++pool.gen;
    pthread_cond_broadcast(&pool.start);
    pthread_mutex_unlock(&pool.mu);
    orc_worker(&job);
    pthread_mutex_lock(&pool.mu);
    while (pool.pending > 0) pthread_cond_wait(&pool.done, &pool.mu);
    pthread_mutex_unlock(&pool.mu);
}

typedef struct {
    const orc_marcher *m; const float *ins; const float *angles; float *outs; int32_t *steps;
    int num_rays; float fov; int64_t pose_stride_rows;
} orc_call;

static void many_body(void *vc, int64_t b, int64_t e)
{
    orc_call *c = (orc_call *)vc;
    const orc_marcher *m = c->m;
    for (int64_t i = b; i < e; ++i) {
        grid_pose g = world_to_grid(m, c->ins[3 * i], c->ins[3 * i + 1], c->ins[3 * i + 2]);
        c->outs[i] = march_grid(m, g.y, g.x, g.theta, c->steps ? c->steps + i : NULL) * m->world_scale;
    }
}

ORC_EXPORT void orc_calc_range_many(const orc_marcher *m, const float *ins, float *outs,
                                    int64_t n, int32_t *steps /* nullable */, int threads)
{
    orc_call c = { m, ins, NULL, outs, steps, 0, 0.0f, 0 };
    orc_parallel_for(n, 2048, threads, many_body, &c);
}

/* 4-arg fork form (scripts/scan_simulator.py:103-106, :130-133): `n` rows, pose k */
/* in row k*num_rays, beam j heading theta - fov/2 + j*fov/num_rays (A.5).          */
/* pose_stride_rows = num_rays for the reference layout, 1 for compact (B,3).       */
static void fan_body(void *vc, int64_t b, int64_t e)
{
    orc_call *c = (orc_call *)vc;
    const orc_marcher *m = c->m;
    const int num_rays = c->num_rays;
    const float *ins = c->ins;
    float *outs = c->outs;
    int32_t *steps = c->steps;
    const int64_t pose_stride_rows = c->pose_stride_rows;
    const float inc = c->fov / (float)num_rays;
    const float half = -0.5f * c->fov;
    for (int64_t k = b; k < e; ++k) {
        const float *p = ins + 3 * k * pose_stride_rows;
        float x = (p[0] - m->world_origin_x) * m->inv_world_scale;
        float y = (p[1] - m->world_origin_y) * m->inv_world_scale;
        float gx = fmaf(m->world_cos_angle, x, -(m->world_sin_angle * y));
        float gy = fmaf(m->world_sin_angle, x, m->world_cos_angle * y);
        for (int j = 0; j < num_rays; ++j) {
            float thw = p[2] + fmaf((float)j, inc, half);
            float thg = -thw + m->rotation_const;
            int64_t o = k * num_rays + j;
            outs[o] = march_grid(m, gy, gx, thg, steps ? steps + o : NULL) * m->world_scale;
        }
    }
}

ORC_EXPORT void orc_calc_range_fan(const orc_marcher *m, const float *ins, float *outs,
                                   int64_t num_poses, int num_rays, float fov,
                                   int64_t pose_stride_rows, int32_t *steps, int threads)
{
    orc_call c = { m, ins, NULL, outs, steps, num_rays, fov, pose_stride_rows };
    orc_parallel_for(num_poses, 4, threads, fan_body, &c);
}

/* calc_range_repeat_angles(ins, angles, outs): heading theta_g - angles[a]. */
static void angles_body(void *vc, int64_t b, int64_t e)
{
    orc_call *c = (orc_call *)vc;
    const orc_marcher *m = c->m;
    const int num_angles = c->num_rays;
    const float *ins = c->ins, *angles = c->angles;
    float *outs = c->outs;
    int32_t *steps = c->steps;
    for (int64_t i = b; i < e; ++i) {
        grid_pose g = world_to_grid(m, ins[3 * i], ins[3 * i + 1], ins[3 * i + 2]);
        for (int a = 0; a < num_angles; ++a) {
            int64_t o = i * num_angles + a;
            outs[o] = march_grid(m, g.y, g.x, g.theta - angles[a], steps ? steps + o : NULL) *
                      m->world_scale;
        }
    }
}

ORC_EXPORT void orc_calc_range_repeat_angles(const orc_marcher *m, const float *ins,
                                             const float *angles, float *outs,
                                             int64_t num_poses, int num_angles,
                                             int32_t *steps, int threads)
{
    orc_call c = { m, ins, angles, outs, steps, num_angles, 0.0f, 1 };
    orc_parallel_for(num_poses, 64, threads, angles_body, &c);
}

/* ------------------------------------------------------------------------- */
/* A.7  convention variants -- NOT the oracle.                               */
/* The scan half of this file restates range_libc from its published          */
/* algorithm (PARITY UNPINNED); SURVEY.md A.5 / A.7 list the fp32 details the  */
/* restatement had to CHOOSE.  This function marches the fork's 4-arg fan with */
/* any subset of those choices flipped, so that a test can say how far the     */
/* results would move if upstream had chosen otherwise                        */
/* (tests/test_convention_sensitivity.py).                                    */
/*   1  no fused multiply-add anywhere (sample, hit distance, rotation, fan)  */
/*   2  (p - origin) / scale instead of * (1 / scale)                         */
/*   4  beam heading accumulated in double, rounded to float once             */
/*   8  cos / sin evaluated in double, rounded to float                       */
/*  16  beam heading by repeated fp32 addition (angle += inc)                 */
/*  32  heading = (-theta - world_angle) - 3 pi / 2 in two fp32 steps         */
/* ------------------------------------------------------------------------- */
typedef struct {
    const orc_marcher *m; const float *ins; float *outs; int num_rays; float fov; unsigned mask;
} orc_variant_call;

static inline float variant_march(const orc_marcher *m, float x0, float y0, float heading, unsigned mask)
{
    float dx, dy;
    if (mask & 8u) { dx = (float)cos((double)heading); dy = (float)sin((double)heading); }
    else { dx = cosf(heading); dy = sinf(heading); }
    const int nofma = (mask & 1u) != 0;
    const float mr = m->max_range;
    const float fw = (float)m->width, fh = (float)m->height;
    float t = 0.0f;
    while (t < mr) {
        float fx = nofma ? x0 + dx * t : fmaf(dx, t, x0);
        float fy = nofma ? y0 + dy * t : fmaf(dy, t, y0);
        if (!(fx > -1.0f && fx < fw && fy > -1.0f && fy < fh)) return mr;
        int px = (int)fx, py = (int)fy;
        float d = m->dist[(size_t)px * m->height + py];
        if (d <= 0.0f) {
            float xd = (float)px - x0;
            float yd = (float)py - y0;
            return nofma ? sqrtf(xd * xd + yd * yd) : sqrtf(fmaf(xd, xd, yd * yd));
        }
        t += fmaxf(d * 0.999f, 1.0f);
    }
    return mr;
}

static void variant_body(void *vc, int64_t b, int64_t e)
{
    orc_variant_call *c = (orc_variant_call *)vc;
    const orc_marcher *m = c->m;
    const unsigned mask = c->mask;
    const int n = c->num_rays;
    const float inc = c->fov / (float)n;
    const float half = -0.5f * c->fov;
    const double dinc = (double)c->fov / (double)n, dhalf = -0.5 * (double)c->fov;
    for (int64_t k = b; k < e; ++k) {
        const float *p = c->ins + 3 * k;
        float x, y;
        if (mask & 2u) { x = (p[0] - m->world_origin_x) / m->world_scale; y = (p[1] - m->world_origin_y) / m->world_scale; }
        else { x = (p[0] - m->world_origin_x) * m->inv_world_scale; y = (p[1] - m->world_origin_y) * m->inv_world_scale; }
        float gx, gy;
        if (mask & 1u) {
            gx = m->world_cos_angle * x - m->world_sin_angle * y;
            gy = m->world_sin_angle * x + m->world_cos_angle * y;
        } else {
            gx = fmaf(m->world_cos_angle, x, -(m->world_sin_angle * y));
            gy = fmaf(m->world_sin_angle, x, m->world_cos_angle * y);
        }
        float run = p[2] + half;
        for (int j = 0; j < n; ++j) {
            float thw;
            if (mask & 4u) thw = (float)((double)p[2] + (dhalf + (double)j * dinc));
            else if (mask & 16u) { thw = run; run += inc; }
            else if (mask & 1u) thw = p[2] + ((float)j * inc + half);
            else thw = p[2] + fmaf((float)j, inc, half);
            float thg;
            if (mask & 32u) thg = (-thw - m->world_angle) - (float)(3.0 * M_PI / 2.0);
            else thg = -thw + m->rotation_const;
            c->outs[k * n + j] = variant_march(m, gy, gx, thg, mask) * m->world_scale;
        }
    }
}

ORC_EXPORT void orc_variant_fan(const orc_marcher *m, const float *poses, float *outs, int64_t num_poses,
                                int num_rays, float fov, unsigned mask, int threads)
{
    orc_variant_call c = { m, poses, outs, num_rays, fov, mask };
    orc_parallel_for(num_poses, 4, threads, variant_body, &c);
}
