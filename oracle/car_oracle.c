/*
 * oracle/car_oracle.c -- CPU restatement of the reference vehicle model for the
 * fused rollout stage (north_star (c)).  TEST INFRASTRUCTURE ONLY (see the
 * header of rangelib_oracle.c).
 *
 * PINNED: unlike range_libc, this part of the reference is vendored and
 * compiles here, so oracle/Makefile also builds the UNMODIFIED
 * /root/reference/racecar/src/racecar.cpp into oracle/_ref/libracecar_ref.so
 * and tests/test_car_oracle.py checks this restatement against it bit for bit
 * (and against tests/golden/car_*.npz on boxes where /root/reference is absent).
 *
 * Follows racecar/src/racecar.cpp:
 *   :10-51   constructor, KP = 2*MAX_ACCEL/MAX_SPEED
 *   :53-98   updatePosition (model switch with hysteresis, accumulators, clamps)
 *   :118-169 computeFromInput (speed P-controller, bang-bang steering rate)
 *   :171-194 updateNormal (kinematic single track)
 *   :196-237 updateSingle (dynamic single track)
 *   :239-292 setCarEdgeDistances (angle incremented BEFORE use, PI = 3.145)
 *   :305-328 isCrashed (first crashed pose, else -(poses+1))
 *   :330-376 setState/getState (11 doubles)
 *   :378-387 getScanPose
 * and racecar/include/racecar.hpp:112-117 for K_THRESH/ST_THRESH/G/PI.
 * Build with -ffp-contract=off: every product and sum below rounds separately,
 * which is what the reference does when compiled without -mfma/-ffast-math.
 */
#include <math.h>
#include <stdint.h>

#define ORC_EXPORT __attribute__((visibility("default")))

/* parameter block, in the reference constructor's argument order */
typedef struct {
    double wb, fc, h_cg, l_f, l_r, cs_f, cs_r, mass, i_z;
    double crash_thresh, width, length;
    double max_steer_vel, max_steer_ang, max_speed, max_accel, max_decel;
} orc_car_params;

#define CAR_K_THRESH 0.5
#define CAR_ST_THRESH 0.53
#define CAR_G 9.81
#define CAR_PI 3.145 /* sic, racecar.hpp:117 */

static double clampd(double v, double lo, double hi) { return fmin(fmax(v, lo), hi); }

/* One updatePosition(dt) on an 11-double state with the current (speed, steer) targets. */
ORC_EXPORT void orc_car_step(const orc_car_params *p, double *st, double in_speed,
                             double in_steer, double dt)
{
    double x = st[0], y = st[1], th = st[2], v = st[3], sa = st[4];
    double w = st[5], beta = st[6];
    int dyn = st[7] > 0.0;

    /* computeFromInput */
    const double kp = 2.0 * p->max_accel / p->max_speed;
    double dv = in_speed - v, accel;
    if (v > 0) accel = (dv > 0) ? clampd(kp * dv, -p->max_accel, p->max_accel) : -p->max_decel;
    else       accel = (dv > 0) ? p->max_decel : clampd(kp * dv, -p->max_accel, p->max_accel);
    double ds = in_steer - sa, sv = 0;
    if (fabs(ds) > 0.0001) sv = (ds > 0) ? p->max_steer_vel : -p->max_steer_vel;

    const double px = x, py = y;
    const double thresh = dyn ? CAR_ST_THRESH : CAR_K_THRESH;
    if (v < thresh) {
        /* updateNormal */
        double xd = v * cos(th), yd = v * sin(th);
        double thd = v / p->wb * tan(sa);
        x += xd * dt; y += yd * dt; th += thd * dt;
        v += accel * dt; sa += sv * dt;
        w = 0; beta = 0; dyn = 0;
    } else {
        /* updateSingle */
        double xd = v * cos(th + beta), yd = v * sin(th + beta);
        double thd = w;
        double rv = CAR_G * p->l_r - accel * p->h_cg;
        double fv = CAR_G * p->l_f + accel * p->h_cg;
        double ratio = w / v;
        double first = p->fc / (v * (p->l_r + p->l_f));
        double wdd = (p->fc * p->mass / (p->i_z * p->wb)) *
                     (p->l_f * p->cs_f * sa * rv +
                      beta * (p->l_r * p->cs_r * fv - p->l_f * p->cs_f * rv) -
                      ratio * (pow(p->l_f, 2) * p->cs_f * rv + pow(p->l_r, 2) * p->cs_r * fv));
        double bd = first * (p->cs_f * sa * (rv) - beta * (p->cs_r * fv + p->cs_f * rv) +
                             ratio * (p->cs_r * p->l_r * fv - p->cs_f * p->l_f * rv)) - w;
        x += xd * dt; y += yd * dt; th += thd * dt;
        v += accel * dt; sa += sv * dt;
        w += wdd * dt; beta += bd * dt; dyn = 1;
    }
    double ddx = px - x, ddy = py - y;
    st[8] += sqrt(ddx * ddx + ddy * ddy);
    st[9] += v;
    st[10] = (double)((int)st[10] + 1);
    v = clampd(v, -p->max_speed, p->max_speed);
    sa = clampd(sa, -p->max_steer_ang, p->max_steer_ang);
    st[0] = x; st[1] = y; st[2] = th; st[3] = v; st[4] = sa;
    st[5] = w; st[6] = beta; st[7] = dyn ? 1.0 : 0.0;
}

/* MCTS.rollout's vehicle half for many cars (scripts/mcts.py:202-235): `steps` updatePosition(dt) per car with a
 * new (speed, steer) from actions[(c, i / every)] every `every`-th step; records the base-link pose after each
 * step narrowed to fp32 where the reference narrows it (the f32 pose buffer, mcts.py:211, :229-231), step-major
 * (steps, n, 3) like the device rollout.  states (n, 11) is updated in place. */
ORC_EXPORT void orc_car_rollout_poses(const orc_car_params *p, double *states, const double *actions, int64_t n,
                                      int steps, int every, double dt, float *poses)
{
    const int n_act = (steps + every - 1) / every;
    for (int64_t c = 0; c < n; ++c) {
        double *st = states + 11 * c;
        for (int i = 0; i < steps; ++i) {
            const double *a = actions + 2 * ((int64_t)n_act * c + i / every);
            orc_car_step(p, st, a[0], a[1], dt);
            float *o = poses + 3 * ((int64_t)i * n + c);
            o[0] = (float)st[0];
            o[1] = (float)st[1];
            o[2] = (float)st[2];
        }
    }
}

ORC_EXPORT void orc_car_scan_pose(const double *st, double scan_dist_to_base, double *pose)
{
    pose[0] = st[0] + scan_dist_to_base * cos(st[2]);
    pose[1] = st[1] + scan_dist_to_base * sin(st[2]);
    pose[2] = st[2];
}

ORC_EXPORT void orc_car_edge_distances(const orc_car_params *p, int num_rays, double min_ang,
                                       double inc, double scan_dist_to_base, double *edge)
{
    const double side = p->width / 2.0;
    const double front = p->wb - scan_dist_to_base;
    const double back = scan_dist_to_base;
    double a = min_ang;
    for (int i = 0; i < num_rays; ++i) {
        a += inc;
        if (a > 0.0) {
            if (a < CAR_PI / 2.0) edge[i] = fmin(side / sin(a), front / cos(a));
            else edge[i] = fmin(side / sin(a - CAR_PI / 2.0), back / cos(a - CAR_PI / 2.0));
        } else {
            if (a == 0.0) a += 0.0001;
            if (a > -CAR_PI / 2.0) edge[i] = fmin(side / sin(-a), front / cos(-a));
            else edge[i] = fmin(side / sin(-a - CAR_PI / 2.0), back / cos(-a - CAR_PI / 2.0));
        }
    }
}

ORC_EXPORT int orc_car_is_crashed(const float *rays, const double *edge, int num_rays,
                                  int poses, double crash_thresh)
{
    for (int i = 0; i < poses; ++i)
        for (int j = 0; j < num_rays; ++j)
            if (((double)rays[(int64_t)i * num_rays + j] - edge[j]) < crash_thresh) return i;
    return -(poses + 1);
}
