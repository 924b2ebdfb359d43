// car.cu -- the vehicle-model half of the fused rollout (north_star (c)) and the crash test that
// follows every scan in the reference:
//   Car::updatePosition / computeFromInput / updateNormal / updateSingle   racecar/src/racecar.cpp:53-237
//   Car::getScanPose                                                       racecar/src/racecar.cpp:378-387
//   Car::setCarEdgeDistances                                               racecar/src/racecar.cpp:239-292
//   Car::isCrashed                                                         racecar/src/racecar.cpp:305-328
//   MCTS.rollout (steps, action every 10th step, checkCollisionMany)       scripts/mcts.py:202-245
// State layout is the reference's 11 doubles (racecar.cpp:330-376).  Everything is fp64 like the
// reference; compiled with -fmad=false so products and sums round separately, as they do in the
// reference built without -ffast-math (the oracle's build).  Device cos/sin/tan are CUDA's, which
// may differ from the host libm in the last ulp: state parity is to a stated tolerance, the crash
// test itself is exact on identical ranges.
#include <cmath>
#include <new>
#include <vector>

#include "glibc_trig.cuh"
#include "march.cuh"
#include "marcher.h"

struct CarParams {
    double wb, fc, h_cg, l_f, l_r, cs_f, cs_r, mass, i_z, crash_thresh, width, length;
    double max_steer_vel, max_steer_ang, max_speed, max_accel, max_decel;
};

struct rl_car {
    CarParams p{};
    int device = 0;
    int num_rays = 0;
    std::vector<double> edge;   // host copy (computed with the host libm, like the reference)
    double *d_edge = nullptr;
};

namespace {

constexpr double K_THRESH = 0.5, ST_THRESH = 0.53, GRAV = 9.81;
constexpr double REF_PI = 3.145;  // sic: racecar/include/racecar.hpp:117
constexpr uint32_t NO_CRASH = 0xffffffffu;   // what cudaMemsetAsync(.., 0xFF, ..) leaves: the identity of the unsigned atomicMin

__device__ __forceinline__ double clampd(double v, double lo, double hi) { return fmin(fmax(v, lo), hi); }

// One Car::updatePosition(dt) on a register-resident state.
struct CarState { double x, y, th, v, sa, w, beta, travel, total_v; int dyn, count; };

__device__ __forceinline__ void car_update(const CarParams &p, CarState &s, double in_speed,
                                           double in_steer, double dt)
{
    // computeFromInput (racecar.cpp:118-169)
    const double kp = 2.0 * p.max_accel / p.max_speed;
    const double dv = in_speed - s.v;
    double accel;
    if (s.v > 0) accel = (dv > 0) ? clampd(kp * dv, -p.max_accel, p.max_accel) : -p.max_decel;
    else accel = (dv > 0) ? p.max_decel : clampd(kp * dv, -p.max_accel, p.max_accel);
    const double ds = in_steer - s.sa;
    double sv = 0.0;
    if (fabs(ds) > 0.0001) sv = (ds > 0) ? p.max_steer_vel : -p.max_steer_vel;

    const double px = s.x, py = s.y;
    const double thresh = s.dyn ? ST_THRESH : K_THRESH;
    if (s.v < thresh) {   // updateNormal (racecar.cpp:171-194)
        const double xd = s.v * cos(s.th), yd = s.v * sin(s.th);
        const double thd = s.v / p.wb * tan(s.sa);
        s.x += xd * dt; s.y += yd * dt; s.th += thd * dt;
        s.v += accel * dt; s.sa += sv * dt;
        s.w = 0; s.beta = 0; s.dyn = 0;
    } else {              // updateSingle (racecar.cpp:196-237)
        const double xd = s.v * cos(s.th + s.beta), yd = s.v * sin(s.th + s.beta);
        const double thd = s.w;
        const double rv = GRAV * p.l_r - accel * p.h_cg;
        const double fv = GRAV * p.l_f + accel * p.h_cg;
        const double ratio = s.w / s.v;
        const double first = p.fc / (s.v * (p.l_r + p.l_f));
        const double wdd = (p.fc * p.mass / (p.i_z * p.wb)) *
                           (p.l_f * p.cs_f * s.sa * rv + s.beta * (p.l_r * p.cs_r * fv - p.l_f * p.cs_f * rv) -
                            ratio * ((p.l_f * p.l_f) * p.cs_f * rv + (p.l_r * p.l_r) * p.cs_r * fv));
        const double bd = first * (p.cs_f * s.sa * (rv) - s.beta * (p.cs_r * fv + p.cs_f * rv) +
                                   ratio * (p.cs_r * p.l_r * fv - p.cs_f * p.l_f * rv)) - s.w;
        s.x += xd * dt; s.y += yd * dt; s.th += thd * dt;
        s.v += accel * dt; s.sa += sv * dt;
        s.w += wdd * dt; s.beta += bd * dt; s.dyn = 1;
    }
    const double ddx = px - s.x, ddy = py - s.y;
    s.travel += sqrt(ddx * ddx + ddy * ddy);
    s.total_v += s.v;
    s.count += 1;
    s.v = clampd(s.v, -p.max_speed, p.max_speed);
    s.sa = clampd(s.sa, -p.max_steer_ang, p.max_steer_ang);
}

__device__ __forceinline__ CarState load_state(const double *st)
{
    CarState s;
    s.x = st[0]; s.y = st[1]; s.th = st[2]; s.v = st[3]; s.sa = st[4]; s.w = st[5]; s.beta = st[6];
    s.dyn = st[7] > 0.0; s.travel = st[8]; s.total_v = st[9]; s.count = (int)st[10];
    return s;
}

__device__ __forceinline__ void store_state(double *st, const CarState &s)
{
    st[0] = s.x; st[1] = s.y; st[2] = s.th; st[3] = s.v; st[4] = s.sa; st[5] = s.w; st[6] = s.beta;
    st[7] = s.dyn ? 1.0 : 0.0; st[8] = s.travel; st[9] = s.total_v; st[10] = (double)s.count;
}

// T updatePosition steps per car; a new (speed, steer) target every `action_every` steps
// (scripts/mcts.py:216-222).  Records the pose scanned after every step, narrowed to fp32 exactly
// where the reference narrows it (the f32 pose buffer, scripts/mcts.py:211, :229-231), and the
// running sum of the post-step velocities (the rollout reward, scripts/mcts.py:235).
__global__ void __launch_bounds__(128)
car_rollout_kernel(CarParams p, double *__restrict__ states, const double *__restrict__ actions,
                   int64_t n_cars, int steps, int action_every, double dt, int lidar_pose,
                   double scan_dist, float *__restrict__ poses, double *__restrict__ vsum,
                   uint32_t *__restrict__ first)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cars) return;
    first[c] = NO_CRASH;   // the scan kernel that follows takes the minimum crashed step into it
    CarState s = load_state(states + 11 * c);
    const int n_actions = (steps + action_every - 1) / action_every;
    const double *act = actions + 2 * n_actions * c;
    double speed = 0.0, steer = 0.0, acc = 0.0;
    for (int i = 0; i < steps; ++i) {
        if (i % action_every == 0) { speed = act[2 * (i / action_every)]; steer = act[2 * (i / action_every) + 1]; }
        car_update(p, s, speed, steer, dt);
        float *o = poses + 3 * ((int64_t)i * n_cars + c);   // step-major: all cars' step i are contiguous
        if (lidar_pose) {   // Car::getScanPose
            o[0] = (float)(s.x + scan_dist * cos(s.th));
            o[1] = (float)(s.y + scan_dist * sin(s.th));
        } else {            // base link, as MCTS.rollout records it
            o[0] = (float)s.x;
            o[1] = (float)s.y;
        }
        o[2] = (float)s.th;
        acc += s.v;
        vsum[c * steps + i] = acc;
    }
    store_state(states + 11 * c, s);
}

// Batched Car::control + Car::updatePosition: one step, per-car targets.
__global__ void __launch_bounds__(128)
car_step_kernel(CarParams p, double *__restrict__ states, const double *__restrict__ speed,
                const double *__restrict__ steer, int64_t n_cars, double dt)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cars) return;
    CarState s = load_state(states + 11 * c);
    car_update(p, s, speed[c], steer[c], dt);
    store_state(states + 11 * c, s);
}

// Rollout action schedule generated on the device: MCTS.rollout draws
// rand_steer = uniform(-max_steer_ang, max_steer_ang) and then rand_speed = uniform(0, max_speed) on
// every action_every-th step (scripts/mcts.py:216-222).  Here every (car, action) pair owns one
// Philox4x32-10 block -- counter (action, car_lo, car_hi, stream_id), key (seed_lo, seed_hi) -- so the
// schedule is reproducible whatever the launch shape and identical to oracle/philox_oracle.c: words
// 0,1 make the steer variate, words 2,3 the speed variate, each as the 53-bit double
// ((a >> 5) * 2^26 + (b >> 6)) / 2^53 scaled the way numpy's uniform() does, lo + (hi - lo) * u.
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1, uint32_t out[4])
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__device__ __forceinline__ double unit_double(uint32_t a, uint32_t b)
{
    return __dmul_rn(__dadd_rn(__dmul_rn((double)(a >> 5), 67108864.0), (double)(b >> 6)), 1.0 / 9007199254740992.0);
}

__global__ void __launch_bounds__(256)
rollout_actions_kernel(double *__restrict__ actions, int64_t n_cars, int n_actions, uint64_t seed,
                       uint32_t stream_id, int64_t car_offset, double speed_lo, double speed_hi, double steer_lo, double steer_hi)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_cars * n_actions) return;
    const int64_t c = i / n_actions;
    const uint32_t a = (uint32_t)(i - c * n_actions);
    const uint64_t gc = (uint64_t)(c + car_offset);   // global car index
    uint32_t w[4];
    philox4x32_10(a, (uint32_t)(gc & 0xffffffffu), (uint32_t)(gc >> 32), stream_id,
                  (uint32_t)(seed & 0xffffffffu), (uint32_t)(seed >> 32), w);
    const double steer = __dadd_rn(steer_lo, __dmul_rn(__dsub_rn(steer_hi, steer_lo), unit_double(w[0], w[1])));
    const double speed = __dadd_rn(speed_lo, __dmul_rn(__dsub_rn(speed_hi, speed_lo), unit_double(w[2], w[3])));
    actions[2 * i] = speed;
    actions[2 * i + 1] = steer;
}

// first[g] : NO_CRASH -> -(poses_per_group + 1), Car::isCrashed's "no crash" value
__global__ void finalize_first_kernel(int32_t *first, int64_t groups, int poses_per_group,
                                      const double *__restrict__ vsum, double *__restrict__ reward)
{
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= groups) return;
    int32_t f = first[g];
    if ((uint32_t)f == NO_CRASH) f = -(poses_per_group + 1);
    first[g] = f;
    if (reward) {   // sum(rewards[:index]) if index >= 0 else sum(rewards)   (scripts/mcts.py:240-245)
        const int upto = f < 0 ? poses_per_group : f;
        reward[g] = upto > 0 ? vsum[g * poses_per_group + upto - 1] : 0.0;
    }
}

// The value MCTS.rollout returns for a node (scripts/mcts.py:240-245): sum(rewards[:index]) / abs(node.action).
__global__ void rollout_value_kernel(const double *__restrict__ reward, const double *__restrict__ node_action,
                                     int64_t n, double *__restrict__ value)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c < n) value[c] = __ddiv_rn(reward[c], fabs(node_action[c]));
}

// Car::isCrashed over ranges that already exist: ray i of pose k of group g.
__global__ void __launch_bounds__(256)
crash_from_rays_kernel(const float *__restrict__ rays, const double *__restrict__ edge, int64_t total,
                       int num_rays, int poses_per_group, double thresh, uint32_t *first)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int64_t k = i / num_rays;
    const int j = (int)(i - k * num_rays);
    if (((double)rays[i] - edge[j]) < thresh) {
        const int64_t g = k / poses_per_group;
        atomicMin(first + g, (uint32_t)(k - g * poses_per_group));
    }
}

// Fan march with the crash test as its epilogue.  WRITE: also store the ranges.  Without WRITE a
// ray whose group already crashed at an earlier pose is skipped (it cannot change the minimum).
// Pose order: group-major (pose k = g*poses_per_group + p, the scanMany layout) or step-major
// (k = p*groups + g, used by the rollout so that a car's earlier steps are scanned -- and its crash
// known -- long before its later steps are scheduled).
// Launch shape: the same 128-thread CTAs, multiply-high beam index and L2 access-policy window as
// march_pose_kernel (profiles/r01_tuning.md sections 2 and 4).  The ray index is split so that it never
// needs a 64-bit divide: blockIdx.y/z select the OUTER unit (the step when step-major, the group when
// group-major), blockIdx.x * 128 + thread runs over the inner_poses * num_rays rays of that unit.
template <bool WRITE, bool STEP_MAJOR, bool PADDED>
__global__ void __launch_bounds__(rl::MARCH_CTA_THREADS)
march_crash_kernel(rl::MarchParams P, const float *__restrict__ poses, const double *__restrict__ edge,
                   uint32_t inner_rays, int num_rays, rl::FastDiv div, int64_t inner_poses, int64_t outer_count,
                   float fov, float inc, double thresh, uint32_t *first, float *__restrict__ outs)
{
    asm volatile("griddepcontrol.launch_dependents;");
    const uint32_t idx = blockIdx.x * rl::MARCH_CTA_THREADS + threadIdx.x;
    const int64_t outer = (int64_t)blockIdx.z * gridDim.y + blockIdx.y;
    if (idx >= inner_rays || outer >= outer_count) return;
    const uint32_t q = num_rays >= 2 ? rl::fast_div(idx, div) : idx;   // pose inside the outer unit
    const int j = (int)(idx - q * (uint32_t)num_rays);
    const int64_t k = outer * inner_poses + q;
    const int64_t g = STEP_MAJOR ? (int64_t)q : outer;
    const uint32_t pose_in_group = STEP_MAJOR ? (uint32_t)outer : q;
    if (!WRITE && __ldcg(first + g) < pose_in_group) return;
    const float *p = poses + 3 * k;
    const float thw = __ldg(p + 2);
    const rl::GridPose gp = rl::world_to_grid(P.w, __ldg(p), __ldg(p + 1), thw);
    const float thg = __fadd_rn(-__fadd_rn(thw, fmaf((float)j, inc, -0.5f * fov)), P.w.rotation_const);
    const rl::FirstSample f0 = rl::first_sample(P, gp.y, gp.x);
    float s, c;
    rl::glibc_sincosf(thg, &s, &c);
    uint32_t steps = 0;
    const float r = __fmul_rn(rl::march_ray<false, PADDED>(P, gp.y, gp.x, c, s, steps, f0), P.w.scale);
    if (WRITE) outs[k * num_rays + j] = r;
    if (((double)r - __ldg(edge + j)) < thresh) atomicMin(first + g, pose_in_group);
}

// Car::setCarEdgeDistances on the host, with the host libm like the reference.
void edge_distances(const CarParams &p, int num_rays, double min_ang, double inc, double scan_dist_to_base,
                    std::vector<double> &edge)
{
    edge.assign(num_rays, 0.0);
    const double side = p.width / 2.0;
    const double front = p.wb - scan_dist_to_base;
    const double back = scan_dist_to_base;
    double a = min_ang;
    for (int i = 0; i < num_rays; ++i) {
        a += inc;   // incremented BEFORE use: edge[i] belongs to min_ang + (i+1)*inc
        if (a > 0.0) {
            if (a < REF_PI / 2.0) edge[i] = std::fmin(side / std::sin(a), front / std::cos(a));
            else edge[i] = std::fmin(side / std::sin(a - REF_PI / 2.0), back / std::cos(a - REF_PI / 2.0));
        } else {
            if (a == 0.0) a += 0.0001;
            if (a > -REF_PI / 2.0) edge[i] = std::fmin(side / std::sin(-a), front / std::cos(-a));
            else edge[i] = std::fmin(side / std::sin(-a - REF_PI / 2.0), back / std::cos(-a - REF_PI / 2.0));
        }
    }
}

inline unsigned blocks_for(int64_t n, int bs) { return (unsigned)((n + bs - 1) / bs); }

rl::FastDiv fast_div_for(int d)
{
    rl::FastDiv f{0, 0, (uint32_t)d};
    if (d >= 2) {
        int s = 1;
        while ((1u << s) < (uint32_t)d) ++s;
        f.magic = (uint32_t)((((uint64_t)1 << (31 + s)) + d - 1) / d);
        f.shift = (uint32_t)(s - 1);
    }
    return f;
}

// outer_count units of inner_poses poses each (see march_crash_kernel)
template <bool WRITE, bool STEP_MAJOR>
int32_t launch_march_crash(rl_marcher *m, const rl_car *car, const float *d_poses, int64_t inner_poses,
                           int64_t outer_count, float fov, uint32_t *d_first, float *d_ranges, cudaStream_t s,
                           const char *who)
{
    const int64_t inner_rays = inner_poses * car->num_rays;
    if (inner_rays >= ((int64_t)1 << 31) || outer_count > (int64_t)65535 * 65535)
        return rl::fail(RL_ERR_BAD_ARG, std::string(who) + ": too many rays for one call");
    const unsigned gy = (unsigned)(outer_count < 65535 ? outer_count : 65535);
    const unsigned gz = (unsigned)((outer_count + gy - 1) / gy);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(blocks_for(inner_rays, rl::MARCH_CTA_THREADS), gy, gz);
    cfg.blockDim = dim3(rl::MARCH_CTA_THREADS);
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    if (m->l2_window_bytes) {   // keep the distance field pinned in L2, as every march launch does
        attr[0].id = cudaLaunchAttributeAccessPolicyWindow;
        attr[0].val.accessPolicyWindow.base_ptr = const_cast<float *>(m->d_field ? m->d_field : m->P.dist);
        attr[0].val.accessPolicyWindow.num_bytes = m->l2_window_bytes;
        attr[0].val.accessPolicyWindow.hitRatio = m->l2_hit_ratio;
        attr[0].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr[0].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
    }
    if (m->P.pad > 0)
        RL_CUDA(cudaLaunchKernelEx(&cfg, march_crash_kernel<WRITE, STEP_MAJOR, true>, m->P, d_poses, (const double *)car->d_edge,
                                   (uint32_t)inner_rays, car->num_rays, fast_div_for(car->num_rays), inner_poses, outer_count,
                                   fov, fov / (float)car->num_rays, car->p.crash_thresh, d_first, d_ranges));
    else
        RL_CUDA(cudaLaunchKernelEx(&cfg, march_crash_kernel<WRITE, STEP_MAJOR, false>, m->P, d_poses, (const double *)car->d_edge,
                                   (uint32_t)inner_rays, car->num_rays, fast_div_for(car->num_rays), inner_poses, outer_count,
                                   fov, fov / (float)car->num_rays, car->p.crash_thresh, d_first, d_ranges));
    return RL_OK;
}

int32_t check_car(const rl_car *car, const char *who)
{
    if (!car) return rl::fail(RL_ERR_BAD_ARG, std::string(who) + ": null car");
    if (car->num_rays <= 0 || !car->d_edge)
        return rl::fail(RL_ERR_BAD_ARG, std::string(who) + ": call rl_car_set_edge_distances first");
    return RL_OK;
}

}  // namespace

extern "C" {

RL_API int32_t rl_car_create(const double *params17, int32_t device, rl_car **out)
{
    if (!params17 || !out) return rl::fail(RL_ERR_BAD_ARG, "rl_car_create: null pointer");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return rl::fail(RL_ERR_NO_DEVICE, "rl_car_create: no CUDA device (this library has no CPU path)");
    if (device < 0 || device >= ndev) return rl::fail(RL_ERR_NO_DEVICE, "rl_car_create: bad device index");
    rl_car *c = new (std::nothrow) rl_car();
    if (!c) return rl::fail(RL_ERR_OOM, "rl_car_create: host allocation failed");
    const double *q = params17;
    c->p = CarParams{q[0], q[1], q[2], q[3], q[4], q[5], q[6], q[7], q[8], q[9], q[10], q[11], q[12], q[13], q[14], q[15], q[16]};
    c->device = device;
    *out = c;
    return RL_OK;
}

RL_API int32_t rl_car_destroy(rl_car *car)
{
    if (!car) return rl::fail(RL_ERR_BAD_ARG, "rl_car_destroy: null car");
    {
        rl::DeviceGuard guard(car->device);
        cudaFree(car->d_edge);
    }
    delete car;
    return RL_OK;
}

RL_API int32_t rl_car_set_edge_distances(rl_car *car, int32_t num_rays, double min_ang, double ang_inc,
                                         double scan_dist_to_base)
{
    if (!car || num_rays <= 0) return rl::fail(RL_ERR_BAD_ARG, "rl_car_set_edge_distances: bad argument");
    rl::DeviceGuard guard(car->device);
    edge_distances(car->p, num_rays, min_ang, ang_inc, scan_dist_to_base, car->edge);
    cudaFree(car->d_edge);
    car->d_edge = nullptr;
    car->num_rays = 0;
    RL_CUDA(cudaMalloc(&car->d_edge, (size_t)num_rays * sizeof(double)));
    RL_CUDA(cudaMemcpy(car->d_edge, car->edge.data(), (size_t)num_rays * sizeof(double), cudaMemcpyHostToDevice));
    car->num_rays = num_rays;
    return RL_OK;
}

RL_API int32_t rl_car_get_edge_distances(const rl_car *car, double *out, int32_t num_rays)
{
    if (!car || !out || num_rays != car->num_rays) return rl::fail(RL_ERR_BAD_ARG, "rl_car_get_edge_distances: bad argument");
    for (int i = 0; i < num_rays; ++i) out[i] = car->edge[i];
    return RL_OK;
}

// Batched control() + updatePosition(dt): d_states (n,11) fp64 in place, per-car targets.
RL_API int32_t rl_car_step(rl_car *car, double *d_states, const double *d_speed, const double *d_steer,
                           int64_t n_cars, double dt, void *stream)
{
    if (!car || n_cars < 0 || (n_cars > 0 && (!d_states || !d_speed || !d_steer)))
        return rl::fail(RL_ERR_BAD_ARG, "rl_car_step: bad argument");
    if (n_cars == 0) return RL_OK;
    rl::DeviceGuard guard(car->device);
    car_step_kernel<<<blocks_for(n_cars, 128), 128, 0, (cudaStream_t)stream>>>(car->p, d_states, d_speed, d_steer, n_cars, dt);
    RL_CUDA(cudaGetLastError());
    return RL_OK;
}

// Car::isCrashed over device ranges: d_first[g] = first crashed pose of group g or -(poses_per_group+1).
RL_API int32_t rl_is_crashed(rl_car *car, const float *d_rays, int64_t groups, int32_t poses_per_group,
                             int32_t *d_first, void *stream)
{
    int32_t rc = check_car(car, "rl_is_crashed");
    if (rc != RL_OK) return rc;
    if (groups < 0 || poses_per_group <= 0 || (groups > 0 && (!d_rays || !d_first)))
        return rl::fail(RL_ERR_BAD_ARG, "rl_is_crashed: bad argument");
    if (groups == 0) return RL_OK;
    rl::DeviceGuard guard(car->device);
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t total = groups * poses_per_group * car->num_rays;
    RL_CUDA(cudaMemsetAsync(d_first, 0xFF, (size_t)groups * sizeof(int32_t), s));   // NO_CRASH
    crash_from_rays_kernel<<<blocks_for(total, 256), 256, 0, s>>>(d_rays, car->d_edge, total, car->num_rays,
                                                                 poses_per_group, car->p.crash_thresh,
                                                                 reinterpret_cast<uint32_t *>(d_first));
    finalize_first_kernel<<<blocks_for(groups, 256), 256, 0, s>>>(d_first, groups, poses_per_group, nullptr, nullptr);
    RL_CUDA(cudaGetLastError());
    return RL_OK;
}

// scanMany + isCrashed in one pass (RacecarSimulator.checkCollisionMany): d_poses (groups *
// poses_per_group, 3) fp32; beams = the car's edge table length; d_ranges may be NULL, in which
// case no range is ever written and poses after a group's first crash are skipped.
RL_API int32_t rl_scan_crash(rl_marcher *m, rl_car *car, const float *d_poses, int64_t groups,
                             int32_t poses_per_group, float fov, int32_t *d_first, float *d_ranges,
                             void *stream)
{
    int32_t rc = check_car(car, "rl_scan_crash");
    if (rc != RL_OK) return rc;
    if (!m || groups < 0 || poses_per_group <= 0 || (groups > 0 && (!d_poses || !d_first)))
        return rl::fail(RL_ERR_BAD_ARG, "rl_scan_crash: bad argument");
    if (m->map->device != car->device) return rl::fail(RL_ERR_BAD_ARG, "rl_scan_crash: car and marcher are on different devices");
    if (groups == 0) return RL_OK;
    rl::DeviceGuard guard(car->device);
    cudaStream_t s = (cudaStream_t)stream;
    RL_CUDA(cudaMemsetAsync(d_first, 0xFF, (size_t)groups * sizeof(int32_t), s));   // NO_CRASH
    uint32_t *first = reinterpret_cast<uint32_t *>(d_first);
    if (d_ranges) rc = launch_march_crash<true, false>(m, car, d_poses, poses_per_group, groups, fov, first, d_ranges, s, "rl_scan_crash");
    else rc = launch_march_crash<false, false>(m, car, d_poses, poses_per_group, groups, fov, first, nullptr, s, "rl_scan_crash");
    if (rc != RL_OK) return rc;
    finalize_first_kernel<<<blocks_for(groups, 256), 256, 0, s>>>(d_first, groups, poses_per_group, nullptr, nullptr);
    RL_CUDA(cudaGetLastError());
    return RL_OK;
}

// d_value[c] = d_reward[c] / |d_node_action[c]|: what MCTS.rollout returns for the node whose action was
// node_action (scripts/mcts.py:240-245; a zero action gives +-inf or nan exactly as numpy's division does).
RL_API int32_t rl_rollout_value(const double *d_reward, const double *d_node_action, int64_t n_cars,
                                double *d_value, int32_t device, void *stream)
{
    if (n_cars < 0 || (n_cars > 0 && (!d_reward || !d_node_action || !d_value)))
        return rl::fail(RL_ERR_BAD_ARG, "rl_rollout_value: bad argument");
    if (n_cars == 0) return RL_OK;
    rl::DeviceGuard guard(device);
    if (!guard.ok) return rl::fail(RL_ERR_NO_DEVICE, "rl_rollout_value: bad device index");
    rollout_value_kernel<<<blocks_for(n_cars, 256), 256, 0, (cudaStream_t)stream>>>(d_reward, d_node_action, n_cars, d_value);
    RL_CUDA(cudaGetLastError());
    return RL_OK;
}

// The action schedule of MCTS.rollout (scripts/mcts.py:216-222) for n_cars cars, drawn on the device:
// d_actions (n_cars, n_actions, 2) fp64 = (speed in [speed_lo, speed_hi), steer in [steer_lo, steer_hi)).
RL_API int32_t rl_rollout_actions(double *d_actions, int64_t n_cars, int32_t n_actions, uint64_t seed,
                                  uint32_t stream_id, int64_t car_offset, double speed_lo, double speed_hi,
                                  double steer_lo, double steer_hi, int32_t device, void *stream)
{
    if (n_cars < 0 || n_actions <= 0 || car_offset < 0 || (n_cars > 0 && !d_actions))
        return rl::fail(RL_ERR_BAD_ARG, "rl_rollout_actions: bad argument");
    if (n_cars == 0) return RL_OK;
    const int64_t total = n_cars * n_actions;
    if ((total + 255) / 256 > 0x7fffffffLL) return rl::fail(RL_ERR_BAD_ARG, "rl_rollout_actions: too many actions for one call");
    rl::DeviceGuard guard(device);
    if (!guard.ok) return rl::fail(RL_ERR_NO_DEVICE, "rl_rollout_actions: bad device index");
    rollout_actions_kernel<<<blocks_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
        d_actions, n_cars, n_actions, seed, stream_id, car_offset, speed_lo, speed_hi, steer_lo, steer_hi);
    RL_CUDA(cudaGetLastError());
    return RL_OK;
}

// MCTS.rollout for n_cars cars at once, nothing leaving the GPU: `steps` updatePosition(dt) per car
// with a new (speed, steer) from d_actions (n_cars, ceil(steps/action_every), 2) every
// action_every-th step, one fan scan per step from the base-link pose (lidar_pose = 0, what
// scripts/mcts.py:228-231 records) or the lidar pose (lidar_pose = 1, Car::getScanPose), crash test
// per scan.  Outputs: d_states updated in place; d_crash_index[c] = first crashed step or
// -(steps+1); d_reward[c] = sum of post-step velocities before the crash (all steps if none).
// d_poses (steps, n_cars, 3) fp32 -- step-major -- and d_vsum (n_cars, steps) fp64 are caller-provided
// scratch that also serve as outputs (the scanned poses, the running velocity sums).
RL_API int32_t rl_rollout(rl_marcher *m, rl_car *car, double *d_states, const double *d_actions,
                          int64_t n_cars, int32_t steps, int32_t action_every, double dt,
                          int32_t lidar_pose, double scan_dist_to_base, float fov,
                          int32_t *d_crash_index, double *d_reward, float *d_poses, double *d_vsum,
                          void *stream)
{
    int32_t rc = check_car(car, "rl_rollout");
    if (rc != RL_OK) return rc;
    if (!m || n_cars < 0 || steps <= 0 || action_every <= 0 ||
        (n_cars > 0 && (!d_states || !d_actions || !d_crash_index || !d_poses || !d_vsum)))
        return rl::fail(RL_ERR_BAD_ARG, "rl_rollout: bad argument");
    if (m->map->device != car->device) return rl::fail(RL_ERR_BAD_ARG, "rl_rollout: car and marcher are on different devices");
    if (n_cars == 0) return RL_OK;
    rl::DeviceGuard guard(car->device);
    cudaStream_t s = (cudaStream_t)stream;
    // three launches: the vehicle steps (which also reset the crash indices), the scans with the crash test
    // as their epilogue, and a 65 536-thread decode of "no crash" + reward.  (Folding the decode into the
    // scan kernel's last CTA would cost one same-address atomic per CTA -- 27.6 M of them at BASELINE config 4.)
    uint32_t *first = reinterpret_cast<uint32_t *>(d_crash_index);
    car_rollout_kernel<<<blocks_for(n_cars, 128), 128, 0, s>>>(car->p, d_states, d_actions, n_cars, steps, action_every, dt,
                                                              lidar_pose, scan_dist_to_base, d_poses, d_vsum, first);
    RL_CUDA(cudaGetLastError());
    rc = launch_march_crash<false, true>(m, car, d_poses, n_cars, steps, fov, first, nullptr, s, "rl_rollout");
    if (rc != RL_OK) return rc;
    finalize_first_kernel<<<blocks_for(n_cars, 256), 256, 0, s>>>(d_crash_index, n_cars, steps, d_vsum, d_reward);
    RL_CUDA(cudaGetLastError());
    return RL_OK;
}

}  // extern "C"
