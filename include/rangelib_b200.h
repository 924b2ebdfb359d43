/*
 * rangelib_b200.h -- C ABI of the B200-native batched lidar scan path.
 *
 * Drop-in boundary for the scan path of felrock/PyRacecarSimulator.  Every entry point
 * replaces one interface of the external `range_libc` extension as the reference binds
 * it (all citations relative to the reference checkout):
 *
 *   range_libc.PyOMap(map_msg)                 scripts/ros_interface.py:210,
 *                                              scripts/mcts_driver.py:278
 *       -> rl_map_from_occupancy / rl_map_from_image / rl_map_from_cells
 *   range_libc.PyRayMarching(omap, mrx)        scripts/scan_simulator.py:72-73
 *   range_libc.PyRayMarchingGPU(omap, mrx)     scripts/scan_simulator.py:75-76
 *       -> rl_marcher_create          (the distance transform upstream builds on the host
 *                                      inside these constructors is built on the GPU by
 *                                      rl_map_from_*)
 *   .calc_range_many(ins, outs)                scripts/two_player/scan.py:69-70
 *       -> rl_calc_range_many[_host]
 *   .calc_range_many(ins, outs, fov, num_rays) scripts/scan_simulator.py:103-106, :130-133
 *       -> rl_calc_range_fan[_host]   (pose_stride_rows = num_rays is the fork's layout)
 *   .calc_range_repeat_angles(ins, angles, outs)   upstream RangeLibc.pyx (north_star)
 *       -> rl_calc_range_repeat_angles[_host]
 *   racecar.PyCar.isCrashed(scans, num_rays, poses) after scanMany
 *                                              scripts/racecar_simulator_v2.py:146-167,
 *                                              racecar/src/racecar.cpp:305-328
 *       -> rl_scan_crash              (fan march with the crash test as its epilogue)
 *   MCTS.rollout (bicycle steps + checkCollisionMany)   scripts/mcts.py:202-245,
 *                                              racecar/src/racecar.cpp:53-237
 *       -> rl_rollout                 (fused step + scan + crash, north_star (c))
 *
 * Conventions
 *   - Plain C: pointers and sizes only.  Functions return RL_OK (0) or a negative
 *     rl_status; rl_last_error() returns a thread-local message.  Nothing here ever
 *     exits the process or falls back to the CPU.
 *   - `d_` pointers are DEVICE pointers on the map's device; the call is enqueued on
 *     `stream` (a cudaStream_t passed as void*, NULL = the legacy default stream) and
 *     returns without synchronising.  `_host` variants take HOST pointers, stage through
 *     pinned memory owned by the marcher, and return when `outs` is filled.
 *   - poses are (x, y, theta) fp32 triples in the map.yaml world frame (metres, radians);
 *     ranges are fp32 metres.  max_range is in pixels (scripts/racecar_simulator_v2.py:196).
 *   - A map is immutable once built and may be shared by any number of marchers and
 *     threads; a marcher serialises its own `_host` calls with an internal mutex and is
 *     otherwise stateless, so concurrent device-pointer calls on different streams are safe.
 */
#ifndef RANGELIB_B200_H
#define RANGELIB_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RL_ABI_VERSION 1

#if defined(__GNUC__)
#define RL_API __attribute__((visibility("default")))
#else
#define RL_API
#endif

typedef enum rl_status {
    RL_OK = 0,
    RL_ERR_BAD_ARG = -1,   /* null pointer, non-positive size, unsupported size */
    RL_ERR_CUDA = -2,      /* a CUDA runtime call failed; see rl_last_error()   */
    RL_ERR_NO_DEVICE = -3, /* no CUDA device / device index out of range        */
    RL_ERR_OOM = -4        /* host or device allocation failed                  */
} rl_status;

/* marcher flags.  The arithmetic has one mode, the bit-exact one (fp32 distance field): a narrower  */
/* (fp16 / u16) field was measured and rejected, see DESIGN.md section 3.                            */
#define RL_FLAG_DEFAULT 0u
/* Do not pin the distance field in L2: by default rl_marcher_create raises the device-wide          */
/* persisting-L2 carve-out (cudaLimitPersistingL2CacheSize) to the size of the field and every march */
/* launch carries an access-policy window over it; rl_marcher_destroy of the last such marcher on a  */
/* device un-pins the lines and restores the previous limit.  An embedding application that manages  */
/* the carve-out itself passes this flag.                                                            */
#define RL_FLAG_NO_L2_WINDOW 1u
/* Do not build the marcher's NaN-padded copy of the march field (it costs (rows + 2p)(cols + 2p) floats with  */
/* p = ceil(max_range_px) + 16 and removes the per-step bounds test); the bounds-tested kernels are used,      */
/* as they are automatically for max_range_px > 2048.  Results are identical either way.                     */
#define RL_FLAG_NO_PADDED_FIELD 2u
/* Large batches of poses are MARCHED in map order by SM territories: one counting sort of the pose indices by  */
/* Morton cell per call (scratch from a pool the marcher owns), then every SM works through one contiguous range */
/* of that order, so neighbouring poses share their field cells in its L1.  Applied when the poses are dense     */
/* (>= one per 16 map cells and >= 24 M rays) or the field is larger than L2 (>= 16 384 poses and >= 16 M rays). */
/* Poses are read and ranges written at the caller's indices; results are identical.  This flag keeps the       */
/* caller's order.                                                                                               */
#define RL_FLAG_NO_POSE_SORT 4u

/* dist2 value of a cell from which no occupied cell is reachable (empty map) */
#define RL_DIST2_INF 0x3fffffff

typedef struct rl_map rl_map;
typedef struct rl_marcher rl_marcher;
typedef struct rl_car rl_car;

RL_API int32_t rl_abi_version(void);
RL_API const char *rl_last_error(void);
RL_API int32_t rl_device_count(int32_t *count);

/* ---- map ingest (GPU: threshold -> occupancy -> exact integer squared EDT -> sqrt) ---- */

/* From image pixels as stored in the PGM (rows top to bottom, `width` columns), applying  */
/* map_server's thresholds (mode: trinary is what every shipped map.yaml uses) and y-flip,  */
/* then (binarise != 0) the reference's `>0 -> 255 else 0` (scripts/ros_interface.py:80-86), */
/* then PyOMap's `> 10` cut.                                                               */
#define RL_MAP_TRINARY 0
#define RL_MAP_SCALE 1
#define RL_MAP_RAW 2
RL_API int32_t rl_map_from_image(const uint8_t *pixels, int32_t width, int32_t height, int32_t negate,
                          double occupied_thresh, double free_thresh, int32_t mode, int32_t binarise,
                          double resolution, double origin_x, double origin_y, double origin_yaw,
                          int32_t device, rl_map **out);

/* Colour / alpha images (launch/simulate.launch:8-9 hands map_server any image it can load): `channels`   */
/* interleaved bytes per pixel (1 grey, 2 grey+alpha, 3 RGB, 4 RGBA).  map_server averages the channels of  */
/* a pixel -- all of them in trinary mode or when the image has no alpha (`has_alpha` = 0), all but the      */
/* last otherwise -- then applies the same rules to the average; in scale mode a pixel whose LAST byte is   */
/* 0 and whose shade lies between the thresholds becomes unknown (ROS map_server image_loader.cpp).         */
RL_API int32_t rl_map_from_image_channels(const uint8_t *pixels, int32_t width, int32_t height, int32_t channels,
                                          int32_t has_alpha, int32_t negate, double occupied_thresh,
                                          double free_thresh, int32_t mode, int32_t binarise, double resolution,
                                          double origin_x, double origin_y, double origin_yaw, int32_t device,
                                          rl_map **out);

/* From OccupancyGrid.data (int8, row-major from the bottom-left cell, width = columns). */
RL_API int32_t rl_map_from_occupancy(const int8_t *data, int32_t width, int32_t height, int32_t binarise,
                              double resolution, double origin_x, double origin_y,
                              double origin_yaw, int32_t device, rl_map **out);

/* From an already-cut boolean grid (non-zero = occupied), same cell order. */
RL_API int32_t rl_map_from_cells(const uint8_t *occupied, int32_t width, int32_t height,
                          double resolution, double origin_x, double origin_y, double origin_yaw,
                          int32_t device, rl_map **out);

RL_API int32_t rl_map_shape(const rl_map *map, int32_t *width, int32_t *height, int32_t *device);
/* Host copies for bit-exact checks: occupancy (uint8 0/1), squared distance (int32,      */
/* RL_DIST2_INF when unreachable) and distance (fp32 pixels), each width*height, row-major */
/* from the bottom-left cell.                                                              */
RL_API int32_t rl_map_get_occupancy(const rl_map *map, uint8_t *out);
RL_API int32_t rl_map_get_dist2(const rl_map *map, int32_t *out);
RL_API int32_t rl_map_get_dist(const rl_map *map, float *out);
/* Device pointer of the fp32 distance field (borrowed; valid until rl_map_destroy). */
RL_API int32_t rl_map_dist_device(const rl_map *map, const float **d_dist);
/* Device time of the last ingest on this map, milliseconds (CUDA events). */
RL_API int32_t rl_map_ingest_ms(const rl_map *map, float *ms);
RL_API int32_t rl_map_destroy(rl_map *map);

/* ---- marcher ---- */
RL_API int32_t rl_marcher_create(const rl_map *map, float max_range_px, uint32_t flags, rl_marcher **out);
RL_API int32_t rl_marcher_destroy(rl_marcher *m);

/* Pipelined launches.  A march has no data dependence on the march before it, but in stream order it    */
/* cannot start until that one has drained, and the drain -- a few hundred warps finishing 100-280-step  */
/* rays one dependent load at a time -- is a third of a launch (profiles/r01_timeline.md).  Callers that  */
/* issue scans back to back (scripts/mcts.py:118-122) can let consecutive device-pointer calls overlap:   */
/*   RL_PIPELINE_STREAMS  calls alternate between two streams owned by the marcher.  Each launch waits    */
/*       for everything enqueued on `stream` before the call (so its inputs are ready) but not for the    */
/*       previous march; `stream` is made to wait for the PREVIOUS call's completion only.  Contract: the */
/*       outputs of a call are valid in `stream` order after the NEXT march call on this marcher, or      */
/*       after rl_marcher_join(m, stream); a call must not read what the call before it writes.           */
/*   RL_PIPELINE_PDL      calls stay on `stream` and are launched with programmatic stream serialization: */
/*       a march may start once every CTA of the preceding march has STARTED.  Same contract, and the     */
/*       outputs of a call are valid after the second-next call or any non-march work on `stream`.        */
/* _host entry points and the *_allgather calls are not affected.  Not capturable into a CUDA graph      */
/* unless rl_marcher_join precedes the end of the capture.                                               */
#define RL_PIPELINE_OFF 0
#define RL_PIPELINE_STREAMS 1
#define RL_PIPELINE_PDL 2
RL_API int32_t rl_marcher_set_pipelined(rl_marcher *m, int32_t mode);
RL_API int32_t rl_marcher_join(rl_marcher *m, void *stream);

/* one (x, y, theta) row per ray: outs[i] = range(ins[i]) */
RL_API int32_t rl_calc_range_many(rl_marcher *m, const float *d_ins, float *d_outs, int64_t num_rays_total,
                           void *stream);
RL_API int32_t rl_calc_range_many_host(rl_marcher *m, const float *ins, float *outs, int64_t num_rays_total);

/* pose k at row k*pose_stride_rows of d_poses; beam j heads theta - fov/2 + j*fov/num_rays; */
/* outs[k*num_rays + j].  pose_stride_rows = 1: compact (B,3); = num_rays: the fork's layout.  */
RL_API int32_t rl_calc_range_fan(rl_marcher *m, const float *d_poses, int64_t pose_stride_rows,
                          float *d_outs, int64_t num_poses, int32_t num_rays, float fov,
                          void *stream);
RL_API int32_t rl_calc_range_fan_host(rl_marcher *m, const float *poses, int64_t pose_stride_rows,
                               float *outs, int64_t num_poses, int32_t num_rays, float fov);

/* outs[i*num_angles + a] = range(x_i, y_i, theta_i + angles[a]) */
RL_API int32_t rl_calc_range_repeat_angles(rl_marcher *m, const float *d_poses, const float *d_angles,
                                    float *d_outs, int64_t num_poses, int32_t num_angles,
                                    void *stream);
RL_API int32_t rl_calc_range_repeat_angles_host(rl_marcher *m, const float *poses, const float *angles,
                                         float *outs, int64_t num_poses, int32_t num_angles);

/* ---- fused march + all-gather over NVLink peer memory (one process per GPU) ---- */
/* rl_peer_alloc: device buffer + 64-byte CUDA IPC handle to hand to the other ranks;           */
/* rl_peer_open: map another rank's buffer into this process (peer access enabled lazily).      */
RL_API int32_t rl_peer_alloc(int32_t device, int64_t bytes, void **d_ptr, uint8_t *handle64);
RL_API int32_t rl_peer_open(int32_t device, const uint8_t *handle64, void **d_ptr);
RL_API int32_t rl_peer_close(int32_t device, void *d_ptr);
RL_API int32_t rl_peer_free(int32_t device, void *d_ptr);
/* rl_calc_range_fan with the all-gather fused in: every range is stored straight into slot      */
/* `rank` (offset rank*slot_rays) of each of the `world` gathered buffers peer_bufs[0..world)     */
/* (peer_bufs is a HOST array of device pointers as mapped in this process).  Ranks synchronise  */
/* afterwards with any stream-ordered collective before reading.  With RL_GATHER_MULTICAST      */
/* peer_bufs[0] is an NVLS multicast address bound to all `world` buffers: one multimem.st per    */
/* range, replicated by the NVSwitch.  When rank*slot_rays is a multiple of 4 the ranges of a CTA    */
/* leave as 16-byte stores (a quarter of the NVLink packets).  Reuse: a faster rank's NEXT call      */
/* starts storing into every GPU's buffer at once, so either barrier before the call as well or     */
/* alternate between two sets of buffers (pyracecarsimulator_b200.sharded.PeerGather does the latter). */
#define RL_GATHER_MULTICAST 1u
/* with RL_GATHER_MULTICAST: multimem.st.weak instead of .relaxed.sys (the kernel boundary and the barrier that */
/* follows publish the stores either way)                                                                   */
#define RL_GATHER_WEAK 2u
RL_API int32_t rl_calc_range_fan_allgather(rl_marcher *m, const float *d_poses, int64_t pose_stride_rows,
                                           void *const *peer_bufs, int32_t world, int32_t rank,
                                           int64_t slot_rays, int64_t num_poses, int32_t num_rays,
                                           float fov, uint32_t flags, void *stream);

/* The same for calc_range_repeat_angles (the particle-filter shape, BASELINE.json configs[2]):       */
/* slot `rank` of every gathered buffer receives outs[i*num_angles + a] of this rank's poses.         */
RL_API int32_t rl_calc_range_repeat_angles_allgather(rl_marcher *m, const float *d_poses, const float *d_angles,
                                                     void *const *peer_bufs, int32_t world, int32_t rank,
                                                     int64_t slot_rays, int64_t num_poses, int32_t num_angles,
                                                     uint32_t flags, void *stream);

/* The all-gather alone, for ranges that already exist on this GPU: d_src[0..n) -> slot `rank` of every       */
/* gathered buffer (same stores as the fused calls; d_src and rank*slot_rays 16-byte aligned).               */
RL_API int32_t rl_allgather_ranges(int32_t device, const float *d_src, void *const *peer_bufs, int32_t world,
                                   int32_t rank, int64_t slot_rays, int64_t n, uint32_t flags, void *stream);

/* Number of distance-field loads ("march steps") the last *_host call performed, when the */
/* marcher was asked to count them (rl_marcher_count_steps(m, 1)); used by the roofline.    */
RL_API int32_t rl_marcher_count_steps(rl_marcher *m, int32_t enable);
RL_API int32_t rl_marcher_last_steps(rl_marcher *m, uint64_t *steps);

/* ---- vehicle model, crash test and fused rollout (north_star (c)) ---- */

/* params17: the 17 doubles of the reference Car constructor, in its order                     */
/* (racecar/src/racecar.cpp:10-13): WB, FC, H_CG, L_F, L_R, CS_F, CS_R, MASS, I_Z, CRASH_THRESH, */
/* WIDTH, LENGTH, MAX_STEER_VEL, MAX_STEER_ANG, MAX_SPEED, MAX_ACCEL, MAX_DECEL.               */
RL_API int32_t rl_car_create(const double *params17, int32_t device, rl_car **out);
RL_API int32_t rl_car_destroy(rl_car *car);
/* Car::setCarEdgeDistances (racecar.cpp:239-292): fixes the beam count of every later call. */
RL_API int32_t rl_car_set_edge_distances(rl_car *car, int32_t num_rays, double min_ang, double ang_inc,
                                         double scan_dist_to_base);
RL_API int32_t rl_car_get_edge_distances(const rl_car *car, double *out, int32_t num_rays);
/* Batched Car::control + Car::updatePosition (racecar.cpp:53-98, :294-303): d_states is        */
/* (n_cars, 11) fp64 in the reference's getState layout, updated in place.                     */
RL_API int32_t rl_car_step(rl_car *car, double *d_states, const double *d_speed, const double *d_steer,
                           int64_t n_cars, double dt, void *stream);
/* Car::isCrashed (racecar.cpp:305-328) over existing ranges, `groups` independent batches of   */
/* `poses_per_group` scans: d_first[g] = first crashed pose (0-based) or -(poses_per_group+1). */
RL_API int32_t rl_is_crashed(rl_car *car, const float *d_rays, int64_t groups, int32_t poses_per_group,
                             int32_t *d_first, void *stream);
/* RacecarSimulator.checkCollisionMany (scripts/racecar_simulator_v2.py:146-167): fan scan of   */
/* d_poses (groups*poses_per_group, 3) with the crash test as the march epilogue.  d_ranges may */
/* be NULL: then no range is written and poses after a group's first crash are skipped.        */
RL_API int32_t rl_scan_crash(rl_marcher *m, rl_car *car, const float *d_poses, int64_t groups,
                             int32_t poses_per_group, float fov, int32_t *d_first, float *d_ranges,
                             void *stream);
/* MCTS.rollout (scripts/mcts.py:202-245) for n_cars cars without leaving the GPU: `steps`       */
/* updatePosition(dt), a new (speed, steer) from d_actions (n_cars, ceil(steps/action_every), 2) */
/* every action_every-th step, a fan scan after every step from the base-link pose               */
/* (lidar_pose = 0, what mcts.py:228-231 records) or the lidar pose (lidar_pose = 1,             */
/* Car::getScanPose), crash test per scan.  d_crash_index[c] = first crashed step or             */
/* -(steps+1); d_reward[c] (nullable) = sum of post-step velocities before the crash.            */
/* d_poses (steps, n_cars, 3) fp32 and d_vsum (n_cars, steps) fp64 are caller-provided scratch   */
/* that double as outputs (scanned poses, step-major; running velocity sums).                    */
RL_API int32_t rl_rollout(rl_marcher *m, rl_car *car, double *d_states, const double *d_actions,
                          int64_t n_cars, int32_t steps, int32_t action_every, double dt,
                          int32_t lidar_pose, double scan_dist_to_base, float fov,
                          int32_t *d_crash_index, double *d_reward, float *d_poses, double *d_vsum,
                          void *stream);

/* The action schedule MCTS.rollout draws on the host (scripts/mcts.py:216-222: every 10th step      */
/* rand_steer = uniform(-max_steer_ang, max_steer_ang), then rand_speed = uniform(0, max_speed)),    */
/* generated on the device instead so that no host tensor feeds rl_rollout: d_actions (n_cars,       */
/* n_actions, 2) fp64 = (speed, steer).  Counter-based (Philox4x32-10): block (action, car_lo,        */
/* car_hi, stream_id) under key (seed_lo, seed_hi); words 0,1 -> steer, words 2,3 -> speed, each the  */
/* 53-bit double ((a>>5)*2^26 + (b>>6))/2^53 mapped as lo + (hi-lo)*u.  Reproducible for any launch   */
/* shape and GPU count (a rank generates its own car range by passing global car indices through     */
/* car_offset).                                                                                       */
RL_API int32_t rl_rollout_actions(double *d_actions, int64_t n_cars, int32_t n_actions, uint64_t seed,
                                  uint32_t stream_id, int64_t car_offset, double speed_lo, double speed_hi,
                                  double steer_lo, double steer_hi, int32_t device, void *stream);

/* The value MCTS.rollout returns for a node (scripts/mcts.py:240-245):                               */
/* d_value[c] = d_reward[c] / |d_node_action[c]| (IEEE division, as numpy does it).                    */
RL_API int32_t rl_rollout_value(const double *d_reward, const double *d_node_action, int64_t n_cars,
                                double *d_value, int32_t device, void *stream);

/* FollowGap(ws, max_distance, max_angle, angle_inc).eval(scan, num_rays) for `num_scans` scans at */
/* once (followgap/followgap.hpp:104-129; caller scripts/mcts.py:262-267): d_scans is             */
/* (num_scans, num_rays) fp32 ranges in metres, d_out[s] the steering angle.  num_rays >= 10.      */
RL_API int32_t rl_follow_gap(const float *d_scans, int64_t num_scans, int32_t num_rays,
                             float max_distance, float max_angle, float angle_inc, float *d_out,
                             void *stream);

/* ---- caller-owned host buffers ---- */
/* The *_host entry points use page-locked buffers in place (ranges are stored straight into `outs` by  */
/* the kernel; no staging copy) and stage pageable ones through pinned memory owned by the marcher.    */
/* A caller that reuses its numpy-style buffers (the reference allocates them once:                     */
/* scripts/scan_simulator.py:32-40, scripts/two_player/scan.py:51-53) can page-lock them once with      */
/* rl_host_register and release them with rl_host_unregister BEFORE freeing the memory.  *was_pinned    */
/* is set to 1 when the range was page-locked already (nothing registered, do not unregister).          */
RL_API int32_t rl_host_register(int32_t device, void *ptr, int64_t bytes, int32_t *was_pinned);
RL_API int32_t rl_host_unregister(int32_t device, void *ptr);

/* ---- measurement support (not on the product path) ---- */
/* Throughput, in GB/s at 4 bytes per gather, of independent random 4-byte gathers from a    */
/* `buffer_bytes` L2-resident buffer: the denominator of the L2-gather roofline.             */
RL_API int32_t rl_gather_bandwidth(int32_t device, int64_t buffer_bytes, int32_t rounds, int32_t iters,
                                   float *gbytes_per_s);

/* Random FULL-SECTOR reads from a `buffer_bytes` L2-resident buffer (each 32-byte sector requested   */
/* once per warp instruction, every byte used, L1 bypassed): 10^9 sectors per second.  x 32 B is the   */
/* L2 -> SM bandwidth a random-access reader can reach; the march's ncu lts__t_sectors per second are  */
/* reported as a fraction of it (bench.py roofline.l2).                                                */
RL_API int32_t rl_l2_sector_bandwidth(int32_t device, int64_t buffer_bytes, int32_t rounds, int32_t iters,
                                      float *gsectors_per_s);

/* Demote all persisting L2 lines to normal.  The march launches pin the distance field in L2 with an  */
/* access-policy window (it survives unrelated traffic between calls); a benchmark that wants a truly  */
/* cold cache calls this together with its L2 flush.                                                  */
RL_API int32_t rl_l2_reset_persisting(int32_t device);

/* The device's sinf/cosf (glibc's algorithm, csrc/glibc_trig.cuh) over an array on the      */
/* current device, for the parity tests: d_sin[i] = sinf(d_in[i]), d_cos[i] = cosf(d_in[i]). */
RL_API int32_t rl_probe_sincosf(const float *d_in, float *d_sin, float *d_cos, int64_t n, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* RANGELIB_B200_H */
