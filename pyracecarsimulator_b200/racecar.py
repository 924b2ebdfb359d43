"""Batched vehicle model, crash test and fused rollout on the GPU -- the callers either side of the
scan in the reference:

  * ``racecar.PyCar`` (racecar/pywrapper/racecar.pyx:75-114 over racecar/src/racecar.cpp):
    ``control`` + ``updatePosition`` -> :meth:`BatchedCar.step`, ``setCarEdgeDistances``,
    ``isCrashed``, ``getScanPose``;
  * ``RacecarSimulator.checkCollisionMany`` (scripts/racecar_simulator_v2.py:146-167)
    -> :meth:`BatchedCar.scan_crash` (scan + crash test in one kernel, ranges optional);
  * ``MCTS.rollout`` (scripts/mcts.py:202-245) -> :meth:`BatchedCar.rollout` for many cars at once.

torch tensors are only the device-buffer carrier; all arithmetic is in csrc/car.cu behind the C ABI.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native
from .range_libc import PyRayMarchingGPU, _current_stream_ptr

# order of the reference Car constructor (racecar/src/racecar.cpp:10-13) and the config keys
# RacecarSimulator feeds it (scripts/racecar_simulator_v2.py:34-43)
CAR_PARAM_ORDER = ("wb", "fc", "h_cg", "l_f", "l_r", "cs_f", "cs_r", "mass", "I_z", "ttc_thresh", "width",
                   "length", "max_steer_vel", "max_steer_ang", "max_speed", "max_accel", "max_decel")

# params.yaml:1-22 (wheelbase ... moment_inertia), :31-36
DEFAULT_CAR_CONFIG = dict(wb=0.3302, fc=1.0, h_cg=0.08255, l_f=0.15875, l_r=0.17145, cs_f=2.3, cs_r=2.3,
                          mass=3.17, I_z=0.0398378, ttc_thresh=0.001, width=0.2032, length=0.4064,
                          max_steer_vel=5.0, max_steer_ang=0.4189, max_speed=7.0, max_accel=3.0,
                          max_decel=20.0)


def car_params(config=None) -> np.ndarray:
    cfg = dict(DEFAULT_CAR_CONFIG)
    if config:
        cfg.update({k: config[k] for k in CAR_PARAM_ORDER if k in config})
    return np.array([float(cfg[k]) for k in CAR_PARAM_ORDER], dtype=np.float64)


def _dev_tensor(t, name, dtype, device):
    import torch
    if not isinstance(t, torch.Tensor) or not t.is_cuda or t.dtype != dtype or not t.is_contiguous():
        raise ValueError(f"{name}: expected a contiguous CUDA tensor of dtype {dtype}")
    if t.device.index != device:
        raise ValueError(f"{name}: tensor is on cuda:{t.device.index}, car is on cuda:{device}")
    return t


class BatchedCar:
    """N cars sharing one parameter set; state is an (N, 11) float64 CUDA tensor in the reference's
    getState layout: x, y, theta, velocity, steer_angle, angular_velocity, slip_angle, st_dyn,
    travel_dist, total_velo, update_count (racecar/src/racecar.cpp:357-376)."""

    def __init__(self, config=None, device: int = 0):
        self.device = int(device)
        self.params = car_params(config)
        h = C.c_void_p()
        _native.check(_native.lib().rl_car_create(self.params.ctypes.data, self.device, C.byref(h)), "BatchedCar")
        self._h = h
        self.num_rays = 0

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and _native._lib is not None:
            _native._lib.rl_car_destroy(h)

    # ---- Car::setCarEdgeDistances ----
    def setCarEdgeDistances(self, num_rays, ang_min, scan_ang_inc, scan_dist_to_base):
        _native.check(_native.lib().rl_car_set_edge_distances(self._h, int(num_rays), float(ang_min),
                                                              float(scan_ang_inc), float(scan_dist_to_base)),
                      "setCarEdgeDistances")
        self.num_rays = int(num_rays)

    def edge_distances(self) -> np.ndarray:
        out = np.empty(self.num_rays, dtype=np.float64)
        _native.check(_native.lib().rl_car_get_edge_distances(self._h, out.ctypes.data, self.num_rays))
        return out

    # ---- Car::control + Car::updatePosition, batched ----
    def step(self, states, speed, steer, dt=0.01):
        import torch
        st = _dev_tensor(states, "states", torch.float64, self.device)
        sp = _dev_tensor(speed, "speed", torch.float64, self.device)
        sa = _dev_tensor(steer, "steer", torch.float64, self.device)
        n = st.shape[0]
        if st.dim() != 2 or st.shape[1] != 11 or sp.numel() != n or sa.numel() != n:
            raise ValueError("states must be (N, 11); speed and steer (N,)")
        _native.check(_native.lib().rl_car_step(self._h, st.data_ptr(), sp.data_ptr(), sa.data_ptr(), n, float(dt),
                                                _current_stream_ptr(self.device)), "step")
        return states

    @staticmethod
    def scan_pose(states, scan_dist_to_base):
        """Car::getScanPose (racecar.cpp:378-387) for every row of ``states``: (N, 3) float64."""
        import torch
        x, y, th = states[:, 0], states[:, 1], states[:, 2]
        return torch.stack([x + scan_dist_to_base * torch.cos(th), y + scan_dist_to_base * torch.sin(th), th], dim=1)

    # ---- Car::isCrashed ----
    def is_crashed_many(self, rays, groups, poses_per_group):
        """``rays``: float32 CUDA tensor of groups*poses_per_group*num_rays ranges.
        Returns an int32 CUDA tensor (groups,): first crashed pose or -(poses_per_group+1)."""
        import torch
        r = _dev_tensor(rays, "rays", torch.float32, self.device)
        if r.numel() != groups * poses_per_group * self.num_rays:
            raise ValueError("rays must hold groups * poses_per_group * num_rays ranges")
        first = torch.empty(groups, dtype=torch.int32, device=r.device)
        _native.check(_native.lib().rl_is_crashed(self._h, r.data_ptr(), groups, poses_per_group, first.data_ptr(),
                                                  _current_stream_ptr(self.device)), "isCrashed")
        return first

    def isCrashed(self, rays, num_rays, poses) -> int:
        """Upstream signature (racecar.pyx: isCrashed(rays, num_rays, poses)): host or device ranges of
        ``poses`` scans; returns the 0-based index of the first crashed pose or -(poses+1)."""
        import torch
        if int(num_rays) != self.num_rays:
            raise ValueError("num_rays differs from setCarEdgeDistances")
        if isinstance(rays, np.ndarray):
            rays = torch.from_numpy(np.ascontiguousarray(rays, dtype=np.float32)).to(f"cuda:{self.device}")
        return int(self.is_crashed_many(rays.reshape(-1)[:poses * self.num_rays].contiguous(), 1, int(poses)).item())

    # ---- scan + crash in one kernel ----
    def scan_crash(self, marcher: PyRayMarchingGPU, poses, groups, poses_per_group, fov, want_ranges=False):
        """``poses``: (groups*poses_per_group, 3) float32 CUDA tensor, group-major.
        Returns (first_crash int32 (groups,), ranges or None)."""
        import torch
        p = _dev_tensor(poses, "poses", torch.float32, self.device)
        if p.numel() != groups * poses_per_group * 3:
            raise ValueError("poses must be (groups * poses_per_group, 3)")
        first = torch.empty(groups, dtype=torch.int32, device=p.device)
        ranges = torch.empty(groups * poses_per_group * self.num_rays, dtype=torch.float32, device=p.device) if want_ranges else None
        _native.check(_native.lib().rl_scan_crash(marcher._h, self._h, p.data_ptr(), groups, poses_per_group, float(fov),
                                                  first.data_ptr(), ranges.data_ptr() if want_ranges else None,
                                                  _current_stream_ptr(self.device)), "scan_crash")
        return first, ranges

    # ---- the random action schedule of MCTS.rollout, drawn on the device ----
    def random_actions(self, n_cars, n_actions, seed=42, stream_id=0, car_offset=0, speed_range=None,
                       steer_range=None):
        """(n_cars, n_actions, 2) float64 CUDA tensor of (speed, steer) targets, the device-side
        counterpart of scripts/mcts.py:216-222 (steer ~ U(-max_steer_ang, max_steer_ang), then
        speed ~ U(0, max_speed), one pair per 10 steps).  Counter-based Philox4x32-10: the value for
        (seed, stream_id, car_offset + car, action) does not depend on batch shape or GPU count."""
        import torch
        cfg = dict(zip(CAR_PARAM_ORDER, self.params))
        s_lo, s_hi = speed_range if speed_range is not None else (0.0, cfg["max_speed"])
        a_lo, a_hi = steer_range if steer_range is not None else (-cfg["max_steer_ang"], cfg["max_steer_ang"])
        out = torch.empty((int(n_cars), int(n_actions), 2), dtype=torch.float64, device=f"cuda:{self.device}")
        _native.check(_native.lib().rl_rollout_actions(out.data_ptr(), int(n_cars), int(n_actions), int(seed),
                                                       int(stream_id), int(car_offset), float(s_lo), float(s_hi),
                                                       float(a_lo), float(a_hi), self.device,
                                                       _current_stream_ptr(self.device)), "random_actions")
        return out

    # ---- MCTS.rollout for many cars ----
    def rollout(self, marcher: PyRayMarchingGPU, states, actions, steps, fov, action_every=10, dt=0.01,
                lidar_pose=False, scan_dist_to_base=0.275, seed=42, stream_id=0, car_offset=0, node_action=None):
        """``states`` (N, 11) float64 CUDA (updated in place), ``actions`` (N, ceil(steps/action_every), 2)
        float64 CUDA with (speed, steer) targets, or ``None`` to draw the reference's random schedule on
        the device (:meth:`random_actions` with ``seed``/``stream_id``/``car_offset``; returned under
        ``"actions"``).  Returns dict(crash_index int32 (N,), reward float64 (N,),
        poses float32 (steps, N, 3), vsum float64 (N, steps)); with ``node_action`` ((N,) float64 CUDA, the
        action of the node each rollout starts from) also ``value = reward / abs(node_action)``, what
        ``MCTS.rollout`` returns (scripts/mcts.py:240-245)."""
        import torch
        st = _dev_tensor(states, "states", torch.float64, self.device)
        n = st.shape[0]
        n_act = (steps + action_every - 1) // action_every
        if actions is None:
            actions = self.random_actions(n, n_act, seed=seed, stream_id=stream_id, car_offset=car_offset)
        ac = _dev_tensor(actions, "actions", torch.float64, self.device)
        if st.dim() != 2 or st.shape[1] != 11 or tuple(ac.shape) != (n, n_act, 2):
            raise ValueError(f"states must be (N, 11) and actions (N, {n_act}, 2)")
        dev = st.device
        crash = torch.empty(n, dtype=torch.int32, device=dev)
        reward = torch.empty(n, dtype=torch.float64, device=dev)
        poses = torch.empty((steps, n, 3), dtype=torch.float32, device=dev)
        vsum = torch.empty((n, steps), dtype=torch.float64, device=dev)
        _native.check(_native.lib().rl_rollout(marcher._h, self._h, st.data_ptr(), ac.data_ptr(), n, int(steps),
                                               int(action_every), float(dt), int(bool(lidar_pose)),
                                               float(scan_dist_to_base), float(fov), crash.data_ptr(),
                                               reward.data_ptr(), poses.data_ptr(), vsum.data_ptr(),
                                               _current_stream_ptr(self.device)), "rollout")
        out = dict(crash_index=crash, reward=reward, poses=poses, vsum=vsum, actions=actions)
        if node_action is not None:
            na = _dev_tensor(node_action, "node_action", torch.float64, self.device)
            if na.numel() != n:
                raise ValueError("node_action must be (N,)")
            out["value"] = torch.empty(n, dtype=torch.float64, device=dev)
            _native.check(_native.lib().rl_rollout_value(reward.data_ptr(), na.data_ptr(), n, out["value"].data_ptr(),
                                                         self.device, _current_stream_ptr(self.device)), "rollout value")
        return out
