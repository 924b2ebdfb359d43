"""``RacecarSimulator`` -- the reference's car + lidar facade (scripts/racecar_simulator_v2.py:4-204)
restated for Python 3 on the GPU path: same constructor config keys, same method names and return
values.  The single car is a ``BatchedCar`` of size one (the vehicle model runs in csrc/car.cu like
everything else; there is no CPU implementation), the lidar is ``ScanSimulator2D``.

Hot-path rows: ``setMap`` (a8: ``max_range_px = int(scan_max_range / resolution)``),
``runScan`` (a9: scan from ``Car::getScanPose``), ``checkCollisionMany`` (a10: scanMany + isCrashed --
here one fused kernel, ranges never written).
"""
from __future__ import annotations

import math

import numpy as np

from .racecar import BatchedCar
from .scan_simulator import ScanSimulator2D


class RacecarSimulator:

    def __init__(self, config, verbose=False, device: int = 0):
        import torch
        self.verbose = verbose
        self.map_frame, self.base_frame, self.scan_frame = "map", "base_link", "laser"
        self.config = config
        # attribute <- config key, the names the reference facade exposes
        # (scripts/racecar_simulator_v2.py:16-32)
        for attr, key in (("scan_dist_to_base", "scan_dist_to_base"), ("max_speed", "max_speed"),
                          ("max_accel", "max_accel"), ("max_steer_ang", "max_steer_ang"),
                          ("max_steer_vel", "max_steer_vel"), ("max_decel", "max_decel"), ("width", "width"),
                          ("length", "length"), ("batch_size", "batch_size"), ("num_rays", "scan_beams"),
                          ("scan_fov", "scan_fov"), ("scan_std", "scan_std"), ("scan_max_range", "scan_max_range"),
                          ("free_thresh", "free_thresh"), ("ttc_thresh", "ttc_thresh")):
            setattr(self, attr, config[key])

        self.device = int(device)
        self._torch = torch
        self.car = BatchedCar(config, device=self.device)
        # scripts/racecar_simulator_v2.py:47-50
        self.car.setCarEdgeDistances(self.num_rays, -self.scan_fov / 2.0, self.scan_fov / self.num_rays,
                                     self.scan_dist_to_base)
        self.scan_simulator = ScanSimulator2D(self.num_rays, self.scan_fov, self.scan_std, self.batch_size)
        self.scan = np.zeros(self.num_rays, dtype=np.float32)
        self.desired_speed = 0.0
        self.desired_steer_ang = 0.0
        dev = f"cuda:{self.device}"
        self._state = torch.zeros((1, 11), dtype=torch.float64, device=dev)
        self._speed = torch.zeros(1, dtype=torch.float64, device=dev)
        self._steer = torch.zeros(1, dtype=torch.float64, device=dev)
        self._poses_dev = torch.zeros((self.batch_size, 3), dtype=torch.float32, device=dev)

    # ---- state (11 doubles, racecar/src/racecar.cpp:330-376) ----
    def setState(self, state):
        self._state.copy_(self._torch.as_tensor(np.asarray(state, dtype=np.float64).reshape(1, 11)))

    def getState(self):
        return self._state[0].cpu().numpy().copy()

    def getMeanVelocity(self):
        s = self.getState()
        return s[9] / s[10]

    def getTravelDistance(self):
        return float(self.getState()[8])

    def getScan(self):
        return self.scan

    # ---- scan ----
    def runScan(self):
        s = self.getState()
        # Car::getScanPose (racecar.cpp:378-387), fp64 on the host like the reference's call
        x = s[0] + self.scan_dist_to_base * math.cos(s[2])
        y = s[1] + self.scan_dist_to_base * math.sin(s[2])
        self.scan = self.scan_simulator.scan(x, y, s[2])

    def drive(self, desired_speed, desired_steer_ang):
        self.desired_speed = desired_speed
        self.desired_steer_ang = desired_steer_ang

    def updatePose(self, dt=0.01):
        self._speed.fill_(float(self.desired_speed))
        self._steer.fill_(float(self.desired_steer_ang))
        self.car.step(self._state, self._speed, self._steer, dt)

    def checkCollision(self):
        return self.car.isCrashed(np.array(self.scan, dtype=np.float32), self.num_rays, 1)

    def checkCollisionMany(self, poses, want_ranges=False):
        """Scan ``batch_size`` poses and return the index of the first crashed one, or
        ``-(batch_size + 1)``: one fused kernel, only 4 bytes come back.

        Difference from the reference: its ``scanMany`` call also leaves all ``batch_size * num_rays``
        ranges in ``scan_simulator.output_vector_many`` (scripts/racecar_simulator_v2.py:153); nothing in
        the reference reads them afterwards, and here poses after the first crash are not even scanned, so
        that buffer is left untouched.  ``want_ranges=True`` restores the side effect (every pose scanned,
        the ranges copied into ``output_vector_many``)."""
        p = np.ascontiguousarray(np.asarray(poses, dtype=np.float32)[:self.batch_size, :3])
        self._poses_dev.copy_(self._torch.from_numpy(p))
        first, ranges = self.car.scan_crash(self.scan_simulator.scan_method, self._poses_dev, 1, self.batch_size,
                                            self.scan_fov, want_ranges=want_ranges)
        if want_ranges:
            self._torch.from_numpy(self.scan_simulator.output_vector_many).copy_(ranges)
        return int(first.item())

    def stop(self):
        self._state.zero_()
        self.desired_speed = 0.0
        self.desired_steer_ang = 0.0

    # ---- map ----
    def setMap(self, ros_map, resolution, origin):
        max_range_px = int(self.scan_max_range / resolution)   # scripts/racecar_simulator_v2.py:196
        self.scan_simulator.setMap(ros_map, max_range_px, resolution, origin)

    def setRaytracingMethod(self, method="RMGPU"):
        self.scan_simulator.setRaytracingMethod(method)
