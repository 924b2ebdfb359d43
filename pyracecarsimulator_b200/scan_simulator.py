"""``ScanSimulator2D`` -- the reference's lidar facade (scripts/scan_simulator.py:11-135), same
constructor, methods, argument meaning and return values, running on the GPU marcher.

Kept from the reference on purpose:
  * ``scan()`` returns the SAME cached ``output_vector`` object every call and ``scanMany()`` the
    same ``output_vector_many`` (callers alias them: scripts/mcts.py:194; copy if you keep one);
  * ``scanMany`` reads exactly ``batch_size`` poses, ``poses[i][0..2]``;
  * no noise is added (``scan_std`` is stored, the noise line is commented out at :109);
  * beam ``j`` of a scan heads ``theta - fov/2 + j*fov/num_rays``.
Changed: no ``print`` + ``sys.exit()`` (:67-69, :77-79) -- errors raise; the buffers are pinned
host memory when a GPU is present so ranges are DMA-ed straight into the arrays callers see.
"""
from __future__ import annotations

import math

import numpy as np

from . import range_libc


def _pinned_zeros(shape, dtype=np.float32) -> np.ndarray:
    """numpy view of page-locked host memory (plain numpy when torch/CUDA is unavailable).  The
    returned array keeps its backing tensor alive through ``.base``."""
    try:
        import torch
        if torch.cuda.is_available():
            return torch.zeros(shape, dtype=getattr(torch, np.dtype(dtype).name), pin_memory=True).numpy()
    except ImportError:
        pass
    return np.zeros(shape, dtype=dtype)


class ScanSimulator2D:

    def __init__(self, num_rays, fov, scan_std, batch_size=100):
        self.num_rays, self.batch_size = int(num_rays), int(batch_size)
        self.fov, self.scan_std = fov, scan_std
        self.twopi = 2.0 * math.pi
        n, b = self.num_rays, self.batch_size
        # single-scan buffers: only row 0 of input_vector is ever meaningful (fork's 4-arg layout)
        self.input_vector = np.zeros((n, 3), dtype=np.float32)
        self.output_vector = _pinned_zeros(n)
        self.noise = np.zeros(n, dtype=np.float32)
        # batch buffers: a compact (batch_size, 3) pose block goes to the GPU; the reference's
        # (batch_size*num_rays, 3) block, of which every num_rays-th row is meaningful, is
        # materialised on demand by the property below
        self.poses_many = _pinned_zeros((b, 3))
        self.output_vector_many = _pinned_zeros(b * n)
        self.hasMap, self.scan_method = False, None

    @property
    def input_vector_many(self) -> np.ndarray:
        """The reference's (batch_size*num_rays, 3) input block (scripts/scan_simulator.py:39-40)."""
        full = np.zeros((self.batch_size * self.num_rays, 3), dtype=np.float32)
        full[::self.num_rays] = self.poses_many
        return full

    def setMap(self, ros_map, max_range_px, resolution, origin):
        """ros_map: a ``range_libc.PyOMap``; max_range_px in pixels; origin (x, y, yaw)."""
        self.omap, self.mrx, self.res = ros_map, max_range_px, resolution
        self.origin_x, self.origin_y = origin[0], origin[1]
        self.origin_c, self.origin_s = math.cos(origin[2]), math.sin(origin[2])
        self.hasMap = True

    def setRaytracingMethod(self, method="RM"):
        if not self.hasMap:
            raise RuntimeError("for setRaytracingMethod use setMap first")
        if method == "RM":
            self.scan_method = range_libc.PyRayMarching(self.omap, self.mrx)
        elif method == "RMGPU":
            self.scan_method = range_libc.PyRayMarchingGPU(self.omap, self.mrx)
        else:
            raise ValueError("Only ray marching is supported")

    def updateMap(self, ros_map):
        self.ros_map = ros_map

    def laser_scan_metadata(self, scan_max_range=None):
        """The LaserScan fields the reference publishes with every scan (scripts/ros_interface.py:342-345)."""
        return dict(angle_min=-self.fov / 2.0, angle_max=self.fov / 2.0, angle_increment=self.fov / self.num_rays,
                    range_max=scan_max_range if scan_max_range is not None else self.mrx * self.res)

    def scan(self, x, y, theta):
        if not self.hasMap or self.scan_method is None:
            raise RuntimeError("Doing a scan without a defined map / ray tracing method")
        self.input_vector[0] = (x, y, theta)
        self.scan_method.calc_range_many(self.input_vector, self.output_vector, self.fov, self.num_rays)
        return self.output_vector

    def scanMany(self, poses):
        if not self.hasMap or self.scan_method is None:
            raise RuntimeError("Doing a scan without a defined map / ray tracing method")
        if isinstance(poses, np.ndarray) and poses.ndim == 2:
            self.poses_many[:] = poses[:self.batch_size, :3]
        else:
            for i in range(self.batch_size):
                self.poses_many[i] = poses[i][0], poses[i][1], poses[i][2]
        self.scan_method.calc_range_fan(self.poses_many, self.output_vector_many, self.fov,
                                        self.num_rays)
        return self.output_vector_many
