"""Batched follow-the-gap (``followgap.PyFollowGap``, followgap/followgap.pyx:23-31 over
followgap/followgap.hpp): same constructor and ``eval(lidar, size)``, plus ``eval_many`` for a
(num_scans, num_rays) CUDA tensor of scans straight from the marcher."""
from __future__ import annotations

import numpy as np

from . import _native
from .range_libc import _current_stream_ptr


class PyFollowGap:

    def __init__(self, ws, md, ma, angle_inc, device: int = 0):
        self.window_size, self.max_distance, self.max_angle, self.angle_inc = int(ws), float(md), float(ma), float(angle_inc)
        self.device = int(device)

    def eval_many(self, scans):
        import torch
        if not (isinstance(scans, torch.Tensor) and scans.is_cuda and scans.dtype == torch.float32
                and scans.is_contiguous() and scans.dim() == 2):
            raise ValueError("scans: expected a contiguous (num_scans, num_rays) float32 CUDA tensor")
        out = torch.empty(scans.shape[0], dtype=torch.float32, device=scans.device)
        _native.check(_native.lib().rl_follow_gap(scans.data_ptr(), scans.shape[0], scans.shape[1], self.max_distance,
                                                  self.max_angle, self.angle_inc, out.data_ptr(),
                                                  _current_stream_ptr(scans.device.index)), "follow_gap")
        return out

    def eval(self, lidar, size) -> float:
        import torch
        l = np.ascontiguousarray(np.asarray(lidar, dtype=np.float32)[:size])
        return float(self.eval_many(torch.from_numpy(l).to(f"cuda:{self.device}").reshape(1, -1)).item())
