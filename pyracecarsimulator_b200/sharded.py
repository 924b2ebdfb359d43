"""Pose-sharded scans across the GPUs of one box: one process per GPU under
``torch.distributed``, the map and its distance field replicated on every GPU (each rank runs the
ingest kernels itself -- it is a sub-millisecond job), the pose batch split into contiguous index
ranges, and ONE collective: an all-gather (or gather-to-root) of the fp32 ranges over NCCL/NVLink.
There is no other exchange step -- every ray is independent.

The march itself is injected as a callable so that the partition / padding / gather logic can be
exercised on CPU with the ``gloo`` backend (tests/test_sharded_gloo.py); the product binding is
:func:`gpu_march_fn`.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch
import torch.distributed as dist


def shard_bounds(n: int, world: int, rank: int):
    """Contiguous pose range [lo, hi) of `rank`: ceil(n / world) poses per rank, the last ranks
    possibly short or empty (pose-major output stays contiguous per rank)."""
    per = -(-n // world) if n > 0 else 0
    lo = min(n, rank * per)
    hi = min(n, lo + per)
    return lo, hi


class _DevicePtr:
    """Zero-copy torch view of a raw device allocation (``__cuda_array_interface__``)."""

    def __init__(self, ptr: int, n_floats: int):
        self.__cuda_array_interface__ = {"shape": (n_floats,), "typestr": "<f4", "data": (int(ptr), False),
                                         "version": 2}


class PeerGather:
    """The fused march + all-gather: one gathered-ranges buffer per GPU (``world * slot_rays`` floats)
    mapped into every process, so that ``rl_calc_range_fan_allgather`` can store each range straight
    into slot ``rank`` of all ``world`` buffers over NVLink while it marches (no separate collective,
    no staging copy).  ``sync()`` is the stream-ordered barrier after which every rank may read
    ``tensor()``.

    backend "symm":  torch symmetric memory does the plumbing (allocation, rendezvous, signal-pad
                     barrier); with ``multicast=True`` and NVLS support the kernel issues ONE
                     ``multimem.st`` per range to the multicast address and the NVSwitch replicates it.
    backend "ipc":   buffers allocated by the C ABI and exchanged as CUDA IPC handles; barrier = a
                     4-byte NCCL all_reduce.  Used when symmetric memory is unavailable.
    """

    def __init__(self, device_index: int, slot_rays: int, group: Optional[dist.ProcessGroup] = None,
                 backend: str = "auto", multicast: bool = True):
        import ctypes as C
        from . import _native
        self._native = _native
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.device_index = int(device_index)
        self.slot_rays = int(slot_rays)
        self.flags = 0
        self._hdl = None
        self._own = None
        self._opened = []
        n_floats = self.world * self.slot_rays
        if backend in ("auto", "symm"):
            try:
                self._init_symm(n_floats, multicast)
                self.backend = "symm"
            except Exception as e:   # noqa: BLE001 - any failure of the optional plumbing -> IPC path
                if backend == "symm":
                    raise
                self._symm_error = repr(e)
                self._hdl = None
        if self._hdl is None:
            self._init_ipc(n_floats)
            self.backend = "ipc"

    def _init_symm(self, n_floats, multicast):
        import ctypes as C
        import torch.distributed._symmetric_memory as symm_mem
        dev = torch.device("cuda", self.device_index)
        t = symm_mem.empty(n_floats, dtype=torch.float32, device=dev)
        grp = self.group if self.group is not None else dist.group.WORLD
        hdl = symm_mem.rendezvous(t, grp)
        ptrs = [int(p) for p in hdl.buffer_ptrs]
        mc = int(hdl.multicast_ptr or 0) if multicast else 0   # 0 when the fabric has no NVLS multicast
        # every rank must take the same decision
        ok = torch.tensor([1 if mc else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
        if int(ok.item()) == 1:
            self.flags = 1   # RL_GATHER_MULTICAST
            self.ptrs = (C.c_void_p * self.world)(*([mc] + ptrs[1:]))
        else:
            self.ptrs = (C.c_void_p * self.world)(*ptrs)
        self._view = t
        self._hdl = hdl

    def _init_ipc(self, n_floats):
        import ctypes as C
        L = self._native.lib()
        own = C.c_void_p()
        handle = (C.c_uint8 * 64)()
        self._native.check(L.rl_peer_alloc(self.device_index, n_floats * 4, C.byref(own), handle), "rl_peer_alloc")
        self._own = own
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle), group=self.group)
        ptrs = []
        for q in range(self.world):
            if q == self.rank:
                ptrs.append(own.value)
                continue
            p = C.c_void_p()
            buf = (C.c_uint8 * 64).from_buffer_copy(handles[q])
            self._native.check(L.rl_peer_open(self.device_index, buf, C.byref(p)), "rl_peer_open")
            self._opened.append(p)
            ptrs.append(p.value)
        self.ptrs = (C.c_void_p * self.world)(*ptrs)
        self._flag = torch.zeros(1, dtype=torch.int32, device=f"cuda:{self.device_index}")
        self._view = torch.as_tensor(_DevicePtr(own.value, n_floats), device=f"cuda:{self.device_index}")

    @property
    def multicast(self) -> bool:
        return bool(self.flags & 1)

    def tensor(self) -> torch.Tensor:
        """This GPU's gathered buffer: (world * slot_rays,) float32, slot r = ranges of rank r."""
        return self._view

    def march(self, marcher, poses: torch.Tensor, fov: float, num_rays: int, stream_ptr: int):
        n = poses.shape[0]
        self._native.check(self._native.lib().rl_calc_range_fan_allgather(
            marcher._h, poses.data_ptr(), 1, self.ptrs, self.world, self.rank, self.slot_rays, n, int(num_rays),
            float(fov), self.flags, stream_ptr), "rl_calc_range_fan_allgather")

    def sync(self):
        """Stream-ordered barrier: returns (on the stream) once every rank's march has completed."""
        if self._hdl is not None:
            self._hdl.barrier()
        else:
            dist.all_reduce(self._flag, group=self.group)

    def close(self):
        """Collective: unmap every peer buffer on every rank before any rank frees its own."""
        L = self._native.lib()
        torch.cuda.synchronize(self.device_index)
        dist.barrier(group=self.group)
        for p in self._opened:
            L.rl_peer_close(self.device_index, p)
        self._opened = []
        dist.barrier(group=self.group)
        self._view = None
        self._hdl = None
        if self._own is not None:
            L.rl_peer_free(self.device_index, self._own)
            self._own = None


def gpu_march_fn(marcher, fov: float, num_rays: int) -> Callable:
    """The product march: ``PyRayMarchingGPU.calc_range_fan`` on device tensors."""
    def run(poses: torch.Tensor, out: torch.Tensor):
        if poses.shape[0]:
            marcher.calc_range_fan(poses, out, fov, num_rays)
    return run


class ShardedScanner:
    """``scan(poses)`` over the whole process group.

    poses:   (N, 3) float32, identical on every rank (only the rank's own rows are read).
    returns: (N * num_rays,) float32 ranges, pose-major -- on every rank for ``gather="all"``, on
             ``root`` only (None elsewhere) for ``gather="root"``, and just the local shard
             (``hi - lo`` poses) for ``gather="none"``.
    """

    def __init__(self, march_fn: Callable, num_rays: int, device: torch.device,
                 group: Optional[dist.ProcessGroup] = None, chunks: int = 1):
        self.march_fn = march_fn
        self.num_rays = int(num_rays)
        self.device = device
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.chunks = max(1, int(chunks))

    def scan(self, poses: torch.Tensor, gather: str = "all", root: int = 0):
        if gather not in ("all", "root", "none"):
            raise ValueError("gather must be 'all', 'root' or 'none'")
        n = poses.shape[0]
        lo, hi = shard_bounds(n, self.world, self.rank)
        per = -(-n // self.world) if n > 0 else 0
        R = self.num_rays
        # equal, padded shards: all_gather needs the same count from every rank
        local = torch.zeros(per * R, dtype=torch.float32, device=self.device)
        mine = poses[lo:hi].to(self.device).contiguous()
        self.march_fn(mine, local[:(hi - lo) * R])
        if gather == "none" or self.world == 1:
            out = local[:(hi - lo) * R]
            return out if (gather != "root" or self.rank == root) else None
        if gather == "all":
            full = torch.empty(self.world * per * R, dtype=torch.float32, device=self.device)
            dist.all_gather_into_tensor(full, local, group=self.group)
            return full[:n * R]
        parts = [torch.empty_like(local) for _ in range(self.world)] if self.rank == root else None
        dist.gather(local, parts, dst=root, group=self.group)
        if self.rank != root:
            return None
        return torch.cat(parts)[:n * R]

    def scan_fused(self, poses: torch.Tensor, marcher, fov: float, peer: "PeerGather" = None):
        """All-gather fused into the march kernel (``rl_calc_range_fan_allgather``): every rank's ranges
        are stored straight into every GPU's gathered buffer over NVLink while the march runs.
        Returns the (N * num_rays,) gathered ranges (a view of the peer buffer, valid until the next
        call) on every rank.  Pass a long-lived :class:`PeerGather` to avoid re-creating the mappings."""
        n = poses.shape[0]
        lo, hi = shard_bounds(n, self.world, self.rank)
        per = -(-n // self.world) if n > 0 else 0
        R = self.num_rays
        if peer is None:
            peer = self._peer if getattr(self, "_peer", None) is not None and self._peer.slot_rays >= per * R else None
            if peer is None:
                peer = self._peer = PeerGather(self.device.index, max(per * R, 1), self.group)
        mine = poses[lo:hi].to(self.device).contiguous()
        stream = int(torch.cuda.current_stream(self.device.index).cuda_stream)
        peer.march(marcher, mine, fov, R, stream)
        peer.sync()
        full = peer.tensor()
        if peer.slot_rays == per * R:
            return full[:n * R]
        # slots padded beyond this batch's shard size: compact the used parts
        return torch.cat([full[r * peer.slot_rays: r * peer.slot_rays + max(0, min(n, (r + 1) * per) - r * per) * R]
                          for r in range(self.world)])


def gpu_rollout_fn(car, marcher, fov: float, action_every: int = 10, dt: float = 0.01,
                   lidar_pose: bool = False, scan_dist_to_base: float = 0.275) -> Callable:
    """The product rollout: ``BatchedCar.rollout`` with the action schedule drawn on the device for
    the rank's global car range (``car_offset``), so every rank sees the schedule it would see in a
    single-GPU run of the whole job."""
    def run(states: torch.Tensor, car_offset: int, steps: int, seed: int, stream_id: int):
        out = car.rollout(marcher, states, None, steps, fov, action_every=action_every, dt=dt,
                          lidar_pose=lidar_pose, scan_dist_to_base=scan_dist_to_base, seed=seed,
                          stream_id=stream_id, car_offset=car_offset)
        return out["crash_index"], out["reward"]
    return run


class ShardedRollout:
    """Config 4 across the GPUs of one box: cars split into contiguous ranges, every
    rank rolls its own cars out on its replica of the map (steps + scans + crash test, nothing leaves
    the GPU), and the only exchange is the all-gather of ``(crash_index int32, reward float64)`` per
    car -- 12 bytes per car instead of 4 bytes per ray.

    ``rollout_fn(states_local, car_offset, steps, seed, stream_id) -> (crash_index, reward)`` is
    injected (product: :func:`gpu_rollout_fn`; the gloo tests pass a CPU stand-in).
    """

    def __init__(self, rollout_fn: Callable, device: torch.device, group: Optional[dist.ProcessGroup] = None):
        self.rollout_fn = rollout_fn
        self.device = device
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0

    def rollout(self, states: torch.Tensor, steps: int, seed: int = 42, stream_id: int = 0, gather: str = "all"):
        """``states``: (N, 11) float64, identical on every rank (only the rank's rows are read; they
        are NOT updated in place -- each rank works on a copy of its shard).  Returns
        ``(crash_index (N,) int32, reward (N,) float64)`` on every rank (``gather="all"``) or just the
        rank's shard (``gather="none"``)."""
        if gather not in ("all", "none"):
            raise ValueError("gather must be 'all' or 'none'")
        n = states.shape[0]
        lo, hi = shard_bounds(n, self.world, self.rank)
        per = -(-n // self.world) if n > 0 else 0
        crash = torch.full((per,), -(steps + 1), dtype=torch.int32, device=self.device)
        reward = torch.zeros(per, dtype=torch.float64, device=self.device)
        if hi > lo:
            mine = states[lo:hi].to(self.device).contiguous().clone()
            c, r = self.rollout_fn(mine, lo, int(steps), int(seed), int(stream_id))
            crash[:hi - lo] = c
            reward[:hi - lo] = r
        if gather == "none" or self.world == 1:
            return crash[:hi - lo], reward[:hi - lo]
        all_crash = torch.empty(self.world * per, dtype=torch.int32, device=self.device)
        all_reward = torch.empty(self.world * per, dtype=torch.float64, device=self.device)
        dist.all_gather_into_tensor(all_crash, crash, group=self.group)
        dist.all_gather_into_tensor(all_reward, reward, group=self.group)
        return all_crash[:n], all_reward[:n]
