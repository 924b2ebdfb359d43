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


def gather_set_layout(world: int, slot_rays: int, nbuf: int):
    """Float offsets of the ``nbuf`` gathered-buffer sets inside one allocation and its total size.  A set holds
    ``world * slot_rays`` floats; sets start on 256-byte boundaries, because the fused gather stores 16 bytes at a
    time (``multimem.st.v4`` / ``float4``): with an odd ``slot_rays`` a second set placed right behind the first
    would start on an 8- or 4-byte boundary only."""
    used = int(world) * int(slot_rays)
    stride = -(-used // 64) * 64
    return [b * stride for b in range(int(nbuf))], max(1, int(nbuf)) * stride


class _DevicePtr:
    """Zero-copy torch view of a raw device allocation (``__cuda_array_interface__``)."""

    def __init__(self, ptr: int, n_floats: int):
        self.__cuda_array_interface__ = {"shape": (n_floats,), "typestr": "<f4", "data": (int(ptr), False),
                                         "version": 2}


class PeerGather:
    """The fused march + all-gather: gathered-ranges buffers on every GPU (``world * slot_rays`` floats
    each) mapped into every process, so that ``rl_calc_range_fan_allgather`` /
    ``rl_calc_range_repeat_angles_allgather`` can store each range straight into slot ``rank`` of all
    ``world`` buffers over NVLink while it marches (no separate collective, no staging copy).
    ``sync()`` is the stream-ordered barrier after which every rank may read ``tensor()``.

    Reuse: there are ``nbuf`` (default 2) sets of buffers and consecutive calls alternate between them.
    A faster rank's next march starts storing into every GPU's buffer as soon as ITS stream gets there, so
    with one set it could overwrite ranges a slower rank is still reading (write-after-read across ranks);
    with two, a set is rewritten only after the barrier of the call in between, which no rank passes before
    every rank has finished that call's march -- and a rank enqueues that march after its own consumers of
    the earlier result.  ``tensor()`` of a call therefore stays valid until the call after next ON ANY
    RANK; consumers must be enqueued on the marching stream (or synchronised with it) before the next call.

    backend "symm":  torch symmetric memory does the plumbing (allocation, rendezvous, signal-pad
                     barrier); with ``multicast=True`` and NVLS support the kernel issues
                     ``multimem.st`` to the multicast address and the NVSwitch replicates it.
    backend "ipc":   buffers allocated by the C ABI and exchanged as CUDA IPC handles; barrier = a
                     4-byte NCCL all_reduce.  Used when symmetric memory is unavailable.
    """

    def __init__(self, device_index: int, slot_rays: int, group: Optional[dist.ProcessGroup] = None,
                 backend: str = "auto", multicast: Optional[bool] = None, nbuf: int = 2):
        import os
        from . import _native
        # RL_GATHER_MODE = auto | mc (NVLS multicast: multimem.st) | mc_weak (multimem.st.weak) | uc (one store per peer).
        # auto: multicast from 3 GPUs up.  Between 2 GPUs a multimem.st buys nothing (one remote copy either way) and
        # the NVLS path moved 17.7 MB in 0.071 ms against 0.044 ms for plain peer stores; config 5's gathered pieces
        # ran at 92.6 vs 165 Grays/s (profiles/r02_bench_n2*.json).  At 8 GPUs multicast wins (0.226 vs 0.284 ms).
        gmode = os.environ.get("RL_GATHER_MODE", "auto")
        self.world = dist.get_world_size(group)
        if multicast is None:
            multicast = gmode in ("mc", "mc_weak") or (gmode == "auto" and self.world >= 3)
        self._weak = gmode == "mc_weak"
        self._native = _native
        self.group = group
        self.rank = dist.get_rank(group)
        self.device_index = int(device_index)
        self.slot_rays = int(slot_rays)
        self.nbuf = max(1, int(nbuf))
        self.flags = 0
        self._hdl = None
        self._own = None
        self._opened = []
        self._next = 0
        self._last = 0
        self.buf_floats = self.world * self.slot_rays
        self._set_offsets, n_floats = gather_set_layout(self.world, self.slot_rays, self.nbuf)
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
        import ctypes as C
        # per buffer set: the `world` base pointers the kernel stores through
        self._ptr_sets = [(C.c_void_p * self.world)(*[p + off * 4 for p in self._base_ptrs]) for off in self._set_offsets]
        self.ptrs = self._ptr_sets[0]

    def _init_symm(self, n_floats, multicast):
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
            self.flags = 3 if self._weak else 1   # RL_GATHER_MULTICAST (| RL_GATHER_WEAK)
            self._base_ptrs = [mc] + ptrs[1:]
        else:
            self._base_ptrs = ptrs
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
        self._base_ptrs = ptrs
        self._flag = torch.zeros(1, dtype=torch.int32, device=f"cuda:{self.device_index}")
        self._view = torch.as_tensor(_DevicePtr(own.value, n_floats), device=f"cuda:{self.device_index}")

    @property
    def multicast(self) -> bool:
        return bool(self.flags & 1)

    def tensor(self, buf: Optional[int] = None) -> torch.Tensor:
        """This GPU's gathered buffer of the last call (or of set ``buf``): (world * slot_rays,) float32,
        slot r = ranges of rank r."""
        b = self._last if buf is None else int(buf)
        return self._view[self._set_offsets[b]:self._set_offsets[b] + self.buf_floats]

    def _take(self, buf):
        b = self._next if buf is None else int(buf)
        self._last = b
        self._next = (b + 1) % self.nbuf
        return self._ptr_sets[b]

    def march(self, marcher, poses: torch.Tensor, fov: float, num_rays: int, stream_ptr: int,
              buf: Optional[int] = None):
        """Fan scan of this rank's poses into slot ``rank`` of every GPU's next buffer set."""
        n = poses.shape[0]
        self._native.check(self._native.lib().rl_calc_range_fan_allgather(
            marcher._h, poses.data_ptr(), 1, self._take(buf), self.world, self.rank, self.slot_rays, n, int(num_rays),
            float(fov), self.flags, stream_ptr), "rl_calc_range_fan_allgather")

    def march_angles(self, marcher, poses: torch.Tensor, angles: torch.Tensor, stream_ptr: int,
                     buf: Optional[int] = None):
        """calc_range_repeat_angles of this rank's poses into slot ``rank`` of every GPU's next buffer set."""
        n, na = poses.shape[0], angles.shape[0]
        self._native.check(self._native.lib().rl_calc_range_repeat_angles_allgather(
            marcher._h, poses.data_ptr(), angles.data_ptr(), self._take(buf), self.world, self.rank, self.slot_rays,
            n, int(na), self.flags, stream_ptr), "rl_calc_range_repeat_angles_allgather")

    def gather(self, ranges: torch.Tensor, stream_ptr: int, buf: Optional[int] = None):
        """The all-gather alone: ranges that already exist on this GPU go to slot ``rank`` of every GPU's next
        buffer set through the same NVLink stores (``rl_allgather_ranges``)."""
        self._native.check(self._native.lib().rl_allgather_ranges(
            self.device_index, ranges.data_ptr(), self._take(buf), self.world, self.rank, self.slot_rays,
            ranges.numel(), self.flags, stream_ptr), "rl_allgather_ranges")

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


def gpu_march_angles_fn(marcher, angles: torch.Tensor) -> Callable:
    """The product march of the particle-filter shape (BASELINE config 3):
    ``PyRayMarchingGPU.calc_range_repeat_angles`` on device tensors; ``num_rays = len(angles)``."""
    def run(poses: torch.Tensor, out: torch.Tensor):
        if poses.shape[0]:
            marcher.calc_range_repeat_angles(poses, angles, out)
    return run


class ShardedScanner:
    """``scan(poses)`` over the whole process group.

    poses:   (N, 3) float32, identical on every rank (only the rank's own rows are read).
    returns: (N * num_rays,) float32 ranges, pose-major -- on every rank for ``gather="all"``, on
             ``root`` only (None elsewhere) for ``gather="root"``, and just the local shard
             (``hi - lo`` poses) for ``gather="none"``.
    chunks:  ``gather="all"`` only: the rank's shard is marched in this many pieces and the all-gather of
             piece k-1 (``async_op`` on the collective's own stream) overlaps the march of piece k
             (SURVEY.md section 5).  The fused path (:meth:`scan_fused`) needs no such pipeline: its
             stores travel while the march runs.
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
        self._peer = None

    def scan(self, poses: torch.Tensor, gather: str = "all", root: int = 0):
        if gather not in ("all", "root", "none"):
            raise ValueError("gather must be 'all', 'root' or 'none'")
        n = poses.shape[0]
        lo, hi = shard_bounds(n, self.world, self.rank)
        per = -(-n // self.world) if n > 0 else 0
        R = self.num_rays
        cnt = hi - lo
        # equal, padded shards: all_gather needs the same count from every rank
        local = torch.zeros(per * R, dtype=torch.float32, device=self.device)
        mine = poses[lo:hi].to(self.device).contiguous()
        if gather == "all" and self.world > 1 and self.chunks > 1 and per > 0:
            full = torch.empty(self.world * per * R, dtype=torch.float32, device=self.device)
            C = min(self.chunks, per)
            cuts = [per * i // C for i in range(C + 1)]   # in poses of the padded shard
            pending = []
            for a, b in zip(cuts[:-1], cuts[1:]):
                m0, m1 = min(a, cnt), min(b, cnt)
                if m1 > m0:
                    self.march_fn(mine[m0:m1], local[m0 * R:m1 * R])
                outs = [full[r * per * R + a * R: r * per * R + b * R] for r in range(self.world)]
                pending.append(dist.all_gather(outs, local[a * R:b * R], group=self.group, async_op=True))
            for w in pending:
                w.wait()
            return full[:n * R]
        self.march_fn(mine, local[:cnt * R])
        if gather == "none" or self.world == 1:
            out = local[:cnt * R]
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

    def _peer_for(self, peer, need_rays):
        if peer is not None:
            return peer
        if self._peer is None or self._peer.slot_rays < need_rays:
            if self._peer is not None:
                self._peer.close()
            self._peer = PeerGather(self.device.index, max(need_rays, 1), self.group)
        return self._peer

    def scan_fused(self, poses: torch.Tensor, marcher, fov: float = 0.0, peer: "PeerGather" = None,
                   angles: Optional[torch.Tensor] = None):
        """All-gather fused into the march kernel (``rl_calc_range_fan_allgather``, or
        ``rl_calc_range_repeat_angles_allgather`` when ``angles`` is given -- then ``num_rays`` must be
        ``len(angles)``): every rank's ranges are stored straight into every GPU's gathered buffer over
        NVLink while the march runs.  Returns the (N * num_rays,) gathered ranges on every rank: a view of
        the peer buffer, valid until the call AFTER NEXT on any rank (see :class:`PeerGather`).  Pass a
        long-lived :class:`PeerGather` to avoid re-creating the mappings."""
        n = poses.shape[0]
        lo, hi = shard_bounds(n, self.world, self.rank)
        per = -(-n // self.world) if n > 0 else 0
        R = self.num_rays
        if angles is not None and angles.shape[0] != R:
            raise ValueError("scan_fused: len(angles) must equal num_rays")
        peer = self._peer_for(peer, per * R)
        mine = poses[lo:hi].to(self.device).contiguous()
        stream = int(torch.cuda.current_stream(self.device.index).cuda_stream)
        if angles is None:
            peer.march(marcher, mine, fov, R, stream)
        else:
            peer.march_angles(marcher, mine, angles, stream)
        peer.sync()
        full = peer.tensor()
        if peer.slot_rays == per * R:
            return full[:n * R]
        # slots padded beyond this batch's shard size: compact the used parts
        return torch.cat([full[r * peer.slot_rays: r * peer.slot_rays + max(0, min(n, (r + 1) * per) - r * per) * R]
                          for r in range(self.world)])

    def scan_fused_chunks(self, poses: torch.Tensor, marcher, fov: float, chunk_poses: int,
                          peer: "PeerGather" = None, angles: Optional[torch.Tensor] = None):
        """Generator over a batch too large to gather into one buffer (BASELINE config 5: 17.3 GB of
        ranges): the rank's shard is marched ``chunk_poses`` poses at a time, each piece all-gathered by the
        fused kernel into the peer buffers (which alternate, so piece k+1 is marched while piece k is being
        consumed) and yielded as ``(first_pose_in_shard, count_per_rank, gathered)`` with ``gathered`` the
        (world, chunk_poses * num_rays) view: row r = ranges of rank r's poses
        ``[r*per + first, r*per + first + count_r)``, where ``count_r`` is ``count_per_rank[r]``."""
        n = poses.shape[0]
        lo, hi = shard_bounds(n, self.world, self.rank)
        per = -(-n // self.world) if n > 0 else 0
        R = self.num_rays
        chunk_poses = max(1, int(chunk_poses))
        peer = self._peer_for(peer, chunk_poses * R)
        mine = poses[lo:hi].to(self.device).contiguous()
        stream = int(torch.cuda.current_stream(self.device.index).cuda_stream)
        for first in range(0, per, chunk_poses):
            a, b = min(first, hi - lo), min(first + chunk_poses, hi - lo)
            if b > a:
                if angles is None:
                    peer.march(marcher, mine[a:b], fov, R, stream)
                else:
                    peer.march_angles(marcher, mine[a:b], angles, stream)
            else:
                peer._take(None)   # nothing of mine in this piece: still advance to the buffer set the others write
            peer.sync()
            counts = [max(0, min(first + chunk_poses, min(n, (r + 1) * per) - r * per) - first) for r in range(self.world)]
            yield first, counts, peer.tensor().view(self.world, peer.slot_rays)


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
