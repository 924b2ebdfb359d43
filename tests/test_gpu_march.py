"""The march kernels (csrc/march.cu) through the range_libc-compatible shim, against the oracle
on the same seeded inputs and against the committed golden vectors.  Bar (north_star): every
range within max(1e-4 rel, 0.5 cell), >= 99.9 % of beams bit-identical, max_range / out-of-map
handling identical."""
import numpy as np
import pytest

from pyracecarsimulator_b200 import maps, range_libc
from gpu_util import assert_ranges_match, build_synth

pytestmark = pytest.mark.gpu

FOV = 4.71


@pytest.fixture(scope="module")
def col(orc, colombia, colombia_scan):
    grid = colombia_scan["grid"]
    binar = np.where(grid > 0, 255, 0).ravel()
    omap = range_libc.PyOMap(maps.OccupancyGrid.make(binar, 435, 350, colombia["resolution"], colombia["origin"]))
    dist = orc.edt_float(colombia_scan["occ"])
    return dict(omap=omap, rm=range_libc.PyRayMarchingGPU(omap, 300),
                orc=orc.Marcher(dist, 300, colombia["resolution"], colombia["origin"]), dist=dist,
                res=colombia["resolution"], origin=colombia["origin"])


@pytest.fixture(scope="module")
def big(orc):
    omap, y, occ, dist = build_synth(orc, 2049, 1234)
    return dict(omap=omap, rm=range_libc.PyRayMarchingGPU(omap, 300),
                orc=orc.Marcher(dist, 300, y.resolution, y.origin), dist=dist, res=y.resolution,
                origin=y.origin)


def test_device_trig_matches_twin(orc):
    """csrc/glibc_trig.cuh on the device == oracle/trig_twin.c == host libm, bit for bit."""
    import ctypes
    import torch
    from pyracecarsimulator_b200 import _native
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-20, 20, 2_000_000), rng.uniform(-130, 130, 200_000),
                        rng.standard_normal(100_000) * 1e-3, rng.standard_normal(50_000) * 1e5,
                        rng.uniform(-1e-5, 1e-5, 10_000),
                        [0.0, -0.0, np.inf, -np.inf, np.nan, 120.0, -120.0, 119.99999, 0.785398185, 1e38, -3e38,
                         -4.712389]]).astype(np.float32)
    d = torch.from_numpy(x).cuda()
    s, c = torch.empty_like(d), torch.empty_like(d)
    _native.check(_native.lib().rl_probe_sincosf(d.data_ptr(), s.data_ptr(), c.data_ptr(), d.numel(), None))
    torch.cuda.synchronize()
    with np.errstate(invalid="ignore"):
        ws, wc = np.sin(x), np.cos(x)     # numpy float32 sin/cos call libm sinf/cosf... checked below
    got_s, got_c = s.cpu().numpy(), c.cpu().numpy()
    twin = np.array([orc.twin_sincosf(v) for v in x[::997]], dtype=np.float32)
    assert np.array_equal(got_s[::997].view(np.uint32)[~np.isnan(twin[:, 0])], twin[:, 0].view(np.uint32)[~np.isnan(twin[:, 0])])
    assert np.array_equal(got_c[::997].view(np.uint32)[~np.isnan(twin[:, 1])], twin[:, 1].view(np.uint32)[~np.isnan(twin[:, 1])])
    assert np.isnan(got_s[np.isinf(x) | np.isnan(x)]).all() and np.isnan(got_c[np.isinf(x) | np.isnan(x)]).all()
    # ranges then inherit bit identity: identical ray directions + IEEE-exact march arithmetic
    del ws, wc


def test_golden_fan(col, colombia_scan):
    g = colombia_scan
    out = np.zeros(g["fan"].size, np.float32)
    col["rm"].calc_range_fan(g["poses"], out, FOV, 1080)
    assert_ranges_match(out, g["fan"], col["res"])
    # out-of-map / max_range beams are exactly max_range * scale
    assert np.array_equal(out == np.float32(15.0), g["fan"] == np.float32(15.0))


def test_golden_many_and_repeat_angles(col, colombia_scan):
    g = colombia_scan
    out = np.zeros(g["many"].size, np.float32)
    col["rm"].calc_range_many(g["rays"], out)
    assert_ranges_match(out, g["many"], col["res"])
    out = np.zeros(g["rep"].size, np.float32)
    col["rm"].calc_range_repeat_angles(g["poses"], g["angles"], out)
    assert_ranges_match(out, g["rep"], col["res"])


def test_survey_known_answer(col):
    out = np.zeros(1080, np.float32)
    ins = np.zeros((1080, 3), np.float32)
    ins[0] = (0.275, 0.0, 0.0)
    col["rm"].calc_range_many(ins, out, FOV, 1080)           # the fork's single-pose call
    for j, v in {0: 1.2391100, 270: 1.1963634, 540: 3.7185178, 810: 1.8073667, 1079: 3.5077786}.items():
        assert abs(out[j] - v) < 1e-6
    assert abs(out.sum(dtype=np.float64) - 3282.2590) < 2e-2


def test_fork_layout_ignores_dead_rows(col, colombia_scan):
    poses = colombia_scan["poses"][:7]
    wide = np.full((7 * 1080, 3), 77.0, np.float32)
    wide[::1080] = poses
    out = np.zeros(7 * 1080, np.float32)
    col["rm"].calc_range_many(wide, out, FOV, 1080)
    assert_ranges_match(out, colombia_scan["fan"][:7 * 1080], col["res"])


@pytest.mark.parametrize("num_rays", [1, 31, 32, 33, 60, 270, 1080, 1081, 4097])
def test_fan_ragged_beam_counts(col, num_rays):
    poses = maps.sample_free_poses(col["dist"], 37, num_rays, col["res"], col["origin"])
    out = np.full(37 * num_rays + 5, -1.0, np.float32)
    col["rm"].calc_range_fan(poses, out[:37 * num_rays], FOV, num_rays)
    assert_ranges_match(out[:37 * num_rays], col["orc"].calc_range_fan(poses, num_rays, FOV), col["res"])
    assert np.all(out[37 * num_rays:] == -1.0)   # nothing written past the end


def test_empty_batches(col):
    col["rm"].calc_range_fan(np.zeros((0, 3), np.float32), np.zeros(0, np.float32), FOV, 1080)
    col["rm"].calc_range_many(np.zeros((0, 3), np.float32), np.zeros(0, np.float32))
    col["rm"].calc_range_repeat_angles(np.zeros((0, 3), np.float32), np.zeros(4, np.float32), np.zeros(0, np.float32))
    col["rm"].calc_range_repeat_angles(np.zeros((3, 3), np.float32), np.zeros(0, np.float32), np.zeros(0, np.float32))


def test_out_of_map_nan_and_inside_wall(col, orc):
    rays = np.array([[-100.0, 0.0, 0.3], [1e9, 1e9, 0.0], [np.nan, 0.0, 0.0], [0.0, np.nan, 0.0],
                     [0.275, 0.0, np.nan], [np.inf, 0.0, 0.0], [-5.70654, -2.020793, 0.5],
                     [-5.70654 - 0.03, -2.020793 + 1.0, 0.0]], np.float32)
    out = np.zeros(len(rays), np.float32)
    col["rm"].calc_range_many(rays, out)
    want = col["orc"].calc_range_many(rays)
    assert np.array_equal(out, want), (out, want)
    assert np.all(out[:6] == np.float32(15.0))


def test_calc_range_single(col, colombia_scan):
    p = colombia_scan["rays"][100]
    assert col["rm"].calc_range(*p) == pytest.approx(float(colombia_scan["many"][100]), abs=1e-6)


def test_mcts_rollout_batch_on_stand_in_map(big):
    # BASELINE config 2 at reduced pose count: the oracle finishes in seconds
    poses = maps.sample_free_poses(big["dist"], 512, 42, big["res"], big["origin"])
    out = np.zeros(512 * 1080, np.float32)
    big["rm"].calc_range_fan(poses, out, FOV, 1080)
    want, steps = big["orc"].calc_range_fan(poses, 1080, FOV, steps=True, threads=0)
    same = assert_ranges_match(out, want, big["res"])
    print(f"config-2 slice: {same*100:.4f}% bit-identical, {steps.mean():.2f} steps/ray")


def test_particle_filter_shape_repeat_angles(big):
    # BASELINE config 3 shape: many poses x 60 angles
    poses = maps.sample_free_poses(big["dist"], 20000, 43, big["res"], big["origin"])
    angles = np.linspace(-FOV / 2, FOV / 2, 60, endpoint=False).astype(np.float32)
    out = np.zeros(20000 * 60, np.float32)
    big["rm"].calc_range_repeat_angles(poses, angles, out)
    assert_ranges_match(out, big["orc"].calc_range_repeat_angles(poses, angles, threads=0), big["res"])


def test_step_counter_matches_oracle(big):
    poses = maps.sample_free_poses(big["dist"], 64, 44, big["res"], big["origin"])
    out = np.zeros(64 * 1080, np.float32)
    big["rm"].count_steps(True)
    try:
        big["rm"].calc_range_fan(poses, out, FOV, 1080)
        got = big["rm"].last_steps()
    finally:
        big["rm"].count_steps(False)
    _, steps = big["orc"].calc_range_fan(poses, 1080, FOV, steps=True)
    assert abs(got - int(steps.sum())) <= 0.002 * steps.sum()


def test_full_size_config2_properties(big):
    # BASELINE config 2 at full size (4096 x 1080): size-independent properties + oracle parity
    poses = maps.sample_free_poses(big["dist"], 4096, 45, big["res"], big["origin"])
    out = np.zeros(4096 * 1080, np.float32)
    big["rm"].calc_range_fan(poses, out, FOV, 1080)
    assert np.all(np.isfinite(out)) and out.min() >= 0.0
    assert out.max() <= 15.0 + 2 * 0.05 * 1.5           # hit ranges may exceed max_range slightly (A.6)
    again = np.zeros_like(out)
    big["rm"].calc_range_fan(poses, again, FOV, 1080)
    assert np.array_equal(out, again)                    # deterministic
    # a permutation of the poses permutes the scans
    perm = np.random.default_rng(0).permutation(4096)
    pout = np.zeros_like(out)
    big["rm"].calc_range_fan(np.ascontiguousarray(poses[perm]), pout, FOV, 1080)
    assert np.array_equal(pout.reshape(4096, 1080), out.reshape(4096, 1080)[perm])
    want = big["orc"].calc_range_fan(poses, 1080, FOV, threads=0)
    assert_ranges_match(out, want, big["res"])


def test_torch_device_tensors_match_host_path(big):
    import torch
    poses = maps.sample_free_poses(big["dist"], 300, 46, big["res"], big["origin"])
    host = np.zeros(300 * 1080, np.float32)
    big["rm"].calc_range_fan(poses, host, FOV, 1080)
    dp = torch.from_numpy(poses).cuda()
    do = torch.zeros(300 * 1080, dtype=torch.float32, device="cuda")
    big["rm"].calc_range_fan(dp, do, FOV, 1080)
    assert np.array_equal(do.cpu().numpy(), host)
    # upstream shapes on device too
    wide = torch.zeros((300 * 1080, 3), dtype=torch.float32, device="cuda")
    wide[::1080] = dp
    do2 = torch.zeros_like(do)
    big["rm"].calc_range_many(wide, do2, FOV, 1080)
    assert torch.equal(do, do2)
    rows = torch.from_numpy(poses).cuda()
    o3 = torch.zeros(300, dtype=torch.float32, device="cuda")
    big["rm"].calc_range_many(rows, o3)
    h3 = np.zeros(300, np.float32)
    big["rm"].calc_range_many(poses, h3)
    assert np.array_equal(o3.cpu().numpy(), h3)
    with pytest.raises(ValueError):
        big["rm"].calc_range_fan(dp, host, FOV, 1080)     # mixed host/device
    with pytest.raises(ValueError):
        big["rm"].calc_range_fan(dp.double(), do, FOV, 1080)


def test_reused_numpy_buffers_get_page_locked_and_results_do_not_change(big):
    """Plain (pageable) numpy buffers passed twice are registered by the shim and then written in place
    by the kernel; a first-time buffer goes through the staged path.  Same ranges either way, including
    the 2-arg row-per-ray form whose INPUT block is the large one."""
    import time
    reg = range_libc._HOST_REGISTRY
    range_libc.release_host_buffers()
    n = 400
    poses = maps.sample_free_poses(big["dist"], n, 77, big["res"], big["origin"])
    want = big["orc"].calc_range_fan(poses, 1080, FOV, threads=0)
    outs = np.zeros(n * 1080, np.float32)                      # 1.7 MB, pageable
    times = []
    for i in range(4):
        outs[:] = -1.0
        t0 = time.perf_counter()
        big["rm"].calc_range_fan(poses, outs, FOV, 1080)
        times.append(time.perf_counter() - t0)
        assert np.array_equal(outs, want), i
        assert (reg.registered_bytes() > 0) == (i >= 1)          # registered on the second sighting
    assert reg.registered_bytes() == outs.nbytes
    view = outs[: 200 * 1080]                                     # a shorter view of registered memory: not re-registered
    view[:] = -1.0
    big["rm"].calc_range_fan(poses[:200], view, FOV, 1080)
    assert np.array_equal(view, want[: 200 * 1080]) and reg.registered_bytes() == outs.nbytes
    # 2-arg form: (N, 3) rows in, N ranges out
    rows = np.zeros((n * 1080, 3), np.float32)                    # 5.2 MB input block
    inc = np.float32(FOV) / np.float32(1080)
    ang = (np.arange(1080, dtype=np.float32) * inc + np.float32(-0.5) * np.float32(FOV)).astype(np.float32)
    rows[:, 0] = np.repeat(poses[:, 0], 1080)
    rows[:, 1] = np.repeat(poses[:, 1], 1080)
    rows[:, 2] = (poses[:, 2][:, None] + ang[None, :]).astype(np.float32).ravel()
    want2 = big["orc"].calc_range_many(rows, threads=0)
    o2 = np.zeros(n * 1080, np.float32)
    for i in range(3):
        o2[:] = -1.0
        big["rm"].calc_range_many(rows, o2)
        assert np.array_equal(o2, want2), i
    assert reg.registered_bytes() == outs.nbytes + rows.nbytes + o2.nbytes
    range_libc.release_host_buffers()
    assert reg.registered_bytes() == 0
    outs[:] = -1.0
    big["rm"].calc_range_fan(poses, outs, FOV, 1080)            # pageable again: staged path
    assert np.array_equal(outs, want)
    range_libc.release_host_buffers()
    print("pageable outs, call times (ms):", [round(t * 1e3, 3) for t in times])


def test_rotated_origin_map(orc):
    # non-zero origin yaw exercises the world->grid rotation (A.4)
    rng = np.random.default_rng(8)
    occ = np.zeros((200, 260), np.uint8)
    occ[:3, :] = occ[-3:, :] = occ[:, :3] = occ[:, -3:] = 1
    for _ in range(10):
        r, c = rng.integers(10, 180), rng.integers(10, 240)
        occ[r:r + 6, c:c + 9] = 1
    origin = (1.5, -2.0, 0.6)
    msg = maps.OccupancyGrid.make(np.where(occ, 100, 0).astype(np.int8).ravel(), 260, 200, 0.1, origin)
    omap = range_libc.PyOMap(msg)
    dist = orc.edt_float(occ)
    assert np.array_equal(omap.dist(), dist)
    rm = range_libc.PyRayMarching(omap, 150)
    m = orc.Marcher(dist, 150, 0.1, origin)
    # world poses: rotate grid-frame samples by +yaw about the origin
    gx, gy = rng.uniform(20, 240, 400) * 0.1, rng.uniform(20, 180, 400) * 0.1
    c, s = np.cos(0.6), np.sin(0.6)
    rays = np.stack([origin[0] + c * gx - s * gy, origin[1] + s * gx + c * gy,
                     rng.uniform(-np.pi, np.pi, 400)], axis=1).astype(np.float32)
    out = np.zeros(400, np.float32)
    rm.calc_range_many(rays, out)
    assert_ranges_match(out, m.calc_range_many(rays), 0.1)


def test_fused_allgather_kernel_two_virtual_ranks(big):
    """rl_calc_range_fan_allgather on one GPU: two "ranks" (two gathered buffers on the same device)
    each march their own shard and store into both buffers; both end up equal to the concatenation
    of the two plain scans.  (The real multi-process path over NVLink is checked inside bench.py.)"""
    import ctypes as C
    import torch
    from pyracecarsimulator_b200 import _native
    L = _native.lib()
    B, R = 64, 1080
    poses = [maps.sample_free_poses(big["dist"], B, 900 + r, big["res"], big["origin"]) for r in range(2)]
    slot = B * R + 128                      # padded slots, like uneven shards
    bufs = []
    for r in range(2):
        p, h = C.c_void_p(), (C.c_uint8 * 64)()
        _native.check(L.rl_peer_alloc(0, 2 * slot * 4, C.byref(p), h))
        bufs.append(p)
    ptrs = (C.c_void_p * 2)(bufs[0].value, bufs[1].value)
    try:
        for r in range(2):
            dp = torch.from_numpy(poses[r]).cuda()
            _native.check(L.rl_calc_range_fan_allgather(big["rm"]._h, dp.data_ptr(), 1, ptrs, 2, r, slot, B, R, FOV, 0, None))
        torch.cuda.synchronize()
        want = []
        for r in range(2):
            o = np.zeros(B * R, np.float32)
            big["rm"].calc_range_fan(poses[r], o, FOV, R)
            want.append(o)
        for r in range(2):
            from pyracecarsimulator_b200.sharded import _DevicePtr   # zero-copy view of the raw allocation
            got = torch.as_tensor(_DevicePtr(bufs[r].value, 2 * slot), device="cuda").cpu().numpy()
            assert np.array_equal(got[:B * R], want[0]) and np.array_equal(got[slot:slot + B * R], want[1])
        # argument checks: slot too small, bad rank, too many peers
        assert L.rl_calc_range_fan_allgather(big["rm"]._h, dp.data_ptr(), 1, ptrs, 2, 0, 10, B, R, FOV, 0, None) == _native.RL_ERR_BAD_ARG
        assert L.rl_calc_range_fan_allgather(big["rm"]._h, dp.data_ptr(), 1, ptrs, 2, 2, slot, B, R, FOV, 0, None) == _native.RL_ERR_BAD_ARG
        assert L.rl_calc_range_fan_allgather(big["rm"]._h, dp.data_ptr(), 1, ptrs, 17, 0, slot, B, R, FOV, 0, None) == _native.RL_ERR_BAD_ARG
    finally:
        for p in bufs:
            L.rl_peer_free(0, p)


def test_device_calls_on_concurrent_streams(big):
    """Device-pointer calls are stateless: several host threads enqueue scans of different batches on
    their own CUDA streams through ONE marcher; every result equals the serial one."""
    import threading
    import torch
    batches = [torch.from_numpy(maps.sample_free_poses(big["dist"], 256, 700 + i, big["res"], big["origin"])).cuda()
               for i in range(4)]
    serial = []
    for p in batches:
        o = torch.empty(256 * 1080, dtype=torch.float32, device="cuda")
        big["rm"].calc_range_fan(p, o, FOV, 1080)
        serial.append(o)
    torch.cuda.synchronize()
    outs = [torch.zeros(256 * 1080, dtype=torch.float32, device="cuda") for _ in batches]
    errs = []

    def work(i):
        try:
            s = torch.cuda.Stream()
            with torch.cuda.stream(s):
                for _ in range(10):
                    big["rm"].calc_range_fan(batches[i], outs[i], FOV, 1080)
            s.synchronize()
        except Exception as e:  # pragma: no cover
            errs.append(e)

    ts = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs, errs
    for a, b in zip(outs, serial):
        assert torch.equal(a, b)
