"""The rollout action generator's CPU restatement (oracle/philox_oracle.c) against the published
Philox4x32-10 known answers, an independent numpy restatement, and the numpy uniform() mapping.
The reference draws its rollout actions from numpy's unseeded global generator
(scripts/mcts.py:216-222), so the pin here is the generator's own known-answer vectors."""
import numpy as np

import oracle

# (counter, key) -> output: zero block, all-ones block, pi-digits block (Random123's known-answer
# set for philox4x32-10; reproduced with the CUDA toolkit's curand_Philox4x32_10 built for the host)
KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
    ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
    ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
     (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
]


def np_philox(ctr, key):
    """Vectorised numpy restatement: ctr (..., 4) uint64-held words, key (2,)."""
    c = [np.asarray(ctr[..., i], dtype=np.uint64) for i in range(4)]
    k0, k1 = np.uint64(key[0]), np.uint64(key[1])
    m32 = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = np.uint64(0xD2511F53) * c[0]
        p1 = np.uint64(0xCD9E8D57) * c[2]
        c = [(p1 >> np.uint64(32)) ^ c[1] ^ k0, p1 & m32, (p0 >> np.uint64(32)) ^ c[3] ^ k1, p0 & m32]
        k0 = (k0 + np.uint64(0x9E3779B9)) & m32
        k1 = (k1 + np.uint64(0xBB67AE85)) & m32
    return np.stack(c, axis=-1)


def np_actions(n_cars, n_actions, seed, stream_id, car_offset, speed_range, steer_range):
    car = (np.arange(n_cars, dtype=np.uint64) + np.uint64(car_offset))[:, None]
    act = np.arange(n_actions, dtype=np.uint64)[None, :]
    ctr = np.empty((n_cars, n_actions, 4), dtype=np.uint64)
    ctr[..., 0] = act
    ctr[..., 1] = car & np.uint64(0xFFFFFFFF)
    ctr[..., 2] = car >> np.uint64(32)
    ctr[..., 3] = stream_id
    w = np_philox(ctr, (seed & 0xFFFFFFFF, seed >> 32))

    def unit(a, b):
        return ((a >> np.uint64(5)).astype(np.float64) * 67108864.0 + (b >> np.uint64(6)).astype(np.float64)) / 9007199254740992.0

    out = np.empty((n_cars, n_actions, 2), dtype=np.float64)
    out[..., 1] = steer_range[0] + (steer_range[1] - steer_range[0]) * unit(w[..., 0], w[..., 1])
    out[..., 0] = speed_range[0] + (speed_range[1] - speed_range[0]) * unit(w[..., 2], w[..., 3])
    return out


def test_known_answers():
    for ctr, key, want in KAT:
        got = oracle.philox4x32_10(ctr, key)
        assert tuple(int(x) for x in got) == want
        assert tuple(int(x) for x in np_philox(np.array(ctr, dtype=np.uint64), key)) == want


def test_actions_match_numpy_restatement():
    for (n, a, seed, sid, off) in [(1, 1, 42, 0, 0), (257, 5, 42, 0, 0), (64, 7, 2**40 + 3, 9, 2**33 + 5),
                                   (1000, 5, 0, 0xFFFFFFFF, 123456789)]:
        got = oracle.rollout_actions(n, a, seed=seed, stream_id=sid, car_offset=off)
        want = np_actions(n, a, seed, sid, off, (0.0, 7.0), (-0.4189, 0.4189))
        assert np.array_equal(got, want)


def test_actions_are_shape_independent_and_in_range():
    full = oracle.rollout_actions(4096, 5, seed=42)
    part = oracle.rollout_actions(1024, 5, seed=42, car_offset=2048)
    assert np.array_equal(full[2048:3072], part)           # a rank's slice == the same cars of the whole job
    assert np.array_equal(full[:, :3], oracle.rollout_actions(4096, 3, seed=42))
    assert full[..., 0].min() >= 0.0 and full[..., 0].max() < 7.0
    assert full[..., 1].min() >= -0.4189 and full[..., 1].max() < 0.4189
    assert np.unique(full.reshape(-1)).size == full.size  # 53-bit variates: no repeats in 40 960 draws
    assert not np.array_equal(full, oracle.rollout_actions(4096, 5, seed=43))
    assert not np.array_equal(full, oracle.rollout_actions(4096, 5, seed=42, stream_id=1))
    # loose uniformity check (each mean is within 5 sigma)
    assert abs(full[..., 0].mean() - 3.5) < 5 * 7.0 / np.sqrt(12 * 20480)
    assert abs(full[..., 1].mean()) < 5 * 0.8378 / np.sqrt(12 * 20480)


def test_seed42_first_values_are_frozen():
    """Golden values of the shipped default (seed 42, stream 0): a change of block layout, word
    assignment or double conversion shows up here."""
    got = oracle.rollout_actions(2, 2, seed=42)
    want = np_actions(2, 2, 42, 0, 0, (0.0, 7.0), (-0.4189, 0.4189))
    assert np.array_equal(got, want)
    frozen = np.load(__file__.replace("test_philox_oracle.py", "golden/philox_seed42.npy"))
    assert np.array_equal(oracle.rollout_actions(8, 5, seed=42), frozen)
