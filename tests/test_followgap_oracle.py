"""Follow-the-gap restatement (oracle/followgap_oracle.c) against the unmodified reference header
(oracle/_ref/libfollowgap_ref.so) and the survey's known answer."""
import numpy as np
import pytest


def scans(rng, count):
    for t in range(count):
        n = int(rng.integers(20, 1200))
        l = rng.uniform(0.05, 20, n).astype(np.float32)
        if t % 3 == 0:
            l[rng.integers(0, n, n // 4)] = 0.0
        if t % 5 == 0:
            a = rng.integers(0, n - 5)
            l[a:a + rng.integers(1, n // 2)] = rng.uniform(0.2, 1.7)
        yield l


def test_survey_known_answer(orc):
    # PyFollowGap(10, 15.0, 0.4189, 0.004).eval(l, 1080), l = 3.0, l[200:300] = 1.0
    l = np.full(1080, 3.0, np.float32)
    l[200:300] = 1.0
    assert orc.followgap_eval(l) == 0.4000000059604645


def test_against_live_reference(orc):
    if not orc.followgap_ref_available():
        pytest.skip("oracle/_ref not built (no /root/reference on this box)")
    for l in scans(np.random.default_rng(0), 1500):
        a, b = orc.ref_followgap_eval(l), orc.followgap_eval(l)
        assert a == b or (a != a and b != b)
