"""CPU: the restated Boost 1.54 unordered_map iteration order (sibelia_b200/csrc/boost_order.h, reached through the
host-only test hook of the C ABI) against the reference's vendored header (oracle/_ref)."""
import numpy as np
import pytest

from oracle import ref
from sibelia_b200 import binding


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (needs /root/reference)")
def test_iteration_order_matches_vendored_boost(built):
    rng = np.random.default_rng(3)
    for n in list(range(0, 40)) + [63, 64, 65, 100, 129, 500, 1025, 5000]:
        for space in (50, 100_000, 2 ** 32):
            if n > space:
                continue
            keys = rng.choice(space, size=n, replace=False).astype(np.uint64) if space < 10 ** 7 else \
                np.unique(rng.integers(0, space, size=n * 2, dtype=np.uint64))[:n]
            rng.shuffle(keys)
            got = binding.debug_unordered_order(keys)
            want = ref.boost_order(keys)
            assert np.array_equal(got, want), (n, space)


def test_iteration_order_golden(built):
    """Pinned vector (generated with the vendored header): keys 0..19 inserted in order."""
    got = binding.debug_unordered_order(np.arange(20, dtype=np.uint64))
    want = [int(x) for x in open(__file__.replace("test_boost_order.py", "golden/boost_order_0_19.txt")).read().split()]
    assert list(map(int, got)) == want
