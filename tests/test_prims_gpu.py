"""Device primitives behind the pipelines (single-pass look-back scans incl. the non-commutative max-plus composition,
onesweep stable pair sort, sparse-table RMQ / previous-smaller) against the host's std:: algorithms, bit-exact,
at ragged sizes around the tile boundaries and at cfg2-like sizes."""
import pytest

pytestmark = pytest.mark.gpu


def test_prims_small_and_ragged(gpu):
    for n in [0, 1, 2, 31, 32, 33, 255, 256, 257, 2047, 2048, 2049, 4095, 4096, 4097, 8191, 8193, 100_003]:
        for bits in (1, 8, 9, 21, 32):
            assert gpu.selftest_prims(n, seed=2 * n + bits, bits=bits) == 0, (n, bits)
            assert gpu.selftest_prims(n, seed=2 * n + bits + 1, bits=bits) == 0, (n, bits, "sorted-ish")


def test_prims_large(gpu):
    assert gpu.selftest_prims(3_000_017, seed=4, bits=22) == 0      # 8-item tiles (fits one wave group)
    assert gpu.selftest_prims(12_000_001, seed=7, bits=20) == 0     # 16-item tiles, path-step sized, sorted-ish keys
    assert gpu.selftest_prims(5_000_000, seed=10, bits=32) == 0
    for rep in range(20):                                           # repeated launches reuse the epoch-tagged scratch
        assert gpu.selftest_prims(300_000 + rep, seed=100 + rep, bits=32) == 0


def check_fired_arcs(pv):
    """Which class re-encounters still find their previous occurrence on the reference's stack (decompose D6): the device
    fixpoint against the stack machine itself on random class sequences full of crossing arcs.  Real class stacks have no
    crossing arcs at all (LPA, every fixture, thousands of random graphs: 0 rounds), so this is the only place the rounds run."""
    settled = 0
    for n, k in [(10, 3), (50, 10), (200, 30), (1000, 100), (3000, 1500)]:
        for seed in range(12):
            bad, rounds = pv.selftest_arcs(n, k, seed * 977 + n)
            assert bad == 0, (n, k, seed, bad, rounds)
            settled += rounds >= 0
    assert settled >= 40


def test_gpu_fired_arcs_fixpoint(gpu):
    check_fired_arcs(gpu)
