"""GPU parity: sibgpu_enumerate (through the C ABI) against the oracle and the committed golden fixtures."""
import glob
import os

import numpy as np
import pytest

import helpers
from oracle import restate
from sibelia_b200 import synth
from test_oracle import GOLDEN_FILES, load_golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("path", GOLDEN_FILES, ids=[os.path.basename(p)[:-4] for p in GOLDEN_FILES])
def test_golden(ctx, path):
    chrs, k, z = load_golden(path)
    got = ctx.enumerate(chrs, k)
    helpers.assert_tables_equal(got, (int(z["count"]), z["pos"], z["neg"]), "golden")


def test_random_tiny(ctx):
    rng = np.random.default_rng(2024)
    for it in range(300):
        chrs, k = helpers.random_case(rng, max_rec=5, max_len=60, kmax=12)
        helpers.assert_tables_equal(ctx.enumerate(chrs, k), restate.enumerate_bifurcations(chrs, k), "case %d k=%d" % (it, k))


def test_edge_cases(ctx):
    A = lambda s: np.frombuffer(s, dtype=np.uint8)
    cases = [
        ([A(b"")], 3), ([A(b""), A(b"")], 2), ([A(b"AC")], 3), ([A(b"ACG")], 3), ([A(b"A" * 100)], 5),
        ([A(b"ACGT" * 30)], 4), ([A(b"ACGT" * 30), A(b"")], 4), ([A(b"AT" * 40)], 6), ([A(b"T" * 64)], 32),
        ([A(b"T" * 64), A(b"A" * 64)], 32), ([A(b"ACGTTGCA" * 8)], 8), ([A(b"G")], 1), ([A(b"GATTACA")], 7),
    ]
    for chrs, k in cases:
        helpers.assert_tables_equal(ctx.enumerate(chrs, k), restate.enumerate_bifurcations(chrs, k), "edge k=%d" % k)


def test_zero_chromosomes(ctx):
    count, pos, neg = ctx.enumerate([], 5)
    assert count == 0 and len(pos) == 0 and len(neg) == 0


@pytest.mark.parametrize("k", [2, 11, 15, 16, 17, 25, 28, 29, 31, 32])
def test_strains_exact_k(ctx, k):
    st = helpers.strain_case(4, 60_000, seed=77)
    helpers.assert_tables_equal(ctx.enumerate(st, k), restate.enumerate_bifurcations(st, k), "strains k=%d" % k)


@pytest.mark.parametrize("k", [33, 34, 48, 63, 64, 65, 100, 128, 1000, 5000])
def test_strains_fingerprint_k(ctx, k):
    """k > 32 goes through the fingerprint path (+ string ranking and per-instance verification)."""
    st = helpers.strain_case(4, 60_000, seed=78)
    helpers.assert_tables_equal(ctx.enumerate(st, k), restate.enumerate_bifurcations(st, k), "strains k=%d" % k)


def test_long_palindromes(ctx):
    """Even k > 32 with k-mers equal to their own reverse complement (one vertex id, instances on both strands)."""
    rng = np.random.default_rng(3)
    w = synth.random_genome(40, 1)
    pal = np.concatenate([w, synth.revcomp(w)])                      # an 80-mer that is its own reverse complement
    x, y, z = (synth.random_genome(300, s) for s in (2, 3, 4))
    chrs = [np.concatenate([x, pal, y]), np.concatenate([z, pal, x[:100]]), synth.revcomp(np.concatenate([y, pal]))]
    for k in (34, 40, 60, 80):
        helpers.assert_tables_equal(ctx.enumerate(chrs, k), restate.enumerate_bifurcations(chrs, k), "pal k=%d" % k)
    for k in (8, 20, 32):
        helpers.assert_tables_equal(ctx.enumerate(chrs, k), restate.enumerate_bifurcations(chrs, k), "pal k=%d" % k)


def test_k_longer_than_everything(ctx):
    st = helpers.strain_case(2, 3000, seed=1)
    count, pos, neg = ctx.enumerate(st, 4000)
    assert count == 0 and len(pos) == 0 and len(neg) == 0


def test_many_small_chromosomes(ctx):
    rng = np.random.default_rng(5)
    base = synth.random_genome(3000, 9)
    chrs = []
    for i in range(200):
        o = int(rng.integers(0, 2500))
        L = int(rng.integers(0, 400))
        piece = base[o:o + L]
        chrs.append(synth.revcomp(piece) if rng.random() < 0.4 else piece.copy())
    for k in (5, 25, 30, 40, 150):
        helpers.assert_tables_equal(ctx.enumerate(chrs, k), restate.enumerate_bifurcations(chrs, k), "many k=%d" % k)


def test_multi_partition_path(built):
    """Force many hash partitions (SIBGPU_PART_RECORDS) so the scatter / per-partition tables are exercised."""
    import sibelia_b200 as sb
    os.environ["SIBGPU_PART_RECORDS"] = "4096"
    try:
        c = sb.Context(0)
    finally:
        del os.environ["SIBGPU_PART_RECORDS"]
    st = helpers.strain_case(4, 100_000, seed=99)
    for k in (25, 31):
        helpers.assert_tables_equal(c.enumerate(st, k), restate.enumerate_bifurcations(st, k), "parts k=%d" % k)
    c.close()


def _ctx_with_env(**env):
    import sibelia_b200 as sb
    old = {k: os.environ.get(k) for k in env}
    os.environ.update({k: str(v) for k, v in env.items()})
    try:
        return sb.Context(0)
    finally:
        for k, v in old.items():
            if v is None:
                del os.environ[k]
            else:
                os.environ[k] = v


def test_partition_overflow_falls_back_to_exact_sizes(built):
    """Fixed-capacity hash partitions (no histogram pass) overflow when one k-mer is repeated very often; the
    enumeration must notice, redo the partitioning with exact sizes and still match the oracle."""
    c = _ctx_with_env(SIBGPU_PART_RECORDS=4096, SIBGPU_PART_SLACK=64)
    rnd = synth.random_genome(60_000, 4)
    poly = np.full(50_000, ord("A"), dtype=np.uint8)                  # 50 000 identical k-mers -> one partition
    chrs = [np.concatenate([rnd[:30_000], poly, rnd[30_000:]]), synth.revcomp(rnd[10_000:40_000])]
    before = c.partition_fallbacks()
    for k in (25, 30):
        helpers.assert_tables_equal(c.enumerate(chrs, k), restate.enumerate_bifurcations(chrs, k), "overflow k=%d" % k)
    assert c.partition_fallbacks() >= before + 2
    # balanced input on the same context: no fallback
    st = helpers.strain_case(4, 50_000, seed=98)
    before = c.partition_fallbacks()
    helpers.assert_tables_equal(c.enumerate(st, 25), restate.enumerate_bifurcations(st, 25), "balanced")
    assert c.partition_fallbacks() == before
    c.close()


def test_bucket_overflow_falls_back_to_l2_tables(built):
    """Shared-memory grouping: a bucket (~1 Ki records, fixed capacity 1664) overflows when one k-mer occurs thousands of
    times; the enumeration must notice and regroup the intact level-1 partitions through the L2-resident tables."""
    import sibelia_b200 as sb
    c = sb.Context(0)
    rnd = synth.random_genome(300_000, 14)
    poly = np.full(6_000, ord("A"), dtype=np.uint8)
    chrs = [np.concatenate([rnd[:100_000], poly, rnd[100_000:]]), synth.revcomp(rnd[50_000:90_000])]
    for k in (25, 28):
        before = c.bucket_fallbacks()
        helpers.assert_tables_equal(c.enumerate(chrs, k), restate.enumerate_bifurcations(chrs, k), "bucket overflow k=%d" % k)
        assert c.bucket_fallbacks() == before + 1
    st = helpers.strain_case(4, 50_000, seed=96)
    before = c.bucket_fallbacks()
    helpers.assert_tables_equal(c.enumerate(st, 25), restate.enumerate_bifurcations(st, 25), "balanced")
    assert c.bucket_fallbacks() == before and c.partition_fallbacks() == 0
    c.close()


def test_vertex_key_list_regrows(built):
    """k_group counts every bifurcation class even when the key list is full; the host regrows it and groups again."""
    c = _ctx_with_env(SIBGPU_CKEYS_INIT=8)
    st = helpers.strain_case(4, 60_000, seed=95)
    for k in (25, 16):
        helpers.assert_tables_equal(c.enumerate(st, k), restate.enumerate_bifurcations(st, k), "regrow k=%d" % k)
    c.close()


@pytest.mark.parametrize("k", [11, 25, 28])
def test_l2_table_grouping_still_matches(built, k):
    """SIBGPU_GROUP_SMEM=0: the per-partition L2-table grouping (the fallback of the shared-memory path)."""
    c = _ctx_with_env(SIBGPU_GROUP_SMEM=0, SIBGPU_PART_RECORDS=16384)
    st = helpers.strain_case(4, 60_000, seed=94)
    helpers.assert_tables_equal(c.enumerate(st, k), restate.enumerate_bifurcations(st, k), "l2 tables k=%d" % k)
    c.close()


def test_exact_histogram_mode_matches(built):
    c = _ctx_with_env(SIBGPU_PART_RECORDS=8192, SIBGPU_EXACT_HIST=1)
    st = helpers.strain_case(3, 70_000, seed=97)
    for k in (25, 32, 40):
        helpers.assert_tables_equal(c.enumerate(st, k), restate.enumerate_bifurcations(st, k), "exact-hist k=%d" % k)
    c.close()


def test_pipelined_upload_many_pieces(ctx):
    """sibgpu_enumerate streams the text in 8 Mi-position pieces (copy / pack / scatter overlapped); chromosome and
    piece boundaries must not matter: same tables as upload + enumerate_resident, which packs in one launch."""
    g = synth.random_genome(9_000_000, 77)

    def mutated(x, seed):
        y = x.copy()
        at = np.random.default_rng(seed).integers(0, len(y), len(y) // 100)
        y[at] = helpers.ACGT[(np.searchsorted(helpers.ACGT, y[at]) + 1) % 4]
        return y
    chrs = [g[:8_388_000], synth.revcomp(mutated(g[1_000_000:1_400_000], 1)), g[8_388_000:],
            mutated(g[8_200_000:8_500_000], 2)]                        # a chromosome boundary near the first piece edge
    for k in (25, 31):
        got = ctx.enumerate(chrs, k)
        ctx.upload(chrs)
        count, ninst = ctx.enumerate_resident(k)
        pos, neg = ctx.download()
        helpers.assert_tables_equal(got, (count, pos, neg), "pipelined vs resident k=%d" % k)
        assert count > 1000
    helpers.assert_tables_equal(got, restate.enumerate_bifurcations(chrs, 31), "pipelined vs oracle")


def test_rejects_unsanitised_input(ctx):
    import sibelia_b200 as sb
    with pytest.raises(sb.SibgpuError) as e:
        ctx.enumerate([np.frombuffer(b"ACGTNNACGT" * 10, dtype=np.uint8)], 5)
    assert e.value.status == 3


def test_rejects_every_illegal_byte_in_every_lane(ctx):
    """the pack kernel's SWAR legality test: any byte other than A, C, G, T must raise SIBGPU_ERR_INPUT, whatever its
    position inside the 16-byte load ('$' is the text separator and therefore not distinguishable on the device)"""
    import sibelia_b200 as sb
    base = np.frombuffer(b"ACGTTGCAGGATCCAT" * 8, dtype=np.uint8)
    ok = set(b"ACGT$")
    for b in range(256):
        if b in ok:
            continue
        a = base.copy()
        a[(b * 7) % len(a)] = b
        with pytest.raises(sb.SibgpuError) as e:
            ctx.enumerate([a], 5)
        assert e.value.status == 3, "byte 0x%02x" % b
    helpers.assert_tables_equal(ctx.enumerate([base], 5), restate.enumerate_bifurcations([base], 5), "legal text")


def test_staged_api_reuses_upload(ctx):
    st = helpers.strain_case(3, 40_000, seed=5)
    ctx.upload(st)
    for k in (20, 25):
        count, ninst = ctx.enumerate_resident(k)
        pos, neg = ctx.download()
        want = restate.enumerate_bifurcations(st, k)
        helpers.assert_tables_equal((count, pos, neg), want, "staged k=%d" % k)
        assert ninst == len(pos)


def test_large_random_properties(ctx):
    """BASELINE configs[1] shape at reduced length (16 Mb random single contig): a random genome has almost no
    repeated 25-mers, so the vertices are the two chromosome-end k-mers plus rare repeats; size-independent
    invariants: strand symmetry, sortedness, id range."""
    g = synth.random_genome(16_000_000, 12345)
    count, pos, neg = ctx.enumerate([g], 25)
    assert len(pos) == len(neg)
    assert count >= 4
    key = pos["chr"].astype(np.int64) << 32 | pos["pos"]
    assert np.all(np.diff(key) > 0)
    keyn = neg["chr"].astype(np.int64) << 32 | neg["pos"]
    assert np.all(np.diff(keyn) > 0)
    assert pos["bifId"].max() < count and neg["bifId"].max() < count
    # every + instance at p has its - twin at len - p - k
    assert np.array_equal(np.sort(len(g) - pos["pos"].astype(np.int64) - 25), neg["pos"].astype(np.int64))
    want = restate.enumerate_bifurcations([g[:2_000_000]], 25)
    helpers.assert_tables_equal(ctx.enumerate([g[:2_000_000]], 25), want, "2 Mb prefix")
