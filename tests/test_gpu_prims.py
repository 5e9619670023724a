"""The scan / compaction / sort kernels in isolation, against numpy.
Index work: bit-exact.  Sizes cover empty, single, tile-boundary (tile = 2048
for the scans, 4096 for the sorts) and multi-million element inputs."""
import numpy as np
import pytest

from mygpuraytracer_b200 import api

pytestmark = pytest.mark.gpu
SIZES = [0, 1, 2, 31, 32, 33, 2047, 2048, 2049, 4095, 4096, 4097, 100_000, 3_000_001]


@pytest.mark.parametrize("n", SIZES)
def test_scan_exclusive(n):
    rng = np.random.default_rng(n)
    a = rng.integers(0, 50, n).astype(np.int32)
    want = np.concatenate([[0], np.cumsum(a, dtype=np.int64)[:-1]]).astype(np.int32) if n else a
    assert np.array_equal(api.scan_exclusive(a), want)


@pytest.mark.parametrize("n", SIZES)
def test_compact_nonzero(n):
    rng = np.random.default_rng(n + 1)
    a = (rng.integers(0, 3, n) * rng.integers(1, 1000, n)).astype(np.int32)
    assert np.array_equal(api.compact_nonzero(a), a[a != 0])


@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("p", [0.0, 0.3, 1.0])
def test_partition_perm_is_stable(n, p):
    """thrust::stable_partition(isTerminate), apps/src/pathtrace.cu:649."""
    rng = np.random.default_rng(n + 7)
    f = (rng.random(n) < p).astype(np.uint8)
    perm, kept = api.partition_perm(f)
    idx = np.arange(n, dtype=np.int32)
    assert kept == int(f.sum())
    assert np.array_equal(perm, np.concatenate([idx[f != 0], idx[f == 0]]))


@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("nmat", [1, 7, 256])
def test_sort_desc_perm_equals_stable_sort(n, nmat):
    """thrust::sort_by_key(sortByMaterial), apps/src/pathtrace.cu:512-516,612."""
    rng = np.random.default_rng(n + nmat)
    k = rng.integers(0, nmat, n).astype(np.int32)
    want = np.argsort(-k.astype(np.int64), kind="stable").astype(np.int32)
    assert np.array_equal(api.sort_desc_perm(k), want)


@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("nmat,general", [(1, False), (3, False), (6, False), (8, False), (6, True), (9, False), (200, False)])
def test_material_sort_with_compaction_ranks(n, nmat, general):
    """The kernels the renderer launches per depth (k_sort_material_few for <= 8 materials, k_sort_material
    otherwise): thrust::sort_by_key(sortByMaterial) + the live prefix of stable_partition(isTerminate)."""
    rng = np.random.default_rng(n * 31 + nmat)
    m = rng.integers(0, nmat, n).astype(np.uint8)
    if n > 100:
        m[n // 4: n // 2] = nmat - 1  # a long run of one material crossing tiles
    live = (rng.random(n) < 0.6).astype(np.uint8)
    perm, rank, kept = api.sort_material_ranks(m, live, nmat, general)
    want = np.argsort(-m.astype(np.int64), kind="stable").astype(np.int32)
    assert np.array_equal(perm, want)
    ls = live[want].astype(np.int64)
    assert np.array_equal(rank, (np.cumsum(ls) - ls).astype(np.int32))
    assert kept == int(live.sum())


@pytest.mark.parametrize("n", SIZES)
def test_radix_sort_pairs(n):
    rng = np.random.default_rng(n + 3)
    k = rng.integers(0, 1 << 30, n, dtype=np.int64).astype(np.uint32)
    if n > 10:
        k[: n // 3] = k[0]  # many duplicates: stability matters
    v = np.arange(n, dtype=np.uint32)
    ks, vs = api.radix_sort_pairs(k, v)
    order = np.argsort(k, kind="stable")
    assert np.array_equal(ks, k[order]) and np.array_equal(vs, v[order])


def test_sort_rejects_out_of_range_keys():
    with pytest.raises(api.B2ptError):
        api.sort_desc_perm(np.array([0, 300], np.int32))
