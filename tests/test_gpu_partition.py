"""GPU: merge-path coordinates from the device search kernel must be BIT-EXACT with the
reference's CPU MergePathSearch (merge_based.hpp:22-44) -- checked against the golden fixture
(compiled reference), the SURVEY known answers, and the oracle on seeded random structures."""
import numpy as np
import pytest

from conftest import named_matrix

pytestmark = pytest.mark.gpu


def test_golden_coordinates(gpu, golden):
    cache = {}
    for rec in golden["partition"]:
        ro = cache.setdefault(rec["matrix"], named_matrix(gpu, rec["matrix"]))[0]
        got = gpu.merge_path_partition(ro, rec["threads"])
        assert got.tolist() == rec["coords"], (rec["matrix"], rec["threads"])


def test_survey_known_answers(gpu):
    ro, _, _ = gpu.gen_wheel(1000000)
    want = [[0, 0], [0, 375001], [0, 750002], [62502, 1062501], [250002, 1250002], [437503, 1437502],
            [625003, 1625003], [812504, 1812503], [1000001, 2000000]]
    assert gpu.merge_path_partition(ro, 8).tolist() == want
    ro, _, _ = gpu.gen_grid3d(150, True)
    want = [[0, 0], [423978, 2934147], [845158, 5871092], [1266338, 8808037], [1687500, 11745000],
            [2108661, 14681964], [2529841, 17618909], [2951021, 20555854], [3375000, 23490000]]
    assert gpu.merge_path_partition(ro, 8).tolist() == want
    assert gpu.merge_path_partition(ro, 2).tolist() == [[0, 0], [1687500, 11745000], [3375000, 23490000]]


@pytest.mark.parametrize("seed", range(6))
def test_random_structures_against_oracle(gpu, orc, seed):
    rng = np.random.default_rng(seed)
    m = int(rng.integers(1, 5000))
    deg = rng.integers(0, 12, size=m)
    deg[rng.random(m) < 0.25] = 0                 # empty rows
    if seed % 2:
        deg[int(rng.integers(0, m))] = 20000      # one very long row
    ro = np.concatenate([[0], np.cumsum(deg)]).astype(np.int32)
    for parts in (1, 2, 7, 64, 1000):
        assert np.array_equal(gpu.merge_path_partition(ro, parts), orc.merge_partition(ro, parts))
    # fixed-size tiles, as the kernels use them
    assert np.array_equal(gpu.merge_path_partition(ro, 50, 128), orc.merge_partition(ro, 50, 128))


def test_degenerate_shapes(gpu, orc):
    for ro in ([0], [0, 0], [0, 0, 0, 0], [0, 5], [0, 0, 3, 3]):
        ro = np.array(ro, dtype=np.int32)
        if len(ro) == 1:
            continue  # m = 0 handled below
        for parts in (1, 3):
            assert np.array_equal(gpu.merge_path_partition(ro, parts), orc.merge_partition(ro, parts))


def test_kernel_tile_coordinates_are_the_reference_search(gpu, orc):
    """the coordinates cached in the CSR handle (what the SpMV/SpMM kernel really consumes)"""
    for gen in (lambda: gpu.gen_grid3d(24, True), lambda: gpu.gen_wheel(50000), lambda: gpu.gen_rmat(12, 16)):
        ro, ci, va = gen()
        a = gpu.CsrMatrix(ro, ci, va)
        for k in (1, 8, 32):      # SpMV tiling, and the tiling of whichever SpMM kernel takes (matrix, k)
            coords, items = a.tile_coords(k)
            assert np.array_equal(coords, orc.merge_partition(ro, len(coords) - 1, items)), k
        a.close()
