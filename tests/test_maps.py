"""Parent <-> sub-mesh maps of quadrature arrays (fenics_constitutive_b200/solver/maps.py), restating
reference tests/solver/test_maps.py on CPU tensors (the maps are torch index operations, the device does not
matter): scalar / vector / tensor values share one cell map, sub -> parent round trips, identity map."""
import numpy as np
import pytest
import torch

from fenics_constitutive_b200.solver.maps import IdentityMap, SubSpaceMap, build_subspace_map

NUM_CELLS = 5 * 7 * 11 * 6  # create_unit_cube(5, 7, 11) has 6 tets per cube (reference :31, :129)
WIDTHS = [1, 3, 9]          # value_shape (1,), (3,), (3, 3) of the reference's quadrature elements (:15-25)


def test_subspace_vector_map_vector_equals_tensor_map():
    """reference test_maps.py:29-74: the cell map does not depend on the value shape."""
    rng = np.random.default_rng(42)
    sample = rng.choice(np.arange(NUM_CELLS), NUM_CELLS // 2, replace=False)
    maps = [build_subspace_map(sample, NUM_CELLS, "cpu") for _ in WIDTHS]
    assert all(isinstance(m, SubSpaceMap) for m in maps)
    assert all(np.array_equal(m.cell_map, maps[0].cell_map) for m in maps)


@pytest.mark.parametrize("width", WIDTHS)
@pytest.mark.parametrize("nq", [1, 4])
def test_subspace_map_evaluation(width, nq):
    """reference test_maps.py:77-122: map_to_sub then map_to_parent reproduces the sampled cells' rows
    (and leaves the others alone), ten random cell samples."""
    rng = np.random.default_rng(42)
    values = torch.from_numpy(rng.random(NUM_CELLS * nq * width))
    for _ in range(10):
        sample = rng.choice(np.arange(NUM_CELLS), NUM_CELLS // 2, replace=False)
        m = build_subspace_map(sample, NUM_CELLS, "cpu")
        sub = torch.full((sample.size * nq * width,), float("nan"), dtype=torch.float64)
        back = torch.full_like(values, -1.0)
        m.map_to_sub(values, sub)
        assert torch.equal(sub.view(sample.size, -1), values.view(NUM_CELLS, -1)[torch.from_numpy(sample)])
        m.map_to_parent(sub, back)
        v, b = values.view(NUM_CELLS, -1).numpy(), back.view(NUM_CELLS, -1).numpy()
        assert np.all(v[sample] == b[sample])
        rest = np.setdiff1d(np.arange(NUM_CELLS), sample)
        assert np.all(b[rest] == -1.0)


@pytest.mark.parametrize("width", WIDTHS)
def test_identity_map_evaluation(width):
    """reference test_maps.py:125-156: a law on every cell gets the IdentityMap."""
    cells = np.arange(NUM_CELLS, dtype=np.int32)
    m = build_subspace_map(cells, NUM_CELLS, "cpu")
    assert isinstance(m, IdentityMap)
    rng = np.random.default_rng(42)
    values = torch.from_numpy(rng.random(NUM_CELLS * width))
    sub, back = torch.zeros_like(values), torch.zeros_like(values)
    m.map_to_sub(values, sub)
    m.map_to_parent(sub, back)
    assert torch.equal(values, back)


def test_empty_cell_list_is_a_no_op():
    m = SubSpaceMap(np.zeros(0, dtype=np.int64), NUM_CELLS, "cpu")
    parent = torch.ones(NUM_CELLS * 3, dtype=torch.float64)
    m.map_to_sub(parent, torch.zeros(0, dtype=torch.float64))
    m.map_to_parent(torch.zeros(0, dtype=torch.float64), parent)
    assert torch.all(parent == 1.0)
