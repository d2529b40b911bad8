"""Parent <-> sub-mesh maps of quadrature arrays (fenics_constitutive_b200/solver/maps.py), restating
reference tests/solver/test_maps.py: scalar / vector / tensor values share one cell map, sub -> parent round
trips (the row gather / scatter kernels of csrc/fcx_maps.cu: GPU test, compared with numpy indexing),
identity map."""
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


@pytest.mark.gpu
@pytest.mark.parametrize("width", WIDTHS + [6, 36])  # + Mandel stress / tangent rows (s = 6)
@pytest.mark.parametrize("nq", [1, 4])
def test_subspace_map_evaluation(width, nq):
    """reference test_maps.py:77-122: map_to_sub then map_to_parent reproduces the sampled cells' rows
    (and leaves the others alone), ten random cell samples -- through fcx_map_rows_to_sub /
    fcx_map_rows_to_parent, bit for bit against numpy indexing."""
    rng = np.random.default_rng(42)
    host = rng.random(NUM_CELLS * nq * width)
    values = torch.from_numpy(host).cuda()
    for _ in range(10):
        sample = rng.choice(np.arange(NUM_CELLS), NUM_CELLS // 2, replace=False)
        m = build_subspace_map(sample, NUM_CELLS, "cuda")
        sub = torch.full((sample.size * nq * width,), float("nan"), dtype=torch.float64, device="cuda")
        back = torch.full_like(values, -1.0)
        m.map_to_sub(values, sub)
        assert np.array_equal(sub.cpu().numpy().reshape(sample.size, -1), host.reshape(NUM_CELLS, -1)[sample])
        m.map_to_parent(sub, back)
        v, b = host.reshape(NUM_CELLS, -1), back.cpu().numpy().reshape(NUM_CELLS, -1)
        assert np.all(v[sample] == b[sample])
        rest = np.setdiff1d(np.arange(NUM_CELLS), sample)
        assert np.all(b[rest] == -1.0)


@pytest.mark.gpu
def test_subspace_map_unaligned_views_and_odd_rows():
    """Rows with an odd number of doubles and views that are not 16-byte aligned take the 8-byte path."""
    rng = np.random.default_rng(7)
    nc, row = 1001, 9
    host = rng.random(nc * row + 1)
    dev = torch.from_numpy(host).cuda()
    sample = rng.choice(np.arange(nc), 333, replace=False)
    m = SubSpaceMap(sample, nc, "cuda")
    sub = torch.zeros(sample.size * row + 1, dtype=torch.float64, device="cuda")
    m.map_to_sub(dev[1:], sub[1:])  # both views start 8 bytes off a 16-byte boundary
    assert np.array_equal(sub.cpu().numpy()[1:].reshape(-1, row), host[1:].reshape(nc, row)[sample])
    back = torch.zeros(nc * row, dtype=torch.float64, device="cuda")
    m.map_to_parent(sub[1:], back)
    assert np.array_equal(back.cpu().numpy().reshape(nc, row)[sample], host[1:].reshape(nc, row)[sample])


def test_subspace_map_refuses_host_tensors():
    """No CPU fallback: the maps are CUDA kernels."""
    m = SubSpaceMap(np.array([0, 2]), 4, "cpu")
    with pytest.raises(ValueError):
        m.map_to_sub(torch.zeros(8, dtype=torch.float64), torch.zeros(4, dtype=torch.float64))


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
