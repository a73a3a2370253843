"""CPU tests of the drop-in boundary: the C-ABI library loads and exports exactly what include/fctpdeco.h
declares; host-side logic (mesh bookkeeping, pattern embedding, reorder helpers) matches the oracle; the
product refuses to run without a GPU instead of falling back."""
import ctypes
import os
import re

import numpy as np
import pytest
import scipy.sparse as sp

import fem_fct_pdeco_b200 as fp
from fem_fct_pdeco_b200 import _lib, helpers
from fem_fct_pdeco_b200.mesh import RectMeshP1
from oracle.p1mesh import RectMesh, reorder_vector_from_dof, reorder_vector_to_dof

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "fctpdeco.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fct_[a-zA-Z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    names = _header_symbols()
    assert len(names) >= 40
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/fctpdeco.h but not exported"


def test_python_binding_covers_header():
    assert sorted(_lib.SIGNATURES) == _header_symbols()


def test_no_cpu_fallback_without_gpu():
    if fp.device_count() > 0:
        pytest.skip("a GPU is visible")
    m = RectMeshP1(4)
    with pytest.raises(fp.FctError):
        fp.FctContext(m.rowptr, m.colidx)
    M = sp.identity(5, format="csr")
    with pytest.raises(fp.FctError):
        helpers.ChebSI(np.ones(5), M, np.ones(5))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "fem-fct-pdeco_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), fn


@pytest.mark.parametrize("n", [1, 2, 3, 7, 40, 80])
def test_mesh_builder_matches_oracle_bit_exact(n):
    m = RectMeshP1(n, -1.0, 1.0)
    o = RectMesh(n, -1.0, 1.0)
    rp, ci = o.pattern()
    assert np.array_equal(m.vertex_to_dof, o.vertex_to_dof)
    assert np.array_equal(m.cells, o.cells)
    assert np.array_equal(m.rowptr, rp) and np.array_equal(m.colidx, ci)
    assert np.array_equal(m.dof_xy, o.dof_xy)
    assert m.dof_neighbors() == o.dof_neighbors()


def test_mesh_sizes_4096():
    nodes, cells, nnz = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
    assert _lib.lib.fct_mesh_rect_sizes(4096, ctypes.byref(nodes), ctypes.byref(cells), ctypes.byref(nnz)) == 0
    assert (nodes.value, cells.value, nnz.value) == (16785409, 33554432, 117465089)
    assert _lib.lib.fct_mesh_rect_sizes(0, None, None, None) != 0
    assert b"n must be" in _lib.lib.fct_last_error()


def test_reorder_helpers_match_reference_semantics():
    m = RectMeshP1(5)
    v = np.random.default_rng(1).random(3 * m.nodes)
    d = helpers.reorder_vector_to_dof(v, 3, m.nodes, m.vertex_to_dof)
    assert np.array_equal(d, reorder_vector_to_dof(v, 3, m.nodes, m.vertex_to_dof))
    assert np.array_equal(helpers.reorder_vector_from_dof(d, 3, m.nodes, m.vertex_to_dof), v)
    assert np.array_equal(helpers.reorder_vector_from_dof_time(d, 3, m.nodes, m.vertex_to_dof),
                          reorder_vector_from_dof(d, 3, m.nodes, m.vertex_to_dof))
    b, bd = helpers.generate_boundary_nodes(m.nodes, m.vertex_to_dof)
    assert len(b) == 4 * 5 and sorted(bd) == sorted(int(m.vertex_to_dof[i]) for i in b)


def test_cost_functional_rejects_bad_optim_before_touching_the_gpu():
    with pytest.raises(ValueError, match="Invalid value for 'optim'"):
        helpers.cost_functional(np.zeros(4), np.zeros(4), np.zeros(4), 0, 0.1, sp.identity(4), 0.1, "sometime")


def test_pattern_embedding_reinserts_pruned_zeros():
    from fem_fct_pdeco_b200.pattern import HostPattern
    m = RectMeshP1(6)
    pat = HostPattern(m.rowptr, m.colidx)
    rng = np.random.default_rng(3)
    vals = rng.standard_normal(pat.nnz)
    vals[rng.random(pat.nnz) < 0.3] = 0.0
    A = pat.to_scipy(vals)
    assert np.array_equal(pat.embed(A), vals)                  # identical pattern: direct
    A.eliminate_zeros()
    assert A.nnz < pat.nnz
    assert np.array_equal(pat.embed(A), vals)                  # pruned CSR
    assert np.array_equal(pat.embed(sp.lil_matrix(A)), vals)   # lil, as the reference passes M
    assert np.array_equal(pat.embed(A.toarray()), vals)        # dense
    bad = sp.lil_matrix(A)
    bad[0, m.nodes - 1] = 1.0
    with pytest.raises(ValueError, match="outside the fixed P1 pattern"):
        pat.embed(bad)


def test_trajectory_io_csv_and_npy(tmp_path):
    """import_data_final / extract_data (helpers.py:1874-1956): the reference's comma-separated text and the binary
    .npy path give the same arrays"""
    m = RectMeshP1(3)
    nodes = m.nodes
    x = np.random.default_rng(2).random(3 * nodes)
    helpers.export_trajectory(str(tmp_path / "a.csv"), x)          # the reference's np.tofile(sep=",") text
    helpers.export_trajectory(str(tmp_path / "a.npy"), x)          # binary, same flat DoF-ordered layout
    assert np.array_equal(np.genfromtxt(tmp_path / "a.csv", delimiter=","), x)
    r1, d1 = helpers.import_data_final(str(tmp_path / "a.csv"), nodes, m.vertex_to_dof, num_steps=2)
    r2, d2 = helpers.import_data_final(str(tmp_path / "a.npy"), nodes, m.vertex_to_dof, num_steps=2)
    assert np.array_equal(r1, r2) and np.array_equal(d1, x[2 * nodes:]) and np.array_equal(d2, d1)
    assert r1.shape == (4, 4) and np.array_equal(r1.ravel(), d1[m.vertex_to_dof])
    r3, d3 = helpers.import_data_final(str(tmp_path / "a.npy"), nodes, m.vertex_to_dof, num_steps=2, time_dep=True)
    assert np.array_equal(d3, x) and np.array_equal(r3, helpers.reorder_vector_from_dof(x, 3, nodes, m.vertex_to_dof))
    helpers.extract_data(str(tmp_path), "a", 0.2, 0.1, nodes, m.vertex_to_dof)
    assert np.array_equal(np.load(tmp_path / "a_T0.2.npy"), x[2 * nodes:])
