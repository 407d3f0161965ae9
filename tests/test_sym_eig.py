"""Host-side check of the product's tridiagonal-QL eigen-solver (csrc/sym_eig.h) against numpy.

The same template runs inside marg_solve_kernel with a CTA context; here it is instantiated single-threaded."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    out = tmp_path_factory.mktemp("symeig") / "libsymeig.so"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off",
                           os.path.join(ROOT, "tests", "hostlib", "sym_eig_host.cpp"), "-o", str(out)])
    l = ctypes.CDLL(str(out))
    l.host_sym_eig.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    return l


def run(lib, A):
    n = A.shape[0]
    A = np.ascontiguousarray(A, dtype=np.float64)
    w = np.zeros(n)
    V = np.zeros((n, n))
    lib.host_sym_eig(A.ctypes.data, n, w.ctypes.data, V.ctypes.data)
    return w, V


@pytest.mark.parametrize("n", [1, 2, 3, 6, 15, 75, 76, 136])
def test_random_symmetric(lib, n):
    rng = np.random.default_rng(n)
    B = rng.standard_normal((n, n))
    A = B + B.T
    w, V = run(lib, A)
    scale = np.abs(A).max()
    assert np.abs(V.T @ V - np.eye(n)).max() < 1e-13
    assert np.abs(A @ V - V * w).max() < 1e-13 * scale * n
    assert np.abs(np.sort(w) - np.linalg.eigvalsh(A)).max() < 1e-13 * scale * n


def test_rank_deficient_prior_like(lib):
    """PSD, huge dynamic range, 4-dimensional null space: the shape of the marginalisation prior."""
    rng = np.random.default_rng(7)
    n = 75
    J = rng.standard_normal((n - 4, n)) * np.logspace(-2, 4, n - 4)[:, None]
    A = J.T @ J
    w, V = run(lib, A)
    ref = np.linalg.eigvalsh(A)
    scale = ref.max()
    assert np.abs(np.sort(w) - ref).max() < 1e-13 * scale
    assert np.abs(A @ V - V * w).max() < 1e-12 * scale
    assert (np.sort(w)[:4] < 1e-9 * scale).all()


def test_diagonal_and_zero(lib):
    w, V = run(lib, np.diag([3.0, -1.0, 2.0, 0.0]))
    assert np.allclose(np.sort(w), [-1, 0, 2, 3])
    assert np.abs(np.abs(V).sum(0) - 1).max() < 1e-15
    w, V = run(lib, np.zeros((5, 5)))
    assert np.all(w == 0) and np.abs(V.T @ V - np.eye(5)).max() < 1e-15
