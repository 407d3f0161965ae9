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
    l.host_prior_floor.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
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


# ---- the prior's eigenvalue floor (csrc/prior_floor.h): partial route (bisection + inverse iteration for the eigenpairs at
# the noise floor only) and full decomposition against the definition evaluated with numpy

def floor(lib, A, b, eps=1e-8, full=False):
    A = np.array(A, dtype=np.float64, order="C")
    g = np.array(b, dtype=np.float64)
    c0, st = np.zeros(1), np.zeros(2, np.int32)
    lib.host_prior_floor(A.ctypes.data, g.ctypes.data, A.shape[0], eps, int(full), c0.ctypes.data, st.ctypes.data)
    return A, g, c0[0], st


def floor_ref(A, b, eps=1e-8):
    """marginalization_factor.cpp:268-297 in information form: A+ = V S+ V^T, g0 = V 1+ V^T b, c0 = b^T V S+^-1 V^T b."""
    w, V = np.linalg.eigh((A + A.T) / 2)
    keep = w > eps
    vb = V.T @ b
    return (V[:, keep] * w[keep]) @ V[:, keep].T, V[:, keep] @ vb[keep], float(np.sum(vb[keep] ** 2 / w[keep])), w


def test_prior_floor_on_a_real_prior(lib):
    A = np.fromfile(os.path.join(ROOT, "harness", "micro", "prior75.bin"), dtype=np.float64)[:5625].reshape(75, 75)
    w, V = np.linalg.eigh((A + A.T) / 2)
    rng = np.random.default_rng(0)
    b = V[:, 3:] @ (rng.standard_normal(72) * np.sqrt(w[3:])) * 3  # in the range of the kept part, as b' = J^T r is
    rA, rg, rc, _ = floor_ref(A, b)
    for full in (True, False):
        Ap, g0, c0, st = floor(lib, A, b, full=full)
        assert (st[0] == -1) if full else (st[0] == 3), st  # three eigenvalues at the 1e-7 noise floor, next one 2e-2
        assert st[1] == 1                                   # one of them below eps
        assert np.abs(Ap - rA).max() < 1e-13 * np.abs(A).max()
        assert np.abs(g0 - rg).max() < 1e-11 * np.abs(b).max()
        assert abs(c0 - rc) < 1e-8 * rc


@pytest.mark.parametrize("seed", range(24))
def test_prior_floor_synthetic(lib, seed):
    """Rank-deficient J^T J with 6..10 decades of dynamic range; the partial route must agree with the definition wherever
    it runs, and must hand over to the full decomposition when too many eigenvalues sit below tau."""
    rng = np.random.default_rng(seed)
    n = int(rng.choice([75, 76, 69, 30, 12, 136]))
    null = int(rng.integers(0, 8))
    rows = n - null
    J = rng.standard_normal((rows, n)) * np.logspace(rng.uniform(-3, 0), rng.uniform(2, 4.5), rows)[:, None]
    A = J.T @ J
    A = A + rng.standard_normal((n, n)) * 1e-17 * np.abs(A).max()  # assembled, not exactly symmetric
    b = J.T @ rng.standard_normal(rows)
    rA, rg, rc, w = floor_ref(A, b)
    scale = np.abs(A).max()
    ambiguous = np.any(np.abs(w - 1e-8) < 1e-13 * scale * n)  # an eigenvalue within rounding of eps: either side is right
    for full in (False, True):
        Ap, g0, c0, st = floor(lib, A, b, full=full)
        assert np.abs(Ap - rA).max() < 1e-13 * scale
        if not ambiguous:
            assert np.abs(g0 - rg).max() < 1e-9 * np.abs(b).max()
            assert abs(c0 - rc) < 1e-5 * rc
    _, _, _, st = floor(lib, A, b)
    tau = max(1e-12 * np.abs(A).sum(1).max(), 4e-8)
    if (w < 0.3 * tau).sum() > 16:
        assert st[0] == -1
    if (w < 3 * tau).sum() <= 16:
        assert st[0] >= 0


def test_prior_floor_full_rank_and_tiny(lib):
    rng = np.random.default_rng(5)
    B = rng.standard_normal((40, 40))
    A = B @ B.T + 40 * np.eye(40)
    b = rng.standard_normal(40)
    Ap, g0, c0, st = floor(lib, A, b)
    assert st[0] == 0 and st[1] == 0
    assert np.abs(Ap - (A + A.T) / 2).max() < 1e-13 * np.abs(A).max() and np.abs(g0 - b).max() < 1e-14
    assert abs(c0 - b @ np.linalg.solve(A, b)) < 1e-12 * c0
    Ap, g0, c0, st = floor(lib, np.array([[2.0]]), np.array([3.0]))
    assert st[0] == -1 and Ap[0, 0] == 2.0 and g0[0] == 3.0 and abs(c0 - 4.5) < 1e-15
    Ap, g0, c0, st = floor(lib, np.array([[1e-9]]), np.array([3.0]))
    assert Ap[0, 0] == 0.0 and g0[0] == 0.0 and c0 == 0.0
