"""Field solve at the benchmark's own grid sizes (VERDICT r1 "parity on the benchmarked configuration").

The reference's dense matrix cannot exist at 1025^2 / 2049^2 nodes (8.8 / 141 TB, generalized_poisson.jl:37), so the
check is matrix free: the device phi is put through the reference's discrete operator -- 5-point stencil over the
neighbours that exist, divided by dh[1]^2 (:44-65), "periodic" coupling first <-> last node (:286-324), identity rows on
Dirichlet nodes (:205-215) -- and the residual against b = -rho/eps0 (:375) must vanish to 1e-10 of |b|.  E is
checked against the stencil of calculate_electric_field! (:398-410) evaluated on the device phi.
"""
import numpy as np
import pytest

from oracle import pic_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ib():
    import iskra_b200
    return iskra_b200


def _apply_reference_operator(phi, dx, periodic):
    """A*phi for every node, rows of the assembled operator before apply_dirichlet."""
    out = np.zeros_like(phi)
    out[:-1, :] += phi[1:, :] - phi[:-1, :]        # neighbour i+1 exists
    out[1:, :] += phi[:-1, :] - phi[1:, :]         # neighbour i-1
    out[:, :-1] += phi[:, 1:] - phi[:, :-1]        # j+1
    out[:, 1:] += phi[:, :-1] - phi[:, 1:]         # j-1
    if 1 in periodic:                              # apply_periodic(ps, 1): couples j = 1 <-> j = ny
        out[:, 0] += phi[:, -1] - phi[:, 0]
        out[:, -1] += phi[:, 0] - phi[:, -1]
    if 2 in periodic:                              # apply_periodic(ps, 2): couples i = 1 <-> i = nx
        out[0, :] += phi[-1, :] - phi[0, :]
        out[-1, :] += phi[0, :] - phi[-1, :]
    return out / dx ** 2


def _reference_efield(phi, dx):
    ex = np.empty_like(phi)
    ey = np.empty_like(phi)
    ex[1:-1, :] = (phi[:-2, :] - phi[2:, :]) / (2.0 * dx)
    ex[0, :] = (phi[0, :] - phi[1, :]) / dx
    ex[-1, :] = (phi[-2, :] - phi[-1, :]) / dx
    ey[:, 1:-1] = (phi[:, :-2] - phi[:, 2:]) / (2.0 * dx)
    ey[:, 0] = (phi[:, 0] - phi[:, 1]) / dx
    ey[:, -1] = (phi[:, -2] - phi[:, -1]) / dx
    return ex, ey


@pytest.mark.parametrize("n,periodic,edges,path", [
    (2049, (1,), (("l", 450.0), ("r", 0.0)), "fft"),        # C5: 11_rf_discharge.jl:76-78 at 2049^2, DST by FFT, cyclic Thomas
    (1025, (1,), (("l", -3.0), ("r", 7.0)), "fft"),
    (1025, (), (("l", 1.0), ("r", 2.0)), "fft"),            # open in j: plain Thomas
    (1025, (1, 2), (), "gemm"),                             # C4: 10_two_streams.jl:53-54 at 1025^2, singular operator
    (513, (2,), (("b", 5.0),), "gemm"),                     # one Dirichlet edge, ring in i
])
def test_field_solve_residual_at_scale(ib, n, periodic, edges, path):
    FDM = ib.finite_difference_method
    dx = 5.234375e-4
    g = ib.regular_grids.create_uniform_grid(np.arange(n) * dx, np.arange(n) * dx)
    ps = FDM.create_poisson_solver(g, O.eps0)
    for ax in periodic:
        FDM.apply_periodic(ps, ax)
    isdir = np.zeros((n, n), bool)
    dval = np.zeros((n, n))
    for name, val in edges:
        m = np.zeros((n, n), bool)
        if name == "l":
            m[0, :] = True
        elif name == "r":
            m[n - 1, :] = True
        elif name == "b":
            m[:, 0] = True
        else:
            m[:, n - 1] = True
        FDM.apply_dirichlet(ps, m, val)
        isdir |= m
        dval[m] = val
    rng = np.random.default_rng(n)
    rho = rng.standard_normal((n, n)) * 1e-6
    singular = not edges
    if singular:
        rho -= rho.mean()            # H3: the fully periodic operator only has solutions for a mean-free right-hand side
    rt = g._rt
    rt.set_fields(rho=rho)
    ib._lib.check(rt.lib.iskb_field_solve(rt.h))
    _, phi, E = rt.fields(rho=False)
    b = -rho / O.eps0
    r = _apply_reference_operator(phi, dx, periodic) - b
    r[isdir] = phi[isdir] - dval[isdir]
    if singular:
        r -= r.mean()
    scale = np.abs(b).max()
    assert np.abs(r[~isdir]).max() <= 1e-10 * scale
    if edges:
        assert np.abs(r[isdir]).max() == 0.0
    ex, ey = _reference_efield(phi, dx)
    assert np.abs(E[:, :, 0] - ex).max() <= 1e-13 * np.abs(ex).max()
    assert np.abs(E[:, :, 1] - ey).max() <= 1e-13 * np.abs(ey).max()
    assert np.all(E[:, :, 2] == 0.0)


def test_device_dense_inversion_is_bit_identical_to_the_host_one(ib):
    """The dense Poisson path inverts its operator on the device from 512 unknowns on (csrc/poisson.cu
    invert_dense_device): same pivots, same arithmetic per element as the host Gauss-Jordan, so the two inverses agree
    bit for bit -- on a random matrix, on a row-equilibrated axial operator with pivoting, and both flag a singular one."""
    import ctypes as C
    import time
    from iskra_b200 import _lib as L
    from oracle import axial_oracle as AX
    g = ib.regular_grids.create_uniform_grid(np.arange(33) * 1e-3, np.arange(33) * 1e-3)      # a context to run on
    host, dev = L.lib().iskb_debug_invert_dense, L.lib().iskb_debug_invert_dense_device
    host.argtypes, host.restype = [C.POINTER(C.c_double), C.c_int64], C.c_int32
    dev.argtypes, dev.restype = [C.c_void_p, C.POINTER(C.c_double), C.c_int64], C.c_int32
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    rng = np.random.default_rng(1)
    n = 700
    M = rng.standard_normal((n, n))
    M[rng.random((n, n)) < 0.5] = 0.0                       # zeros exercise the skipped updates
    M = np.asfortranarray(M + 3.0 * np.diag(rng.standard_normal(n)))
    A, B = M.copy(order="F"), M.copy(order="F")
    assert host(dp(A), n) == 0 and dev(g._rt.h, dp(B), n) == 0
    assert np.array_equal(A, B)
    assert np.abs(B @ M - np.eye(n)).max() < 1e-8
    og = AX.AxialGrid2(np.arange(25) * 0.0025, np.arange(49) * 0.0025)                       # 1225 unknowns
    ps = AX.PoissonSolver(og, 1.0)
    bot, top = np.zeros((25, 49), bool), np.zeros((25, 49), bool)
    bot[:, 0], top[:, 48] = True, True
    O.apply_dirichlet(ps, bot, 0.0)
    O.apply_dirichlet(ps, top, 1.0)
    As = np.asfortranarray(ps.A / np.abs(ps.A).max(axis=1)[:, None])
    A, B = As.copy(order="F"), As.copy(order="F")
    t0 = time.time()
    assert host(dp(A), As.shape[0]) == 0
    t1 = time.time()
    assert dev(g._rt.h, dp(B), As.shape[0]) == 0
    t2 = time.time()
    assert np.array_equal(A, B) and np.abs(B @ As - np.eye(As.shape[0])).max() < 1e-12
    print("dense inverse of %d unknowns: host %.2f s, device %.2f s" % (As.shape[0], t1 - t0, t2 - t1))
    S = np.zeros((600, 600), order="F")
    S[:599, :599] = np.eye(599)
    assert dev(g._rt.h, dp(S), 600) == L.E_SINGULAR
