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
