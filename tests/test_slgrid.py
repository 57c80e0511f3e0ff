"""The SL table generator (host-side table provenance) against the reference's own
self-checks: orthoCheck identity (exputil/SLGridMP2.cc:1775-1836) and the known
Hernquist-Ostriker spectrum."""
import numpy as np

from conftest import get_table
from exp_b200 import slgrid
from oracle import oracle as O


def test_orthogonality(table_A):
    for l in range(table_A.lmax + 1):
        D = slgrid.ortho_check(table_A, l) - np.eye(table_A.nmax)
        assert np.abs(D).max() < (2e-3 if l == 0 else 1e-5)


def test_hernquist_spectrum(table_A):
    """eigenvalues -> K_nl = n(n+4l+3)/2 + (l+1)(2l+1) (Hernquist & Ostriker 1992)"""
    for l in range(1, 5):
        for n in range(6):
            K = n * (n + 4 * l + 3) / 2 + (l + 1) * (2 * l + 1)
            assert abs(table_A.ev[l, n] / K - 1) < 1e-2   # finite [rmin,rmax] shifts the spectrum slightly


def test_lowest_function_is_the_model_potential(table_A):
    """l=0,n=0: u ~ const, i.e. the lowest basis potential is proportional to Phi_0"""
    u = table_A.ef[0, 0, 50:-50]
    assert np.std(u) / abs(np.mean(u)) < 0.02
    assert abs(table_A.ev[0, 0] - 1.0) < 5e-3


def test_oracle_interpolation_reproduces_nodes(table_A):
    g = O.Grid(table_A)
    for i in (1, 10, 500, 1500, 1998):
        r = table_A.r[i]
        m = g.get_pot(r)
        exact = table_A.ef[:, :, i] / np.sqrt(table_A.ev) * table_A.p0[i]
        assert np.allclose(m, exact, rtol=1e-9, atol=1e-13)
    # force: 3-point derivative of a smooth function ~ analytic derivative of the interpolant
    r = 0.5 * (table_A.r[700] + table_A.r[701])
    h = 1e-6 * r
    fd = (g.get_pot(r + h) - g.get_pot(r - h)) / (2 * h)
    assert np.allclose(g.get_force(r), fd, rtol=2e-3, atol=1e-8)
