"""GPU parity of the covariance (pcavar) accumulation against the oracle's restatement of
src/SphericalBasis.cc:501-581: per jackknife bin the count, mass, coefficient share (complex)
and the nmax x nmax covariance matrices (computed on the FP64 tensor cores)."""
import numpy as np
import pytest

from conftest import get_table
from exp_b200 import ics
from exp_b200.component import Component
from exp_b200.sphere import Sphere
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("kind,L,N,n,sampT,flags", [
    ("hernquist", 4, 10, 40000, 1, {}),
    ("hernquist", 4, 10, 40000, 7, {}),
    ("nfw", 6, 18, 60000, 3, {}),
    ("nfw", 6, 18, 30000, 100, {}),
    ("hernquist", 2, 10, 5000, 4, dict(M0_only=True)),
])
def test_sphere_covariance_parity(kind, L, N, n, sampT, flags):
    tab = get_table(kind, L, N, rmax=50.0 if kind == "hernquist" else 20.0)
    gen = ics.hernquist_halo if kind == "hernquist" else ics.nfw_halo
    pos, vel, mass = gen(n, seed=23)
    pos = pos + np.array([0.01, -0.02, 0.005])
    pos[:5] *= 100.0
    g = O.Grid(tab)
    S = O.ParticleSet(pos, vel, mass)
    P = O.sph_params(rmin=tab.rmin, rmax=tab.rmax, **flags)
    ref = O.sph_coefficients_cov(g, P, S, sampT)

    c = Component(n)
    c.upload(pos, vel, mass)
    s = Sphere(tab, **flags)
    out = s.covariance(c, sampT)
    assert np.array_equal(out["countT"], ref["countT"])
    assert _rel(out["massT"], ref["massT"]) <= 1e-10      # serial vs tree summation of n masses
    assert _rel(out["coefT"], ref["coefT"]) <= 1e-10
    assert _rel(out["coefM"], ref["coefM"]) <= 1e-10
    # the bins add up to the coefficients themselves (a checksum of checksums)
    coef = s.determine_coefficients(c)
    tot = out["coefT"].sum(axis=0)
    for l in range(L + 1):
        for m in range(l + 1):
            iC, row = l * (l + 1) // 2 + m, l * l + (2 * m - 1 if m else 0)
            assert np.abs(tot[iC].real - coef[row]).max() <= 1e-12 * np.abs(coef).max()
            if m:
                assert np.abs(tot[iC].imag - coef[row + 1]).max() <= 1e-12 * np.abs(coef).max()
    # and the run is reproducible bit for bit
    again = s.covariance(c, sampT)
    assert again["coefM"].tobytes() == out["coefM"].tobytes()


def test_covariance_after_reordering_uses_host_slots():
    """Bins follow the host slot (Particle::indx), not the storage order."""
    tab = get_table("hernquist", 2, 10)
    n, sampT = 20000, 5
    pos, vel, mass = ics.hernquist_halo(n, seed=5)
    g = O.Grid(tab)
    S = O.ParticleSet(pos, vel, mass)
    ref = O.sph_coefficients_cov(g, O.sph_params(rmin=tab.rmin, rmax=tab.rmax), S, sampT)
    c = Component(n)
    c.upload(pos, vel, mass)
    s = Sphere(tab)
    s.determine_coefficients(c)            # sorts and re-orders the store
    out = s.covariance(c, sampT)
    assert np.array_equal(out["countT"], ref["countT"])
    assert _rel(out["coefM"], ref["coefM"]) <= 1e-10


# ---- cylinder (`covar` branch of EmpCylSL::accumulate, exputil/EmpCylSL.cc:4123-4141) ----
@pytest.mark.parametrize("shape,n,sampT,rot", [
    ((4, 6, 48, 24), 20000, 1, False),
    ((4, 6, 48, 24), 20000, 7, True),
    ((6, 18, 64, 32), 30000, 3, False),
    ((6, 18, 64, 32), 30000, 173, False),      # ~ sqrt(N) bins: most (bin, cell) groups are singletons
    ((0, 4, 16, 8), 3000, 2, False),
    ((9, 5, 24, 12), 6000, 4, False),          # two harmonic groups (mmax + 1 > 8)
])
def test_cylinder_covariance_parity(shape, n, sampT, rot):
    from exp_b200 import cyltable
    from exp_b200.cylinder import Cylinder
    from test_emul_cyl import disk, rotations
    tab = cyltable.synthetic_cyl_table(mmax=shape[0], norder=shape[1], numx=shape[2], numy=shape[3])
    pos, vel, mass = disk(n, seed=61)
    body, orig = rotations() if rot else (None, None)
    g = O.CylGridPy(tab)
    P = O.cyl_params(rcylmax=20.0, acyl=tab.ascale, body=body, orig=orig)
    S = O.ParticleSet(pos, vel, mass)
    ref = O.cyl_coefficients_cov(g, P, S, sampT)

    c = Component(n)
    c.upload(pos, vel, mass)
    cyl = Cylinder(tab)
    if rot:
        cyl.set_orientation(body, orig)
    cs, sn = cyl.determine_coefficients(c)          # sorts (and may re-order) the store first
    out = cyl.covariance(c, sampT)
    assert np.array_equal(out["numbT"], ref["numbT"])
    assert out["numbT"].sum() <= ref["used"]      # `used` also counts particles off the table
    assert _rel(out["massT"], ref["massT"]) <= 1e-10
    assert _rel(out["VC"], ref["VC"]) <= 1e-10
    assert _rel(out["MV"], ref["MV"]) <= 1e-10
    # Hermitian, and the bins add up to the coefficients: Re vec = vC cos + vS sin, so
    # sum_T Re VC = cosN + sinN (a checksum of checksums)
    assert np.abs(out["MV"] - np.conj(np.swapaxes(out["MV"], -1, -2))).max() <= 1e-14 * np.abs(out["MV"]).max()
    tot = out["VC"].sum(axis=0)
    sn0 = sn.copy()
    sn0[0] = 0.0
    assert np.abs(tot.real - (cs + sn0)).max() <= 1e-11 * np.abs(cs).max()
    again = cyl.covariance(c, sampT)
    assert again["MV"].tobytes() == out["MV"].tobytes()      # reproducible bit for bit
