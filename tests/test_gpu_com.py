"""Centre of mass / centring pass: expb200_component_fix_positions
against the oracle's restatement of Component::fix_positions_cpu (src/Component.cc:3280-3590)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _setup(n, seed, multistep, rtrunc=1e20, center=(0.0, 0.0, 0.0)):
    from exp_b200 import ics
    from exp_b200.component import Component
    from oracle import oracle as O
    rng = np.random.default_rng(seed)
    pos, vel, mass = ics.hernquist_halo(n, seed=seed)
    pos += np.array([0.3, -0.2, 0.1])
    mass = mass * rng.uniform(0.5, 1.5, n)
    acc = rng.normal(size=(n, 3))
    pot = rng.normal(size=n)
    level = rng.integers(0, multistep + 1, n).astype(np.uint32)
    c = Component(n)
    c.upload(pos, vel, mass, level=level)
    c.set_acc_pot(acc, pot)
    c.set_center(center, rtrunc)
    S = O.ParticleSet(pos, vel, mass, center=center, rtrunc=rtrunc, level=level)
    S.acc[:] = acc
    return c, S


@pytest.mark.parametrize("multistep,mlevel,dim", [(0, 0, 3), (4, 0, 3), (4, 2, 3), (3, 1, 2)])
def test_fix_positions_matches_oracle(multistep, mlevel, dim):
    from oracle import oracle as O
    n = 200_003
    c, S = _setup(n, 11 + multistep, multistep, rtrunc=8.0, center=(0.3, -0.2, 0.1))
    # rows below mlevel carry over from an earlier call in both implementations
    prev = np.arange((multistep + 1) * 10, dtype=np.float64).reshape(multistep + 1, 10) * 1e-3
    lev_o, out_o = O.fix_positions(S, mlevel, multistep, lev_sums=prev.copy(), dim=dim)
    lev_g, out_g = c.fix_positions(mlevel, multistep, lev_sums=prev.copy(), dim=dim)
    assert np.array_equal(lev_g[:mlevel], prev[:mlevel])
    # summation order only: errors are relative to the sum of the magnitudes of the terms
    m = S.mass[:, None]
    scale = np.concatenate([[S.mass.sum()], (m * np.abs(S.pos)).sum(0), (m * np.abs(S.vel)).sum(0),
                            (m * np.abs(S.acc)).sum(0)])
    assert (np.abs(lev_g - lev_o) / scale).max() <= 1e-13
    assert (np.abs(out_g - out_o)[1:] / (scale[1:] / scale[0])).max() <= 1e-12
    if dim == 2:
        assert np.all(lev_g[mlevel:, [3, 6, 9]] == 0.0)


def test_fix_positions_freeze_and_determinism():
    from oracle import oracle as O
    n = 100_000
    c, S = _setup(n, 5, 2, rtrunc=1.0)
    lev_o, _ = O.fix_positions(S, 0, 2)
    a, _ = c.fix_positions(0, 2)
    b, _ = c.fix_positions(0, 2)
    assert np.array_equal(a, b)                                  # no atomics: bit-reproducible
    assert np.abs(a[:, 0] - lev_o[:, 0]).max() <= 1e-13 * lev_o[:, 0].max()
    # the frozen particles (r > rtrunc) are excluded: the mass is well below the total
    assert a[:, 0].sum() < 0.8 * S.mass.sum()


def test_fix_positions_after_steps_uses_storage_order():
    """The store is physically re-ordered by the sort; the sums must not care."""
    from exp_b200 import slgrid
    from exp_b200.sphere import Sphere
    from oracle import oracle as O
    n = 50_000
    c, S = _setup(n, 9, 0)
    tab = slgrid.make_slgrid(slgrid.HernquistModel(), lmax=2, nmax=6, numr=400, rmin=1e-3, rmax=50.0,
                             refine=4, richardson=False)
    s = Sphere(tab)
    s.determine_coefficients(c)
    s.set_coefficients(None)
    c.zero_pot_acc()
    s.get_acceleration_and_potential(c)
    out = c.download()
    S2 = O.ParticleSet(out["pos"], out["vel"], S.mass)
    S2.acc[:] = out["acc"]
    lev_o, out_o = O.fix_positions(S2, 0, 0)
    lev_g, out_g = c.fix_positions(0, 0)
    assert np.abs(out_g - out_o).max() <= 1e-12 * max(1.0, np.abs(out_o).max())
