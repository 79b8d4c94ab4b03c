"""CPU checks that pin the oracle's small-matrix, kernel and constants restatements against independent
implementations (numpy / scipy / closed forms).  The reference ships no golden vectors."""
import numpy as np
import pytest

from oracle import oracle as orc


@pytest.mark.parametrize("dim", [2, 3])
def test_qr_inverse_matches_numpy(dim):
    rng = np.random.default_rng(0)
    for _ in range(200):
        a = rng.normal(size=(dim, dim))
        ok, inv = orc.qr_inverse(a)
        assert ok
        np.testing.assert_allclose(inv, np.linalg.inv(a), rtol=1e-9, atol=1e-9 * np.abs(np.linalg.inv(a)).max())


@pytest.mark.parametrize("dim", [2, 3])
def test_qr_inverse_singular_is_not_invertible(dim):
    z = np.zeros((dim, dim))
    assert not orc.qr_inverse(z)[0]
    a = np.ones((dim, dim))  # rank 1
    assert not orc.qr_inverse(a)[0]
    if dim == 3:
        r = np.array([[1.0, 2.0, 3.0], [2.0, 4.0, 6.0], [0.0, 1.0, 5.0]])  # rank 2
        assert not orc.qr_inverse(r)[0]
    # a single neighbour gives L = c * R R^T: rank 1 -> identity fallback in dSPH_PreStep
    rvec = np.arange(1.0, dim + 1.0)
    assert not orc.qr_inverse(0.3 * np.outer(rvec, rvec))[0]


@pytest.mark.parametrize("dim", [2, 3])
def test_min_eigenvalue_matches_eigvalsh(dim):
    rng = np.random.default_rng(1)
    for _ in range(300):
        a = rng.normal(size=(dim, dim))
        s = a + a.T
        lam = orc.min_eigenvalue(s)
        ref = np.linalg.eigvalsh(s)[0]
        assert abs(lam - ref) <= 1e-12 * max(1.0, np.abs(s).max())
    # lower triangle only is read (Eigen selfadjointView<Lower>)
    s = np.array([[2.0, 99.0], [0.5, 1.0]]) if dim == 2 else np.array([[2.0, 9.0, 9.0], [0.5, 1.0, 9.0], [0.1, 0.2, 3.0]])
    low = np.tril(s) + np.tril(s, -1).T
    assert abs(orc.min_eigenvalue(s) - np.linalg.eigvalsh(low)[0]) < 1e-12
    assert orc.min_eigenvalue(np.eye(dim)) == pytest.approx(1.0, abs=1e-15)
    assert orc.min_eigenvalue(np.zeros((dim, dim))) == 0.0


@pytest.mark.parametrize("dim", [2, 3])
def test_wendland_kernel_is_normalised(dim):
    H = 0.7
    Wc = 7.0 / (4.0 * np.pi * H * H) if dim == 2 else 21.0 / (16.0 * np.pi * H**3)
    r = np.linspace(0.0, 2.0 * H, 20001)
    w = np.array([orc.kernel(x, H, Wc, dim) for x in r])
    shell = 2.0 * np.pi * r if dim == 2 else 4.0 * np.pi * r * r
    integral = np.trapezoid(w * shell, r)
    assert integral == pytest.approx(1.0, rel=1e-6)
    assert orc.kernel(2.0 * H, H, Wc, dim) == 0.0
    assert orc.kernel(0.0, H, Wc, dim) == pytest.approx(Wc)


def test_set_values_constants_3d():
    dx, c, rho0 = 0.0015, 100.0, 810.0
    p = orc.default_params(3, particle_step=dx, speed_sound=c, rho_rest=rho0, mu=0.000142, sig=0.0256)
    assert p.B == pytest.approx(rho0 * c * c / 7.0)
    assert p.H == 2.0 * dx and p.H_sq == p.H * p.H and p.sr == 4.0 * p.H_sq
    assert p.dx == pytest.approx(dx)  # press_pipe = 0 -> rho_pipe = rho_rest
    assert p.sim_mass == pytest.approx(rho0 * dx**3) and p.bnd_mass == p.sim_mass
    assert p.rho_max == pytest.approx(1.5 * rho0) and p.rho_min == pytest.approx(0.5 * rho0)
    assert p.dsph_cont == pytest.approx(2.0 * 0.1 * p.H * c)
    assert p.nu == pytest.approx(0.000142 / rho0)
    assert p.W_correc == pytest.approx(21.0 / (16.0 * np.pi * p.H**3))
    assert p.W_dx == pytest.approx(p.W_correc * (1 - 0.25) ** 4 * 2.0)
    assert p.nb_beta == 0.25 and p.nb_gamma == 0.5
    assert p.aero_L == pytest.approx(dx * np.cbrt(3.0 / (4.0 * np.pi)))
    assert p.A_sphere == pytest.approx(np.pi * p.aero_L**2) and p.A_plate == pytest.approx(dx * dx)
    assert p.interp_fac == 2.0
    assert p.sos == pytest.approx(np.sqrt(298.0 * 287.0 * 1.403))
    # strict '<' on a 4dx-radius lattice ball: 251 interior points + up to 6 axis ties decided in the last bit
    assert 251 <= p.n_full <= 257 and p.i_n_full == 1.0 / p.n_full
    assert p.delta_t == 2e-10


def test_set_values_constants_2d():
    p = orc.default_params(2, particle_step=0.02, speed_sound=125.0)
    assert p.W_correc == pytest.approx(7.0 / (4.0 * np.pi * p.H**2))
    assert p.sim_mass == pytest.approx(1000.0 * 0.02**2)
    assert p.aero_L == pytest.approx(0.02 / np.sqrt(np.pi))
    assert list(p.grav)[:2] == [0.0, -9.81]
    assert 45 <= p.n_full <= 49
