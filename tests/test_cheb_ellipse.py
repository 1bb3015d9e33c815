"""Host logic of the inner F solve (no GPU): the ellipse form of the Chebyshev-Jacobi polynomial that replaces the
ILU-preconditioned inner GMRES of PreconditionASIMPLE::vmult (reference src/NavierStokes.cpp:978-981), and the small
dense eigenproblem behind the estimate of the imaginary extent of D^-1 F (csrc/nsb_capi.cu: cheb_ellipse_coeffs,
skew_radius_host)."""
import numpy as np
import pytest


def residual_poly(dev, k, lmax, ratio, imag, lam):
    """pi(lambda) = 1 - lambda q(lambda) of the sweep recurrence on the scalar problem lambda z = 1 (Dinv = 1)"""
    it, c1, c2 = dev.cheb_coeffs(k, lmax, ratio, imag)
    lam = np.asarray(lam, complex)
    zold = np.zeros_like(lam)
    z = it * np.ones_like(lam)
    for i in range(1, k):
        zn = z + c1[i] * (z - zold) + c2[i] * (1.0 - lam * z)
        zold, z = z, zn
    return 1.0 - lam * z


@pytest.mark.parametrize("k", [2, 3, 4, 8, 12])
def test_interval_form_is_the_classical_chebyshev_polynomial(pkg, k):
    lmax, ratio = 2.9, max(6.0, (k / 1.1) ** 2)
    lmin = lmax / ratio
    theta, delta = 0.5 * (lmax + lmin), 0.5 * (lmax - lmin)
    lam = np.linspace(lmin, lmax, 101)
    ref = np.polynomial.chebyshev.Chebyshev.basis(k)((theta - lam) / delta) / np.polynomial.chebyshev.Chebyshev.basis(k)(theta / delta)
    got = residual_poly(pkg.device, k, lmax, ratio, 0.0, lam)
    assert np.max(np.abs(got.imag)) == 0.0
    assert np.max(np.abs(got.real - ref)) < 1e-12
    # classical coefficients (three-term form with rho, sigma)
    it, c1, c2 = pkg.device.cheb_coeffs(k, lmax, ratio, 0.0)
    sigma = theta / delta
    rho = 1.0 / sigma
    assert it == pytest.approx(1.0 / theta, rel=1e-15)
    for i in range(1, k):
        rn = 1.0 / (2 * sigma - rho)
        assert c1[i] == pytest.approx(rn * rho, rel=1e-13) and c2[i] == pytest.approx(2 * rn / delta, rel=1e-13)
        rho = rn


@pytest.mark.parametrize("k", [3, 4, 6, 9])
@pytest.mark.parametrize("imag", [0.5, 1.26, 2.0])
def test_ellipse_form_contracts_on_the_whole_ellipse(pkg, k, imag):
    """The interval polynomial exceeds one at the complex eigenvalues of a convection-dominated F (1.26 + 1.49i next
    to lmax = 2.9: the failure of the NACA 10-degree case); the ellipse polynomial is bounded by its convergence
    factor ((a + b) / (theta + sqrt(theta^2 - c2)))^k-ish on and inside the ellipse, upright or not."""
    lmax, ratio = 2.9, max(6.0, (k / 1.1) ** 2)
    lmin = lmax / ratio
    theta, a = 0.5 * (lmax + lmin), 0.5 * (lmax - lmin)
    phi = np.linspace(0, 2 * np.pi, 721)
    boundary = theta + a * np.cos(phi) + 1j * imag * np.sin(phi)
    got = np.abs(residual_poly(pkg.device, k, lmax, ratio, imag, boundary))
    c2 = a * a - imag * imag
    r = (a + imag) / (theta + np.sqrt(theta * theta - c2))
    assert r < 1
    assert got.max() <= 2.0 * r ** k / (1 + r ** (2 * k)) * (1 + 1e-9) + 1e-12
    inside = theta + 0.6 * (boundary - theta)
    assert np.abs(residual_poly(pkg.device, k, lmax, ratio, imag, inside)).max() <= got.max() * (1 + 1e-9)


def test_interval_polynomial_amplifies_the_naca_eigenvalue_and_the_ellipse_one_does_not(pkg):
    lam = np.array([1.26 + 1.49j])
    lmax, k = 2.9, 4
    ratio = (k / 1.1) ** 2
    assert abs(residual_poly(pkg.device, k, lmax, ratio, 0.0, lam)[0]) > 1.5
    assert abs(residual_poly(pkg.device, k, lmax, ratio, 1.25 * 1.5, lam)[0]) < 0.6


def test_skew_radius(pkg):
    rng = np.random.default_rng(5)
    for m in (2, 5, 16, 32):
        H = rng.standard_normal((m, m))
        N = 0.5 * (H - H.T)
        ref = np.linalg.svd(N, compute_uv=False)[0]
        sig, y = pkg.device.skew_radius(H)
        assert sig == pytest.approx(ref, rel=1e-6)
        assert np.linalg.norm(y) == pytest.approx(1.0, rel=1e-12)
        assert np.linalg.norm(N @ y) == pytest.approx(ref, rel=1e-4)
    sig, _ = pkg.device.skew_radius(np.diag([1.0, 2.0, 3.0]))  # symmetric: no skew part
    assert sig == 0.0
