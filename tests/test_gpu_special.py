"""Device special functions against SciPy (the third-party arithmetic the reference uses)."""
import numpy as np
import pytest
from scipy import special, stats

pytestmark = pytest.mark.gpu


def _plan(backend, copula="student", nu=5.3):
    from cvar_b200.inputs import make_inputs
    inp = make_inputs(copula, "single", 64, nu=nu, sigma=np.ones((1, 2)))
    return backend.VarPlan(inp)


@pytest.fixture(scope="module")
def backend(cuda_device):
    from cvar_b200 import backend as be
    return be


def _mp_t_quantile(u, nu):
    """Student-t quantile to 30 digits (mpmath): root of the regularised incomplete beta form of the cdf."""
    import mpmath as mp
    mp.mp.dps = 30
    um = mp.mpf(u)
    low = um < mp.mpf("0.5")
    p = um if low else 1 - um
    if p == mp.mpf("0.5"):
        return 0.0
    nu_m = mp.mpf(nu)
    cdf_lower = lambda t: mp.betainc(nu_m / 2, mp.mpf("0.5"), 0, nu_m / (nu_m + t * t), regularized=True) / 2
    guess = abs(float(stats.t.ppf(float(p), df=nu)))
    tau = mp.findroot(lambda t: mp.log(cdf_lower(t)) - mp.log(p), guess)
    return float(-tau if low else tau)


def _u_grid():
    rng = np.random.default_rng(0)
    u = np.concatenate([
        rng.uniform(0, 1, 4000),
        10.0 ** rng.uniform(-17, -1, 2000),
        1.0 - 10.0 ** rng.uniform(-15.9, -1, 2000),
        [5.551115123125783e-17, 1.1102230246251565e-16, 0.5, 0.25, 0.75, 1 - 1.1102230246251565e-16],
    ])
    return u


@pytest.mark.parametrize("nu", [2.01, 2.5, 4.0, 5.3, 8.11, 30.0, 50.0])
def test_student_quantile_table_and_iterative(backend, nu):
    u = _u_grid()
    ref = stats.t.ppf(u, df=nu)
    with _plan(backend, nu=nu) as plan:
        assert plan.info().tq_table_max_rel_err < 5e-14
        fast = plan.special(0, u)
        slow = plan.special(1, u)
        edge = plan.special(0, np.array([0.0, 1.0, 1.5, -0.1, np.nan]))
    # relative accuracy in the tails, absolute accuracy (in units of 0.1) around the median, where the
    # quantile crosses zero and only its absolute error enters the cell exponents
    scale = np.maximum(np.abs(ref), 0.1)
    err_fast, err_slow = np.abs(fast - ref) / scale, np.abs(slow - ref) / scale
    # the two device routines are independent of each other (table in the normal score vs Newton on the cdf)
    assert np.max(np.abs(fast - slow) / scale) < 5e-14
    # SciPy itself is only good to ~1e-11 at a few spots (closed forms for nu = 4 around the median, nu = 2.5
    # near u = 0.21): agree with it to 1e-10 everywhere, to 1e-13 on 99 % of the grid, and let mpmath
    # arbitrate on the points of largest disagreement.
    assert max(err_fast.max(), err_slow.max()) < 1e-10
    assert np.quantile(err_fast, 0.99) < 1e-13 and np.quantile(err_slow, 0.99) < 1e-13
    worst = np.argsort(err_fast)[-12:]
    exact = np.array([_mp_t_quantile(float(u[i]), nu) for i in worst])
    assert np.max(np.abs(fast[worst] - exact) / np.maximum(np.abs(exact), 0.1)) < 1e-13
    assert np.max(np.abs(slow[worst] - exact) / np.maximum(np.abs(exact), 0.1)) < 1e-13
    assert edge[0] == -np.inf and edge[1] == np.inf and np.all(np.isnan(edge[2:]))


def test_exp2_log2_rcp_primitives(backend):
    rng = np.random.default_rng(1)
    t = np.concatenate([rng.uniform(-1000, 1000, 20000), rng.uniform(-2, 2, 20000), [0.0, -0.5, 0.5, 1023.0, -1021.0]])
    x = np.concatenate([10.0 ** rng.uniform(-300, 300, 20000), rng.uniform(0.5, 2.0, 20000), [1.0, 2.0, 0.75, np.sqrt(2)]])
    with _plan(backend, "gaussian") as plan:
        e = plan.special(2, t)
        l = plan.special(3, x)
        r = plan.special(6, x)
        tiny = plan.special(2, np.array([-1100.0, -5000.0]))
        et = plan.special(7, t)                      # table-assisted variant used by the Gaussian / Student cells
        tiny_t = plan.special(7, np.array([-1100.0, -5000.0]))
    assert np.max(np.abs(et / np.exp2(t) - 1)) < 7e-16
    assert np.all(tiny_t >= 0) and np.all(tiny_t < 1e-300)
    assert np.max(np.abs(e / np.exp2(t) - 1)) < 7e-16
    assert np.max(np.abs(l - np.log2(x)) / np.maximum(np.abs(np.log2(x)), 1e-3)) < 6e-16
    assert np.max(np.abs(r * x - 1)) < 3e-16
    assert np.all(tiny >= 0) and np.all(tiny < 1e-300)


def test_phi_via_erf_reproduces_reference_formula(backend):
    z = np.concatenate([np.linspace(-9, 9, 20001), [-8.3, 8.3, -8.4, 8.4]])
    ref = 0.5 * (1 + special.erf(z / np.sqrt(2)))
    with _plan(backend, "gaussian") as plan:
        got = plan.special(4, z)
        q = plan.special(5, np.array([0.0, 1.0, 0.5, 1e-300, 5.551115123125783e-17]))
    assert np.max(np.abs(got - ref)) < 2.3e-16          # a 2-ulp erf at most moves Phi by one spacing of 2^-53
    # saturation to exactly 0 / 1 (|z| ~ 8.3) happens at the same grid points, give or take a rounding tie
    assert np.sum((got == 0) != (ref == 0)) + np.sum((got == 1) != (ref == 1)) <= 2
    assert q[0] == -np.inf and q[1] == np.inf and q[2] == 0.0
    np.testing.assert_allclose(q[3:], stats.norm.ppf([1e-300, 5.551115123125783e-17]), rtol=1e-14)
