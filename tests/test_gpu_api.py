"""The drop-in Python API on the GPU: `ValueAtRiskCalcualtion` mirror, the in-place patcher, hooks, error paths."""
import types

import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu(cuda_device):
    return cuda_device


def _calculator_for(inp):
    from utils.factory import ValueAtRiskCalculationFactory as F
    est = "garch" if inp.marginal == "single" else "msm"
    return F.create_var_calculator(inp.copula, est)


def _driver_from_golden(name):
    from utils.calc_var_class import ValueAtRiskCalcualtion
    inp, g = load_golden(name)
    kw = dict(sigma=inp.sigma) if inp.marginal == "single" else dict(probs_by_state=inp.probs, sigma_states=inp.sigma_states)
    v = ValueAtRiskCalcualtion.from_forecasts(_calculator_for(inp), inp.copula_params(), num_points=inp.n,
                                              weights=inp.weights, ptf_mean=inp.ptf_mean, **kw)
    return v, inp, g


@pytest.mark.parametrize("name", ["kat1_gaussian_single", "kat1_student_mixture", "plackett_mr_w37_n80",
                                  "student_msm8_w46_n40", "c1_gaussian_garch_n100_T250"])
def test_calc_var_through_the_mirror_api_matches_reference(gpu, name):
    v, inp, g = _driver_from_golden(name)
    for a in g["alphas"]:
        var = v.calc_var(obj_var=float(a))
        assert isinstance(var, np.ndarray) and var.shape == (inp.T,) and var.dtype == np.float64
        assert var.tobytes() == g[f"ref_var_{a}"].tobytes()
    both = v.calc_var_multi([float(a) for a in g["alphas"]])
    for k, a in enumerate(g["alphas"]):
        assert both[k].tobytes() == g[f"ref_var_{a}"].tobytes()          # alpha fusion changes nothing


def test_compute_integral_and_host_driven_bisection(gpu):
    """compute_integral is the reference's seam; driving the reference's own host loop over it must land on
    the same VaR as the in-kernel state machine."""
    v, inp, g = _driver_from_golden("kat2_gaussian_n100")
    np.testing.assert_allclose(v.compute_integral(g["bounds"]), g["ref_strip_mass"], rtol=0, atol=2e-13)
    alpha = 0.05
    T = inp.T
    first = v.compute_integral(np.column_stack([np.full(T, -100.0), np.full(T, -3.0)]))
    lo = np.where(first >= alpha, -3.5, -3.0)
    hi = np.where(first < alpha, -2.0, -3.0)
    b = np.column_stack([lo, hi])
    prev_upper = np.where(lo == -3.5, -3.5, -3.0)
    cur = v.adjust_integral(v.compute_integral(b), first, b, np.full(T, -3.0))
    bb = np.empty((T, 2))
    bb[cur > alpha] = (-7.5, -3.5)
    bb[(cur < alpha) & (hi == -3.0)] = (-3.5, -3.0)
    bb[(cur < alpha) & (hi == -2.0)] = (-2.0, 0.0)
    bb[(cur > alpha) & (hi == -2.0)] = (-3.0, -2.0)
    stack = ~np.isin(bb[:, 1], [-3.5, -2.0])
    host = v.bisection_algorithm(alpha, bb, cur, stack, prev_upper) + v.ptf_mean
    assert host.tobytes() == g["ref_var_0.05"].tobytes()
    assert v.calc_var(alpha).tobytes() == host.tobytes()


def test_dropin_patches_a_reference_shaped_object(gpu):
    """`cvar_b200.dropin.install` reads the reference's attribute layout; emulate the reference's class."""
    from cvar_b200 import dropin
    inp, g = load_golden("kat1_student_single")

    class StudentCopulaVaR:                      # class name is how the reference's calculator is recognised
        pass

    class ValueAtRiskCalcualtion:                # stand-in with the reference's attribute names
        def calc_var(self, obj_var=0.05, first_guess=-3, second_guess=(-3.5, -2)):
            raise AssertionError("should have been replaced")

        def compute_integral(self, bounds):
            raise AssertionError("should have been replaced")

    v = ValueAtRiskCalcualtion()
    v.VaRCalculationMethod = StudentCopulaVaR()
    v.num_points, v.weights, v.dim, v.ptf_mean = inp.n, inp.weights, 2, inp.ptf_mean
    v.copula_params = inp.copula_params()
    v.integrations_params_t = [inp.sigma]
    v.integrations_params_static = None
    v.grids_generations_params = (np.ones((2, 1, inp.n)), inp.x, inp.dx, np.zeros((1, 2)))
    dropin.install(ValueAtRiskCalcualtion)
    try:
        for a in g["alphas"]:
            assert v.calc_var(obj_var=float(a)).tobytes() == g[f"ref_var_{a}"].tobytes()
        np.testing.assert_allclose(v.compute_integral(g["bounds"]), g["ref_strip_mass"], rtol=0, atol=2e-13)
        # one cached plan for all of the above; a refit of the copula parameters must not reuse it
        assert len(v._cvar_b200_plans) == 1
        first = next(iter(v._cvar_b200_plans.values()))
        before = v.calc_var(obj_var=0.05)
        assert next(iter(v._cvar_b200_plans.values())) is first
        v.copula_params = np.array([inp.nu + 3.0, 0.1])
        after = v.calc_var(obj_var=0.05)
        assert len(v._cvar_b200_plans) == 1 and next(iter(v._cvar_b200_plans.values())) is not first
        assert not np.array_equal(before, after)
    finally:
        dropin.uninstall(ValueAtRiskCalcualtion)


@pytest.mark.parametrize("copula,kw", [("gaussian", dict(rho=0.6)), ("gaussian", dict(rho=-0.5)),
                                       ("student", dict(rho=0.6, nu=5.3)), ("student", dict(rho=-0.4, nu=2.5)),
                                       ("plackett", dict(theta=4.2)), ("plackett", dict(theta=0.5))])
def test_copula_density_hook_matches_reference_formulas(gpu, copula, kw):
    from cvar_b200.density import copula_density_gpu
    from oracle import var_oracle as vo
    rng = np.random.default_rng(5)
    u = np.vstack([rng.uniform(0, 1, (2000, 2)), 10.0 ** rng.uniform(-12, -1, (500, 2)), [[0.0, 0.3], [0.4, 1.0], [0.5, 0.5]]])
    got = copula_density_gpu(copula, u, **kw)
    want = vo.copula_density(copula, u[:, 0], u[:, 1], **kw)
    assert np.array_equal(np.isnan(got), np.isnan(want))
    np.testing.assert_allclose(got, want, rtol=5e-12, equal_nan=True)


def test_calculator_hooks_are_callable(gpu):
    from utils.factory import ValueAtRiskCalculationFactory as F
    from oracle import var_oracle as vo
    u = np.array([[0.2, 0.3], [0.7, 0.1]])
    s = F.create_var_calculator("student", "garch")
    nu, corr = s.unpack_copula_params(np.array([5.3, 0.6]))
    np.testing.assert_allclose(s.copula_density(cdf=u, nu=nu, corr_matrix=corr), vo.copula_density("student", u[:, 0], u[:, 1], nu=5.3, rho=0.6), rtol=1e-12)
    p = F.create_var_calculator("plackett", "mean_reverting")
    np.testing.assert_allclose(p.copula_density(cdf=u, nu=4.2, corr_matrix=None), vo.plackett_copula_density(u[:, 0], u[:, 1], 4.2), rtol=1e-14)


def test_error_paths(gpu):
    from cvar_b200.backend import VarPlan
    from cvar_b200._lib import CvarError
    from cvar_b200.inputs import make_inputs
    inp = make_inputs("gaussian", "single", 64, sigma=np.ones((3, 2)))
    with VarPlan(inp) as plan:
        with pytest.raises(CvarError) as e:
            plan.solve(inp.day_params(), [0.0])
        assert e.value.status == -4
        with pytest.raises(CvarError) as e:
            plan.solve(inp.day_params(), [0.01] * 9)
        assert e.value.status == -5
        with pytest.raises(ValueError):
            plan.solve(np.ones((3, 3)), [0.01])
        with pytest.raises(CvarError):
            plan.solve(inp.day_params(), [0.01], forced_iterations=40)
        empty = plan.solve(np.empty((0, 2)), [0.01])
        assert empty.var.shape == (1, 0)
    huge = make_inputs("gaussian", "single", 5000, sigma=np.ones((1, 2)))       # beyond one SM's shared memory
    with pytest.raises(CvarError) as e:
        VarPlan(huge)
    assert e.value.status == -8
    bad = make_inputs("gaussian", "single", 64, rho=1.5, sigma=np.ones((3, 2)))
    with pytest.raises(CvarError) as e:
        VarPlan(bad)
    assert e.value.status == -4


@pytest.mark.parametrize("name", ["kat1_student_mixture", "kat2_gaussian_n100", "plackett_mr_w37_n80"])
def test_calc_grids_and_integrals_results_seam(gpu, name):
    """The lowest pure-function seam (reference: utils/calc_integral/calc_integral.py:8-119), called the way the
    reference's compute_integral calls it (calc_var_class.py:194-212): np.unique'd bounds + inverse indices."""
    from utils.calc_integral.calc_integral import calc_grids_and_integrals_results

    v, inp, g = _driver_from_golden(name)
    bounds = np.asarray(g["bounds"], float)
    uniq, inverse = np.unique(bounds, axis=0, return_inverse=True)
    got = calc_grids_and_integrals_results(
        T=inp.T, unique_var_values=uniq, unique_indices=inverse, num_points=v.num_points, dim=v.dim, var_function=None,
        lower_bound=-5, upper_bound=5, grids_generations_params=v.grids_generations_params,
        integrations_params_t=v.integrations_params_t, integrations_params_static=v.integrations_params_static,
        copula_params=v.copula_params, integrated_function=v.integrated_function, copula_density=v.copula_function,
        unpack_copula_params=v.unpack_copula_params, weights=v.weights)
    assert isinstance(got, np.ndarray) and got.shape == (inp.T,)
    ref = np.asarray(g["ref_strip_mass"], float)
    np.testing.assert_allclose(got, ref, rtol=2e-13, atol=1e-300)
    np.testing.assert_array_equal(got, v.compute_integral(bounds))


def test_quantile_table_budget_and_interleaved_plans(gpu, monkeypatch):
    """(1) A Student-t plan whose quantile table misses its accuracy budget is refused (CVAR_ERR_TABLE), not used silently.
    (2) Two live plans of the same kernel variant with different grids: the dynamic shared-memory opt-in belongs to the
    kernel instantiation, not to a plan, so creating the small plan must not break later launches of the large one."""
    from cvar_b200.backend import VarPlan
    from cvar_b200._lib import CvarError
    from cvar_b200 import synthetic as syn
    from cvar_b200.inputs import make_inputs

    small_in = make_inputs("student", "single", 100, rho=0.6, nu=5.3, sigma=syn.garch_sigma_path(4))
    monkeypatch.setenv("CVAR_TQ_BUDGET", "1e-30")
    with pytest.raises(CvarError) as e:
        VarPlan(small_in)
    assert e.value.status == -9
    monkeypatch.delenv("CVAR_TQ_BUDGET")
    big_in = make_inputs("student", "single", 2048, rho=0.6, nu=5.3, sigma=syn.garch_sigma_path(3))
    with VarPlan(big_in) as big, VarPlan(big_in) as reference_plan:
        expected = reference_plan.solve(big_in.day_params(), [0.01]).var
        with VarPlan(small_in) as small:                          # same variant, ~20x less shared memory
            assert small.info().kernel_variant == big.info().kernel_variant
            assert small.info().tq_table_max_rel_err <= 1e-11
            first = small.solve(small_in.day_params(), [0.01]).var
            np.testing.assert_array_equal(big.solve(big_in.day_params(), [0.01]).var, expected)
            np.testing.assert_array_equal(small.solve(small_in.day_params(), [0.01]).var, first)
        np.testing.assert_array_equal(big.solve(big_in.day_params(), [0.01]).var, expected)
