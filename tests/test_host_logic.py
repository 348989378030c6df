"""Host-side logic of the drop-in boundary: axis, input packing, factory, calculators, day sharding (CPU only)."""
import numpy as np
import pytest

from cvar_b200 import msm_layout
from cvar_b200.axis import build_axis, segment_counts
from cvar_b200.inputs import HotPathInputs, make_inputs
from cvar_b200.distributed import shard_bounds
from cvar_b200 import synthetic as syn


def test_axis_shape_and_segments():
    for n, marg, counts in ((100, "single", (12, 20, 36)), (2048, "mixture", (512, 292, 440)), (64, "single", (8, 12, 24))):
        assert segment_counts(n, marg) == counts
        x, dx = build_axis(n, marg)
        assert x.shape == dx.shape == (n,)
        assert x[0] == -5.0 and x[-1] == 5.0 and np.all(np.diff(x) > 0)
        assert dx[0] == dx[1] and np.array_equal(dx[1:], np.diff(x))


def test_inputs_validation():
    x, dx = build_axis(32, "single")
    with pytest.raises(ValueError):
        HotPathInputs("clayton", "single", 32, x, dx, sigma=np.ones((2, 2)))
    with pytest.raises(ValueError):
        HotPathInputs("gaussian", "single", 32, x, dx)
    with pytest.raises(ValueError):
        HotPathInputs("gaussian", "single", 32, x, dx, sigma=np.ones((2, 3)))
    with pytest.raises(ValueError):
        HotPathInputs("gaussian", "single", 32, x, dx, sigma=np.ones((2, 2)), weights=np.ones(3) / 3)
    with pytest.raises(ValueError):
        HotPathInputs("student", "mixture", 32, x, dx, probs=np.ones((2, 2, 3)), sigma_states=np.ones((2, 4)))
    inp = make_inputs("student", "mixture", 32, probs=np.full((5, 2, 3), 1 / 3), sigma_states=np.ones((2, 3)))
    assert (inp.T, inp.q) == (5, 3) and inp.take_days(slice(1, 3)).T == 2
    assert np.array_equal(inp.copula_params(), [5.3, 0.6])


def test_msm_layout_functions():
    rng = np.random.default_rng(3)
    vs = np.array([syn.msm_vol_states(3, 0.4, 1.1), syn.msm_vol_states(3, 0.55, 1.4)])
    pr = rng.random((2, 6, 8))
    pr /= pr.sum(axis=2, keepdims=True)
    pbs, lv = msm_layout.merge_states(vs, pr)
    assert pbs.shape == (6, 2, 4) and lv.shape == (2, 4)                 # binomial MSM(k): k + 1 distinct levels
    np.testing.assert_allclose(pbs.sum(axis=2), 1.0, atol=1e-15)
    pairs = msm_layout.state_index_pairs(2, 4)
    joint = msm_layout.pair_probabilities(pbs)
    assert pairs.shape == (16, 2) and joint.shape == (6, 16)
    for l, (s0, s1) in enumerate(pairs):
        np.testing.assert_array_equal(joint[:, l], pbs[:, 0, s0] * pbs[:, 1, s1])
    dens, x, dx = msm_layout.state_densities(lv, 48)
    assert dens.shape == (2, 4, 48)
    np.testing.assert_allclose(dens[1, 2], np.exp(-0.5 * (x / lv[1, 2]) ** 2) / (np.sqrt(2 * np.pi) * lv[1, 2]), rtol=1e-15)


def test_synthetic_generators_are_seeded_and_in_range():
    s1, s2 = syn.garch_sigma_path(250), syn.garch_sigma_path(250)
    assert np.array_equal(s1, s2) and s1.shape == (250, 2)
    assert 0.4 < s1.min() and s1.max() < 3.0
    k = syn.kalman_sigma_path(50)
    assert k.shape == (50, 2) and np.all(k > 0)
    pbs, lv = syn.msm_day_params(12, 4)
    assert pbs.shape == (12, 2, 5) and lv.shape == (2, 5)
    np.testing.assert_allclose(pbs.sum(axis=2), 1.0, atol=1e-12)
    P = syn.msm_transition_matrix(3, 0.4, 3.0, 0.3)
    np.testing.assert_allclose(P.sum(axis=1), 1.0, atol=1e-14)


def test_factory_matches_reference_table_including_quirk_q11():
    from utils.factory import ValueAtRiskCalculationFactory as F
    from utils.model_estimation.copula.gaussian_estimation import GaussianCopulaVaR
    from utils.model_estimation.copula.plackett_estimation import PlackettCopulaVaR
    from utils.model_estimation.copula.student_estimation import StudentCopulaVaR
    from utils.model_estimation.model.garch_estimation import GarchEstimation
    from utils.model_estimation.model.mean_reverting_estimation import MeanRevertingEstimation
    from utils.model_estimation.model.msm_estimation import MSMEstimation
    table = {("student", "msm"): (StudentCopulaVaR, MSMEstimation), ("student", "garch"): (StudentCopulaVaR, GarchEstimation),
             ("student", "mean_reverting"): (StudentCopulaVaR, MeanRevertingEstimation),
             ("gaussian", "msm"): (GaussianCopulaVaR, MSMEstimation), ("gaussian", "garch"): (GaussianCopulaVaR, GarchEstimation),
             ("gaussian", "mean_reverting"): (PlackettCopulaVaR, MeanRevertingEstimation),      # Q11
             ("plackett", "msm"): (PlackettCopulaVaR, MSMEstimation), ("plackett", "garch"): (PlackettCopulaVaR, GarchEstimation),
             ("plackett", "mean_reverting"): (PlackettCopulaVaR, MeanRevertingEstimation)}
    for (cop, est), (ccls, mcls) in table.items():
        calc = F.create_var_calculator(cop, est)
        assert type(calc) is ccls and type(calc.estimation_method) is mcls
    assert type(F.create_var_calculator("gaussian", "mean_reverting", strict_reference_quirks=False)) is GaussianCopulaVaR
    for bad in (("clayton", "garch"), ("student", "egarch")):
        with pytest.raises(ValueError, match="Unsupported estimation type."):
            F.create_var_calculator(*bad)


def test_copula_param_packing_round_trips():
    from utils.factory import ValueAtRiskCalculationFactory as F
    R = np.array([[1.0, 0.37], [0.37, 1.0]])
    s = F.create_var_calculator("student", "garch")
    packed = s.copula_integrations_params({"optimized_params": [4.5, 0.0], "corr_matrix": R})
    assert np.array_equal(packed, [4.5, 0.37])
    nu, corr = s.unpack_copula_params(packed)
    assert nu == 4.5 and np.array_equal(corr, R)
    g = F.create_var_calculator("gaussian", "garch")
    nu, corr = g.unpack_copula_params(g.copula_integrations_params({"corr_matrix": R}))
    assert nu is None and np.array_equal(corr, R)
    p = F.create_var_calculator("plackett", "garch")
    assert p.unpack_copula_params(p.copula_integrations_params({"theta": 4.2})) == (4.2, None)


def test_fit_stages_are_flagged_out_of_scope():
    from utils.calc_var_ABC import OutOfScopeStage, VaRCalculationMethod
    from utils.factory import ValueAtRiskCalculationFactory as F
    for est in ("garch", "mean_reverting", "msm"):
        calc = F.create_var_calculator("student", est)
        assert isinstance(calc, VaRCalculationMethod)
        with pytest.raises(OutOfScopeStage):
            calc.model_params_insample({"A": np.zeros(5)})
        with pytest.raises(OutOfScopeStage):
            calc.copula_or_correl_params_insample(np.zeros((5, 2)), np.zeros((5, 2)))


def test_from_forecasts_reproduces_the_reference_attribute_layout():
    """No GPU needed: only the packing into the reference's attribute names / shapes is checked."""
    from utils.calc_var_class import ValueAtRiskCalcualtion, hot_path_inputs_from_attributes
    from utils.factory import ValueAtRiskCalculationFactory as F
    sigma = syn.garch_sigma_path(7)
    v = ValueAtRiskCalcualtion.from_forecasts(F.create_var_calculator("student", "garch"), np.array([5.3, 0.6]),
                                              sigma=sigma, num_points=64, ptf_mean=0.01)
    dens, x, dx, params = v.grids_generations_params
    assert dens.shape == (2, 1, 64) and np.all(dens == 1) and params.shape == (1, 2)
    assert v.integrations_params_static is None and v.integrations_params_t[0] is not None
    assert (v.out_sample_N, v.dim, v.num_points, v.ptf_mean) == (7, 2, 64, 0.01)
    inp = hot_path_inputs_from_attributes(v)
    assert (inp.copula, inp.marginal, inp.nu, inp.rho) == ("student", "single", 5.3, 0.6)
    assert np.array_equal(inp.sigma, sigma)

    vs = np.array([syn.msm_vol_states(2, 0.4, 1.1), syn.msm_vol_states(2, 0.55, 1.4)])
    pr = np.full((2, 5, 4), 0.25)
    m = ValueAtRiskCalcualtion.from_forecasts(F.create_var_calculator("plackett", "msm"), 4.2, state_probs=pr,
                                              vol_states=vs, num_points=56)
    fbs, joint = m.integrations_params_t
    assert fbs.shape == (5, 2, 3) and joint.shape == (5, 9) and m.integrations_params_static.shape == (2, 3)
    dens, x, dx, combos = m.grids_generations_params
    assert dens.shape == (2, 3, 56) and combos.shape == (9, 2)
    inp = hot_path_inputs_from_attributes(m)
    assert (inp.copula, inp.marginal, inp.theta, inp.q) == ("plackett", "mixture", 4.2, 3)


def test_adjust_integral_semantics():
    from utils.calc_var_class import ValueAtRiskCalcualtion as V
    out = V.adjust_integral(np.array([1.0, 2.0]), np.array([10.0, 20.0]), np.array([[-3.0, -2.0], [-2.5, -2.0]]), np.array([-3.0, -3.0]))
    assert np.array_equal(out, [11.0, 18.0])


def test_shard_bounds_cover_all_days_once():
    for T, world in ((1000, 8), (1001, 8), (7, 8), (0, 2), (250, 3)):
        spans = [shard_bounds(T, world, r) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == T
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert max(hi - lo for lo, hi in spans) <= -(-T // world) if T else True


def test_work_model_matches_survey_figures():
    from cvar_b200 import workmodel as wm
    from oracle import var_oracle as vo
    assert wm.F_CELL == vo.F_CELL and wm.F_AXIS_SINGLE == vo.F_AXIS_SINGLE
    assert wm.flops_per_axis_point("student", "mixture", 9) == 2612 + 78 * 9 - 80
    # SURVEY worked example: n = 2048, Student + single-normal, C = 0.382 n^2  ->  ~1.39e8 flop per solve
    c = 0.382 * 2048 ** 2
    assert abs(wm.algorithmic_flops("student", "single", 1, 2048, [c]) - 1.39e8) < 0.02e8
    assert wm.algorithmic_flops("gaussian", "single", 1, 100, [10, 20]) == 35 * 30 + 2 * 2 * 100 * 262


def test_window_slices_cover_the_series_and_match_the_day_shards():
    """Returns slice of a rank == exactly what its rolling windows read (forecast -> solve chaining)."""
    from cvar_b200.distributed import shard_bounds, window_slice
    for T, N, stride, world in [(10, 4, 1, 2), (7, 5, 1, 3), (1000, 1135, 1, 8), (9, 3, 3, 4), (2, 6, 1, 4), (5, 1, 1, 8)]:
        L = (T - 1) * stride + N
        seen_days = []
        for rank in range(world):
            d0, d1, r0, r1 = window_slice(T, N, world, rank, stride)
            assert (d0, d1) == shard_bounds(T, world, rank)
            if d1 == d0:
                assert r1 == r0
                continue
            assert 0 <= r0 < r1 <= L and (r1 - r0 - N) % stride == 0 and (r1 - r0 - N) // stride + 1 == d1 - d0
            # window w of the full series == window w - d0 of the slice
            for w in (d0, d1 - 1):
                assert r0 + (w - d0) * stride == w * stride and w * stride + N <= r1
            seen_days += list(range(d0, d1))
        assert seen_days == list(range(T))


def test_factory_backend_selection(monkeypatch):
    """`create_var_calculator(copula, estimation, backend=None)` + CVAR_BACKEND; this package holds the
    B200 path only, so anything else raises instead of falling back to a CPU path (reference: utils/factory.py:9-31)."""
    from utils.factory import ValueAtRiskCalculationFactory as F

    monkeypatch.delenv("CVAR_BACKEND", raising=False)
    assert F.create_var_calculator("student", "msm").backend == "b200"
    assert F.create_var_calculator("student", "msm", backend="B200").backend == "b200"
    assert F.create_var_calculator("plackett", "garch", "b200").backend == "b200"      # third positional argument
    monkeypatch.setenv("CVAR_BACKEND", "b200")
    assert F.create_var_calculator("gaussian", "garch").backend == "b200"
    monkeypatch.setenv("CVAR_BACKEND", "reference")
    with pytest.raises(ValueError, match="no CPU fallback"):
        F.create_var_calculator("gaussian", "garch")
    assert F.create_var_calculator("gaussian", "garch", backend="b200").backend == "b200"   # the argument wins
    with pytest.raises(ValueError, match="Unsupported backend"):
        F.create_var_calculator("gaussian", "garch", backend="cpu")
    with pytest.raises(ValueError, match="Unsupported estimation type"):
        F.create_var_calculator("clayton", "garch", backend="b200")


def test_dropin_adds_the_backend_keyword_to_a_two_argument_factory(monkeypatch):
    """`install_factory` on a factory with the reference's signature; the patched driver methods dispatch per object."""
    import types
    import cvar_b200.dropin as dropin

    class Factory:                               # the reference's shape: a staticmethod of two arguments
        @staticmethod
        def create_var_calculator(copula_type, estimation_type):
            return types.SimpleNamespace(copula_type=copula_type, estimation_type=estimation_type)

    class Driver:
        def calc_var(self, obj_var=0.05, first_guess=-3, second_guess=(-3.5, -2)):
            return ("cpu path", obj_var)

        def compute_integral(self, bounds):
            return ("cpu path", bounds)

    dropin.install_factory(Factory)
    dropin.install_factory(Factory)              # idempotent
    dropin.install(Driver)
    try:
        monkeypatch.delenv("CVAR_BACKEND", raising=False)
        v = Driver()
        v.VaRCalculationMethod = Factory.create_var_calculator("student", "garch", backend="reference")
        assert dropin.backend_of(v) == "reference" and v.calc_var(obj_var=0.01) == ("cpu path", 0.01)
        assert v.compute_integral("b") == ("cpu path", "b")
        v.VaRCalculationMethod = Factory.create_var_calculator("student", "garch")
        assert dropin.backend_of(v) == "b200"                      # installed drop-in: the GPU unless told otherwise
        monkeypatch.setenv("CVAR_BACKEND", "cpu")
        assert dropin.backend_of(v) == "reference" and v.calc_var() == ("cpu path", 0.05)
        v.VaRCalculationMethod = Factory.create_var_calculator("student", "garch", backend="b200")
        assert dropin.backend_of(v) == "b200"
        with pytest.raises(ValueError, match="Unsupported backend"):
            Factory.create_var_calculator("student", "garch", backend="tpu")
    finally:
        dropin.uninstall(Driver)
        dropin.uninstall_factory(Factory)
    assert Driver().calc_var() == ("cpu path", 0.05) and "_cvar_b200_create" not in Factory.__dict__
    assert Factory.create_var_calculator("a", "b").copula_type == "a"


def test_calc_integral_seam_recognises_the_copula_family():
    """utils.calc_integral.calc_integral (reference: calc_integral.py:8-25): family from the hooks it is handed."""
    from utils.calc_integral.calc_integral import calc_grids_and_integrals_results, copula_family_of
    from utils.factory import ValueAtRiskCalculationFactory as F
    import inspect

    ref_args = ["T", "unique_var_values", "unique_indices", "num_points", "dim", "var_function", "lower_bound", "upper_bound",
                "grids_generations_params", "integrations_params_t", "integrations_params_static", "copula_params",
                "integrated_function", "copula_density", "unpack_copula_params", "weights"]
    assert list(inspect.signature(calc_grids_and_integrals_results).parameters) == ref_args
    for copula, params in (("student", np.array([5.3, 0.6])), ("gaussian", np.array([0.6])), ("plackett", 4.2)):
        m = F.create_var_calculator(copula, "garch")
        assert copula_family_of(m.copula_density, m.unpack_copula_params, params) == copula
        # anonymous hooks: fall back to the shape of what unpack_copula_params returns
        unpack = m.unpack_copula_params
        assert copula_family_of(lambda *a, **k: None, lambda p: unpack(p), params) == copula
    with pytest.raises(NotImplementedError):
        calc_grids_and_integrals_results(1, np.zeros((1, 2)), np.zeros(1, int), 10, 3, None, -5, 5, None, None, None, None,
                                         None, None, None, np.ones(3) / 3)
