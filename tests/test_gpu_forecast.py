"""GPU forecast producers against the reference's golden vectors and the oracle."""
import numpy as np
import pytest

from conftest import GOLDEN_DIR

pytestmark = pytest.mark.gpu

G = dict(np.load(GOLDEN_DIR / "forecast_producers.npz"))
MSM = sorted({k.split("__")[0] for k in G if k.startswith("msm")})
GARCH = sorted({k.split("__")[0] for k in G if k.startswith("garch")})
KALMAN = sorted({k.split("__")[0] for k in G if k.startswith("kalman")})


def case(name):
    return {k.split("__")[1]: v for k, v in G.items() if k.startswith(name + "__")}


@pytest.fixture(scope="module")
def fc(cuda_device):
    from cvar_b200 import forecast
    return forecast


@pytest.mark.parametrize("name", MSM)
def test_msm_state_filter_matches_reference(fc, name):
    """Kronecker-factored butterflies sum in a different order than the reference's dense rows: 1e-12 relative."""
    c = case(name)
    prm = fc.MsmParams(float(c["m0"]), float(c["sigma_bar"]), float(c["b"]), float(c["gamma"]))
    pbs, sig, sp, info = fc.msm_forecast(c["series"][None, :], [prm], int(c["k"]), int(c["N"]), return_state_probs=True)
    assert not info["degenerate"] and sp.shape == (1,) + c["ref"].shape
    np.testing.assert_allclose(sp[0], c["ref"], rtol=1e-12, atol=1e-300)
    # merged levels == the reference's sum_forecast_by_state on the reference's own probabilities
    from cvar_b200.msm_layout import merge_states, msm_vol_states
    vols = msm_vol_states(int(c["k"]), prm.m0, prm.sigma_bar)[None, :]
    want_pbs, want_sig = merge_states(vols, c["ref"][None, :, :])
    assert np.array_equal(sig, want_sig)
    np.testing.assert_allclose(pbs, want_pbs, rtol=1e-12, atol=1e-300)


@pytest.mark.parametrize("name", GARCH)
def test_garch_forecast_matches_reference(fc, name):
    c = case(name)
    sigma, _ = fc.garch_forecast(c["series"][None, :], [float(c["omega"])], [c["alpha"]], [c["beta"]], int(c["N"]))
    np.testing.assert_allclose(sigma[:, 0], c["ref"], rtol=4e-16, atol=0)


@pytest.mark.parametrize("name", KALMAN)
def test_kalman_forecast_matches_reference(fc, name):
    c = case(name)
    sigma, info = fc.kalman_forecast(c["series"][None, :], [float(c["a"])], [float(c["l"])], [float(c["q"])], int(c["N"]))
    assert not info["failed"]
    np.testing.assert_allclose(sigma[:, 0], c["ref"], rtol=1e-12)


def test_kalman_adapter_runs_on_the_gpu(fc):
    from utils.model_estimation.model.mean_reverting_estimation import MeanRevertingEstimation
    c = case("kalman_b")
    N, T = int(c["N"]), int(c["T"])
    windows = {f"d{t}": {"A": c["series"][t:t + N]} for t in range(T)}
    params = {"A": {"optimal_params": {"a": float(c["a"]), "l": float(c["l"]), "q": float(c["q"])}}}
    (sigma,) = MeanRevertingEstimation().compute_forecast(windows, params)
    np.testing.assert_allclose(sigma[:, 0], c["ref"], rtol=1e-12)


def test_independent_windows_mode_equals_rolling_mode(fc):
    c = case("msm_k4")
    prm = [fc.MsmParams(float(c["m0"]), float(c["sigma_bar"]), float(c["b"]), float(c["gamma"]))]
    N, T = int(c["N"]), int(c["T"])
    rolling, _, _ = fc.msm_forecast(c["series"][None, :], prm, 4, N)
    windows = np.array([c["series"][t:t + N] for t in range(T)])
    assert np.array_equal(fc.rolling_series(windows), c["series"])
    separate, _, _ = fc.msm_forecast(windows.reshape(1, -1), prm, 4, N, window_stride=N)
    assert np.array_equal(rolling, separate)


def test_returns_to_var_pipeline_matches_the_oracle_pipeline(fc):
    """Two assets, MSM(k=4): GPU filter -> GPU solve  vs  oracle filter -> reference merge -> oracle solve."""
    from cvar_b200 import synthetic as syn
    from cvar_b200.backend import VarPlan
    from cvar_b200.inputs import make_inputs
    from cvar_b200.msm_layout import merge_states
    from oracle import forecast_oracle as fo, var_oracle as vo
    k, N, T = 4, 80, 6
    prm = [fc.MsmParams(0.4, 1.1, 3.0, 0.3), fc.MsmParams(0.55, 1.4, 5.0, 0.2)]
    series = np.array([syn.msm_simulate_returns(T + N - 1, k, p.m0, p.sigma_bar, p.b, p.gamma, 40 + i) for i, p in enumerate(prm)])
    pbs, sig, _ = fc.msm_forecast(series, prm, k, N)
    inp = make_inputs("student", "mixture", 96, rho=0.6, nu=5.3, probs=pbs, sigma_states=sig)
    with VarPlan(inp) as plan:
        gpu = plan.solve(inp.day_params(), [0.01, 0.05])
    sp = np.array([fo.msm_forecast(series[a], k, prm[a].m0, prm[a].sigma_bar, prm[a].b, prm[a].gamma, N, T) for a in range(2)])
    vols = np.array([fo.msm_tables(k, p.m0, p.sigma_bar, p.b, p.gamma)[0] for p in prm])
    o_pbs, o_sig = merge_states(vols, sp)
    o_inp = make_inputs("student", "mixture", 96, rho=0.6, nu=5.3, probs=o_pbs, sigma_states=o_sig)
    for j, a in enumerate((0.01, 0.05)):
        assert np.max(np.abs(gpu.var[j] - vo.calc_var(o_inp, a).var)) <= 1e-7


def test_mirror_adapters_run_their_forecast_stage_on_the_gpu(fc):
    """MSMEstimation.forecasts_array / GarchEstimation.compute_forecast with the reference's argument layout."""
    from utils.model_estimation.model.garch_estimation import GarchEstimation
    from utils.model_estimation.model.msm_estimation import MSMEstimation
    c = case("msm_k3")
    N, T = int(c["N"]), int(c["T"])
    windows = {f"d{t}": {"A": c["series"][t:t + N], "B": c["series"][t:t + N] * 1.0} for t in range(T)}
    params = {tk: {"optimal_params": {"m_0": float(c["m0"]), "sig": float(c["sigma_bar"]), "b": float(c["b"]), "gamma": float(c["gamma"])}}
              for tk in ("A", "B")}
    fa = MSMEstimation().forecasts_array(windows, params, 3)
    assert fa.shape == (2, T, 8)
    np.testing.assert_allclose(fa[0], c["ref"], rtol=1e-12)
    np.testing.assert_allclose(fa[1], c["ref"], rtol=1e-12)
    g = case("garch_21")
    N, T = int(g["N"]), int(g["T"])
    windows = {f"d{t}": {"A": g["series"][t:t + N]} for t in range(T)}
    params = {"A": {"optimal_params": {"best_pq": (2, 1), "best_params": [float(g["omega"]), *g["alpha"], *g["beta"]], "best_bic": 0.0}}}
    (sigma,) = GarchEstimation().compute_forecast(windows, params)
    np.testing.assert_allclose(sigma[:, 0], g["ref"], rtol=4e-16)


def test_from_returns_runs_returns_to_var_on_the_gpu(fc):
    """`ValueAtRiskCalcualtion.from_returns`: centring, rolling windows, GPU forecasts, GPU solve -- against the
    oracle pipeline fed with the reference's data-preparation rules (load_data.py:105-137)."""
    import pandas as pd
    from cvar_b200.inputs import make_inputs
    from oracle import forecast_oracle as fo, var_oracle as vo
    from utils.calc_var_class import ValueAtRiskCalcualtion
    from utils.factory import ValueAtRiskCalculationFactory as F
    rng = np.random.default_rng(8)
    N, T = 120, 7
    raw = pd.DataFrame(rng.standard_normal((N + T, 2)) * [1.0, 1.3] + [0.03, -0.02], columns=["A", "B"])
    params = {"A": {"optimal_params": {"best_pq": (1, 1), "best_params": [0.02, 0.09, 0.89], "best_bic": 0}},
              "B": {"optimal_params": {"best_pq": (2, 1), "best_params": [0.03, 0.05, 0.04, 0.88], "best_bic": 0}}}
    w = np.array([0.4, 0.6])
    v = ValueAtRiskCalcualtion.from_returns(F.create_var_calculator("gaussian", "garch"), raw, N, params, np.array([0.55]),
                                            num_points=80, weights=w)
    got = v.calc_var(obj_var=0.05)
    vals = raw.to_numpy()
    mean = vals[:N].mean(axis=0)
    centred = vals - mean
    sigma = np.column_stack([fo.garch_forecast(centred[:, 0], 0.02, [0.09], [0.89], N, T),
                             fo.garch_forecast(centred[:, 1], 0.03, [0.05, 0.04], [0.88], N, T)])
    inp = make_inputs("gaussian", "single", 80, rho=0.55, weights=w, sigma=sigma, ptf_mean=float(np.sum(mean * w)))
    want = vo.calc_var(inp, 0.05).var
    assert got.shape == (T,) and np.max(np.abs(got - want)) <= 1e-7
    assert v.out_sample_N == T and len(v.out_sample_data) == T and abs(v.ptf_mean - float(np.sum(mean * w))) < 1e-15


def test_device_resident_forecast_feeds_the_solve_without_host_round_trip(fc, cuda_device):
    import torch
    from cvar_b200 import synthetic as syn
    from cvar_b200.backend import VarPlan
    from cvar_b200.inputs import make_inputs
    k, N, T = 5, 90, 8
    prm = [fc.MsmParams(0.4, 1.1, 3.0, 0.3), fc.MsmParams(0.55, 1.4, 5.0, 0.2)]
    series = np.array([syn.msm_simulate_returns(T + N - 1, k, p.m0, p.sigma_bar, p.b, p.gamma, 60 + i) for i, p in enumerate(prm)])
    host_pbs, host_sig, _ = fc.msm_forecast(series, prm, k, N)
    d_pbs, sig, status = fc.msm_forecast_device(torch.from_numpy(series).to(cuda_device), prm, k, N)
    inp = make_inputs("gaussian", "mixture", 64, rho=0.5, probs=host_pbs, sigma_states=host_sig)
    with VarPlan(inp) as plan:
        traj = plan.solve_device(d_pbs, [0.01])
        var, case, iters = plan.finalize_device(traj)
        want = plan.solve(host_pbs, [0.01])
        torch.cuda.synchronize()
    assert int(status.item()) == 0 and np.array_equal(sig, host_sig)
    assert np.array_equal(d_pbs.cpu().numpy(), host_pbs)
    assert np.array_equal(var.cpu().numpy(), want.var)


def test_device_resident_garch_and_kalman_equal_their_host_entry_points(fc, cuda_device):
    import torch
    rng = np.random.default_rng(77)
    N, T = 60, 37
    series = rng.standard_normal((2, T + N - 1)) * np.array([[0.9], [1.3]])
    d_series = torch.from_numpy(series).to(cuda_device)
    g_host, _ = fc.garch_forecast(series, [0.02, 0.03], [[0.09], [0.05, 0.04]], [[0.89], [0.88]], N)
    g_dev = fc.garch_forecast_device(d_series, [0.02, 0.03], [[0.09], [0.05, 0.04]], [[0.89], [0.88]], N)
    k_host, info = fc.kalman_forecast(series, [0.97, 0.95], [0.0, 0.1], [0.15, 0.2], N)
    k_dev, status = fc.kalman_forecast_device(d_series, [0.97, 0.95], [0.0, 0.1], [0.15, 0.2], N)
    assert g_dev.shape == (T, 2) and np.array_equal(g_dev.cpu().numpy(), g_host)
    assert np.array_equal(k_dev.cpu().numpy(), k_host) and bool(status.item()) == info["failed"]
    with pytest.raises(ValueError):
        fc.garch_forecast_device(torch.from_numpy(series), [0.02, 0.03], [[0.09], [0.05]], [[0.89], [0.88]], N)   # host tensor


@pytest.mark.parametrize("family", ["msm", "garch", "kalman"])
def test_returns_to_var_driver_equals_host_forecast_plus_host_solve(fc, cuda_device, family):
    """`var_from_returns_sharded` on one rank: upload returns once, forecast and solve on the device."""
    import torch
    from cvar_b200 import synthetic as syn
    from cvar_b200.backend import VarPlan
    from cvar_b200.distributed import var_from_returns_sharded
    from cvar_b200.inputs import make_inputs
    N, T = 80, 11
    if family == "msm":
        k = 4
        prm = [fc.MsmParams(0.4, 1.1, 3.0, 0.3), fc.MsmParams(0.55, 1.4, 5.0, 0.2)]
        series = np.array([syn.msm_simulate_returns(T + N - 1, k, p.m0, p.sigma_bar, p.b, p.gamma, 70 + i) for i, p in enumerate(prm)])
        pbs, sig, _ = fc.msm_forecast(series, prm, k, N)
        inp = make_inputs("student", "mixture", 72, rho=0.5, nu=6.0, probs=pbs, sigma_states=sig)
        producer = lambda r: fc.msm_forecast_device(r, prm, k, N)[0]
    else:
        series = np.random.default_rng(5).standard_normal((2, T + N - 1)) * np.array([[0.8], [1.2]])
        if family == "garch":
            args = ([0.02, 0.03], [[0.09], [0.08]], [[0.89], [0.90]])
            sigma, _ = fc.garch_forecast(series, *args, N)
            producer = lambda r: fc.garch_forecast_device(r, *args, N)
        else:
            args = ([0.97, 0.95], [0.0, 0.1], [0.15, 0.2])
            sigma, _ = fc.kalman_forecast(series, *args, N)
            producer = lambda r: fc.kalman_forecast_device(r, *args, N)[0]
        inp = make_inputs("plackett" if family == "kalman" else "gaussian", "single", 72, rho=0.5, theta=4.2, sigma=sigma)
    with VarPlan(inp) as plan:
        want = plan.solve(inp.day_params(), [0.01, 0.05], ptf_mean=0.01)
        var, case, iters = var_from_returns_sharded(plan, producer, series, N, [0.01, 0.05], ptf_mean=0.01)
        torch.cuda.synchronize()
    assert var.cpu().numpy().tobytes() == want.var.tobytes()
    assert np.array_equal(case.cpu().numpy(), want.case) and np.array_equal(iters.cpu().numpy(), want.iterations)
