"""GPU parity: the CUDA path (through the C ABI) against the reference's golden vectors and the oracle.

Tolerances (north_star): VaR within 1e-7 absolute in return units with identical exceedance counts.  The VaR
is a dyadic midpoint fixed by ~22 comparison outcomes, so the tests additionally require it to be bit-identical
on every golden case; strip masses (floating-point sums in a different association order and with factored
cell weights) must agree to 2e-13 absolute + 2e-12 relative (the relative part only matters for the
Plackett theta=20 sweep, where the reference's formula is near-singular and "masses" reach ~200).
"""
import numpy as np
import pytest

from conftest import golden_names, load_golden

pytestmark = pytest.mark.gpu

NAMES = golden_names()
VAR_TOL = 1e-7
MASS_TOL = 2e-13
MASS_RTOL = 2e-12


@pytest.fixture(scope="module")
def backend(cuda_device):
    from cvar_b200 import backend as be
    return be


@pytest.mark.parametrize("name", NAMES)
def test_strip_mass_matches_reference(backend, name):
    inp, g = load_golden(name)
    if "bounds" not in g:
        pytest.skip("case has no strip-mass vector")
    with backend.VarPlan(inp) as plan:
        got, cells = plan.strip_mass(inp.day_params(), g["bounds"], return_cells=True)
    ref = g["ref_strip_mass"]
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    np.testing.assert_allclose(got, ref, rtol=MASS_RTOL, atol=MASS_TOL, equal_nan=True)
    from oracle import var_oracle as vo
    want_cells = [int(np.sum(np.subtract(*vo.strip_ranges(inp, lo, hi)[::-1]))) for lo, hi in g["bounds"]]
    assert cells.tolist() == want_cells


@pytest.mark.parametrize("name", NAMES)
def test_var_matches_reference(backend, name):
    inp, g = load_golden(name)
    alphas = [float(a) for a in g["alphas"]]
    with backend.VarPlan(inp) as plan:
        res = plan.solve(inp.day_params(), alphas, ptf_mean=inp.ptf_mean)
    for k, a in enumerate(alphas):
        ref = g[f"ref_var_{a}"]
        assert np.max(np.abs(res.var[k] - ref)) <= VAR_TOL, (name, a)
        assert res.var[k].tobytes() == ref.tobytes(), (name, a, res.var[k], ref)


@pytest.mark.parametrize("name", ["c1_gaussian_garch_n100_T250", "student_garch_n100_T24", "cases_student_single",
                                  "student_msm8_n48", "plackett_mr_w37_n80"])
def test_trace_matches_oracle(backend, name):
    """Bracket ids, iteration counts and evaluated-cell counts equal the oracle's, alpha by alpha."""
    from oracle import var_oracle as vo
    inp, g = load_golden(name)
    alphas = [float(a) for a in g["alphas"]]
    with backend.VarPlan(inp) as plan:
        res = plan.solve(inp.day_params(), alphas, ptf_mean=inp.ptf_mean)
        max_iter = plan.max_iter
    for k, a in enumerate(alphas):
        tr = vo.calc_var(inp, a)
        assert res.iterations[k] == tr.iterations
        assert np.array_equal(res.case[k], tr.case)
        # the kernel always records max_iter iterations; the oracle stops at K
        tr_full = vo.calc_var(inp, a, forced_iterations=max_iter)
        assert np.array_equal(res.cells[k].astype(np.int64), tr_full.cells)


def test_exceedance_counts_config1(backend):
    from oracle import var_oracle as vo
    inp, g = load_golden("c1_gaussian_garch_n100_T250")
    rng = np.random.default_rng(7)
    r_ptf = (rng.standard_normal((inp.T, 2)) * inp.sigma).mean(axis=1)
    alphas = [float(a) for a in g["alphas"]]
    with backend.VarPlan(inp) as plan:
        res = plan.solve(inp.day_params(), alphas)
    for k, a in enumerate(alphas):
        assert vo.exceedances(res.var[k], r_ptf) == vo.exceedances(g[f"ref_var_{a}"], r_ptf)


def test_forced_iterations_reproduce_subset_solves(backend):
    """Q7: a day's value depends on the batch-wide iteration count; forcing K reproduces any batch."""
    from oracle import var_oracle as vo
    inp, g = load_golden("cases_gaussian_single")
    with backend.VarPlan(inp) as plan:
        for K in (19, 20, 21, 22):
            res = plan.solve(inp.day_params(), [0.01], forced_iterations=K)
            tr = vo.calc_var(inp, 0.01, forced_iterations=K)
            assert res.var[0].tobytes() == tr.var.tobytes()


def test_device_resident_path_equals_host_path(backend, cuda_device):
    import torch
    inp, g = load_golden("student_garch_n100_T24")
    alphas = [0.05, 0.01]
    with backend.VarPlan(inp) as plan:
        host = plan.solve(inp.day_params(), alphas, ptf_mean=0.25)
        d_day = torch.from_numpy(inp.day_params()).to(cuda_device)
        traj = plan.solve_device(d_day, alphas)
        var, case, iters = plan.finalize_device(traj, ptf_mean=0.25)
        torch.cuda.synchronize()
    assert np.array_equal(var.cpu().numpy(), host.var)
    assert np.array_equal(case.cpu().numpy(), host.case)
    assert np.array_equal(iters.cpu().numpy(), host.iterations)
