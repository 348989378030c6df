"""The drop-in on REAL reference objects, end to end on the GPU: the reference's own `ValueAtRiskCalcualtion` (installed
unmodified under baseline/_ref by baseline/install_reference.sh, or the mounted checkout) is built through the reference's
own factory, patched with `cvar_b200.dropin`, and each object is solved twice -- `backend="reference"` runs the reference's
numba/joblib path (utils/calc_var_class.py:95-177), `backend="b200"` the CUDA library -- in the same process.
Skipped where no reference install is reachable."""
import subprocess
import sys
import textwrap

import pytest

from reference_runner import BUILD_OBJECT, reference_env, reference_root

SCRIPT = BUILD_OBJECT + textwrap.dedent("""
    import contextlib, io, json
    import utils, cvar_b200.dropin as dropin
    assert "copula-msm-and-copula-garch-var_b200" not in utils.__path__[0], utils.__path__      # the reference's package

    dropin.install(ValueAtRiskCalcualtion)
    dropin.install_factory(ValueAtRiskCalculationFactory)
    rng = np.random.default_rng(5)
    cases = [
        dict(copula_type="student", estimation="garch", n=48, copula_params=np.array([5.3, 0.6]), weights=(0.5, 0.5),
             ptf_mean=0.013, sigma=0.8 + 0.8 * rng.random((3, 2))),
        dict(copula_type="gaussian", estimation="msm", n=40, copula_params=np.array([0.45]), weights=(0.4, 0.6),
             probs=rng.dirichlet(np.ones(3), size=(2, 2)), sigma_states=np.array([[0.5, 1.0, 2.1], [0.6, 1.3, 2.4]])),
        dict(copula_type="plackett", estimation="mean_reverting", n=44, copula_params=4.2, sigma=0.7 + rng.random((3, 2))),
    ]
    report = []
    for case in cases:
        out = {}
        for backend in ("reference", "b200"):
            v = build_reference_object(backend=backend, **case)
            assert type(v).__module__ == "utils.calc_var_class" and dropin.backend_of(v) == backend
            with contextlib.redirect_stdout(io.StringIO()):      # the reference prints timings
                var = np.asarray(v.calc_var(obj_var=0.02))
                bounds = np.column_stack((np.full(v.out_sample_N, -3.0), np.full(v.out_sample_N, -1.25)))
                mass = np.asarray(v.compute_integral(bounds))
            out[backend] = (var, mass)
        dvar = float(np.max(np.abs(out["b200"][0] - out["reference"][0])))
        dmass = float(np.max(np.abs(out["b200"][1] - out["reference"][1]) / np.abs(out["reference"][1])))
        report.append(dict(case=case["copula_type"] + "/" + case["estimation"], max_abs_dvar=dvar, max_rel_dmass=dmass))
    print("REPORT " + json.dumps(report))
""")


@pytest.mark.gpu
@pytest.mark.skipif(reference_root() is None, reason="no reference install (baseline/install_reference.sh) on this box")
def test_reference_objects_solve_identically_on_both_backends(cuda_device, tmp_path):
    import json

    out = subprocess.run([sys.executable, "-c", SCRIPT], env=reference_env(tmp_path / "stubs"), cwd=tmp_path,
                         capture_output=True, text=True, timeout=1500)
    assert out.returncode == 0, out.stderr[-3000:]
    line = [ln for ln in out.stdout.splitlines() if ln.startswith("REPORT ")][-1]
    for row in json.loads(line[len("REPORT "):]):
        assert row["max_abs_dvar"] == 0.0, row          # VaR levels: bit-identical to the reference's own CPU path
        assert row["max_rel_dmass"] < 2e-13, row        # strip masses: summation order only
