"""Build-container only (skipped where /root/reference is absent, e.g. on the GPU box): `cvar_b200.dropin` reads the
hot-path inputs from a REAL reference `ValueAtRiskCalcualtion` object, next to the reference's own `utils` package."""
import os
import subprocess
import sys
import textwrap
from pathlib import Path

import pytest

from conftest import PKG_ROOT

REFERENCE = Path("/root/reference")

SCRIPT = textwrap.dedent("""
    import numpy as np
    from utils.factory import ValueAtRiskCalculationFactory          # reference
    from utils.calc_var_class import ValueAtRiskCalcualtion          # reference
    import utils, cvar_b200.dropin as dropin
    assert utils.__path__[0].startswith("/root/reference"), utils.__path__

    def make(copula, est, n, **kw):
        m = ValueAtRiskCalculationFactory.create_var_calculator(copula, est)
        v = object.__new__(ValueAtRiskCalcualtion)
        v.VaRCalculationMethod = m
        v.num_points, v.weights, v.dim, v.ptf_mean = n, np.array([0.3, 0.7]), 2, 0.02
        if est == "msm":
            uvs = np.array([[0.5, 1.0, 2.0], [0.6, 1.2, 2.4]]); fbs = np.full((4, 2, 3), 1 / 3)
            dens, x, dx = m.compute_normal_densities(uvs, n)
            v.grids_generations_params = (dens, x, dx, m.create_vol_combinations(uvs))
            v.integrations_params_t = (fbs, m.compute_forecast_combinations(fbs)); v.integrations_params_static = uvs
        else:
            dens, x, dx = m.compute_normal_densities(2, n)
            v.grids_generations_params = (dens, x, dx, np.zeros((1, 2)))
            v.integrations_params_t = [np.array([[1.0, 1.1], [0.9, 1.3]])]; v.integrations_params_static = None
        v.copula_params = kw["cp"]
        return v, x, dx

    v, x, dx = make("student", "garch", 50, cp=np.array([5.3, 0.6]))
    inp = dropin.inputs_from_reference_object(v)
    assert (inp.copula, inp.marginal, inp.n, inp.nu, inp.rho, inp.ptf_mean) == ("student", "single", 50, 5.3, 0.6, 0.02)
    assert inp.x.tobytes() == x.tobytes() and inp.dx.tobytes() == dx.tobytes() and inp.T == 2
    v, x, dx = make("plackett", "msm", 42, cp=4.2)
    inp = dropin.inputs_from_reference_object(v)
    assert (inp.copula, inp.marginal, inp.theta, inp.q, inp.T) == ("plackett", "mixture", 4.2, 3, 4)
    v, _, _ = make("gaussian", "mean_reverting", 40, cp=np.array([0.4]))      # Q11: the factory hands back Plackett
    assert type(v.VaRCalculationMethod).__name__ == "PlackettCopulaVaR"
    orig = dropin.install(ValueAtRiskCalcualtion)
    assert ValueAtRiskCalcualtion.calc_var is dropin.calc_var and "compute_integral" in orig
    dropin.uninstall(ValueAtRiskCalcualtion)
    assert ValueAtRiskCalcualtion.calc_var is orig["calc_var"]
    print("OK")
""")


@pytest.mark.skipif(not REFERENCE.exists(), reason="the reference checkout is only mounted in the build container")
def test_dropin_reads_a_real_reference_object(tmp_path):
    stubs = tmp_path / "stubs"
    (stubs / "matplotlib").mkdir(parents=True)
    (stubs / "yfinance.py").write_text("")
    (stubs / "matplotlib" / "__init__.py").write_text("")
    (stubs / "matplotlib" / "pyplot.py").write_text("")
    env = dict(os.environ, PYTHONPATH=f"{stubs}:{REFERENCE}:{PKG_ROOT}")
    out = subprocess.run([sys.executable, "-c", SCRIPT], env=env, cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and out.stdout.strip().endswith("OK"), out.stderr[-2000:]
