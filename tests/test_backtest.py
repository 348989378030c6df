"""Backtest layer: exceedance definition (main.py:73 semantics) and the coverage tests (CPU)."""
import numpy as np
import pandas as pd

from cvar_b200 import backtest as bt
from oracle import var_oracle as vo


def test_portfolio_return_is_the_unweighted_column_mean():
    df = pd.DataFrame({"A": [1.0, -2.0, 0.5], "B": [3.0, -4.0, 1.5]})
    assert np.array_equal(bt.portfolio_returns(df), [2.0, -3.0, 1.0])
    assert np.array_equal(bt.portfolio_returns(df.to_numpy()), [2.0, -3.0, 1.0])


def test_exceedances_match_the_oracle_definition():
    rng = np.random.default_rng(0)
    var = -np.abs(rng.normal(2, 0.3, (2, 500)))
    r = rng.normal(0, 1.2, 500)
    assert bt.exceedances(var[0], r) == vo.exceedances(var[0], r)
    assert list(bt.exceedances(var, r)) == [vo.exceedances(v, r) for v in var]
    assert bt.exceedances(np.array([-1.0, -1.0]), np.array([-1.0, -1.0000001])) == 1      # strict inequality


def test_kupiec_and_christoffersen_known_values():
    # 250 days, 1 % VaR, 7 exceedances: LR_pof = 5.497, rejected at the 5 % level
    k = bt.kupiec_pof(7, 250, 0.01)
    assert abs(k.statistic - 5.497) < 5e-3 and k.p_value < 0.05
    assert bt.kupiec_pof(5, 500, 0.01).statistic < 1e-12                       # exactly on target
    h = np.zeros(500, bool)
    h[[10, 11, 12, 200, 201, 400]] = True                                        # clustered hits
    ind = bt.christoffersen_independence(h)
    assert ind.statistic > 6 and ind.p_value < 0.02
    spread = np.zeros(500, bool)
    spread[::100] = True
    assert bt.christoffersen_independence(spread).p_value > 0.5
    cc = bt.conditional_coverage(h, 0.01)
    assert abs(cc.statistic - (bt.kupiec_pof(6, 500, 0.01).statistic + ind.statistic)) < 1e-12 and cc.dof == 2


def test_report_rows():
    rng = np.random.default_rng(1)
    r = rng.normal(0, 1, 1000)
    var = np.vstack([np.full(1000, -2.326), np.full(1000, -1.645)])
    rows = bt.backtest_report(var, r, [0.01, 0.05])
    assert [row["alpha"] for row in rows] == [0.01, 0.05]
    assert rows[0]["exceedances"] == int(np.sum(r < -2.326)) and rows[1]["days"] == 1000
    assert all(0 <= row["kupiec_p"] <= 1 for row in rows)
