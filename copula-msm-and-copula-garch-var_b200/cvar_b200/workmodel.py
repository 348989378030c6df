"""Algorithmic work model of the VaR solve -- the numerator of `bench.py`'s roofline figure.

    flops(solve) = C * f_cell(copula) + 2 n * f_axis(copula, marginal, q)

C = grid cells the reference's own strip scheme evaluates for that (day, alpha) solve (the kernel counts them);
conventions: DFMA = 2, DADD/DMUL = 1, exp = 30, log = 42, division = 14.  These are properties of the ALGORITHM, fixed
before any kernel was written; the kernel's own FP64 instruction counts per cell are lower.
"""
F_CELL = {"gaussian": 35, "student": 80, "plackett": 29}
F_AXIS_SINGLE = {"gaussian": 262, "student": 2612, "plackett": 80}


def flops_per_axis_point(copula: str, marginal: str, q: int) -> int:
    f = F_AXIS_SINGLE[copula]
    if marginal == "mixture":
        f += 78 * q - 80
    return f


def algorithmic_flops(copula: str, marginal: str, q: int, n: int, cells) -> float:
    """Total algorithmic flops of solves whose evaluated-cell counts are `cells` (scalar or array)."""
    import numpy as np

    cells = np.asarray(cells, dtype=np.float64)
    return float(F_CELL[copula] * cells.sum() + cells.size * 2.0 * n * flops_per_axis_point(copula, marginal, q))
