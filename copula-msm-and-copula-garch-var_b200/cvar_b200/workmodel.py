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


# ------------------------------------------------------------------------------------------------
# ISSUED FP64 instructions per cell of the solve kernel's cell loops, by kernel variant (cvar_plan_info_t.kernel_variant):
# what the FP64 pipe actually executes, as opposed to the algorithmic flop convention above.  Counted in the SASS of the
# built library (DFMA / DADD / DMUL per loop trip divided by the cells in flight per trip); tests/test_abi.py re-counts
# them with tools/sass_loops.py so that they cannot go stale.  For the Student-t power variants the first number is the
# one-lookup form of the cell (quadratic form below 2^pow_octaves), the second the two-table form.
FP64_PER_CELL = {0: (10, 10), 1: (19, 19), 2: (8, 8), 3: (10, 11), 4: (11, 12), 5: (12, 13), 6: (13, 14)}
CELLS_IN_FLIGHT = {0: 8, 1: 4, 2: 8, 3: 4, 4: 4, 5: 4, 6: 4}


def issued_fp64_per_cell(kernel_variant: int, pow_octaves: int = 0) -> int:
    fast, slow = FP64_PER_CELL[int(kernel_variant)]
    return fast if pow_octaves > 0 else slow


def cell_fp64_pipe_fraction(kernel_variant: int, pow_octaves: int, cells, kernel_seconds: float, peak_tflops: float) -> float:
    """Share of the FP64 pipe's time that the CELL instructions of one launch account for: cells x issued FP64
    instructions per cell (one instruction = one pipe slot per lane, whatever its flop count) over the slots the pipe
    offers in `kernel_seconds` (peak_tflops is measured with DFMA = 2 flop, hence the factor 2).  A lower bound of the
    pipe utilisation ncu reports (sm__pipe_fp64_cycles_active): it leaves out the axis stage, the row set-up and the
    slots of lanes that are masked off inside a warp's instruction."""
    import numpy as np

    lane_instr = float(np.asarray(cells, dtype=np.float64).sum()) * issued_fp64_per_cell(kernel_variant, pow_octaves)
    return lane_instr / (kernel_seconds * peak_tflops * 1e12 / 2.0)
