"""CPU ORACLE of the per-day VaR solve  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this module; the product
path (``copula-msm-and-copula-garch-var_b200/``) never does and has no CPU
fallback.

It is a NumPy/SciPy restatement of the reference's algorithm for the hot path
(SURVEY.md §8 rows a1-a16, App. A), each function citing the reference lines
it follows.  Third-party arithmetic the reference delegates to SciPy
(`scipy.special.erf`, `scipy.stats.norm.ppf`, `scipy.stats.t.ppf`; the
reference does not pin a SciPy version, this image has 1.18.1) is delegated
to the same SciPy functions here.

PARITY PIN: the reference ships no tests or golden vectors for this path.
The pin is the reference itself, executed unmodified in the build container
by ``tests/golden/make_golden.py``; its outputs are committed under
``tests/golden/*.npz`` and ``tests/test_oracle_golden.py`` checks this module
against every one of them (VaR vectors bit-for-bit, strip masses to 1e-15).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
from scipy import special as _sp
from scipy import stats as _st

# bracket / bisection constants  (reference: utils/calc_var_class.py:95,111-112,257)
FIRST_GUESS = -3.0
SECOND_GUESS = (-3.5, -2.0)
MIN_VAR = -7.5
MAX_VAR = 0.0
TOL = 1e-6
NEG_INF_PROXY = -100.0          # calc_var_class.py:114
CLIP_LO = -5.0                  # calc_var_class.py:201  (lower_bound of the grid)

CASE_A, CASE_B, CASE_C, CASE_D, CASE_UNDEFINED = 0, 1, 2, 3, 4


# ----------------------------------------------------------------------------
# a11  utils/utils.py:4-42
# ----------------------------------------------------------------------------
def norm_cdf(z):
    """Phi via erf, absolute (not relative) tail accuracy -- quirk Q14 (utils/utils.py:17-22)."""
    return 0.5 * (1 + _sp.erf(z / np.sqrt(2)))


def norm_pdf_std(z):
    """(1/(1*sqrt(2 pi))) * exp(-z^2/2)   (utils/utils.py:36-42 with std=1)."""
    return (1 / (1 * np.sqrt(2 * np.pi))) * np.exp(-0.5 * z ** 2)


# ----------------------------------------------------------------------------
# a12-a14 copula densities on per-cell arrays
# ----------------------------------------------------------------------------
def _corr(rho):
    return np.array([[1.0, rho], [rho, 1.0]])


def gaussian_quantile(u):
    """norm.ppf  (copulas/gaussian/gaussian.py:43-44)."""
    return _st.norm.ppf(u)


def student_quantile(u, nu):
    """t.ppf(u, df=nu)  (copulas/student/student.py:100-102)."""
    return _st.t.ppf(u, df=nu)


def gaussian_copula_density_from_y(y0, y1, rho):
    """N2(y; R) / (phi(y0) phi(y1))  (copulas/gaussian/gaussian.py:47-117)."""
    R = _corr(rho)
    sinv = np.linalg.inv(R)
    det = np.linalg.det(R)
    with np.errstate(all="ignore"):
        v0 = y0 * sinv[0, 0] + y1 * sinv[1, 0]
        v1 = y0 * sinv[0, 1] + y1 * sinv[1, 1]
        quad = v0 * y0 + v1 * y1
        term1 = 1 / (np.sqrt((2 * np.pi) ** 2 * det))
        mv = term1 * np.exp(-0.5 * quad)
        uni0 = (1 / np.sqrt(2 * np.pi)) * np.exp(-0.5 * y0 ** 2)
        uni1 = (1 / np.sqrt(2 * np.pi)) * np.exp(-0.5 * y1 ** 2)
        return mv / (uni0 * uni1)


def student_copula_density_from_y(y0, y1, nu, rho):
    """t2(y; R, nu) / (t1(y0) t1(y1)); non-finite y -> 0/0 = NaN
    (copulas/student/student.py:106-174, quirk Q10)."""
    R = _corr(rho)
    sinv = np.linalg.inv(R)
    det = np.linalg.det(R)
    d = 2
    finite = np.isfinite(y0) & np.isfinite(y1)
    with np.errstate(all="ignore"):
        v0 = y0 * sinv[0, 0] + y1 * sinv[1, 0]
        v1 = y0 * sinv[0, 1] + y1 * sinv[1, 1]
        quad = v0 * y0 + v1 * y1
        term1 = math.gamma((nu + d) / 2) / (math.gamma(nu / 2) * ((nu * np.pi) ** (d / 2)) * np.sqrt(det))
        mv = np.where(finite, term1 * (1 + quad / nu) ** (-(nu + d) / 2), 0.0)
        gterm = math.gamma((nu + 1) / 2) / (np.sqrt(nu * np.pi) * math.gamma(nu / 2))
        uni0 = np.where(np.isfinite(y0), gterm * (1 + (y0 ** 2 / nu)) ** (-(nu + 1) / 2), 0.0)
        uni1 = np.where(np.isfinite(y1), gterm * (1 + (y1 ** 2 / nu)) ** (-(nu + 1) / 2), 0.0)
        return mv / (uni0 * uni1)


def plackett_copula_density(u, v, theta):
    """The reference's (non-textbook, quirk Q9) Plackett density
    (copulas/plackett/plackett.py:66-69)."""
    with np.errstate(all="ignore"):
        num = theta * (1 + (theta - 1) * (u + v - 2 * u * v))
        denom = ((1 + (theta - 1) * (u + v)) * (1 + (theta - 1) * (1 - u - v))) ** 2
        return num / denom


def copula_density(copula, u0, u1, *, rho=None, nu=None, theta=None):
    """c(u0, u1) on arrays of PIT values (used by special-function tests)."""
    if copula == "gaussian":
        return gaussian_copula_density_from_y(gaussian_quantile(u0), gaussian_quantile(u1), rho)
    if copula == "student":
        return student_copula_density_from_y(student_quantile(u0, nu), student_quantile(u1, nu), nu, rho)
    return plackett_copula_density(u0, u1, theta)


# ----------------------------------------------------------------------------
# per-day axis quantities  (App. A.1)
# ----------------------------------------------------------------------------
@dataclass
class DayAxes:
    """Everything that depends on (day, axis point) but not on the cell."""
    u: tuple            # (u0[n], u1[n])   PIT values per grid column
    y: tuple            # copula quantiles of u (None for Plackett)
    pdf: tuple          # single: phi(z)/sigma per column; mixture: None
    mixw: tuple         # mixture: a_d[i] = dx[i] * sum_s p[d,s] N(x[i];0,sigma[1-d,s])  (Q3)


def day_axes(inp, t) -> DayAxes:
    x = inp.x
    if inp.marginal == "single":
        # integration_functions/garch_integration_function.py:27-38
        sig = inp.sigma[t]
        z = [x / sig[0], x / sig[1]]
        u = tuple(norm_cdf(zz) for zz in z)
        pdf = tuple(norm_pdf_std(z[d]) / sig[d] for d in range(2))
        mixw = None
    else:
        # integration_functions/msm_integration_function.py:32-36 ; weights from
        # create_grids.py:121,143 with densities[current_dim - 1] (quirk Q3)
        p = inp.probs[t]                      # (2, q)
        s = inp.sigma_states                  # (2, q)
        u = tuple(np.sum(p[d][None, :] * norm_cdf(x[:, None] / s[d][None, :]), axis=1) for d in range(2))
        dens = [(1 / (np.sqrt(2 * np.pi) * s[d][None, :])) * np.exp(-0.5 * (x[:, None] / s[d][None, :]) ** 2)
                for d in range(2)]            # msm_estimation.py:322-328, dens[d][i, state]
        mixw = tuple(inp.dx * np.sum(p[d][None, :] * dens[1 - d], axis=1) for d in range(2))
        pdf = None
    if inp.copula == "gaussian":
        y = tuple(gaussian_quantile(uu) for uu in u)
    elif inp.copula == "student":
        y = tuple(student_quantile(uu, inp.nu) for uu in u)
    else:
        y = None
    return DayAxes(u=u, y=y, pdf=pdf, mixw=mixw)


# ----------------------------------------------------------------------------
# a6  strip membership  (create_grids.py:102-110, integration_algo.py:20)
# ----------------------------------------------------------------------------
def inner_bound(x_outer, q, w):
    """g(q) = (q - x_outer*w[1]) / w[0], each operation rounded (quirk Q1)."""
    return (q - x_outer * w[1]) / w[0]


def strip_ranges(inp, lo, hi):
    """Per outer point i0 the half-open inner index range [j0, j1) with
    max(g(lo), -5) < x[j] <= g(hi)   (quirk Q2)."""
    x, w = inp.x, inp.weights
    g_hi = inner_bound(x, hi, w)
    g_lo = np.maximum(inner_bound(x, lo, w), CLIP_LO)
    j1 = np.searchsorted(x, g_hi, side="right")          # #{x <= g_hi}
    j0 = np.searchsorted(x, g_lo, side="right")          # #{x <= g_lo}
    return j0, np.maximum(j1, j0)


def half_plane_count(inp, q):
    """N(q): number of grid cells in the strip (-100, q]  (SURVEY §8(d))."""
    j0, j1 = strip_ranges(inp, NEG_INF_PROXY, q)
    return int(np.sum(j1 - j0))


def _flatten(j0, j1):
    lens = j1 - j0
    total = int(lens.sum())
    i0 = np.repeat(np.arange(len(j0)), lens)
    starts = np.repeat(j0 - (np.cumsum(lens) - lens), lens)
    i1 = np.arange(total) + starts
    return i0, i1


# ----------------------------------------------------------------------------
# a4-a10  strip mass
# ----------------------------------------------------------------------------
def _cell_copula(inp, ax, i0, i1):
    if inp.copula == "gaussian":
        return gaussian_copula_density_from_y(ax.y[0][i0], ax.y[1][i1], inp.rho)
    if inp.copula == "student":
        return student_copula_density_from_y(ax.y[0][i0], ax.y[1][i1], inp.nu, inp.rho)
    return plackett_copula_density(ax.u[0][i0], ax.u[1][i1], inp.theta)


def strip_mass(inp, t, lo, hi, ax: DayAxes | None = None, faithful_mixture: bool = False):
    """S(lo, hi) for day t: what `compute_integral` returns for that day
    (calc_var_class.py:179-212 -> calc_integral.py:8-171 -> integrand).

    ``faithful_mixture`` evaluates the mixture integrand state pair by state
    pair in the reference's order (q^2 columns, msm_integration_function.py:
    41-45); the default uses the algebraically identical separable form of
    SURVEY App. A.1 (same VaR bits, ~q^2 times cheaper).
    """
    if ax is None:
        ax = day_axes(inp, t)
    j0, j1 = strip_ranges(inp, lo, hi)
    i0, i1 = _flatten(j0, j1)
    if i0.size == 0:
        return 0.0
    c = _cell_copula(inp, ax, i0, i1)
    dx = inp.dx
    with np.errstate(all="ignore"):
        if inp.marginal == "single":
            # garch_integration_function.py:38-50
            val = np.nan_to_num(c * (ax.pdf[0][i0] * ax.pdf[1][i1]))
            return float(np.sum(val * ((1.0 * (1.0 * dx[i0])) * (1.0 * dx[i1]))))
        if not faithful_mixture:
            return float(np.sum(c * (ax.mixw[0][i0] * ax.mixw[1][i1])))
        # faithful: per state pair l = s0*q + s1
        p = inp.probs[t]
        s = inp.sigma_states
        q = inp.q
        x = inp.x
        dens = [(1 / (np.sqrt(2 * np.pi) * s[d][:, None])) * np.exp(-0.5 * (x[None, :] / s[d][:, None]) ** 2)
                for d in range(2)]            # dens[d][state, i]
        total = np.empty(q * q)
        for s0 in range(q):
            for s1 in range(q):
                step = (1.0 * (dens[1][s0, i0] * dx[i0])) * (dens[0][s1, i1] * dx[i1])
                total[s0 * q + s1] = np.sum(c * step) * (p[0, s0] * p[1, s1])
        return float(np.sum(total))


def compute_integral(inp, bounds, axes=None, **kw):
    """Vector of strip masses, one (lo, hi) pair per day (a4)."""
    T = inp.T
    out = np.empty(T)
    for t in range(T):
        out[t] = strip_mass(inp, t, bounds[t, 0], bounds[t, 1], None if axes is None else axes[t], **kw)
    return out


# ----------------------------------------------------------------------------
# a1-a3  bracket + bisection  (App. A.4)
# ----------------------------------------------------------------------------
@dataclass
class SolveTrace:
    var: np.ndarray             # (T,) solved q + ptf_mean
    case: np.ndarray            # (T,) 0..3 = A..D, 4 = undefined (R == alpha or NaN)
    iterations: int             # global K actually run (quirk Q7)
    decisions: np.ndarray       # (T,) bit k set <=> after iteration k the running mass was < alpha
    zero_bits: np.ndarray       # (T,) bit k set <=> running mass after iteration k was exactly 0
    cells: np.ndarray           # (T,) grid cells evaluated (the algorithmic work C(solve))
    mass: np.ndarray            # (T,) running mass R after the last iteration


def needed_iterations(width, tol=TOL):
    """Smallest K with width / 2**K <= tol (the while condition of
    calc_var_class.py:278 for one day; widths are dyadic so this is exact)."""
    k = 0
    while width > tol:
        width = width / 2
        k += 1
    return k


def calc_var(inp, alpha, first_guess=FIRST_GUESS, second_guess=SECOND_GUESS, tol=TOL, days=None,
             forced_iterations=None, faithful_mixture=False) -> SolveTrace:
    """VaR vector of `ValueAtRiskCalcualtion.calc_var(obj_var=alpha)`
    (calc_var_class.py:95-177) with the vectorised bisection of :250-309.

    ``forced_iterations`` overrides the global iteration count K (needed to
    reproduce a day's value when it was solved inside a larger batch, Q7).
    """
    T = inp.T
    idx = range(T) if days is None else days
    nT = len(idx)
    axes = [day_axes(inp, t) for t in idx]
    cells = np.zeros(nT, dtype=np.int64)

    def S(k, lo, hi):
        j0, j1 = strip_ranges(inp, lo, hi)
        cells[k] += int(np.sum(j1 - j0))
        return strip_mass(inp, idx[k], lo, hi, axes[k], faithful_mixture=faithful_mixture)

    R = np.empty(nT)
    lower = np.empty(nT)
    upper = np.empty(nT)
    prev_upper = np.empty(nT)
    case = np.empty(nT, dtype=np.int64)
    for k in range(nT):
        f3 = S(k, NEG_INF_PROXY, first_guess)                       # :114-119
        if f3 >= alpha:                                             # :125-129
            lo, hi = second_guess[0], first_guess
        else:
            lo, hi = first_guess, second_guess[1]
        pu = second_guess[0] if lo == second_guess[0] else first_guess      # :132 (Q6)
        s2 = S(k, lo, hi)
        r = f3 + s2 if lo == first_guess else f3 - s2               # adjust_integral vs upper=-3 (:138-142)
        if r > alpha and hi == second_guess[1]:                     # :147-155
            c = CASE_D
        elif r > alpha:
            c = CASE_A
        elif r < alpha and hi == first_guess:
            c = CASE_B
        elif r < alpha and hi == second_guess[1]:
            c = CASE_C
        else:
            c = CASE_UNDEFINED                                      # Q8: np.empty garbage in the reference
        case[k] = c
        R[k] = r
        prev_upper[k] = pu
        if c == CASE_UNDEFINED:
            lower[k] = upper[k] = np.nan
        else:
            lower[k], upper[k] = _case_bracket(c, first_guess, second_guess)
    stack = ~np.isin(upper, list(second_guess))                     # :160
    decisions = np.zeros(nT, dtype=np.int64)
    zero_bits = np.zeros(nT, dtype=np.int64)
    it = 0
    with np.errstate(invalid="ignore"):
        while (np.any(upper - lower > tol) if forced_iterations is None else it < forced_iterations):   # :278
            mid = (lower + upper) / 2
            newR = np.empty(nT)
            for k in range(nT):
                if case[k] == CASE_UNDEFINED:
                    newR[k] = np.nan
                    continue
                a, b = (lower[k], mid[k]) if stack[k] else (mid[k], upper[k])    # :282
                s = S(k, a, b)
                newR[k] = R[k] + s if a == prev_upper[k] else R[k] - s           # :241-246 (Q6)
            zero_bits |= (newR == 0).astype(np.int64) << it
            if forced_iterations is None and np.all(newR == 0):                  # :293-295
                break
            stack = newR < alpha                                                 # :298
            decisions |= stack.astype(np.int64) << it
            lower = np.where(~stack, lower, mid)
            upper = np.where(stack, upper, mid)
            R = newR
            prev_upper = mid
            it += 1
    var = (lower + upper) / 2 + inp.ptf_mean                        # :306, :171
    return SolveTrace(var=var, case=case, iterations=it, decisions=decisions, zero_bits=zero_bits,
                      cells=cells, mass=R)


def _case_bracket(c, first_guess, second_guess):
    return {
        CASE_A: (MIN_VAR, second_guess[0]),
        CASE_B: (second_guess[0], first_guess),
        CASE_C: (second_guess[1], MAX_VAR),
        CASE_D: (first_guess, second_guess[1]),
    }[c]


# ----------------------------------------------------------------------------
# backtest helper (SURVEY §7 "Exceedance counts")
# ----------------------------------------------------------------------------
def exceedances(var, r_ptf):
    """Number of days with realised portfolio return below the VaR level."""
    return int(np.sum(np.asarray(r_ptf) < np.asarray(var)))


# ----------------------------------------------------------------------------
# algorithmic work model; shared by tests and bench.py
# ----------------------------------------------------------------------------
F_CELL = {"gaussian": 35, "student": 80, "plackett": 29}
F_AXIS_SINGLE = {"gaussian": 262, "student": 2612, "plackett": 80}


def flops_per_axis_point(copula, marginal, q):
    f = F_AXIS_SINGLE[copula]
    if marginal == "mixture":
        f += 78 * q - 80
    return f


def algorithmic_flops(copula, marginal, q, n, cells):
    """flops(solve) = C * f_cell + 2 n f_axis."""
    return cells * F_CELL[copula] + 2 * n * flops_per_axis_point(copula, marginal, q)
