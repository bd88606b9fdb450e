"""CPU check of the error bound the tensor-core rank pass relies on (cornac_b200/csrc/rank_tc.cu: pow2_scale,
user_scale, row_eps).  The kernel nominates every item whose approximate score is >= tau - 2 eps, tau <= the k-th best
approximate score; that is complete iff |approx - exact| <= eps for every (user, item).  Here the approximate pass is
emulated with numpy (operands scaled by the same powers of two, rounded to float16, f32 accumulation, base carried as
a float16 hi/lo pair) on ordinary and on hostile magnitudes, and both the bound and the completeness of the
nomination are asserted.  The formulas are restated from the kernel source; the GPU test
tests/test_rank_tc_gpu.py::test_tensor_pass_matches_fp16_matmul ties the real tensor pass to the same emulation."""
import numpy as np
import pytest


def pow2_scale(m):
    if not m > 0:
        return np.float32(1.0)
    e = np.frexp(np.float32(m))[1]
    return np.float32(2.0) ** np.float32(min(60, max(-60, 14 - int(e))))


def user_scale(row_absmax, sV, sB, bmax):
    sU, cA = pow2_scale(row_absmax), np.float32(0.0)
    if bmax > 0:
        cA = np.float32(sU * sV / sB)
        if cA > 32768.0:
            sU = np.float32(sU * (np.float32(32768.0) / cA))
            cA = np.float32(32768.0)
        if cA < 5.9604645e-8:
            cA = np.float32(0.0)
    return sU, cA, np.float32(sU * sV)


def row_eps(un, vmax, bmax, sU, sV, cA, S, k):
    eps = (0.00098 * un * vmax + 2.99e-8 * np.sqrt(k) * (vmax / sU + un / sV) + 2e-6 * (un * vmax + bmax)
           + 4.8e-7 * bmax + 6e-8 * cA / S)
    if cA == 0:
        eps += bmax * 1.0001
    return eps


def emulate(U, V, base):
    """approx scores (unscaled) and eps per row, as the kernel computes them"""
    n_q, k = U.shape
    vnorm = np.linalg.norm(V.astype(np.float64), axis=1) * 1.0001
    vmax, bmax = float(vnorm.max()), float(np.abs(base).max())
    sV, sB = pow2_scale(np.abs(V).max()), pow2_scale(bmax)
    with np.errstate(over="ignore", under="ignore"):
        Vh = (V * sV).astype(np.float16).astype(np.float32)
        bs = (base * sB).astype(np.float32)
        hi = bs.astype(np.float16).astype(np.float32)
        lo = (bs - hi).astype(np.float16).astype(np.float32)
    approx = np.empty((n_q, V.shape[0]), np.float64)
    eps = np.empty(n_q)
    for q in range(n_q):
        sU, cA, S = user_scale(np.abs(U[q]).max(), sV, sB, bmax)
        with np.errstate(over="ignore", under="ignore"):
            uh = (U[q] * sU).astype(np.float16).astype(np.float32)
        assert np.all(np.isfinite(uh)) and np.all(np.isfinite(Vh))           # the scaling keeps fp16 in range
        acc = (Vh @ uh).astype(np.float32) + np.float32(cA) * hi + np.float32(cA) * lo
        approx[q] = acc.astype(np.float64) / float(S)
        un = float(np.linalg.norm(U[q].astype(np.float64))) * 1.0001
        eps[q] = row_eps(un, vmax, bmax, float(sU), float(sV), float(cA), float(S), k)
    return approx, eps


CASES = ["normal", "tiny", "huge", "mixed_rows", "base_dominates", "base_negligible", "wide_elements", "no_base",
         "adversarial_rounding"]


@pytest.mark.parametrize("case", CASES)
def test_fp16_pass_error_bound_and_nomination_completeness(case):
    rng = np.random.RandomState(len(case))
    n_q, n_items, k, topk = 40, 3000, 64, 50
    U = rng.normal(0, 0.3, (n_q, k)).astype(np.float32)
    V = rng.normal(0, 0.3, (n_items, k)).astype(np.float32)
    base = rng.normal(0, 0.3, n_items).astype(np.float32)
    if case == "tiny":
        U *= np.float32(1e-9); V *= np.float32(1e-7); base *= np.float32(1e-16)
    elif case == "huge":
        U *= np.float32(3e6); V *= np.float32(1e7); base *= np.float32(1e13)
    elif case == "mixed_rows":
        U *= (10.0 ** rng.uniform(-6, 3, (n_q, 1))).astype(np.float32)
        V *= (10.0 ** rng.uniform(-3, 2, (n_items, 1))).astype(np.float32)
    elif case == "base_dominates":
        base *= np.float32(1e9)
    elif case == "base_negligible":
        base *= np.float32(1e-12)
    elif case == "wide_elements":
        U *= (10.0 ** rng.uniform(-8, 0, (n_q, k))).astype(np.float32)
        V *= (10.0 ** rng.uniform(-8, 0, (n_items, k))).astype(np.float32)
    elif case == "no_base":
        base[:] = 0
    elif case == "adversarial_rounding":
        # every element sits half-way between two fp16 neighbours (worst relative rounding error) with aligned signs
        grid = (1.0 + (2 * rng.randint(0, 512, (n_q, k)) + 1) / 2048.0).astype(np.float32)
        U = (grid * rng.choice([0.25, 0.5, 1.0], (n_q, k))).astype(np.float32)
        gridv = (1.0 + (2 * rng.randint(0, 512, (n_items, k)) + 1) / 2048.0).astype(np.float32)
        V = (gridv * rng.choice([0.25, 0.5, 1.0], (n_items, k))).astype(np.float32)
    approx, eps = emulate(U, V, base)
    exact = U.astype(np.float64) @ V.astype(np.float64).T + base.astype(np.float64)[None, :]
    err = np.abs(approx - exact)
    assert np.all(err <= eps[:, None]), (case, float((err / eps[:, None]).max()))
    # the bound is not vacuous on ordinary data: within 50x of the worst observed error
    if case in ("normal", "adversarial_rounding"):
        assert eps.min() < 50 * err.max()
    # nomination completeness: items with approx >= (k-th best approx) - 2 eps contain the exact top-k
    for q in range(n_q):
        tau = np.sort(approx[q])[-topk]
        keep = approx[q] >= tau - 2 * eps[q]
        top = np.argsort(-exact[q], kind="stable")[:topk]
        assert keep[top].all(), (case, q)
