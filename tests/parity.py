"""Parity metrics of SURVEY.md 8(c), per scattering order.

For every order o (S0, S1, S2) with coefficient block b_o of the float64 oracle and a_o of the CUDA path:
  floored   max |a - b| / max(|b|, tau_o),  tau_o = 1e-3 * max|b_o| per signal  — the pass/fail figure (<= 1e-4)
  unfloored max |a - b| / |b| over the entries with |b| > tau_o                 — reported
  linf      max|a - b| / max|b| per signal, worst signal                        — reported
The floor is per order: order-2 coefficients (3-7e-4 on natural patches) are judged against their own scale,
not against S0's 0.4.  Test infrastructure only."""
import numpy as np

TOL = 1e-4


def order_slices(J, L, max_order=2):
    out = {0: slice(0, 1), 1: slice(1, 1 + J * L)}
    if max_order >= 2:
        out[2] = slice(1 + J * L, 1 + J * L + L * L * J * (J - 1) // 2)
    return out


def parity_report(a, b, J, L, max_order=2):
    """a, b: [nsig, K, ...] (maps [nsig, K, h, w] or pooled [nsig, K]).  Returns {order: {floored, unfloored, linf}}."""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    rep = {}
    for o, sl in order_slices(J, L, max_order).items():
        ao = a[:, sl].reshape(a.shape[0], -1); bo = b[:, sl].reshape(b.shape[0], -1)
        if bo.shape[1] == 0:
            continue
        peak = np.abs(bo).max(axis=1, keepdims=True)
        tau = np.maximum(1e-3 * peak, 1e-30)
        err = np.abs(ao - bo)
        big = np.abs(bo) > tau
        rep[o] = {
            "floored": float((err / np.maximum(np.abs(bo), tau)).max()),
            "unfloored": float((err[big] / np.abs(bo)[big]).max()) if big.any() else 0.0,
            "linf": float((err.max(axis=1, keepdims=True) / np.maximum(peak, 1e-30)).max()),
        }
    return rep


def assert_parity(a, b, J, L, max_order=2, tol=TOL, what=""):
    rep = parity_report(a, b, J, L, max_order)
    for o, r in rep.items():
        assert r["floored"] <= tol, "%s order %d: floored rel err %.3g > %g (%s)" % (what, o, r["floored"], tol, rep)
    return rep


def pooled(ref_maps):
    """(mean, population std) over the spatial axes, like train_and_save_model.py:371-372."""
    return ref_maps.mean(axis=(-2, -1)), ref_maps.std(axis=(-2, -1))
