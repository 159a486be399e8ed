"""Shared helpers for the parity tests (test code only)."""
import numpy as np

M48 = 2**48 - 1
ALPHA_INV_CACHE = {}


def seed_for_state_at(oracle, target_seed_at_g: int, g: int) -> int:
    """Initial step seed S such that the chain's seed *before the draw at gid g* equals
    target (mod 2^48): s_g = alpha^g S + c(g)  =>  S = (s_g - c(g)) * (alpha^g)^-1."""
    c = oracle.lib().sqo_jump(0, 0, g)
    ag = (oracle.lib().sqo_jump(1, 0, g) - c) & M48
    inv = pow(ag, -1, 2**48)
    return ((target_seed_at_g - c) * inv) & M48


def seed_with_retry_at(oracle, g: int, x: int = 0x1234) -> int:
    """Step seed whose draw at gid g has t1 = x < 2^16 (v1 = 0 -> inf -> redraw, tau_kernel.cl:282)."""
    A, B = 0x5DEECE66D, 0xB
    inv_a = pow(A, -1, 2**48)
    s_g = (((x - B) * inv_a) - g) & M48  # (s_g + g) * A + B == x
    S = seed_for_state_at(oracle, s_g, g)
    # sanity: literal chain reaches it without an earlier event
    import ctypes as C
    s = C.c_uint64(S)
    for k in range(g):
        rec = oracle.Draw()
        oracle.lib().sqo_random(C.byref(s), k, C.byref(rec))
        assert rec.ndraws == 1 and rec.plus_branch == 0
    rec = oracle.Draw()
    oracle.lib().sqo_random(C.byref(s), g, C.byref(rec))
    assert rec.ndraws >= 2
    return S


def maxabs(a, b):
    return float(np.max(np.abs(np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64))))
