"""The INT8 oracle's exact fp32 FMA (oracle/int8_forward.py::fma32) against libm's fmaf, including constructed
double-rounding traps; and the requantisation contract on hand-computed values."""
import ctypes
import ctypes.util

import numpy as np

from oracle.int8_forward import _requant, fma32


def _libm_fmaf():
    lib = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
    lib.fmaf.restype = ctypes.c_float
    lib.fmaf.argtypes = [ctypes.c_float] * 3
    return lib.fmaf


def test_fma32_matches_libm_on_random_and_adversarial_inputs():
    fmaf = _libm_fmaf()
    rng = np.random.default_rng(0)
    n = 20000
    a = rng.integers(-2 ** 31, 2 ** 31, n).astype(np.float32)                     # accumulators as the kernel converts them
    b = (rng.random(n, dtype=np.float32) * np.float32(0.02)).astype(np.float32)   # requantisation multipliers
    c = (rng.standard_normal(n) * 50).astype(np.float32)
    # traps: c chosen so that a*b + c sits within a few float64 ulps of a float32 rounding tie
    p = a[:4000].astype(np.float64) * b[:4000].astype(np.float64)
    near = p.astype(np.float32)
    tie = (near.astype(np.float64) + np.nextafter(near, np.float32(np.inf)).astype(np.float64)) / 2
    c_trap = (tie - p).astype(np.float32)
    # genuine double-rounding traps: (1 + 2^-12)^2 = 1 + 2^-11 + 2^-24 is EXACTLY a float32 tie; an addend far below one
    # float64 ulp disappears from the float64 sum, and only its sign says which neighbour the fused result is
    sc = np.float32(2.0) ** rng.integers(-20, 20, 64).astype(np.float32)
    a_t = (np.float32(1 + 2.0 ** -12) * sc).astype(np.float32)
    b_t = np.full(64, 1 + 2.0 ** -12, np.float32)
    c_t = (np.float32(2.0 ** -80) * sc * np.where(np.arange(64) % 2, 1, -1)).astype(np.float32)
    naive = (a_t.astype(np.float64) * b_t.astype(np.float64) + c_t.astype(np.float64)).astype(np.float32)
    assert (fma32(a_t, b_t, c_t).view(np.uint32) != naive.view(np.uint32)).any()   # the tie repair is exercised
    a = np.concatenate([a, a[:4000], a_t]); b = np.concatenate([b, b[:4000], b_t]); c = np.concatenate([c, c_trap, c_t])
    got = fma32(a, b, c)
    want = np.array([fmaf(float(x), float(y), float(z)) for x, y, z in zip(a, b, c)], dtype=np.float32)
    bad = np.flatnonzero(got.view(np.uint32) != want.view(np.uint32))
    assert bad.size == 0, (bad[:5], a[bad[:5]], b[bad[:5]], c[bad[:5]], got[bad[:5]], want[bad[:5]])
    # and the unfused evaluation is NOT the same function (the test would be vacuous otherwise)
    unfused = (a * b + c).astype(np.float32)
    assert (unfused.view(np.uint32) != want.view(np.uint32)).any()


def test_requant_contract_on_small_values():
    op = dict(m=np.array([0.5, 0.25], np.float32), b=np.array([0.25, -0.75], np.float32), r=np.float32(2.0), relu=True)
    acc = np.array([[[[3]], [[-10]]]], dtype=np.int64)        # [1, 2, 1, 1]
    res = np.array([[[[1]], [[1]]]], dtype=np.int32)
    # c0: 3*0.5+0.25 = 1.75, +1*2 = 3.75 -> rint 4 ; c1: -10*0.25-0.75 = -3.25, +2 = -1.25 -> relu 0
    assert _requant(acc, op, res).reshape(-1).tolist() == [4, 0]
    op["relu"] = False
    assert _requant(acc, op, res).reshape(-1).tolist() == [4, -1]
    assert _requant(np.array([[[[10 ** 6]], [[-(10 ** 6)]]]], dtype=np.int64), op, None).reshape(-1).tolist() == [127, -127]
    # half-way cases round to even: 2.5 -> 2, 3.5 -> 4
    op2 = dict(m=np.array([0.5], np.float32), b=np.array([0.0], np.float32), r=None, relu=False)
    assert _requant(np.array([[[[5]]]], dtype=np.int64), op2, None).item() == 2
    assert _requant(np.array([[[[7]]]], dtype=np.int64), op2, None).item() == 4
