"""Numerical analysis of the three FP64 exponentials the kernels use instead of libm's exp (gpx_b200/csrc/kfun.cuh:
exp_nonpos, exp_tab, exp2_tab256), done on the CPU: each algorithm is restated operation by operation with exactly
rounded IEEE arithmetic (fma = one rounding of the exact rational a*b + c, via fractions.Fraction) and compared with a
50-digit reference.  The constants below are the ones in the .cuh file (the test reads them from the source so that an
edit there cannot go unnoticed).  This bounds the entry-level error of every kernel matrix the device builds."""
import math
import os
import re
import struct
from decimal import Decimal, getcontext
from fractions import Fraction

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = open(os.path.join(ROOT, "gpx_b200", "csrc", "kfun.cuh")).read()
SHIFT = 6755399441055744.0  # 1.5 * 2^52
getcontext().prec = 50


def fma(a, b, c):
    return float(Fraction(a) * Fraction(b) + Fraction(c))  # exact, rounded once (Fraction -> float rounds to nearest)


def lo_int(d):  # __double2loint
    v = struct.unpack("<q", struct.pack("<d", d))[0] & 0xFFFFFFFF
    return v - (1 << 32) if v & 0x80000000 else v


def scale_pow2(p, k):  # __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p)) for normal results
    bits = struct.unpack("<q", struct.pack("<d", p))[0]
    return struct.unpack("<d", struct.pack("<q", bits + (k << 52)))[0]


def body(name):
    m = re.search(r"double " + name + r"\(.*?\n}\n", SRC, flags=re.S)
    assert m, name
    return re.sub(r"//[^\n]*", "", m.group(0))  # comments carry numbers too


def floats_in(text):
    return [float(v) for v in re.findall(r"(?<![\w.])-?\d+\.\d+(?:e[-+]?\d+)?", text)]


def ulp_err(got, x_exact_exp):
    ref = float(x_exact_exp)
    if ref == 0.0:
        return 0.0
    return abs(Decimal(got) - x_exact_exp) / Decimal(math.ulp(ref))


def exp_nonpos(x):
    c = floats_in(body("exp_nonpos"))
    log2e, ln2_hi, ln2_lo = c[1], c[2], c[3]
    coef = c[4:18]  # 1/13! ... 1/2!, 1, 1
    assert c[0] == SHIFT and abs(log2e - 1.4426950408889634) < 1e-15 and len(coef) == 14 and coef[-2:] == [1.0, 1.0]
    kf = fma(x, log2e, SHIFT)
    ki = lo_int(kf)
    kd = kf - SHIFT
    r = fma(kd, ln2_hi, x)
    r = fma(kd, ln2_lo, r)
    p = coef[0]
    for a in coef[1:]:
        p = fma(p, r, a)
    return 0.0 if x < -708.0 else scale_pow2(p, ki)


def exp_tab(x):
    c = floats_in(body("exp_tab"))
    inv, hi, lo = c[1], c[2], c[3]
    coef = c[4:10]
    assert c[0] == SHIFT and abs(inv - 64 / math.log(2)) < 1e-12 and len(coef) == 6 and coef[-2:] == [1.0, 1.0]
    kf = fma(x, inv, SHIFT)
    ki = lo_int(kf)
    kd = kf - SHIFT
    r = fma(kd, hi, x)
    r = fma(kd, lo, r)
    p = coef[0]
    for a in coef[1:]:
        p = fma(p, r, a)
    p = p * (2.0 ** ((ki & 63) / 64.0))  # table entry exp2(j/64): correctly rounded here, <= 1 ulp on the device
    return 0.0 if ki < -65372 else scale_pow2(p, ki >> 6)


def exp2_tab256(u):
    c = floats_in(body("exp2_tab256"))
    coef = c[1:6]
    assert c[0] == SHIFT and len(coef) == 5 and coef[-1] == 1.0 and abs(coef[-2] - math.log(2) / 256) < 1e-18
    kf = u + SHIFT
    ki = lo_int(kf)
    r = u - (kf - SHIFT)
    p = coef[0]
    for a in coef[1:]:
        p = fma(p, r, a)
    p = p * (2.0 ** ((ki & 255) / 256.0))
    return 0.0 if ki < -261632 else scale_pow2(p, ki >> 8)


def test_exp_nonpos_within_two_ulp():
    rng = np.random.default_rng(0)
    xs = np.concatenate([-rng.uniform(0, 50, 1500), -rng.uniform(50, 700, 300), -np.logspace(-12, 0, 100), [0.0, -1e-300]])
    worst = max(ulp_err(exp_nonpos(float(x)), Decimal(float(x)).exp()) for x in xs)
    assert worst < 2.0, worst


def test_exp_tab_within_two_ulp():
    rng = np.random.default_rng(1)
    xs = np.concatenate([-rng.uniform(0, 50, 1500), -rng.uniform(50, 700, 300), -np.logspace(-12, 0, 100), [0.0]])
    worst = max(ulp_err(exp_tab(float(x)), Decimal(float(x)).exp()) for x in xs)
    assert worst < 2.0, worst


def test_exp2_tab256_within_two_ulp_of_two_to_the_u_over_256():
    """The mat-vec forms u = -d2 * 256 log2(e) with one FMA and evaluates 2^(u/256): the error of the kernel entry is
    this function's error plus the rounding of u (|u| * 2^-53 * ln2/256 relative, i.e. below 1e-14 for d2 <= 700)."""
    rng = np.random.default_rng(2)
    us = np.concatenate([-rng.uniform(0, 256 * 60, 1500), -rng.uniform(0, 256 * 1000, 300), -np.logspace(-9, 2, 100), [0.0]])
    ln2_256 = Decimal(2).ln() / 256
    worst = max(ulp_err(exp2_tab256(float(u)), (Decimal(float(u)) * ln2_256).exp()) for u in us)
    assert worst < 2.0, worst
    scale = floats_in(re.search(r"EXP2_256_SCALE = [^;]*;", SRC).group(0))[0]
    assert abs(scale - 256 / math.log(2)) < 1e-12
