"""Error model of the tensor-core FILTER of label propagation (wesup_b200/csrc/label_propagate_tc.cu), on the CPU.

The kernel keeps a labeled row j as a candidate for row u when its APPROXIMATE distance
    approx = |a|^2 - 2 * (a_hi.b_hi + a_hi.b_lo + a_lo.b_hi + n_hi + n_lo),   n = -|b|^2 / 2,
is within 2 * KAPPA * (|a|^2 + max|b|^2) (+ a slack) of the smallest approximate distance seen, KAPPA = 2^-16, and then
evaluates the candidates exactly.  The arg-max of the exact evaluation is the reference's iff the approximate distance
errs by less than KAPPA * (|a|^2 + |b|^2).  This file emulates the split (cvt.rna.tf32), the products and an fp32
accumulation in numpy and checks that bound -- with the margin of 2 the GPU tests require of the measured error -- on
feature distributions far outside the ones the GPU tests draw (large common offsets, tiny and huge scales, collapsed
rows).  Test infrastructure only."""
import numpy as np
import pytest

KAPPA = 2.0 ** -16


def tf32_rna(x: np.ndarray) -> np.ndarray:
    """cvt.rna.tf32.f32: round to 10 explicit mantissa bits, ties away from zero (sign-magnitude: add half, truncate)."""
    bits = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    return ((bits + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)


def split(x):
    hi = tf32_rna(x)
    lo = tf32_rna((x - hi).astype(np.float32))
    return hi, lo


def approx_d2(a: np.ndarray, b: np.ndarray, order: np.ndarray) -> np.ndarray:
    """(n_u, n_l) approximate distances with the kernel's operands; the 98 products of a pair are added in fp32 in the
    given order (the tensor core's own order is not documented: several orders are tried)."""
    a_hi, a_lo = split(a)
    b_hi, b_lo = split(b)
    nb = np.zeros(len(b), np.float32)
    for k in range(b.shape[1]):                                   # the prep kernel's fmaf chain, ascending k
        nb = (b[:, k] * b[:, k] + nb).astype(np.float32)
    n_hi, n_lo = split((-0.5 * nb).astype(np.float32))
    na = np.zeros(len(a), np.float32)
    for k in range(a.shape[1]):
        na = (a[:, k] * a[:, k] + na).astype(np.float32)
    # operand rows of the contraction: [a_hi|a_hi|a_lo|1 1] . [b_hi|b_lo|b_hi|n_hi n_lo]
    A = np.concatenate([a_hi, a_hi, a_lo, np.ones((len(a), 2), np.float32)], axis=1)
    B = np.concatenate([b_hi, b_lo, b_hi, n_hi[:, None], n_lo[:, None]], axis=1)
    acc = np.zeros((len(a), len(b)), np.float32)
    for k in order:                                                # products of TF32 values are exact in fp32 (<= 22 bits)
        acc = (acc + A[:, k, None] * B[None, :, k]).astype(np.float32)
    return (na[:, None] - 2.0 * acc).astype(np.float32), na, nb


def exact_d2(a, b):
    return ((a[:, None, :].astype(np.float64) - b[None, :, :].astype(np.float64)) ** 2).sum(-1)


CASES = {
    "bench (|N(0,1)| * 0.06)": lambda g, n: np.abs(g.standard_normal((n, 32))) * 0.06,
    "relu-like, unit scale": lambda g, n: np.maximum(g.standard_normal((n, 32)), 0.0),
    "uniform [0, 1)": lambda g, n: g.random((n, 32)),
    "large common offset": lambda g, n: 5.0 + 0.01 * g.standard_normal((n, 32)),
    "huge scale": lambda g, n: 300.0 * g.random((n, 32)),
    "tiny scale": lambda g, n: 1e-4 * g.random((n, 32)),
    "sparse": lambda g, n: g.random((n, 32)) * (g.random((n, 32)) < 0.1),
    "collapsed": lambda g, n: np.repeat(g.random((1, 32)), n, axis=0),
}


@pytest.mark.parametrize("name", list(CASES))
def test_filter_error_stays_inside_the_bound(name):
    g = np.random.default_rng(len(name))
    a = CASES[name](g, 96).astype(np.float32)
    b = CASES[name](g, 160).astype(np.float32)
    d2 = exact_d2(a, b)
    worst = 0.0
    for order in (np.arange(98), np.arange(98)[::-1], g.permutation(98)):
        approx, na, nb = approx_d2(a, b, order)
        scale = na[:, None].astype(np.float64) + nb[None, :]
        ratio = np.abs(approx.astype(np.float64) - d2) / np.maximum(scale, 1e-300)
        worst = max(worst, float(ratio[scale > 0].max()) if (scale > 0).any() else 0.0)
    assert worst < KAPPA / 2, (name, worst, KAPPA / 2)


def test_true_argmax_is_always_a_candidate():
    """End to end on the emulation: with the kernel's candidate rule the exact nearest row (lowest index on ties of the
    fp32 similarity) is always among the candidates."""
    g = np.random.default_rng(7)
    for name in ("bench (|N(0,1)| * 0.06)", "large common offset", "collapsed", "sparse"):
        a = CASES[name](g, 64).astype(np.float32)
        b = CASES[name](g, 256).astype(np.float32)
        approx, na, nb = approx_d2(a, b, np.arange(98))
        d2 = exact_d2(a, b).astype(np.float32)
        sim = np.exp(-d2)
        best = sim.argmax(axis=1)                                  # first arg-max = lowest index on ties
        run_min = np.full(len(a), np.inf, np.float32)
        nb_max = np.float32(0)
        cand = np.zeros_like(d2, dtype=bool)
        for t0 in range(0, len(b), 128):                          # tiles in order, running minimum as in the kernel
            sl = slice(t0, t0 + 128)
            nb_max = max(nb_max, nb[sl].max())
            gate = np.minimum(run_min, approx[:, sl].min(axis=1))
            thr = gate + (2.0 * KAPPA * (na + nb_max) + 2.0e-6)
            cand[:, sl] = approx[:, sl] <= thr[:, None]
            run_min = gate
        assert cand[np.arange(len(a)), best].all(), name
