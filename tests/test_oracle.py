"""Pins the oracle: oracle/refcpu.py must reproduce, bit for bit, what the REAL reference
(imported from /root/reference by oracle/make_golden.py) produced for every workload."""
import os

import numpy as np
import pytest

import workloads as wl
from oracle import refcpu
from util import assert_bits_equal

GOLDEN = np.load(os.path.join(os.path.dirname(__file__), "golden", "workloads.npz"))


def test_axpy_matches_golden():
    i = wl.make_inputs("axpy", 2048)
    assert_bits_equal(wl.axpy(refcpu, i["a"], refcpu.leaf(i["x"]), refcpu.leaf(i["y"])).get(),
                      GOLDEN["axpy"], "axpy")


def test_black_scholes_matches_golden():
    i = wl.make_inputs("black_scholes", 2048)
    call, put = wl.black_scholes(refcpu, *(refcpu.leaf(i[k]) for k in ("S", "K", "T")))
    assert_bits_equal(call.get(), GOLDEN["bs_call"], "call")
    assert_bits_equal(put.get(), GOLDEN["bs_put"], "put")


def test_reductions_match_golden():
    i = wl.make_inputs("l2", 2048)
    a, b = refcpu.leaf(i["a"]), refcpu.leaf(i["b"])
    assert_bits_equal(np.asarray(wl.l2_distance(refcpu, a, b)), GOLDEN["l2"], "l2")
    assert_bits_equal(wl.dot(refcpu, a, b).get(), GOLDEN["dot"], "dot")
    assert_bits_equal(np.asarray(np.sqrt(wl.dot(refcpu, a, a).get())), GOLDEN["norm"], "norm")


def test_heat_matches_golden():
    u = refcpu.leaf(wl.make_inputs("heat", 64)["u"].copy())
    assert_bits_equal(wl.heat(refcpu, u, 5).get(), GOLDEN["heat"], "heat")


def test_nbody_matches_golden():
    i = wl.make_inputs("nbody", 128)
    acc = wl.nbody_acc(refcpu, refcpu.leaf(i["pos"]), refcpu.leaf(i["m"]))
    assert_bits_equal(np.asarray(acc), GOLDEN["nbody"], "nbody")


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_integer_power_association_matches_golden(dt):
    """x**3 is ((x*x)*x) in the reference (delayarray.py:316-324), not np.power(x, 3)."""
    rng = np.random.default_rng(6)
    name = np.dtype(dt).name
    if dt is np.float64:                       # the generator draws f32 first, then f64
        rng.standard_normal(2048)
    x = rng.standard_normal(2048).astype(dt)
    assert_bits_equal((refcpu.leaf(x) ** 3).get(), GOLDEN[f"pow3_{name}"], "pow3")
    assert_bits_equal((refcpu.leaf(x) ** 5).get(), GOLDEN[f"pow5_{name}"], "pow5")
    assert_bits_equal((np.sin(refcpu.leaf(x)) ** 2 + np.cos(refcpu.leaf(x)) ** 2).get(),
                      GOLDEN[f"fuse_{name}"], "fuse")
    # and the association order is observable: it differs from np.power for some inputs
    assert not np.array_equal((x * x) * x, np.power(x, 3))


def test_oracle_tree_recursion_recomputes_shared_nodes():
    """cpu.py:19-26 has no memo below the root; the port must time the same amount of work."""
    calls = []
    orig = np.sqrt

    class Counting(np.lib.mixins.NDArrayOperatorsMixin):
        pass
    x = refcpu.leaf(np.arange(4.0))
    s = np.sqrt(x)
    expr = s + s * s
    n0 = refcpu.Lazy._eval
    count = {"n": 0}

    def counting(self):
        if self.kind == "op" and self.func is np.sqrt:
            count["n"] += 1
        return n0(self)
    refcpu.Lazy._eval = counting
    try:
        expr.get()
    finally:
        refcpu.Lazy._eval = n0
    assert count["n"] == 3
