"""ORACLE -- test infrastructure only, never on the product path.

A from-scratch CPU restatement (NumPy, one ufunc call per node) of what the reference
computes on its own CPU backend for the capture -> evaluate path:

    capture   /root/reference/delayrepay/delayarray.py:46-61   (__array_ufunc__)
    builders  /root/reference/delayrepay/delayarray.py:316-336 (pow_ex / create_ex)
    evaluate  /root/reference/delayrepay/cpu.py:13-31          (CpuVisitor, run)
    eager ops /root/reference/delayrepay/delayarray.py:72-90,511-568 (dot, sum, ...)
    views     /root/reference/delayrepay/delayarray.py:114-128 (__setitem__/__getitem__)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.  It is pinned ("parity pinned") in two ways:
  * tests/golden/*.npz were produced by importing the real reference
    (DELAY_CPU=1, PYTHONPATH=/root/reference) with oracle/make_golden.py; and
    tests/test_oracle.py checks this restatement against them bit for bit;
  * oracle/make_golden.py --check runs reference and restatement side by side
    (only possible in the build container, where /root/reference exists).

What is restated (and why it matters for numbers or for the CPU timing):
  * every captured node is evaluated by ONE NumPy ufunc call with NumPy's own promotion
    (cpu.py:19-26) -- unfused, one temporary per node;
  * evaluation is TREE-recursive with no memo (cpu.py:19-26 + visitor.py:7-21): a node
    shared by k parents is recomputed k times; only the root caches (delayarray.py:38-44);
  * x ** k for a Python-int k is rewritten before evaluation into the left-associated
    chain ((x*x)*x)... (delayarray.py:316-324); np.square(x) -> x*x (delayarray.py:330-331);
  * a binary ufunc whose two array operands differ in shape leaves the lazy world: it is
    evaluated eagerly and returns a raw ndarray (delayarray.py:47-55), only for
    matmul/add/multiply/subtract/true_divide (delayarray.py:188-194; KeyError otherwise);
  * np.sum / np.dot / @ / np.max ... force their operand and call NumPy eagerly
    (delayarray.py:72-90, 511-568);
  * x[key] forces x and wraps the NumPy view (delayarray.py:123-128); x[key] = rhs forces
    rhs into a temporary, then assigns (delayarray.py:114-121) -> Jacobi semantics.
Deliberately NOT restated: the `.dot()` method bug (delayarray.py:98-99), the `x**0/1/-1`
quirk and the Scalar hash collisions -- they are reference defects,
listed as documented divergences in DESIGN.md.
"""
from numbers import Number

import numpy as np
import numpy.lib.mixins

_EAGER_BINARY = {   # delayarray.py:188-194
    "matmul": np.matmul, "add": np.add, "multiply": np.multiply,
    "subtract": np.subtract, "true_divide": np.true_divide, "divide": np.true_divide,
}


class Lazy(numpy.lib.mixins.NDArrayOperatorsMixin):
    """A captured expression node (leaf, constant or ufunc application)."""

    __slots__ = ("kind", "func", "kids", "value", "shape", "_cache")

    def __init__(self, kind, func=None, kids=(), value=None, shape=None):
        self.kind, self.func, self.kids, self.value = kind, func, kids, value
        self.shape = shape
        self._cache = None

    # ---- evaluation: cpu.py:13-31 (tree recursion, no memo below the root)
    def _eval(self):
        if self.kind == "leaf":
            return self.value
        if self.kind == "const":
            return self.value
        return self.func(*[k._eval() for k in self.kids])

    def __array__(self, dtype=None, copy=None):     # delayarray.py:38-44 (root caches)
        if self.kind == "leaf":
            return self.value
        if self._cache is None:
            self._cache = self._eval()
        return self._cache

    def get(self):                                  # delayarray.py:101-106
        return self.__array__()

    def run(self):
        self.__array__()

    def __repr__(self):
        return str(self.__array__())

    # ---- capture: delayarray.py:46-61
    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        if len(inputs) > 1:
            lhs, rhs = inputs[0], inputs[1]
            if not isinstance(lhs, Number) and not isinstance(rhs, Number):
                if lhs.shape != rhs.shape and lhs.shape != (0,) and rhs.shape != (0,):
                    return _EAGER_BINARY[ufunc.__name__](np.asarray(lhs), np.asarray(rhs))
        kids = tuple(_as_node(a) for a in inputs)
        return _build(ufunc, kids)

    def __array_function__(self, func, types, args, kwargs):   # delayarray.py:87-90
        if func is np.dot:
            return _dot(*args)
        forced = [np.asarray(a) if isinstance(a, Lazy) else a for a in args]
        out = func(*forced, **kwargs)
        if func in _WRAPPED_RESULTS and isinstance(out, np.ndarray):
            return leaf(out)
        return out

    def __matmul__(self, other):                    # delayarray.py:69-70
        return _dot(self, other)

    def dot(self, other):
        return _dot(self, other)

    def sum(self, *a, **k):                         # delayarray.py:133-134
        return np.sum(self, *a, **k)

    def astype(self, dt):                           # delayarray.py:401-408 (leaf only)
        assert self.kind == "leaf"
        self.value = self.value.astype(dt)
        return self

    @property
    def dtype(self):
        return np.asarray(self).dtype if self.kind != "const" else None

    @property
    def T(self):                                    # delayarray.py:139-143
        return self if len(self.shape) == 1 else leaf(np.transpose(np.asarray(self)))

    def __len__(self):
        return self.shape[0]

    def reshape(self, *a, **k):                     # delayarray.py:111-112
        return leaf(np.asarray(self).reshape(*a, **k))

    # ---- views: delayarray.py:114-128
    def __getitem__(self, key):
        if isinstance(key, Lazy):
            key = np.asarray(key)
        return leaf(np.asarray(self)[key])

    def __setitem__(self, key, item):
        if isinstance(key, Lazy):
            key = np.asarray(key)
        if isinstance(item, Lazy):
            item = np.asarray(item)                 # RHS fully evaluated into a temporary
        np.asarray(self)[key] = item


_WRAPPED_RESULTS = {np.transpose, np.roll, np.repeat, np.tile, np.diagflat}


def leaf(arr):
    arr = np.asarray(arr)
    return Lazy("leaf", value=arr, shape=arr.shape)


def _as_node(x):                                    # delayarray.py:470-479
    if isinstance(x, Lazy):
        return x
    if isinstance(x, Number):
        return Lazy("const", value=x, shape=(0,))   # (0,) sentinel: delayarray.py:441
    if isinstance(x, np.ndarray):
        return leaf(x)
    raise NotImplementedError(type(x))


def _shape(a, b):                                   # delayarray.py:197-213
    return b.shape if a.shape == (0,) else a.shape


def _build(ufunc, kids):                            # delayarray.py:316-336
    name = ufunc.__name__
    if name == "square":
        return Lazy("op", np.multiply, (kids[0], kids[0]), shape=kids[0].shape)
    if name == "power" and kids[1].kind == "const" and isinstance(kids[1].value, int) \
            and not isinstance(kids[1].value, bool) and kids[1].value >= 2:
        base, acc = kids[0], kids[0]
        for _ in range(kids[1].value - 1):
            acc = Lazy("op", np.multiply, (acc, base), shape=base.shape)
        return acc
    if len(kids) == 1:
        return Lazy("op", ufunc, kids, shape=kids[0].shape)
    return Lazy("op", ufunc, kids, shape=_shape(kids[0], kids[1]))


def _dot(a, b):                                     # delayarray.py:72-85
    return leaf(np.dot(np.asarray(_as_node(a)), np.asarray(_as_node(b))))


# ---- drop-in module surface used by the workloads (delayarray.py:571-644)
def _wrap(fn):
    def made(*a, **k):
        return leaf(fn(*a, **k))
    made.__name__ = fn.__name__
    return made


array, asarray, ones, zeros, full, empty, arange, linspace = (
    _wrap(f) for f in (np.array, np.asarray, np.ones, np.zeros, np.full, np.empty,
                       np.arange, np.linspace))
add, subtract, multiply, dot, sum = np.add, np.subtract, np.multiply, np.dot, np.sum
sqrt, exp, log, sin, cos, tan, power, square = (np.sqrt, np.exp, np.log, np.sin, np.cos,
                                                np.tan, np.power, np.square)
tanh, sinh, cosh, arctan2, abs = np.tanh, np.sinh, np.cosh, np.arctan2, np.abs
newaxis, pi, float32, double = np.newaxis, np.pi, np.float32, np.double
