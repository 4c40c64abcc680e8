"""Shape / dtype parity of the NumPy surface, checked WITHOUT a GPU: every handler is called with
random shapes, dtypes, axes and keyword arguments on NumPy arrays and on DelayArrays (inside
engine.dry_run: planning, code generation and NVRTC compilation run, kernels do not), and the
result's shape and dtype must agree with NumPy's.  Values are covered by the GPU tests.

usage: python tools/fuzz_shapes.py [--n 2000] [--seed 0]
"""
import argparse
import os
import sys
import traceback

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

DTYPES = [np.float32, np.float64, np.int32, np.int64, np.uint8, np.bool_, np.int16]


def rand_shape(rng, nd=None):
    nd = int(rng.integers(1, 4)) if nd is None else nd
    return tuple(int(rng.choice([1, 2, 3, 5, 8, 33])) for _ in range(nd))


def make(rng, shape, dtype):
    if dtype == np.bool_:
        return rng.integers(0, 2, shape).astype(bool)
    if np.issubdtype(dtype, np.integer):
        return rng.integers(0, 50, shape).astype(dtype)
    return rng.standard_normal(shape).astype(dtype)


def cases(rng):
    """yield (name, fn(xp_array_maker) -> result) pairs; fn gets `A(host)` that wraps a host array"""
    shape = rand_shape(rng)
    dt = DTYPES[int(rng.integers(len(DTYPES)))]
    dt2 = DTYPES[int(rng.integers(len(DTYPES)))]
    x = make(rng, shape, dt)
    y = make(rng, shape, dt2)
    nd = len(shape)
    ax = int(rng.integers(-nd, nd))
    kd = bool(rng.integers(0, 2))
    axes2 = tuple(sorted({int(rng.integers(0, nd)), int(rng.integers(0, nd))}))
    which = str(rng.choice(["reduce", "argred", "cumsum", "concat", "stack", "reshape", "squeeze", "swap", "take",
                            "compress", "diff", "round", "isclose", "like", "outer", "where", "clip", "anyall",
                            "matmul", "tilerep", "roll", "var", "inplace", "bcast", "transpose"]))
    if which == "reduce":
        red = getattr(np, str(rng.choice(["sum", "prod", "max", "min", "mean"])))
        a = rng.choice([None, ax, "t"])
        a = None if a is None else (axes2 if a == "t" else int(a))
        yield f"{red.__name__}(axis={a}, keepdims={kd}) {dt.__name__}{shape}", lambda A: red(A(x), axis=a, keepdims=kd)
        yield f"sum(dtype=float32) {dt.__name__}{shape}", lambda A: np.sum(A(x), axis=ax, dtype=np.float32)
        yield f"count_nonzero(axis={ax}) {dt.__name__}{shape}", lambda A: np.count_nonzero(A(x), axis=ax)
        yield f"ptp {dt.__name__}{shape}", lambda A: np.ptp(A(x), axis=ax) if dt != np.bool_ else np.sum(A(x))
    elif which == "argred":
        f = getattr(np, str(rng.choice(["argmax", "argmin"])))
        a = None if rng.random() < 0.3 else ax
        yield f"{f.__name__}(axis={a}, keepdims={kd}) {dt.__name__}{shape}", lambda A: f(A(x), axis=a, keepdims=kd)
    elif which == "cumsum":
        a = None if rng.random() < 0.3 else ax
        yield f"cumsum(axis={a}) {dt.__name__}{shape}", lambda A: np.cumsum(A(x), axis=a)
    elif which == "concat":
        a = None if rng.random() < 0.2 else ax
        yield f"concatenate(axis={a}) {dt.__name__},{dt2.__name__}{shape}", lambda A: np.concatenate([A(x), A(y), A(x)], axis=a)
        if nd <= 2:
            yield f"vstack {shape}", lambda A: np.vstack([A(x), A(y)])
            yield f"hstack {shape}", lambda A: np.hstack([A(x), A(y)])
    elif which == "stack":
        a = int(rng.integers(-nd - 1, nd + 1))
        yield f"stack(axis={a}) {shape}", lambda A: np.stack([A(x), A(y)], axis=a)
    elif which == "reshape":
        yield f"ravel {shape}", lambda A: np.ravel(A(x))
        yield f"reshape(-1, last) {shape}", lambda A: np.reshape(A(x), (-1, shape[-1]))
        yield f"flatten {shape}", lambda A: A(x).flatten()
        yield f"expand_dims({ax}) {shape}", lambda A: np.expand_dims(A(x), ax)
    elif which == "squeeze":
        s1 = tuple(1 if rng.random() < 0.5 else n for n in shape)
        z = make(rng, s1, dt)
        yield f"squeeze {s1}", lambda A: np.squeeze(A(z))
        ones = [i for i, n in enumerate(s1) if n == 1]
        if ones:
            yield f"squeeze(axis={ones[0]}) {s1}", lambda A: np.squeeze(A(z), axis=ones[0])
    elif which == "swap":
        b = int(rng.integers(-nd, nd))
        yield f"swapaxes({ax},{b}) {shape}", lambda A: np.swapaxes(A(x), ax, b)
        yield f"moveaxis({ax},{b}) {shape}", lambda A: np.moveaxis(A(x), ax, b)
        yield f"T {shape}", lambda A: A(x).T
    elif which == "take":
        idx = rng.integers(-shape[0], shape[0], rand_shape(rng, int(rng.integers(1, 3))))
        yield f"x[idx{idx.shape}] {dt.__name__}{shape}", lambda A: A(x)[A(idx)]
        yield f"take(axis={ax}) {shape}", lambda A: np.take(A(x), [0, -1], axis=ax)
        yield f"take(flat) {shape}", lambda A: np.take(A(x), idx % x.size)
    elif which == "compress":
        m = make(rng, shape, np.bool_)
        yield f"x[mask] {dt.__name__}{shape}", lambda A: A(x)[A(m)].shape[1:]       # count is data dependent
        yield f"nonzero {shape}", lambda A: len(np.nonzero(A(m)))
        rowm = make(rng, (shape[ax % nd],), np.bool_)
        yield f"compress(axis={ax}) {shape}", lambda A: np.compress(rowm, A(x), axis=ax).shape[:ax % nd]
    elif which == "diff":
        if shape[ax % nd] >= 2:
            yield f"diff(axis={ax}) {dt.__name__}{shape}", lambda A: np.diff(A(x), axis=ax)
    elif which == "round":
        d = int(rng.integers(-1, 4))
        if dt in (np.float32, np.float64) or d >= 0:
            yield f"round({d}) {dt.__name__}{shape}", lambda A: np.round(A(x), d)
    elif which == "isclose":
        if dt != np.bool_ and dt2 != np.bool_:
            yield f"isclose {dt.__name__},{dt2.__name__}{shape}", lambda A: np.isclose(A(x), A(y))
    elif which == "like":
        f = getattr(np, str(rng.choice(["zeros_like", "ones_like", "empty_like"])))
        yield f"{f.__name__} {dt.__name__}{shape}", lambda A: f(A(x))
        yield f"full_like {dt.__name__}{shape}", lambda A: np.full_like(A(x), 3)
        yield f"zeros_like(dtype) {shape}", lambda A: np.zeros_like(A(x), dtype=dt2)
    elif which == "outer":
        yield f"outer {dt.__name__},{dt2.__name__}{shape}", lambda A: np.outer(A(x), A(y))
    elif which == "where":
        m = make(rng, shape, np.bool_)
        s = rng.choice([0.5, 2, True])
        s = float(s) if s == 0.5 else (int(s) if s == 2 else bool(s))
        yield f"where(m, x, {s!r}) {dt.__name__}{shape}", lambda A: np.where(A(m), A(x), s)
        yield f"where(m, x, y) {dt.__name__},{dt2.__name__}{shape}", lambda A: np.where(A(m), A(x), A(y))
    elif which == "clip":
        if dt != np.bool_:
            yield f"clip {dt.__name__}{shape}", lambda A: np.clip(A(x), 1, 3)
    elif which == "anyall":
        yield f"any(axis={ax}, keepdims={kd}) {dt.__name__}{shape}", lambda A: np.any(A(x), axis=ax, keepdims=kd)
        yield f"all {dt.__name__}{shape}", lambda A: np.all(A(x))
    elif which == "matmul":
        k = shape[-1]
        b2 = make(rng, (k, int(rng.choice([1, 3, 8]))), dt2)
        v = make(rng, (k,), dt2)
        if dt != np.bool_ and dt2 != np.bool_ and nd <= 2:
            yield f"x @ B {dt.__name__}{shape} {dt2.__name__}{b2.shape}", lambda A: A(x) @ A(b2)
            yield f"x @ v {dt.__name__}{shape}", lambda A: A(x) @ A(v)
            yield f"dot {dt.__name__}", lambda A: np.dot(A(v), A(v))
    elif which == "tilerep":
        yield f"tile {shape}", lambda A: np.tile(A(x), 2)
        yield f"repeat(axis={ax}) {shape}", lambda A: np.repeat(A(x), 3, axis=ax)
    elif which == "roll":
        yield f"roll(axis={ax}) {shape}", lambda A: np.roll(A(x), 2, axis=ax)
        yield f"roll(flat) {shape}", lambda A: np.roll(A(x), -1)
    elif which == "var":
        if dt != np.bool_:
            yield f"var(axis={ax}) {dt.__name__}{shape}", lambda A: np.var(A(x), axis=ax)
            yield f"std {dt.__name__}{shape}", lambda A: np.std(A(x))
            yield f"average {dt.__name__}{shape}", lambda A: np.average(A(x), axis=ax)
    elif which == "inplace":
        def f(A):
            a = A(x.astype(np.float64))
            a += A(y)
            a *= 2
            return a
        yield f"+= *= {dt2.__name__}{shape}", f
        yield f"out= {shape}", lambda A: np.add(A(x), A(y), out=A(np.zeros(shape, np.result_type(dt, dt2)
                                                                           if np.result_type(dt, dt2) != np.bool_ else np.bool_)))
    elif which == "bcast":
        bs = (int(rng.choice([2, 4])),) + shape
        yield f"broadcast_to {shape}->{bs}", lambda A: np.broadcast_to(A(x), bs)
        yield f"x[None] + y[:, None] {shape}", lambda A: A(x)[None] + A(y)[:, None] if nd == 1 else A(x) + A(y)
    elif which == "transpose":
        perm = tuple(int(p) for p in rng.permutation(nd))
        yield f"transpose{perm} {shape}", lambda A: np.transpose(A(x), perm)
        yield f"T.copy {shape}", lambda A: A(x).T.copy()


def describe(r):
    if isinstance(r, (tuple, list)):
        return tuple(describe(t) for t in r)
    if hasattr(r, "shape") and hasattr(r, "dtype"):
        return (tuple(r.shape), np.dtype(r.dtype).name)
    return r


def force(r):
    if isinstance(r, (tuple, list)):
        for t in r:
            force(t)
    elif hasattr(r, "run"):
        r.run()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=2000)
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()
    import delayrepay_b200 as dr
    from delayrepay_b200 import engine
    bad = total = 0
    with engine.dry_run():
        for s in range(a.seed, a.seed + a.n):
            rng = np.random.default_rng(s)
            for name, fn in cases(rng):
                total += 1
                try:
                    with np.errstate(all="ignore"):
                        want = describe(fn(lambda h: h.copy()))
                except Exception:                               # noqa: BLE001  (not a valid NumPy call)
                    continue
                if any("float16" in str(w) for w in (want if isinstance(want, tuple) else (want,))):
                    continue
                try:
                    got_r = fn(lambda h: dr.array(h))
                    force(got_r)
                    got = describe(got_r)
                except Exception as ex:                         # noqa: BLE001
                    tb = traceback.extract_tb(ex.__traceback__)
                    where = next((f"{os.path.basename(t.filename)}:{t.lineno}" for t in reversed(tb)
                                  if "delayrepay_b200" in t.filename), "?")
                    if isinstance(ex, TypeError) and "float16" in str(ex):
                        continue
                    bad += 1
                    print(f"EXC seed={s} {name}: {type(ex).__name__}: {str(ex)[:150]} @ {where}", flush=True)
                    continue
                if got != want:
                    bad += 1
                    print(f"MISMATCH seed={s} {name}: got {got} want {want}", flush=True)
    print(f"fuzz_shapes: {bad} failing of {total}")


if __name__ == "__main__":
    main()
