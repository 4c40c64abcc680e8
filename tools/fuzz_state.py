"""Stateful differential fuzzer: random PROGRAMS OF STATEMENTS over a few arrays -- slice and
masked assignment (including right-hand sides that read shifted views of the target), in-place
operators, views created before and used after a mutation, lazy temporaries, reductions --
executed on NumPy arrays and on DelayArrays side by side.  Checks memo invalidation (buffer
versions), the plan cache, the stencil / hazard paths of engine.assign and view aliasing.

usage: python tools/fuzz_state.py [--n 300] [--seed 0] [--only K] [-v]
Lazy temporaries are forced before every mutation: like the reference, the engine evaluates an
expression when it is needed, so `t = a + 1; a[0] = 5; print(t)` sees the new a (reference
delayarray.py:38-44: evaluation happens in __array__).  Arithmetic is + - * / max / min / where
with bounded values, so every comparison is bit-exact; sums are checked to rtol 1e-12.
"""
import argparse
import os
import sys
import traceback

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np


class World:
    def __init__(self, rng, dr):
        self.rng, self.dr = rng, dr
        self.log = []
        nd = int(rng.choice([1, 2]))
        self.shape = (int(rng.choice([5, 64, 257, 4099])),) if nd == 1 else \
            (int(rng.choice([4, 33, 130])), int(rng.choice([8, 65, 256, 300])))
        self.dtype = rng.choice([np.float32, np.float64])
        self.h, self.d = {}, {}
        for name in "abc"[:int(rng.integers(2, 4))]:
            x = rng.uniform(-2, 2, self.shape).astype(self.dtype)
            self.h[name] = x
            self.d[name] = dr.array(x)
        self.temps = []             # (host value, lazy device node, description)
        self.views = []             # (host view, device view, description)

    # ---- random pieces
    def region(self):
        """a tuple of slices selecting a non-empty box, plus the box shape"""
        key = []
        for n in self.shape:
            if self.rng.random() < 0.3 or n < 4:
                key.append(slice(None))
            else:
                lo = int(self.rng.integers(0, n // 2))
                hi = int(self.rng.integers(lo + 1, n + 1))
                step = int(self.rng.choice([1, 1, 1, 2]))
                key.append(slice(lo, hi, step))
        return tuple(key)

    def shifted(self, key, name):
        """the same box moved by -1/0/+1 along each axis where that stays in bounds"""
        out = []
        for k, n in zip(key, self.shape):
            lo, hi, st = k.indices(n)
            s = int(self.rng.choice([-1, 0, 1]))
            if lo + s < 0 or hi + s > n:
                s = 0
            out.append(slice(lo + s, hi + s, st))
        return tuple(out)

    def operand(self, key, target=None):
        """(host, device, text) operand over the box `key`"""
        r = self.rng.random()
        if r < 0.2:
            v = float(self.rng.choice([0.5, -0.25, 1.5, 2.0, 0.1]))
            return v, v, repr(v)
        name = str(self.rng.choice(list(self.h)))
        k = self.shifted(key, name) if (name == target and self.rng.random() < 0.7) or self.rng.random() < 0.3 else key
        return self.h[name][k], self.d[name][k], f"{name}[{_fmt(k)}]"

    def expr(self, key, target=None, depth=2):
        if depth == 0 or self.rng.random() < 0.25:
            return self.operand(key, target)
        (ha, da, ta), (hb, db, tb) = self.expr(key, target, depth - 1), self.expr(key, target, depth - 1)
        if np.isscalar(ha) and np.isscalar(hb):
            ha, da, ta = self.operand(key, target)
        op = str(self.rng.choice(["+", "-", "*", "max", "min", "where", "avg"]))
        if op == "+":
            return ha + hb, da + db, f"({ta} + {tb})"
        if op == "-":
            return ha - hb, da - db, f"({ta} - {tb})"
        if op == "*":
            return ha * hb * 0.5, da * db * 0.5, f"({ta} * {tb} * 0.5)"
        if op == "max":
            return np.maximum(ha, hb), np.maximum(da, db), f"max({ta}, {tb})"
        if op == "min":
            return np.minimum(ha, hb), np.minimum(da, db), f"min({ta}, {tb})"
        if op == "avg":
            return (ha + hb) * 0.5, (da + db) * 0.5, f"(({ta} + {tb}) * 0.5)"
        return np.where(ha > hb, ha, hb * 0.5), np.where(da > db, da, db * 0.5), f"where({ta} > {tb}, .., ..)"

    def force_temps(self):
        for _, node, _ in self.temps:
            if hasattr(node, "run"):
                node.run()

    # ---- statements
    def step(self):
        rng = self.rng
        r = rng.random()
        name = str(rng.choice(list(self.h)))
        if r < 0.22:                                        # slice assignment
            key = self.region()
            h, d, t = self.expr(key, target=name)
            self.force_temps()
            self.log.append(f"{name}[{_fmt(key)}] = {t}")
            self.h[name][key] = h
            self.d[name][key] = d
        elif r < 0.34:                                      # in-place operator
            key = tuple(slice(None) for _ in self.shape)
            h, d, t = self.expr(key, target=name, depth=1)
            self.force_temps()
            op = str(rng.choice(["+=", "-=", "*="]))
            self.log.append(f"{name} {op} {t}")
            ha, da = self.h[name], self.d[name]
            if op == "+=":
                ha += h; da += d
            elif op == "-=":
                ha -= h; da -= d
            else:
                ha *= 0.5; da *= 0.5
            self.d[name] = da
        elif r < 0.44:                                      # masked assignment
            self.force_temps()
            v = float(rng.choice([0.0, 1.0, -0.5]))
            thr = float(rng.choice([-1.0, 0.0, 1.0]))
            self.log.append(f"{name}[{name} > {thr}] = {v}")
            self.h[name][self.h[name] > thr] = v
            self.d[name][self.d[name] > thr] = v
        elif r < 0.56:                                      # a view, kept for later
            key = self.region()
            self.log.append(f"view{len(self.views)} = {name}[{_fmt(key)}]")
            self.views.append((self.h[name][key], self.d[name][key], f"{name}[{_fmt(key)}]"))
        elif r < 0.64 and self.views:                       # write through an old view
            hv, dv, t = self.views[int(rng.integers(len(self.views)))]
            self.force_temps()
            v = float(rng.choice([0.25, -1.0, 3.0]))
            self.log.append(f"({t})[...] = ({t}) * 0.5 + {v}")
            hv[...] = hv * 0.5 + v
            dv[...] = dv * 0.5 + v
        elif r < 0.8:                                       # lazy temporary
            key = self.region()
            h, d, t = self.expr(key, depth=3)
            if not np.isscalar(h) and getattr(d, "kind", "leaf") != "leaf":     # (a bare view aliases)
                self.log.append(f"t{len(self.temps)} = {t}")
                self.temps.append((np.array(h, copy=True), d, t))
        else:                                               # a check in the middle of the program
            return self.check(one=True)
        return None

    def check(self, one=False):
        if getattr(self, "dry", False):
            self.force_temps()
            return None
        items = [(self.h[k], self.d[k], k) for k in self.h] + [(h, d, "view " + t) for h, d, t in self.views] + \
                [(h, d, "temp " + t) for h, d, t in self.temps]
        if one:
            items = [items[int(self.rng.integers(len(items)))]]
        for h, d, what in items:
            got = d.get()
            if got.shape != h.shape or got.dtype != h.dtype:
                return f"SHAPE/DTYPE {what}: {got.shape} {got.dtype} vs {h.shape} {h.dtype}"
            if got.tobytes() != h.tobytes() and not np.array_equal(got, h, equal_nan=True):
                bad = np.argwhere(got != h)
                return f"VALUE {what}: {len(bad)} mismatches, first at {bad[0].tolist()}: got {got[tuple(bad[0])]!r} want {h[tuple(bad[0])]!r}"
            if one and h.size:
                s_got, s_want = float(np.sum(d)), float(np.sum(h.astype(np.float64)))
                if abs(s_got - s_want) > (1e-12 if h.dtype == np.float64 else 1e-5) * float(np.sum(np.abs(h.astype(np.float64))) + 1e-300):
                    return f"SUM {what}: {s_got!r} vs {s_want!r}"
        # temps keep their value once compared: a later mutation of their operands must not change it
        return None


def _fmt(key):
    return ", ".join(f"{k.start if k.start is not None else ''}:{k.stop if k.stop is not None else ''}"
                     + (f":{k.step}" if k.step not in (None, 1) else "") for k in key)


def run_one(seed, verbose=False, dry=False):
    import delayrepay_b200 as dr
    rng = np.random.default_rng(seed)
    w = World(rng, dr)
    msg = None
    try:
        w.dry = dry                 # no values without a device: statements only, no comparisons
        for _ in range(int(rng.integers(6, 24))):
            msg = w.step()
            if msg:
                break
        if not msg and not dry:
            msg = w.check()
    except Exception as ex:                                     # noqa: BLE001
        tb = traceback.extract_tb(ex.__traceback__)
        where = next((f"{os.path.basename(t.filename)}:{t.lineno}" for t in reversed(tb)
                      if "delayrepay_b200" in t.filename), f"{os.path.basename(tb[-1].filename)}:{tb[-1].lineno}")
        msg = f"EXC {type(ex).__name__}: {str(ex)[:160]} @ {where}"
    if verbose or msg:
        prog = "; ".join(w.log)
        if msg:
            return f"{msg} | seed={seed} shape={w.shape} {np.dtype(w.dtype).name} | {prog[-900:]}"
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=300)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--only", type=int, default=None)
    ap.add_argument("-v", action="store_true")
    ap.add_argument("--dry", action="store_true", help="no GPU: plan, generate and compile only")
    a = ap.parse_args()
    import delayrepay_b200 as dr
    if a.dry:
        from delayrepay_b200 import engine
        engine.dry_run().__enter__()
    else:
        dr.set_device(0)
    seeds = [a.only] if a.only is not None else list(range(a.seed, a.seed + a.n))
    bad = 0
    for s in seeds:
        msg = run_one(s, a.v, a.dry)
        if msg:
            bad += 1
            print(msg, flush=True)
    print(f"fuzz_state: {bad} failing of {len(seeds)}")


if __name__ == "__main__":
    main()
