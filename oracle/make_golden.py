"""Generate tests/golden/*.npz by running the REAL reference (CPU backend) in this container.

    python oracle/make_golden.py            # write fixtures
    python oracle/make_golden.py --check    # also run oracle/refcpu.py side by side (bitwise)

The reference is imported from /root/reference with DELAY_CPU=1
(/root/reference/delayrepay/backend.py:3-5 selects cpu.py).  It cannot travel to the GPU
box, so its outputs on small seeded inputs are committed as fixtures together with this
script.  ORACLE / test infrastructure only.
"""
import os
import sys

os.environ["DELAY_CPU"] = "1"
os.environ.pop("DELAY_LIFT", None)
sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, "/root/reference")
sys.path.insert(1, ROOT)

import contextlib
import io

import numpy as np

with contextlib.redirect_stdout(io.StringIO()):       # backend.py:4 prints on import
    import delayrepay as ref
    from delayrepay.delayarray import reset as ref_reset

assert ref.__file__.startswith("/root/reference"), ref.__file__

import workloads as wl          # noqa: E402
from oracle import refcpu                            # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def run_all(xp, wrap, n=2048, grid=64, bodies=128, steps=5):
    """Evaluate every workload through module ``xp``; ``wrap`` turns ndarray -> xp array."""
    res = {}
    i = wl.make_inputs("axpy", n)
    res["axpy"] = np.asarray(wl.axpy(xp, i["a"], wrap(i["x"]), wrap(i["y"])).get())

    i = wl.make_inputs("black_scholes", n)
    call, put = wl.black_scholes(xp, wrap(i["S"]), wrap(i["K"]), wrap(i["T"]))
    res["bs_call"], res["bs_put"] = np.asarray(call.get()), np.asarray(put.get())

    i = wl.make_inputs("l2", n)
    res["l2"] = np.asarray(wl.l2_distance(xp, wrap(i["a"]), wrap(i["b"])))
    res["dot"] = np.asarray(wl.dot(xp, wrap(i["a"]), wrap(i["b"])).get())
    res["norm"] = np.asarray(np.sqrt(wl.dot(xp, wrap(i["a"]), wrap(i["a"])).get()))

    i = wl.make_inputs("heat", grid)
    u = wrap(i["u"].copy())
    res["heat"] = np.asarray(wl.heat(xp, u, steps).get()).copy()

    i = wl.make_inputs("nbody", bodies)
    acc = wl.nbody_acc(xp, wrap(i["pos"]), wrap(i["m"]))
    res["nbody"] = np.asarray(acc.get() if hasattr(acc, "get") else acc)

    # association order of the integer-power expansion (delayarray.py:316-324)
    rng = np.random.default_rng(6)
    for dt in (np.float32, np.float64):
        x = rng.standard_normal(n).astype(dt)
        res[f"pow3_{np.dtype(dt).name}"] = np.asarray((wrap(x) ** 3).get())
        res[f"pow5_{np.dtype(dt).name}"] = np.asarray((wrap(x) ** 5).get())
        res[f"fuse_{np.dtype(dt).name}"] = np.asarray(
            (np.sin(wrap(x)) ** 2 + np.cos(wrap(x)) ** 2).get())
        res[f"chain_{np.dtype(dt).name}"] = np.asarray(
            (np.tanh(wrap(x)) * np.arctan2(wrap(x), wrap(x) + 2) / (np.abs(wrap(x)) + 1)).get())
    return res


def main():
    check = "--check" in sys.argv
    os.makedirs(OUT, exist_ok=True)
    ref_reset()
    golden = run_all(ref, ref.NPArray)
    np.savez_compressed(os.path.join(OUT, "workloads.npz"), **golden)
    print("wrote", os.path.join(OUT, "workloads.npz"), {k: v.shape for k, v in golden.items()})
    if check:
        for kw in (dict(), dict(n=4099, grid=97, bodies=67, steps=3)):
            ref_reset()
            a = run_all(ref, ref.NPArray, **kw)
            b = run_all(refcpu, refcpu.leaf, **kw)
            for k in a:
                same = a[k].shape == b[k].shape and a[k].dtype == b[k].dtype and \
                    a[k].tobytes() == b[k].tobytes()
                print(f"  {k:14s} {str(a[k].dtype):8s} {'bit-identical' if same else 'MISMATCH'}")
                assert same, k
        print("oracle/refcpu.py == reference CPU backend, bit for bit")


if __name__ == "__main__":
    main()
