"""The leading-axis sharding layer (delayrepay_b200/sharding.py) against the oracle.

In-process meshes whose ranks share ONE device run the whole protocol -- partitioning,
localisation, the halo-pushing stencil kernel with its release/acquire flags, the partial
reductions and their combination -- on a single GPU, so these tests run on the 1-GPU test box.
The SPMD tests (one process per GPU: CUDA IPC mappings, TCP rendezvous, libdrcuda's NCCL
communicator) need two devices and skip otherwise; bench.py exercises the same path at
WORLD_SIZE > 1 and verifies its results.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

import workloads as wl
from oracle import refcpu
from util import assert_bits_equal

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture()
def mesh3(gpu):
    from delayrepay_b200 import sharding
    m = sharding.init(devices=[0, 0, 0])
    yield m
    sharding.shutdown()


def _oracle_heat(h, steps, before=None):
    u = refcpu.leaf(h.copy())
    if before is not None:
        before(u)
    wl.heat(refcpu, u, steps)
    return u.get()


@pytest.mark.parametrize("shape,steps", [((403, 512), 40), ((96, 1024), 7), ((1000, 260), 25)])
def test_sharded_heat_bit_exact_one_gpu_three_ranks(gpu, mesh3, shape, steps):
    """C4 through the drop-in API on a sharded array: every step is ONE stencil launch per block,
    halo rows are pushed by the kernel.  Bit-exact against the oracle (Jacobi semantics)."""
    from delayrepay_b200 import engine
    dr = gpu
    rng = np.random.default_rng(4)
    h = rng.random(shape, dtype=np.float32)
    u = dr.shard(h)
    assert u.array.base.H == 1 and u.shape == shape
    l0 = engine.stats["launches"]
    wl.heat(dr, u, steps)
    assert engine.stats["launches"] - l0 == 3 * steps, "one launch per block per step, nothing else"
    assert all(l.epoch == steps and not l.dirty for l in u.array.base.links.values())
    assert_bits_equal(u.get(), _oracle_heat(h, steps), f"heat {shape} x{steps}")


def test_sharded_heat_with_boundary_writes_and_fallback_exchange(gpu, mesh3):
    """Writes that are not the stencil kernel (boundary conditions) leave the neighbours' halo
    copies stale: the next stencil step starts with the peer-copy exchange."""
    dr = gpu
    rng = np.random.default_rng(5)
    h = rng.random((300, 256), dtype=np.float32)

    def bc(u):
        u[0, :] = 1.0
        u[:, -1] = 0.5
        u[100:200, 3:9] = 0.25          # crosses the block boundary rows 100 / 200
    u = dr.shard(h)
    want = refcpu.leaf(h.copy())
    for _ in range(3):
        bc(u)
        bc(want)
        wl.heat(dr, u, 5)
        wl.heat(refcpu, want, 5)
    assert_bits_equal(u.get(), want.get(), "heat with boundary writes")


def test_other_consumers_of_pushed_halo_rows(gpu, mesh3):
    """Halo rows pushed by the neighbours' stencil kernels are also read by kernels that are not
    the halo stencil (a difference of shifted views, a reduction over them, a stencil into ANOTHER
    array): those are preceded by a device-side wait on the neighbours' flags."""
    from delayrepay_b200 import engine
    dr = gpu
    rng = np.random.default_rng(21)
    h = rng.random((300, 256), dtype=np.float32)
    u, want = dr.shard(h), h.copy()
    v = dr.shard(np.zeros_like(h))
    for _ in range(4):
        wl.heat(dr, u, 3)
        for _ in range(3):
            want[1:-1, 1:-1] = want[1:-1, 1:-1] + np.float32(0.1) * (
                want[2:, 1:-1] + want[:-2, 1:-1] + want[1:-1, 2:] + want[1:-1, :-2] - np.float32(4.0) * want[1:-1, 1:-1])
        l0 = engine.stats["launches"]
        grad = u[2:, 1:-1] - u[:-2, 1:-1]                    # reads one row into each halo
        assert_bits_equal(grad.get(), want[2:, 1:-1] - want[:-2, 1:-1], "vertical difference")
        assert engine.stats["launches"] - l0 >= 3 + 2, "a wait kernel per block with a neighbour"
        got = float(np.sum(np.abs(u[2:, :] - u[:-2, :])))
        ref = float(np.sum(np.abs(want[2:, :] - want[:-2, :]).astype(np.float64)))
        assert abs(got - ref) <= 1e-5 * ref
        v[1:-1, :] = u[2:, :] + u[:-2, :]
        assert_bits_equal(v.get()[1:-1], want[2:, :] + want[:-2, :], "stencil into another array")


def test_sharded_heat_not_stencil_eligible_shape(gpu, mesh3):
    """203 columns (not a multiple of the vector width): the blocks take the generic
    temporary + copy path and the explicit exchange every step -- slow, but exact."""
    dr = gpu
    rng = np.random.default_rng(6)
    h = rng.random((150, 203), dtype=np.float32)
    u = dr.shard(h)
    wl.heat(dr, u, 6)
    assert_bits_equal(u.get(), _oracle_heat(h, 6), "heat 150x203")


def test_sharded_wide_stencil_halo2(gpu, mesh3):
    dr = gpu
    rng = np.random.default_rng(7)
    h = rng.random((240, 512), dtype=np.float64)

    def step(u):
        u[2:-2, 2:-2] = 0.2 * (u[4:, 2:-2] + u[:-4, 2:-2] + u[2:-2, 4:] + u[2:-2, :-4] + u[2:-2, 2:-2])
    u = dr.shard(h, halo=2)
    want = refcpu.leaf(h.copy())
    for _ in range(9):
        step(u)
        step(want)
    assert_bits_equal(u.get(), want.get(), "5-point stencil with 2-row reach, float64")
    with pytest.raises((ValueError, NotImplementedError)):      # reads 2 rows away, holds 1
        step(dr.shard(h, halo=1))


def test_sharded_reductions_and_elementwise(gpu, mesh3):
    dr = gpu
    i = wl.make_inputs("l2", (1 << 18) + 5)
    a, b = dr.shard(i["a"]), dr.shard(i["b"])
    ra, rb = refcpu.leaf(i["a"]), refcpu.leaf(i["b"])

    def val(x):
        return float(x.get()) if hasattr(x, "get") else float(x)
    for name, fn in (("l2", wl.l2_distance), ("dot", wl.dot)):
        got, want = float(fn(dr, a, b)), val(fn(refcpu, ra, rb))
        assert abs(got - want) <= 1e-12 * abs(want), (name, got, want)
    got, want = float(wl.norm(dr, a)), val(wl.norm(refcpu, ra))
    assert abs(got - want) <= 1e-12 * abs(want)
    assert float(np.max(a)) == i["a"].max() and float(np.min(a - b)) == (i["a"] - i["b"]).min()
    assert abs(float(np.mean(a)) - i["a"].mean()) <= 1e-12
    e = wl.axpy(dr, 1.5, a, b)
    assert type(e._force()).__name__ == "ShardView"
    assert_bits_equal(e.get(), 1.5 * i["a"] + i["b"], "sharded axpy")
    # a replicated (ordinary) operand of the global shape is row-sliced per rank
    assert_bits_equal((a + dr.array(i["b"])).get(), i["a"] + i["b"], "sharded + replicated")
    assert_bits_equal((a * i["b"]).get(), i["a"] * i["b"], "sharded * host array")
    m = np.arange(12.0 * 7).reshape(12, 7)
    s = dr.shard(m, halo=0)
    np.testing.assert_array_equal(np.sum(s, axis=1).get(), m.sum(1))
    np.testing.assert_array_equal(np.sum(s, axis=0).get(), m.sum(0))
    np.testing.assert_array_equal((s + np.arange(7.0)).get(), m + np.arange(7.0))
    np.testing.assert_array_equal(s[3:9, 1:5].get(), m[3:9, 1:5])
    np.testing.assert_array_equal(s[5].get(), m[5])


def test_sharded_numpy_surface(gpu, mesh3):
    """The NumPy calls a script makes around the hot path, on sharded operands, against NumPy."""
    dr = gpu
    rng = np.random.default_rng(12)
    m, v = rng.standard_normal((50, 12)), rng.standard_normal(50)
    s, sv = dr.shard(m, halo=0), dr.shard(v, halo=0)
    w = np.arange(12.0)
    cases = {
        "where": (lambda x, y: np.where(x > 0, x, -x), True),
        "compare": (lambda x, y: x >= 0.5, True),
        # (on a LEAF astype converts in place, like the reference's delayarray.py:401-408; a lazy
        # node casts)
        "astype": (lambda x, y: (x + 0).astype(np.float32) * 2, True),
        # x ** 3 is the reference's multiply chain (x*x)*x (delayarray.py:316-324), not np.power
        "pow3": (lambda x, y: (x * x) * x if isinstance(x, np.ndarray) else x ** 3, True),
        "mean0": (lambda x, y: np.mean(x, axis=0), False),
        "max0": (lambda x, y: np.max(x, axis=0), True),
        "sum1_keepdims": (lambda x, y: np.sum(x, axis=1, keepdims=True), False),
        "sum_keepdims": (lambda x, y: np.sum(x, keepdims=True), False),
        "column_broadcast": (lambda x, y: x * y[:, None], True),
        "neg": (lambda x, y: -x + 1, True),
        "sqrt_abs": (lambda x, y: np.sqrt(np.abs(x)), True),
        "min": (lambda x, y: np.min(x), True),
        "matvec": (lambda x, y: x @ w, False),
        "norm": (lambda x, y: np.linalg.norm(y), False),
        "var": (lambda x, y: np.var(y), False),
        "std": (lambda x, y: np.std(x), False),
        "any": (lambda x, y: np.any(x > 3), True),
        "count_nonzero": (lambda x, y: np.count_nonzero(x > 0), True),
    }
    for name, (fn, exact) in cases.items():
        got, want = np.asarray(fn(s, sv).get()), np.asarray(fn(m, v))
        assert got.shape == want.shape and got.dtype == want.dtype, (name, got.shape, want.shape, got.dtype, want.dtype)
        if exact:
            assert_bits_equal(got, want, name)
        else:
            np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-13, err_msg=name)
    s += 1.0                                   # in place: the leaf keeps its blocks
    assert type(s).__name__ == "NPArray"
    assert_bits_equal(s.get(), m + 1.0, "in-place add")
    a, b = s * 2, s + 3
    dr.evaluate(a, b)
    assert_bits_equal(a.get(), (m + 1.0) * 2, "evaluate a")
    assert_bits_equal(b.get(), (m + 1.0) + 3, "evaluate b")
    assert_bits_equal(s.copy().get(), m + 1.0, "copy")
    assert len(s) == 50 and s.ndim == 2 and s.size == 600
    with pytest.raises(NotImplementedError):
        (s[3:] + s[:-3]).get()                 # rows 3 away are not resident (halo 0)
    with pytest.raises(AttributeError):
        np.cumsum(s)                           # not a sharded operation: a clear error, no silent gather


def test_sharded_black_scholes_is_one_kernel_per_block(gpu, mesh3):
    """C2 through the layer: call and put are co-evaluated per row block (dr.evaluate ->
    sharding.run_many), bit-identical to the unsharded engine, no communication."""
    from delayrepay_b200 import engine
    dr = gpu
    n = (1 << 16) + 12
    h = wl.make_inputs("black_scholes", n)
    ref = wl.black_scholes(dr, *(dr.array(h[k]) for k in ("S", "K", "T")))
    dr.evaluate(*ref)
    S, K, T = (dr.shard(h[k]) for k in ("S", "K", "T"))
    l0 = engine.stats["launches"]
    call, put = wl.black_scholes(dr, S, K, T)
    dr.evaluate(call, put)
    assert engine.stats["launches"] - l0 == 3, "one two-output kernel per block"
    assert type(call._force()).__name__ == "ShardView"
    names = {k.name for k in engine._kernels.values() if k.name == engine.last_kernel_name()}
    assert any("dr_bulk_load_s(" in k.source for k in engine._kernels.values() if k.name in names), \
        "the blocks run the vectorised staged kernel"
    # same kernel as unsharded; only the <= 3 tail elements of each block take the scalar path
    for got, want in ((call.get(), ref[0].get()), (put.get(), ref[1].get())):
        same = np.mean(got.view(np.uint32) == want.view(np.uint32))
        assert same >= 0.9998, same
        bar = 16 * np.finfo(np.float32).eps * np.maximum(h["S"], h["K"])
        assert np.all(np.abs(got.astype(np.float64) - want) <= bar)


def test_sharded_1d_stencil_generic_path(gpu, mesh3):
    """1-d arrays shard with a halo on request; their stencils take the generic temporary + copy
    path with the peer-copy exchange before each step."""
    dr = gpu
    h = np.random.default_rng(3).standard_normal(1000)
    v, want = dr.shard(h, halo=1), h.copy()
    assert v.array.base.H == 1 and dr.shard(h).array.base.H == 0
    for _ in range(4):
        v[1:-1] = 0.25 * (v[2:] + v[:-2]) + 0.5 * v[1:-1]
        want[1:-1] = 0.25 * (want[2:] + want[:-2]) + 0.5 * want[1:-1]
    assert_bits_equal(v.get(), want, "1-d three-point stencil")


def test_sharded_nbody_matches_unsharded(gpu, mesh3):
    """C5 on a sharded ``pos``: rows of W are sharded, pos / m are gathered once."""
    dr = gpu
    from delayrepay_b200 import engine
    i = wl.make_inputs("nbody", 1536)
    acc = wl.nbody_acc(dr, dr.shard(i["pos"], halo=0), dr.array(i["m"]))
    got = acc.get()
    _check_nbody(i, got)
    # W @ pos and W.sum(1) share ONE pass over the all-pairs producer per block
    names = [k.name for k in engine._kernels.values()]
    assert any(n.startswith("dr_mm_skinny_") for n in names)


def _check_nbody(i, got):
    """|error| <= 1e-5 of the term scale sum|w||pos| + |pos| sum|w| against a float64 evaluation
    (the bar of tests/test_parity_gpu.py: acc is a difference of two large sums)."""
    p, m = i["pos"].astype(np.float64), i["m"].astype(np.float64)
    d = p[None, :, :] - p[:, None, :]
    w = m[None, :] * ((d ** 2).sum(-1) + 1e-3) ** -1.5
    want = w @ p - p * w.sum(1)[:, None]
    scale = np.abs(w) @ np.abs(p) + np.abs(p) * np.abs(w).sum(1)[:, None]
    assert np.max(np.abs(got - want) / scale) <= 1e-5


def test_inprocess_mesh_over_two_devices(gpu):
    """One process driving two GPUs: peer access between the blocks, ncclCommInitAll communicators,
    grouped all-reduce, peer-copy gather."""
    from delayrepay_b200 import _lib, sharding
    if _lib.init() < 2:
        pytest.skip("needs two GPUs")
    dr = gpu
    mesh = sharding.init(devices=[0, 1])
    try:
        assert len(mesh.comms) == 2
        rng = np.random.default_rng(9)
        h = rng.random((300, 512), dtype=np.float32)
        u = dr.shard(h)
        wl.heat(dr, u, 30)
        u[0, :] = 2.0
        wl.heat(dr, u, 3)
        want = refcpu.leaf(h.copy())
        wl.heat(refcpu, want, 30)
        want[0, :] = 2.0
        wl.heat(refcpu, want, 3)
        assert_bits_equal(u.get(), want.get(), "in-process 2-device heat")
        i = wl.make_inputs("l2", (1 << 18) + 1)
        a, b = dr.shard(i["a"]), dr.shard(i["b"])
        got, ref = float(wl.l2_distance(dr, a, b)), float(np.sqrt(np.sum((i["a"] - i["b"]) ** 2)))
        assert abs(got - ref) <= 1e-12 * ref
        assert_bits_equal((a * 2 + b).get(), i["a"] * 2 + i["b"], "in-process 2-device axpy")
        j = wl.make_inputs("nbody", 1024)
        _check_nbody(j, wl.nbody_acc(dr, dr.shard(j["pos"], halo=0), dr.array(j["m"])).get())
    finally:
        sharding.shutdown()
        dr.set_device(0)


SPMD = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np
import delayrepay_b200 as dr
from delayrepay_b200 import sharding, engine
import workloads as wl
from oracle import refcpu
rank = int(os.environ["RANK"])
dr.set_device(int(os.environ["LOCAL_RANK"]))
mesh = sharding.init()
assert mesh.spmd and mesh.world == 2
rng = np.random.default_rng(4)
h = rng.random((515, 768), dtype=np.float32)
u = dr.shard(h)
l0 = engine.stats["launches"]
wl.heat(dr, u, 60)
assert engine.stats["launches"] - l0 == 60
want = refcpu.leaf(h.copy()); wl.heat(refcpu, want, 60)
assert u.get().tobytes() == want.get().tobytes(), "sharded heat differs"
g = (u[2:, 1:-1] - u[:-2, 1:-1]).get()                     # non-stencil consumer of pushed halo rows
wn = want.get()
assert g.tobytes() == (wn[2:, 1:-1] - wn[:-2, 1:-1]).tobytes(), "difference of shifted views differs"
u[0, :] = 1.0; want[0, :] = 1.0
wl.heat(dr, u, 3); wl.heat(refcpu, want, 3)
assert u.get().tobytes() == want.get().tobytes(), "sharded heat after a boundary write differs"
i = wl.make_inputs("l2", (1 << 20) + 3)
a, b = dr.shard(i["a"]), dr.shard(i["b"])
got = float(wl.l2_distance(dr, a, b)); ref = float(np.sqrt(np.sum((i["a"] - i["b"]) ** 2)))
assert abs(got - ref) <= 1e-12 * ref, (got, ref)
got = float(wl.dot(dr, a, b)); ref = float(np.dot(i["a"], i["b"]))
assert abs(got - ref) <= 1e-12 * abs(ref), (got, ref)
assert (a * 2 + b).get().tobytes() == (i["a"] * 2 + i["b"]).tobytes()
j = wl.make_inputs("nbody", 1024)
got = wl.nbody_acc(dr, dr.shard(j["pos"], halo=0), dr.array(j["m"])).get()
p64, m64 = j["pos"].astype(np.float64), j["m"].astype(np.float64)
dd = p64[None, :, :] - p64[:, None, :]
w64 = m64[None, :] * ((dd ** 2).sum(-1) + 1e-3) ** -1.5
want64 = w64 @ p64 - p64 * w64.sum(1)[:, None]
scale = np.abs(w64) @ np.abs(p64) + np.abs(p64) * np.abs(w64).sum(1)[:, None]
assert np.max(np.abs(got - want64) / scale) <= 1e-5
mesh.barrier()
sharding.shutdown()
print("SPMD-OK", rank, flush=True)
"""


def test_spmd_two_processes_two_gpus(gpu, tmp_path):
    from delayrepay_b200 import _lib
    if _lib.init() < 2:
        pytest.skip("needs two GPUs")
    script = tmp_path / "spmd.py"
    script.write_text(SPMD.format(root=ROOT))
    port = 29000 + os.getpid() % 2000
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), PYTHONPATH=ROOT + os.pathsep + os.path.join(ROOT, "tests"))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = []
    for p in procs:
        try:
            outs.append(p.communicate(timeout=300)[0])
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"SPMD-OK {r}" in o, o[-3000:]
