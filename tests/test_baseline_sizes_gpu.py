"""Parity AT the BASELINE sizes (operands >= 4 GiB: element offsets past 2^30, byte offsets past
2^32, the streaming-hint variants, persistent loops with thousands of tiles per CTA).

The oracle cannot evaluate 2^30 elements in seconds, so the device operands repeat a seeded
chunk: the result must (i) repeat with the same period -- compared on the device over ALL
positions, bitwise -- and (ii) equal the oracle on the chunk, under the same bars as
tests/test_parity_gpu.py.  bench.py applies the same checks to what it times.
"""
import numpy as np
import pytest

import workloads as wl
from oracle import refcpu
from util import assert_bits_equal

pytestmark = pytest.mark.gpu


def _val(x):
    return float(x.get()) if hasattr(x, "get") else float(x)


def _tiled(dr, host, reps):
    return dr.tile(dr.array(host), reps)


def _periodic(np_mod, got, period):
    blocks = got.reshape(got.size // period, period)
    return bool(np_mod.all(np_mod.equal(blocks, blocks[0:1])))


def test_flat_light_body_f32_5gib_operand(gpu):
    """flat family, vectorised streaming variant; 1.25 * 2^30 float32 per operand = 5 GiB, so the
    last quarter lives at byte offsets past 2^32."""
    dr = gpu
    chunk, reps = 1 << 22, 320
    rng = np.random.default_rng(11)
    hx, hy = rng.standard_normal(chunk).astype(np.float32), rng.standard_normal(chunk).astype(np.float32)
    x, y = _tiled(dr, hx, reps), _tiled(dr, hy, reps)
    r = (x * 2.5 + y) / (y * y + 1.0)
    assert r.size * 4 > (1 << 32)
    want = (hx * np.float32(2.5) + hy) / (hy * hy + np.float32(1.0))
    assert_bits_equal(r[:chunk].get(), want, "first block")
    assert_bits_equal(r[-chunk:].get(), want, "last block (past 4 GiB)")
    assert _periodic(np, r, chunk)
    # a view that STARTS past 4 GiB as an operand
    tail = x[(1 << 30) + 5:(1 << 30) + 5 + (1 << 20)] + 1.0
    assert_bits_equal(tail.get(), np.resize(hx, x.size)[(1 << 30) + 5:(1 << 30) + 5 + (1 << 20)] + np.float32(1.0), "offset view")


def test_black_scholes_2pow30_staged_kernel(gpu):
    """The headline configuration itself: 2^30 options, staged (TMA ring) heavy-body kernel with
    the streaming hints, 3 x 4 GiB in, 2 x 4 GiB out."""
    import bench
    dr = gpu
    chunk, n = 1 << 22, 1 << 30
    host = wl.make_inputs("black_scholes", chunk)
    S, K, T = (_tiled(dr, host[k], n // chunk) for k in ("S", "K", "T"))
    call, put = wl.black_scholes(dr, S, K, T)
    dr.evaluate(call, put)
    ver = bench.Verifier()
    bench.verify_black_scholes(ver, dr, wl, host, call, put, n, chunk)
    res = ver.results["black_scholes_f32"]
    print("   black-scholes 2^30:", res)
    assert res["ok"], res


def test_full_reductions_f64_2pow30(gpu):
    dr = gpu
    chunk, n = 1 << 22, 1 << 30
    i = wl.make_inputs("l2", chunk)
    a, b = _tiled(dr, i["a"], n // chunk), _tiled(dr, i["b"], n // chunk)
    ra, rb = refcpu.leaf(i["a"]), refcpu.leaf(i["b"])
    reps = n // chunk
    got = float(wl.l2_distance(dr, a, b))
    want = np.sqrt(reps) * _val(wl.l2_distance(refcpu, ra, rb))
    assert abs(got - want) <= 1e-12 * want
    got = float(wl.dot(dr, a, b))
    want = reps * _val(wl.dot(refcpu, ra, rb))
    assert abs(got - want) <= 1e-12 * abs(want)
    # max / argmax land on the last occurrence-independent value; the index must be the FIRST one
    assert float(np.max(a)) == float(i["a"].max())
    assert int(np.argmax(a)) == int(i["a"].argmax())


def test_heat_stencil_5gib_grid(gpu):
    """stencil family on a 40960 x 32768 float32 grid (5 GiB, the last rows past 2^32 bytes):
    3 steps, bit-exact against the oracle on blocks at both corners, at the 4 GiB line and in
    the centre."""
    dr = gpu
    rows, cols, blk, steps = 40960, 32768, 2048, 3
    h0 = wl.make_inputs("heat", blk)["u"]
    u = dr.tile(dr.array(h0), (rows // blk, cols // blk))

    def u0(r0, r1, c0, c1):
        return h0[np.ix_(np.arange(r0, r1) % blk, np.arange(c0, c1) % blk)]

    def oracle_block(r0, r1, c0, c1):
        R0, R1, C0, C1 = max(r0 - steps, 0), min(r1 + steps, rows), max(c0 - steps, 0), min(c1 + steps, cols)
        blk_ = refcpu.leaf(u0(R0, R1, C0, C1).copy())
        wl.heat(refcpu, blk_, steps)
        return blk_.get()[r0 - R0:r1 - R0, c0 - C0:c1 - C0]
    wl.heat(dr, u, steps)
    line = (1 << 32) // (cols * 4)            # the row whose first byte is at offset 2^32
    for (r0, r1, c0, c1) in [(0, 64, 0, 300), (rows - 64, rows, cols - 300, cols),
                             (line - 32, line + 32, 1000, 1300), (rows // 2, rows // 2 + 64, cols // 2, cols // 2 + 300)]:
        assert_bits_equal(u[r0:r1, c0:c1].get(), oracle_block(r0, r1, c0, c1), f"block {(r0, r1, c0, c1)}")
