"""Host logic without a GPU: capture, hash-consing, planning, code generation and NVRTC
compilation of every workload (engine.dry_run: placeholder addresses, launches recorded)."""
import os

import numpy as np
import pytest

import delayrepay_b200 as dr
from delayrepay_b200 import engine, planner
import workloads as wl


@pytest.fixture()
def dry():
    with engine.dry_run() as log:
        del log[:]
        yield log


def test_memoise_leaf_and_expression(dry):
    """reference tests/test.py:143-162 (TestMeta)."""
    arr = np.array([1, 2, 3])
    assert dr.NPArray(arr) is dr.NPArray(arr)
    a = dr.full((3,), 5).astype(np.float32)
    b = dr.full((3,), 3).astype(np.float32)
    assert a is not b
    x = dr.array([1, 2, 3])
    assert np.sin(x) is np.sin(x)
    assert (x + 1) is (x + 1)
    assert (x + 1) is not (x + 1.0)          # 1 / 1.0 / True no longer collide
    assert (x * 2) is not (x * True)


def test_dtype_promotion_follows_numpy(dry):
    f32 = dr.array(np.ones(8, np.float32))
    i64 = dr.array(np.ones(8, np.int64))
    assert (f32 * 0.1).dtype == np.float32            # NEP 50: weak Python scalar
    assert (f32 * np.float64(0.1)).dtype == np.float64
    assert (i64 * 0.5).dtype == np.float64
    assert (i64 / i64).dtype == np.float64
    assert (f32 > 0.5).dtype == np.bool_
    assert np.sqrt(i64).dtype == np.float64
    assert (f32 + i64).dtype == np.float64
    assert np.sum(dr.array(np.ones(4, np.int8))).dtype == np.int64
    assert np.sum(f32).dtype == np.float32


def test_broadcast_shapes_stay_lazy(dry):
    x = dr.array(np.ones(5, np.float32))
    d = x[None, :] - x[:, None]
    assert isinstance(d, dr.DelayArray) and d.shape == (5, 5)
    with pytest.raises(ValueError):
        dr.array(np.ones(4)) + dr.array(np.ones(5))


def test_integer_power_expands_left_associated(dry):
    x = dr.array(np.ones(8))
    p = x ** 3
    assert isinstance(p, dr.BinaryNumpyEx) and p.op == "multiply"
    assert p.children[1] is x and p.children[0].children[0] is x
    assert (x ** 2.5).op == "power"
    assert np.square(x) is x * x


def test_unknown_ufunc_raises_keyerror(dry):
    x = dr.array(np.ones(8))
    with pytest.raises(KeyError):
        np.spacing(x)
    with pytest.raises(KeyError):
        np.argsort(x)


def test_structural_key_ignores_values_and_sizes(dry):
    def plan(n, a):
        x, y = dr.array(np.ones(n)), dr.array(np.ones(n))
        return planner.build_program([a * x + y]).key()
    assert plan(64, 2.0) == plan(4096, -7.5)
    before = engine.stats["compiled"] + engine.stats["disk_hits"]
    for n, a in ((64, 2.0), (4096, -7.5), (1 << 20, 0.1)):
        x, y = dr.array(np.ones(n)), dr.array(np.ones(n))
        (a * x + y).run()
    assert engine.stats["compiled"] + engine.stats["disk_hits"] - before <= 1


def test_shared_subexpressions_are_planned_once(dry):
    x = dr.array(np.ones(64))
    e = x
    for _ in range(18):                 # 2^18 paths through the DAG; reference: 262143 statements
        e = e + e
    prog = planner.build_program([e])
    assert len(prog.instrs) == 18


def test_all_workloads_generate_and_compile(dry):
    i = wl.make_inputs("axpy", 4096)
    wl.axpy(dr, i["a"], dr.array(i["x"]), dr.array(i["y"])).run()
    i = wl.make_inputs("black_scholes", 4096)
    call, put = wl.black_scholes(dr, *(dr.array(i[k]) for k in ("S", "K", "T")))
    n0 = len(dry)
    dr.evaluate(call, put)
    assert len(dry) == n0 + 1, "call and put must share ONE fused kernel"
    kern = dry[-1][0]
    body = kern.source[kern.source.index('extern "C" __global__'):]
    # staged per-warp TMA rings + lockstep body with the guards the interval analysis left
    assert "dr_bulk_load_s(" in body and "dr_elect()" in body and "__syncthreads();          //" not in body
    assert body.count("dr_div4_r<false, false>(") == 2 and "dr_erf4_s<false>(" in body      # generation-3 erf
    assert "dr_log4_t<false>(" in body and "dr_sqrt4_r<false>(" in body and "dr_exp4_t<2>(" in body
    assert "dr_rg.pos4(v0[u].v); dr_rg.pos4(v1[u].v); dr_rg.pos4(v2[u].v);" in body
    i = wl.make_inputs("l2", 4096)
    a, b = dr.array(i["a"]), dr.array(i["b"])
    n0 = len(dry)
    wl.dot(dr, a, b).run()
    assert len(dry) == n0 + 1, "dot = multiply fused into the reduction kernel"
    assert "dr_grid_reduce" in dry[-1][0].source
    wl.l2_distance(dr, a, b).run()
    wl.norm(dr, a).run()
    u = dr.array(wl.make_inputs("heat", 64)["u"])
    wl.heat(dr, u, 2)
    i = wl.make_inputs("nbody", 64)
    wl.nbody_acc(dr, dr.array(i["pos"]), dr.array(i["m"])).run()
    for kern, grid, block in dry:
        assert kern.cubin[:4] == b"\x7fELF"


def test_interval_analysis_places_guards(dry):
    """ranges.analyse: operands tested once at the top, per-operation tests only where the
    propagated interval does not imply the fast form's precondition."""
    from delayrepay_b200 import codegen, ranges
    F = np.float32
    S, K, T, X = (dr.NPArray(dr.DeviceArray.empty((4096,), F)) for _ in range(4))
    call, put = wl.black_scholes(dr, S, K, T)
    prog = planner.build_program([call, put])
    guarded = set(codegen._LANE4_R)
    an = ranges.analyse(prog, "ccc", ranges.scalar_classes(prog), guarded)
    assert an.pos_inputs == [0, 1, 2] and an.any_inputs == []
    ops = {k: prog.instrs[k][0] for k in an.check}
    by_op = {}
    for k, flags in an.check.items():
        by_op.setdefault(ops[k], []).append(flags)
    assert by_op.get("true_divide", by_op.get("divide")) == [(False, False), (False, False)]
    assert by_op["sqrt"] == [(False,)] and by_op["log"] == [(False,)]
    assert by_op["exp"] == [(2,)], "|-r T| < 87 does not follow from T < 2^30 (finite: one max-test)"
    assert by_op["erf"] == [(False,), (False,)], "finite arguments: no nan test"
    # a scalar outside 2^-24 .. 2^24 is not trusted: the second division keeps its tests
    call, put = wl.black_scholes(dr, S, K, T, v=1e-30)
    prog = planner.build_program([call, put])
    scl = ranges.scalar_classes(prog)
    assert "u" in scl
    an = ranges.analyse(prog, "ccc", scl, guarded)
    divs = [f for k, f in an.check.items() if prog.instrs[k][0] in ("true_divide", "divide")]
    assert divs[0] == (False, False) and divs[1] != (False, False)
    # nothing is known about an operand that feeds no guarded operation directly
    prog = planner.build_program([np.exp(X) / (X + 1.0)])
    an = ranges.analyse(prog, "c", ranges.scalar_classes(prog), guarded)
    assert an.pos_inputs == [] and an.any_inputs == []
    assert all(all(f) for f in an.check.values())
    # the abstract domain itself
    a, b = ranges.R(-30, 30, neg=False), ranges.R(-24, 24)
    m = ranges._mul(a, b)
    assert (m.lo, m.hi, m.zero) == (-55, 54, False) and m.neg and m.pos
    s_ = ranges._addsub(ranges.R(-25, 7, zero=True), m, False)
    assert s_.zero and not s_.nz and s_.lo == -55 - 24 and s_.hi == 55
    assert ranges._mul(ranges.R(0, 100), ranges.R(0, 100)) is None, "may overflow: unknown"


def test_plan_cache_skips_planning_on_replay(dry, monkeypatch):
    """A second evaluation of the same expression STRUCTURE (new nodes, new scalar value) must
    not plan again: engine.evaluate_nodes launches the prepared plan."""
    x, y = dr.array(np.ones(1 << 12)), dr.array(np.ones(1 << 12))
    (1.5 * x + y).run()
    calls = []
    real = planner.build_program
    monkeypatch.setattr(planner, "build_program", lambda nodes: calls.append(1) or real(nodes))
    n0 = len(dry)
    (2.5 * x + y).run()
    (1.5 * y + x).run()
    assert calls == [] and len(dry) == n0 + 2
    assert dry[-1][0] is dry[n0 - 1][0], "same kernel object"
    (2.5 * x + x).run()                       # other sharing pattern: one operand -> planned
    assert len(calls) == 1
    (x[1:] * 2.5 + y[1:]).run()               # other layout (unaligned view) -> planned
    assert len(calls) == 2
    monkeypatch.setattr(engine, "_PLAN_CACHE", False)
    (2.5 * x + y).run()
    assert len(calls) == 3


def test_cubin_is_sm100a_with_vector_ldst(dry, tmp_path):
    import subprocess
    x, y = dr.array(np.ones(1 << 12)), dr.array(np.ones(1 << 12))
    (1.5 * x + y).run()
    kern = dry[-1][0]
    p = tmp_path / "k.cubin"
    p.write_bytes(kern.cubin)
    sass = subprocess.run(["cuobjdump", "-sass", str(p)], capture_output=True, text=True).stdout
    assert "sm_100a" in sass or "SM100a" in sass.upper() or "EF_CUDA_SM100" in sass
    assert "LDG.E.128" in sass and "STG.E.128" in sass
    assert "DFMA" not in sass, "fmad=false: a*x+y must not be contracted"


def test_no_cpu_fallback_without_device():
    from delayrepay_b200 import _lib
    if _lib.gpu_available():
        pytest.skip("a GPU is present")
    x = dr.array(np.ones(8))               # host-backed leaf: graph building only
    with pytest.raises(_lib.DrcError):
        (x + 1).get()


# ------------------------------------------------------------------ host logic of this round's kernels
def test_axis_reduction_dispatch_vectors_splits_and_transposed_views():
    """rows / cols selection, 128-bit operand classes, the split reduced axis (two launches, the
    second folds the partials) and the transposed dispatch, observed through dry-run launches."""
    import delayrepay_b200 as dr
    from delayrepay_b200 import engine

    def families(fn):
        n0 = len(engine.dry_log)
        fn().run()
        return [k[0].name.split("_")[1] for k in engine.dry_log[n0:]], engine.dry_log[n0:]

    with engine.dry_run():
        X = dr.array(np.ones((2048, 512), np.float32))
        v0, v1 = dr.array(np.ones(2048, np.float32)), dr.array(np.ones(512, np.float32))
        fam, launches = families(lambda: np.sum(X, axis=0))
        assert len(fam) >= 2 and set(fam) == {"cols"} and launches[0][1][1] > 1, \
            "tall matrix: reduced axis split over grid.y, partials folded by the same family"
        assert "dr_ld<false, float, 4>" in launches[0][0].source, "contiguous operand read with 128-bit loads"
        assert families(lambda: np.sum(X, axis=1))[0] == ["rows"]
        assert families(lambda: X @ v1)[0] == ["rows"]
        assert set(families(lambda: v0 @ X)[0]) == {"cols"}, "v @ X walks X's rows contiguously"
        assert set(families(lambda: np.sum(X.T, axis=1))[0]) == {"cols"}, "transposed view: column kernel"
        fam, launches = families(lambda: np.sum(X[:, 1:], axis=0))
        assert "dr_ld<false, float, 4>" not in launches[0][0].source, "misaligned rows: scalar loads"
        wide = dr.array(np.ones((4, 1 << 16), np.float64))
        assert families(lambda: np.max(wide, axis=0))[0] == ["cols"], "enough columns: no split"


def test_internal_promotions_leave_the_leaf_alone_and_integers_wrap():
    import delayrepay_b200 as dr
    from delayrepay_b200 import engine
    with engine.dry_run():
        x = dr.array(np.arange(12, dtype=np.int32).reshape(3, 4))
        for use in (lambda: np.sum(x), lambda: np.mean(x, axis=0), lambda: np.sum(x, dtype=np.float32),
                    lambda: x @ dr.array(np.ones(4)), lambda: np.cumsum(x, dtype=np.float64)):
            use().run()
            assert x.dtype == np.int32 and x._force().dtype == np.int32
        n0 = len(engine.dry_log)
        ((x * x) * x - x).run()
        src = engine.dry_log[n0][0].source
        assert "(unsigned int)x0 * (unsigned int)x0" in src, "signed arithmetic must be emitted unsigned (wraps)"
        assert x.astype(np.float32) is x and x.dtype == np.float32        # the user-facing astype is in place


def test_transposed_source_detection():
    from delayrepay_b200 import engine, extras
    import delayrepay_b200 as dr
    with engine.dry_run():
        a = dr.array(np.ones((64, 96), np.float32))._force()
        assert extras.transposed_source(a.T) == (1, 64, 96)
        assert extras.transposed_source(a) is None
        assert extras.transposed_source(a[:, 8:72].T) == (1, 64, 64), "column block: pitch > width"
        assert extras.transposed_source(a[::2].T) == (1, 32, 96), "every second row: a larger pitch"
        assert extras.transposed_source(a[:, ::2].T) is None, "strided columns are not a transpose"
        b = dr.array(np.ones((3, 40, 50), np.float64))._force()
        assert extras.transposed_source(b.transpose(0, 2, 1)) == (3, 40, 50)
        assert extras.transposed_source(dr.array(np.ones((8, 8), np.float32))._force().T) is None, "too small"
        assert extras.transposed_source(dr.array(np.ones((64, 64), np.int8))._force().T) is None, "1-byte words"


def test_numpy_surface_registry_and_module_passthrough():
    import delayrepay as dnp
    from delayrepay_b200 import HANDLED_FUNCTIONS
    for fn in (np.sum, np.prod, np.max, np.min, np.mean, np.var, np.std, np.average, np.linalg.norm, np.where,
               np.clip, np.any, np.all, np.count_nonzero, np.argmax, np.argmin, np.ptp, np.trace, np.cumsum,
               np.transpose, np.roll, np.repeat, np.tile, np.diag, np.diagflat, np.reshape, np.ravel, np.squeeze,
               np.expand_dims, np.swapaxes, np.moveaxis, np.broadcast_to, np.concatenate, np.stack, np.vstack,
               np.hstack, np.outer, np.inner, np.vdot, np.diff, np.round, np.around, np.isclose, np.allclose,
               np.array_equal, np.take, np.compress, np.extract, np.nonzero, np.flatnonzero, np.argwhere,
               np.zeros_like, np.ones_like, np.empty_like, np.full_like, np.shape, np.size, np.ndim, np.copy,
               np.matmul):
        assert fn in HANDLED_FUNCTIONS, fn.__name__
    assert dnp.floor is np.floor and dnp.inf == np.inf and dnp.int8 is np.int8 and dnp.linalg is np.linalg
    with pytest.raises(AttributeError):
        dnp.definitely_not_a_numpy_name
    x = dnp.array(np.ones(4))
    with pytest.raises(KeyError):               # no device implementation: loud, never a host fallback
        np.sort(x)


def test_numpy_surface_shapes_and_dtypes_match_numpy_without_a_device():
    """tools/fuzz_shapes.py: every handler of the NumPy surface with random shapes, dtypes, axes
    and keywords, planned and compiled in dry-run mode; result shapes and dtypes must be NumPy's
    (values are the GPU tests' job).  Found: a row mask on an (n, 1) array dropped the unit axis."""
    import subprocess
    import sys
    tool = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "fuzz_shapes.py")
    out = subprocess.run([sys.executable, tool, "--n", "150"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    assert out.stdout.strip().splitlines()[-1].startswith("fuzz_shapes: 0 failing"), out.stdout[-3000:]


def test_tile_family_for_transposed_operands_and_generation2_scans():
    """Planning of the section 8(f) kernels added in round 2 (compile-only): a transposed operand
    beside row-major ones selects the shared-memory tile family (two elements per thread when
    every access is 8-byte aligned), in-place and reduced regions keep the strided kernel; 1-d
    cumsum takes the one-pass chained scan, row scans the vectorised kernel."""
    import delayrepay_b200 as dr
    from delayrepay_b200 import engine, planner

    def launches(fn):
        n0 = len(engine.dry_log)
        r = fn()
        if hasattr(r, "run"):
            r.run()
        return engine.dry_log[n0:]

    with engine.dry_run():
        X = dr.array(np.ones((256, 320), np.float32))
        Y = dr.array(np.ones((320, 256), np.float32))
        (k, grid, threads), = launches(lambda: (Y.T + X) * 2.0)
        assert k.name.startswith("dr_tile_") and threads == 256 and grid == 4 * 5
        assert "tile0[64][65]" in k.source and "dr_ld<true, float, 2>" in k.source, "64 x 64 tile, pairs"
        (k, _, _), = launches(lambda: Y.T[:, 1:] + X[:, 1:])
        assert k.name.startswith("dr_tile_") and "dr_ld<true, float, 2>" not in k.source, "misaligned rows: W = 1"
        D = dr.array(np.ones((100, 70)))
        E = dr.array(np.ones((70, 100)))
        v = dr.array(np.ones(70))
        (k, _, _), = launches(lambda: np.exp(E.T) - D * v)
        assert k.name.startswith("dr_tile_") and "tile0[32][33]" in k.source, "8-byte words: 32 x 32 tile"
        (k, _, _), = launches(lambda: Y.T.copy())
        assert k.name.startswith("dr_tile_"), "a 2-d transpose copy is a one-operand tile region"
        small = dr.array(np.ones((16, 320), np.float32))
        assert not launches(lambda: small.T + dr.array(np.ones((320, 16), np.float32)))[0][0].name.startswith("dr_tile_")
        assert all(not k.name.startswith("dr_tile_") for k, _, _ in launches(lambda: np.sum(Y.T + X)))
        # the class table itself
        prog = planner.build_program([Y.T + X])
        from delayrepay_b200.device import DeviceArray
        out = DeviceArray.empty((256, 320), np.float32)
        lay = planner.resolve_layout(prog, [out])
        assert lay.family == "nd" and planner.tile_classes(prog, [out], lay) == (("t", "v"), 64, 2)

        x = dr.array(np.ones(1 << 23, np.float32))
        names = [k.name for k, _, _ in launches(lambda: np.cumsum(x))]
        assert len(names) == 1 and names[0].endswith("_chain"), names
        (k, grid, threads), = launches(lambda: np.cumsum(x))
        assert (grid, threads) == (148 * 2, 512), "two co-resident CTAs per SM"
        assert "st.relaxed.gpu.global.b128" in k.source, "totals travel as single 16-byte records"
        M = dr.array(np.ones((256, 2048), np.float32))
        assert [k.name.split("_")[-1] for k, _, _ in launches(lambda: np.cumsum(M, axis=1))] == ["rowscan2"]
        odd = dr.array(np.ones((256, 2050), np.float32))
        assert [k.name.split("_")[-1] for k, _, _ in launches(lambda: np.cumsum(odd, axis=1))] == ["rowscan"], \
            "rows that are not a multiple of 16 bytes keep the scalar row scan"
        short = dr.array(np.ones(5000, np.float32))
        assert all(not k.name.endswith("_chain") for k, _, _ in launches(lambda: np.cumsum(short)))


def test_emission_rules_found_by_the_late_fuzz_batches():
    """Compile-only guards for three round-2 emission rules: integer negation is a subtraction from
    a zero ptxas cannot fold (it loses a literal negation inside a fused VIMNMX3), np.maximum /
    np.minimum put their second operand first in dr_max / dr_min (NumPy returns the second operand
    when the two compare equal), np.clip chooses the operand order by the kind of its bounds."""
    import delayrepay_b200 as dr
    from delayrepay_b200 import engine

    def source(fn):
        n0 = len(engine.dry_log)
        fn().run()
        return engine.dry_log[n0][0].source

    with engine.dry_run():
        xi = dr.array(np.arange(64, dtype=np.int32))
        xu = dr.array(np.arange(64, dtype=np.uint16))
        a = dr.array(np.ones(64, np.float64))
        b = dr.array(np.ones(64, np.float64))
        assert "(unsigned int)(gridDim.z - 1u) - (unsigned int)x0" in source(lambda: np.minimum(-xi, xi))
        assert "(gridDim.z - 1u) - x0" in source(lambda: -xu + xu)
        assert "dr_max<double>(x1, x0)" in source(lambda: np.maximum(a, b)), "second operand first"
        assert "dr_min<double>(x1, x0)" in source(lambda: np.minimum(a, b))
        s = source(lambda: np.clip(a, 0.0, 1.0))
        assert "dr_max<double>(x0, s0)" in s and "dr_min<double>(t0, s1)" in s, "two scalar bounds keep `a` on equality"
        s = source(lambda: np.clip(a, b, None))
        assert "dr_max<double>(x1, x0)" in s, "an array bound wins on equality"
        with pytest.raises(ValueError):
            engine.launch(engine._kernels[next(iter(engine._kernels))], 0, (1, 1, 2), 32, engine.Args())
