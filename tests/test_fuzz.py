"""Differential fuzzing (tools/fuzz_diff.py): random NumPy programs -- mixed dtypes, Python and
NumPy scalars, broadcast / offset / strided / reversed / transposed operands, `where`,
comparisons, integer powers, one transcendental or one reduction at the root -- evaluated by the
engine and by plain NumPy with the reference's expansion rules (its CPU backend is NumPy per node,
cpu.py:13-31; integer powers and `square` are multiply chains, delayarray.py:316-336).

Bars as everywhere else: arithmetic bit-exact, transcendentals <= 2 ulp, reductions rtol 1e-12 /
1e-5 of the sum of magnitudes, integer and boolean results exact, shapes and dtypes identical.
The CPU test plans, generates and NVRTC-compiles the same programs without a device.
"""
import os
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import fuzz_diff  # noqa: E402


def test_random_programs_plan_and_compile_without_a_device():
    from delayrepay_b200 import engine
    with engine.dry_run():
        bad = [m for m in (fuzz_diff.run_one(s, dry=True) for s in range(40)) if m]
    assert not bad, "\n".join(bad)


@pytest.mark.gpu
@pytest.mark.parametrize("first", [0, 1000])          # basic operation set / extended set (seeds >= 1000)
def test_random_programs_match_numpy(gpu, first):
    bad = []
    for s in range(first, first + 300):
        m = fuzz_diff.run_one(s)
        if m:
            bad.append(m)
    assert not bad, "\n".join(bad)
