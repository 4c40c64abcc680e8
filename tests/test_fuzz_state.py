"""Stateful differential fuzzing (tools/fuzz_state.py): random programs of slice / masked
assignments (right-hand sides that read shifted views of the target included), in-place operators,
views used after mutations, lazy temporaries and reductions, executed on NumPy arrays and on
DelayArrays side by side; every comparison is bit-exact.  Exercises memo invalidation (buffer
versions), snapshot semantics of evaluated nodes, the plan cache and engine.assign's hazard /
stencil paths."""
import os
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import fuzz_state  # noqa: E402


def test_random_statement_programs_plan_and_compile_without_a_device():
    from delayrepay_b200 import engine
    with engine.dry_run():
        bad = [m for m in (fuzz_state.run_one(s, dry=True) for s in range(25)) if m]
    assert not bad, "\n".join(bad)


@pytest.mark.gpu
def test_random_statement_programs_match_numpy(gpu):
    bad = [m for m in (fuzz_state.run_one(s) for s in range(200)) if m]
    assert not bad, "\n".join(bad)
