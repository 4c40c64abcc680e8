#!/bin/bash
# Rebuild the working set of the cubin cache and drop everything else (stale generations of
# kernels whose source changed, one-off experiments): touches every entry that build(), the CPU
# test suite and the seeds of the GPU fuzz tests need, then deletes the untouched ones.
set -e
cd "$(dirname "$0")/.."
touch /tmp/dr_prune_mark
sleep 1
export DR_CACHE_TOUCH=1
python __graft_entry__.py > /tmp/dr_prune_build.log 2>&1
python -m pytest tests -x -q -m "not gpu" > /tmp/dr_prune_tests.log 2>&1
python tools/fuzz_diff.py --n 300 --seed 0 --dry > /dev/null 2>&1
python tools/fuzz_diff.py --n 300 --seed 1000 --dry > /dev/null 2>&1
python tools/fuzz_state.py --n 200 --seed 0 --dry > /dev/null 2>&1
before=$(ls delayrepay_b200/_cache | wc -l)
find delayrepay_b200/_cache -name '*.cubin' ! -newer /tmp/dr_prune_mark -delete
echo "cubin cache: $before -> $(ls delayrepay_b200/_cache | wc -l) entries, $(du -sh delayrepay_b200/_cache | cut -f1)"
