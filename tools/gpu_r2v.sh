#!/bin/bash
# round-2 GPU session V (1 GPU): compute-sanitizer memcheck over every round-2 kernel form (small decks)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_small.py > $O/r02_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -12 $O/r02_sanitizer_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"
tail -4 $O/r02_sanitizer_racecheck.log
