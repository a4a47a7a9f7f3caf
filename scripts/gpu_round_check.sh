#!/bin/bash
# One GPU-box visit at the end of a round: the established GPU suite, the default bench line, the newest tests on their
# own (so that a failure there cannot hide the rest), compute-sanitizer over every kernel family.  Everything lands in
# gpurun_out/ under the tag given as $1; each step runs under its own timeout.
tag=${1:-check}
out=gpurun_out
mkdir -p $out
NEW='shape_sweep or large_k or candidate_limit or long_lists or adversarial or stub_runs'
t0=$(date +%s)
timeout -k 10 600 python -m pytest tests -q -m gpu -k "not ($NEW)" -p no:cacheprovider --timeout 300 > $out/${tag}_gpu_tests.log 2>&1
echo "[$(( $(date +%s) - t0 ))s] established suite rc=$? $(tail -n 1 $out/${tag}_gpu_tests.log)"
timeout -k 10 420 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err
echo "[$(( $(date +%s) - t0 ))s] bench rc=$? $(head -c 300 $out/${tag}_bench.json)"
timeout -k 10 300 python -m pytest tests -q -m gpu -k "$NEW" -p no:cacheprovider --timeout 200 --tb=short > $out/${tag}_gpu_tests_new.log 2>&1
echo "[$(( $(date +%s) - t0 ))s] new tests rc=$? $(tail -n 1 $out/${tag}_gpu_tests_new.log)"
SANITIZE_TIMEOUT=${SANITIZE_TIMEOUT:-240} bash scripts/sanitize.sh $out "${SANITIZE_TOOLS:-memcheck racecheck synccheck}" 2>&1 | sed "s/^/[sanitize] /"
for f in $out/sanitize_*; do mv "$f" "$out/${tag}_$(basename $f)"; done
echo "[$(( $(date +%s) - t0 ))s] done"
