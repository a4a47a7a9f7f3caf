#!/bin/bash
# compute-sanitizer over every kernel family at small sizes (scripts/sanitize_paths.py).  Usage (on a B200):
#   bash scripts/sanitize.sh [outdir=gpurun_out] [tools="memcheck synccheck racecheck"]
# Only this library's kernels are instrumented (mangled names contain "4dewi"); racecheck skips the tcgen05 sweeps,
# whose shared memory is written by TMA (asynchronous proxy, ordered by mbarriers that racecheck does not model).
# Each tool runs under its own timeout; logs land in <outdir>/sanitize_<tool>.log, one summary line per tool on stdout.
out=${1:-gpurun_out}
tools=${2:-"memcheck synccheck racecheck"}
mkdir -p "$out"
CS=${COMPUTE_SANITIZER:-/usr/local/cuda/bin/compute-sanitizer}
cd "$(dirname "$0")/.."
timeout 300 python scripts/sanitize_paths.py > "$out/sanitize_plain.log" 2>&1
echo "sanitize plain rc=$? $(tail -n 1 "$out/sanitize_plain.log" | cut -c1-200)"
for tool in $tools; do
  extra=""
  [ "$tool" = racecheck ] && extra="--kernel-name-exclude kns=search_tc --racecheck-report analysis"
  [ "$tool" = memcheck ] && extra="--leak-check no"
  timeout ${SANITIZE_TIMEOUT:-420} "$CS" --tool "$tool" --kernel-name kns=4dewi $extra --error-exitcode 9 --print-limit 40 \
    --log-file "$out/sanitize_$tool.log" python scripts/sanitize_paths.py > "$out/sanitize_${tool}_stdout.log" 2>&1
  rc=$?
  echo "sanitize $tool rc=$rc $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' "$out/sanitize_$tool.log" | tail -n 1) | $(tail -n 1 "$out/sanitize_${tool}_stdout.log" | cut -c1-160)"
done
