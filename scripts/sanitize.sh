#!/bin/bash
# compute-sanitizer over a tiny-shape sweep of every C-ABI kernel family ("racecheck / memcheck on every
# kernel"): run on a B200 box,   gpurun --timeout 1500 -- bash scripts/sanitize.sh   -> gpurun_out/r2_sanitizer_*.txt.
# memcheck: every family.  racecheck / synccheck: the shared-memory + mbarrier protocols (GEMM, attention, LayerNorm).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
summary=gpurun_out/${SANITIZE_SUMMARY:-r2_sanitizer_summary.txt}
: > "$summary"
run() {   # tool, family
  local out=gpurun_out/r2_sanitizer_$1_$2.txt
  timeout 240 "$CS" --tool "$1" --print-limit 20 python scripts/sanitize_sweep.py "$2" > "$out" 2>&1
  local rc=$?
  echo "[$1 / $2] exit $rc : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|done' "$out" | tr '\n' ' ')" | tee -a "$summary"
}
FAMS=${SANITIZE_FAMILIES:-"gemm attention rowwise heads"}      # e.g. SANITIZE_FAMILIES="attention heads" for a partial re-run
for fam in $FAMS; do run memcheck "$fam"; done
for fam in $FAMS; do case $fam in gemm|gemm_r2n|attention|rowwise) run racecheck "$fam";; esac; done
for fam in $FAMS; do case $fam in gemm|gemm_r2n|attention) run synccheck "$fam";; esac; done
cat "$summary"
