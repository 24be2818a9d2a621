#!/bin/bash
# compute-sanitizer over one small case per kernel family (selftest sanitize <family>): memcheck, racecheck (shared-memory
# hazards between the TMA / MMA / epilogue warps), synccheck (barrier misuse) and initcheck.  Summaries (last lines of every
# run) go to $OUT (default gpurun_out/sanitizer); the judged copy is profiles/r02_sanitizer.txt.
set -u
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
OUT="${1:-$ROOT/gpurun_out/sanitizer}"
mkdir -p "$OUT"
BIN="$ROOT/stabletriton_b200/csrc/selftest"
SUMMARY="$OUT/summary.txt"
: > "$SUMMARY"
for tool in memcheck racecheck synccheck initcheck; do
  for fam in gemm conv attn norm misc; do
    log="$OUT/${tool}_${fam}.log"
    timeout 900 compute-sanitizer --tool "$tool" --print-limit 20 "$BIN" sanitize "$fam" > "$log" 2>&1
    rc=$?
    echo "== $tool $fam (exit $rc): $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|failure\(s\)' "$log" | tr '\n' ' ')" >> "$SUMMARY"
  done
done
cat "$SUMMARY"
