# compute-sanitizer over every kernel family (run on the GPU box from the repo root):
#   bash profiles/microbench/sanitize.sh "fused two pipe policy f32 generic open4 spawn warp" "memcheck racecheck"
# One log per (tool, path) under gpurun_out/sanitize/; the summary line of each goes to gpurun_out/sanitize/SUMMARY.txt.
PATHS=${1:-"fused two pipe policy f32 generic open4 spawn warp"}
TOOLS=${2:-"memcheck racecheck"}
mkdir -p gpurun_out/sanitize
: > gpurun_out/sanitize/SUMMARY.txt
for tool in $TOOLS; do
  for p in $PATHS; do
    log=gpurun_out/sanitize/${tool}_${p}.log
    timeout 600 compute-sanitizer --tool $tool --print-limit 20 python profiles/microbench/sanitize_paths.py $p > $log 2>&1
    rc=$?
    echo "$tool $p rc=$rc :: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $log | tail -1) :: $(grep -c '^ok' $log) ok-lines" >> gpurun_out/sanitize/SUMMARY.txt
  done
done
cat gpurun_out/sanitize/SUMMARY.txt
