mkdir -p gpurun_out; cd "${GRAFT_REPO_ROOT:-.}"
for tool in memcheck racecheck synccheck; do
  MG_GEN_SLICES=2 timeout 1200 compute-sanitizer --tool $tool python scripts/sanitize_small.py > gpurun_out/sanitizer_$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|^ok|Error|hazard" gpurun_out/sanitizer_$tool.log | sort | uniq -c | sort -rn | head -12
done
