mkdir -p gpurun_out; cd "${GRAFT_REPO_ROOT:-.}"
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool python scripts/sanitize_small.py > gpurun_out/sanitizer_$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|^ok|Error|hazard" gpurun_out/sanitizer_$tool.log | head -12
done
