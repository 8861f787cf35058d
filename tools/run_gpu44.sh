for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool python tools/sanitizer_case.py > gpurun_out/r02_sanitizer_$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Error" gpurun_out/r02_sanitizer_$tool.log | sort | uniq -c | head -8
done
