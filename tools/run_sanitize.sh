for tool in memcheck racecheck synccheck initcheck; do
  echo "== $tool"
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 7 python tools/sanitize_scene.py 2>&1 | tail -4
  echo "exit $?"
done 2>&1 | tee gpurun_out/r02_sanitizers.txt
