set -x
for t in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $t python tools/sanitize_small.py > gpurun_out/san_$t.log 2>&1
  echo "$t : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/san_$t.log | tail -1)"
  tail -2 gpurun_out/san_$t.log | cut -c1-200
done
