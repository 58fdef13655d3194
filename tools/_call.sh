set -x
for t in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $t python tools/sanitize_small.py > gpurun_out/r02_san_$t.txt 2>&1
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize driver done" gpurun_out/r02_san_$t.txt | tail -2
done
