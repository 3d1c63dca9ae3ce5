out=gpurun_out/r2_sanitizers.txt; : > $out
for mode in g8 r8; do
  if [ $mode = r8 ]; then export RR_B200_R8=1 RR_B200_R8_MIN_FILL=0; else unset RR_B200_R8 RR_B200_R8_MIN_FILL; fi
  for tool in racecheck memcheck synccheck; do
    echo "== compute-sanitizer --tool $tool python tools/sanitize_large_path.py  (Gram sweeps: $mode; n = 262444, 384 candidates, d = 20)" >> $out
    timeout 600 compute-sanitizer --tool $tool python tools/sanitize_large_path.py 2>&1 | grep -v "^$" | tail -4 >> $out
    echo "$tool rc=$?" >> $out
  done
done
cat $out
