#!/bin/sh
mkdir -p gpurun_out; L=gpurun_out/r2e_bisect.log; : > $L
T=tests/test_gpu_parity.py
SEL="not full_size and not 1024 and not bathroom2_full and not nee_over"
timeout 300 python -m pytest $T -q -m gpu -x -k "$SEL" > /tmp/o.log 2>&1; rc=$?
echo "== plain -> exit $rc: $(tail -1 /tmp/o.log)" >> $L
for k in "material_testball or directional" "water_caustic or directional" "big_scenes or directional"; do
  timeout 300 python -m pytest $T -q -m gpu -x -k "$k" > /tmp/o.log 2>&1
  echo "== -k '$k' -> exit $?: $(tail -1 /tmp/o.log)" >> $L
done
if [ $rc -ne 0 ]; then
  timeout 1200 compute-sanitizer --tool memcheck --print-limit 3 python -m pytest $T -q -m gpu -x -k "$SEL" > /tmp/s.log 2>&1
  echo "== sanitizer exit $?" >> $L
  grep -a -v "^$" /tmp/s.log | grep -a -A45 "Invalid\|=== ERROR\|Error" | head -120 >> $L
  tail -5 /tmp/s.log >> $L
fi
cat $L | cut -c1-250 | head -150
