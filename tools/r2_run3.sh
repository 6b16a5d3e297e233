#!/bin/sh
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu -s > gpurun_out/r2i_pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/r2i_pytest_gpu.log
grep -a "passed\|failed\|rel L2\|spp\|Error" gpurun_out/r2i_pytest_gpu.log | tail -12
python tools/sweep.py r2i base n2:lib=n2 n2a:lib=n2a sd:lib=sd mm3:lib=mm3 combo:lib=combo combo2:lib=combo2 st:lib=st base2
for v in n2 sd mm3 combo combo2 st; do
  echo "== parity with variant $v: $(FERMAT_B200_LIB=$PWD/fermat_b200/variants/libfermat_b200_$v.so timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k 'closest_hits or shadow_rays or render_matches or big_scenes or deterministic' 2>&1 | tail -1)"
done
