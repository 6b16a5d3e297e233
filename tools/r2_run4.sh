#!/bin/sh
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "split_shade or big_scenes" -s 2>&1 | grep -a "passed\|failed\|different path\|Error" | tail -12
python tools/sweep.py r2l base split8:FB200_SHADE_SPLIT=1 split6:lib=pb6:FB200_SHADE_SPLIT=1 split5:lib=pb5:FB200_SHADE_SPLIT=1 base2 split8b:FB200_SHADE_SPLIT=1
