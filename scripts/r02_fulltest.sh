#!/bin/bash
timeout 1200 python -m pytest tests -m gpu -x -q -s > gpurun_out/r02_gpu_tests.log 2>&1; echo "tests rc=$?"
grep -E "c4 full size|KAT|group arm|passed|failed|Error" gpurun_out/r02_gpu_tests.log | cut -c1-400 | tail -60
