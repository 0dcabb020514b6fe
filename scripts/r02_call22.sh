#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02v_gpu_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r02v_gpu_tests.log | cut -c1-200; grep -n "^E  " gpurun_out/r02v_gpu_tests.log | head -5 | cut -c1-200
