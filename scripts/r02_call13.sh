#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02m_gpu_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02m_gpu_tests.log | cut -c1-300; grep -n "^E  " gpurun_out/r02m_gpu_tests.log | head -8 | cut -c1-220
