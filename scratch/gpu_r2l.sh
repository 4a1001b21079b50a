#!/bin/bash
# round 2, run L: lj/long/coul/long + special bonds on data.spce (in.spce non-bonded + k-space force), then the whole GPU suite
mkdir -p gpurun_out
python -m pytest tests/test_gpu_spce.py -x -q -m gpu > gpurun_out/r2l_spce.log 2>&1; echo "spce rc=$?" >> gpurun_out/r2l_spce.log
tail -25 gpurun_out/r2l_spce.log
python -m pytest tests -q -m gpu > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2l_pytest.log
tail -6 gpurun_out/r2l_pytest.log
