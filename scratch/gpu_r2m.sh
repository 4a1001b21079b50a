#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_host.py -x -q -m gpu -k spce > gpurun_out/r2m_host_spce.log 2>&1; echo "rc=$?" >> gpurun_out/r2m_host_spce.log
tail -15 gpurun_out/r2m_host_spce.log
