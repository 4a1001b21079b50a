#!/bin/bash
mkdir -p gpurun_out
python -m pytest ${PYTEST_ARGS:-tests} -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -${TAILN:-30} gpurun_out/pytest_gpu.log
