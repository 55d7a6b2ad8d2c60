#!/bin/bash
# GPU tests with the product library, the golden / parity suite again under the -DRAMCU_CHECK build (in-kernel index
# assertions; tools/build_variant.sh check "-DRAMCU_CHECK"), then the default bench. Outputs under gpurun_out/r2h/.
O=gpurun_out/r2h
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
RAMCU_LIB=$PWD/tools/variants/libramcu_check.so python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py tests/test_seed_extend.py -m gpu -q \
  > $O/pytest_check_build.log 2>&1; echo "check build pytest rc=$?"; tail -3 $O/pytest_check_build.log
python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; cat $O/bench.json; tail -4 $O/bench.err
