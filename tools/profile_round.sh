#!/bin/bash
# tools/profile_round.sh TAG -- the measurement sequence behind profiles/ (run on the GPU box through gpurun):
#   1. GPU tests, 2. the default bench (the number), 3. ncu launch list of a short bench run, 4. one ncu --set full
#   capture of the config-4 launch, 5. ncu --set full of every extend_kernel launch of a 96-family batch.
# Every ncu pass runs only after the same command has exited 0 without ncu. Outputs under gpurun_out/TAG/.
TAG=${1:-prof}
O=gpurun_out/$TAG
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest.log
python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; tail -3 $O/bench.err
python bench.py --steps 3 --warmup 3 --families 96 > $O/bench_short.json 2> $O/bench_short.err && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches.csv \
      python bench.py --steps 3 --warmup 3 --families 96 > $O/ncu_bench.log 2>&1
echo "launch list rc=$?"
python tools/pair_one.py 10000 5000 2 > $O/pair_one.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:extend_kernel -c 1 -f -o $O/pair \
      python tools/pair_one.py 10000 5000 1 > $O/ncu_pair.log 2>&1
echo "pair capture rc=$?"; cat $O/pair_one.log
ncu -i $O/pair.ncu-rep --page raw --csv > $O/pair_raw.csv 2>/dev/null
ncu -i $O/pair.ncu-rep --page source --csv --print-source cuda,sass > $O/pair_source.csv 2>/dev/null
python tools/batch_one.py 96 1 > $O/batch_one.log 2>&1 && \
  ncu --set full --clock-control none -k regex:extend_kernel -c 12 -f -o $O/batch \
      python tools/batch_one.py 96 0 > $O/ncu_batch.log 2>&1
echo "batch capture rc=$?"; cat $O/batch_one.log
ncu -i $O/batch.ncu-rep --page raw --csv > $O/batch_raw.csv 2>/dev/null
rm -f $O/batch.ncu-rep
ls -la $O
