O=gpurun_out/r2f
mkdir -p $O
python tools/pair_one.py 10000 5000 2 > $O/pair_one.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:extend_kernel -c 1 -f -o $O/pair \
      python tools/pair_one.py 10000 5000 1 > $O/ncu_pair.log 2>&1
echo "pair capture rc=$?"; cat $O/pair_one.log
ncu -i $O/pair.ncu-rep --page raw --csv > $O/pair_raw.csv 2>/dev/null
ncu -i $O/pair.ncu-rep --page source --csv --print-source cuda,sass > $O/pair_source.csv 2>/dev/null
ncu -i $O/pair.ncu-rep --page source --csv > $O/pair_source_default.csv 2>/dev/null
python bench.py --steps 3 --warmup 3 --families 96 > $O/bench_short.json 2> $O/bench_short.err && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches.csv \
      python bench.py --steps 3 --warmup 3 --families 96 > $O/ncu_bench.log 2>&1
echo "launch list rc=$?"
timeout 300 python tools/cli_bench.py > $O/cli_bench.json 2> $O/cli_bench.err; echo "cli rc=$?"; cat $O/cli_bench.json
timeout 400 python tools/batch_cli_bench.py 24 200 300 > $O/batch_cli.json 2> $O/batch_cli.err; echo "batchcli rc=$?"; cat $O/batch_cli.json
ls -la $O
