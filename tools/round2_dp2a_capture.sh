# final capture of round 2 (after the IDP.2A sweep): default bench line, ncu --set full of the config-4 launch,
# launch list of a short bench run. Every ncu pass only after the same command has exited 0 without ncu.
O=gpurun_out/r2n
mkdir -p $O
python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
python tools/pair_one.py 10000 5000 2 > $O/pair_one.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:extend_kernel -c 1 -f -o $O/pair \
      python tools/pair_one.py 10000 5000 1 > $O/ncu_pair.log 2>&1
echo "pair capture rc=$?"; cat $O/pair_one.log
ncu -i $O/pair.ncu-rep --page raw --csv > $O/pair_raw.csv 2>/dev/null
ncu -i $O/pair.ncu-rep --page source --csv --print-source cuda,sass > $O/pair_source.csv 2>/dev/null
rm -f $O/pair.ncu-rep
python bench.py --steps 3 --warmup 3 --families 96 --no-cli > $O/bench_short.json 2> $O/bench_short.err && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches.csv \
      python bench.py --steps 3 --warmup 3 --families 96 --no-cli > $O/ncu_bench.log 2>&1
echo "launch list rc=$?"
ls -la $O
