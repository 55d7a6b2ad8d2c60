#!/bin/bash
# one product CLI run on config-4 files with timing lines (development aid)
set -x
cd "$(dirname "$0")/.."
python - <<'PY'
import sys; sys.path.insert(0,'.')
from repeatafterme_b200 import synth
fam = synth.make_family(4, int(sys.argv[1]) if len(sys.argv)>1 else 10000, 5000, 0.15)
synth.write_family_files(fam, "/tmp/f.2bit", "/tmp/f.tsv")
PY
export RAMX_TIMING=1
( time timeout 120 repeatafterme_b200/bin/RAMExtend -twobit /tmp/f.2bit -ranges /tmp/f.tsv -bandwidth 40 -matrix 14p43g -minimprovement 30 -L 20000 -outtsv /tmp/o.tsv -outfa /tmp/o.fa -cons /tmp/c.fa > /tmp/out.txt ) 2>&1 | tail -20
echo "rc=$?"; tail -3 /tmp/out.txt; grep -c . /tmp/out.txt
