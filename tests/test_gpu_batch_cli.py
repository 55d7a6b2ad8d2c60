"""RAMExtendBatch (csrc/ramx_batch.c): a manifest of RAMExtend command lines run in one process on
one CUDA context -- the families of a chunk share the GPU in one ramcu_batch_run -- must write exactly
what separate runs of the reference RAMExtend write: stdout, -cons, -outtsv, -outfa (goldens produced
by the reference binary, tests/golden/cli_*)."""
import os
import subprocess

import pytest

from repeatafterme_b200 import synth
from tests import golden_io as gio
from tests.cases import filter_stdout

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "repeatafterme_b200", "bin", "RAMExtendBatch")


@pytest.mark.parametrize("chunk,gpus,procs", [(512, 1, 1), (4, 1, 1), (512, 1, 3)])
def test_manifest_of_cli_cases_and_ce10_families(tmp_path, chunk, gpus, procs):
    if not os.path.exists(EXE):
        pytest.skip("bin/RAMExtendBatch not built (needs the reference's host objects)")
    from tests import ce10_standin as cs
    lines, checks = [], []
    for name in gio.cli_case_names():
        case = gio.load_cli(name)
        d = tmp_path / name
        d.mkdir()
        tb, rg = str(d / "in.2bit"), str(d / "in.tsv")
        synth.write_twobit(tb, case["names"], case["seqs"])
        synth.write_ranges(rg, case["bed"])
        lines.append(" ".join([f"stdout={d / 'stdout.txt'}", "-twobit", tb, "-ranges", rg, "-outtsv", str(d / "out.tsv"),
                               "-outfa", str(d / "out.fa"), "-cons", str(d / "cons.fa")] + case["flags"]))
        checks.append((str(d), case["want"]))
    gdir = tmp_path / "ce10"
    gdir.mkdir()
    twobit, tsvs = cs.build(str(gdir))
    for fam in (1, 2, 3):
        d = tmp_path / f"fam{fam}"
        d.mkdir()
        lines.append(" ".join([f"stdout={d / 'stdout.txt'}", "-twobit", twobit, "-ranges", tsvs[fam], "-outtsv",
                               str(d / "out.tsv"), "-outfa", str(d / "out.fa"), "-cons", str(d / "cons.fa")] + cs.STK_FLAGS))
        checks.append((str(d), gio.load_cli_outputs(f"ce10_fam{fam}"), str(gdir)))
    manifest = tmp_path / "manifest.txt"
    manifest.write_text("# one RAMExtend argument list per line\n" + "\n".join(lines) + "\n")
    r = subprocess.run([EXE, "-chunk", str(chunk), "-gpus", str(gpus), "-procs", str(procs), str(manifest)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "families in" in r.stderr
    for chk in checks:
        d, want = chk[0], chk[1]
        out = open(os.path.join(d, "stdout.txt")).read()
        out = filter_stdout(out, chk[2] if len(chk) > 2 else d)
        if len(chk) > 2:
            out = out.replace(d, chk[2])   # output paths of the golden run lived next to the genome
        assert out == want["stdout.txt"], d
        for f in ("out.tsv", "out.fa", "cons.fa"):
            p = os.path.join(d, f)
            assert (open(p).read() if os.path.exists(p) else "") == want[f], (d, f)
