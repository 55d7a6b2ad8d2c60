"""Host-side logic that needs no GPU: the C-ABI library loads and exports every symbol the header
declares, struct layouts agree between C and ctypes, the product fails loudly without CUDA, the
2bit writer/reader, the synthetic generators and the family sharding."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

from repeatafterme_b200 import abi, batch, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ramcu.h")


def _has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_library_exports_every_declared_symbol():
    text = open(HEADER).read()
    declared = set(re.findall(r"\b(ramcu_[a-z_0-9]+)\s*\(", text))
    assert declared == set(abi.ABI_SYMBOLS), declared ^ set(abi.ABI_SYMBOLS)
    lib = abi.load_ramcu()
    for name in sorted(declared):
        assert getattr(lib, name) is not None
    assert lib.ramcu_abi_version() == 3


def test_struct_layouts_match_the_header(tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "ramcu.h"\nint main(){printf("%zu %zu %zu %zu\\n",'
                   "sizeof(ramcu_core_t),sizeof(ramcu_scoring_t),sizeof(ramcu_params_t),"
                   "sizeof(ramcu_result_t));return 0;}\n")
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    sizes = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert sizes == [abi.CORE_DTYPE.itemsize, ctypes.sizeof(abi.Scoring), ctypes.sizeof(abi.Params),
                     ctypes.sizeof(abi.ExtendResult)]


def test_builtin_scoring_tables():
    lib = abi.load_ramcu()
    s = abi.scoring_by_name(lib, "14p43g")
    assert s.as_array()[0].tolist() == [9, -18, -10, -21, -1] and (s.gapopen, s.gapextn) == (-33, -7)
    s = abi.scoring_by_name(lib, "25p43g")
    assert s.as_array()[3].tolist() == [-15, -6, -13, 8, -1] and (s.gapopen, s.gapextn) == (-25, -5)
    s = abi.scoring_by_name(lib, "repeatscout", (1, -1, -5))
    assert s.as_array()[2].tolist() == [-1, -1, 1, -1, -1] and (s.gapopen, s.gapextn) == (0, -5)
    with pytest.raises(abi.RamCuError):
        abi.scoring_by_name(lib, "nope")


@pytest.mark.skipif(_has_cuda(), reason="only meaningful on a box without a GPU")
def test_product_fails_loudly_without_cuda():
    with pytest.raises(abi.RamCuError, match="no CUDA device"):
        abi.RamCu(0)
    exe = os.path.join(ROOT, "repeatafterme_b200", "bin", "RAMExtend")
    tb = os.path.join(ROOT, "tests", "golden", "cli_config1_default")
    if os.path.exists(exe):
        from tests import golden_io as gio
        case = gio.load_cli("config1_default")
        import tempfile
        with tempfile.TemporaryDirectory() as tmp:
            synth.write_twobit(os.path.join(tmp, "a.2bit"), case["names"], case["seqs"])
            synth.write_ranges(os.path.join(tmp, "a.tsv"), case["bed"])
            r = subprocess.run([exe, "-twobit", os.path.join(tmp, "a.2bit"), "-ranges",
                                os.path.join(tmp, "a.tsv")], capture_output=True, text=True)
        assert r.returncode == 1 and "no CPU fallback" in r.stdout and "Extended right" not in r.stdout
    assert os.path.isdir(tb)


def test_missing_library_is_an_error(tmp_path):
    with pytest.raises(abi.RamCuError, match="no CPU fallback"):
        abi.load_ramcu(str(tmp_path / "libramcu.so"))


def test_twobit_roundtrip(tmp_path):
    rng = np.random.default_rng(0)
    names = ["a", "chrB", "s0000003"]
    seqs = [rng.integers(0, 4, n, dtype=np.uint8) for n in (1, 1023, 4096 + 3)]
    seqs[1][100:140] = 99
    seqs[2][0:5] = 99
    seqs[2][-3:] = 99
    p = str(tmp_path / "t.2bit")
    synth.write_twobit(p, names, seqs)
    n2, s2 = synth.read_twobit(p)
    assert n2 == names
    for a, b in zip(seqs, s2):
        assert np.array_equal(a, b)


def test_mutate_statistics_and_family_geometry():
    rng = np.random.default_rng(1)
    anc = rng.integers(0, 4, 20000, dtype=np.uint8)
    m = synth.mutate(rng, anc, 0.15)
    assert abs(m.size - anc.size) < 0.02 * anc.size       # insertions ~ deletions
    fam = synth.make_family(3, 50, 120, 0.15, one_sided_frac=0.4, n_frac=0.01)
    c = fam.cores
    assert np.all(c["lower_seq_bound"][1:] == c["upper_seq_bound"][:-1] + 1)   # windows tile sequence[]
    assert c["upper_seq_bound"][-1] == fam.sequence.size - 1
    core_lo = np.minimum(c["left_seq_pos"], c["right_seq_pos"])
    core_hi = np.maximum(c["left_seq_pos"], c["right_seq_pos"])
    assert np.all(core_hi - core_lo == 19)
    assert np.all((core_lo >= c["lower_seq_bound"]) & (core_hi <= c["upper_seq_bound"]))
    fwd = c["orient"] == 0
    assert np.all(c["left_seq_pos"][fwd] < c["right_seq_pos"][fwd])
    assert np.all(c["left_seq_pos"][~fwd] > c["right_seq_pos"][~fwd])
    # a core that is not extendable on one side has no window on that side (sequence.c:625-743)
    noleft_fwd = fwd & (c["left_extendable"] == 0)
    assert np.all(c["lower_seq_bound"][noleft_fwd] == core_lo[noleft_fwd])
    assert (c["left_extendable"] == 0).any() and (c["right_extendable"] == 0).any()


def test_config5_specs_are_deterministic_and_in_range():
    a, b = synth.config5_specs(5, 50), synth.config5_specs(5, 50)
    assert a == b
    for s in a:
        assert 100 <= s["n"] <= 10000 and 500 <= s["flank"] <= 5000
        assert s["params"]["bandwidth"] in (20, 40, 60, 80, 100)
        assert s["params"]["matrix"] in ("14p43g", "18p43g", "20p43g", "25p43g")


def test_lpt_shard_partitions_every_family_once():
    rng = np.random.default_rng(2)
    costs = rng.integers(1, 1000, 37).astype(float).tolist()
    for world in (1, 2, 4, 8):
        shards = batch.lpt_shard(costs, world)
        flat = sorted(i for s in shards for i in s)
        assert flat == list(range(37))
        loads = [sum(costs[i] for i in s) for s in shards]
        assert max(loads) - min(loads) <= max(costs)


def test_twobit_window_spec_matches_the_written_file(tmp_path):
    """synth.family_twobit_windows (input of ramcu_library_from_twobit) describes exactly the packed
    DNA that write_twobit puts into the .2bit records, and decodes back to the loader's bytes"""
    fam = synth.make_family(7, 9, 57, 0.1, one_sided_frac=0.4, n_frac=0.05, pad=5)
    blob, wins, ns, nl, total = synth.family_twobit_windows(fam)
    assert total == fam.sequence.size and wins[:, 1].sum() == total
    two2ram = np.array([3, 1, 0, 2], dtype=np.uint8)
    out = np.zeros(total, dtype=np.uint8)
    for dst, ln, sbyte, sbase in wins.tolist():
        idx = sbase + np.arange(ln)
        byte = blob[sbyte + (idx >> 2)]
        out[dst:dst + ln] = two2ram[(byte >> (6 - 2 * (idx & 3))) & 3]
    for a, b in zip(ns.tolist(), nl.tolist()):
        out[a:a + b] = 99
    assert np.array_equal(out, fam.sequence)
    # the same packed bytes are what the .2bit file holds
    p = str(tmp_path / "f.2bit")
    synth.write_family_files(fam, p, str(tmp_path / "f.tsv"))
    data = open(p, "rb").read()
    rec0 = synth.pack_twobit_record(fam.seqs[0]).tobytes()
    assert rec0 in data
