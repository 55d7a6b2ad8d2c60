"""The product's flank loader (csrc/ramx_loader.c: a strong loadSequenceSubsetMinimal for the CLI)
against the reference's own loader object code (oracle/_ref/libramref.so) on the same files: library
bytes, core records, offsets, boundaries -- and the packed windows it keeps for the device against the
bytes. CPU only."""
import ctypes as C
import os
import tempfile

import numpy as np
import pytest

from repeatafterme_b200 import synth
from repeatafterme_b200.abi import CORE_DTYPE
from tests import cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "repeatafterme_b200", "host", "_refobj", "libramxload.so")


class ProductLoader:
    def __init__(self):
        if not os.path.exists(LIB):
            pytest.skip("host/_refobj/libramxload.so not built (needs /root/reference at build time)")
        self.lib = C.CDLL(LIB)
        vp, i32, u64 = C.c_void_p, C.c_int32, C.c_uint64
        self.lib.ramx_load_flat.argtypes = [C.c_char_p, C.c_char_p, i32, C.POINTER(i32), C.POINTER(u64), vp, vp, vp, vp,
                                            vp, C.POINTER(u64), C.POINTER(u64)]
        self.lib.ramx_load_source.argtypes = [C.c_char_p, vp, C.POINTER(u64), vp, vp, vp]

    def load(self, twobit, ranges, max_flanking_bp):
        n, slen, nw, nn = C.c_int32(), C.c_uint64(), C.c_uint64(), C.c_uint64()
        self.lib.ramx_load_flat(twobit.encode(), ranges.encode(), max_flanking_bp, C.byref(n), C.byref(slen), None,
                                None, None, None, None, None, None)
        cores = np.zeros(n.value, dtype=CORE_DTYPE)
        seq = np.zeros(slen.value, dtype=np.uint8)
        offsets = np.zeros(n.value, dtype=np.uint64)
        bounds = np.zeros(n.value, dtype=np.uint64)
        flags = np.zeros((n.value, 3), dtype=np.int32)
        self.lib.ramx_load_flat(twobit.encode(), ranges.encode(), max_flanking_bp, C.byref(n), C.byref(slen),
                                cores.ctypes.data, seq.ctypes.data, offsets.ctypes.data, bounds.ctypes.data,
                                flags.ctypes.data, C.byref(nw), C.byref(nn))
        bb = C.c_uint64()
        self.lib.ramx_load_source(b"", None, C.byref(bb), None, None, None)
        blob = np.zeros(bb.value, dtype=np.uint8)
        wins = np.zeros((nw.value, 4), dtype=np.uint64)
        ns, nl = np.zeros(nn.value, dtype=np.uint64), np.zeros(nn.value, dtype=np.uint64)
        self.lib.ramx_load_source(b"", blob.ctypes.data, C.byref(bb), wins.ctypes.data, ns.ctypes.data, nl.ctypes.data)
        return dict(seq=seq, cores=cores, offsets=offsets, bounds=bounds, flags=flags, blob=blob, windows=wins,
                    n_start=ns, n_len=nl)


def _decode_windows(src, total):
    """numpy decode of the packed windows + N runs -> byte codes (what the device kernels produce)"""
    out = np.zeros(total, dtype=np.uint8)
    two2ram = np.array([3, 1, 0, 2], dtype=np.uint8)
    for dst, ln, src_byte, src_base in src["windows"].tolist():
        idx = src_base + np.arange(ln)
        byte = src["blob"][src_byte + (idx >> 2)]
        out[dst: dst + ln] = two2ram[(byte >> (6 - 2 * (idx & 3)).astype(np.uint8)) & 3]
    for a, b in zip(src["n_start"].tolist(), src["n_len"].tolist()):
        out[a: a + b] = 99
    return out


def _inputs(name):
    if name == "satellite":
        n, s, b = cases.satellite_inputs()
        return n, s, b, 10014
    if name == "tandem":
        n, s, b = cases.tandem_inputs()
        return n, s, b, 10020
    if name == "overlap":
        n, s, b = cases.overlap_inputs()
        return n, s, b, 10020
    if name == "short_flank":      # windows limited by L, by the sequence ends and by neighbours at once
        n, s, b = cases.tandem_inputs(seed=5)
        return n, s, b, 60
    if name == "n_runs":
        f = synth.make_family(61, 25, 300, 0.1, pad=40, one_sided_frac=0.3, n_frac=0.05)
        return f.names, f.seqs, f.bed, 250
    if name == "many_n_blocks":    # dozens of N blocks per record, several windows on each record: the block search
        n, s, b = cases.tandem_inputs(seed=9)
        rng = np.random.default_rng(9)
        for q in s:
            for st in rng.integers(0, q.size - 12, 60).tolist():
                q[st: st + int(rng.integers(1, 12))] = 99
        return n, s, b, 150
    if name == "unsorted_duplicates":
        f = synth.make_family(62, 12, 120, 0.1, pad=10, one_sided_frac=0.3)
        bed = list(f.bed)[::-1] + [f.bed[3], f.bed[3][:3] + (0, 1, "-")]
        return f.names, f.seqs, bed, 10014
    raise KeyError(name)


def test_product_loader_equals_reference_loader_ce10_standin(ref):
    """BASELINE config 2-3 inputs: 100 cores per family on six chromosomes of ce10 size, 38 neighbour
    pairs of fam1 closer than L + bandwidth (CORE_BOUNDARY / midpoint windows)"""
    from tests import ce10_standin
    with tempfile.TemporaryDirectory() as tmp:
        tb, tsvs = ce10_standin.build(tmp)
        for fam in (1, 2, 3):
            want_seq, want_cores, want_off, want_bounds = ref.load(tb, tsvs[fam], 20040)
            got = ProductLoader().load(tb, tsvs[fam], 20040)
            assert np.array_equal(got["seq"], want_seq)
            assert got["cores"].tobytes() == want_cores.tobytes()
            assert np.array_equal(got["offsets"], want_off) and np.array_equal(got["bounds"], want_bounds)
            assert np.array_equal(_decode_windows(got, want_seq.size), want_seq)


@pytest.mark.parametrize("name", ["satellite", "tandem", "overlap", "short_flank", "n_runs", "many_n_blocks",
                                  "unsorted_duplicates"])
def test_product_loader_equals_reference_loader(ref, name):
    names, seqs, bed, max_flank = _inputs(name)
    with tempfile.TemporaryDirectory() as tmp:
        tb, rg = os.path.join(tmp, "in.2bit"), os.path.join(tmp, "in.tsv")
        synth.write_twobit(tb, names, seqs)
        synth.write_ranges(rg, bed)
        want_seq, want_cores, want_off, want_bounds = ref.load(tb, rg, max_flank)
        got = ProductLoader().load(tb, rg, max_flank)
    assert np.array_equal(got["seq"], want_seq)
    assert got["cores"].tobytes() == want_cores.tobytes()
    assert np.array_equal(got["offsets"], want_off) and np.array_equal(got["bounds"], want_bounds)
    # the windows kept for the device decode to the same library
    assert np.array_equal(_decode_windows(got, want_seq.size), want_seq)
    assert got["windows"][:, 0].tolist() == [0] + want_bounds[:-1].tolist()
