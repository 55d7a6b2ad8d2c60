"""In-tree build of the native pieces (nvcc cross-compiles sm_100a without a GPU).

  csrc/libramcu.so   CUDA kernels + extern "C" ABI (include/ramcu.h)          -- the product
  bin/RAMExtend      the reference's own host code (compiled out of tree into host/_refobj when
                     /root/reference is present) with extend_alignment(), the flank loader and the DP
                     scratch allocation replaced (csrc/ramx_boundary.c, csrc/ramx_loader.c)  -- the product CLI
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources if os.path.exists(s))


KERNEL_UNITS = ["kernels_b14", "kernels_b20", "kernels_b40", "kernels_b60", "kernels_b80", "kernels_b100",
                "kernels_generic"]


def _nvcc_obj(unit, verbose=False):
    src, obj = os.path.join(CSRC, unit + ".cu"), os.path.join(CSRC, unit + ".o")
    cmd = [NVCC, *ARCH, "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-c", src, "-o", obj]
    if verbose:
        cmd[1:1] = ["-Xptxas", "-v"]
    log = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if log.returncode != 0:
        raise RuntimeError(f"nvcc failed for {unit}:\n{log.stdout}\n{log.stderr}")
    return obj, log.stderr


def build_libramcu(force=False, verbose=False):
    """nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo, one object per translation unit
    (compiled in parallel), linked into csrc/libramcu.so."""
    from concurrent.futures import ThreadPoolExecutor

    out = os.path.join(CSRC, "libramcu.so")
    common = [os.path.join(CSRC, f) for f in ("extend_kernel.cuh", "aux_kernels.cuh", "flatten.h", "kernels.h")]
    common.append(os.path.join(ROOT, "include", "ramcu.h"))
    units = ["ramcu"] + KERNEL_UNITS
    todo = [u for u in units
            if force or _newer(os.path.join(CSRC, u + ".o"), [os.path.join(CSRC, u + ".cu")] + common)]
    logs = {}
    if todo:
        with ThreadPoolExecutor(max_workers=min(len(todo), os.cpu_count() or 1)) as ex:
            for u, (obj, err) in zip(todo, ex.map(lambda u: _nvcc_obj(u, verbose), todo)):
                logs[u] = err
    objs = [os.path.join(CSRC, u + ".o") for u in units]
    if todo or not os.path.exists(out):
        subprocess.check_call([NVCC, *ARCH, "--shared", "-o", out, *objs], cwd=CSRC)
    if verbose:
        for u, err in logs.items():
            print(err)
    return out


REF_HOST_SRC = ["ram_extend", "common", "report", "score_system", "cmd_line_opts", "version", "sequence"]
KENT_SRC = ["twoBitNew", "common", "memalloc", "hash", "errAbort", "obscure", "localmem", "dnautil", "dlist", "dystring",
            "linefile", "options", "verbose", "fa", "dnaseq", "bits"]


def build_cli(force=False):
    """bin/RAMExtend = the reference's own host code (main, flag parsing, report, every output; compiled
    out of tree from /root/reference into host/_refobj, nothing copied) with three symbols replaced:
    extend_alignment() by csrc/ramx_boundary.c (the GPU path), loadSequenceSubsetMinimal() and
    allocate_score() by csrc/ramx_loader.c. The reference's DP object (bnw_extend.o) is not linked.
    Also host/_refobj/libramxload.so: the product loader behind a flat entry point, for the tests."""
    ref = os.environ.get("RAM_REFERENCE", "/root/reference")
    obj_dir = os.path.join(HERE, "host", "_refobj")
    bin_dir = os.path.join(HERE, "bin")
    out = os.path.join(bin_dir, "RAMExtend")
    have_ref = os.path.exists(os.path.join(ref, "ram_extend.c"))
    if not have_ref:
        return out if os.path.exists(out) else None   # GPU box: the prebuilt binary travels
    os.makedirs(obj_dir, exist_ok=True)
    os.makedirs(bin_dir, exist_ok=True)
    gcc = "/usr/bin/gcc"
    ours = [os.path.join(CSRC, f) for f in ("ramx_boundary.c", "ramx_loader.c", "ramx_loader_flat.c", "ramx_loader.h",
                                            "ramx_boundary.h", "ramx_batch.c")]
    ours.append(os.path.join(ROOT, "include", "ramcu.h"))
    if not force and os.path.exists(out) and not _newer(out, ours + [os.path.join(CSRC, "libramcu.so")]):
        return out
    jobs = []
    for name in REF_HOST_SRC:   # the reference's flags (Makefile:16)
        o = os.path.join(obj_dir, name + ".o")
        if force or not os.path.exists(o):
            jobs.append([gcc, "-std=gnu11", "-fpic", "-mtune=native", "-g", "-O3", "-w", "-I" + os.path.join(ref, "minunit"),
                         "-I" + ref, "-c", os.path.join(ref, name + ".c"), "-o", o])
    for name in KENT_SRC:
        o = os.path.join(obj_dir, "kent_" + name + ".o")
        if force or not os.path.exists(o):
            jobs.append([gcc, "-g", "-fpic", "-O2", "-w", "-I" + os.path.join(ref, "kentsrc"), "-c",
                         os.path.join(ref, "kentsrc", name + ".c"), "-o", o])
    inc = ["-I" + ref, "-I" + os.path.join(ref, "minunit"), "-I" + os.path.join(ROOT, "include"), "-I" + CSRC]
    for name in ("ramx_boundary", "ramx_loader", "ramx_loader_flat", "ramx_batch"):
        jobs.append([gcc, "-std=gnu11", "-O2", "-fpic", "-Wall", *inc, "-c", os.path.join(CSRC, name + ".c"),
                     "-o", os.path.join(obj_dir, name + ".o")])
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=os.cpu_count() or 1) as ex:
        list(ex.map(subprocess.check_call, jobs))
    weak_main = os.path.join(obj_dir, "ram_extend_weak.o")
    subprocess.check_call(["objcopy", "--weaken-symbol=extend_alignment", os.path.join(obj_dir, "ram_extend.o"), weak_main])
    weak_seq = os.path.join(obj_dir, "sequence_weak.o")
    subprocess.check_call(["objcopy", "--weaken-symbol=loadSequenceSubsetMinimal", os.path.join(obj_dir, "sequence.o"), weak_seq])
    kent = os.path.join(obj_dir, "libTwoBit.a")
    if os.path.exists(kent):
        os.remove(kent)
    subprocess.check_call(["ar", "rcs", kent] + [os.path.join(obj_dir, "kent_" + k + ".o") for k in KENT_SRC])
    objs = [weak_main, weak_seq] + [os.path.join(obj_dir, o + ".o") for o in REF_HOST_SRC if o not in ("ram_extend", "sequence")]
    subprocess.check_call([gcc, "-o", out, os.path.join(obj_dir, "ramx_boundary.o"), os.path.join(obj_dir, "ramx_loader.o"),
                           *objs, "-L" + obj_dir, "-lTwoBit", "-lm", "-L" + CSRC, "-lramcu", "-Wl,-rpath,$ORIGIN/../csrc"])
    # bin/RAMExtendBatch: the same objects with the reference's main() renamed, driven by csrc/ramx_batch.c
    batch_main = os.path.join(obj_dir, "ram_extend_batch.o")
    subprocess.check_call(["objcopy", "--redefine-sym", "main=ramref_main", weak_main, batch_main])
    subprocess.check_call([gcc, "-o", os.path.join(bin_dir, "RAMExtendBatch"), os.path.join(obj_dir, "ramx_batch.o"),
                           os.path.join(obj_dir, "ramx_boundary.o"), os.path.join(obj_dir, "ramx_loader.o"), batch_main,
                           *objs[1:], "-L" + obj_dir, "-lTwoBit", "-lm", "-L" + CSRC, "-lramcu",
                           "-Wl,-rpath,$ORIGIN/../csrc"])
    subprocess.check_call([gcc, "-shared", "-o", os.path.join(obj_dir, "libramxload.so"),
                           os.path.join(obj_dir, "ramx_loader.o"), os.path.join(obj_dir, "ramx_loader_flat.o")])
    return out


def build_emulator():
    """tests/emul/libramemul.so: the kernel source compiled for the CPU (test infrastructure; built
    here so that a kernel edit that breaks the emulation shim shows up in the build check). Built with
    -DRAMCU_CHECK: every CPU test of the kernel source also runs its index assertions."""
    emul = os.path.join(ROOT, "tests", "emul")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-DRAMCU_EMUL", "-DRAMCU_CHECK", "-I.", "-fPIC", "-shared", "-pthread",
                           "-o", "libramemul.so", "emul_driver.cpp"], cwd=emul)
    return os.path.join(emul, "libramemul.so")


def build_all(force=False):
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "all"])
    lib = build_libramcu(force=force)
    cli = build_cli(force=force)
    build_emulator()
    return lib, cli


if __name__ == "__main__":
    print(build_all(force="--force" in sys.argv))
