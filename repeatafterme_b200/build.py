"""In-tree build of the native pieces (nvcc cross-compiles sm_100a without a GPU).

  csrc/libramcu.so   CUDA kernels + extern "C" ABI (include/ramcu.h)          -- the product
  bin/RAMExtend      the reference's own host objects with extend_alignment() replaced by
                     csrc/ramx_boundary.c (only when oracle/_ref/obj exists)     -- the product CLI
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


def build_cli(force=False):
    """Link bin/RAMExtend = reference host objects (compiled out of tree by oracle/Makefile from
    /root/reference, extend_alignment weakened) + our boundary + libramcu.so."""
    obj_dir = os.path.join(ROOT, "oracle", "_ref", "obj")
    boundary = os.path.join(CSRC, "ramx_boundary.c")
    if not os.path.isdir(obj_dir) or not os.path.exists(boundary):
        return None
    bin_dir = os.path.join(HERE, "bin")
    os.makedirs(bin_dir, exist_ok=True)
    out = os.path.join(bin_dir, "RAMExtend")
    ref_objs = ["ram_extend", "common", "report", "score_system", "cmd_line_opts", "version",
                "sequence", "bnw_extend"]
    srcs = [boundary, os.path.join(CSRC, "libramcu.so")] + [os.path.join(obj_dir, o + ".o") for o in ref_objs]
    if not force and not _newer(out, srcs):
        return out
    weak = os.path.join(bin_dir, "ram_extend_weak.o")
    subprocess.check_call(["objcopy", "--weaken-symbol=extend_alignment",
                           os.path.join(obj_dir, "ram_extend.o"), weak])
    bobj = os.path.join(bin_dir, "ramx_boundary.o")
    ref_inc = os.environ.get("RAM_REFERENCE", "/root/reference")
    if not os.path.exists(os.path.join(ref_inc, "common.h")):
        return out if os.path.exists(out) else None
    subprocess.check_call(["/usr/bin/gcc", "-std=gnu11", "-O2", "-fpic", "-Wall", "-I" + ref_inc,
                           "-I" + os.path.join(ref_inc, "minunit"), "-I" + os.path.join(ROOT, "include"),
                           "-c", boundary, "-o", bobj])
    objs = [weak] + [os.path.join(obj_dir, o + ".o") for o in ref_objs if o != "ram_extend"]
    subprocess.check_call(["/usr/bin/gcc", "-o", out, bobj, *objs,
                           "-L" + os.path.join(ROOT, "oracle", "_ref"), "-lTwoBit", "-lm",
                           "-L" + CSRC, "-lramcu", "-Wl,-rpath,$ORIGIN/../csrc"])
    return out


def build_emulator():
    """tests/emul/libramemul.so: the kernel source compiled for the CPU (test infrastructure; built
    here so that a kernel edit that breaks the emulation shim shows up in the build check)"""
    emul = os.path.join(ROOT, "tests", "emul")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-DRAMCU_EMUL", "-I.", "-fPIC", "-shared", "-pthread",
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
