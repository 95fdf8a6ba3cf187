"""In-tree build of libqw_b200.so (sm_100a only).

    python -m qasr_ijcnlp_b200.build [--force] [--verbose]

Each (dtype, n_qubits) instantiation of the QuantumConv1d kernels is its own nvcc invocation so the
build parallelises over the host cores; objects are cached under ``qasr_ijcnlp_b200/_build/`` keyed by a
hash of the sources.  The shared library lands next to this file so it travels with the repo snapshot.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libqw_b200.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC_FLAGS = [
    "-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=...)")


def _units():
    """(object name, source, extra flags)"""
    units = [("qw_api.o", "qw_api.cu", [f'-DQW_BUILD_STAMP="{_source_hash()}"']), ("qw_conv1d.o", "qw_conv1d.cu", []), ("qw_logmel.o", "qw_logmel.cu", []),
             ("qw_conv1d_fast.o", "qw_conv1d_fast.cu", []), ("qw_conv1d_general.o", "qw_conv1d_general.cu", []), ("qw_dp.o", "qw_dp.cu", []),
             ("qw_stem.o", "qw_stem.cu", []), ("qw_stem_train.o", "qw_stem_train.cu", []), ("qw_collapsed.o", "qw_collapsed.cu", [])]
    for tname, t in (("f32", "float"), ("f64", "double")):
        for q in (1, 2, 3, 4):
            units.append((f"qw_conv1d_inst_{tname}_q{q}.o", "qw_conv1d_inst.cu", [f"-DQW_T={t}", f"-DQW_Q={q}"]))
    units.append(("qw_circuit_warp.o", "qw_circuit_warp.cu", []))
    for tname, t in (("f32", "float"), ("f64", "double")):
        for q in range(1, 13):
            units.append((f"qw_circuit_warp_inst_{tname}_q{q}.o", "qw_circuit_warp_inst.cu", [f"-DQW_T={t}", f"-DQW_Q={q}"]))
    return units


def _source_hash() -> str:
    h = hashlib.sha256()
    files = sorted(os.listdir(CSRC)) + [os.path.join(INCLUDE, f) for f in sorted(os.listdir(INCLUDE))]
    for f in files:
        p = f if os.path.isabs(f) else os.path.join(CSRC, f)
        if os.path.isfile(p):
            h.update(f.encode())
            with open(p, "rb") as fh:
                h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()[:16]


def _unit_hash(src: str, extra) -> str:
    """Hash of what one object depends on: its own .cu, every header of the tree, the flags."""
    h = hashlib.sha256()
    deps = [os.path.join(CSRC, src)]
    deps += sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h")))
    deps += [os.path.join(INCLUDE, f) for f in sorted(os.listdir(INCLUDE))]
    for p in deps:
        h.update(os.path.basename(p).encode())
        with open(p, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS + list(extra)).encode())
    return h.hexdigest()[:16]


def is_stale() -> bool:
    """True when libqw_b200.so does not match the sources next to it (or is missing)."""
    stamp = os.path.join(BUILD, "stamp.txt")
    return not (os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read().strip() == _source_hash())


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(BUILD, exist_ok=True)
    stamp = os.path.join(BUILD, "stamp.txt")
    want = _source_hash()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read().strip() == want:
        return LIB
    nvcc = _nvcc()
    units = _units()

    def compile_one(u):
        obj, src, extra = u
        out = os.path.join(BUILD, obj)
        hfile = out + ".hash"
        uh = _unit_hash(src, extra)
        if not force and not verbose and os.path.exists(out) and os.path.exists(hfile) and open(hfile).read().strip() == uh:
            return obj, 0, ""  # object is current: only the units whose sources changed recompile
        cmd = [nvcc, *NVCC_FLAGS, *extra, "-I", INCLUDE, "-c", os.path.join(CSRC, src), "-o", out]
        if verbose:
            cmd.insert(1, "-Xptxas")
            cmd.insert(2, "-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode == 0:
            with open(hfile, "w") as fh:
                fh.write(uh)
        return obj, r.returncode, r.stdout + r.stderr

    with ThreadPoolExecutor(max_workers=min(len(units), os.cpu_count() or 4)) as ex:
        results = list(ex.map(compile_one, units))
    failed = [(o, out) for o, rc, out in results if rc != 0]
    if verbose or failed:
        for o, rc, out in results:
            if out.strip():
                print(f"---- {o}\n{out}", file=sys.stderr)
    if failed:
        raise RuntimeError("nvcc failed for: " + ", ".join(o for o, _ in failed))
    objs = [os.path.join(BUILD, o) for o, _, _ in units]
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs, "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    with open(stamp, "w") as fh:
        fh.write(want)
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
