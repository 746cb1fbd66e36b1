"""Build libdssm_b200.so (sm_100a) in-tree with nvcc.  `python -m dssm_b200.build [--force]`."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIBDIR = PKG / "lib"
LIB = LIBDIR / "libdssm_b200.so"
SOURCES = ["api.cu", "spmm.cu", "bn.cu", "fc.cu", "fc_tc.cu", "cosloss.cu", "adam.cu", "topk.cu", "topk_tc.cu", "topk_bf16.cu", "nvlink.cu", "metrics.cu", "hostbatch.cu", "tower.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-O2,-Wall,-Wno-unused-function", "--expt-relaxed-constexpr",
]


HASH_FILE = LIBDIR / "libdssm_b200.hash"


def source_hash() -> str:
    """Content hash of everything the library is built from (mtimes do not survive the copy to the GPU box)."""
    import hashlib

    h = hashlib.sha256()
    files = sorted(CSRC.glob("*.cu")) + sorted(CSRC.glob("*.cuh")) + [PKG.parent / "include" / "dssm_b200.h"]
    for f in files:
        h.update(f.name.encode())
        h.update(f.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def up_to_date() -> bool:
    return LIB.exists() and HASH_FILE.exists() and HASH_FILE.read_text().strip() == source_hash()


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: libdssm_b200.so cannot be built")
    return exe


def _stale(out: Path, deps) -> bool:
    if not out.exists():
        return True
    t = out.stat().st_mtime
    return any(Path(d).stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    LIBDIR.mkdir(exist_ok=True)
    if not force and not verbose and up_to_date():
        return LIB
    objdir = PKG / "build"
    objdir.mkdir(exist_ok=True)
    headers = [CSRC / "common.cuh", PKG.parent / "include" / "dssm_b200.h"] + sorted(CSRC.glob("*.cuh"))
    objs, procs = [], []
    for src in SOURCES:
        s = CSRC / src
        o = objdir / (s.stem + ".o")
        objs.append(o)
        if force or not up_to_date() or _stale(o, [s, *headers]):
            cmd = [nvcc(), *NVCC_FLAGS, "-c", str(s), "-o", str(o)]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"--- nvcc failed on {src} ---\n{out}\n")
        elif verbose or out.strip():
            sys.stderr.write(f"--- {src} ---\n{out}\n")
    if failed:
        raise RuntimeError("nvcc compilation failed")
    if force or procs or _stale(LIB, objs):
        cmd = [nvcc(), "-shared", "-o", str(LIB), *map(str, objs), "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
        subprocess.run(cmd, check=True)
    HASH_FILE.write_text(source_hash())
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
