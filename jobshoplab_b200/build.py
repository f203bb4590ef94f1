"""Build libjsl_b200.so in-tree with nvcc for sm_100a (no JIT, no torch extension machinery).

    python -m jobshoplab_b200.build [--force]

Each translation unit is compiled by its own nvcc process (they run in parallel) and linked into
jobshoplab_b200/csrc/libjsl_b200.so, which travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

CSRC = Path(__file__).resolve().parent / "csrc"
INCLUDE = Path(__file__).resolve().parent.parent / "include"
LIB = Path(os.environ.get("JSL_LIB_OUT", CSRC / "libjsl_b200.so"))
LOG_LIB = CSRC / "libjsl_b200_log.so"
EXTRA = os.environ.get("JSL_NVCC_EXTRA", "").split()
OBJ_TAG = os.environ.get("JSL_OBJ_TAG", "")
UNITS = ["jsl_abi.cu", "jsl_kernels_jw1.cu", "jsl_kernels_jw2.cu", "jsl_kernels_jw4.cu"]
HEADERS = [CSRC / "jsl_engine.cuh", CSRC / "jsl_kernels.cuh", CSRC / "jsl_randomize.cuh", CSRC / "jsl_tobj.hpp", INCLUDE / "jobshoplab_b200.h"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--fmad=false",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(exe).exists():
        raise RuntimeError("nvcc not found: the CUDA engine cannot be built")
    return exe


def _compile(unit: str, force: bool, tag: str = "", extra=()) -> Path:
    tag = OBJ_TAG + tag
    src, obj = CSRC / unit, CSRC / (unit[:-3] + tag + ".o")
    newest = max(p.stat().st_mtime for p in [src, *HEADERS])
    if not force and obj.exists() and obj.stat().st_mtime >= newest:
        return obj
    log = subprocess.run([nvcc(), *NVCC_FLAGS, *EXTRA, *extra, "-c", str(src), "-o", str(obj)], capture_output=True, text=True)
    (CSRC / (unit[:-3] + tag + ".ptxas.log")).write_text(log.stderr)
    if log.returncode != 0:
        raise RuntimeError(f"nvcc failed for {unit}:\n{log.stderr[-4000:]}")
    return obj


def _link(lib: Path, objs, force: bool):
    if force or not lib.exists() or lib.stat().st_mtime < max(o.stat().st_mtime for o in objs):
        subprocess.check_call([nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", str(lib),
                               *map(str, objs), "-lcudart"])


def build(force: bool = False) -> Path:
    """libjsl_b200.so (throughput build) and libjsl_b200_log.so (same sources with -DJSL_LEG_LOG: keeps the
    transport-task log env.render() needs; used for batches created with record_ops)."""
    jobs = [(u, "", ()) for u in UNITS] + ([] if OBJ_TAG else [(u, "_log", ("-DJSL_LEG_LOG",)) for u in UNITS])
    with ThreadPoolExecutor(len(jobs)) as ex:
        objs = list(ex.map(lambda j: _compile(j[0], force, j[1], j[2]), jobs))
    _link(LIB, objs[:len(UNITS)], force)
    if not OBJ_TAG:
        _link(LOG_LIB, objs[len(UNITS):], force)
    return LIB


def build_variant(tag: str, extra_flags, lib: Path, force: bool = False) -> Path:
    """The same sources with extra nvcc flags into their own library, e.g. a user reward hook:
    build_variant("_user", ['-DJSL_USER_HEADER="examples/jsl_user_example.cuh"'], CSRC / "libjsl_b200_user.so")."""
    with ThreadPoolExecutor(len(UNITS)) as ex:
        objs = list(ex.map(lambda u: _compile(u, force, tag, tuple(extra_flags)), UNITS))
    _link(lib, objs, force)
    return lib


USER_EXAMPLE_LIB = CSRC / "libjsl_b200_user.so"


def build_user_example(force: bool = False) -> Path:
    """libjsl_b200_user.so: the throughput build with csrc/examples/jsl_user_example.cuh as reward hook (tests)."""
    hdr = CSRC / "examples" / "jsl_user_example.cuh"
    if hdr not in HEADERS:
        HEADERS.append(hdr)
    return build_variant("_user", ['-DJSL_USER_HEADER="examples/jsl_user_example.cuh"'], USER_EXAMPLE_LIB, force)


if __name__ == "__main__":
    print(build("--force" in sys.argv))
