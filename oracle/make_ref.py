"""Recipe for oracle/_ref: the UNMODIFIED reference, installed so that it can travel to the GPU box.

TEST / BASELINE INFRASTRUCTURE ONLY.  /root/reference exists in the build container only; oracle/_ref is
git-ignored (no reference source ever enters the history) but not gpurun-ignored, so the GPU box gets it with the
repo snapshot, next to the import shims of oracle/shims (gymnasium, hydra, omegaconf, plotly, dash*, readchar are
absent from the image).  bench.py's cpu_baseline / --impl reference legs and the seam-2 GPU test
(tests/test_middleware.py) import the reference from here.

    python oracle/make_ref.py [--force]

Steps (the base contract's install recipe, with --no-deps because the dependency list cannot be resolved offline
and from a copy under /tmp because the build writes into the source tree):
    pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target oracle/_ref <copy>
    cp -r /root/reference/data oracle/_ref/data            (instances + configs the reference loads by path)
"""
from __future__ import annotations

import shutil
import subprocess
import sys
import tempfile
from pathlib import Path

HERE = Path(__file__).resolve().parent
SRC = Path("/root/reference")
DST = HERE / "_ref"


def available() -> bool:
    return (DST / "jobshoplab" / "__init__.py").exists()


def build(force: bool = False) -> bool:
    """Returns True when oracle/_ref is usable afterwards."""
    if available() and not force:
        return True
    if not SRC.exists():
        return False
    if DST.exists():
        shutil.rmtree(DST)
    with tempfile.TemporaryDirectory() as tmp:
        copy = Path(tmp) / "reference"
        shutil.copytree(SRC, copy, ignore=shutil.ignore_patterns(".git", "docs", "jupyter"))
        r = subprocess.run([sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
                            "--find-links", "/opt/wheelhouse", "--target", str(DST), str(copy)],
                           capture_output=True, text=True)
        if r.returncode != 0 or not available():
            # setuptools missing or similar: the package is pure Python, a plain copy is the same thing
            DST.mkdir(parents=True, exist_ok=True)
            shutil.copytree(SRC / "jobshoplab", DST / "jobshoplab", dirs_exist_ok=True)
    for extra in ("docs", "jupyter", "scripts", "tests"):  # stray top-level packages of the wheel
        shutil.rmtree(DST / extra, ignore_errors=True)
    shutil.copytree(SRC / "data", DST / "data", dirs_exist_ok=True)
    return available()


if __name__ == "__main__":
    ok = build("--force" in sys.argv)
    print("oracle/_ref", "ready" if ok else "unavailable (no /root/reference here)")
