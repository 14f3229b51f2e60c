"""Build libafa_sm100.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)                      # .../diffbinaural-binaural-audio-generation_b200
_REPO = os.path.dirname(_ROOT)
_CSRC = os.path.join(_ROOT, "csrc")
_INCLUDE = os.path.join(_REPO, "include")
LIB_NAME = "libafa_sm100.so"


def library_path() -> str:
    return os.path.join(_HERE, LIB_NAME)


_COMMON = [os.path.join(_INCLUDE, "afa_b200.h"), os.path.join(_CSRC, "afa_internal.h")]
# translation unit -> the headers only it includes (an object is rebuilt when its unit, its headers or _COMMON change)
_UNITS = {
    "afa_capi.cu": ["afa_kernels.cuh", "afa_cl_kernels.cuh", "afa_actconv_kernels.cuh"],
    "afa_mel.cu": [],
    "afa_tc.cu": ["afa_tc_kernels.cuh"],
    "afa_tc_cl.cu": ["afa_tc_kernels.cuh", "afa_tc_cl_kernels.cuh"],
    "afa_ingest.cu": [],
}


def _sources():
    return [os.path.join(_CSRC, u) for u in _UNITS]


def _deps():
    return _sources() + [os.path.join(_CSRC, h) for hs in _UNITS.values() for h in hs] + _COMMON


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build libafa_sm100.so (there is no CPU fallback)")


LAST_BUILD = {"action": None, "compiled_units": []}   # what the most recent build_library() call did (for build()'s report)


def build_library(force: bool = False, verbose: bool = False, out: str | None = None, defines=()) -> str:
    """Compile csrc/ -> afa_b200/libafa_sm100.so.  Rebuilds only when a source is newer (AFA_FORCE_REBUILD=1 rebuilds
    everything).  `out` / `defines` build an experimental variant elsewhere (tuning sweeps)."""
    out = out or library_path()
    force = force or os.environ.get("AFA_FORCE_REBUILD", "0") == "1"
    LAST_BUILD["compiled_units"] = []
    if not force and os.path.exists(out):
        t = os.path.getmtime(out)
        if all(os.path.getmtime(d) <= t for d in _deps()):
            LAST_BUILD["action"] = "reused: the library is newer than every source and header"
            return out
    nvcc = _nvcc()
    flags = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
             "-Xptxas", "-v" if verbose else "-O3", "-I", _INCLUDE, "-I", _CSRC, "-Xcompiler", "-fPIC"] + [f"-D{d}" for d in defines]

    def run(cmd):
        proc = subprocess.run(cmd, capture_output=True, text=True)
        if proc.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + proc.stdout + proc.stderr)
        if verbose:
            print(proc.stderr)

    # one object per translation unit, cached next to the library (variants built elsewhere keep their own objects)
    objdir = os.path.join(os.path.dirname(out), "_obj" if out == library_path() else "_obj_" + os.path.basename(out))
    os.makedirs(objdir, exist_ok=True)
    objs = []
    for unit, headers in _UNITS.items():
        src = os.path.join(_CSRC, unit)
        obj = os.path.join(objdir, unit[:-3] + ".o")
        deps = [src] + [os.path.join(_CSRC, h) for h in headers] + _COMMON
        if force or defines or not os.path.exists(obj) or any(os.path.getmtime(d) > os.path.getmtime(obj) for d in deps):
            run([nvcc] + flags + ["-c", src, "-o", obj])
            LAST_BUILD["compiled_units"].append(unit)
        objs.append(obj)
    run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "--shared", "-o", out] + objs)
    LAST_BUILD["action"] = ("compiled " + ", ".join(LAST_BUILD["compiled_units"]) if LAST_BUILD["compiled_units"] else "relinked cached objects") + \
        " with nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo"
    return out
