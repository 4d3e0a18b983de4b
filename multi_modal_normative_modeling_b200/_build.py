"""In-tree nvcc build of libnmb.so for sm_100a (B200).  No torch headers are involved:
the library is a plain C-ABI shared object (include/nmb.h)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libnmb.so")
SOURCES = ["nmb_train.cu", "nmb_train_tcp.cu", "nmb_deviation.cu", "nmb_prologue.cu", "nmb_csv.cu", "nmb_api.cu"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: libnmb.so cannot be built (there is no CPU fallback)")
    return exe


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, extra_flags=None, out_name: str = None) -> str:
    """Compile every CUDA source for sm_100a and link lib/libnmb.so.  Returns its path.
    extra_flags / out_name: dev variants for A/B timing in one GPU session (tools load them through NMB_LIB)."""
    os.makedirs(LIB_DIR, exist_ok=True)
    obj_dir = os.path.join(HERE, "build" if not out_name else "build_" + out_name)
    lib_path = LIB_PATH if not out_name else os.path.join(LIB_DIR, "libnmb_%s.so" % out_name)
    flags = FLAGS + list(extra_flags or [])
    os.makedirs(obj_dir, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "nmb.h"))
    nvcc = _nvcc()
    objs, procs = [], []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(obj_dir, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + headers):
            cmd = [nvcc] + ARCH + flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose and out:
            sys.stderr.write(out)
        if p.returncode != 0:
            raise RuntimeError("nvcc failed: %s\n%s" % (" ".join(cmd), out))
    if force or procs or _stale(lib_path, objs):
        cmd = [nvcc] + ARCH + ["-shared", "-o", lib_path] + objs
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed: %s\n%s" % (" ".join(cmd), r.stdout))
    return lib_path


if __name__ == "__main__":
    variant = [a.split("=", 1)[1] for a in sys.argv if a.startswith("--variant=")]
    defs = [a for a in sys.argv if a.startswith("-D")]
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, extra_flags=defs, out_name=variant[0] if variant else None))
