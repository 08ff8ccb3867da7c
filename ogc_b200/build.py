"""Build libogc_b200.so (hand-written sm_100a CUDA behind a C ABI) in-tree.

    python -m ogc_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU.  The .so lands in ogc_b200/csrc/ (git-ignored, but shipped to
the GPU box by gpurun).
"""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(CSRC, "libogc_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
         "-Xcompiler", "-fPIC,-fvisibility=hidden", "--expt-relaxed-constexpr"] + os.environ.get("OGC_NVCC_FLAGS", "").split()


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _stale():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(HERE, "..", "include", "*.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return OUT
    objs = []
    procs = []
    objdir = os.path.join(CSRC, "_obj")
    os.makedirs(objdir, exist_ok=True)
    for src in sources():
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(f"--- {os.path.basename(src)} ---\n{out}\n")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    subprocess.check_call([NVCC, "-shared", "-o", OUT] + objs + ["-Xcompiler", "-fPIC"])
    return OUT


PROBE_SRC = os.path.join(HERE, "..", "tests", "csrc")
PROBE_OUT = os.path.join(PROBE_SRC, "libogc_probe.so")


def build_probe(force: bool = False) -> str:
    """Test-only diagnostics (tests/csrc/*.cu: the tcgen05 descriptor probe, the MMA issue-rate probe) -> their own
    library, so that nothing test-only is linked into libogc_b200.so."""
    srcs = sorted(glob.glob(os.path.join(PROBE_SRC, "*.cu")))
    deps = srcs + glob.glob(os.path.join(CSRC, "*.cuh"))
    if not force and os.path.exists(PROBE_OUT) and all(os.path.getmtime(d) <= os.path.getmtime(PROBE_OUT) for d in deps):
        return PROBE_OUT
    subprocess.check_call([NVCC] + FLAGS + ["-shared", "-o", PROBE_OUT] + srcs)
    return PROBE_OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
    print(build_probe(force="--force" in sys.argv))
