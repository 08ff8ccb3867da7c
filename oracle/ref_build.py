"""TEST INFRASTRUCTURE -- build the REFERENCE's own pointnet2 CUDA extension into oracle/_ref/.

    python oracle/ref_build.py [--sass]

The sources are compiled from where they lie under /root/reference/pointnet2/src: a build-time
copy goes to a temp dir (the reference tree is read-only and four .cpp files need a 2-line include
shim because torch >= 1.11 no longer ships THC/THC.h); the .cu kernels are compiled UNCHANGED with
the reference's own flags (nvcc -O2, pointnet2/setup.py:19-20) plus the sm_100a gencode.  Only the
built module lands in the repo (oracle/_ref/pointnet2_cuda_ref.so, git-ignored, shipped by gpurun).
No reference source is copied into the repository.

The result is (a) the GPU oracle the CUDA kernels and the CPU restatement are checked against on a
B200 (tests/test_gpu_vs_refext.py) and (b) the "reference CUDA extension" timing arm of bench.py.
`--sass` prints the FADD/FMUL/FFMA sequence of the distance expression for the record
(rounding order).
"""
import glob
import os
import re
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/pointnet2/src"
OUT_DIR = os.path.join(HERE, "_ref")
NAME = "pointnet2_cuda_ref"


def build(verbose=False):
    if not os.path.isdir(REF_SRC):
        return None
    os.makedirs(OUT_DIR, exist_ok=True)
    out = os.path.join(OUT_DIR, NAME + ".so")
    if os.path.exists(out):
        return out
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
    from torch.utils import cpp_extension
    tmp = tempfile.mkdtemp(prefix="ogc_refbuild_")
    src = os.path.join(tmp, "src")
    shutil.copytree(REF_SRC, src)
    for p in glob.glob(os.path.join(src, "*.cpp")):
        os.chmod(p, 0o644)
        text = open(p).read()
        text = text.replace("#include <THC/THC.h>", "#include <ATen/cuda/CUDAContext.h>")
        text = re.sub(r"^extern THCState \*state;\s*$", "", text, flags=re.M)
        open(p, "w").write(text)
    files = sorted(glob.glob(os.path.join(src, "*.cpp")) + glob.glob(os.path.join(src, "*.cu")))
    build_dir = os.path.join(tmp, "build")
    os.makedirs(build_dir)
    cpp_extension.load(
        name=NAME, sources=files, build_directory=build_dir, verbose=verbose, is_python_module=False,
        extra_cflags=["-g"],
        extra_cuda_cflags=["-O2", "-gencode", "arch=compute_100a,code=sm_100a"])
    shutil.copy(os.path.join(build_dir, NAME + ".so"), out)
    shutil.rmtree(tmp, ignore_errors=True)
    return out


def sass():
    so = build()
    txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    for fn in ["ball_query_kernel_fast", "three_nn_kernel_fast", "furthest_point_sampling_kernelILj1024"]:
        m = re.search(r"Function : (\S*%s\S*)(.*?)(?=Function :|\Z)" % fn, txt, flags=re.S)
        if not m:
            continue
        ops = re.findall(r"\b(FADD|FMUL|FFMA)\b[^;]*;", m.group(2))
        print(m.group(1), " ".join(ops[:12]))


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
    if "--sass" in sys.argv:
        sass()
