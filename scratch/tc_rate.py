"""GPU: issue rate of tcgen05.mma kind::tf32 for the operand sources of the fused SA kernels (tests/csrc/tc_rate.cu).

    gpurun -- python scratch/tc_rate.py > gpurun_out/tc_rate.txt
"""
import ctypes
import os

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = ctypes.CDLL(os.path.join(ROOT, "tests", "csrc", "libogc_probe.so"))
NAMES = {0: "ts 3xTF32", 1: "ss 3xTF32", 2: "ts single", 3: "ss single", 4: "ts 3x, 2 accumulators", 5: "ts 3x, unrolled", 6: "ts 3x + concurrent LDTM", 7: "ts 3x + concurrent STTM", 8: "ts 3x + commit every 4 k-steps"}


def run(mode, n, k, reps, ctas):
    cyc = torch.zeros(ctas, dtype=torch.int64, device="cuda")
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    args = (mode, n, k, reps, ctas, ctypes.c_void_p(cyc.data_ptr()), st)
    rc = lib.ogc_tc_rate(*args)
    if rc != 0:
        return None
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    lib.ogc_tc_rate(*args)
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e)
    per = 3 if mode in (0, 1, 4, 5, 6, 7, 8) else 1
    n_mma = reps * (k // 8) * per
    flops = 2.0 * 128 * n * 8 * n_mma * ctas
    return float(cyc.float().mean()) / n_mma, ms, flops / (ms * 1e-3) / 1e12


def main():
    print("mode                      N    K  ctas  cycles/MMA  floor(N/2)  TFLOP/s(tf32, chip)")
    for ctas in (148,):
        for mode in (0, 8):
            for n in (64, 128):
                r = run(mode, n, 64, 400, ctas)
                if r is None:
                    continue
                print(f"{NAMES[mode]:24s} {n:4d} {64:4d} {ctas:5d}  {r[0]:10.1f}  {n / 2:10.1f}  {r[2]:8.1f}")


if __name__ == "__main__":
    main()
