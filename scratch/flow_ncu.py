"""One eager FlowStep3D training step (configs[2]: 16 pairs x 2048 points, iters 4) for ncu.

    ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,\
dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,\
sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active \
        --clock-control none --csv --log-file gpurun_out/flow_launches.csv python scratch/flow_ncu.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["OGC_FLOW_EAGER"] = "1"
import torch
import bench

step, _ = bench.flow_step_fn(2048, 16, 4, torch.device("cuda", 0))
for i in range(2):
    step(i, host_inputs=False)
torch.cuda.synchronize()
torch.cuda.profiler.start()
step(2, host_inputs=False)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
