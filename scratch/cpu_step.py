import sys, time; sys.path.insert(0,'.')
import torch
import bench
bench.N_POINT = int(sys.argv[1]) if len(sys.argv)>1 else 1024
t=time.time(); print(bench.run_cpu_port(1, 0, 1, 8), time.time()-t)
