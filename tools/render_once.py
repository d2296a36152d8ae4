"""One dense 800x800 render of the bench's random-init model (for `ncu --metrics gpu__time_duration.sum` launch lists
and CUDA-event stage timing):  PYTHONPATH=. python tools/render_once.py [n_frames]"""
import sys
import torch
import bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
dev = torch.device("cuda", 0)
model, rk, cfg = bench.build_problem(160, dev)
r = bench.render_metric(model, rk, dev, n, 65536, "dense random-init")
print(r)
