"""Dense 800x800 render time against the chunk size and the number of streams the chunks are pipelined over.
PYTHONPATH=. python tools/render_sweep.py"""
import torch
import bench
from directvoxgo_b200 import synthetic as syn
from directvoxgo_b200.fused import FusedRenderer

dev = torch.device("cuda", 0)
model, rk, cfg = bench.build_problem(160, dev)
H = W = syn.BLENDER["H"]
K = syn.intrinsics(H, W)
poses = syn.random_poses(3, seed=4242)
for chunk in (32768, 65536, 131072):
    for streams in (1, 2, 3, 4):
        r = FusedRenderer(model, rk)
        r.render_view(H, W, K, poses[0], chunk=chunk, streams=streams)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for c2w in poses[1:]:
            img = r.render_view(H, W, K, c2w, chunk=chunk, streams=streams)["rgb_marched"]
        e1.record()
        torch.cuda.synchronize()
        print("chunk %6d streams %d: %.2f ms/frame  (mean rgb %.6f)" % (chunk, streams, e0.elapsed_time(e1) / 2, float(img.mean())), flush=True)
        del r
        torch.cuda.empty_cache()
