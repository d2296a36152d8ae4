"""TMEM -> register read bandwidth per SM (tcgen05.ld), by warp count and instruction shape.
PYTHONPATH=. python tools/ldtm_rate.py"""
import torch
from directvoxgo_b200 import ext

for mode, name in ((0, "32x32b.x16 x4"), (1, "32x32b.x32 x2"), (2, "16x256b.x4 x2")):
    for nw in (1, 4, 8, 16, 32):
        o = ext.tc_ldtm_rate(nw, 200, mode)
        torch.cuda.synchronize()
        cyc, byt = int(o[0]), int(o[1])
        print("%-14s warps=%2d: %8d cycles for %9d bytes -> %6.1f B/cycle" % (name, nw, cyc, byt, byt / cyc))
