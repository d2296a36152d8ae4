"""What a running tensor pipe takes away from the other warps of the SM (one CTA, warp 0 issues MMAs, 8 warps run a
SIMT loop).   PYTHONPATH=. python tools/contention.py"""
import torch
from directvoxgo_b200 import ext

mma = {0: "no MMA", 1: "SS N=128", 2: "TS N=128", 3: "SS N=16", 4: "TS N=16"}
simt = {0: "FFMA chain x64", 1: "st.shared.v4 x4", 2: "ld.shared.v4 x4", 3: "ld.global.v4 x4", 4: "tcgen05.ld x16 x4"}
reps = {0: 400, 1: 2000, 2: 2000, 3: 400, 4: 400}
for sm in simt:
    base = None
    for mm in mma:
        n_mma = 0 if mm == 0 else (4000 if mm <= 2 else 16000)
        o = ext.tc_contention(mm, n_mma, sm, reps[sm])
        torch.cuda.synchronize()
        t_s, t_m = int(o[0]), int(o[1])
        base = base or t_s
        print("%-18s | %-9s: SIMT %8d cyc (%.2fx)   MMA %8d cyc (%s per MMA)" % (
            simt[sm], mma[mm], t_s, t_s / base, t_m, "%.1f" % (t_m / n_mma) if n_mma else "-"))
