"""Turn an `ncu --set full` report of the bench command into the per-stage DRAM traffic table bench.py prints as
`roofline.traffic` (dram__bytes_read.sum + dram__bytes_write.sum per launch), keyed by the commit it was taken at.

    ncu --set full --clock-control none --import-source on -k regex:"mlp_|march_|k0_|sweep_kernel" -s <warm-up launches> \
        -c <N> -o gpurun_out/r02_full python bench.py --steps 2 --warmup 3 --ramp-s 0 --no-cpu-baseline --no-ref-gpu \
        --no-extras --no-render          (on the GPU box, after the same command exited 0 without ncu)
    python tools/ncu_traffic.py gpurun_out/r02_full.ncu-rep profiles/r02_ncu_traffic.json profiles/r02_ncu_kernels.md
"""
import csv
import io
import json
import subprocess
import sys

STAGE_OF = {"march_fwd_kernel": "march_fwd", "k0_gather_kernel": "march_fwd", "k0_gather_tiles_kernel": "march_fwd", "mlp_fwd_kernel": "mlp_fwd",
            "mlp_bwd_kernel": "mlp_bwd", "k0_scatter_kernel": "march_bwd", "march_bwd_kernel": "march_bwd", "sweep_rows_kernel": "sweep",
            "sweep_kernel": "sweep"}
COLS = {"gpu__time_duration.sum": "time_us", "dram__bytes_read.sum": "dram_rd", "dram__bytes_write.sum": "dram_wr",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed": "tensor_pct",
        "launch__registers_per_thread": "regs", "launch__grid_size": "grid", "launch__block_size": "block",
        "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
        "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
        "lts__t_sector_hit_rate.pct": "l2_hit_pct", "smsp__inst_executed.sum": "warp_insts",
        "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed": "l1_lsu_wavefront_pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "smem_lsu_wavefronts",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum": "smem_tc_wavefronts",
        "sm__cycles_elapsed.max": "sm_cycles"}
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e3, "us": 1.0, "ns": 1e-3, "s": 1e6}


def main(rep, out_json, out_md):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    per = {}
    for r in rows[2:]:
        name = r[ix["Kernel Name"]]
        short = name.split("(")[0].split("::")[-1].replace("void ", "").strip()
        base = short.split("<")[0]
        rec = per.setdefault(short, {"launches": 0})
        rec["launches"] += 1
        for col, key in COLS.items():
            if col in ix and r[ix[col]] not in ("", "n/a"):
                v = float(r[ix[col]].replace(",", "")) * UNIT.get(units[ix[col]], 1.0)
                rec[key] = rec.get(key, 0.0) + v
        rec["stage"] = STAGE_OF.get(base)
    for rec in per.values():
        for k in list(rec):
            if k not in ("launches", "stage"):
                rec[k] /= rec["launches"]
    stages = {}
    for rec in per.values():
        if rec["stage"]:
            stages[rec["stage"]] = stages.get(rec["stage"], 0.0) + rec.get("dram_rd", 0) + rec.get("dram_wr", 0)
    commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    json.dump({"commit": commit, "workload": "cfg2: 160^3, 8192 rays, near random-init state (ncu capture of bench.py --steps 2 --warmup 3)",
               "what": "dram__bytes_read.sum + dram__bytes_write.sum per launch, summed over the kernels of each bench stage",
               "kernels": stages, "per_kernel": per}, open(out_json, "w"), indent=1)
    with open(out_md, "w") as f:
        f.write("# ncu --set full, per-kernel averages (%s, commit %s)\n\n" % (rep, commit))
        f.write("| kernel | launches | time us | dram rd MB | dram wr MB | dram % | sm % | tensor pipe % | issue active % | warps active % | L2 hit % | regs | grid x block | L1 LSU wavefronts % | smem wavefronts / SM / cycle (LSU + tensor) |\n")
        f.write("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|\n")
        for k, r in sorted(per.items(), key=lambda kv: -kv[1].get("time_us", 0)):
            cyc = max(r.get("sm_cycles", 0), 1.0) * 148
            f.write("| `%s` | %d | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %d | %d x %d | %.1f | %.2f + %.2f |\n" % (
                k, r["launches"], r.get("time_us", 0), r.get("dram_rd", 0) / 1e6, r.get("dram_wr", 0) / 1e6, r.get("dram_pct", 0),
                r.get("sm_pct", 0), r.get("tensor_pct", 0), r.get("issue_active_pct", 0), r.get("warps_active_pct", 0),
                r.get("l2_hit_pct", 0), r.get("regs", 0), r.get("grid", 0), r.get("block", 0),
                r.get("l1_lsu_wavefront_pct", 0), r.get("smem_lsu_wavefronts", 0) / cyc, r.get("smem_tc_wavefronts", 0) / cyc))
    print(json.dumps(stages))


if __name__ == "__main__":
    main(*sys.argv[1:4])
