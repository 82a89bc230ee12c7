#!/usr/bin/env python
"""Turn the round's ncu captures (gpurun_out/<tag>_*.ncu-rep, <tag>_launches.csv) into the tracked summaries under
profiles/: <tag>_launches.md (per-kernel share of the bench step), <tag>_kernels.md (key metrics per captured
kernel) and <tag>_ellipse_full_traffic.json (read by bench.py for roofline.traffic)."""
import collections, csv, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TAG = sys.argv[1] if len(sys.argv) > 1 else "r01"
GO, PR = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
METRICS = [
 "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
 "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
 "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
 "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
 "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
 "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
 "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
 "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
 "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.avg.per_second",
 "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
 "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
 "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
 "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
 "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
 "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
 "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
]


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        res.append({h: (r[i], units[i]) for i, h in enumerate(hdr)})
    return res


def num(v):
    return float(v.replace(",", ""))


def to_bytes(v, u):
    return num(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]


def launches():
    path = os.path.join(GO, f"{TAG}_launches.csv")
    if not os.path.exists(path):
        return
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        name = r[ik].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        t = num(r[iv]) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[r[iu]]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += t
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(PR, f"{TAG}_launches.md"), "w") as f:
        f.write(f"# {TAG}: launch list of `python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --months 120 "
                "--month-steps 1 --no-kriging-1deg --no-stochastic --no-ellipse-fit --no-notebook --no-fp64-peaks`\n\n`ncu --metrics gpu__time_duration.sum --clock-control none -c 900` (per-launch times are "
                "cold-cache and serialised: read the SHARES).  Includes warm-up launches: 3 + 2 covariance builds and the 64-row "
                "probe of the builder, 2 + 1 month batches of 120.\n\n"
                "| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
        for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k[:90]}` | {c} | {t:.3f} | {100 * t / tot:.1f}% |\n")
        gemm = sum(t for k, (c, t) in agg.items() if any(x in k for x in ("trsm_", "potrf_colupdate", "potrf_panel", "potrf_update")))
        month = sum(t for k, (c, t) in agg.items() if not k.startswith(("ellipse_", "matern_")))
        tile = sum(t for k, (c, t) in agg.items() if k.startswith("ellipse_cov_tile_kernel<3"))
        cov = sum(t for k, (c, t) in agg.items() if k.startswith(("ellipse_", "matern_")))
        f.write("\nPer bench step the captured launches are: one covariance build = 1 tile kernel + prepare; one batch of 120 "
                "months = 13 solve steps, 12 column updates, 13 diagonal blocks, 12 panels, 13 coefficient strips, gather, "
                "reduce, finalize.\n"
                f"Share of the covariance step held by the tile kernel: {100 * tile / cov:.1f} %; share of the month step held "
                f"by the DMMA GEMM kernels (solve steps + column updates + panels + strips): {100 * gemm / month:.0f} %.\n")


def kernels():
    lines = [f"# {TAG}: `ncu --set full --clock-control none --import-source on` captures (one launch each)\n"]
    traffic = None
    for fn in sorted(os.listdir(GO)):
        if not (fn.startswith(TAG + "_") and fn.endswith(".ncu-rep")):
            continue
        for k in raw(os.path.join(GO, fn)):
            lines.append(f"\n## {fn}\n\nkernel: `{k['Kernel Name'][0][:140]}`\n\n| metric | value | unit |\n|---|---:|---|")
            for m in METRICS:
                if m in k:
                    lines.append(f"| {m} | {k[m][0]} | {k[m][1]} |")
            if "ellipse_full_nu2.2" in fn:
                w = to_bytes(*k["dram__bytes_write.sum"]) + to_bytes(*k["dram__bytes_read.sum"])
                traffic = {"nu": 2.2, "kernel": k["Kernel Name"][0][:120], "dram_bytes_per_launch": w,
                           "algorithmic_bytes_per_launch": 8 * 64800 * 64800,
                           "kernel_ms_under_ncu": num(k["gpu__time_duration.sum"][0]),
                           "fp64_pipe_pct_of_peak": num(k["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"][0]) / 100,
                           "issue_active_pct": num(k["smsp__issue_active.avg.pct_of_peak_sustained_active"][0]) / 100,
                           "source": f"gpurun_out/{fn} (ncu --set full, one launch of the full 1-degree build)"}
    if len(lines) == 1:
        return                              # no capture of this tag under gpurun_out/: keep the tracked summary
    with open(os.path.join(PR, f"{TAG}_kernels.md"), "w") as f:
        f.write("\n".join(lines) + "\n")
    if traffic:
        with open(os.path.join(PR, f"{TAG}_ellipse_full_traffic.json"), "w") as f:
            json.dump(traffic, f, indent=1)


if __name__ == "__main__":
    os.makedirs(PR, exist_ok=True)
    launches()
    kernels()
