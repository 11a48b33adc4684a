"""Turns the ncu artefacts under gpurun_out/ into the tracked summaries under profiles/.

    python tools/summarize_ncu.py r01

* gpurun_out/launches_<R>.csv      -> profiles/launches_<R>_summary.csv (per kernel: launches, total, share)
* gpurun_out/<name>_<R>.ncu-rep    -> profiles/ncu_<name>_<R>.csv (selected metrics per captured launch)
                                      and profiles/ncu_summary_<R>.json (read by bench.py for `traffic`)
"""
import collections, csv, glob, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = sys.argv[1] if len(sys.argv) > 1 else "r01"
OUT = os.path.join(ROOT, "profiles")
os.makedirs(OUT, exist_ok=True)

KEEP = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__cluster_dim_x",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__ops_path_tensor_src_fp64.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fp64.sum",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum", "l1tex__m_l1tex2xbar_write_bytes_mem_global_op_tma_st.sum",
    "l1tex__m_xbar2l1tex_read_bytes_mem_dshared.sum", "smsp__pcsamp_warps_issue_stalled_barrier",
    "smsp__pcsamp_warps_issue_stalled_wait", "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle",
    "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_short_scoreboard",
    "smsp__pcsamp_warps_issue_stalled_membar", "smsp__pcsamp_warps_issue_stalled_selected",
]


def short(name):
    n = name.split("(")[0]
    for tok in ("void ", "unnamed>::", "t4b::dla::", "<unnamed>::"):
        n = n.replace(tok, "")
    return n.strip()


def launches():
    src = os.path.join(ROOT, "gpurun_out", f"launches_{R}.csv")
    if not os.path.exists(src):
        return
    rows = list(csv.reader(open(src, errors="replace")))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h = rows[hi]
    kn, mv = h.index("Kernel Name"), h.index("Metric Value")
    agg = collections.OrderedDict()
    total = 0.0
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        try:
            v = float(r[mv].replace(",", ""))
        except ValueError:
            continue
        k = short(r[kn])
        a = agg.setdefault(k, [0, 0.0, 1e30, 0.0])
        a[0] += 1; a[1] += v; a[2] = min(a[2], v); a[3] = max(a[3], v)
        total += v
    dst = os.path.join(OUT, f"launches_{R}_summary.csv")
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "launches", "total_ms", "share", "avg_us", "min_us", "max_us"])
        for k, (n, t, mn, mx) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            w.writerow([k, n, round(t / 1e6, 3), round(t / total, 4), round(t / n / 1e3, 2), round(mn / 1e3, 2), round(mx / 1e3, 2)])
        w.writerow(["# command: tools/ncu_capture.sh (ncu --metrics gpu__time_duration.sum --clock-control none [--launch-skip S] -c N --csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline); %d launches captured" % sum(a[0] for a in agg.values())])
        w.writerow(["# times under ncu are serialised / cold-cache: compare the SHARE with bench.py's kernel_profile_ms"])
    print("wrote", dst)


def reports():
    summary = {}
    for rep in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", f"*_{R}.ncu-rep"))):
        name = os.path.basename(rep)[: -len(f"_{R}.ncu-rep")]
        p = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True)
        rows = list(csv.reader(p.stdout.splitlines()))
        if len(rows) < 3:
            continue
        h, units = rows[0], rows[1]
        cols = [i for i, c in enumerate(h) if c in KEEP or c == "Kernel Name"]
        dst = os.path.join(OUT, f"ncu_{name}_{R}.csv")
        with open(dst, "w", newline="") as f:
            w = csv.writer(f)
            w.writerow(["metric", "unit"] + [short(r[h.index("Kernel Name")])[:48] for r in rows[2:]])
            for i in cols:
                if h[i] == "Kernel Name":
                    continue
                w.writerow([h[i], units[i]] + [r[i] for r in rows[2:]])
        print("wrote", dst)
        for r in rows[2:]:
            k = short(r[h.index("Kernel Name")])
            def val(metric):
                if metric not in h:
                    return None
                i = h.index(metric)
                try:
                    v = float(r[i].replace(",", ""))
                except ValueError:
                    return None
                u = units[i].lower()
                mult = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
                return v * mult
            rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
            summary.setdefault(name, []).append({
                "kernel": k, "dram_bytes": (rd or 0) + (wr or 0),
                "fp64_tensor_pct": val("sm__ops_path_tensor_src_fp64.avg.pct_of_peak_sustained_elapsed"),
                "duration": val("gpu__time_duration.sum")})
    if summary:
        dst = os.path.join(OUT, f"ncu_summary_{R}.json")
        json.dump(summary, open(dst, "w"), indent=1)
        print("wrote", dst)


launches()
reports()
