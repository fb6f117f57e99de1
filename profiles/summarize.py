"""Distil gpurun_out/*.ncu-rep + launches.csv into tracked summaries under profiles/ (run on the CPU box)."""
import collections, csv, io, json, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
TAG = sys.argv[1] if len(sys.argv) > 1 else "r02"
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lts__t_bytes.sum", "smsp__cycles_active.avg",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio", "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
        "smsp__average_warp_latency_issue_stalled_barrier.ratio", "smsp__average_warp_latency_issue_stalled_mio_throttle.ratio",
        "smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio", "smsp__average_warp_latency_issue_stalled_not_selected.ratio",
        "smsp__average_warp_latency_issue_stalled_wait.ratio", "smsp__average_warp_latency_issue_stalled_dispatch_stall.ratio",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]

def raw(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr = rows[0]
    out = []
    for r in rows[2:]:
        out.append(dict(zip(hdr, r)))
    return hdr, rows[1], out

lines = ["# ncu summaries (%s)" % TAG, "", "Captured with `profiles/run_profiles.sh` under gpurun on one B200 (`--set full --clock-control none`);",
         "per-launch times are cold-cache / serialised — compare shares, not absolutes.", ""]
traffic = {}
for name in ("prof_loss", "prof_rollout", "prof_fwd", "prof_adam", "prof_env", "prof_k1", "prof_gae", "prof_pack", "prof_gather", "prof_dqn"):
    rep = os.path.join(OUT, name + ".ncu-rep")
    if not os.path.exists(rep):
        continue
    hdr, units, recs = raw(rep)
    for rec in recs:
        kn = re.sub(r"\(.*", "", rec.get("Kernel Name", "?"))
        lines.append("## %s — `%s`" % (name, kn))
        lines.append("")
        lines.append("| metric | value | unit |")
        lines.append("|---|---|---|")
        ui = dict(zip(hdr, units))
        for k in KEYS:
            if k in rec:
                lines.append("| %s | %s | %s |" % (k, rec[k], ui.get(k, "")))
        try:
            MULT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}   # ncu picks a unit per metric: convert each one separately
            t = (float(rec["dram__bytes_read.sum"].replace(",", "")) * MULT.get(ui.get("dram__bytes_read.sum", "byte"), 1)
                 + float(rec["dram__bytes_write.sum"].replace(",", "")) * MULT.get(ui.get("dram__bytes_write.sum", "byte"), 1))
            mult = 1
            mk = re.search(r"(ac_loss_grad_tc_kernel|ac_loss_grad_kernel|rollout_tc_kernel|forward_tc_kernel|forward_kernel|reduce_clip_adam_kernel|env_step_kernel|scan_series_fastest|dqn_loss_grad_kernel|pack_records_kernel|sample_gather_kernel)", kn)
            traffic[mk.group(1) if mk else kn] = t * mult
            lines.append("| dram traffic (read+write) | %.3f | MB |" % (t * mult / 1e6))
        except Exception as e:
            pass
        lines.append("")
        break  # first captured launch is enough
# launch lists
for fname, title in (("launches.csv", "python bench.py --steps 2 --warmup 1 (eager launches)"), ("launches_c5.csv", "python bench.py --config c5 (DQN update loop)")):
    lp = os.path.join(OUT, fname)
    if not os.path.exists(lp):
        continue
    rows = list(csv.reader(open(lp, errors="replace")))
    hdr = None; data = []
    for r in rows:
        if "Kernel Name" in r: hdr = r; continue
        if hdr and len(r) == len(hdr): data.append(dict(zip(hdr, r)))
    tot = collections.defaultdict(float); cnt = collections.Counter()
    for d in data:
        nm = re.sub(r"\(.*", "", d["Kernel Name"])
        head, _, targs = nm.partition("<")
        nm = head.split("::")[-1] + ("<" + re.sub(r"(\(anonymous namespace\)|envdev)::", "", targs)[:48] if targs else "")
        v = float(d["Metric Value"].replace(",", "")); u = d["Metric Unit"]
        v *= {"ns": 1, "nsecond": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6}.get(u, 1)
        tot[nm] += v; cnt[nm] += 1
    T = sum(tot.values())
    lines += ["## launch list (`ncu --metrics gpu__time_duration.sum`, %s)" % title, "",
              "| kernel | launches | total ms | avg us | share |", "|---|---|---|---|---|"]
    for k, v in sorted(tot.items(), key=lambda x: -x[1]):
        lines.append("| %s | %d | %.3f | %.2f | %.1f%% |" % (k, cnt[k], v / 1e6, v / cnt[k] / 1e3, 100 * v / T))
    lines.append("")
open(os.path.join(ROOT, "profiles", TAG + "_ncu_summary.md"), "w").write("\n".join(lines))
if traffic:
    json.dump(traffic, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
print("\n".join(lines))
