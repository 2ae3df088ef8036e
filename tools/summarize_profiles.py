#!/usr/bin/env python3
"""Turn gpurun_out/ artefacts (ncu launch list CSV, ncu --set full report) into the small text
summaries committed under profiles/.   usage: summarize_profiles.py <round-tag>"""
import collections, csv, os, re, subprocess, sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out_dir = os.path.join(root, "profiles")
os.makedirs(out_dir, exist_ok=True)

def short(name):
    m = re.search(r"body_kernel(?:_strided)?<zk::(\w+)(<zk::(G\d)[^>]*>)?", name)
    if m:
        return m.group(1) + (f"<{m.group(3)}>" if m.group(3) else "")
    return re.sub(r"\(.*", "", name.replace("void ", ""))[:60]

lst = os.path.join(root, "gpurun_out", f"{tag}_launches.csv")
if os.path.exists(lst):
    rows = [r for r in csv.reader(l for l in open(lst) if not l.startswith("==")) if len(r) > 14 and r[0].isdigit()]
    agg = collections.OrderedDict()
    for r in rows:
        a = agg.setdefault(short(r[4]), [0, 0.0])
        a[0] += 1
        a[1] += float(r[14]) / 1e6
    total = sum(v[1] for v in agg.values())
    with open(os.path.join(out_dir, f"{tag}_ncu_launch_list_summary.txt"), "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none -c 900 python bench.py --steps 2 --warmup 3 --skip-cpu-baseline\n")
        f.write(f"# {len(rows)} launches, {total:.2f} ms summed (cold-cache, serialised: compare SHARES, not absolutes)\n")
        f.write(f"{'kernel':34s} {'launches':>8s} {'ms':>10s} {'share':>7s}\n")
        for k, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k:34s} {c:8d} {ms:10.3f} {100 * ms / total:6.1f}%\n")
        names = ("RecodeCount", "Scatter", "Accumulate", "FixupLevel", "BucketReduce", "PairSum", "Finish", "scan_block_sums",
                 "scan_top_level", "scan_apply", "bucket_reduce_kernel", "pair_sum_kernel", "row_sum_kernel", "BatchedAddRound", "PairCount",
                 "bucket_acc_kernel", "finish_kernel", "PoisonPartial", "BucketSizeCheck", "FixupDirect")
        msm = {k: v for k, v in agg.items() if any(nm in k for nm in names)}
        t2 = sum(v[1] for v in msm.values())
        f.write(f"\n# MSM pipeline kernels only ({t2:.2f} ms): share of one MSM step\n")
        for k, (c, ms) in sorted(msm.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k:34s} {c:8d} {ms:10.3f} {100 * ms / t2:6.1f}%\n")
    print("wrote launch list summary")

rep = os.path.join(root, "gpurun_out", f"{tag}_prof_acc.ncu-rep")
if not os.path.exists(rep):
    rep = os.path.join(root, "gpurun_out", f"{tag}_prof_hot.ncu-rep")      # round 2: hot kernels of one warm step
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
            "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
            "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "smsp__warps_eligible.avg.per_cycle_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
            "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "dram__bytes_read.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "smsp__pcsamp_warps_issue_stalled_wait", "smsp__pcsamp_warps_issue_stalled_long_scoreboard",
            "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle", "smsp__pcsamp_warps_issue_stalled_not_selected", "smsp__pcsamp_warps_issue_stalled_short_scoreboard",
            "smsp__pcsamp_warps_issue_stalled_no_instructions", "smsp__pcsamp_warps_issue_stalled_dispatch_stall", "smsp__pcsamp_sample_buffer_full"]
    with open(os.path.join(out_dir, f"{tag}_accumulate_ncu_full.txt"), "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k <hot kernels> -s <three warm steps> -c <one step's launches> python bench.py --steps 2 --warmup 3 --skip-cpu-baseline --g2-logn 0 --groth16-logn 0   (tools/ncu_capture.sh)\n")
        for r in data:
            f.write(f"\n== {r[hdr.index('Kernel Name')][:100]}\n")
            for k in keys:
                if k in hdr:
                    i = hdr.index(k)
                    f.write(f"{k:75s} {r[i]:>18s} {units[i]}\n")
    print("wrote accumulate summary")
    # per-launch DRAM traffic of the dominant kernel, consumed by bench.py's roofline.traffic
    import json
    def col(r, k):
        i = hdr.index(k)
        v = float(r[i])
        u = units[i].lower()
        return v * {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1}.get(u, 1)
    # the capture holds the accumulation group of ONE MSM step (batched-affine rounds + Accumulate): bytes are summed
    group = [r for r in data if "BatchedAddRound" in r[hdr.index("Kernel Name")] or "bucket_acc" in r[hdr.index("Kernel Name")]
             or "Accumulate" in r[hdr.index("Kernel Name")]]
    rd = sum(col(r, "dram__bytes_read.sum") for r in group)
    wr = sum(col(r, "dram__bytes_write.sum") for r in group)
    data = group
    n = int(os.environ.get("TRAFFIC_N", str(1 << 20)))
    json.dump({"kernel": "bucket accumulation group (BatchedAddRound<G1> rounds + AccumulateBuckets<G1>), one MSM step", "launches": len(data), "n": n, "n_gpus": 1, "dram_bytes_read": round(rd), "dram_bytes_write": round(wr),
               "dram_bytes_total": round(rd + wr), "algorithmic_bytes": n * 128, "source": f"profiles/{tag}_accumulate_ncu_full.txt",
               "note": "first round: random gathers of 96-byte points from the precomputed tables (x in the forward pass, x and y on the way back) plus prefix scratch and 96 B of output per addition; later rounds stream"},
              open(os.path.join(out_dir, "accumulate_traffic.json"), "w"), indent=1)
    print("wrote accumulate_traffic.json")
