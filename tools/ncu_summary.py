"""Text summary of an .ncu-rep (raw page): the metrics the roofline discussion uses, one block per captured launch, plus
the hottest SASS lines by stall samples.   python tools/ncu_summary.py file.ncu-rep > profiles/xyz.txt"""
import csv, io, re, subprocess, sys, collections
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.sum",
        "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]
seen = collections.Counter()
for r in rows[2:]:
    name = r[idx["Kernel Name"]]
    short = re.sub(r"\(.*", "", name).replace("void ", "").replace("leida::", "")
    seen[short] += 1
    print(f"=== {short}  (launch #{seen[short]} of this kernel in the capture; id {r[idx['ID']]})")
    for w in WANT:
        if w in idx and r[idx[w]] != "":
            print(f"  {w:86s} {r[idx[w]]} {units[idx[w]]}")
    print()
# hottest instructions of the first launch of each kernel
for short in seen:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + short.split("<")[0],
                          "--launch-count", "1"], capture_output=True, text=True).stdout
    lines = [r for r in csv.reader(io.StringIO(src)) if len(r) > 5 and r[0].startswith("0x")]
    if not lines:
        continue
    def I(x):
        try: return int(x)
        except ValueError: return 0
    uniq = {}
    for r in lines: uniq[r[0]] = r
    lines = list(uniq.values())
    tot = sum(I(r[2]) for r in lines) or 1
    by = collections.Counter()
    for r in lines:
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[1])
        by[m.group(2) if m else "?"] += I(r[2])
    print(f"--- {short}: warp-stall samples by opcode (total {tot}): " + ", ".join(f"{op} {100*c/tot:.1f}%" for op, c in by.most_common(10)))
    for r in sorted(lines, key=lambda r: -I(r[2]))[:8]:
        print(f"      {100*I(r[2])/tot:5.1f}%  executed {r[5]:>9s}  {r[1].strip()[:90]}")
    print()
