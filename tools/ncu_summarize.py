"""Per-kernel summary of an .ncu-rep (first launch of each kernel) + launch-list aggregation; writes markdown."""
import csv, subprocess, sys, collections

rep, launches = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__sass_inst_executed_op_local_st.sum",
        "smsp__sass_inst_executed_op_local_ld.sum"]
seen = set()
out = []
for r in rows[2:]:
    name = [r[i] for i, h in enumerate(hdr) if h == "Kernel Name"][0]
    if name in seen:
        continue
    seen.add(name)
    out.append("### " + name)
    for w in want:
        for i, h in enumerate(hdr):
            if h == w:
                out.append(f"- {w}: {r[i]} {rows[1][i]}")
    st = [(float(r[i].replace(",", "")), h) for i, h in enumerate(hdr)
          if "smsp__average_warps_issue_stalled" in h and "_per_issue_active" in h and "not_issued" not in h]
    st.sort(reverse=True)
    out.append("- stall cycles per issue: " + ", ".join(
        f"{h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')}={v:.2f}" for v, h in st[:7]))
    out.append("")
rows = list(csv.reader(open(launches)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h2 = rows[hi]; ci = {h: i for i, h in enumerate(h2)}
agg = collections.OrderedDict()
for r in rows[hi + 2:]:
    if len(r) < len(h2):
        continue
    a = agg.setdefault(r[ci["Kernel Name"]][:90], [0, 0.0])
    a[0] += 1; a[1] += float(r[ci["Metric Value"]].replace(",", "")) / 1e6
tot = sum(v[1] for k, v in agg.items() if "k_dfma" not in k)
out.append("## launch list (ncu --metrics gpu__time_duration.sum), aggregated")
out.append("| kernel | launches | total ms | share (without k_dfma) |\n|---|---:|---:|---:|")
for k, (c, ms) in agg.items():
    out.append(f"| `{k}` | {c} | {ms:.3f} | {'-' if 'k_dfma' in k else f'{100 * ms / tot:.1f} %'} |")
print("\n".join(out))
