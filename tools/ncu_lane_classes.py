"""Reading of an `ncu -i report.ncu-rep --page source --csv` export for one kernel: instruction mix per warp-step, share of
instructions / stall samples by active-lane class, stall reasons, hottest low-lane regions and `no_instruction` rows.
What profiles/r02_stageA_speed_of_light.md and profiles/r02_stm_ncu.md were written from.
usage: ncu_lane_classes.py source.csv [warp_steps]   (warp_steps: default = the most common execution count of an instruction)"""
import collections, csv, re, sys

rows = list(csv.reader(open(sys.argv[1])))
h = rows[1]
ix = {c: i for i, c in enumerate(h)}
data = [r for r in rows[2:] if len(r) > ix["# Samples"]]
first = data[0][ix["Address"]] if data else None      # some ncu versions list the kernel's instructions twice
again = [i for i, r in enumerate(data) if i and r[ix["Address"]] == first]
if again:
    data = data[:again[0]]
stall_cols = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]


def num(r, c):
    try:
        return int(r[ix[c]])
    except ValueError:
        return 0


cnt = collections.Counter(num(r, "Instructions Executed") for r in data if num(r, "Instructions Executed") > 0)
W = float(sys.argv[2]) if len(sys.argv) > 2 else float(cnt.most_common(1)[0][0])
ops, cls_inst, cls_samp, cls_stall, stalls = (collections.Counter(), collections.Counter(), collections.Counter(),
                                              collections.defaultdict(collections.Counter), collections.Counter())
tot_i = tot_t = tot_s = 0
for r in data:
    ie, te, s = num(r, "Instructions Executed"), num(r, "Thread Instructions Executed"), num(r, "# Samples")
    if ie == 0:
        continue
    src = re.sub(r"^@!?U?P\d+\s+", "", r[ix["Source"]].strip())
    ops[src.split()[0].split(".")[0]] += ie
    a = te / ie
    c = "<= 8" if a <= 8 else "9-20" if a <= 20 else "21-27" if a < 27.5 else ">= 27.5"
    cls_inst[c] += ie
    cls_samp[c] += s
    tot_i, tot_t, tot_s = tot_i + ie, tot_t + te, tot_s + s
    for sc in stall_cols:
        v = num(r, sc)
        cls_stall[c][sc] += v
        stalls[sc] += v
print(f"warp-steps {W:.0f}; warp instructions per warp-step {tot_i / W:.0f}; lanes per instruction {tot_t / tot_i:.2f}")
fp64 = sum(ops[o] for o in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"))
print(f"FP64-pipe instructions per warp-step {fp64 / W:.0f}")
print("opcode mix per warp-step:", ", ".join(f"{k} {v / W:.0f}" for k, v in ops.most_common(14)))
for c in (">= 27.5", "21-27", "9-20", "<= 8"):
    top = ", ".join(f"{k[6:]} {100 * v / max(tot_s, 1):.1f}" for k, v in cls_stall[c].most_common(4))
    print(f"lanes {c:8s}: {100 * cls_inst[c] / tot_i:5.1f} % of instructions, {100 * cls_samp[c] / max(tot_s, 1):5.1f} % of samples ({top})")
print("stall samples:", ", ".join(f"{k[6:]} {100 * v / max(tot_s, 1):.1f} %" for k, v in stalls.most_common(10)))
