"""Top stall-sample SASS lines of the first kernel in an `ncu --page source --csv` export."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]; ci = {h: i for i, h in enumerate(hdr)}
body = []
for r in rows[2:]:
    if r and r[0] == "Kernel Name":
        break
    if len(r) >= len(hdr) - 2 and r[0].startswith("0x"):
        body.append(r)
samp = [int(r[ci["# Samples"]]) for r in body]
tot = sum(samp)
print("total samples", tot, "instrs", len(body))
top = sorted(range(len(body)), key=lambda i: -samp[i])[:ntop]
for i in sorted(top):
    r = body[i]
    print(i, samp[i], f"{100*samp[i]/tot:.1f}%", r[ci["Source"]].strip()[:80], "| long", r[ci["stall_long_sb"]], "wait", r[ci["stall_wait"]],
          "short", r[ci["stall_short_sb"]], "lg", r[ci["stall_lg"]], "noinst", r[ci["stall_no_inst"]], "exec", r[ci["Instructions Executed"]])
