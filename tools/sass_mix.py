"""Instruction mix per kernel from `cuobjdump -sass` text: total, FP64-pipe (DADD/DMUL/DFMA/DSETP/DMNMX), MUFU, local/shared
memory, branches.  usage: sass_mix.py dump.sass [name-substring]"""
import re, sys, collections
txt = open(sys.argv[1]).read().split("Function : ")
pat = sys.argv[2] if len(sys.argv) > 2 else ""
for blk in txt[1:]:
    name = blk.split("\n", 1)[0].strip()
    if pat not in name:
        continue
    ops = collections.Counter()
    for m in re.finditer(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", blk):
        ops[m.group(1).split(".")[0]] += 1
    tot = sum(ops.values())
    f64 = sum(v for k, v in ops.items() if k in ("DADD", "DMUL", "DFMA", "DSETP", "DMNMX"))
    print(name[-70:], "total", tot, "fp64", f64, "DADD", ops["DADD"], "DMUL", ops["DMUL"], "DFMA", ops["DFMA"], "DSETP", ops["DSETP"],
          "MUFU", ops["MUFU"], "LDL", ops["LDL"], "STL", ops["STL"], "LDS", ops["LDS"], "STS", ops["STS"], "BRA", ops["BRA"],
          "MOV", ops["MOV"] + ops["IMAD"], "SEL", ops["SEL"] + ops["FSEL"], "CALL", ops["CALL"])
