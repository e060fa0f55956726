"""Probe: propagate-only (hb_cr3bp_propagate) and section2 stage times on one C5 tube; used to compare kernel build
variants (HITEN_B200_LIB=tools/variants/...).  usage: gpu_probe_prop.py [n_total] [arith]"""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import hiten_b200 as hb
from hiten_b200 import synodic, propagate as P, workloads as W

n_total = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
arith = sys.argv[2] if len(sys.argv) > 2 else "parity"
ics, mu = W.c5_batch(n_total)
integ = hb.make_integ(arith=arith)
key = "l1"
x = ics[key]; n = len(x); te = W.c5_grid(key)
y0 = torch.from_numpy(np.ascontiguousarray(x.T)).cuda()
ws = P.workspace(y0.device)
best = 1e9
for _ in range(4):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); r = hb.cr3bp_propagate(y0, mu, float(te[-1]), forward=-1, flip=(0, 6), integ=integ, ws=ws); e1.record()
    torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
steps = int((r.n_acc.sum() + r.n_rej.sum()).item())
chk = float(r.yf.double().sum().item())
print(json.dumps({"what": "propagate_only", "n": n, "ms": best, "steps_per_s": steps / best * 1e3, "checksum": chk}))
if os.environ.get("PROBE_PROP_ONLY") == "1":      # large batches: the section pipeline's step scratch would not fit
    sys.exit(0)
run = synodic.TubeSectionRunner(n, mu, te, W.c5_section(key, mu), forward=-1, flip=(0, 6), integ=integ, steps_capacity=192)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
run.set_stage_events(ev)
for _ in range(3):
    run.launch(y0)
torch.cuda.synchronize()
print(json.dumps({"what": "section2", "stage_ms": [ev[i].elapsed_time(ev[i + 1]) for i in range(4)], "hits": run.hit_count()}))
