"""Probe: stage times of hb_cr3bp_section2 (sparse records) on one C5 tube, for comparing build variants of the
propagation kernel (HITEN_B200_LIB=tools/variants/...).  usage: gpu_probe_stageA.py [n_total] [tube] [records]"""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import hiten_b200 as hb
from hiten_b200 import synodic, propagate as P, workloads as W

n_total = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
key = sys.argv[2] if len(sys.argv) > 2 else "l1"
records = sys.argv[3] if len(sys.argv) > 3 else "near"
ics, mu = W.c5_batch(n_total)
integ = hb.make_integ(arith="parity")
x = ics[key]; n = len(x); te = W.c5_grid(key)
fwd = -1 if key == "l1" else 1
y0 = torch.from_numpy(np.ascontiguousarray(x.T)).cuda()
run = synodic.TubeSectionRunner(n, mu, te, W.c5_section(key, mu), forward=fwd, flip=(0, 6), integ=integ,
                                steps_capacity=128 if records == "near" else 192, records=records)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
run.set_stage_events(ev)
best = None
for _ in range(4):
    run.launch(y0)
    torch.cuda.synchronize()
    ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(4)]
    best = ms if best is None or ms[0] < best[0] else best
print(json.dumps({"lib": os.environ.get("HITEN_B200_LIB", "default"), "tube": key, "n": n, "records": records,
                  "stage_ms": best, "hits": run.hit_count(), "recs": int(run.records_written().sum().item())}))
