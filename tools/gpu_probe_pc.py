"""Probe: hb_cr3bp_section3 (records through shared memory) vs hb_cr3bp_section2 (records through HBM) on the C5 tubes:
identical hits / end states, timings.  usage: gpu_probe_pc.py [n_total]"""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import hiten_b200 as hb
from hiten_b200 import synodic, workloads as W

n_total = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
arith = sys.argv[2] if len(sys.argv) > 2 else "parity"
ics, mu = W.c5_batch(n_total)
integ = hb.make_integ(arith=arith)
for key in ("l1", "l2"):
    x = ics[key]; n = len(x); te = W.c5_grid(key)
    kw = dict(forward=W.C5_TUBES[key]["forward"], flip=(0, 6), integ=integ)
    y0 = torch.from_numpy(np.ascontiguousarray(x.T)).cuda()
    res = {}
    for name, opts in (("section3", dict(pool_records=8)), ("section2", dict(steps_capacity=192))):
        run = synodic.TubeSectionRunner(n, mu, te, W.c5_section(key, mu), **kw, **opts)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4 if name == "section3" else 5)]
        run.set_stage_events(ev)
        for _ in range(2):
            run.launch(y0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run.launch(y0); e1.record(); torch.cuda.synchronize()
        stage = [ev[i].elapsed_time(ev[i + 1]) for i in range(len(ev) - 1)]
        ovf = int((run.status == 4).sum().item())
        h = run.sorted_hits()
        res[name] = (h, run.yf.clone(), run.nacc.clone(), run.nrej.clone())
        print(json.dumps({"tube": key, "path": name, "n": n, "ms": e0.elapsed_time(e1), "stage_ms": stage,
                          "overflowed": ovf, "hits": len(h.times), "status_ok": bool((run.status == 0).all().item())}))
        del run
        torch.cuda.empty_cache()
    a, b = res["section3"], res["section2"]
    same = (np.array_equal(a[0].trajectory_indices, b[0].trajectory_indices) and np.array_equal(a[0].times, b[0].times)
            and np.array_equal(a[0].states, b[0].states) and torch.equal(a[1], b[1]) and torch.equal(a[2], b[2])
            and torch.equal(a[3], b[3]))
    print(json.dumps({"tube": key, "section3_equals_section2_bit_for_bit": bool(same)}))
