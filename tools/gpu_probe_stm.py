"""STM (42-state) kernel throughput vs batch size, parity and fast."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import hiten_b200 as hb

s = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "stm_family.npz"))
for arith in ("parity", "fast"):
    integ = hb.make_integ(arith=arith)
    for rep in (16, 128, 1024, 4096):
        x0 = torch.from_numpy(np.ascontiguousarray(np.tile(s["x0"], (rep, 1)).T)).cuda()
        T = torch.from_numpy(np.tile(s["period"], rep)).cuda()
        for _ in range(2):
            r = hb.cr3bp_stm(x0, float(s["mu"]), 0.0, tf_per_traj=T, integ=integ)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            r = hb.cr3bp_stm(x0, float(s["mu"]), 0.0, tf_per_traj=T, integ=integ)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        st = int((r.n_acc.sum() + r.n_rej.sum()).item())
        print(f"{arith} n={x0.shape[1]:7d}: {ms:8.2f} ms  {st / ms * 1e3:.3e} steps/s  {st * 9500 / ms * 1e3 / 1e12:.2f} TFLOP/s", flush=True)
