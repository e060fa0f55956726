import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import hiten_b200 as hb
s = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "stm_family.npz"))
x0 = torch.from_numpy(np.ascontiguousarray(np.tile(s["x0"], (512, 1)).T)).cuda()
T = torch.from_numpy(np.tile(s["period"], 512)).cuda()
for _ in range(2):
    hb.cr3bp_stm(x0, float(s["mu"]), 0.0, tf_per_traj=T, integ=hb.make_integ(arith="parity"))
torch.cuda.synchronize()
