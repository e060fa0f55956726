"""2+ ranks: DistributedTubeSection's peer exchange -- copy-engine form and kernel form (hb_peer_put) -- against the NCCL gather
on the same launch (same records on rank 0),
and the time of both.  Run: python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/gpu_probe_peer.py"""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import hiten_b200 as hb
from hiten_b200 import sharded, synodic, workloads as W

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n_total = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
cap = int(sys.argv[2]) if len(sys.argv) > 2 else 128          # a small step capacity forces reruns with the fused kernel
ics, mu = W.c5_batch(n_total * world, rank, world)
res = {}
for mode in ("nccl", "peer", "peer_kernel"):
    ds = {}
    for key in ("l1", "l2"):
        ds[key] = sharded.DistributedTubeSection(n_total * world // 2, mu, W.c5_grid(key), W.c5_section(key, mu),
                                                 forward=W.C5_TUBES[key]["forward"], flip=(0, 6), steps_capacity=cap,
                                                 exchange=mode.split("_")[0], device_put=(mode == "peer_kernel"))
    y0 = {key: torch.from_numpy(np.ascontiguousarray(ics[key].T)).cuda() for key in ds}
    out = None
    for it in range(4):
        dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
        for key in ds:
            ds[key].launch(y0[key])
        if mode != "nccl":
            for key in ds:
                assert ds[key].start_gather()
            out = {key: ds[key].finish_gather() for key in ds}
        else:
            out = {key: ds[key].gather_device() for key in ds}
        torch.cuda.synchronize(); dist.barrier(); dt = time.perf_counter() - t0
    local = torch.tensor([ds[key].runner.hit_count() for key in ds], dtype=torch.int64, device="cuda")
    dist.all_reduce(local)
    if rank == 0:
        assert [int(out[key][1].sum()) for key in ds] == local.tolist(), (mode, local.tolist())
        res[mode] = {key: ([h.clone() for h in out[key][0]], out[key][1].clone(), [y.clone() for y in out[key][2]]) for key in ds}
        print(mode, "ms per step", 1e3 * dt, "hits", {key: int(out[key][1].sum()) for key in ds})
if rank == 0:
    for mode, key in [(m, k) for m in ("peer", "peer_kernel") for k in ("l1", "l2")]:
        a, b = res["nccl"][key], res[mode][key]
        assert torch.equal(a[1].cpu(), b[1].cpu())
        for r in range(world):
            k = int(a[1][r])
            def ordered(buf):                      # records are appended through an atomic counter: order by (traj, seq)
                rec = buf[: 9 * k].view(k, 9)
                return rec[torch.argsort(rec[:, 0].view(torch.int64) * 65536 + rec[:, 1].view(torch.int64))]
            assert torch.equal(ordered(a[0][r]), ordered(b[0][r])), (key, r)
            nl = b[2][r].shape[1]
            assert torch.equal(a[2][r][:, :nl], b[2][r]), (key, r)
    print("peer exchange == nccl gather: OK (copy-engine form and hb_peer_put form)")
dist.destroy_process_group()
