"""torchrun check: the NCCL multi-domain run equals the single-domain run.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/check_multi_gpu.py

Every rank routes its tributary domains (rank 0 also the mainstem, after the NCCL hand-off) with the blocking and
with the pipelined driver, and compares REACH_Q of the reaches it owns with a single-domain run of the whole
network on its own GPU.  Expected: bit-identical (same kernels, same operation order)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from mizuroute_b200 import synth
from mizuroute_b200.multi import DomainSet
from mizuroute_b200.network import RouteOptions, RouteParams
from mizuroute_b200.route import Router

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
# size / lakes / methods / step by environment: MR_CHECK_N=3000000 MR_CHECK_LAKES=10000 MR_CHECK_ROUTE=2 MR_CHECK_DT=86400 is C5
net = synth.conus_like(int(os.environ.get("MR_CHECK_N", "40000")), seed=3)
opts = RouteOptions(dt=float(os.environ.get("MR_CHECK_DT", "3600")), route_opt=os.environ.get("MR_CHECK_ROUTE", "12"), runoffMin=1e-15)
n_lakes = int(os.environ.get("MR_CHECK_LAKES", "0"))
if n_lakes:
    synth.add_lakes(net, n_lakes, np.random.default_rng(103))
    opts.is_lake_sim = True
    opts.LakeInputOption = 1
params = RouteParams()
K, B = int(os.environ.get("MR_CHECK_K", "24")), int(os.environ.get("MR_CHECK_B", "8"))
ro = synth.runoff_series(net, K, seed=11, dt=opts.dt)
single = Router(net, params, opts, device=local, max_batch=K).route_batch(ro)

worst = 0.0
for mode in ("blocking", "pipelined", "pipelined end-to-end"):
    dom = DomainSet(net, params, opts, B, rank, world, device=local)
    sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
    if dom.trib is not None:
        dom.trib.set_stream(sa.cuda_stream)
    if dom.main is not None:
        dom.main.set_stream(sb.cuda_stream if mode != "blocking" else sa.cuda_stream)
    if mode == "pipelined end-to-end":                     # pinned host buffers in and out, all batches in flight
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        nm = len(opts.route_opt)
        ins_t = [pin(ro[s:s + B][:, dom.trib_net.meta["hru_index"]]) for s in range(0, K, B)]
        outs_t = [torch.empty((nm, B, dom.trib_net.nRch), dtype=torch.float64).pin_memory() for _ in range(0, K, B)]
        ins_m = [pin(ro[s:s + B][:, dom.main_net.meta["hru_index"]]) if dom.main is not None else None for s in range(0, K, B)]
        outs_m = [torch.empty((nm, B, dom.main_net.nRch), dtype=torch.float64).pin_memory() if dom.main is not None else None for _ in range(0, K, B)]
        for i in range(len(ins_t)):
            dom.route_batch_pipelined(ins_t[i], outs_t[i], ins_m[i], outs_m[i], sa, sb)
        sa.wait_stream(sb)
        dom.wait()
        torch.cuda.synchronize()
        for i, s in enumerate(range(0, K, B)):
            assert np.array_equal(outs_t[i].numpy(), single[:, s:s + B][:, :, dom.dec.trib[rank]]), f"rank {rank} tributaries differ ({mode})"
            if dom.main is not None:
                keep = ~dom.main_net.meta["ghost_mask"]
                assert np.array_equal(outs_m[i].numpy()[:, :, keep], single[:, s:s + B][:, :, dom.main_net.meta["reach_index"][keep]]), f"mainstem differs ({mode})"
    for s in (range(0, K, B) if mode != "pipelined end-to-end" else []):
        dom.upload_runoff(ro[s:s + B])
        with torch.cuda.stream(sa):
            if mode == "blocking":
                dom.route_resident(B)
            else:
                dom.route_resident_pipelined(B, sa, sb)
                sa.wait_stream(sb)
                dom.wait()
        torch.cuda.synchronize()
        if dom.trib is not None:
            q = dom.trib.download_q(B)
            ref = single[:, s:s + B][:, :, dom.dec.trib[rank]]
            worst = max(worst, float(np.max(np.abs(q - ref) / np.maximum(np.abs(ref), 1e-300))))
            assert np.array_equal(q, ref), f"rank {rank} tributaries differ ({mode})"
        if dom.main is not None:
            q = dom.main.download_q(B)
            keep = ~dom.main_net.meta["ghost_mask"]
            ref = single[:, s:s + B][:, :, dom.main_net.meta["reach_index"][keep]]
            assert np.array_equal(q[:, :, keep], ref), f"mainstem differs ({mode})"
    dist.barrier()
    if rank == 0:
        print(f"{mode}: {net.nRch} reaches ({n_lakes} lakes), route_opt {opts.route_opt}, world {world}, mainstem {dom.dec.mainstem.size} reaches, {dom.dec.outlets.size} outlets handed over by NCCL: "
              f"REACH_Q bit-identical to the single-domain run on every rank")
dist.destroy_process_group()
