"""world_size-2 gloo test of the multi-rank path bench.py uses: families are sharded by LPT over
ranks with no data-path collective; only the timing (MAX) and the cell-update count (SUM) are
reduced. The compute stand-in on CPU is the oracle (tests may call it; the product never does)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from repeatafterme_b200 import batch, synth
from repeatafterme_b200.abi import make_params


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _specs():
    specs = synth.config5_specs(5, 9)
    return [dict(s, n=max(8, s["n"] // 200), flank=max(40, s["flank"] // 40)) for s in specs]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tests import support
    ora = support.Oracle()
    specs = _specs()
    costs = [batch.family_cost(s["n"], s["params"]["bandwidth"], s["flank"]) for s in specs]
    mine = batch.lpt_shard(costs, world)[rank]
    cu, bps = 0.0, {}
    for i in mine:
        fam = synth.family_from_spec(specs[i])
        p = {k: v for k, v in fam.params.items() if k != "matrix"}
        sc = ora.scoring(fam.params["matrix"])
        for d in (1, 0):
            o = ora.extend(d, fam.cores, fam.sequence, sc, make_params(**p))
            cu += o["rows_processed"] * o["n_extendable"] * (2 * p["bandwidth"] + 1) * 4.0
            bps[(i, d)] = o["extended_bp"]
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)      # stand-in for the device time
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    c = torch.tensor([cu], dtype=torch.float64)
    dist.all_reduce(c, op=dist.ReduceOp.SUM)
    gathered = [None] * world
    dist.all_gather_object(gathered, (mine, bps))
    if rank == 0:
        torch.save(dict(tmax=float(t), cu=float(c), gathered=gathered), os.path.join(out_dir, "out.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_family_sharding(tmp_path, oracle):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    out = torch.load(os.path.join(str(tmp_path), "out.pt"), weights_only=False)
    assert out["tmax"] == 2.0
    specs = _specs()
    seen = sorted(i for mine, _ in out["gathered"] for i in mine)
    assert seen == list(range(len(specs)))          # every family exactly once
    # sharded results == single-process results
    cu = 0.0
    for mine, bps in out["gathered"]:
        for i in mine:
            fam = synth.family_from_spec(specs[i])
            p = {k: v for k, v in fam.params.items() if k != "matrix"}
            sc = oracle.scoring(fam.params["matrix"])
            for d in (1, 0):
                o = oracle.extend(d, fam.cores, fam.sequence, sc, make_params(**p))
                assert o["extended_bp"] == bps[(i, d)]
                cu += o["rows_processed"] * o["n_extendable"] * (2 * p["bandwidth"] + 1) * 4.0
    assert abs(cu - out["cu"]) < 1e-6 * max(cu, 1.0)
