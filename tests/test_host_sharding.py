"""N>1 host logic on CPU: loci are sharded over ranks with the LPT partition (no data-path
collective), every rank builds its shard independently, results are gathered and merged in input
order, and the merged output is identical to the single-rank output.  World size 2 over gloo; the
per-shard compute is the oracle here (no GPU in this container) -- what is under test is the
sharding / gather / ordering code that `from_msa --gpus N` and `bench.py --gpus N` rely on."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

REPO = Path(__file__).resolve().parent.parent


def test_lpt_partition_properties():
    from make_prg_b200.engine import lpt_partition

    rng = np.random.default_rng(0)
    costs = rng.integers(1, 1000, 57).tolist()
    for n in (1, 2, 4, 8):
        parts = lpt_partition(costs, n)
        assert sorted(i for p in parts for i in p) == list(range(57))
        loads = [sum(costs[i] for i in p) for p in parts]
        assert max(loads) - min(loads) <= max(costs)  # LPT bound on imbalance
    assert lpt_partition([], 3) == [[], [], []]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, str(REPO))
    sys.path.insert(0, str(REPO / "oracle"))
    import torch.distributed as dist

    import make_prg_oracle as mo
    from make_prg_b200 import synth
    from make_prg_b200.engine import lpt_partition

    dist.init_process_group("gloo", rank=rank, world_size=world)
    shapes = [(12, 60), (30, 200), (8, 40), (25, 150), (10, 90), (16, 120), (6, 30)]
    mats = [synth.synth_msa(r, c, 500 + i, var_frac=0.08, n_dels=2) for i, (r, c) in enumerate(shapes)]
    parts = lpt_partition([r * c for r, c in shapes], world)
    mine = {i: mo.build_prg_from_matrix([f"s{k}" for k in range(mats[i].shape[0])], mats[i], 5, 7)[0]
            for i in parts[rank]}
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    if rank == 0:
        merged = {}
        for g in gathered:
            merged.update(g)
        ordered = [merged[i] for i in range(len(mats))]
        single = [mo.build_prg_from_matrix([f"s{k}" for k in range(m.shape[0])], m, 5, 7)[0] for m in mats]
        (Path(out_dir) / "ok").write_text("1" if ordered == single else "0")
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_shard_gather_is_order_invariant(tmp_path):
    import torch.multiprocessing as mp

    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok").read_text() == "1"


def test_default_lanes_follow_the_core_share(monkeypatch):
    """device.default_lanes: builds in flight per GPU = this process's share of the host cores minus one, between
    3 and 6; MPRG_BUILD_LANES overrides (no device needed)."""
    import os

    from make_prg_b200 import device

    monkeypatch.delenv("MPRG_BUILD_LANES", raising=False)
    for cores, ranks, want in ((16, 1, 6), (24, 2, 6), (32, 4, 6), (32, 8, 3), (8, 2, 3), (2, 1, 3), (12, 2, 5)):
        monkeypatch.setattr(os, "sched_getaffinity", lambda pid, n=cores: set(range(n)))
        monkeypatch.setenv("LOCAL_WORLD_SIZE", str(ranks))
        assert device.default_lanes() == want, (cores, ranks)
    monkeypatch.setenv("MPRG_BUILD_LANES", "9")
    assert device.default_lanes() == 9
