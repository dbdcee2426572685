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


def test_build_pipeline_lane_logic_with_stub_contexts(monkeypatch):
    """device.BuildPipeline without a device: contexts replaced by stubs.  Submissions go to the lanes round
    robin, a lane runs its submissions in order, consume runs before batch and result are freed, without consume
    the caller gets (batch, result) un-freed, resident batches are never freed by the pipeline, an exception of a
    build reaches the future and the lane keeps working."""
    import threading
    import time

    from make_prg_b200 import device

    log = []

    class Obj:
        def __init__(self, name):
            self.name, self.freed = name, False

        def free(self):
            self.freed = True
            log.append(("free", self.name))

    class StubContext:
        n = 0

        def __init__(self, dev=0):
            self.id = StubContext.n
            StubContext.n += 1
            self.workers = self.wait = None
            self.thread_ids = set()

        def set_workers(self, n):
            self.workers = n

        def set_wait_mode(self, mode):
            self.wait = mode

        def build_packed(self, tag, *a):
            self.thread_ids.add(threading.get_ident())
            if tag == "boom":
                raise device.MprgError(-2, "injected")
            time.sleep(0.01 if self.id == 0 else 0.0)  # lane 0 is the slow one
            log.append(("build", self.id, tag))
            return Obj(f"batch{tag}"), Obj(f"res{tag}")

        def build(self, batch, *a):
            log.append(("build_resident", self.id, batch.name))
            return Obj("res_" + batch.name)

        def launch_count(self):
            return 1

        def copy_stats(self, reset=False):
            return {"h2d_bytes": 10, "d2h_bytes": 1}

        def close(self):
            log.append(("close", self.id))

    monkeypatch.setattr(device, "Context", StubContext)
    monkeypatch.delenv("MPRG_LANE_WAIT", raising=False)
    pipe = device.BuildPipeline(0, depth=3)
    assert [c.workers for c in pipe.contexts] == [1, 1, 1] and {c.wait for c in pipe.contexts} == {"yield"}
    seen = []

    def consume(batch, res):
        assert not batch.freed and not res.freed
        seen.append(res.name)
        return res.name

    futs = [pipe.submit_packed(k, None, None, None, None, 5, 7, consume=consume) for k in range(9)]
    assert [f.result() for f in futs] == [f"res{k}" for k in range(9)]
    builds = [e for e in log if e[0] == "build"]
    assert sorted((lane, tag) for _, lane, tag in builds) == sorted((k % 3, k) for k in range(9))
    for lane in range(3):  # in order per lane
        assert [tag for _, l, tag in builds if l == lane] == [k for k in range(9) if k % 3 == lane]
    assert sum(1 for e in log if e[0] == "free") == 18
    assert all(len(c.thread_ids) == 1 for c in pipe.contexts)  # one host thread per lane
    # no consume: the caller owns both objects
    batch, res = pipe.submit_packed("x", None, None, None, None, 5, 7).result()
    assert not batch.freed and not res.freed
    # resident batches stay the caller's
    mine = Obj("mine")
    assert pipe.submit_resident(mine, 5, 7, consume=lambda b, r: (b is mine, r.name)).result() == (True, "res_mine")
    assert not mine.freed
    # a failing build: the error reaches the future, the lane goes on
    lane = pipe.next_lane
    bad = pipe.submit_packed("boom", None, None, None, None, 5, 7, consume=consume)
    with pytest.raises(device.MprgError):
        bad.result()
    for _ in range(2):
        pipe.submit_packed("skip", None, None, None, None, 5, 7, consume=consume).result()
    assert pipe.next_lane == lane
    assert pipe.submit_packed("after", None, None, None, None, 5, 7, consume=consume).result() == "resafter"
    assert pipe.launch_count() == 3 and pipe.copy_stats() == {"h2d_bytes": 30, "d2h_bytes": 3}
    pipe.close()
    assert sum(1 for e in log if e[0] == "close") == 3
    single = device.BuildPipeline(0, depth=1)
    assert single.contexts[0].wait == "spin"
    single.close()
