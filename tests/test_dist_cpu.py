"""World-size-2 gloo tests of the multi-GPU host logic (masp_b200/sharding.py):
proof-level sharding and the base-split MSM with gathered partials.  Device
work runs in the host-compiled emulation; the collectives are real."""
import os
import sys

import pytest
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from util import load_emu, ib, rand_scalars, assignment, oracle_proofs
        from masp_b200 import sharding, synthetic as syn
        from oracle import c_oracle as co
        pv = load_emu()
        # base-split MSM
        n = 101
        logs = syn.fr_uniform(syn.MASTER_SEED, 12, n)
        bases = co.g1_gen_mul(syn.limbs_to_bytes(logs), n)
        sc = ib(rand_scalars(n, 4))
        got = sharding.split_msm_g1(pv, bases, sc, n)
        ok_msm = got == co.msm_g1(bases, sc, n)
        # proof sharding, uneven split (5 proofs over 2 ranks)
        sh = syn.micro_shape()
        kb = co.params_from_logs(syn.key_logs(sh))
        dens = sh.densities()
        P = pv.Parameters.read(kb, dens)
        ws = [syn.witness(sh, i, co.fr_mul) for i in range(5)]
        proofs = sharding.prove_sharded(pv, P, [assignment(pv, w) for w in ws], [w["r"] for w in ws],
                                        [w["s"] for w in ws])
        ok_prove = proofs == oracle_proofs(co, kb, sh, dens, ws)
        lo, hi = sharding.shard_range(5, rank, world)
        q.put((rank, ok_msm, ok_prove, (lo, hi)))
    finally:
        dist.destroy_process_group()


@pytest.mark.slow
def test_world_size_2_gloo():
    from masp_b200.build import build_emu
    from oracle import c_oracle
    build_emu()
    c_oracle.build()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert res == [(0, True, True, (0, 3)), (1, True, True, (3, 5))]


def test_shard_helpers():
    from masp_b200 import sharding
    for n in (0, 1, 7, 256, 1024):
        for world in (1, 2, 4, 8):
            cover = []
            for r in range(world):
                lo, hi = sharding.shard_range(n, r, world)
                cover += list(range(lo, hi))
                assert hi - lo in (n // world, n // world + 1)
            assert cover == list(range(n))
    q = sharding.longest_first_queue({"output": 2, "spend": 2, "convert": 1})
    assert [c for c, _ in q] == ["spend", "spend", "convert", "output", "output"]
