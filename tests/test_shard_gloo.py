"""CPU suite: the N>1 host logic (batch sharding, shared-parameter gradient all-reduce, result
gather) on a world_size-2 gloo group.  The per-shard solve itself is stood in for by the oracle:
what is under test is the sharding/exchange code of b200qp/dist.py, not the solver."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, nb, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "diff-qp-mpc_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from b200qp import dist as D
        from oracle import qp_oracle as O
        torch.set_num_threads(1)
        Q, p, G, h, A, b = O.random_qp(nb, 6, 8, 0, seed=3)
        Gs = G[0].clone()  # G shared across the batch; keep every problem feasible
        z0 = torch.randn(nb, 6, dtype=torch.float64, generator=torch.Generator().manual_seed(1))
        h = (Gs @ z0.T).T + 0.5
        params = (Q, p, Gs, h, A, b)
        mine = D.shard_params(params, nb, rank, world)
        sl = D.shard_slice(nb, rank, world)
        assert mine[0].shape[0] == sl.stop - sl.start and mine[2].dim() == 2
        fwd = O.qp_forward(*[t.clone() for t in mine])
        gr = O.qp_backward(fwd, *mine, torch.ones_like(fwd["zhat"]))
        dG_global = D.allreduce_shared_grad(gr["dG"], sl.stop - sl.start)
        zhat = D.gather_batch(fwd["zhat"])
        # data-parallel policy training (deqmpc/train.py:165-175): one flat-bucket all-reduce averages the gradients
        lin = torch.nn.Sequential(torch.nn.Linear(3, 4), torch.nn.LayerNorm(4))
        for i, prm in enumerate(lin.parameters()):
            prm.grad = torch.full_like(prm, float(rank + 1) * (i + 1))
        D.allreduce_gradients(list(lin.parameters()))
        for i, prm in enumerate(lin.parameters()):
            assert torch.allclose(prm.grad, torch.full_like(prm, 1.5 * (i + 1))), "allreduce_gradients"
        if rank == 0:
            # the same thing computed by one process, shard by shard
            zs, num = [], 0
            for r in range(world):
                pr = D.shard_params(params, nb, r, world)
                f = O.qp_forward(*[t.clone() for t in pr])
                g = O.qp_backward(f, *pr, torch.ones_like(f["zhat"]))
                n_r = f["zhat"].shape[0]
                zs.append(f["zhat"])
                num = num + g["dG"] * n_r
            ok = torch.equal(zhat, torch.cat(zs)) and torch.allclose(dG_global, num / nb, rtol=1e-12, atol=1e-14)
            q.put(("ok" if ok else "mismatch", tuple(zhat.shape)))
    except Exception as ex:  # pragma: no cover
        q.put(("error", repr(ex)))
    finally:
        dist.destroy_process_group()


def test_shard_slice_is_a_partition():
    from b200qp.dist import shard_slice
    for nb in (1, 7, 8, 128, 1001):
        for world in (1, 2, 4, 8):
            idx = []
            for r in range(world):
                s = shard_slice(nb, r, world)
                idx.extend(range(s.start, s.stop))
            assert idx == list(range(nb))
    with pytest.raises(ValueError):
        shard_slice(8, 2, 2)


def test_two_rank_gloo_shard_and_exchange():
    world, nb = 2, 7  # ragged: 4 + 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nb, q)) for r in range(world)]
    for p in procs:
        p.start()
    status, detail = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
    assert status == "ok", detail
    assert detail == (nb, 6)
