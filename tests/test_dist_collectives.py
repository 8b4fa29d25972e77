"""CPU, world_size 2, gloo: the host-side collectives of mico_b200.mico (concat_all_gather, all_gather_with_grad) against
a fixture produced by the reference's own data/utils/distributed.py on 2 gloo ranks (oracle/make_golden.py gen_dist):
same gathered tensors, same gradient (all-reduce(SUM) of the stacked gradients, own slice -- NOT divided by world)."""
import os
import sys

import torch
import torch.multiprocessing as mp

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, REPO)
    import torch.distributed as dist
    from mico_b200.mico import all_gather_with_grad, concat_all_gather
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(100 + rank)
    x = torch.randn(3, 5, generator=g, requires_grad=True)
    ids = torch.randint(0, 50, (3, 4), generator=g)
    gathered = all_gather_with_grad(x)
    ids_all = concat_all_gather(ids)
    w = torch.arange(1, gathered.numel() + 1, dtype=torch.float32).view_as(gathered) * (rank + 1)
    (gathered * w).sum().backward()
    q.put((rank, gathered.detach(), ids_all, x.grad.clone()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_matches_reference(golden_dir):
    gold = torch.load(os.path.join(golden_dir, "dist_gather_2rank.pt"), weights_only=False)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, 29641, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=180) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
    for r in range(2):
        assert torch.equal(res[r][1], gold["gathered"][r])
        assert torch.equal(res[r][2], gold["ids_all"][r])
        assert torch.allclose(res[r][3], gold["x_grad"][r], rtol=1e-6, atol=1e-6)
