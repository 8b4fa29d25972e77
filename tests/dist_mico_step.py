"""Multi-GPU check of the MiCo retrieval step (ITC with all-gathered negatives, ITM with all_gather_with_grad) --
run under torchrun, one rank per GPU over NCCL:

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 tests/dist_mico_step.py

Every rank runs 'ret%tv' on its own batch; gradients are SUMmed across ranks like the reference loop
(data/utils/pipeline.py:93-99).  Rank 0 then evaluates the CPU oracle on the simulated world (oracle/mico.py) and
compares per-rank losses (1e-3 relative) and the summed gradients."""
import os
import sys

import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))


def main():
    from test_gpu_mico import make_cfg, make_rank_batch, oracle_params
    from mico_b200.mico import MiCo, _AttrDict
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.manual_seed(0)
    model = MiCo.from_pretrained(make_cfg(), {})
    p = oracle_params(model) if rank == 0 else None
    model = model.cuda().train()
    b = 4
    g = torch.Generator().manual_seed(55)
    negs = [(torch.randint(0, world * b, (b,), generator=g), torch.randint(0, world * b, (b,), generator=g))
            for _ in range(world)]
    for k in range(world):      # a negative may not be the sample itself
        for t in negs[k]:
            own = torch.arange(k * b, k * b + b)
            t[t == own] = (t[t == own] + 1) % (world * b)
    r = make_rank_batch(rank, b=b)
    batch = dict(vision_pixels=r["pixels"].cuda(),
                 caption_tokens=_AttrDict(input_ids=r["ids"].cuda(), attention_mask=r["att"].cuda()),
                 itm_neg_cond_tv=negs[rank][0], itm_neg_text_tv=negs[rank][1])
    out = model(batch, "ret%tv", compute_loss=True)
    sum(out.values()).backward()
    losses = torch.stack([out["loss_itc"].detach(), out["loss_itm"].detach()])
    all_losses = [torch.empty_like(losses) for _ in range(world)]
    dist.all_gather(all_losses, losses)
    names, grads = [], []
    for k, v in model.named_parameters():
        if v.grad is not None:
            names.append(k)
            dist.all_reduce(v.grad)          # SUM, no divide (pipeline.py:93-99)
            grads.append(v.grad)
    ok = True
    if rank == 0:
        from oracle import eva_vit as OV
        from oracle import mico as OM
        ranks = [dict(make_rank_batch(k, b=b), neg_c=negs[k][0], neg_t=negs[k][1]) for k in range(world)]
        ref = OM.retrieval_caption_step(p, ranks, OV.vit_cfg(width=176, depth=2, heads=2, mlp=352), layers=2, heads=2,
                                        itm_ratio=0.1, task="ret%tv")
        sum(sum(d.values()) for d in ref).backward()
        for k in range(world):
            for j, nm in enumerate(("loss_itc", "loss_itm")):
                a, e = all_losses[k][j].item(), ref[k][nm].item()
                print(f"rank {k} {nm}: {a:.6f} vs oracle {e:.6f}")
                ok &= abs(a - e) <= 1e-3 * abs(e)
        errs = []
        for nm, gr in zip(names, grads):
            rg = p[nm].grad
            if rg is None or rg.norm().item() < 1e-7 or nm.endswith("self.key.bias") or nm.endswith("decoder.weight"):
                continue
            errs.append((((gr.cpu() - rg).norm() / rg.norm()).item(), nm))
        errs.sort(reverse=True)
        print("worst summed-gradient errors:", [(f"{e:.3e}", n) for e, n in errs[:4]], "median", f"{errs[len(errs) // 2][0]:.3e}")
        ok &= errs[0][0] < 1e-1 and errs[len(errs) // 2][0] < 3e-2
        print("DIST_MICO_STEP", "PASS" if ok else "FAIL", f"world={world}")
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
