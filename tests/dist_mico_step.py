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
    ok2 = reference_two_rank_check(rank, world)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if (ok and ok2) else 1)


def reference_two_rank_check(rank, world):
    """World size 2 only: replay tests/golden/losses_2rank.pt -- the reference's own forward_ret / forward_cap run on two
    gloo ranks -- through MiCo.forward over NCCL: per-rank losses (1e-3 relative) and the gradient that all_gather_with_grad
    returns to every rank's fusion input (3e-2 rel-L2)."""
    if world != 2:
        return True
    import torch.nn.functional as F
    from test_gpu_mico import make_cfg
    from mico_b200.mico import MiCo, _AttrDict
    gd = os.path.join(REPO, "tests", "golden")
    one = torch.load(os.path.join(gd, "losses_tiny.pt"), weights_only=False)
    r = torch.load(os.path.join(gd, "losses_2rank.pt"), weights_only=False)["ranks"][rank]
    torch.manual_seed(0)
    cfg = make_cfg()
    cfg["contra_dim"] = 32
    model = MiCo.from_pretrained(cfg, {})
    model.load_state_dict(one["state_dict"], strict=False)
    model = model.cuda().train()
    raw_t, raw_v, cond = (r[k].clone().cuda().requires_grad_(True) for k in ("raw_t", "raw_v", "cond"))
    batch = dict(feat_t=F.normalize(raw_t, dim=-1), feat_v=F.normalize(raw_v, dim=-1), condition_feats_v=cond,
                 caption_tokens=_AttrDict(input_ids=r["ids"].cuda(), attention_mask=r["att"].cuda()),
                 cap_input_ids=r["cap_ids"].cuda(), cap_labels=r["cap_labels"].cuda(),
                 itm_neg_cond_tv=r["neg_c"], itm_neg_text_tv=r["neg_t"])
    out = model(batch, "ret%tv_cap%tv", compute_loss=True)
    sum(out.values()).backward()
    ok = True
    for k in ("loss_itc", "loss_itm", "loss_cap"):
        a, e = out[k].item(), r[k].item()
        print(f"[reference 2-rank fixture] rank {rank} {k}: {a:.6f} vs {e:.6f}")
        ok &= abs(a - e) <= 1e-3 * max(1.0, abs(e))
    for name, got, want in (("d_cond", cond.grad, r["d_cond"]), ("d_raw_t", raw_t.grad, r["d_raw_t"]), ("d_raw_v", raw_v.grad, r["d_raw_v"])):
        e = ((got.cpu() - want).norm() / want.norm()).item()
        print(f"[reference 2-rank fixture] rank {rank} {name} rel-L2 {e:.3e}")
        ok &= e < 3e-2
    flag = torch.tensor([1.0 if ok else 0.0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("DIST_REFERENCE_FIXTURE", "PASS" if flag.item() > 0 else "FAIL")
    return flag.item() > 0


if __name__ == "__main__":
    main()
