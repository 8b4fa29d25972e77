"""Per-shape device time of the omni-modal bench step (developer / evidence tool): wraps the ops.* entry points with a CUDA
event pair keyed by (op, shape, flags), runs bench.py's own omni step and prints the table sorted by time.

    python scripts/omni_shapes.py [--batch 64] [--config omni] [--top 60]
"""
import argparse
import collections
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch

import bench
from mico_b200 import dp, ops, optim
from mico_b200.audioprocessor import AudioProcessor
from mico_b200.mico import MiCo, _AttrDict


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--config", default="omni")
    ap.add_argument("--top", type=int, default=70)
    ap.add_argument("--light-blocks", type=int, default=bench.DEFAULT_LIGHT_BLOCKS)
    ap.add_argument("--qkv-blocks", type=int, default=bench.DEFAULT_QKV_BLOCKS)
    ap.add_argument("--attn-blocks", type=int, default=bench.DEFAULT_ATTN_BLOCKS)
    ap.add_argument("--out", default="")
    ap.add_argument("--plain", type=int, default=0, help="run this many plain steps and exit (target for an ncu launch list)")
    a = ap.parse_args()
    sp = bench.SPECS[a.config]
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    torch.manual_seed(0)
    cfg = bench.model_cfg(ckpt=sp["ckpt"], tower=sp["tower"])
    with torch.device(dev):
        model = MiCo.from_pretrained(cfg, {})
    model = model.to(dev).train()
    tower = getattr(model.vision_encoder, "visual", None)
    if tower is not None:
        tower.ckpt_light_blocks, tower.ckpt_qkv_blocks, tower.ckpt_attn_blocks = a.light_blocks, a.qkv_blocks, a.attn_blocks
    flat = dp.FlatGrads(model)
    no_decay = ("bias", "LayerNorm.bias", "LayerNorm.weight")
    named = list(model.named_parameters())
    opt = optim.AdamW([dict(params=[p for k, p in named if not any(s in k for s in no_decay)], weight_decay=0.01),
                       dict(params=[p for k, p in named if any(s in k for s in no_decay)], weight_decay=0.0)],
                      lr=1e-6, betas=(0.9, 0.98))
    if tower is not None:
        opt.attach_bf16_sinks(tower)
    audio = AudioProcessor(melbins=224, target_length=224, sample_num=max(sp["n_a"], 1), training=True, device=dev)
    host = bench.host_batch(a.batch, 0, n_v=sp["n_v"], n_d=sp["n_d"], wave=sp["n_a"] > 0)
    res = {k: v.to(dev) for k, v in host.items() if k not in ("input_ids", "attention_mask") and v is not None}

    def step():
        flat.zero_grad()
        batch = dict(vision_pixels=res["vision_pixels"],
                     caption_tokens=_AttrDict(input_ids=host["input_ids"], attention_mask=host["attention_mask"]))
        if "depth_pixels" in res:
            batch["depth_pixels"] = res["depth_pixels"]
        if "audio_waveforms" in res:
            batch["audio_spectrograms"] = audio.batch(res["audio_waveforms"])
        out = model(batch, sp["task"], compute_loss=True)
        sum(out.values()).backward()
        flat.detach_unused()
        opt.step()

    if a.plain:
        from mico_b200 import _lib
        for i in range(a.plain):
            _lib.reset_launch_count()
            if i == a.plain - 1:
                torch.cuda.profiler.start()        # ncu --profile-from-start off: only the last step is captured
            step()
            torch.cuda.synchronize()
            if i == a.plain - 1:
                torch.cuda.profiler.stop()
            print(f"plain step {i}: {_lib.launch_count()} mico_b200 launches", flush=True)
        return
    for _ in range(2):
        step()
    torch.cuda.synchronize()

    recs = []

    def wrap(name, keyfn):
        orig = getattr(ops, name)

        def f(*args, **kw):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = orig(*args, **kw)
            e1.record()
            recs.append((name, keyfn(*args, **kw), e0, e1))
            return r
        setattr(ops, name, f)

    def gemm_key(a_, b_, **kw):
        K, M = a_.shape if kw.get("a_mn") else a_.shape[::-1]
        N = b_.shape[1] if kw.get("b_mn") else b_.shape[0]
        fl = [k for k in ("a_mn", "b_mn", "bias", "residual", "row_scale", "aux_out", "aux_in", "accumulate") if kw.get(k) is not None and kw.get(k) is not False]
        od = "f32" if (kw.get("out_dtype") == torch.float32 or (kw.get("out") is not None and kw["out"].dtype == torch.float32)) else "bf16"
        return (M, N, K, od, kw.get("act", 0), ",".join(fl)), 2.0 * M * N * K

    def attn_key(q, k, v, *rest, **kw):
        B, Sq, H, D = q.shape
        mask = kw.get("mask")
        return (B, H, Sq, k.shape[1], D, "mask%dd" % mask.dim() if mask is not None else "plain",
                "drop" if kw.get("dropout") else "", "kvidx%d" % k.shape[0] if kw.get("kv_index") is not None else ""), \
            4.0 * B * H * Sq * k.shape[1] * D

    def ln_key(x, *r, **kw):
        return tuple(x.shape), float(x.numel())

    wrap("gemm", gemm_key)
    wrap("attention_fwd", attn_key)
    wrap("attention_bwd", lambda q, k, v, o, lse, dout, scale, **kw: (lambda kk: (kk[0], kk[1] * 2.5))(attn_key(q, k, v, **kw)))
    wrap("layernorm_fwd", ln_key)
    wrap("layernorm_bwd", lambda dy, x, *r, **kw: ln_key(x))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    step()
    e1.record()
    torch.cuda.synchronize()
    total = e0.elapsed_time(e1)
    agg = collections.OrderedDict()
    for name, (key, work), s, e in recs:
        d = agg.setdefault((name, key), [0, 0.0, 0.0])
        d[0] += 1
        d[1] += s.elapsed_time(e)
        d[2] += work
    rows = sorted(agg.items(), key=lambda kv: -kv[1][1])
    fam = collections.defaultdict(float)
    for (name, key), (n, ms, w) in rows:
        fam[name] += ms
    lines = [f"step {total:.1f} ms (events recorded around every call: includes their overhead); by op: "
             + ", ".join(f"{k} {v:.1f}" for k, v in fam.items())]
    for (name, key), (n, ms, w) in rows[:a.top]:
        rate = w / (ms * 1e-3) / 1e12 if ms > 0 else 0
        unit = "TF/s" if name in ("gemm", "attention_fwd", "attention_bwd") else "Gelt/s x1000"
        lines.append(f"{ms:9.2f} ms  {n:5d} calls  {ms / n * 1e3:9.1f} us/call  {rate:8.1f} {unit}  {name} {key}")
    txt = "\n".join(lines)
    print(txt)
    if a.out:
        open(a.out, "w").write(txt + "\n")


if __name__ == "__main__":
    main()
