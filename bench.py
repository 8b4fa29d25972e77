#!/usr/bin/env python
"""bench.py -- MiCo hot-path benchmark on B200 (contract in the task statement).

Workloads
  --config omni (default; BASELINE.json configs[4], one rank's slice): the full omni-modal pretraining step
      model(batch, "ret%tv%ta%tva%td_cap%tv%ta%tva") -> sum(losses).backward() -> gradient SUM over ranks -> fused AdamW
      per rank 64 samples x (video 8 frames + audio 3 spectrogram slices computed from 10 s waveforms by the fbank kernel +
      depth 1 frame) through the ViT-g/14 tower, 128 text tokens through the BERT-base text / fusion encoder, ITC + ITM +
      caption losses with the NCCL feature gathers at N > 1 (data/model/vast.py:317-512, data/utils/pipeline.py:35-111).
  --config vitg (BASELINE.json configs[1]): ViT-g/14 image-only fwd+bwd, bs 64 per GPU (round 1's headline).
  --config imgtext (configs[2]): image + text ITC / ITM / caption step, bs 32 per GPU.
  --config trimodal (configs[3]): Swin-B video + log-mel audio + text step, bs 4 per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config omni|vitg] [--batch B]

N > 1 is launched by torchrun, one rank per GPU; ranks process independent batches (weak scaling).
Output: ONE JSON line on rank 0 (contract keys + "roofline", "cpu_baseline", "e2e", "clocks", "selfcheck").
`--impl reference` times the reference's CPU eager path restated in oracle/ (fp32, all host threads) on a bounded sample.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

T_IMG = 257                    # tokens per 224x224 frame through ViT-g/14
NS_TOKENS = dict(image=257, video=1568, text=128)     # BASELINE.json north_star accounting
VIT_G = dict(img_size=224, patch_size=14, embed_dim=1408, depth=40, num_heads=16, mlp_ratio=4.3637, qkv_bias=True,
             drop_path_rate=0.4, num_classes=1024, use_mean_pooling=False)
# algorithmic FLOPs per frame, forward: 40 blocks x (2*257*1408*(4224+1408+2*6144) + 4*257^2*1408)
# + patch embed 2*256*588*1408; backward = 2x forward.
FWD_FLOPS_PER_FRAME = 40 * (2 * 257 * 1408 * (4224 + 1408 + 2 * 6144) + 4 * 257 * 257 * 1408) + 2 * 256 * 588 * 1408
OMNI_TASK = "ret%tv%ta%tva%td_cap%tv%ta%tva"
# tower checkpoint levels, counted from the last block backwards (mico_b200/eva_vit.py:_launch_forward): every block keeps
# its attention output + log-sum-exp, the last DEFAULT_QKV_BLOCKS + DEFAULT_LIGHT_BLOCKS also qkv, the last
# DEFAULT_LIGHT_BLOCKS also x1
DEFAULT_LIGHT_BLOCKS = int(os.environ.get("MICO_BENCH_LIGHT_BLOCKS", "0"))
DEFAULT_QKV_BLOCKS = int(os.environ.get("MICO_BENCH_QKV_BLOCKS", "14"))
DEFAULT_ATTN_BLOCKS = int(os.environ.get("MICO_BENCH_ATTN_BLOCKS", "-1"))
N_V, N_A, N_D, S_TXT = 8, 3, 1, 128
WAVE_SAMPLES = 160000          # 10 s at 16 kHz
# MiCo.forward workloads: BASELINE.json configs[4] (the metric's own), configs[2] and configs[3] (builder-run evidence lines)
SPECS = {
    "omni": dict(task=OMNI_TASK, batch=64, n_v=N_V, n_a=N_A, n_d=N_D, tower="evaclip01_giant", ckpt=True,
                 frame_tokens=T_IMG, frame_flops=None,
                 workload="full omni-modal (video 8f + audio 3 slices + depth + text S=128) ViT-g/14 + BERT-base pretraining step, "
                          "bs 64 per GPU (BASELINE configs[4]: 512 global on 8 GPUs), task " + OMNI_TASK),
    "imgtext": dict(task="ret%tv_cap%tv", batch=32, n_v=1, n_a=0, n_d=0, tower="evaclip01_giant", ckpt=False,
                    frame_tokens=T_IMG, frame_flops=None,
                    workload="image + text contrastive / matching / caption step (EVA-CLIP ViT-g image tower + BERT text), bs 32 per "
                             "GPU (BASELINE configs[2]: 256 global on 8 GPUs, NCCL feature all-gather), task ret%tv_cap%tv"),
    "trimodal": dict(task="ret%tv%ta%tva_cap%tv%ta%tva", batch=4, n_v=8, n_a=3, n_d=0, tower="swin_base_22k", ckpt=False,
                     frame_tokens=49, frame_flops=30.9e9,
                     workload="video (Swin-B, 8 frames) + audio (10 s waveform -> log-mel, 3 slices) + text tri-modal step, bs 4 per GPU "
                              "(BASELINE configs[3]: 32 global on 8 GPUs), task ret%tv%ta%tva_cap%tv%ta%tva"),
}
METRIC = {"omni": "omni-modal pretrain tokens/sec @ ViT-g/14",
          "imgtext": "omni-modal pretrain tokens/sec @ ViT-g/14 (image + text step, bs 32/GPU)",
          "trimodal": "omni-modal pretrain tokens/sec (Swin-B video + audio + text step, bs 4/GPU)",
          "vitg": "omni-modal pretrain tokens/sec @ ViT-g/14 (image-only fwd+bwd, bs 64/GPU)"}
WORKLOAD = {"omni": SPECS["omni"]["workload"], "imgtext": SPECS["imgtext"]["workload"], "trimodal": SPECS["trimodal"]["workload"],
            "vitg": "ViT-g/14 image-only fwd+bwd, bs=64 synthetic 224x224 per GPU (BASELINE configs[1])"}


def peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(src="measured (MEASURED_PEAKS.json)", hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"],
                    tf_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]))
    return dict(src="fallback (B200_PROFILING.md)", hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc = index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        self.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["unavailable"], samples=0)
        return dict(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))


# --------------------------------------------------------------------------------------------- workload description
def omni_tokens(b, n_v=N_V, n_a=N_A, n_d=N_D, S=S_TXT, frame_tokens=T_IMG):
    """(north-star tokens, processed tokens) per step of b samples: north_star counts 1568 per video sample, 257 per image
    (each spectrogram slice and the depth map are images through the tower), 128 per text; processed = what the encoders
    actually see (257 per frame: an 8-frame video is 2056 tower tokens)."""
    vis = NS_TOKENS["image"] if n_v == 1 else NS_TOKENS["video"] * n_v / N_V
    ns = b * (vis + NS_TOKENS["image"] * (n_a + n_d) + S)
    processed = b * (frame_tokens * (n_v + n_a + n_d) + S)
    return ns, processed


def omni_flops(b, n_v=N_V, n_a=N_A, n_d=N_D, S=S_TXT, task=OMNI_TASK, frame_tokens=T_IMG, frame_flops=None, tower_dim=1408):
    """Algorithmic forward FLOPs of one MiCo.forward step as the reference computes it (2MNK per linear layer, 4*Sq*Sk*D per
    attention; no recompute counted): (tower, fusion inputs, BERT + LM head).  fwd+bwd = 3x."""
    H, F, V, L = 768, 3072, 30522, 12
    frames = b * (n_v + n_a + n_d)
    tower = frames * (frame_flops if frame_flops else FWD_FLOPS_PER_FRAME)
    fusion = 2 * frames * frame_tokens * tower_dim * H

    def bert(n, Sk, lm):
        M = n * S
        per = 2 * M * H * 3 * H + 4 * n * S * S * H + 2 * M * H * H + 2 * 2 * M * H * F
        if Sk:
            per += 2 * M * H * H + 2 * n * Sk * H * 2 * H + 4 * n * S * Sk * H + 2 * M * H * H
        tot = L * per
        if lm:
            tot += 2 * M * H * H + 2 * M * H * V
        return tot
    sk = {"tv": n_v * frame_tokens, "ta": n_a * frame_tokens, "tva": (n_v + n_a) * frame_tokens, "td": n_d * frame_tokens}
    text = bert(b, 0, False)
    for part in task.split("_"):
        for st in part.split("%")[1:]:
            text += bert(3 * b, sk[st], False) if part.startswith("ret") else bert(b, sk[st], True)
    return tower, fusion, text


def make_config(args):
    """Identical in the product and the reference arm (the driver compares them)."""
    if args.config == "vitg":
        return dict(workload=WORKLOAD["vitg"], batch_per_gpu=args.batch, tokens_per_image=T_IMG, drop_path_rate=0.4,
                    loss="tokens.pow(2).mean()", weight_cast_in_step=True,
                    e2e_inputs="pinned host pixels, H2D every step on a side stream one step ahead (the reference's "
                               "PrefetchLoader, data/utils/loader.py:100-142)",
                    l2="working set per step (~35 GB of activations) exceeds the 126 MB L2; no flush needed",
                    grad_sync=(f"nccl all_reduce(SUM, {args.grad_dtype}) per {args.bucket_blocks}-block bucket of the flat "
                               "gradient buffer, overlapped with backward on a side stream (mico_b200/dp.py)")
                    if args.gpus > 1 else "none (1 GPU)")
    sp = SPECS[args.config]
    ns, processed = omni_tokens(args.batch, sp["n_v"], sp["n_a"], sp["n_d"], frame_tokens=sp["frame_tokens"])
    ckpt = sp["ckpt"] and not args.no_ckpt
    return dict(workload=sp["workload"], batch_per_gpu=args.batch, video_frames=sp["n_v"], audio_slices=sp["n_a"],
                depth_frames=sp["n_d"], text_len=S_TXT, task=sp["task"],
                tower="EVA01-CLIP-g-14 (1408 x 40, 16 heads x 88, MLP 6144), DropPath 0.4" if sp["tower"].startswith("eva")
                else "Swin-B (embed 128, depths 2/2/18/2, window 7), DropPath 0.2",
                text_encoder="bert-base-uncased-crossattn (12 layers, dropout 0.1), tied LM head",
                optimizer="AdamW (fused), gradient SUM over ranks",
                token_accounting=f"value counts north_star tokens (1568 / 8-frame video, 257 / image incl. each audio slice and depth map, "
                                 f"128 / text) = {int(ns / args.batch)} per sample; the encoders process {int(processed / args.batch)} per sample",
                l2="working set per step (tens of GB of activations) exceeds the 126 MB L2; no flush needed",
                activation_checkpointing=("none" if not ckpt else "tower blocks keep their input only (the reference's "
                                          "config.checkpointing, eva_vit_model.py:635-637)")
                + ((f"; kept besides the input, from the last block backwards: {args.light_blocks} blocks qkv + attention "
                    f"output + x1, {args.qkv_blocks} blocks qkv + attention output, "
                    f"{'all remaining' if args.attn_blocks < 0 else args.attn_blocks} blocks attention output (never re-run "
                    "the attention kernel)") if ckpt else ""),
                schedule="one tower pass for all modalities; ITM / caption sub-tasks differentiated group by group inside "
                         "forward (mico_b200/train_step.py)",
                e2e_inputs="pinned host pixels + waveforms, H2D every step on a side stream one step ahead (the reference's "
                           "PrefetchLoader, data/utils/loader.py:100-142); token ids from the host; 3 losses read back",
                grad_sync=(f"nccl all_reduce(SUM, {args.grad_dtype}) of the flat gradient buffer: non-tower segment when the tower "
                           f"backward starts, then per {args.bucket_blocks} tower blocks, overlapped on a side stream "
                           "(mico_b200/dp.py)") if args.gpus > 1 else "none (1 GPU)")


def model_cfg(ckpt=True, tower="evaclip01_giant"):
    from mico_b200.mico import _AttrDict
    return _AttrDict(vision_encoder_type=tower, vision_resolution=224, checkpointing=ckpt, contra_dim=512,
                     max_vision_sample_num=N_V, max_audio_sample_num=N_A, max_depth_sample_num=N_D, beam_size=3, itm_ratio=0.1,
                     max_omni_caption_len=70, max_caption_len=S_TXT, max_subtitle_len=70, frame_embedding_type="adaptive",
                     pool_video=False)


def host_batch(b, rank, n_v=N_V, n_d=N_D, S=S_TXT, pin=True, wave=True):
    """Seeded synthetic inputs of one rank, on the host: pixels ~ N(0,1) (already normalised),
    waveforms 0.1 N(0,1), token ids uniform in [1000, 30522) with [CLS] / [SEP] / padding."""
    import torch
    g = torch.Generator().manual_seed(1234 + rank)
    lens = torch.randint(8, S + 1, (b,), generator=g)
    att = (torch.arange(S)[None] < lens[:, None]).long()
    ids = torch.randint(1000, 30522, (b, S), generator=g) * att
    ids[:, 0] = 101
    ids[torch.arange(b), lens - 1] = 102
    h = dict(vision_pixels=torch.randn(b, n_v, 3, 224, 224, generator=g),
             depth_pixels=torch.randn(b, n_d, 3, 224, 224, generator=g) if n_d else None,
             audio_waveforms=0.1 * torch.randn(b, WAVE_SAMPLES, generator=g) if wave else None,
             input_ids=ids, attention_mask=att)
    if pin:
        h = {k: (v.pin_memory() if v is not None else None) for k, v in h.items()}
    return h


# --------------------------------------------------------------------------------------------- reference arm (CPU)
def _cpu_threads():
    import torch
    n = os.cpu_count() or 1
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        pass
    torch.set_num_threads(n)          # torchrun exports OMP_NUM_THREADS=1
    return torch.get_num_threads()


def vitg_oracle_step(params, cfg, x, dp):
    """One fwd+bwd of the CPU oracle (restates eva_vit_model.py:611-650; fp32 eager on host cores)."""
    from oracle import eva_vit as O
    for p in params.values():
        p.grad = None
    y = O.forward_features(params, x, cfg, dp_scales=dp)
    y.float().pow(2).mean().backward()
    return y


class OmniOracle:
    """The reference's CPU eager omni step restated in oracle/ (oracle/mico.py:omni_step; /root/reference is not on the GPU
    box): fp32, full ViT-g/14 + BERT-base, on a bounded sample of the per-rank batch."""

    def __init__(self, seed=0):
        import torch
        from oracle import bert as OB
        from oracle import eva_vit as OV
        self.vit_cfg = OV.VIT_G14
        torch.manual_seed(seed)
        p = {"vision_encoder.visual." + k: v for k, v in OV.init_params(self.vit_cfg, seed=seed).items()}
        p.update({"multimodal_encoder." + k: v for k, v in OB.init_params(seed=seed, prefix="").items()})
        H, D, C = 768, 1408, 512
        n = lambda *s: 0.02 * torch.randn(*s)
        for k in "tvad":
            p[f"contra_head_{k}.linear.weight"] = n(C, H if k == "t" else D)
        for k, i in (("va", 2 * D), ("id", 2 * D)):
            p[f"contra_head_{k}.weight"], p[f"contra_head_{k}.bias"] = n(C, i), torch.zeros(C)
        p["contra_temp"] = torch.tensor(0.07)
        p["itm_head.linear1.weight"], p["itm_head.linear1.bias"] = n(H, H), torch.zeros(H)
        p["itm_head.layernorm.weight"], p["itm_head.layernorm.bias"] = torch.ones(H), torch.zeros(H)
        p["itm_head.linear2.weight"], p["itm_head.linear2.bias"] = n(2, H), torch.zeros(2)
        for kind, nmax in (("vision", N_V), ("audio", N_A), ("depth", N_D)):
            p[f"hidden_trans_{kind}_multimodal.0.weight"], p[f"hidden_trans_{kind}_multimodal.0.bias"] = n(H, D), torch.zeros(H)
            p[f"hidden_trans_{kind}_multimodal.1.weight"], p[f"hidden_trans_{kind}_multimodal.1.bias"] = torch.ones(H), torch.zeros(H)
            p[f"{kind}_frame_embedding"] = n(1, nmax, H)
            p[f"{kind}_type_embeddings"] = n(1, 1, H)
        self.p = {k: (v.requires_grad_(True) if v.is_floating_point() else v) for k, v in p.items()}

    def batch(self, b, n_v, n_a, n_d):
        import torch
        from oracle import fbank as OF
        h = host_batch(b, 0, n_v=n_v, n_d=n_d, pin=False, wave=n_a > 0)
        spec = torch.stack([OF.audio_processor(w.unsqueeze(0), melbins=224, target_length=224, sample_num=n_a)
                            for w in h["audio_waveforms"]], dim=0) if n_a else None
        ids, att = h["input_ids"], h["attention_mask"]
        g = torch.Generator().manual_seed(7)
        pick = (torch.rand(ids.shape, generator=g) < 0.6) & (att > 0)
        pick[:, 0] = False
        pick[:, 1] = True
        return dict(vision_pixels=h["vision_pixels"], depth_pixels=h["depth_pixels"], audio_spectrograms=spec, ids=ids, att=att,
                    cap_ids=torch.where(pick, torch.full_like(ids, 103), ids),
                    cap_labels=torch.where(pick, ids, torch.full_like(ids, -100)))

    def step(self, batch, task=OMNI_TASK):
        from oracle import mico as OM
        for v in self.p.values():
            if v.is_floating_point():
                v.grad = None
        out = OM.omni_step(self.p, batch, self.vit_cfg, 12, 12, task)
        sum(out.values()).backward()
        return {k: float(v) for k, v in out.items()}


def run_reference(args):
    """--impl reference: the reference's CPU eager implementation of the path (oracle port) on all host threads.  Each step
    is a bounded sample of the per-rank batch, sized so that --steps + --warmup end within a few minutes."""
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = _cpu_threads()
    budget = float(os.environ.get("MICO_BENCH_REF_BUDGET_S", "200"))     # seconds for the whole --steps + --warmup run
    nsteps = args.steps + args.warmup
    if args.config == "vitg":
        from oracle import eva_vit as O
        cfg = O.VIT_G14
        params = {k: v.requires_grad_(True) for k, v in O.init_params(cfg, seed=0).items()}
        g = torch.Generator().manual_seed(1234)
        x1 = torch.randn(1, 3, 224, 224, generator=g)
        t0 = time.perf_counter()
        vitg_oracle_step(params, cfg, x1, None)
        t1 = time.perf_counter() - t0
        b = int(max(1, min(4, budget / (nsteps * t1))))
        x = torch.randn(b, 3, 224, 224, generator=g)
        dp = torch.ones(cfg["depth"], 2, b)
        run = lambda: vitg_oracle_step(params, cfg, x, dp)
        tokens = b * T_IMG
        sample = f"{b} image(s) of the bs-{args.batch} batch per step, ViT-g/14 fwd+bwd fp32 eager, {args.steps} steps"
    else:
        orc = OmniOracle()
        # probe: one frame through the tower, fwd+bwd, to size the sample
        from oracle import eva_vit as O
        tp = {k[len("vision_encoder.visual."):]: v for k, v in orc.p.items() if k.startswith("vision_encoder.visual.")}
        x1 = torch.randn(1, 3, 224, 224)
        vitg_oracle_step(tp, orc.vit_cfg, x1, None)
        t0 = time.perf_counter()
        vitg_oracle_step(tp, orc.vit_cfg, x1, None)
        t_frame = time.perf_counter() - t0
        # a step of b=2 samples costs ~ 2*(n_v+n_a+n_d) frames * t_frame * ~1.25 (text side); shrink the frame counts
        # (keeping every modality and every sub-task) until the run fits the budget
        if args.config == "trimodal":
            print(json.dumps(dict(impl="reference", unavailable="the CPU oracle has no Swin-B MiCo step (builder-run config)")))
            return
        sp = SPECS[args.config]
        b = 2
        shapes = [(8, 3, 1), (4, 2, 1), (2, 1, 1), (1, 1, 1)] if args.config == "omni" else [(sp["n_v"], sp["n_a"], sp["n_d"])]
        n_v, n_a, n_d = next((s for s in shapes if nsteps * b * sum(s) * t_frame * 1.25 <= budget), shapes[-1])
        batch = orc.batch(b, n_v, n_a, n_d)
        run = lambda: orc.step(batch, sp["task"])
        tokens = omni_tokens(b, n_v, n_a, n_d)[0]
        sample = (f"{b} samples per step with video {n_v}f / audio {n_a} slices / depth {n_d}f / text S={S_TXT} (per-rank batch is "
                  f"{args.batch} x {sp['n_v']} / {sp['n_a']} / {sp['n_d']}), every sub-task of {sp['task']}, full ViT-g/14 + BERT-base, "
                  f"fp32 eager, {args.steps} steps")
    for _ in range(args.warmup):
        run()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        run()
    dt = time.perf_counter() - t0
    tps = args.steps * tokens / dt
    line = dict(impl="reference", metric=METRIC[args.config], value=tps, unit="tokens/s", n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=1e3 * dt / args.steps, higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="f32", data="synthetic", config=make_config(args),
                cpu_baseline=dict(value=tps, unit="tokens/s", cores=cores, kind="port", sample=sample),
                e2e=dict(value=tps, unit="tokens/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


def gpu_eager_leg(config, dev, b=4):
    """Informational context for the product line (SURVEY "facts": the reference's real competitor is PyTorch
    eager on the same GPU): the oracle port -- plain PyTorch ops -> ATen / cuBLAS / cuDNN kernels -- on this B200 under
    torch.autocast(bf16), full ViT-g/14 + BERT-base, the same sub-tasks, on `b` samples per step (eager keeps every
    activation, no checkpointing: b = 64 does not fit).  Nothing of mico_b200 runs here; only bench.py imports the oracle."""
    import torch
    sp = SPECS[config]
    try:
        torch.cuda.empty_cache()
        torch.cuda.reset_peak_memory_stats(dev)
        orc = OmniOracle()
        orc.p = {k: v.detach().to(dev).requires_grad_(v.is_floating_point()) for k, v in orc.p.items()}
        batch = {k: (v.to(dev) if v is not None else None) for k, v in orc.batch(b, sp["n_v"], sp["n_a"], sp["n_d"]).items()}

        def run():
            with torch.autocast("cuda", dtype=torch.bfloat16):
                return orc.step(batch, sp["task"])
        run()
        torch.cuda.synchronize()
        n = 3
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            losses = run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        tok = omni_tokens(b, sp["n_v"], sp["n_a"], sp["n_d"], frame_tokens=sp["frame_tokens"])[0]
        out = dict(value=tok / (ms * 1e-3), unit="tokens/s", ms_per_step=ms, samples_per_step=b, kind="oracle port on cuda, "
                   "torch.autocast(bf16), eager ATen/cuBLAS kernels, no optimizer step, no activation checkpointing",
                   peak_hbm_gb=round(torch.cuda.max_memory_allocated(dev) / 1e9, 1), losses=losses)
        del orc, batch
        torch.cuda.empty_cache()
        return out
    except Exception as e:          # informational leg: never fails the bench
        return dict(unavailable=f"{type(e).__name__}: {str(e)[:200]}")


def cpu_baseline_leg(config, budget_s=25.0):
    """cpu_baseline for the product line: the oracle timed on this box's host cores on a bounded sample (rank 0, N = 1)."""
    import torch
    cores = _cpu_threads()
    if config == "vitg":
        from oracle import eva_vit as O
        cfg = O.VIT_G14
        params = {k: v.requires_grad_(True) for k, v in O.init_params(cfg, seed=0).items()}
        g = torch.Generator().manual_seed(1234)
        x = torch.randn(1, 3, 224, 224, generator=g)
        t0 = time.perf_counter()
        vitg_oracle_step(params, cfg, x, None)           # warm-up (allocator, thread pool)
        t1 = time.perf_counter() - t0
        b = 2 if t1 * 2 * 3 < budget_s else 1
        x = torch.randn(b, 3, 224, 224, generator=g)
        n = int(max(1, min(5, budget_s / (t1 * b))))
        t0 = time.perf_counter()
        for _ in range(n):
            vitg_oracle_step(params, cfg, x, None)
        dt = time.perf_counter() - t0
        return dict(value=n * b * T_IMG / dt, unit="tokens/s", cores=cores, kind="port",
                    sample=f"{n} fwd+bwd steps of {b} image(s) (of the bs-64 batch), full ViT-g/14, fp32 eager, after 1 warm-up")
    orc = OmniOracle()
    from oracle import eva_vit as O
    tp = {k[len("vision_encoder.visual."):]: v for k, v in orc.p.items() if k.startswith("vision_encoder.visual.")}
    x1 = torch.randn(1, 3, 224, 224)
    vitg_oracle_step(tp, orc.vit_cfg, x1, None)          # warm-up
    t0 = time.perf_counter()
    vitg_oracle_step(tp, orc.vit_cfg, x1, None)
    t_frame = time.perf_counter() - t0
    sp = SPECS[config]
    b = 2
    shapes = [(8, 3, 1), (4, 2, 1), (2, 1, 1), (1, 1, 1)] if config == "omni" else [(sp["n_v"], sp["n_a"], sp["n_d"])]
    n_v, n_a, n_d = next((s for s in shapes if b * sum(s) * t_frame * 1.25 <= budget_s), shapes[-1])
    batch = orc.batch(b, n_v, n_a, n_d)
    t0 = time.perf_counter()
    losses = orc.step(batch, sp["task"])
    dt = time.perf_counter() - t0
    return dict(value=omni_tokens(b, n_v, n_a, n_d)[0] / dt, unit="tokens/s", cores=cores, kind="port",
                sample=f"1 step of {b} samples with video {n_v}f / audio {n_a} slices / depth {n_d}f / text S={S_TXT}, every sub-task "
                       f"of {sp['task']}, full ViT-g/14 + BERT-base, fp32 eager (after a 1-frame tower warm-up); losses {losses}")


# --------------------------------------------------------------------------------------------- product arm (B200)
def _dist_setup(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    return world, rank, local, dev


def nccl_fixture_replay(rank, world, dev):
    """Self-check at N > 1 (before timing): ranks 0 and 1 replay tests/golden/losses_2rank.pt -- the reference's OWN
    data/model/vast.py forward_ret / forward_cap run on two gloo ranks (oracle/make_golden.py) -- through MiCo.forward over
    NCCL (feature gathers + all_gather_with_grad): per-rank losses 1e-3 relative, fusion-input gradient 3e-2 rel-L2."""
    import torch
    import torch.distributed as dist
    import torch.nn.functional as F
    from mico_b200 import mico as M
    gd = os.path.join(REPO, "tests", "golden")
    sub = dist.new_group([0, 1])
    ok, detail = True, {}
    if rank < 2:
        one = torch.load(os.path.join(gd, "losses_tiny.pt"), weights_only=False)
        r = torch.load(os.path.join(gd, "losses_2rank.pt"), weights_only=False)["ranks"][rank]
        cfg = M._AttrDict(vision_encoder_type="evaclip01_giant", vision_resolution=224, checkpointing=False, contra_dim=32,
                          max_vision_sample_num=2, max_audio_sample_num=3, max_depth_sample_num=1, beam_size=3, itm_ratio=0.1,
                          max_omni_caption_len=70, max_caption_len=24, max_subtitle_len=70, frame_embedding_type="adaptive",
                          pool_video=False,
                          vision_tower_kwargs=dict(embed_dim=176, depth=2, num_heads=2, mlp_ratio=2.0, drop_path_rate=0.0,
                                                   num_classes=8),
                          bert_config=dict(vocab_size=1000, hidden_size=128, num_hidden_layers=2, num_attention_heads=2,
                                           intermediate_size=256, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0,
                                           max_position_embeddings=64))
        torch.manual_seed(0)
        model = M.MiCo.from_pretrained(cfg, {})
        model.load_state_dict(one["state_dict"], strict=False)
        model = model.to(dev).train()
        prev = M.set_process_group(sub)
        try:
            raw_t, raw_v, cond = (r[k].clone().to(dev).requires_grad_(True) for k in ("raw_t", "raw_v", "cond"))
            batch = dict(feat_t=F.normalize(raw_t, dim=-1), feat_v=F.normalize(raw_v, dim=-1), condition_feats_v=cond,
                         caption_tokens=M._AttrDict(input_ids=r["ids"].to(dev), attention_mask=r["att"].to(dev)),
                         cap_input_ids=r["cap_ids"].to(dev), cap_labels=r["cap_labels"].to(dev),
                         itm_neg_cond_tv=r["neg_c"], itm_neg_text_tv=r["neg_t"])
            out = model(batch, "ret%tv_cap%tv", compute_loss=True)
            sum(out.values()).backward()
            for k in ("loss_itc", "loss_itm", "loss_cap"):
                a, e = out[k].item(), r[k].item()
                detail[k] = abs(a - e) / max(1.0, abs(e))
                ok &= detail[k] <= 1e-3
            for name, got, want in (("d_cond", cond.grad, r["d_cond"]), ("d_raw_t", raw_t.grad, r["d_raw_t"]),
                                    ("d_raw_v", raw_v.grad, r["d_raw_v"])):
                detail[name] = ((got.cpu() - want).norm() / want.norm()).item()
                ok &= detail[name] < 3e-2
        finally:
            M.set_process_group(prev)
        del model
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    return dict(status="pass" if flag.item() > 0 else "FAIL", world=2, fixture="tests/golden/losses_2rank.pt",
                rank0={k: float(f"{v:.3g}") for k, v in detail.items()})


def run_product_omni(args):
    import torch
    import torch.distributed as dist
    from mico_b200 import _lib, dp, ops, optim
    from mico_b200.audioprocessor import AudioProcessor
    from mico_b200.mico import MiCo, _AttrDict

    sys.stdout.flush()
    json_fd = os.dup(1)          # stdout carries exactly ONE JSON line; NCCL prints its banner on fd 1 -> point it at stderr
    os.dup2(2, 1)
    world, rank, local, dev = _dist_setup(args)
    selfcheck = {}
    if world > 1 and not args.no_selfcheck:
        selfcheck["nccl_reference_fixture_2rank"] = nccl_fixture_replay(rank, world, dev)
    sp = SPECS[args.config]
    task = sp["task"]
    B = args.batch
    torch.manual_seed(0)
    cfg = model_cfg(ckpt=sp["ckpt"] and not args.no_ckpt, tower=sp["tower"])
    with torch.device(dev):
        model = MiCo.from_pretrained(cfg, {})
    model = model.to(dev).train()
    tower = getattr(model.vision_encoder, "visual", None)       # None: Swin (the tower module itself is the encoder)
    if tower is not None:
        tower.ckpt_light_blocks, tower.ckpt_qkv_blocks, tower.ckpt_attn_blocks = args.light_blocks, args.qkv_blocks, args.attn_blocks
    flat = dp.FlatGrads(model)
    sync = dp.GradSync(flat, bucket_blocks=args.bucket_blocks,
                       dtype=torch.bfloat16 if args.grad_dtype == "bf16" else torch.float32) if world > 1 else None
    no_decay = ("bias", "LayerNorm.bias", "LayerNorm.weight")
    named = [(k, p) for k, p in model.named_parameters()]
    opt = optim.AdamW([dict(params=[p for k, p in named if not any(s in k for s in no_decay)], weight_decay=0.01),
                       dict(params=[p for k, p in named if any(s in k for s in no_decay)], weight_decay=0.0)],
                      lr=1e-6, betas=(0.9, 0.98))
    if tower is not None:
        opt.attach_bf16_sinks(tower)
    audio = AudioProcessor(melbins=224, target_length=224, sample_num=max(sp["n_a"], 1), training=True, device=dev)
    host = host_batch(B, rank, n_v=sp["n_v"], n_d=sp["n_d"], wave=sp["n_a"] > 0)
    resident = {k: v.to(dev) for k, v in host.items() if k not in ("input_ids", "attention_mask") and v is not None}
    host_loss = torch.empty(3, dtype=torch.float32).pin_memory()

    def step(dev_in, verify=False):
        """One training step on device-resident pixels / waveforms; token ids come from the host like a tokenizer's output."""
        flat.zero_grad()
        if sync is not None:
            sync.begin_step(verify=verify)
        batch = dict(vision_pixels=dev_in["vision_pixels"],
                     caption_tokens=_AttrDict(input_ids=host["input_ids"], attention_mask=host["attention_mask"]))
        if "depth_pixels" in dev_in:
            batch["depth_pixels"] = dev_in["depth_pixels"]
        if "audio_waveforms" in dev_in:
            batch["audio_spectrograms"] = audio.batch(dev_in["audio_waveforms"])
        out = model(batch, task, compute_loss=True)
        loss = sum(out.values())
        loss.backward()
        if sync is not None:
            sync.finish()
        flat.detach_unused()
        opt.step()
        return out

    copy_stream = torch.cuda.Stream(device=dev)
    staged = []
    h2d_keys = tuple(k for k in ("vision_pixels", "depth_pixels", "audio_waveforms") if host[k] is not None)

    def stage_next():
        # the reference's PrefetchLoader (data/utils/loader.py:100-142): the NEXT batch crosses PCIe on a side stream
        copy_stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(copy_stream):
            staged.append({k: host[k].to(dev, non_blocking=True) for k in h2d_keys})

    def e2e_step():
        if not staged:
            stage_next()
        x = staged.pop(0)
        torch.cuda.current_stream().wait_stream(copy_stream)
        for v in x.values():
            v.record_stream(torch.cuda.current_stream())
        stage_next()
        out = step(x)
        host_loss.copy_(torch.stack([out["loss_itc"].detach(), out["loss_itm"].detach(), out["loss_cap"].detach()]),
                        non_blocking=True)
        torch.cuda.current_stream().synchronize()          # the reference loop's .item() (pipeline.py:47)
        return host_loss.tolist()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    W = max(args.warmup, 3)
    for i in range(W):
        out = step(resident, verify=(i == 0 and sync is not None and not args.no_selfcheck))
        if i == 0 and sync is not None and not args.no_selfcheck:
            ok, msg = sync.check()
            flag = torch.tensor([1.0 if ok else 0.0], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            selfcheck["grad_sync_sum"] = dict(status="pass" if flag.item() > 0 else "FAIL", detail=msg,
                                              bytes_reduced_per_step=sync.bytes_reduced)
    losses0 = {k: float(v) for k, v in out.items()}
    barrier()
    torch.cuda.reset_peak_memory_stats(dev)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    _lib.reset_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step(resident)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = _lib.launch_count()
    peak_mem = torch.cuda.max_memory_allocated(dev) / 1e9
    peak_reserved = torch.cuda.max_memory_reserved(dev) / 1e9

    e2e_losses = e2e_step()
    barrier()
    e0.record()
    for _ in range(args.steps):
        e2e_losses = e2e_step()
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None

    nprof = min(args.steps, 3)
    ops.profile_enable(True)
    for _ in range(nprof):
        step(resident)
    fam = ops.profile_collect()
    ops.profile_enable(False)

    if world > 1:
        t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = t.tolist()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk = peaks()
    ns_tok, proc_tok = omni_tokens(B, sp["n_v"], sp["n_a"], sp["n_d"], frame_tokens=sp["frame_tokens"])
    value = world * ns_tok * args.steps / (ms / 1e3)
    e2e = world * ns_tok * args.steps / (ms_e2e / 1e3)
    tower_f, fusion_f, text_f = omni_flops(B, sp["n_v"], sp["n_a"], sp["n_d"], task=task, frame_tokens=sp["frame_tokens"],
                                           frame_flops=sp["frame_flops"], tower_dim=1408 if tower is not None else 1024)
    step_flops = 3 * (tower_f + fusion_f + text_f)
    step_s = ms / args.steps * 1e-3
    g = fam["gemm"]
    gemm_tflops = g["work"] / (g["ms"] * 1e-3) / 1e12 if g["ms"] > 0 else 0.0
    fam_total = sum(v["ms"] for v in fam.values())
    traffic, traffic_src, traffic_note = None, None, None
    tpath = os.path.join(REPO, "profiles", "gemm_traffic.json")
    if os.path.exists(tpath) and tower is not None:
        tj = json.load(open(tpath))
        traffic, traffic_src = tj["avg_dram_bytes_per_launch"], tj["source"]
        traffic_note = ("average over the four representative TOWER launches of the capture (" + tj.get("note", "") + "); their "
                        "algorithmic operand + output bytes average 4.62 GB (5.43 / 5.43 / 2.98 / 4.65): the two GEMMs with a K = 6144 "
                        "or 197 376-row contraction still re-read part of their streamed operand (fc2 fwd 5.1 GB read for 3.5, fc2 wgrad 5.6 for 3.0; "
                        "8.4 / 7.4 GB before the balanced unit schedule, DESIGN.md 4a); `achieved` "
                        "averages all of the family's launches, including the small text-encoder GEMMs")
    is_tf = lambda k: "attention" in k or k == "gemm"
    roofline = dict(bound="tensor", kernel="gemm_bf16_kernel (tcgen05 GEMM: every linear layer fwd / dgrad / wgrad of the tower and "
                                           "the text encoder)",
                    achieved=gemm_tflops, peak=pk["tf_sustained"], unit="TFLOP/s", frac=gemm_tflops / pk["tf_sustained"],
                    traffic=traffic, traffic_unit="bytes/launch (dram read+write)", traffic_source=traffic_src,
                    traffic_note=traffic_note,
                    algorithmic_flops_per_launch=g["work"] / max(g["calls"], 1), peak_source=pk["src"] + ", sustained bf16",
                    launches_per_step=g["calls"] / nprof, avg_launch_ms=g["ms"] / max(g["calls"], 1),
                    share_of_step=g["ms"] / fam_total if fam_total else None,
                    whole_step_algorithmic_tflop=step_flops / 1e12,
                    whole_step_tflops=step_flops / step_s / 1e12,
                    whole_step_frac=step_flops / step_s / 1e12 / pk["tf_sustained"],
                    whole_step_note="algorithmic FLOPs of the reference's step (3 x forward; activation recompute of the "
                                    "checkpointed tower blocks is extra work and not counted)",
                    families={k: dict(ms_per_step=v["ms"] / nprof, calls_per_step=v["calls"] / nprof,
                                      rate=(v["work"] / (v["ms"] * 1e-3) / (1e12 if is_tf(k) else 1e9)) if v["ms"] > 0 else None,
                                      rate_unit="TFLOP/s" if is_tf(k) else "GB/s") for k, v in fam.items()})
    cpu = cpu_baseline_leg(args.config) if (world == 1 and not args.no_cpu_baseline and args.config != "trimodal") else None
    gpu_eager = None
    if world == 1 and args.gpu_eager and args.config in ("omni", "imgtext"):
        import gc
        staged.clear()
        del model, opt, flat, resident, named, tower, out, audio       # the closures above are not called again
        gc.collect()
        gpu_eager = gpu_eager_leg(args.config, dev)
    h2d = sum(host[k].numel() * host[k].element_size() for k in h2d_keys) + 2 * host["input_ids"].numel() * 8 * 2
    config = make_config(args)
    extra = dict(processed_tokens_per_s=world * proc_tok * args.steps / (ms / 1e3), peak_hbm_gb=round(peak_mem, 1),
                 peak_hbm_reserved_gb=round(peak_reserved, 1), losses_after_warmup=losses0)
    line = dict(metric=METRIC[args.config], value=value, unit="tokens/s", n_gpus=world, steps=args.steps, warmup=W,
                ms_per_step=ms / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="bf16",
                data="synthetic", config=config, clocks=clocks,
                e2e=dict(value=e2e, unit="tokens/s", ms_per_step=ms_e2e / args.steps, h2d_bytes_per_step=h2d,
                         d2h_bytes_per_step=12, losses=e2e_losses),
                gpu_launches=launches, roofline=roofline, cpu_baseline=cpu, gpu_eager=gpu_eager, selfcheck=selfcheck or None,
                extra=extra)
    sys.stdout.flush()
    os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


def run_product_vitg(args):
    """BASELINE configs[1]: the image-only tower step (round 1's bench), now on the package's gradient sync."""
    import torch
    import torch.distributed as dist
    from mico_b200 import _lib, dp, ops
    from mico_b200.eva_vit import EVAVisionTransformer

    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    world, rank, local, dev = _dist_setup(args)
    B = args.batch
    torch.manual_seed(0)
    with torch.device(dev):     # random-init weights of the named architecture, created on the device
        tower = EVAVisionTransformer(**VIT_G)
    tower = tower.train()
    gen = torch.Generator().manual_seed(1234 + rank)
    host_pixels = torch.randn(B, 3, 224, 224, generator=gen).pin_memory()
    dev_pixels = host_pixels.to(dev)
    host_loss = torch.empty((), dtype=torch.float32).pin_memory()
    flat = dp.FlatGrads(tower, tower=tower)
    sync = dp.GradSync(flat, bucket_blocks=args.bucket_blocks,
                       dtype=torch.bfloat16 if args.grad_dtype == "bf16" else torch.float32) if world > 1 else None
    selfcheck = {}

    def step(pixels, verify=False):
        tower.invalidate_weight_cache()         # no optimizer in this configuration: the per-step weight cast stays inside
        flat.zero_grad()
        if sync is not None:
            sync.begin_step(verify=verify)
        y = tower(pixels, return_all_features=True)
        loss = y.float().pow(2).mean()
        loss.backward()
        if sync is not None:
            sync.finish()
        return loss

    copy_stream = torch.cuda.Stream(device=dev)
    staged = []

    def stage_next():
        copy_stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(copy_stream):
            staged.append(host_pixels.to(dev, non_blocking=True))

    def e2e_step():
        if not staged:
            stage_next()
        x = staged.pop(0)
        torch.cuda.current_stream().wait_stream(copy_stream)
        x.record_stream(torch.cuda.current_stream())
        stage_next()
        loss = step(x)
        host_loss.copy_(loss.detach(), non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return float(host_loss)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    W = max(args.warmup, 3)
    for i in range(W):
        step(dev_pixels, verify=(i == 0 and sync is not None))
        if i == 0 and sync is not None:
            ok, msg = sync.check()
            selfcheck["grad_sync_sum"] = dict(status="pass" if ok else "FAIL", detail=msg)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    _lib.reset_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step(dev_pixels)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = _lib.launch_count()
    e2e_step()
    barrier()
    e0.record()
    for _ in range(args.steps):
        e2e_step()
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    ops.profile_enable(True)
    for _ in range(args.steps):
        step(dev_pixels)
    fam = ops.profile_collect()
    ops.profile_enable(False)
    if world > 1:
        t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = t.tolist()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk = peaks()
    tokens_per_step = world * B * T_IMG
    value = tokens_per_step * args.steps / (ms / 1e3)
    e2e = tokens_per_step * args.steps / (ms_e2e / 1e3)
    g = fam["gemm"]
    gemm_tflops = g["work"] / (g["ms"] * 1e-3) / 1e12 if g["ms"] > 0 else 0.0
    fam_total = sum(v["ms"] for v in fam.values())
    step_flops = 3 * FWD_FLOPS_PER_FRAME * B
    traffic, traffic_src = None, None
    tpath = os.path.join(REPO, "profiles", "gemm_traffic_vitg.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        traffic, traffic_src = tj["avg_dram_bytes_per_launch"], tj["source"]
    is_tf = lambda k: "attention" in k or k == "gemm"
    roofline = dict(bound="tensor", kernel="gemm_bf16_kernel (tcgen05 GEMM, all linear fwd/dgrad/wgrad)",
                    achieved=gemm_tflops, peak=pk["tf_sustained"], unit="TFLOP/s",
                    frac=gemm_tflops / pk["tf_sustained"], traffic=traffic, traffic_unit="bytes/launch (dram read+write)",
                    traffic_source=traffic_src, algorithmic_flops_per_launch=g["work"] / max(g["calls"], 1),
                    peak_source=pk["src"] + ", sustained bf16",
                    launches_per_step=g["calls"] / args.steps, avg_launch_ms=g["ms"] / max(g["calls"], 1),
                    share_of_step=g["ms"] / fam_total if fam_total else None,
                    whole_step_tflops=step_flops / (ms / args.steps * 1e-3) / 1e12,
                    whole_step_frac=step_flops / (ms / args.steps * 1e-3) / 1e12 / pk["tf_sustained"],
                    families={k: dict(ms_per_step=v["ms"] / args.steps, calls_per_step=v["calls"] / args.steps,
                                      rate=(v["work"] / (v["ms"] * 1e-3) / (1e12 if is_tf(k) else 1e9)) if v["ms"] > 0 else None,
                                      rate_unit="TFLOP/s" if is_tf(k) else "GB/s") for k, v in fam.items()})
    cpu = cpu_baseline_leg("vitg") if (world == 1 and not args.no_cpu_baseline) else None
    config = make_config(args)
    line = dict(metric=METRIC["vitg"], value=value, unit="tokens/s", n_gpus=world, steps=args.steps, warmup=W,
                ms_per_step=ms / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="bf16",
                data="synthetic", config=config, clocks=clocks,
                e2e=dict(value=e2e, unit="tokens/s", ms_per_step=ms_e2e / args.steps,
                         h2d_bytes_per_step=host_pixels.numel() * 4, d2h_bytes_per_step=4),
                gpu_launches=launches, roofline=roofline, cpu_baseline=cpu, selfcheck=selfcheck or None)
    sys.stdout.flush()
    os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="mico_b200", choices=["mico_b200", "reference"])
    ap.add_argument("--config", default="omni", choices=["omni", "imgtext", "trimodal", "vitg"])
    ap.add_argument("--batch", type=int, default=None, help="samples per GPU (default: the configuration's own)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-selfcheck", action="store_true")
    ap.add_argument("--gpu-eager", action="store_true",
                    help="omni / imgtext at N=1: also time the oracle port on the GPU under bf16 autocast (context, not a baseline)")
    ap.add_argument("--no-ckpt", action="store_true", help="omni: keep every tower activation (small --batch only)")
    ap.add_argument("--light-blocks", type=int, default=DEFAULT_LIGHT_BLOCKS,
                    help="omni: the last n tower blocks keep qkv / attention output / x1 instead of their input only")
    ap.add_argument("--qkv-blocks", type=int, default=DEFAULT_QKV_BLOCKS,
                    help="omni: the n tower blocks before those keep qkv / attention output")
    ap.add_argument("--attn-blocks", type=int, default=DEFAULT_ATTN_BLOCKS,
                    help="omni: the n tower blocks before those keep their attention output (-1: all remaining)")
    ap.add_argument("--bucket-blocks", type=int, default=int(os.environ.get("MICO_BENCH_BUCKET_BLOCKS", "5")))
    ap.add_argument("--grad-dtype", default="fp32", choices=["fp32", "bf16"])
    args = ap.parse_args()
    if args.batch is None:
        args.batch = 64 if args.config == "vitg" else SPECS[args.config]["batch"]
    if args.impl == "reference":
        run_reference(args)
    elif args.config in SPECS:
        run_product_omni(args)
    else:
        run_product_vitg(args)


if __name__ == "__main__":
    main()
