#!/usr/bin/env python
"""bench.py -- MiCo hot-path benchmark on B200 (contract in the task statement).

Workload (BASELINE.json configs[1]): ViT-g/14 (EVA01-CLIP-g-14: width 1408, 40 blocks, 16 heads x 88, MLP 6144)
image-only forward + backward at batch 64 of synthetic 224x224 pixels per GPU, training mode (DropPath
linspace(0, 0.4, 40)), loss = tokens.pow(2).mean().  257 tokens per image.
One step = one fwd+bwd pass over one batch, including the fp32->bf16 cast of all weights (they change every
optimizer step in training, so the cast is part of the pass).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--batch 64]

N > 1 is launched by torchrun, one rank per GPU; ranks process independent batches (weak scaling) and SUM
their flat gradient buckets over NCCL like the reference loop does (data/utils/pipeline.py:93-99).

Output: ONE JSON line on rank 0 (keys per the contract, plus "roofline" and "cpu_baseline").
`--impl reference` times the reference's CPU eager path restated in oracle/ (fp32, all host threads) on a bounded
sample of the same workload.
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

TOKENS_PER_IMAGE = 257
VIT_G = dict(img_size=224, patch_size=14, embed_dim=1408, depth=40, num_heads=16, mlp_ratio=4.3637, qkv_bias=True,
             drop_path_rate=0.4, num_classes=1024, use_mean_pooling=False)
# algorithmic FLOPs per image, forward: 40 blocks x (2*257*1408*(4224+1408+2*6144) + 4*257^2*1408)
# + patch embed 2*256*588*1408; backward = 2x forward.
FWD_FLOPS_PER_IMAGE = 40 * (2 * 257 * 1408 * (4224 + 1408 + 2 * 6144) + 4 * 257 * 257 * 1408) + 2 * 256 * 588 * 1408
METRIC = "omni-modal pretrain tokens/sec @ ViT-g/14 (image-only fwd+bwd, bs 64/GPU)"
WORKLOAD = "ViT-g/14 image-only fwd+bwd, bs=64 synthetic 224x224 per GPU (BASELINE configs[1])"


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


# --------------------------------------------------------------------------------------------- reference arm (CPU)
def oracle_step(params, cfg, x, dp):
    """One fwd+bwd of the CPU oracle (restates eva_vit_model.py:611-650; fp32 eager on host cores)."""
    import torch
    from oracle import eva_vit as O
    for p in params.values():
        p.grad = None
    y = O.forward_features(params, x, cfg, dp_scales=dp)
    y.float().pow(2).mean().backward()
    return y


def cpu_reference_setup(seed=0):
    import torch
    from oracle import eva_vit as O
    cfg = O.VIT_G14
    params = {k: v.requires_grad_(True) for k, v in O.init_params(cfg, seed=seed).items()}
    return cfg, params


def run_reference(args):
    """--impl reference: the reference's CPU eager implementation of the path (oracle port; /root/reference is not
    on the GPU box) on all host threads.  Each step is a bounded sample: `b` images of the bs-64 workload."""
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = torch.get_num_threads()
    cfg, params = cpu_reference_setup()
    g = torch.Generator().manual_seed(1234)
    x1 = torch.randn(1, 3, 224, 224, generator=g)
    t0 = time.perf_counter()
    oracle_step(params, cfg, x1, None)
    t1 = time.perf_counter() - t0
    budget = 150.0          # seconds for the whole --steps + --warmup run
    b = int(max(1, min(4, budget / ((args.steps + args.warmup) * t1))))
    x = torch.randn(b, 3, 224, 224, generator=g)
    dp = torch.ones(cfg["depth"], 2, b)
    for _ in range(args.warmup):
        oracle_step(params, cfg, x, dp)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle_step(params, cfg, x, dp)
    dt = time.perf_counter() - t0
    tps = args.steps * b * TOKENS_PER_IMAGE / dt
    sample = f"{b} image(s) of the bs-64 batch per step, ViT-g/14 fwd+bwd fp32 eager, {args.steps} steps"
    line = dict(impl="reference", metric=METRIC, value=tps, unit="tokens/s", n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=1e3 * dt / args.steps, higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="f32", data="synthetic",
                config=dict(workload=WORKLOAD, batch_per_gpu=args.batch, tokens_per_image=TOKENS_PER_IMAGE,
                            drop_path_rate=0.4, loss="tokens.pow(2).mean()",
                            reference_sample=f"{b} image(s) of the bs-{args.batch} batch per step (bounded CPU sample)"),
                cpu_baseline=dict(value=tps, unit="tokens/s", cores=cores, kind="port", sample=sample),
                e2e=dict(value=tps, unit="tokens/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


def cpu_baseline_leg(budget_s=20.0):
    """cpu_baseline for the product line: the oracle timed on this box's host cores on a bounded sample."""
    import torch
    cores = torch.get_num_threads()
    cfg, params = cpu_reference_setup()
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(1, 3, 224, 224, generator=g)
    t0 = time.perf_counter()
    oracle_step(params, cfg, x, None)           # warm-up (allocator, thread pool)
    t1 = time.perf_counter() - t0
    b = 2 if t1 * 2 * 3 < budget_s else 1
    x = torch.randn(b, 3, 224, 224, generator=g)
    n = int(max(1, min(5, budget_s / (t1 * b))))
    t0 = time.perf_counter()
    for _ in range(n):
        oracle_step(params, cfg, x, None)
    dt = time.perf_counter() - t0
    return dict(value=n * b * TOKENS_PER_IMAGE / dt, unit="tokens/s", cores=cores, kind="port",
                sample=f"{n} fwd+bwd steps of {b} image(s) (of the bs-64 batch), full ViT-g/14, fp32 eager, after 1 warm-up")


# --------------------------------------------------------------------------------------------- product arm (B200)
def run_product(args):
    import torch
    import torch.distributed as dist
    from mico_b200 import _lib, ops
    from mico_b200.eva_vit import EVAVisionTransformer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly ONE JSON line.  Native libraries write there too (NCCL prints "NCCL version ..." on fd 1 during
    # communicator setup), so fd 1 is pointed at stderr for the whole run and the line is written to the saved descriptor.
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    reserved = int(os.environ.get("MICO_BENCH_RESERVED_SMS", "0")) if world > 1 else 0
    if world > 1:
        # The overlapped gradient all-reduce runs NCCL kernels next to ours.  MICO_BENCH_RESERVED_SMS=n gives NCCL n SMs of its
        # own (and caps it there) so that its CTAs never push one of our one-CTA-per-SM persistent kernels into a second wave.
        # Measured at N=2 (ms/step; 1 GPU 121.9): n=0 126.5, n=4 129.2, n=8 130.0, n=2 161.2 -- an uncapped NCCL finishes each
        # bucket so quickly that the contention costs less than the SMs given away, so the default stays 0.
        if reserved > 0:
            os.environ.setdefault("NCCL_MAX_CTAS", str(reserved))
            os.environ.setdefault("NCCL_MIN_CTAS", str(min(reserved, 4)))
        dist.init_process_group("nccl", device_id=dev)
        _lib.check(_lib.lib.mico_set_reserved_sms(reserved), "mico_set_reserved_sms")
    B = args.batch
    torch.manual_seed(0)
    with torch.device(dev):     # random-init weights of the named architecture, created on the device
        tower = EVAVisionTransformer(**VIT_G)
    tower = tower.train()
    gen = torch.Generator().manual_seed(1234 + rank)
    host_pixels = torch.randn(B, 3, 224, 224, generator=gen).pin_memory()
    dev_pixels = host_pixels.to(dev)
    host_loss = torch.empty((), dtype=torch.float32).pin_memory()
    comm = torch.cuda.Stream(device=dev) if world > 1 else None

    bucket_blocks = int(os.environ.get("MICO_BENCH_BUCKET_BLOCKS", "5"))   # 5 blocks ~ 0.5 GB per all-reduce: 95.9 vs 94.3 % at N=2
    if world > 1:
        pending = []
        hook_calls = [0]

        def bucket_hook(bucket):
            """DP gradient SUM (pipeline.py:93-99 semantics: no divide), one NCCL all-reduce per `bucket_blocks` finished
            block buckets (contiguous in the flat gradient buffer, later blocks at higher addresses) on a side stream
            while the backward of the earlier blocks keeps the compute stream busy."""
            pending.append(bucket)
            is_last = bucket.data_ptr() == tower._last_flat_grad[0].data_ptr()
            tail = len(tower.blocks) - hook_calls[0] <= 2         # the last buckets stay small: their all-reduce is exposed
            hook_calls[0] = 0 if is_last else hook_calls[0] + 1
            if len(pending) < bucket_blocks and not is_last and not tail:
                return
            lo = min(b.data_ptr() for b in pending)
            n = sum(b.numel() for b in pending)
            flat = tower._last_flat_grad[0]
            off = (lo - flat.data_ptr()) // 4
            merged = flat[off:off + n]
            pending.clear()
            comm.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(comm):
                dist.all_reduce(merged)
        tower.grad_bucket_hook = bucket_hook

    def grad_sync():
        torch.cuda.current_stream().wait_stream(comm)

    def step(pixels):
        tower.invalidate_weight_cache()
        for p in tower.parameters():
            p.grad = None
        y = tower(pixels, return_all_features=True)
        loss = y.float().pow(2).mean()
        loss.backward()
        if world > 1:
            grad_sync()
        return loss

    # The reference's PrefetchLoader (data/utils/loader.py:100-142) copies the NEXT batch on a side stream while the
    # current step computes and hands it over with wait_stream + record_stream; same here: every step's pixels cross
    # PCIe from pinned host memory inside the timed region, one step ahead of their use.
    copy_stream = torch.cuda.Stream(device=dev)
    staged = []

    def stage_next():
        copy_stream.wait_stream(torch.cuda.current_stream())     # the buffer it replaces is no longer read
        with torch.cuda.stream(copy_stream):
            staged.append(host_pixels.to(dev, non_blocking=True))   # H2D of one step's inputs (pinned)

    def e2e_step():
        if not staged:
            stage_next()
        x = staged.pop(0)
        torch.cuda.current_stream().wait_stream(copy_stream)
        x.record_stream(torch.cuda.current_stream())
        stage_next()                                        # next step's inputs travel while this step computes
        loss = step(x)
        host_loss.copy_(loss.detach(), non_blocking=True)   # D2H read of the step's result
        torch.cuda.current_stream().synchronize()           # the reference loop's .item() (pipeline.py:47)
        return float(host_loss)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step(dev_pixels)
    barrier()

    # ---- timed region 1: inputs resident in HBM
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

    # ---- timed region 2: end to end through the module's public call, host buffers
    e2e_step()
    barrier()
    e0.record()
    for _ in range(args.steps):
        e2e_step()
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None

    # ---- per-family device time over the same K steps (cudaEvent pair around every entry point's launches)
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
    tokens_per_step = world * B * TOKENS_PER_IMAGE
    value = tokens_per_step * args.steps / (ms / 1e3)
    e2e = tokens_per_step * args.steps / (ms_e2e / 1e3)
    g = fam["gemm"]
    gemm_tflops = g["work"] / (g["ms"] * 1e-3) / 1e12 if g["ms"] > 0 else 0.0
    fam_total = sum(v["ms"] for v in fam.values())
    step_flops = 3 * FWD_FLOPS_PER_IMAGE * B
    traffic, traffic_src = None, None
    tpath = os.path.join(REPO, "profiles", "gemm_traffic.json")
    if os.path.exists(tpath):       # dram bytes per launch from the committed `ncu --set full` capture (scripts/ncu_to_traffic.py)
        tj = json.load(open(tpath))
        traffic, traffic_src = tj["avg_dram_bytes_per_launch"], tj["source"]
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
                                      rate=(v["work"] / (v["ms"] * 1e-3) / (1e12 if "attention" in k or k == "gemm" else 1e9))
                                      if v["ms"] > 0 else None,
                                      rate_unit="TFLOP/s" if ("attention" in k or k == "gemm") else "GB/s")
                              for k, v in fam.items()})
    cpu = cpu_baseline_leg() if (world == 1 and not args.no_cpu_baseline) else None
    line = dict(metric=METRIC, value=value, unit="tokens/s", n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3),
                ms_per_step=ms / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="bf16",
                data="synthetic",
                config=dict(workload=WORKLOAD,
                            batch_per_gpu=B, tokens_per_image=TOKENS_PER_IMAGE, drop_path_rate=0.4,
                            loss="tokens.pow(2).mean()", weight_cast_in_step=True,
                            e2e_inputs="pinned host pixels, H2D every step on a side stream one step ahead "
                                       "(the reference's PrefetchLoader, data/utils/loader.py:100-142)",
                            l2="working set per step (~35 GB of activations) exceeds the 126 MB L2; no flush needed",
                            grad_sync=(f"nccl all_reduce(SUM) per {bucket_blocks}-block bucket of the flat fp32 gradient buffer, overlapped "
                                       f"with backward on a side stream; {reserved} SMs reserved for NCCL (NCCL_MAX_CTAS)")
                            if world > 1 else "none (1 GPU)"),
                clocks=clocks,
                e2e=dict(value=e2e, unit="tokens/s", ms_per_step=ms_e2e / args.steps,
                         h2d_bytes_per_step=host_pixels.numel() * 4 * 1, d2h_bytes_per_step=4),
                gpu_launches=launches, roofline=roofline, cpu_baseline=cpu)
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
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_product(args)


if __name__ == "__main__":
    main()
