"""CPU: the reference arm of bench.py (`--impl reference`, the oracle port of the reference's CPU eager path) prints ONE JSON line
with the contract's keys and the same metric / config as the product arm; host-side workload arithmetic."""
import json
import os
import subprocess
import sys
import types

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    r = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=1200, cwd=REPO,
                       env=dict(os.environ, MICO_BENCH_REF_BUDGET_S="20"))
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    sys.path.insert(0, REPO)
    import bench
    assert d["impl"] == "reference" and d["metric"] == bench.METRIC["omni"] and d["unit"] == "tokens/s"
    assert d["config"]["workload"] == bench.WORKLOAD["omni"] and d["higher_is_better"] is True
    assert "configs[4]" in d["config"]["workload"] and d["config"]["text_len"] == 128 and d["config"]["batch_per_gpu"] == 64
    # the product arm emits the identical config dict (the driver's same_config check)
    args = types.SimpleNamespace(config="omni", batch=64, gpus=1, no_ckpt=False, light_blocks=bench.DEFAULT_LIGHT_BLOCKS,
                                 qkv_blocks=bench.DEFAULT_QKV_BLOCKS, attn_blocks=bench.DEFAULT_ATTN_BLOCKS,
                                 grad_dtype="fp32", bucket_blocks=5)
    assert d["config"] == bench.make_config(args)
    assert d["value"] > 0 and d["steps"] == 1 and d["n_gpus"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == dict(value=d["value"], unit="tokens/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0)


def test_workload_arithmetic():
    sys.path.insert(0, REPO)
    import bench
    ns, processed = bench.omni_tokens(64)
    assert ns == 64 * (1568 + 4 * 257 + 128) and processed == 64 * (12 * 257 + 128)
    tower, fusion, text = bench.omni_flops(64)
    assert abs(tower / (64 * 12) - 534.06e9) < 0.01e9            # 534.06 GFLOP per frame forward
    # SURVEY.md 8d BERT figures per sample at S = 128: text-only 22.35 G, cross-attention to 257 keys 34.46 G, + LM head 40.61 G
    t1 = bench.omni_flops(1, n_v=1, n_a=0, n_d=0)[2]
    assert text > 0 and t1 > 0
