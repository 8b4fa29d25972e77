"""Full-size omni-modal pretraining step (BASELINE.json configs[4] shape, per-GPU slice): ViT-g/14 tower shared by image /
video / audio / depth, BERT-base fusion encoder, ITC + ITM + caption losses, through mico_b200.mico.MiCo.forward.
Developer / evidence tool (prints one JSON line; not the contract bench):  python scripts/omni_step.py [b] [steps]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from mico_b200.mico import MiCo, _AttrDict


def main():
    b = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    ckpt = (sys.argv[3] == "ckpt") if len(sys.argv) > 3 else b > 12       # activation checkpointing in the tower (config.checkpointing)
    n_v, n_a, n_d, S = 8, 3, 1, int(os.environ.get("OMNI_S", "128"))
    cfg = _AttrDict(vision_encoder_type="evaclip01_giant", vision_resolution=224, checkpointing=ckpt, contra_dim=512,
                    max_vision_sample_num=8, max_audio_sample_num=3, max_depth_sample_num=1, beam_size=3, itm_ratio=0.1,
                    max_omni_caption_len=70, max_caption_len=S, max_subtitle_len=70, frame_embedding_type="adaptive",
                    pool_video=False)
    torch.manual_seed(0)
    dev = torch.device("cuda")
    with torch.device(dev):
        model = MiCo.from_pretrained(cfg, {})
    model = model.to(dev).train()
    g = torch.Generator().manual_seed(1234)
    lens = torch.randint(8, S + 1, (b,), generator=g)
    att = (torch.arange(S)[None] < lens[:, None]).long()
    ids = torch.randint(1000, 30522, (b, S), generator=g) * att
    ids[:, 0] = 101
    ids[torch.arange(b), lens - 1] = 102
    host = dict(vision_pixels=torch.randn(b, n_v, 3, 224, 224, generator=g).pin_memory(),
                audio_spectrograms=torch.randn(b, n_a, 224, 224, generator=g).pin_memory(),
                depth_pixels=torch.randn(b, n_d, 3, 224, 224, generator=g).pin_memory())
    task = "ret%tv%ta%tva%td_cap%tv%ta%tva"

    def step():
        batch = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
        batch["caption_tokens"] = _AttrDict(input_ids=ids.to(dev), attention_mask=att.to(dev))
        for p in model.parameters():
            p.grad = None
        out = model(batch, task, compute_loss=True)
        loss = sum(out.values())
        loss.backward()
        return {k: float(v) for k, v in out.items()}

    for _ in range(2):
        losses = step()
    torch.cuda.synchronize()
    t0 = time.time()
    for _ in range(steps):
        losses = step()        # float() of the losses synchronises, like the reference loop's .item()
    torch.cuda.synchronize()
    dt = (time.time() - t0) / steps
    from mico_b200 import ops
    ops.profile_enable(True)
    step()
    fam = ops.profile_collect()
    ops.profile_enable(False)
    device_ms = {k: round(v["ms"], 2) for k, v in fam.items()}
    device_calls = {k: int(v["calls"]) for k, v in fam.items()}
    frames = b * (n_v + n_a + n_d)
    text_tok = int(att.sum())
    print(json.dumps(dict(workload="omni-modal step: video n=8 + audio n=3 + depth n=1 + text S=" + str(S) + ", ViT-g/14 + BERT-base, task " + task,
                          samples_per_step=b, vit_frames_per_step=frames, tower_activation_checkpointing=ckpt, ms_per_step=dt * 1e3,
                          processed_tokens_per_s=(frames * 257 + text_tok) / dt,
                          north_star_tokens_per_s=b * (1568 + 3 * 257 + 257 + 128) / dt,
                          device_ms_by_family=device_ms, device_ms_total=round(sum(device_ms.values()), 2), launches_by_family=device_calls,
                          losses=losses, peak_mem_gb=torch.cuda.max_memory_allocated() / 1e9,
                          finite=all(x == x and abs(x) < 1e6 for x in losses.values()))))


if __name__ == "__main__":
    main()
