"""Checkpoint interop, CPU: modify_checkpoint vs a fixture produced by the reference's own
MMGeneralModule.modify_checkpoint; ModelSaver -> load_from_pretrained_dir round trip in the reference's directory layout."""
import json
import os

import torch

GOLD = os.path.join(os.path.dirname(__file__), "golden", "modify_checkpoint.pt")


def test_modify_checkpoint_matches_reference_fixture():
    from mico_b200.checkpoint import _AttrDict, modify_checkpoint
    cases = torch.load(GOLD, weights_only=False)
    for name, c in cases.items():
        out = modify_checkpoint({k: v.clone() for k, v in c["inp"].items()}, _AttrDict(c["cfg"]))
        assert set(out.keys()) == set(c["out"].keys()), name
        for k, want in c["out"].items():
            assert out[k].dtype == want.dtype and out[k].shape == want.shape, (name, k)
            assert torch.equal(out[k], want), (name, k)       # same torch ops on the same inputs: bit-exact


def test_saver_and_pretrained_dir_round_trip(tmp_path):
    from mico_b200.checkpoint import ModelSaver, load_from_pretrained_dir
    run = tmp_path / "MiCo-tiny"
    (run / "ckpt").mkdir(parents=True)
    (run / "log").mkdir()
    cfg = dict(model_cfg=dict(frame_embedding_type='adaptive', max_vision_sample_num=4, max_audio_sample_num=2,
                              vision_encoder_type='evaclip01_giant', vision_resolution=56))
    json.dump(cfg, open(run / "log" / "hps.json", "w"))

    class Tiny(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.vision_frame_embedding = torch.nn.Parameter(torch.randn(1, 2, 6))
            self.audio_frame_embedding = torch.nn.Parameter(torch.randn(1, 2, 6))
            self.vision_encoder = torch.nn.Module()
            self.vision_encoder.visual = torch.nn.Module()
            self.vision_encoder.visual.pos_embed = torch.nn.Parameter(torch.randn(1, 1 + 4, 8))
            self.vision_encoder.visual.patch_embed = torch.nn.Module()
            self.vision_encoder.visual.patch_embed.proj = torch.nn.Conv2d(3, 8, 14, 14)

    m = Tiny()
    opt = torch.optim.SGD(m.parameters(), lr=0.1)
    saver = ModelSaver(str(run / "ckpt"))
    saver.save(m, 10, optimizer=opt)
    saver.save(m, 200, optimizer=opt)          # removes step 10 (remove_before_ckpt)
    assert sorted(os.listdir(run / "ckpt")) == ["model_step_200.pt", "optimizer_step_200.pt"]
    sd, model_cfg = load_from_pretrained_dir(str(run))
    assert model_cfg.vision_resolution == 56 and model_cfg["max_vision_sample_num"] == 4
    assert sd["vision_frame_embedding"].shape == (1, 4, 6)                        # nearest: 2 -> 4 frames
    assert torch.equal(sd["vision_frame_embedding"][0, ::2], m.vision_frame_embedding.detach()[0])
    assert sd["vision_encoder.visual.pos_embed"].shape == (1, 1 + 16, 8)          # 2x2 grid -> 4x4 (56 / 14)
    assert torch.equal(sd["vision_encoder.visual.pos_embed"][0, 0], m.vision_encoder.visual.pos_embed.detach()[0, 0])
    assert torch.equal(sd["audio_frame_embedding"], m.audio_frame_embedding.detach())
