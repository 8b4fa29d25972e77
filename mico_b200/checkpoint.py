"""Checkpoint interop: load weights the reference wrote, and write checkpoints it can read.

Mirrors inference_demo.py:14-117 (`load_from_pretrained_dir`), model/mico.py:250-321 (`modify_checkpoint`: key remap,
nearest-neighbour resize of the frame embeddings, bilinear resize of the ViT position embedding) and
data/utils/save.py:9-41 (`ModelSaver`).  Host-side logic on small tensors -- plain torch on whatever device the
checkpoint lives on; nothing here is on the hot path.
"""
import json
import os
from collections import defaultdict
from os.path import join

import torch
import torch.nn.functional as F


class _AttrDict(dict):
    """easydict-style attribute access (the reference wraps hps.json in an EasyDict, inference_demo.py:17)."""

    def __init__(self, d=None):
        super().__init__()
        for k, v in (d or {}).items():
            self[k] = _AttrDict(v) if isinstance(v, dict) else v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v


def modify_checkpoint(checkpoint, config):
    """model/mico.py:250-321 == inference_demo.py:29-95.  Returns a new dict; tensors that need no change are shared."""
    new_ckpt = {}
    for k, v in checkpoint.items():
        if 'video' in k:
            new_ckpt[k.replace('video', 'vision')] = v
        elif 'evaclip_model' in k:
            new_ckpt[k.replace('evaclip_model', 'vision_encoder')] = v
        elif 'clip_model' in k:
            new_ckpt[k.replace('clip_model', 'vision_encoder')] = v
        else:
            new_ckpt[k] = v.float()       # only the untouched keys are cast (mico.py:260)
    checkpoint = new_ckpt

    def resize_frames(key, n):
        e = checkpoint[key]
        if e.shape[1] != n:
            checkpoint[key] = F.interpolate(e.permute(0, 2, 1), n, mode='nearest').permute(0, 2, 1)

    if config.frame_embedding_type == 'adaptive':
        if 'vision_frame_embedding' in checkpoint:
            resize_frames('vision_frame_embedding', config.max_vision_sample_num)
        else:
            resize_frames('vision_perceiver.vision_frame_embedding', config.max_vision_sample_num)
        if 'audio_frame_embedding' in checkpoint:
            resize_frames('audio_frame_embedding', config.max_audio_sample_num)

    def resize_grid(src, patch):
        """[1 + g*g, C] -> [1 + G*G, C], cls row kept, patch rows resampled bilinearly on the 2-D grid."""
        width = src.shape[-1]
        grid = round((src.shape[0] - 1) ** 0.5)
        new_grid = config.vision_resolution // patch
        if new_grid == grid:
            return None
        oth = F.interpolate(src[1:].reshape(grid, grid, width).permute(2, 0, 1).unsqueeze(0), (new_grid, new_grid),
                            mode='bilinear')
        oth = oth[0].permute(1, 2, 0).reshape(-1, width)
        return torch.cat((src[0:1], oth), dim=0)

    if config.vision_encoder_type.startswith('clip'):
        key = "vision_encoder.visual.positional_embedding"
        tgt = resize_grid(checkpoint[key], checkpoint["vision_encoder.visual.conv1.weight"].shape[-1])
        if tgt is not None:
            checkpoint[key] = tgt
    elif config.vision_encoder_type.startswith('evaclip'):
        key = "vision_encoder.visual.pos_embed"
        tgt = resize_grid(checkpoint[key][0], checkpoint["vision_encoder.visual.patch_embed.proj.weight"].shape[-1])
        if tgt is not None:
            checkpoint[key] = tgt.unsqueeze(0)
    return checkpoint


def load_from_pretrained_dir(pretrain_dir, video_resolution=224, return_modal="full"):
    """inference_demo.py:14-117: newest `ckpt/model_step_<n>.pt` of a reference run directory + `log/hps.json`.
    Returns (state_dict, model_cfg) ready for `MiCo.from_pretrained(model_cfg, state_dict)`."""
    checkpoint_dir = os.path.join(pretrain_dir, 'ckpt')
    file_cfg = _AttrDict(json.load(open(os.path.join(pretrain_dir, 'log', 'hps.json'))))
    model_cfg = file_cfg.model_cfg
    steps = sorted(int(i.split('_')[2].split('.')[0]) for i in os.listdir(checkpoint_dir) if i.startswith('model_step'))
    if not steps:
        raise FileNotFoundError(f"no model_step_*.pt under {checkpoint_dir}")
    ckpt_file = os.path.join(checkpoint_dir, 'model_step_' + str(steps[-1]) + '.pt')
    checkpoint = torch.load(ckpt_file, map_location='cpu')
    print(f'load_from_pretrained: {ckpt_file}')
    checkpoint = modify_checkpoint(checkpoint, model_cfg)
    if return_modal == "full":
        new_ckpt = checkpoint
    elif return_modal == "uni":
        new_ckpt = defaultdict()
        for k in checkpoint.keys():
            if "video_encoder" in k:
                new_ckpt[".".join(k.split(".")[1:])] = checkpoint[k]
    elif return_modal == "text":
        new_ckpt = defaultdict()
        for k in checkpoint.keys():
            if "multimodal_encoder" in k:
                new_ckpt[".".join(k.split(".")[1:])] = checkpoint[k]
    else:
        new_ckpt = checkpoint
    return new_ckpt, model_cfg


class ModelSaver(object):
    """data/utils/save.py:9-41: `model_step_<n>.pt` (CPU state_dict), optional `best_<k>.pt`, `optimizer_step_<n>.pt`."""

    def __init__(self, output_dir, prefix='model_step', suffix='pt', remove_before_ckpt=True):
        self.output_dir = output_dir
        self.prefix = prefix
        self.suffix = suffix
        self.remove_before_ckpt = remove_before_ckpt

    def save(self, model, step, optimizer=None, best_indicator=None, save_best=False):
        previous_state = [i for i in os.listdir(self.output_dir) if i.startswith('model')]
        if self.remove_before_ckpt:
            for p in previous_state:
                os.remove(os.path.join(self.output_dir, p))
        output_model_file = join(self.output_dir, f"{self.prefix}_{step}.{self.suffix}")
        state_dict = {k: v.cpu() if isinstance(v, torch.Tensor) else v for k, v in model.state_dict().items()}
        torch.save(state_dict, output_model_file)
        if save_best:
            for k in best_indicator:
                if best_indicator[k]:
                    torch.save(state_dict, join(self.output_dir, f"best_{k}.{self.suffix}"))
        if optimizer is not None:
            previous_state = [i for i in os.listdir(self.output_dir) if i.startswith('optimizer')]
            if self.remove_before_ckpt:
                for p in previous_state:
                    os.remove(os.path.join(self.output_dir, p))
            torch.save(optimizer.state_dict(), f'{self.output_dir}/optimizer_step_{step}.pt')
