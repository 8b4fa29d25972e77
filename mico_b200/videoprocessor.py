"""GPU mirror of model/videoprocessor.py: sample `sample_num` frames of a clip (one per equal segment: random within the
segment when training, the middle frame otherwise), then ToTensor + Resize + Normalize for all frames in ONE launch of the
image-preprocessing kernel (csrc/imageproc.cu) instead of per-frame CPU transforms.

data_format 'frame' reads a directory of frame images (the reference's branch tests an undefined variable,
videoprocessor.py:59, and therefore always returns None; the intended behaviour is implemented here); 'raw' decodes with
decord when it is installed (videoprocessor.py:73-90) and raises otherwise."""
import os
import random

import torch

from ._lib import MicoError
from .imageprocessor import _CLIP_MEAN, _CLIP_STD, _INET_MEAN, _INET_STD, _default_antialias, resize_normalize


def split(frame_name_lists, sample_num):
    """videoprocessor.py:11-15: `sample_num` contiguous, near-equal segments; short clips are padded with the last frame."""
    frames = list(frame_name_lists)
    if len(frames) < sample_num:
        frames += [frames[-1]] * (sample_num - len(frames))
    k, m = divmod(len(frames), sample_num)
    return [frames[i * k + min(i, m):(i + 1) * k + min(i + 1, m)] for i in range(sample_num)]


def sample_indices(segments, training):
    """videoprocessor.py:66-69 / 81-84."""
    if training:
        return [random.choice(s) for s in segments]
    return [s[(len(s) + 1) // 2 - 1] for s in segments]


class VideoProcessor(object):
    def __init__(self, video_resolution, video_encoder_type, sample_num=4, video_transforms='none', data_format="frame",
                 training=True, device="cuda", antialias=None):
        self.training = training
        self.sample_num = sample_num
        self.data_format = data_format
        self.resolution = video_resolution
        self.video_encoder_type = video_encoder_type
        if video_encoder_type.startswith('clip') or video_encoder_type.startswith('evaclip'):
            self.mean, self.std = _CLIP_MEAN, _CLIP_STD
        else:
            self.mean, self.std = _INET_MEAN, _INET_STD
        if video_transforms != 'none':
            raise NotImplementedError("mico_b200.VideoProcessor implements video_transforms='none' (Resize + Normalize)")
        self.device = device
        self.antialias = _default_antialias() if antialias is None else bool(antialias)

    def process_uint8(self, frames):
        """frames: uint8 [n, H, W, 3] -> fp32 [n, 3, R, R] on self.device."""
        return resize_normalize(frames.to(self.device, non_blocking=True), (self.resolution, self.resolution), self.mean,
                                self.std, self.antialias)

    def __call__(self, video_file):
        try:
            if not os.path.exists(video_file):
                print('not have videos', video_file)
                return None
            import numpy as np
            if self.data_format == 'frame':
                from PIL import Image
                names = sorted(os.listdir(video_file))
                picked = sample_indices(split(names, self.sample_num), self.training)
                arrs = [np.asarray(Image.open(os.path.join(video_file, n)).convert('RGB')) for n in picked]
                if len({a.shape for a in arrs}) != 1:
                    raise MicoError("frames of one clip must share a size")
                frames = torch.from_numpy(np.stack(arrs))
            elif self.data_format == 'raw':
                try:
                    import decord
                except ImportError as e:
                    raise MicoError("data_format='raw' needs decord (model/videoprocessor.py:3)") from e
                container = decord.VideoReader(uri=video_file)
                picked = sample_indices(split(list(range(len(container))), self.sample_num), self.training)
                frames = torch.from_numpy(container.get_batch(picked).asnumpy())
            else:
                raise NotImplementedError(self.data_format)
            return self.process_uint8(frames)
        except Exception as e:       # the reference swallows decode errors the same way (videoprocessor.py:99-102)
            print(e)
            print(video_file)
            return None
