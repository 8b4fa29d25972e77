"""GPU mirror of model/imageprocessor.py (and the per-frame transform of model/videoprocessor.py): same constructor, same
mean / std tables, same call (`proc(image_file)` -> (1, 3, R, R) fp32, normalised), but decode -> uint8 HWC upload ->
ONE CUDA kernel (ToTensor + Resize + Normalize, csrc/imageproc.cu) instead of three CPU passes.

`antialias`: torchvision's `Resize` on a tensor anti-aliases by default since 0.17; the reference pins torchvision 0.15.2
(set_env.sh), where it does not.  Default None = follow the installed torchvision's default, so that outputs equal the
reference code run in the same environment; pass False for the pinned-version behaviour."""
import ctypes as C
import os

import torch

from ._lib import MicoError, check, lib

_CLIP_MEAN, _CLIP_STD = [0.48145466, 0.4578275, 0.40821073], [0.26862954, 0.26130258, 0.27577711]
_INET_MEAN, _INET_STD = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]


def _default_antialias():
    try:
        import inspect
        from torchvision.transforms.transforms import Resize
        return inspect.signature(Resize.__init__).parameters["antialias"].default is True
    except Exception:
        return False      # no torchvision to ask: the reference's pinned torchvision 0.15.2 does not anti-alias tensors


def resize_normalize(src, size, mean, std, antialias=True):
    """src: CUDA uint8 [n, H, W, C] (decoder layout) or fp32 [n, C, H, W]; returns fp32 [n, C, size[0], size[1]]."""
    if not src.is_cuda:
        raise MicoError("resize_normalize runs on CUDA tensors only (no CPU fallback)")
    if src.dim() == 3:
        src = src.unsqueeze(0)
    src = src.contiguous()
    if src.dtype == torch.uint8:
        n, H, W, Cc = src.shape
        u8 = 1
    elif src.dtype == torch.float32:
        n, Cc, H, W = src.shape
        u8 = 0
    else:
        raise MicoError("resize_normalize: uint8 HWC or fp32 CHW input")
    Ho, Wo = int(size[0]), int(size[1])
    out = torch.empty((n, Cc, Ho, Wo), device=src.device, dtype=torch.float32)
    m = (C.c_float * Cc)(*[float(x) for x in mean[:Cc]])
    s = (C.c_float * Cc)(*[float(x) for x in std[:Cc]])
    stream = C.c_void_p(torch.cuda.current_stream(src.device).cuda_stream)
    check(lib.mico_resize_normalize(C.c_void_p(src.data_ptr()), u8, n, Cc, H, W, C.c_int64(H * W * Cc), C.c_void_p(out.data_ptr()),
                                    Ho, Wo, m, s, int(bool(antialias)), stream), "mico_resize_normalize")
    return out


class ImageProcessor(object):
    def __init__(self, image_resolution, image_encoder_type, image_transforms='none', training=True, device="cuda",
                 antialias=None):
        self.training = training
        self.resolution = image_resolution
        self.image_encoder_type = image_encoder_type
        if image_encoder_type.startswith('clip') or image_encoder_type.startswith('evaclip'):
            self.mean, self.std = _CLIP_MEAN, _CLIP_STD
        else:
            self.mean, self.std = _INET_MEAN, _INET_STD
        self.image_transforms = image_transforms
        if image_transforms != 'none':      # 'crop_flip' (RandomResizedCrop + flip) is a training-time augmentation
            raise NotImplementedError("mico_b200.ImageProcessor implements image_transforms='none' (Resize + Normalize)")
        self.device = device
        self.antialias = _default_antialias() if antialias is None else bool(antialias)

    def process_uint8(self, frames):
        """frames: uint8 [H, W, 3] or [n, H, W, 3] (any device) -> fp32 [n, 3, R, R] on self.device."""
        return resize_normalize(frames.to(self.device, non_blocking=True), (self.resolution, self.resolution), self.mean,
                                self.std, self.antialias)

    def __call__(self, image_file):
        try:
            if not os.path.exists(image_file):
                print('not have image', image_file)
                return None
            from PIL import Image
            import numpy as np
            img = Image.open(image_file).convert('RGB')
            return self.process_uint8(torch.from_numpy(np.asarray(img).copy()))
        except Exception as e:       # the reference swallows decode errors the same way (imageprocessor.py:62-64)
            print(e)
            return None
