"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference image preprocessing.

model/imageprocessor.py:24-29,52-56: ToTensor (uint8 HWC / 255 -> fp32 CHW) -> Resize((R, R)) -> Normalize(mean, std).  Resize on
a tensor is torch.nn.functional.interpolate(mode='bilinear', align_corners=False, antialias=<torchvision default>); the
interpolation itself lives in ATen (aten/src/ATen/native/cpu/UpSampleKernel.cpp), which is restated here dimension by dimension
in plain numpy: plain bilinear (area_pixel_compute_source_index) and the anti-aliased separable triangle filter
(_compute_indices_weights_aa).  Pinned by tests/golden/imageproc.pt, produced by the reference's own ImageProcessor on PNG files.
"""
import numpy as np

CLIP_MEAN, CLIP_STD = [0.48145466, 0.4578275, 0.40821073], [0.26862954, 0.26130258, 0.27577711]
INET_MEAN, INET_STD = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]


def _weights_1d(n_in, n_out, antialias):
    """(n_out, n_in) fp32 interpolation matrix along one dimension."""
    M = np.zeros((n_out, n_in), np.float32)
    scale = np.float32(n_in) / np.float32(n_out)
    for i in range(n_out):
        if antialias:
            support = scale if scale >= 1 else np.float32(1)
            centre = scale * np.float32(i + 0.5)
            xmin = max(0, int(centre - support + np.float32(0.5)))
            xsize = min(n_in, int(centre + support + np.float32(0.5))) - xmin
            inv = np.float32(1) / scale if scale >= 1 else np.float32(1)
            w = np.array([max(0.0, 1.0 - abs((j + xmin - centre + np.float32(0.5)) * inv)) for j in range(xsize)], np.float32)
            M[i, xmin:xmin + xsize] = w / w.sum(dtype=np.float32)
        else:
            f = max(np.float32(0), scale * np.float32(i + 0.5) - np.float32(0.5))
            x0 = min(int(f), n_in - 1)
            x1 = min(x0 + 1, n_in - 1)
            lam = np.float32(f - x0)
            M[i, x0] += np.float32(1) - lam
            M[i, x1] += lam
    return M


def image_processor(u8_hwc, resolution, encoder_type, antialias=True):
    """uint8 [H, W, 3] -> fp32 [1, 3, R, R], as ImageProcessor.__call__ returns it (image_transforms='none')."""
    x = u8_hwc.astype(np.float32).transpose(2, 0, 1) / np.float32(255)        # ToTensor
    H, W = x.shape[1:]
    My, Mx = _weights_1d(H, resolution, antialias), _weights_1d(W, resolution, antialias)
    r = np.stack([(x[c] @ Mx.T) for c in range(x.shape[0])])                  # horizontal pass first (ATen order), fp32
    r = np.stack([(My @ r[c]) for c in range(r.shape[0])]).astype(np.float32)  # then vertical
    clip = encoder_type.startswith("clip") or encoder_type.startswith("evaclip")
    mean, std = (CLIP_MEAN, CLIP_STD) if clip else (INET_MEAN, INET_STD)
    mean, std = np.array(mean, np.float32)[:, None, None], np.array(std, np.float32)[:, None, None]
    return ((r - mean) / std)[None]
