"""Image preprocessing: oracle vs the fixture produced by the reference's own ImageProcessor (CPU); CUDA kernel
vs fixture (GPU).  fp32 throughout: 1e-5 absolute on O(1) normalised pixels."""
import os

import pytest
import torch


def _gold(golden_dir):
    return torch.load(os.path.join(golden_dir, "imageproc.pt"), weights_only=False)


def test_oracle_matches_reference_image_processor(golden_dir):
    from oracle import imageproc as OI
    g = _gold(golden_dir)
    for img, enc, key, aa in (("big", "evaclip01_giant", "big_evaclip", True), ("big", "swin_base_22k_224", "big_swin", True),
                              ("small", "evaclip01_giant", "small_evaclip", True), ("big", "evaclip01_giant", "big_evaclip_noaa", False)):
        got = torch.from_numpy(OI.image_processor(g[img].numpy(), 224, enc, antialias=aa))
        assert got.shape == g[key].shape == (1, 3, 224, 224)
        assert (got - g[key]).abs().max().item() < 5e-5, key          # normalised pixels are O(1)-O(4): a few fp32 ulp


def test_video_frame_sampler_matches_reference(golden_dir):
    """split() against the reference's own videoprocessor.split; eval-mode sampling takes the middle frame of every segment."""
    pytest.importorskip("mico_b200._lib")
    from mico_b200.videoprocessor import sample_indices, split
    g = _gold(golden_dir)
    for (n, k), want in g["video_split"].items():
        got = split(list(range(n)), k)
        assert got == want, (n, k)
        mid = sample_indices(got, training=False)
        assert mid == [s[(len(s) + 1) // 2 - 1] for s in want]
        assert all(c in s for c, s in zip(sample_indices(got, training=True), got))


@pytest.mark.gpu
def test_cuda_video_processor_frames(golden_dir, tmp_path):
    """A directory of frames -> (sample_num, 3, R, R): every sampled frame equals the image processor's output for that file."""
    from PIL import Image
    from mico_b200.imageprocessor import ImageProcessor
    from mico_b200.videoprocessor import VideoProcessor
    g = _gold(golden_dir)
    d = tmp_path / "clip"
    d.mkdir()
    for i in range(6):
        Image.fromarray(g["big"].roll(7 * i, 1).numpy()).save(str(d / f"img_{i:04d}.png"))
    vp = VideoProcessor(224, "evaclip01_giant", sample_num=3, data_format="frame", training=False, antialias=True)
    out = vp(str(d))
    assert out.shape == (3, 3, 224, 224)
    ip = ImageProcessor(224, "evaclip01_giant", antialias=True)
    for j, i in enumerate((0, 2, 4)):       # middle frames of the segments [0,1] [2,3] [4,5]
        assert torch.equal(out[j], ip(str(d / f"img_{i:04d}.png"))[0])
    assert vp(str(tmp_path / "missing")) is None


@pytest.mark.gpu
def test_cuda_image_processor_matches_reference(golden_dir, tmp_path):
    from mico_b200.imageprocessor import ImageProcessor, resize_normalize
    g = _gold(golden_dir)
    for img, enc, key, aa in (("big", "evaclip01_giant", "big_evaclip", True), ("big", "swin_base_22k_224", "big_swin", True),
                              ("small", "evaclip01_giant", "small_evaclip", True), ("big", "evaclip01_giant", "big_evaclip_noaa", False)):
        proc = ImageProcessor(224, enc, antialias=aa)
        got = proc.process_uint8(g[img]).cpu()
        assert got.shape == (1, 3, 224, 224)
        assert (got - g[key]).abs().max().item() < 5e-5, key          # normalised pixels are O(1)-O(4): a few fp32 ulp
    # file path + video-style batch of frames + fp32 CHW input
    from PIL import Image
    f = str(tmp_path / "x.png")
    Image.fromarray(g["big"].numpy()).save(f)
    one = ImageProcessor(224, "evaclip01_giant", antialias=True)(f).cpu()
    assert (one - g["big_evaclip"]).abs().max().item() < 1e-5
    frames = torch.stack([g["big"], g["big"].flip(1)])
    two = ImageProcessor(224, "evaclip01_giant", antialias=True).process_uint8(frames).cpu()
    assert (two[0] - g["big_evaclip"][0]).abs().max().item() < 1e-5
    assert (two[1] - g["big_evaclip"][0].flip(2)).abs().max().item() < 1e-4       # mirrored taps: same weights, other order
    chw = (g["big"].permute(2, 0, 1).float() / 255).cuda()
    three = resize_normalize(chw[None], (224, 224), [0.48145466, 0.4578275, 0.40821073], [0.26862954, 0.26130258, 0.27577711], True).cpu()
    assert (three - g["big_evaclip"]).abs().max().item() < 1e-5
