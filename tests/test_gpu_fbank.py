"""GPU parity of the fused Kaldi fbank kernel (mico_b200.audioprocessor / csrc/fbank.cu) against torchaudio's own
compliance.kaldi.fbank output (tests/golden/fbank.pt) and the oracle's restatement of audioprocessor.py:38-70.
Tolerance: log-mel values are O(10); 2e-3 absolute (fp32 FFT in a different summation order, fast log)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_fbank_matches_torchaudio_golden(golden_dir):
    from mico_b200 import ops
    from mico_b200.audioprocessor import kaldi_mel_banks, povey_window
    g = torch.load(os.path.join(golden_dir, "fbank.pt"), weights_only=False)
    for bins in (224, 64):
        fb = ops.fbank(g["wave"].cuda(), povey_window().cuda(), kaldi_mel_banks(bins).cuda())[0]
        ref = g[f"fbank_{bins}"]
        err = (fb.cpu() - ref).abs().max().item()
        print(f"fbank {bins} bins: max abs err {err:.2e} (values {ref.min():.1f}..{ref.max():.1f})")
        assert fb.shape == ref.shape and err < 2e-3


def test_audio_processor_matches_oracle(golden_dir):
    from mico_b200.audioprocessor import AudioProcessor
    from oracle import fbank as OF
    g = torch.load(os.path.join(golden_dir, "fbank.pt"), weights_only=False)
    wave = torch.cat([g["wave"], g["wave"].flip(1), g["wave"], g["wave"] * 0.5], dim=1)[:, :160000]     # 10 s
    proc = AudioProcessor(melbins=224, target_length=224, sample_num=3, training=False)
    out = proc(wave)
    ref = OF.audio_processor(wave, 224, 224, 3)
    assert out.shape == (3, 224, 224)
    assert (out.cpu() - ref).abs().max().item() < 2e-4        # normalised by 1 / 13.1
    proc64 = AudioProcessor(melbins=64, target_length=224, sample_num=3, training=False)      # bilinear resize to 224 bins
    ref64 = OF.audio_processor(wave, 64, 224, 3)
    assert (proc64(wave).cpu() - ref64).abs().max().item() < 2e-4
