"""Audio front-end on the GPU: mirror of the reference's ``model/audioprocessor.py`` ``AudioProcessor`` (:15-77).

Same constructor (``melbins, target_length, sample_num, frame_shift, resize_melbin_num, mean, std, training``) and the
same output ``(sample_num, target_length, melbins)`` spectrogram slices, but the Kaldi log-mel filterbank
(``torchaudio.compliance.kaldi.fbank``, :40) and the normalisation (:47) run in one CUDA kernel (csrc/fbank.cu) on a
waveform already resident on the device; decoding / resampling the file stays host I/O (out of scope).
``__call__`` accepts a waveform tensor ``(channels, samples)`` or ``(samples,)`` at 16 kHz, or a file path when
``torchaudio.load`` is usable.
"""
import math
import os
import random

import torch

from . import ops
from .ops import F32, MicoError


def split(frame_name_lists, sample_num):
    """audioprocessor.py:8-12"""
    if len(frame_name_lists) < sample_num:
        frame_name_lists += [frame_name_lists[-1]] * (sample_num - len(frame_name_lists))
    k, m = divmod(len(frame_name_lists), sample_num)
    return [frame_name_lists[i * k + min(i, m):(i + 1) * k + min(i + 1, m)] for i in list(range(sample_num))]


def povey_window(n=400):
    """hann(n, periodic=False) ** 0.85 (Kaldi 'povey' window)"""
    return torch.hann_window(n, periodic=False, dtype=torch.float64).pow(0.85).float()


def kaldi_mel_banks(num_bins, padded=512, sample_freq=16000.0, low_freq=20.0, high_freq=0.0):
    """Kaldi triangular mel filters on the 1127*ln(1+f/700) scale -> [num_bins, padded/2 + 1] (last column zero)."""
    nyq = 0.5 * sample_freq
    if high_freq <= 0.0:
        high_freq += nyq
    ms = lambda f: 1127.0 * math.log(1.0 + f / 700.0)
    lo, hi = ms(low_freq), ms(high_freq)
    delta = (hi - lo) / (num_bins + 1)
    b = torch.arange(num_bins, dtype=torch.float32).unsqueeze(1)
    left, center, right = lo + b * delta, lo + (b + 1.0) * delta, lo + (b + 2.0) * delta
    mel = (1127.0 * (1.0 + (sample_freq / padded) * torch.arange(padded // 2, dtype=torch.float32) / 700.0).log()).unsqueeze(0)
    up, down = (mel - left) / (center - left), (right - mel) / (right - center)
    bins = torch.clamp(torch.min(up, down), min=0.0)
    return torch.nn.functional.pad(bins, (0, 1)).contiguous()


class AudioProcessor(object):
    def __init__(self, melbins, target_length, sample_num, frame_shift=10, resize_melbin_num=224, mean=15.41663,
                 std=6.55582, training=True, device="cuda"):
        self.melbins, self.target_length, self.training = melbins, target_length, training
        self.frame_shift, self.sample_num, self.resize_melbin_num = frame_shift, sample_num, resize_melbin_num
        self.mean, self.std = mean, std
        self.device = torch.device(device)
        self._window = povey_window(400)
        self._mel = kaldi_mel_banks(melbins)

    def fbank(self, waveform):
        """(samples,) or (channels, samples) fp32 at 16 kHz -> normalised log-mel (frames, melbins) on the device."""
        w = waveform if waveform.dim() == 2 else waveform.unsqueeze(0)
        w = w[:1].to(self.device, F32).contiguous()              # kaldi.fbank uses channel 0
        if self._window.device != w.device:
            self._window, self._mel = self._window.to(w.device), self._mel.to(w.device)
        if self.melbins == self.resize_melbin_num:
            return ops.fbank(w, self._window, self._mel, frame_shift=160, norm_sub=self.mean, norm_mul=1.0 / (self.std * 2))[0]
        fb = ops.fbank(w, self._window, self._mel, frame_shift=160)[0]
        fb = torch.nn.functional.interpolate(fb.reshape(1, 1, *fb.shape), size=(fb.size(0), self.resize_melbin_num),
                                             mode='bilinear').reshape(fb.size(0), self.resize_melbin_num)
        return (fb - self.mean) / (self.std * 2)

    def batch(self, waveforms):
        """(clips, samples) fp32 waveforms at 16 kHz -> (clips, sample_num, target_length, melbins): `__call__` for a whole
        batch in ONE fbank launch (the reference runs its CPU front-end clip by clip in the data-loader workers,
        audioprocessor.py:28-77).  Same padding (:54) and slice choice (:60-68) per clip."""
        if self.melbins != self.resize_melbin_num:
            return torch.stack([self(w) for w in waveforms], dim=0)
        w = waveforms.to(self.device, F32).contiguous()
        if self._window.device != w.device:
            self._window, self._mel = self._window.to(w.device), self._mel.to(w.device)
        fb = ops.fbank(w, self._window, self._mel, frame_shift=160, norm_sub=self.mean, norm_mul=1.0 / (self.std * 2))
        src_length = fb.shape[1]
        pad_len = max(self.target_length * self.sample_num - src_length,
                      self.target_length - src_length % self.target_length)
        n_slices = (src_length + pad_len) // self.target_length
        slices = split(list(range(n_slices)), self.sample_num)
        out = torch.zeros((w.shape[0], n_slices * self.target_length, self.melbins), device=w.device, dtype=F32)
        out[:, :src_length] = fb
        out = out.view(w.shape[0], n_slices, self.target_length, self.melbins)
        if self.training:
            idx = torch.tensor([[random.choice(i) for i in slices] for _ in range(w.shape[0])], device=w.device)
        else:
            idx = torch.tensor([[i[(len(i) + 1) // 2 - 1] for i in slices]] * w.shape[0], device=w.device)
        return out[torch.arange(w.shape[0], device=w.device).unsqueeze(1), idx]

    def __call__(self, wav):
        if isinstance(wav, str):
            if not os.path.exists(wav):
                print('not have audios', wav)
                return torch.zeros(self.sample_num, self.target_length, self.melbins)
            import torchaudio
            waveform, sr = torchaudio.load(wav)
            if sr != 16000:
                waveform = torchaudio.transforms.Resample(sr, 16000)(waveform)
        else:
            waveform = wav
        fbank = self.fbank(waveform)
        src_length = fbank.shape[0]
        pad_len = max(self.target_length * self.sample_num - src_length,
                      self.target_length - src_length % self.target_length)      # audioprocessor.py:54
        fbank = torch.nn.functional.pad(fbank, (0, 0, 0, pad_len))
        slices = split(list(range(fbank.shape[0] // self.target_length)), self.sample_num)
        if self.training:
            idx = [random.choice(i) for i in slices]
        else:
            idx = [i[(len(i) + 1) // 2 - 1] for i in slices]
        return torch.stack([fbank[i * self.target_length:(i + 1) * self.target_length] for i in idx], dim=0)
