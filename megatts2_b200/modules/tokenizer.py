"""Mel front end with the reference's surface: ``extract_mel_spec`` and the HIFIGAN_*
constants of modules/tokenizer.py:19-24, 107-125 (speechbrain ``mel_spectogram`` ->
torchaudio ``MelSpectrogram`` -> log(clamp(., 1e-5))).  The G2P /
lhotse pieces of that file are host string processing and stay the reference's.
Device work: libmegatts2_b200 ``mtts_mel_spectrogram_f32``."""
import math

import numpy as np
import torch

from .. import ops

HIFIGAN_SR = 16000
HIFIGAN_HOP_LENGTH = 256
HIFIGAN_WIN_LENGTH = 1024
HIFIGAN_MEL_CHANNELS = 80
HIFIGAN_NFFT = 1024
HIFIGAN_MAX_FREQ = 8000


def slaney_mel_filterbank(n_freqs=HIFIGAN_NFFT // 2 + 1, f_min=0.0, f_max=float(HIFIGAN_MAX_FREQ),
                          n_mels=HIFIGAN_MEL_CHANNELS, sample_rate=HIFIGAN_SR):
    """(n_freqs, n_mels) fp32 triangular filterbank on the Slaney mel scale with Slaney
    (area) normalisation - what mel_scale="slaney", norm="slaney" request at tokenizer.py:121-122.
    Built in float64 on the host: mel(f) = 3f/200 below 1 kHz, 15 + 27 ln(f/1000)/ln 6.4 above."""
    knee_hz, knee_mel, slope, step = 1000.0, 15.0, 3.0 / 200.0, math.log(6.4) / 27.0
    to_mel = lambda f: f * slope if f < knee_hz else knee_mel + math.log(f / knee_hz) / step  # noqa: E731
    mel_pts = np.linspace(to_mel(f_min), to_mel(f_max), n_mels + 2)
    hz_pts = np.where(mel_pts < knee_mel, mel_pts / slope, knee_hz * np.exp(step * (mel_pts - knee_mel)))
    bins = np.linspace(0.0, sample_rate // 2, n_freqs)[:, None]
    left, centre, right = hz_pts[None, :-2], hz_pts[None, 1:-1], hz_pts[None, 2:]
    tri = np.maximum(np.minimum((bins - left) / (centre - left), (right - bins) / (right - centre)), 0.0)
    return (tri * (2.0 / (right - left))).astype(np.float32)


def pack_grouped_filterbank(fb: np.ndarray):
    """(n_bins, n_mels) dense filterbank -> (fb_w float32, fb_off int32 [G + 1], fb_start int32 [4 G]): the grouped banded
    form mtts_mel_spectrogram_f32 takes (include/megatts2_b200.h).  Mels go in groups of 4; each mel's band is read from
    a start bin rounded down to a multiple of 4 (shifted further down where it would run past bin n_bins + 2), for
    len_g taps = the group's longest aligned band rounded up to a multiple of 4.  A group's block is (4, len_g) row-major
    with zeros outside a mel's own band - the kernel's tap loop is uniform across the group and uses 128-bit reads."""
    n_bins, n_mels = fb.shape
    G = (n_mels + 3) // 4
    limit = (n_bins + 3) // 4 * 4                        # the kernel zeroes the magnitudes between n_bins and this
    lo = np.zeros(4 * G, dtype=np.int64)
    ln = np.zeros(4 * G, dtype=np.int64)
    for m in range(n_mels):
        nz = np.nonzero(fb[:, m])[0]
        if nz.size:
            lo[m], ln[m] = int(nz[0]), int(nz[-1]) + 1 - int(nz[0])
    offs, blocks, starts = [0], [], np.zeros(4 * G, dtype=np.int32)
    padded = np.zeros((limit, n_mels), dtype=np.float32)
    padded[:n_bins] = fb
    for g in range(G):
        ms = range(4 * g, 4 * g + 4)
        len_g = max(4, max((int(lo[m]) % 4 + int(ln[m]) + 3) // 4 * 4 for m in ms))
        assert len_g <= limit
        block = np.zeros((4, len_g), dtype=np.float32)
        for i, m in enumerate(ms):
            s0 = min(int(lo[m]) // 4 * 4, limit - len_g)
            starts[m] = s0
            if m < n_mels:
                block[i] = padded[s0:s0 + len_g, m]
        blocks.append(block.reshape(-1))
        offs.append(offs[-1] + 4 * len_g)
    return np.concatenate(blocks), np.asarray(offs, dtype=np.int32), starts


class _MelTables:
    """Window + grouped banded filterbank tables, built once per device."""
    _cache = {}

    @classmethod
    def get(cls, device):
        key = (device.type, device.index)
        t = cls._cache.get(key)
        if t is None:
            fb_w, fb_off, fb_start = pack_grouped_filterbank(slaney_mel_filterbank())      # (513, 80)
            # the exact fp32 table torchaudio's MelSpectrogram uses (window_fn=torch.hann_window, periodic)
            window = torch.hann_window(HIFIGAN_WIN_LENGTH, periodic=True, dtype=torch.float32)
            t = dict(
                window=window.to(device),
                fb_w=torch.from_numpy(fb_w).to(device),
                fb_off=torch.from_numpy(fb_off).to(device),
                fb_start=torch.from_numpy(fb_start).to(device),
            )
            cls._cache[key] = t
        return t


def extract_mel_spec(samples: torch.Tensor, frames_major: bool = False) -> torch.Tensor:
    """samples (L,) or (B, L) fp32 CUDA -> (80, F) / (B, 80, F), F = 1 + L // 256
    (frames_major=True returns (B, F, 80), the layout Megatts.forward transposes to)."""
    squeeze = samples.dim() == 1
    x = samples.unsqueeze(0) if squeeze else samples
    t = _MelTables.get(x.device)
    out = ops.mel_spectrogram(x, t["window"], t["fb_w"], t["fb_off"], t["fb_start"], HIFIGAN_MEL_CHANNELS, 1e-5,
                              frames_major=frames_major)
    return out[0] if squeeze else out


def compute_num_frames(num_samples: int, hop: int = HIFIGAN_HOP_LENGTH) -> int:
    """Frames the reference keeps per clip: ``lhotse.utils.compute_num_frames(duration, frame_shift, sr)``
    (modules/tokenizer.py:149-154; lhotse is an un-vendored, unpinned dependency of the reference - restated from its
    published definition, parity unpinned).  With duration = L / sr and frame_shift = hop / sr both of lhotse's forms
    (round-half-up of duration / frame_shift; (samples + hop // 2) // hop) reduce to this integer expression."""
    return (int(num_samples) + hop // 2) // hop


class MelSpecExtractor:
    """Bulk mel extraction with the surface of the reference's lhotse extractor (modules/tokenizer.py:128-155;
    caller prepare_ds.py:211-217): ``extract(samples, sampling_rate) -> (num_frames, 80)`` numpy, plus
    ``extract_batch`` which runs a whole list of ragged clips as ONE kernel launch.
    The lhotse base class / HDF5 writer plumbing stays the reference's."""
    name = "mel_spec"
    frame_shift = HIFIGAN_HOP_LENGTH / HIFIGAN_SR

    def __init__(self, device="cuda"):
        self.device = torch.device(device)

    def feature_dim(self, sampling_rate: int) -> int:
        return HIFIGAN_MEL_CHANNELS

    def extract(self, samples, sampling_rate: int) -> np.ndarray:
        return self.extract_batch([samples], sampling_rate)[0]

    def extract_batch(self, clips, sampling_rate: int = HIFIGAN_SR):
        assert sampling_rate == HIFIGAN_SR
        clips = [torch.as_tensor(np.asarray(c) if not isinstance(c, torch.Tensor) else c, dtype=torch.float32).squeeze()
                 for c in clips]
        if not clips:
            return []
        lens = [int(c.shape[-1]) for c in clips]
        if min(lens) <= HIFIGAN_NFFT // 2:
            raise ValueError("clips must be longer than n_fft / 2 samples (reflect padding, as torch.stft)")
        L_max = (max(lens) + 3) // 4 * 4                         # 16-byte aligned rows for the vector loads
        host = torch.zeros(len(clips), L_max, dtype=torch.float32, pin_memory=self.device.type == "cuda")
        for i, c in enumerate(clips):
            host[i, :lens[i]] = c
        wav = host.to(self.device, non_blocking=True)
        lens_d = torch.tensor(lens, dtype=torch.int32, device=self.device)
        t = _MelTables.get(wav.device)
        mel = ops.mel_spectrogram(wav, t["window"], t["fb_w"], t["fb_off"], t["fb_start"], HIFIGAN_MEL_CHANNELS, 1e-5,
                                  frames_major=True, lens=lens_d).cpu().numpy()      # (B, F_max, 80)
        return [mel[i, :compute_num_frames(lens[i])] for i in range(len(clips))]
