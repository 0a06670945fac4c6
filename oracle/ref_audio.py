"""CPU restatement of the audio front / back ends around the synthesis path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

* ``load_resample`` / ``peak_normalize`` restate ``librosa.load(path, sr=16000)`` + ``librosa.util.normalize(y)``
  (models/megatts2.py:335-336, prepare_ds.py:113-124).  librosa and its soxr resampler are neither vendored nor pinned by
  the reference (requirements.txt lists torch / torchaudio / lightning / lhotse / h5py only): **parity unpinned** for the
  resampling filter.  The restatement therefore follows the published band-limited interpolation of
  ``torchaudio.functional.resample`` (torchaudio IS pinned) with librosa's former default quality ("kaiser_best"
  parameters); ``tests/test_oracle_golden.py`` checks it against torchaudio itself in this container.
* ``wav_float32_payload`` restates what ``torchaudio.save(path, float32 tensor, sr)`` stores (models/megatts2.py:375):
  the samples verbatim as 32-bit IEEE float PCM.
"""
import math

import numpy as np

KAISER_BEST = dict(lowpass_filter_width=64, rolloff=0.9475937167399596, beta=14.769656459379492)


def _i0(x):
    return np.i0(x)


def resample_filter(orig_freq, new_freq, lowpass_filter_width=64, rolloff=0.9475937167399596, beta=14.769656459379492):
    """Windowed-sinc polyphase filter (up, taps) in fp64 (torchaudio.functional.resample, sinc_interp_kaiser)."""
    g = math.gcd(int(orig_freq), int(new_freq))
    orig, new = int(orig_freq) // g, int(new_freq) // g
    base = min(orig, new) * rolloff
    width = math.ceil(lowpass_filter_width * orig / base)
    idx = np.arange(-width, width + orig, dtype=np.float64)[None, :] / orig
    t = np.arange(0, -new, -1, dtype=np.float64)[:, None] / new + idx
    t = np.clip(t * base, -lowpass_filter_width, lowpass_filter_width)
    window = _i0(beta * np.sqrt(1.0 - (t / lowpass_filter_width) ** 2)) / _i0(beta)
    t = t * math.pi
    with np.errstate(invalid="ignore", divide="ignore"):
        k = np.where(t == 0, 1.0, np.sin(t) / t)
    return new, orig, width, k * window * (base / orig)


def resample(x, orig_freq, new_freq):
    """x (L,) -> (ceil(L * new / orig),), fp64 arithmetic: y[i*up + p] = sum_k xpad[i*down + k] h[p, k]."""
    if int(orig_freq) == int(new_freq):
        return np.asarray(x, dtype=np.float64)
    up, down, width, h = resample_filter(orig_freq, new_freq, **KAISER_BEST)
    x = np.asarray(x, dtype=np.float64)
    n_out = -(-len(x) * up // down)
    frames = -(-n_out // up)
    xp = np.concatenate([np.zeros(width), x, np.zeros(frames * down + h.shape[1])])
    idx = np.arange(frames)[:, None] * down + np.arange(h.shape[1])[None]
    return (xp[idx] @ h.T).reshape(-1)[:n_out]


def peak_normalize(y):
    """librosa.util.normalize(y) with its defaults (norm=inf, threshold=tiny, fill=None): y / max|y|."""
    y = np.asarray(y)
    peak = np.abs(y).max() if y.size else 0.0
    return y if peak < np.finfo(np.float32).tiny else y / peak


def wav_float32_payload(samples):
    return np.asarray(samples, dtype="<f4").tobytes()
