"""Audio front / back ends of ``Megatts.forward`` on the device: resample to 16 kHz + peak
normalisation of the prompt clips (librosa.load / librosa.util.normalize, models/megatts2.py:335-336) and the wav writer
behind ``torchaudio.save`` (:375).  File parsing stays on the host (RIFF headers); everything that touches samples runs
in libmegatts2_b200 (csrc/audio.cu).

Resampler: librosa's default (``soxr_hq``) is an un-vendored, unpinned dependency of the reference, so the
kernel implements the published band-limited interpolation of ``torchaudio.functional.resample`` - torchaudio IS pinned by
the reference (requirements.txt) - with the Kaiser-windowed "kaiser_best" parameters librosa used before soxr; the filter
table is built here in fp64 exactly as torchaudio builds it.  Parity vs soxr: unpinned; vs torchaudio: tested."""
import ctypes as C
import math
import struct

import numpy as np
import torch

from . import _lib as L
from . import ops

KAISER_BEST = dict(lowpass_filter_width=64, rolloff=0.9475937167399596, beta=14.769656459379492)
_kernels = {}


def resample_table(orig_freq, new_freq, lowpass_filter_width=64, rolloff=0.9475937167399596, beta=14.769656459379492):
    """(up, down, width, h (up, taps) fp32) of torchaudio.functional.resample(..., resampling_method="sinc_interp_kaiser"):
    h[p, k] = windowed sinc of output phase p at input offset k - width.  Built on the host with the dtype sequence
    torchaudio itself uses - fp64 grid, but the phase offsets p / up and beta pass through fp32 - so the table is
    bit-identical to torchaudio's (the CPU suite checks that)."""
    g = math.gcd(int(orig_freq), int(new_freq))
    orig, new = int(orig_freq) // g, int(new_freq) // g
    base = min(orig, new) * rolloff
    width = math.ceil(lowpass_filter_width * orig / base)
    idx = torch.arange(-width, width + orig, dtype=torch.float64)[None, :] / orig
    t = (torch.arange(0, -new, -1) / new)[:, None] + idx              # int64 / int -> fp32 phases, promoted to fp64
    t = (t * base).clamp_(-lowpass_filter_width, lowpass_filter_width)
    b32 = torch.tensor(float(beta))                                    # fp32 scalar, as in torchaudio
    window = torch.i0(b32 * torch.sqrt(1 - (t / lowpass_filter_width) ** 2)) / torch.i0(b32)
    t = t * math.pi
    k = torch.where(t == 0, torch.tensor(1.0, dtype=t.dtype), t.sin() / t) * window * (base / orig)
    return new, orig, width, np.ascontiguousarray(k.to(torch.float32).numpy())


def _table(orig_freq, new_freq, device):
    key = (int(orig_freq), int(new_freq), str(device))
    if key not in _kernels:
        up, down, width, h = resample_table(orig_freq, new_freq, **KAISER_BEST)
        _kernels[key] = (up, down, width, torch.from_numpy(h).to(device))
    return _kernels[key]


def resample(x, orig_freq, new_freq, lens=None):
    """x (B, L) fp32 on the device -> (B, ceil(L * new / orig)); ``lens`` (B,) int32: ragged clips (each clip's output
    length is ceil(len * new / orig), the rest is zero).  Returns (y, out_lens | None)."""
    x = ops._dev(x, name="x")
    if x.stride(1) != 1:
        x = x.contiguous()
    if int(orig_freq) == int(new_freq):
        return x, lens
    up, down, width, h = _table(orig_freq, new_freq, x.device)
    B, Lin = x.shape
    Lout = -(-Lin * up // down)
    y = torch.empty(B, Lout, dtype=torch.float32, device=x.device)
    lo = None
    if lens is not None:
        lens = ops._dev(lens, torch.int32, "lens").contiguous()
        lo = (-(-(lens.to(torch.int64) * up) // down)).to(torch.int32)      # ceil(len * up / down), tiny host-free op
    L.check(L.lib().mtts_resample_f32(ops._ptr(x), x.stride(0), B, Lin, ops._ptr(lens), ops._ptr(h), up, down, width,
                                      h.shape[1], ops._ptr(y), y.stride(0), Lout, ops._ptr(lo), ops._stream()))
    return y, lo


def peak_normalize(x, lens=None):
    """in place x[b] /= max |x[b]|  (librosa.util.normalize, norm=inf); returns x."""
    x = ops._dev(x, name="x")
    assert x.dim() == 2 and x.stride(1) == 1
    scratch = torch.empty(x.shape[0], dtype=torch.int32, device=x.device)
    if lens is not None:
        lens = ops._dev(lens, torch.int32, "lens").contiguous()
    L.check(L.lib().mtts_peak_normalize_f32(ops._ptr(x), x.stride(0), x.shape[0], x.shape[1], ops._ptr(lens),
                                            ops._ptr(scratch), ops._stream()))
    return x


def load_prompts(clips, device, target_sr=16000):
    """The audio half of Megatts.forward's prompt loop (:333-338) for a list of (samples float32 numpy (L,), sr): clips of
    one sampling rate are resampled in ONE launch; returns (wav (B, Lmax) on the device, lens (B,) int32)."""
    by_sr = {}
    for i, (y, sr) in enumerate(clips):
        by_sr.setdefault(int(sr), []).append(i)
    outs = [None] * len(clips)
    for sr, ids in by_sr.items():
        Lmax = max(len(clips[i][0]) for i in ids)
        host = torch.zeros(len(ids), Lmax, dtype=torch.float32)
        for j, i in enumerate(ids):
            host[j, :len(clips[i][0])] = torch.from_numpy(np.asarray(clips[i][0], dtype=np.float32))
        lens = torch.tensor([len(clips[i][0]) for i in ids], dtype=torch.int32)
        y, lo = resample(host.to(device), sr, target_sr, lens.to(device))
        lo = lens.to(device) if lo is None else lo
        peak_normalize(y, lo)
        lo_h = lo.tolist()
        for j, i in enumerate(ids):
            outs[i] = y[j, :lo_h[j]]
    Lmax = max(o.shape[0] for o in outs)
    wav = torch.zeros(len(outs), Lmax, dtype=torch.float32, device=device)
    for i, o in enumerate(outs):
        wav[i, :o.shape[0]] = o
    return wav, torch.tensor([o.shape[0] for o in outs], dtype=torch.int32, device=device)


# ------------------------------------------------------------------------------------------ RIFF / WAVE (host framing)
def read_wav(path):
    """Minimal RIFF/WAVE reader (PCM 8/16/24/32 and IEEE float 32/64) -> (mono float32 numpy (L,), sample_rate).
    Multi-channel files are averaged to mono like librosa.load(mono=True)."""
    with open(path, "rb") as f:
        data = f.read()
    if data[:4] != b"RIFF" or data[8:12] != b"WAVE":
        raise ValueError(f"{path}: not a RIFF/WAVE file")
    pos, fmt, payload = 12, None, None
    while pos + 8 <= len(data):
        cid, size = data[pos:pos + 4], struct.unpack("<I", data[pos + 4:pos + 8])[0]
        body = data[pos + 8:pos + 8 + size]
        if cid == b"fmt ":
            tag, ch, sr, _, _, bits = struct.unpack("<HHIIHH", body[:16])
            if tag == 0xFFFE and len(body) >= 26:
                tag = struct.unpack("<H", body[24:26])[0]
            fmt = (tag, ch, sr, bits)
        elif cid == b"data":
            payload = body
        pos += 8 + size + (size & 1)
    if fmt is None or payload is None:
        raise ValueError(f"{path}: missing fmt / data chunk")
    tag, ch, sr, bits = fmt
    if tag == 3:
        x = np.frombuffer(payload, dtype="<f4" if bits == 32 else "<f8").astype(np.float32)
    elif tag == 1 and bits == 16:
        x = np.frombuffer(payload, dtype="<i2").astype(np.float32) / 32768.0
    elif tag == 1 and bits == 32:
        x = np.frombuffer(payload, dtype="<i4").astype(np.float32) / 2147483648.0
    elif tag == 1 and bits == 8:
        x = (np.frombuffer(payload, dtype=np.uint8).astype(np.float32) - 128.0) / 128.0
    elif tag == 1 and bits == 24:
        b = np.frombuffer(payload, dtype=np.uint8).reshape(-1, 3).astype(np.int32)
        v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
        x = (v - ((v & 0x800000) << 1)).astype(np.float32) / 8388608.0
    else:
        raise ValueError(f"{path}: unsupported WAVE format tag {tag} / {bits} bits")
    x = x[: (len(x) // ch) * ch].reshape(-1, ch)
    return (x.mean(axis=1) if ch > 1 else x[:, 0]).astype(np.float32), sr


def wav_bytes(samples, sample_rate, encoding="PCM_F"):
    """RIFF/WAVE bytes of a mono signal.  ``samples``: float32 (PCM_F: 32-bit IEEE float, what torchaudio.save writes for a
    float32 tensor) or int16 (PCM_S: 16-bit PCM) numpy array."""
    samples = np.ascontiguousarray(samples)
    if encoding == "PCM_F":
        assert samples.dtype == np.float32
        tag, bits = 3, 32
    else:
        assert samples.dtype == np.int16
        tag, bits = 1, 16
    payload = samples.tobytes()
    block = bits // 8
    fmt = struct.pack("<HHIIHH", tag, 1, sample_rate, sample_rate * block, block, bits)
    chunks = b"fmt " + struct.pack("<I", len(fmt)) + fmt
    if tag == 3:
        chunks += b"fact" + struct.pack("<II", 4, len(samples))
    chunks += b"data" + struct.pack("<I", len(payload)) + payload
    return b"RIFF" + struct.pack("<I", 4 + len(chunks)) + b"WAVE" + chunks


def save_wav(path, wav, sample_rate=16000, length=None, encoding="PCM_F"):
    """The back end of Megatts.forward (torchaudio.save('test.wav', audio[0], 16000), :375): wav (1, L) | (L,) fp32 on the
    device -> file.  PCM_S quantises on the device (mtts_pcm16_f32); the file framing is written by the host."""
    w = ops._dev(wav.reshape(-1), name="wav")
    if length is not None:
        w = w[:length]
    w = w.contiguous()
    if encoding == "PCM_S":
        q = torch.empty(w.numel(), dtype=torch.int16, device=w.device)
        L.check(L.lib().mtts_pcm16_f32(ops._ptr(w), w.numel(), ops._ptr(q), ops._stream()))
        host = q.cpu().numpy()
    else:
        host = w.cpu().numpy()
    with open(path, "wb") as f:
        f.write(wav_bytes(host, int(sample_rate), encoding))
