"""Stage-2 latent dump (reference: ``DatasetMaker.extract_latent``, prepare_ds.py:224-258).

For every cut the reference calls ``G.s2_latent(phone_tokens, tokens_lens, mel_timbres, mel_targets)`` at batch 1 and
stores ``{'tc_latent': (1, Tp, 512) float32, 'p_code': (1, 1, ceil(Tt / 8)) int64}`` with ``np.save`` under
``<ds_path>/latents/<speaker>/<recording_id>.npy``.  Here the cuts are bucketed by their tensor shapes - the MRTE encoders
are unmasked in the reference (modules/mrte.py:154-171), so only equal-shape cuts may share a batch without changing
results - and each bucket is ONE pass of the CUDA path; files have the reference's layout, dtypes and names."""
import os
from typing import Iterable, Tuple

import numpy as np
import torch


def dump_s2_latents(G, items: Iterable[Tuple[str, str, torch.Tensor, torch.Tensor, torch.Tensor]], ds_path: str,
                    max_batch: int = 64, device=None):
    """items: (recording_id, speaker, phone_tokens (Tp,) int64, mel_timbres (Tm, 80), mel_targets (Tt, 80)), host or device
    tensors.  Returns the list of written paths (input order)."""
    device = device or next(G.parameters()).device
    items = list(items)
    buckets = {}
    for i, (_, _, ph, mt, mg) in enumerate(items):
        buckets.setdefault((ph.shape[0], mt.shape[0], mg.shape[0]), []).append(i)
    paths = [None] * len(items)
    with torch.no_grad():
        for _, ids in sorted(buckets.items()):
            for s in range(0, len(ids), max_batch):
                chunk = ids[s:s + max_batch]
                ph = torch.stack([items[i][2] for i in chunk]).to(device)
                mt = torch.stack([items[i][3] for i in chunk]).to(device, torch.float32)
                mg = torch.stack([items[i][4] for i in chunk]).to(device, torch.float32)
                lens = torch.full((len(chunk),), ph.shape[1], dtype=torch.int32, device=device)
                tc, codes = G.s2_latent(ph, lens, mt, mg)                       # (b, Tp, 512), (1, b, T8)
                tc_h, codes_h = tc.cpu().numpy(), codes.cpu().numpy()
                for j, i in enumerate(chunk):
                    rid, spk = items[i][0], items[i][1]
                    d = os.path.join(ds_path, "latents", str(spk))
                    os.makedirs(d, exist_ok=True)
                    paths[i] = os.path.join(d, f"{rid}.npy")
                    np.save(paths[i], {"tc_latent": tc_h[j:j + 1], "p_code": codes_h[:, j:j + 1]})
    return paths
