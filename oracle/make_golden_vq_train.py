"""Generate tests/golden/vq_train.npz from the REAL reference (container only).

    python -m oracle.make_golden_vq_train

Runs the reference's VectorQuantization (modules/quantization/core_vq.py) in train mode for three steps - k-means
initialisation on the first batch, a dead-code expiry, EMA updates, straight-through output and commitment loss with a
backward pass - while recording the indices its sample_vectors draws (torch.randperm / torch.randint are wrapped, the
reference itself is untouched).  Asserts that oracle/ref_vq_train.py reproduces every buffer, output and gradient from
the same draws, then writes inputs, draws and reference results as a fixture."""
import os
import warnings

import numpy as np
import torch

from . import ref_vq_train as RV
from . import stubs

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def main():
    warnings.simplefilter("ignore")
    stubs.install()
    from modules.quantization import core_vq as ref     # the real reference
    torch.set_num_threads(4)
    K, D, B, N, ITERS, STEPS = 48, 128, 2, 36, 6, 3   # D: the search kernel takes 128 | 256 | 512
    draws = []
    orig_perm, orig_randint = torch.randperm, torch.randint

    def rec_perm(*a, **k):
        r = orig_perm(*a, **k)
        draws.append(r.clone())
        return r

    def rec_randint(*a, **k):
        r = orig_randint(*a, **k)
        draws.append(r.clone())
        return r

    torch.manual_seed(20240924)
    vq = ref.VectorQuantization(dim=D, codebook_size=K, kmeans_init=True, kmeans_iters=ITERS, decay=0.9,
                                threshold_ema_dead_code=2, commitment_weight=0.25)
    vq.train()
    g = torch.Generator().manual_seed(7)
    # clustered data so that k-means has structure, a few codes start (almost) empty and expire at the first step
    centers = torch.randn(20, D, generator=g) * 2.0
    xs, res = [], []
    cb = RV.Codebook(D, K, kmeans_iters=ITERS, decay=0.9, threshold_ema_dead_code=2)
    torch.randperm, torch.randint = rec_perm, rec_randint
    try:
        for step in range(STEPS):
            which = torch.randint(0, 20, (B, N), generator=g) if False else orig_randint(0, 20, (B, N), generator=g)
            x = (centers[which] + 0.3 * torch.randn(B, N, D, generator=g)).transpose(1, 2).contiguous()   # (B, D, N)
            x.requires_grad_(True)
            n0 = len(draws)
            q, ind, loss = vq(x)
            step_draws = draws[n0:]
            wq = torch.randn(q.shape, generator=g)
            ((q * wq).sum() + 3.0 * loss.sum()).backward()
            r = dict(x=x.detach().clone(), wq=wq, q=q.detach().clone(), ind=ind.clone(), loss=loss.detach().clone(),
                     gx=x.grad.clone(), embed=vq._codebook.embed.clone(), embed_avg=vq._codebook.embed_avg.clone(),
                     cluster_size=vq._codebook.cluster_size.clone())
            if step == 0:
                r["init_indices"] = step_draws[0][:K].clone()
                step_draws = step_draws[1:]
            r["expire_pick"] = step_draws[0][:K].clone() if step_draws else torch.zeros(K, dtype=torch.int64)
            r["expired"] = torch.tensor(int(bool(step_draws)))
            res.append(r)
            xs.append(x)
            # ---- the restatement, from the same draws
            x2 = x.detach().clone().requires_grad_(True)
            kw = {}
            if step == 0:
                kw["init_indices"] = r["init_indices"]
            kw["expire_pick"] = r["expire_pick"]
            q2, ind2, loss2, used = RV.vq_forward_train(cb, x2, 0.25, **kw)
            ((q2 * wq).sum() + 3.0 * loss2.sum()).backward()
            assert used == bool(r["expired"]), (step, used)
            assert torch.equal(ind2, ind), f"step {step}: indices differ"
            for name, a, b_ in (("q", q2, q), ("loss", loss2, loss), ("gx", x2.grad, x.grad), ("embed", cb.embed, r["embed"]),
                                ("embed_avg", cb.embed_avg, r["embed_avg"]), ("cluster_size", cb.cluster_size, r["cluster_size"])):
                err = (a.detach() - b_.detach()).abs().max().item()
                print(f"  step {step} {name:13s} max|d| = {err:.3e}")
                assert err <= 1e-6, (step, name, err)
            print(f"step {step}: expired={bool(r['expired'])} loss={loss.item():.6f} "
                  f"dead codes after={(vq._codebook.cluster_size < 2).sum().item()}")
    finally:
        torch.randperm, torch.randint = orig_perm, orig_randint
    out = {"meta": np.array([K, D, B, N, ITERS, STEPS]), "decay": np.array(0.9), "commitment_weight": np.array(0.25)}
    for i, r in enumerate(res):
        for k, v in r.items():
            out[f"s{i}_{k}"] = v.detach().cpu().numpy()
    path = os.path.join(OUT, "vq_train.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.1f} KiB)")


if __name__ == "__main__":
    main()
