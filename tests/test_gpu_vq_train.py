"""GPU: training-mode VQ codebook (second half) against the fixture the REAL reference produced
(oracle/make_golden_vq_train.py: VectorQuantization in train mode, three steps, k-means initialisation, dead-code expiry,
EMA update, straight-through output + commitment loss, backward) with the reference's random draws replayed."""
import os

import numpy as np
import pytest
import torch

import helpers
from megatts2_b200 import vq_train as VT
from megatts2_b200.modules.quantization.core_vq import VectorQuantization
from oracle import ref_vq_train as RV

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]
DEV = "cuda"
GOLD = os.path.join(os.path.dirname(__file__), "golden", "vq_train.npz")


def _t(a):
    return torch.from_numpy(np.asarray(a))


def test_vq_train_steps_match_reference_fixture(monkeypatch):
    g = np.load(GOLD)
    K, D, B, N, iters, steps = (int(v) for v in g["meta"])
    vq = VectorQuantization(dim=D, codebook_size=K, kmeans_init=True, kmeans_iters=iters, decay=float(g["decay"]),
                            threshold_ema_dead_code=2, commitment_weight=float(g["commitment_weight"])).to(DEV)
    vq.train()
    draws = []
    monkeypatch.setattr(VT, "sample_indices", lambda n, num, device: draws.pop(0).to(device))
    for s in range(steps):
        if s == 0:
            draws.append(_t(g["s0_init_indices"]))
        draws.append(_t(g[f"s{s}_expire_pick"]))
        x = _t(g[f"s{s}_x"]).to(DEV).requires_grad_(True)
        wq = _t(g[f"s{s}_wq"]).to(DEV)
        q, ind, loss = vq(x)
        ((q * wq).sum() + 3.0 * loss.sum()).backward()
        assert not draws
        assert torch.equal(ind.cpu(), _t(g[f"s{s}_ind"])), f"step {s}: code indices differ from the reference"
        cb = vq._codebook
        for name, got, tol in (("q", q, 1e-6), ("loss", loss, 1e-6), ("gx", x.grad, 1e-6), ("embed", cb.embed, 2e-5),
                               ("embed_avg", cb.embed_avg, 2e-5), ("cluster_size", cb.cluster_size, 2e-5)):
            ref = _t(g[f"s{s}_{name}"])
            err = (got.detach().cpu() - ref).abs().max().item()
            helpers.record("vq_train", dict(step=s, tensor=name, max_abs_err=err))
            assert err <= tol * max(1.0, ref.abs().max().item()), (s, name, err)
    assert bool(cb.inited.item())


@pytest.mark.parametrize("N,K,D", [(4096, 1024, 256), (777, 100, 20), (16, 64, 32)])
def test_kmeans_pieces_vs_oracle(N, K, D):
    gen = torch.Generator().manual_seed(N + K)
    centers = torch.randn(max(K // 3, 2), D, generator=gen) * 1.5
    x = centers[torch.randint(0, centers.shape[0], (N,), generator=gen)] + 0.4 * torch.randn(N, D, generator=gen)
    init = torch.randperm(N, generator=gen)[:K] if N >= K else torch.randint(0, N, (K,), generator=gen)
    means0 = x[init]
    # assignment: equal to the direct-difference argmax wherever the two best distances are not a rounding apart
    d = ((x[:, None, :].double() - means0[None, :, :].double()) ** 2).sum(-1)
    top2 = d.topk(2, dim=1, largest=False)
    idx = VT.kmeans_assign(x.to(DEV), means0.to(DEV)).cpu()
    gap = (top2.values[:, 1] - top2.values[:, 0]) / top2.values[:, 1].clamp_min(1e-30)
    clear = gap > 1e-5
    assert torch.equal(idx[clear], top2.indices[clear, 0])
    assert ((d.gather(1, idx[:, None])[:, 0] - top2.values[:, 0]) <= 1e-5 * top2.values[:, 0] + 1e-12).all()
    helpers.record("kmeans_assign", dict(N=N, K=K, D=D, exact=float((idx == top2.indices[:, 0]).float().mean())))
    # per-code sums in sample order: bit-identical to a sequential scatter_add_
    s, c = VT.cluster_sum(x.to(DEV), idx.to(DEV), K)
    ref_s = torch.zeros(K, D).scatter_add_(0, idx[:, None].expand(-1, D), x)
    assert torch.equal(c.cpu(), torch.bincount(idx, minlength=K).float())
    assert torch.equal(s.cpu(), ref_s)
    # whole k-means from the same start
    means, bins = VT.kmeans(x.to(DEV), K, 4, init_indices=init.to(DEV))
    rmeans, rbins = RV.kmeans(x, K, 4, init)
    assert (bins.cpu() - rbins).abs().sum().item() <= max(2, N // 500)      # a near-tie may move a sample
    assert (means.cpu() - rmeans).abs().max().item() <= 1e-3 * max(1.0, rmeans.abs().max().item())


def test_expiry_is_a_noop_without_dead_codes():
    K, D, N = 32, 16, 200
    gen = torch.Generator().manual_seed(3)
    embed = torch.randn(K, D, generator=gen).to(DEV)
    before = embed.clone()
    VT.replace_expired(embed, torch.randn(N, D, generator=gen).to(DEV), torch.full((K,), 5.0, device=DEV), 2.0)
    assert torch.equal(embed, before)
    cs = torch.full((K,), 5.0, device=DEV)
    cs[3] = 0.5
    samples = torch.randn(N, D, generator=gen).to(DEV)
    pick = torch.arange(K, device=DEV) * 3
    VT.replace_expired(embed, samples, cs, 2.0, pick)
    assert torch.equal(embed[3], samples[9])
    mask = torch.ones(K, dtype=torch.bool)
    mask[3] = False
    assert torch.equal(embed[mask], before[mask])


def test_eval_mode_unchanged_and_uninitialised_codebook_raises():
    from megatts2_b200 import _lib as L
    vq = VectorQuantization(dim=16, codebook_size=32, kmeans_init=True).to(DEV).eval()
    with pytest.raises(L.MttsError):
        vq(torch.randn(1, 16, 8, device=DEV))
