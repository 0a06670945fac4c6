"""GPU: parity at the sequence lengths BASELINE.json names and the remaining a11 / 8f-2 surfaces.

* One body of the MegaPLM.infer / MegaADM.infer loops (models/megatts2.py:172-178, 264-273) at t in {128, 256, 512}
  - the lengths config C3 reaches, where the K/V length crosses the attention kernel's tiling and the tap-GEMM's
  tile / split-K / CTA-pair heuristics change - teacher-forced on seeded prefixes, product vs the CPU oracle.
  (The free-running decode at T = 512 is 42 TFLOP per sequence: ~hours on the CPU; one step is 0.3 TFLOP.)
* MegaG.forward / MegaG.s2_latent (models/megatts2.py:56-84), which the stage-2 latent dump (prepare_ds.py:224-258)
  calls, vs the oracle.
"""
import pytest
import torch
import torch.nn.functional as F

import helpers
from megatts2_b200 import ops, pack
from megatts2_b200.modules.transformer import run_encoder
from oracle import ref_megatts2 as R
from oracle import weights

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]
DEV = "cuda"


def gen(seed):
    g = torch.Generator()
    g.manual_seed(seed)
    return g


@pytest.fixture(scope="module")
def PLM(weights_cpu):
    return helpers.build_plm(weights_cpu("plm"), DEV)


@pytest.fixture(scope="module")
def ADM(weights_cpu):
    return helpers.build_adm(weights_cpu("adm"), DEV)


def _check_logits(name, t, got, ref, tol):
    err = (got.cpu() - ref).abs().max().item()
    top = ref.topk(2, -1).values
    clear = (top[..., 0] - top[..., 1]) > 1e-3
    same = got.cpu().argmax(-1) == ref.argmax(-1)
    helpers.record(f"{name}_single_step", dict(t=t, logits_max_abs_err=err, rows=int(same.numel()),
                                               argmax_equal=int(same.sum()), rows_with_clear_margin=int(clear.sum())))
    assert err < tol, f"{name} t={t}: logits differ by {err}"
    assert bool(same[clear].all()), f"{name} t={t}: argmax differs on a row whose top-2 margin exceeds 1e-3"
    assert int(clear.sum()) >= 1


@pytest.mark.parametrize("t", [128, 256, 512])
def test_plm_single_step_parity_at_baseline_lengths(weights_cpu, PLM, t):
    """Step t-1 of MegaPLM.infer (non-causal full recompute over t positions, last row only), B = 2."""
    B = 2
    tc = F.relu(torch.randn(B, t, 512, generator=gen(700 + t)))
    codes = torch.cat([torch.full((B, 1), 1024), torch.randint(0, 1024, (B, t - 1), generator=gen(800 + t))], 1)
    ref = R.plm_step_logits(R.SD(weights_cpu("plm")), tc, codes, weights.PLM_CFG)              # (B, 1024)
    x = torch.cat([tc.to(DEV), ops.embed_pe(codes.to(DEV), PLM.pc_embedding.weight.detach())], -1)
    x = PLM.pos(x)
    h = run_encoder(PLM.plm, list(PLM.plm.layers), x, last_row_only=True)                      # what infer() consumes
    got = ops.linear(h[:, 0], pack.pack_linear(PLM.predict_layer.weight))
    _check_logits("plm", t, got, ref, 2e-3)
    # the full-sequence forward agrees with the pruned last row (the exactness claim of the pruning)
    full = run_encoder(PLM.plm, list(PLM.plm.layers), x)
    assert (full[:, -1] - h[:, 0]).abs().max().item() < 5e-4


@pytest.mark.parametrize("t", [128, 256, 512])
def test_adm_single_step_parity_at_baseline_lengths(weights_cpu, ADM, t):
    """Step t-1 of MegaADM.infer: raw-float feedback row, non-causal recompute over t positions, B = 2."""
    B = 2
    sd = R.SD(weights_cpu("adm"))
    tcl = F.relu(torch.randn(B, t, 512, generator=gen(900 + t)))
    p = torch.cat([torch.zeros(B, 1, 1), torch.rand(B, t - 1, 1, generator=gen(950 + t)) * 6 + 1], 1)
    xr = R.sine_pe_add(torch.cat([F.linear(tcl, sd("tc_linear_emb.weight")), F.linear(p, sd("dt_linear_emb.weight"))], -1),
                       sd("pos_emb.alpha"))
    ref = F.linear(R.encoder(sd.sub("adm"), xr, weights.ADM_CFG["n_layers"], weights.ADM_CFG["n_heads"], False),
                   sd("predict_layer.weight"))[:, -1, 0]                                     # (B,)
    x = torch.cat([ops.linear(tcl.to(DEV), pack.pack_linear(ADM.tc_linear_emb.weight)),
                   ops.linear(p.to(DEV), pack.pack_linear(ADM.dt_linear_emb.weight))], -1)
    x = ADM.pos_emb(x)
    h = run_encoder(ADM.adm, list(ADM.adm.layers), x, last_row_only=True)
    got = ops.linear(h[:, 0], pack.pack_linear(ADM.predict_layer.weight))[:, 0]
    err = (got.cpu() - ref).abs().max().item()
    helpers.record("adm_single_step", dict(t=t, raw_max_abs_err=err, ref=[float(v) for v in ref]))
    assert err < 2e-3 * max(1.0, ref.abs().max().item()), err


def test_megag_forward_and_s2_latent_vs_oracle(weights_cpu):
    """a11 / 8f-2: MegaG.s2_latent -> (tc_latent, codes) and MegaG.forward -> (mel, commit_loss, vq_loss) vs the oracle's
    restatement of models/megatts2.py:56-84 (the 3-argument tc_latent call of the reference's own callers)."""
    G = helpers.build_g(weights_cpu("g"), DEV)
    g = R.SD(weights_cpu("g"))
    B, Tp = 2, 7
    phone = torch.randint(0, 320, (B, Tp), generator=gen(71))
    lens = torch.full((B,), Tp, dtype=torch.int32)
    mel_mrte = torch.randn(B, 96, 80, generator=gen(72)) * 2 - 4
    d0 = torch.randint(1, 5, (Tp,), generator=gen(73), dtype=torch.int32)
    dur = torch.stack([d0, d0[torch.randperm(Tp, generator=gen(75))]])               # equal totals: forward() concatenates
    L = int(dur[0].sum())
    mel_vqpe = torch.randn(B, L, 80, generator=gen(74)) * 2 - 4
    # oracle
    tc_ref, _, _ = R.mrte_tc_latent(g.sub("mrte"), phone, mel_mrte, weights.G_CFG)
    zq_ref, commit_ref, vql_ref, codes_ref, _ = R.vqpe_forward(g.sub("vqpe"), mel_vqpe, weights.G_CFG)
    x_ref = torch.cat([R.length_regulate(tc_ref, dur), zq_ref], -1).transpose(1, 2)
    mel_ref = R.convnet(g.sub("decoder"), x_ref, weights.G_CFG["dec_kernel"], weights.G_CFG["dec_n_stack"],
                        weights.G_CFG["dec_n_block"]).transpose(1, 2)
    # product
    tc, codes = G.s2_latent(phone.to(DEV), lens.to(DEV), mel_mrte.to(DEV), mel_vqpe.to(DEV))
    assert torch.equal(codes.cpu(), codes_ref), "stage-2 prosody codes must be bit-exact"
    assert (tc.cpu() - tc_ref).abs().max().item() < 2e-4
    mel, commit, vql = G(dur.to(DEV), phone.to(DEV), lens.to(DEV), mel_mrte.to(DEV), mel_vqpe.to(DEV))
    assert mel.shape == (B, L, 80)
    assert (mel.cpu() - mel_ref).abs().mean().item() < 1e-4 and (mel.cpu() - mel_ref).abs().max().item() < 1e-3
    assert float(commit.abs().max()) == 0.0 and commit.shape == commit_ref.shape
    assert abs(float(vql) - float(vql_ref)) < 1e-5 * max(1.0, float(vql_ref))


def test_stage2_latent_dump_layout_and_values(weights_cpu, tmp_path):
    """8f-2: megatts2_b200.latents.dump_s2_latents writes what prepare_ds.py:224-258 writes - one pickled dict per cut,
    {'tc_latent': (1, Tp, 512) float32, 'p_code': (1, 1, ceil(Tt / 8)) int64} - bucketing equal-shape cuts into one batch;
    values vs the oracle's batch-1 restatement of G.s2_latent."""
    import numpy as np
    from megatts2_b200.latents import dump_s2_latents
    G = helpers.build_g(weights_cpu("g"), DEV)
    g = R.SD(weights_cpu("g"))
    items = []
    for i, (tp, tm, tt) in enumerate([(6, 96, 40), (9, 64, 33), (6, 96, 40)]):
        items.append((f"rec{i}", f"spk{i % 2}", torch.randint(0, 320, (tp,), generator=gen(300 + i)),
                      torch.randn(tm, 80, generator=gen(310 + i)) * 2 - 4, torch.randn(tt, 80, generator=gen(320 + i)) * 2 - 4))
    paths = dump_s2_latents(G, items, str(tmp_path))
    for (rid, spk, ph, mt, mg), pth in zip(items, paths):
        assert pth == str(tmp_path / "latents" / spk / f"{rid}.npy")
        d = np.load(pth, allow_pickle=True).item()
        tc_ref, _, _ = R.mrte_tc_latent(g.sub("mrte"), ph[None], mt[None], weights.G_CFG)
        _, _, _, codes_ref, _ = R.vqpe_forward(g.sub("vqpe"), mg[None], weights.G_CFG)
        assert d["tc_latent"].shape == (1, ph.shape[0], 512) and d["tc_latent"].dtype == np.float32
        assert d["p_code"].shape == (1, 1, (mg.shape[0] + 7) // 8) and d["p_code"].dtype == np.int64
        assert np.array_equal(d["p_code"], codes_ref.numpy())
        assert np.abs(d["tc_latent"] - tc_ref.numpy()).max() < 2e-4
