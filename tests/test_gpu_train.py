"""GPU: training-mode forward + backward through the drop-in modules vs torch autograd over the oracle.

What MegaPLMTrainer / MegaADMTrainer.training_step do (models/trainer.py:243-268, 334-355): module.train(), forward under
bf16 autocast, cross-entropy (sum, ignore_index = 1025) / MSE (sum), loss.backward(), AdamW step.  Here the product modules
run that forward and backward on the library's kernels; the reference gradients come from torch autograd on the CPU oracle
(the reference's own torch ops, fp32) with the same weights, dropout switched off for the comparison."""
import pytest
import torch
import torch.nn.functional as F

import helpers
from oracle import ref_megatts2 as R
from oracle import weights

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]
DEV = "cuda"


def gen(seed):
    g = torch.Generator()
    g.manual_seed(seed)
    return g


def _no_dropout(m):
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
        if hasattr(mod, "dropout") and isinstance(getattr(mod, "dropout"), float):
            mod.dropout = 0.0
        if hasattr(mod, "p_drop"):
            mod.p_drop = 0.0


def _oracle_grads(sd, loss_fn):
    leaves = {k: v.clone().requires_grad_(v.dtype.is_floating_point) for k, v in sd.items()}
    loss = loss_fn(leaves)
    loss.backward()
    return float(loss), {k: v.grad for k, v in leaves.items() if v.grad is not None}


def _compare(model, ref_grads, loss, ref_loss, tol):
    """Per parameter: relative L2 error < tol and max-abs error < 10 * tol of the tensor's largest entry.  (The L2 norm is
    the primary bar: an activation that sits within rounding of the ReLU kink takes either sub-gradient - on the CPU and
    on the GPU alike - which moves single rows of a weight gradient by far more than arithmetic noise, most visibly when a
    batch has few rows.)"""
    assert abs(loss - ref_loss) <= 2e-4 * max(1.0, abs(ref_loss)), (loss, ref_loss)
    worst = 0.0
    n = 0
    # gradients that are zero in exact arithmetic (the key bias of an attention layer: softmax is invariant to a constant
    # shift of every score of a row) are pure rounding noise on both sides: errors are taken relative to at least 1e-5 of
    # the largest gradient in the model
    floor = 1e-5 * max(float(r.abs().max()) for r in ref_grads.values())
    for name, p in model.named_parameters():
        if not p.requires_grad:
            assert p.grad is None, name
            continue
        if name not in ref_grads:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, name
            continue
        assert p.grad is not None, f"no gradient reached {name}"
        g, r = p.grad.detach().cpu().double(), ref_grads[name].double()
        l2 = ((g - r).norm() / max(float(r.norm()), floor * r.numel() ** 0.5)).item()
        mx = (g - r).abs().max().item() / max(r.abs().max().item(), floor)
        worst = max(worst, l2)
        assert l2 < tol, f"{name}: relative L2 gradient error {l2:.3e}"
        assert mx < 10 * tol, f"{name}: relative max gradient error {mx:.3e}"
        n += 1
    assert n >= 10
    return worst


@pytest.mark.parametrize("B,T", [(4, 40), (2, 12)], ids=["tensor-core-rows", "small-rows"])
def test_plm_training_step_gradients_match_oracle_autograd(weights_cpu, B, T):
    sd = weights_cpu("plm")
    plm = helpers.build_plm(sd, DEV)
    plm.train()
    _no_dropout(plm)
    tc = F.relu(torch.randn(B, T, 512, generator=gen(5 + T)))
    codes = torch.cat([torch.full((B, 1), 1024), torch.randint(0, 1024, (B, T), generator=gen(6 + T))], 1)
    lens = torch.tensor([T] + [max(T - 3 * i, 2) for i in range(1, B)], dtype=torch.int32)
    for b in range(B):                                   # padding targets carry the ignore index like the collator's
        codes[b, 1 + int(lens[b]):] = 1025

    def ref_loss(w):
        logits, y = R.plm_forward(R.SD(w), tc, codes, lens, weights.PLM_CFG)
        return F.cross_entropy(logits.transpose(1, 2), y, reduction="sum", ignore_index=1025)
    ref_l, ref_g = _oracle_grads(sd, ref_loss)
    with torch.autocast("cuda", dtype=torch.bfloat16):    # as the trainer does; the Functions run fp32 inside
        logits, y = plm(tc.to(DEV), codes.to(DEV), lens.to(DEV))
        loss = F.cross_entropy(logits.float().transpose(1, 2), y, reduction="sum", ignore_index=1025)
    loss.backward()
    worst = _compare(plm, ref_g, float(loss), ref_l, 5e-3)
    helpers.record("plm_training_grads", dict(B=B, T=T, loss=float(loss), ref_loss=ref_l, worst_rel_grad_err=worst))
    # one optimiser step, then inference through the packed plans picks the new weights up (version counters)
    opt = torch.optim.AdamW(plm.parameters(), lr=1e-3)
    before = plm.predict_layer.weight.detach().clone()
    opt.step()
    assert not torch.equal(before, plm.predict_layer.weight.detach())
    plm.eval()
    ids = plm.infer(tc[:1, :4].to(DEV))
    assert ids.shape == (1, 4)


def test_adm_training_step_gradients_match_oracle_autograd(weights_cpu):
    sd = weights_cpu("adm")
    adm = helpers.build_adm(sd, DEV)
    adm.train()
    _no_dropout(adm)
    adm.pos_emb.alpha.requires_grad_(True)               # exercise the d-alpha path of the positional embedding too
    B, T = 4, 36
    tcl = F.relu(torch.randn(B, T, 512, generator=gen(31)))
    dt = torch.cat([torch.zeros(B, 1, 1), torch.randint(1, 9, (B, T, 1), generator=gen(32)).float()], 1)
    lens = torch.tensor([T, T - 5, T - 11, 7], dtype=torch.int32)

    def ref_loss(w):
        pred, tgt = R.adm_forward(R.SD(w), tcl, dt, lens, weights.ADM_CFG)
        return F.mse_loss(pred, tgt, reduction="sum")
    ref_l, ref_g = _oracle_grads(sd, ref_loss)
    pred, tgt = adm(tcl.to(DEV), dt.to(DEV), lens.to(DEV))
    loss = F.mse_loss(pred, tgt, reduction="sum")
    loss.backward()
    worst = _compare(adm, ref_g, float(loss), ref_l, 5e-3)
    helpers.record("adm_training_grads", dict(B=B, T=T, loss=float(loss), ref_loss=ref_l, worst_rel_grad_err=worst))


def test_training_mode_dropout_runs_and_is_seeded(weights_cpu):
    """With the configured dropout (0.1) the step runs, gradients are finite, two different seeds give different losses and
    the same seed reproduces the loss bit for bit."""
    plm = helpers.build_plm(weights_cpu("plm"), DEV)
    plm.train()
    B, T = 4, 40
    tc = F.relu(torch.randn(B, T, 512, generator=gen(41))).to(DEV)
    codes = torch.cat([torch.full((B, 1), 1024), torch.randint(0, 1024, (B, T), generator=gen(42))], 1).to(DEV)
    lens = torch.full((B,), T, dtype=torch.int32, device=DEV)

    def run(seed):
        torch.manual_seed(seed)
        plm.zero_grad(set_to_none=True)
        logits, y = plm(tc, codes, lens)
        loss = F.cross_entropy(logits.transpose(1, 2), y, reduction="sum", ignore_index=1025)
        loss.backward()
        assert all(torch.isfinite(p.grad).all() for p in plm.parameters() if p.grad is not None)
        return float(loss), logits.detach().clone()
    (la, a), (lb, b), (lc, c) = run(1), run(2), run(1)
    assert not torch.equal(a, b) and la != lb
    assert torch.equal(a, c), "same seed, same masks: the library's forward is bit-reproducible"
    assert abs(la - lc) <= 1e-6 * abs(la)          # torch's own sum reduction of the loss may differ in the last ulp


def test_attention_and_layernorm_functions_vs_torch_autograd():
    """Unit level: AttentionFn (mask + no dropout) and LayerNormFn / LinearFn gradients vs torch's own autograd in fp64."""
    from megatts2_b200 import autograd as A
    g = gen(77)
    B, T, H, dh = 3, 37, 4, 64
    D = H * dh
    q, k, v = (torch.randn(B, T, D, generator=g, dtype=torch.float64, requires_grad=True) for _ in range(3))
    mask = R.attn_mask(torch.tensor([T, T - 4, 9], dtype=torch.int32), H, True).double()
    go = torch.randn(B, T, D, generator=g, dtype=torch.float64)
    qh, kh, vh = (t.view(B, T, H, dh).transpose(1, 2) for t in (q, k, v))
    ref = (torch.softmax(qh @ kh.transpose(-1, -2) / dh ** 0.5 + mask, -1) @ vh).transpose(1, 2).reshape(B, T, D)
    ref.backward(go)
    qd, kd, vd = (t.detach().float().to(DEV).requires_grad_(True) for t in (q, k, v))
    out = A.AttentionFn.apply(qd, kd, vd, H, mask.float().to(DEV), 0.0)
    out.backward(go.float().to(DEV))
    assert (out.detach().cpu().double() - ref.detach()).abs().max().item() < 2e-5
    for a, b in ((qd, q), (kd, k), (vd, v)):
        assert (a.grad.cpu().double() - b.grad).abs().max().item() < 5e-5 * max(1.0, b.grad.abs().max().item())
    # LayerNorm + Linear(ReLU)
    x = torch.randn(300, 768, generator=g, dtype=torch.float64, requires_grad=True)
    ln = torch.nn.LayerNorm(768).double()
    lin = torch.nn.Linear(768, 1024).double()
    torch.nn.init.normal_(ln.weight, 1.0, 0.2); torch.nn.init.normal_(ln.bias, 0.0, 0.2)
    gy = torch.randn(300, 1024, generator=g, dtype=torch.float64)
    torch.relu(lin(ln(x))).backward(gy)
    xd = x.detach().float().to(DEV).requires_grad_(True)
    lnd = torch.nn.LayerNorm(768).to(DEV)
    lind = torch.nn.Linear(768, 1024).to(DEV)
    lnd.load_state_dict({k: v.float() for k, v in ln.state_dict().items()})
    lind.load_state_dict({k: v.float() for k, v in lin.state_dict().items()})
    y = A.linear(A.layernorm(xd, lnd), lind.weight, lind.bias, relu=True)
    y.backward(gy.float().to(DEV))
    for got, want in ((xd.grad, x.grad), (lnd.weight.grad, ln.weight.grad), (lnd.bias.grad, ln.bias.grad),
                      (lind.weight.grad, lin.weight.grad), (lind.bias.grad, lin.bias.grad)):
        assert (got.cpu().double() - want).abs().max().item() < 1e-4 * max(1.0, want.abs().max().item())
