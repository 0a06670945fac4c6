"""GPU: the tcgen05 bf16x3 engine against fp64 and against the exact FFMA engine / golden ids."""
import math

import pytest
import torch

import helpers
from megatts2_b200 import pack

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]
DEV = "cuda"


def gen(seed):
    g = torch.Generator()
    g.manual_seed(seed)
    return g


CASES = [
    dict(M=128, K=64, N=128),
    dict(M=256, K=1024, N=1024, bias=True),
    dict(M=4096, K=1024, N=3072, bias=True),
    dict(M=1000, K=1024, N=4096, bias=True, relu=True),            # M tail
    dict(M=2048, K=4096, N=1024, bias=True, res=True),
    dict(M=300, K=768, N=2304, bias=True),                         # ADM qkv
    dict(M=640, K=72, N=200),                                      # K and N tails
    dict(M=129, K=1024, N=1024, res=True),
]


@pytest.mark.parametrize("c", CASES, ids=lambda c: "-".join(f"{k}{v}" for k, v in c.items()))
def test_linear_tc_vs_fp64(c):
    from megatts2_b200 import ops
    g = gen(c["M"] + c["K"] + c["N"])
    x = torch.randn(c["M"], c["K"], generator=g) * 2
    w = torch.randn(c["N"], c["K"], generator=g) / math.sqrt(c["K"])
    b = torch.randn(c["N"], generator=g) if c.get("bias") else None
    r = torch.randn(c["M"], c["N"], generator=g) if c.get("res") else None
    ref = x.double() @ w.double().t()
    if b is not None:
        ref = ref + b.double()
    if c.get("relu"):
        ref = torch.relu(ref)
    if r is not None:
        ref = ref + r.double()
    y = ops.linear_tc(x.to(DEV), pack.pack_tc_planes(w).to(DEV), b.to(DEV) if b is not None else None,
                      res=r.to(DEV) if r is not None else None, post_act=1 if c.get("relu") else 0)
    torch.cuda.synchronize()
    err = (y.cpu().double() - ref).abs().max().item()
    # fp32-grade: the split drops O(2^-24) terms; accumulation is fp32 in TMEM
    assert err < 3e-5 * max(1.0, ref.abs().max().item()), err
    # and it must be much closer to fp64 than a single-pass bf16 product could be (~4e-3 relative)
    y_ffma = ops.linear(x.to(DEV), pack.pack_linear(w).to(DEV), b.to(DEV) if b is not None else None,
                        res=r.to(DEV) if r is not None else None, post_act=1 if c.get("relu") else 0)
    err_ffma = (y_ffma.cpu().double() - ref).abs().max().item()
    print(f"tc err {err:.3e}  ffma err {err_ffma:.3e}")
    assert err < 20 * max(err_ffma, 1e-7)


def test_split_planes_are_exact():
    w = torch.randn(257, 96, generator=gen(1)) * 3
    p = pack.pack_tc_planes(w).float()
    assert (p.sum(0) - w).abs().max().item() <= 2.0 ** -22 * w.abs().max().item()


def test_plm_tc_engine_matches_golden_ids(golden, weights_cpu):
    g = golden("plm")
    plm = helpers.build_plm(weights_cpu("plm"), DEV)
    plm.plm.engine = 1
    big = torch.cat([g["tc8"]] * 8, 0).to(DEV)                    # B = 16 so that M = B*(t+1) crosses 128
    ids, logits = plm.infer(big, return_logits=True)
    assert torch.equal(ids.cpu(), torch.cat([g["ids"]] * 8, 0)), "PLM ids must stay bit-exact on the tensor-core engine"
    assert (logits.cpu() - torch.cat([g["logits"]] * 8, 0)).abs().max().item() < 2e-3
    plm.plm.engine = 0
    ids0 = plm.infer(big)
    assert torch.equal(ids0, ids)


def test_encoder_tc_vs_ffma(weights_cpu):
    from megatts2_b200.modules.transformer import run_encoder
    plm = helpers.build_plm(weights_cpu("plm"), DEV)
    x = torch.randn(8, 40, 1024, generator=gen(3)).to(DEV)
    plm.plm.engine = 0
    y0 = plm.plm(x)
    plm.plm.engine = 1
    y1 = plm.plm(x)
    assert (y0 - y1).abs().max().item() < 5e-4
    l1 = run_encoder(plm.plm, list(plm.plm.layers), x, last_row_only=True)
    assert (l1[:, 0] - y0[:, -1]).abs().max().item() < 5e-4
