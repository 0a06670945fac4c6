"""One-time weight re-layout for the tap-GEMM kernels (host-side plumbing, torch ops).

The kernels take weights as (k, Cin, Cout) row-major ("tap-major, output-channel
contiguous"); torch stores nn.Linear as (N, K), nn.Conv1d as (Cout, Cin, k) and
nn.ConvTranspose1d as (Cin, Cout, k).
"""
import ctypes as C

import torch

from . import _lib as L


def pack_linear(w):
    """nn.Linear weight (N, K) -> (1, K, N)."""
    return w.detach().t().contiguous().unsqueeze(0).float()


def pack_conv(w):
    """nn.Conv1d weight (Cout, Cin, k) -> (k, Cin, Cout)."""
    return w.detach().permute(2, 1, 0).contiguous().float()


def pack_qkv(wq, wk, wv):
    """three (D, D) projections -> (1, D, 3D) so one GEMM yields [q | k | v]."""
    return pack_linear(torch.cat([wq.detach(), wk.detach(), wv.detach()], 0))


def pack_conv_transpose(w, bias, stride):
    """nn.ConvTranspose1d weight (Cin, Cout, k = 2*stride), padding (k - stride)//2, as the
    2-tap conv the kernel runs:  super-row u: out[u*s + r - pad] = x[u-1] W[..., r+s] + x[u] W[..., r]
    -> weight (2, Cin, s*Cout) with column index r*Cout + co, bias tiled s times."""
    cin, cout, k = w.shape
    s = stride
    assert k == 2 * s, "only kernel == 2*stride transposed convs are on the path (HiFi-GAN V1)"
    w = w.detach().float()
    tap0 = w[:, :, s:].permute(0, 2, 1).reshape(cin, s * cout)   # pairs with x[u-1]
    tap1 = w[:, :, :s].permute(0, 2, 1).reshape(cin, s * cout)   # pairs with x[u]
    wp = torch.stack([tap0, tap1], 0).contiguous()
    bp = bias.detach().float().repeat(s).contiguous() if bias is not None else None
    return wp, bp


def pack_tc_planes(w):
    """(N, K) fp32 weight -> (3, N, K) bf16 planes with w = w1 + w2 + w3 to ~2^-24 (round-to-nearest at
    every step, the same split the kernels apply to activations)."""
    w = w.detach().float()
    p1 = w.to(torch.bfloat16)
    r1 = w - p1.float()
    p2 = r1.to(torch.bfloat16)
    p3 = (r1 - p2.float()).to(torch.bfloat16)
    return torch.stack([p1, p2, p3], 0).contiguous()


def pack_conv_tc_planes(w):
    """nn.Conv1d weight (Cout, Cin, k) fp32 -> (3, k, Cout, Cin) bf16 planes (per-tap K-major B operands)."""
    return pack_tc_planes(w.detach().permute(2, 0, 1).contiguous().reshape(-1, w.shape[1])).reshape(
        3, w.shape[2], w.shape[0], w.shape[1]).contiguous()


def default_engine() -> int:
    """1 = tcgen05 bf16x3 engine for the large GEMMs (default), 0 = fp32 FFMA engine everywhere
    (MEGATTS2_ENGINE=tc|ffma).  Both are fp32-grade; ids are bit-identical between them in the tests."""
    import os
    return 0 if os.environ.get("MEGATTS2_ENGINE", "tc").lower() in ("ffma", "fp32", "0") else 1


def signature(tensors):
    """Cheap change detector for a parameter set (storage address + in-place version)."""
    return tuple((t.data_ptr(), t._version) for t in tensors)


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class PlanMixin:
    """nn.Module mixin: the cached plan holds ctypes structs (raw device pointers) that must
    not be pickled / deep-copied with the module; it is rebuilt lazily instead."""

    def __getstate__(self):
        d = self.__dict__.copy()
        d["_plan"] = None
        return d


class Plan:
    """Keeps packed device tensors and the ctypes structs that point at them alive together."""

    def __init__(self):
        self.keep = []
        self.sig = None

    def hold(self, t):
        self.keep.append(t)
        return t

    def p(self, t):
        """hold a tensor and return its device pointer as an int (for ctypes struct fields)."""
        if t is None:
            return None
        t = t.detach()
        if t.dtype != torch.float32 or not t.is_contiguous():
            t = t.float().contiguous()
        self.keep.append(t)
        return t.data_ptr()


def build_encoder_struct(plan, layers, d_model, n_heads, ff_dim, conv_ff, engine=0):
    """layers: iterable of objects with .norm1 .norm2 .attn(w_q,w_k,w_v,out_proj[0]) .ff"""
    arr = (L.EncoderLayer * len(layers))()
    tc = engine == 1

    def tcc(w):
        t = pack_conv_tc_planes(w)
        plan.keep.append(t)
        return t.data_ptr()

    def tcp(w):
        t = pack_tc_planes(w)
        plan.keep.append(t)
        return t.data_ptr()
    for i, lyr in enumerate(layers):
        a = lyr.attn
        e = arr[i]
        e.ln1_g, e.ln1_b = plan.p(lyr.norm1.weight), plan.p(lyr.norm1.bias)
        e.ln2_g, e.ln2_b = plan.p(lyr.norm2.weight), plan.p(lyr.norm2.bias)
        e.w_qkv = plan.p(pack_qkv(a.w_q.weight, a.w_k.weight, a.w_v.weight))
        e.b_qkv = plan.p(torch.cat([a.w_q.bias.detach(), a.w_k.bias.detach(), a.w_v.bias.detach()]))
        e.w_o, e.b_o = plan.p(pack_linear(a.out_proj[0].weight)), plan.p(a.out_proj[0].bias)
        if conv_ff:
            e.w_ff1, e.b_ff1 = plan.p(pack_conv(lyr.ff[0].weight)), plan.p(lyr.ff[0].bias)
            e.w_ff2, e.b_ff2 = plan.p(pack_conv(lyr.ff[2].weight)), plan.p(lyr.ff[2].bias)
        else:
            e.w_ff1, e.b_ff1 = plan.p(pack_linear(lyr.ff[0].weight)), plan.p(lyr.ff[0].bias)
            e.w_ff2, e.b_ff2 = plan.p(pack_linear(lyr.ff[3].weight)), plan.p(lyr.ff[3].bias)
        if tc:
            e.w_qkv_tc = tcp(torch.cat([a.w_q.weight.detach(), a.w_k.weight.detach(), a.w_v.weight.detach()], 0))
            e.w_o_tc = tcp(a.out_proj[0].weight)
            if conv_ff:
                e.w_ff1_tc, e.w_ff2_tc = tcc(lyr.ff[0].weight), tcc(lyr.ff[2].weight)
            else:
                e.w_ff1_tc, e.w_ff2_tc = tcp(lyr.ff[0].weight), tcp(lyr.ff[3].weight)
    plan.hold(arr)
    enc = L.Encoder()
    enc.n_layers, enc.d_model, enc.n_heads, enc.ff_dim, enc.conv_ff = len(layers), d_model, n_heads, ff_dim, int(conv_ff)
    enc.engine = 1 if tc else 0
    enc.layers = C.cast(arr, C.POINTER(L.EncoderLayer))
    return enc


def fill_conv_blocks(plan, arr, offset, stack, engine=0):
    """stack: ResidualBlockStack; writes n_stacks*n_blocks mtts_conv_block entries at arr[offset:]."""
    i = offset
    for cs in stack.conv_stacks:
        for blk in cs.blocks:
            arr[i].w, arr[i].b = plan.p(pack_conv(blk.conv.weight)), plan.p(blk.conv.bias)
            arr[i].ln_g, arr[i].ln_b = plan.p(blk.norm.weight), plan.p(blk.norm.bias)
            if engine == 1:
                t = pack_conv_tc_planes(blk.conv.weight)
                plan.keep.append(t)
                arr[i].w_tc = t.data_ptr()
            i += 1
    return i
