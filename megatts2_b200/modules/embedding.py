"""TokenEmbedding / SinePositionalEmbedding with the reference's surface
(modules/embedding.py:21-47, 50-98): same ctor kwargs, state_dict keys
(``word_embeddings.weight``; ``alpha``) and forward signatures; device work in
libmegatts2_b200 (mtts_embed_pe_f32 / mtts_add_pe_f32)."""
import math

import torch
import torch.nn as nn

from .. import autograd as A
from .. import ops


def sine_table(n_pos: int, dim: int, reverse: bool = False) -> torch.Tensor:
    """The (n_pos, dim) fp32 table of embedding.py:66-92, built on the host exactly as the
    reference does (interleaved sin/cos of position * exp(-2i ln(1e4)/dim))."""
    cpu = torch.device("cpu")       # always a real host table, even under a meta-device constructor context
    pos = (torch.arange(n_pos - 1, -1, -1.0, dtype=torch.float32, device=cpu) if reverse
           else torch.arange(0, n_pos, dtype=torch.float32, device=cpu))
    freq = torch.exp(torch.arange(0, dim, 2, dtype=torch.float32, device=cpu) * -(math.log(10000.0) / dim))
    ang = pos.unsqueeze(1) * freq
    pe = torch.zeros(n_pos, dim, device=cpu)
    pe[:, 0::2] = torch.sin(ang)
    pe[:, 1::2] = torch.cos(ang)
    return pe


class TokenEmbedding(nn.Module):
    def __init__(self, dim_model: int, vocab_size: int, dropout: float = 0.0):
        super().__init__()
        self.vocab_size = vocab_size
        self.dim_model = dim_model
        self.dropout = nn.Dropout(p=dropout)
        self.word_embeddings = nn.Embedding(vocab_size, dim_model)

    @property
    def weight(self) -> torch.Tensor:
        return self.word_embeddings.weight

    def embedding(self, index: int) -> torch.Tensor:
        return self.word_embeddings.weight[index:index + 1]

    def forward(self, x: torch.Tensor):
        if self.training:
            return A.dropout(A.EmbeddingFn.apply(x, self.word_embeddings.weight), self.dropout.p, True)
        return ops.embed_pe(x, self.word_embeddings.weight.detach())


class SinePositionalEmbedding(nn.Module):
    def __init__(self, dim_model: int, dropout: float = 0.0, scale: bool = False, alpha: bool = False):
        super().__init__()
        self.dim_model = dim_model
        self.x_scale = math.sqrt(dim_model) if scale else 1.0
        self.alpha = nn.Parameter(torch.ones(1), requires_grad=alpha)
        self.dropout = nn.Dropout(p=dropout)
        self.reverse = False
        self.pe = None
        self._pe_host = None
        self.extend_pe(torch.zeros(1, device="cpu").expand(1, 4000))

    def extend_pe(self, x, offset=0):
        need = x.size(1) + offset
        if self._pe_host is None or self._pe_host.size(0) < need:
            self._pe_host = sine_table(need, self.dim_model, self.reverse)
            self.pe = None
        dev = x.device
        if self.pe is None or self.pe.device != dev:
            self.pe = self._pe_host.to(dev).contiguous()

    def alpha_host(self) -> float:
        """alpha as a python float, read back from the device only when the parameter changed."""
        sig = (self.alpha.data_ptr(), self.alpha._version)
        if getattr(self, "_alpha_sig", None) != sig:
            self._alpha_val = float(self.alpha.detach().cpu())
            self._alpha_sig = sig
        return self._alpha_val

    def table(self, device, need):
        self.extend_pe(torch.empty(1, need, device=device))
        return self.pe

    def forward(self, x: torch.Tensor, offset: int = 0) -> torch.Tensor:
        assert self.x_scale == 1.0, "scale=True is not used on the synthesis path"
        self.extend_pe(x, offset)
        out = x.unsqueeze(-1) if x.ndim == 2 else x
        if self.training:
            y = A.SinePosFn.apply(out, self.alpha, self.pe[offset:offset + out.size(1)].contiguous())
            return A.dropout(y, self.dropout.p, True)
        return ops.add_pe(out, self.pe[offset:offset + out.size(1)].contiguous(), self.alpha_host())
