"""Host-side helpers mirroring the reference's utils/utils.py surface that callers of the
hot path use: make_pad_mask / make_attn_mask (utils/utils.py:12-39) and instantiate_class
(utils/utils.py:86-102).  Plumbing only - masks are built with torch ops on the input's device."""
import importlib
from typing import Any, Dict, Tuple, Union

import torch


def make_pad_mask(lengths: torch.Tensor, max_len: int = 0) -> torch.Tensor:
    """True where position >= length. (B,) -> (B, max(max_len, lengths.max()))."""
    assert lengths.ndim == 1, lengths.ndim
    n = max(int(max_len), int(lengths.max()))
    pos = torch.arange(n, device=lengths.device)
    return pos[None, :] >= lengths[:, None]


def make_attn_mask(lengths: torch.Tensor, num_heads: int, causal: bool = False) -> torch.Tensor:
    """Additive fp32 attention mask: 0 = keep, -inf = masked.
    non-causal -> (B, H, 1, T); causal -> (B, H, T, T) (requires all lengths == T, as the reference asserts)."""
    pad = make_pad_mask(lengths)
    b, t = pad.shape
    pad = pad.view(b, 1, 1, t).expand(-1, num_heads, -1, -1)
    if causal:
        assert t == int(lengths.max()), "Causal mask requires all lengths to be equal to max_len"
        future = torch.ones(t, t, dtype=torch.bool, device=pad.device).triu(1).view(1, 1, t, t)
        blocked = future | pad
    else:
        blocked = pad
    return torch.zeros(blocked.shape, dtype=torch.float32, device=pad.device).masked_fill(blocked, float("-inf"))


def instantiate_class(args: Union[Any, Tuple[Any, ...]], init: Dict[str, Any]) -> Any:
    """{"class_path": "pkg.mod.Class", "init_args": {...}} -> Class(*args, **init_args).
    ``class_path`` is the plugin switch: the YAMLs under configs/ point it at megatts2_b200.*"""
    kwargs = init.get("init_args", {})
    if not isinstance(args, tuple):
        args = (args,)
    mod_name, cls_name = init["class_path"].rsplit(".", 1)
    return getattr(importlib.import_module(mod_name), cls_name)(*args, **kwargs)
