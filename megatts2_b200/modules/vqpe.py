"""VQProsodyEncoder with the reference's surface (modules/vqpe.py:13-62): same ctor kwargs,
attributes (.stride .mel_bins .vq.dimension), state_dict keys (``convnet.*``,
``vq.vq.layers.0._codebook.*``) and forward contract
``forward(mel (B,T,80)) -> (zq (B,T,256), commit_loss (1,1), vq_loss (), codes (1,B,ceil(T/8)))``."""
import torch
import torch.nn as nn

from .. import ops
from .convnet import ConvNetDouble
from .quantization import ResidualVectorQuantizer
from .tokenizer import HIFIGAN_MEL_CHANNELS


class VQProsodyEncoder(nn.Module):
    def __init__(self, mel_bins: int = HIFIGAN_MEL_CHANNELS, stride: int = 8, hidden_size: int = 384,
                 kernel_size: int = 5, n_layers: int = 3, n_stacks: int = 5, n_blocks: int = 2,
                 vq_bins: int = 1024, vq_dim: int = 256, activation: str = 'ReLU'):
        super().__init__()
        self.stride = stride
        self.mel_bins = mel_bins
        self.convnet = ConvNetDouble(
            in_channels=mel_bins, out_channels=vq_dim, hidden_size=hidden_size, n_layers=n_layers,
            n_stacks=n_stacks, n_blocks=n_blocks, middle_layer=nn.MaxPool1d(stride, ceil_mode=True),
            kernel_size=kernel_size, activation=activation)
        self.vq = ResidualVectorQuantizer(dimension=vq_dim, n_q=1, bins=vq_bins, decay=0.99)

    def encode_cl(self, mel: torch.Tensor):
        """mel (B,T,>=mel_bins) -> ze (B,N,D) channels-last, codes (B,N) int64."""
        ze = self.convnet.forward_cl(mel[..., :self.mel_bins])
        cb = self.vq.vq.layers[0]._codebook
        cb._require_inference()
        B, N, D = ze.shape
        codes = cb.quantize(ze.reshape(B * N, D)).view(B, N)
        return ze, codes

    def forward(self, mel: torch.Tensor):
        mel_len = mel.size(1)
        ze, codes = self.encode_cl(mel)
        embed = self.vq.vq.layers[0]._codebook.embed
        zq_n = ops.vq_gather(codes, embed)                                     # (B,N,D): one row per code
        zq = ops.vq_gather(codes, embed, t_out=mel_len, repeat=self.stride)    # x8 repeat, truncated (vqpe.py:60-61)
        # vq_loss = mse(ze, zq) (vqpe.py:59): a scalar metric the trainer logs; tiny reduction on device
        vq_loss = torch.mean((ze - zq_n) ** 2)
        commit_loss = torch.zeros(1, 1, device=mel.device)                     # eval: stacked [0.] (core_vq.py:302,346)
        return zq, commit_loss, vq_loss, codes.unsqueeze(0)
