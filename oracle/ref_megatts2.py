"""CPU fp32 restatement of the Mega-TTS 2 synthesis hot path (the ORACLE).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) - never imported by the product.

Style: plain functions over a flat ``state_dict`` (key -> fp32 tensor), torch CPU
ops only.  Each function cites the reference file:line it restates (paths are
relative to the reference repo, LSimon95/megatts2 @ 2ab81a1).  Pinning: every
function here is compared against the real reference, run in the build
container, by ``oracle/make_golden.py``; the resulting fixtures live in
``tests/golden/``.  Exceptions ("parity unpinned"): the speechbrain
mel wrapper (pinned one level down, against torchaudio) and the HiFi-GAN generator
(restated from the published architecture).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


class SD:
    """Prefix view over a flat state_dict."""

    def __init__(self, sd, prefix=""):
        self.sd = sd
        self.p = prefix

    def __call__(self, name):
        return self.sd[self.p + name]

    def sub(self, name):
        return SD(self.sd, self.p + name + ".")

    def has(self, name):
        return (self.p + name) in self.sd


# ----------------------------------------------------------------------------- a1
def hann_periodic(n):
    """torch.hann_window(n) default (periodic): w[i] = 0.5 - 0.5 cos(2 pi i / n)."""
    i = np.arange(n, dtype=np.float64)
    return (0.5 - 0.5 * np.cos(2.0 * np.pi * i / n)).astype(np.float32)


def slaney_fbanks(n_freqs=513, f_min=0.0, f_max=8000.0, n_mels=80, sample_rate=16000):
    """Slaney-scale, slaney-normalised triangular mel filterbank, as requested by
    modules/tokenizer.py:108-123 (norm="slaney", mel_scale="slaney") -> (n_freqs, n_mels).

    Definition (Slaney's Auditory Toolbox): mel(f) = 3f/200 below 1 kHz, and
    15 + 27*ln(f/1000)/ln(6.4) above; n_mels+2 knots equally spaced in mel; filter m is
    the triangle over knots (m, m+1, m+2), scaled by 2/(f_{m+2} - f_m).  Computed in
    float64 and rounded once to fp32; tests pin it against torchaudio's fp32 table to
    <= 2e-7 absolute."""
    lin_slope, knee_hz, knee_mel, log_step = 3.0 / 200.0, 1000.0, 15.0, math.log(6.4) / 27.0

    def to_mel(f):
        return f * lin_slope if f < knee_hz else knee_mel + math.log(f / knee_hz) / log_step

    knots_mel = np.linspace(to_mel(f_min), to_mel(f_max), n_mels + 2)
    knots = np.where(knots_mel < knee_mel, knots_mel / lin_slope,
                     knee_hz * np.exp(log_step * (knots_mel - knee_mel)))
    bins = np.linspace(0.0, sample_rate // 2, n_freqs)[:, None]          # (n_freqs,1)
    lo, mid, hi = knots[None, :-2], knots[None, 1:-1], knots[None, 2:]
    tri = np.minimum((bins - lo) / (mid - lo), (hi - bins) / (hi - mid))
    fb = np.maximum(tri, 0.0) * (2.0 / (hi - lo))
    return torch.from_numpy(fb.astype(np.float32))


def mel_spectrogram(wav, n_fft=1024, hop=256, n_mels=80, f_min=0.0, f_max=8000.0, sr=16000,
                    clamp=1e-5):
    """extract_mel_spec (modules/tokenizer.py:107-125) -> speechbrain mel_spectogram
    -> torchaudio MelSpectrogram(power=1, center=True, reflect pad, periodic Hann,
    slaney fb) -> log(clamp(., 1e-5)).   wav (..., L) fp32 -> (..., n_mels, 1 + L//hop).
    SURVEY.md Appendix B."""
    lead = wav.shape[:-1]
    x = wav.reshape(-1, wav.shape[-1]).to(torch.float32)
    win = torch.hann_window(n_fft, periodic=True, dtype=torch.float32, device=x.device)   # torchaudio's window_fn default
    spec = torch.stft(x, n_fft=n_fft, hop_length=hop, win_length=n_fft, window=win, center=True,
                      pad_mode="reflect", normalized=False, onesided=True, return_complex=True)
    mag = spec.abs()                                        # (N, 513, F)   power = 1
    fb = slaney_fbanks(n_fft // 2 + 1, f_min, f_max, n_mels, sr).to(x.device)
    mel = torch.matmul(mag.transpose(-1, -2), fb).transpose(-1, -2)
    out = torch.log(torch.clamp(mel, min=clamp))
    return out.reshape(*lead, n_mels, out.shape[-1])


# ------------------------------------------------------------------------ a3 / a6 / a7
def layer_norm(x, w, b, eps=1e-5):
    return F.layer_norm(x, (x.shape[-1],), w, b, eps)


def conv_block(sd, x, k):
    """ConvBlock.forward (modules/convnet.py:22-31): ReLU -> (dropout, eval no-op) ->
    Conv1d(C,C,k,same) -> LayerNorm over C.   x (B,C,T)."""
    y = F.conv1d(F.relu(x), sd("conv.weight"), sd("conv.bias"), padding=(k - 1) // 2)
    return layer_norm(y.transpose(1, 2), sd("norm.weight"), sd("norm.bias")).transpose(1, 2)


def residual_stack(sd, x, k, n_stacks, n_blocks):
    """ResidualBlockStack.forward (modules/convnet.py:69-72): x = x + ConvStack(x)."""
    for s in range(n_stacks):
        y = x
        for b in range(n_blocks):
            y = conv_block(sd.sub(f"conv_stacks.{s}.blocks.{b}"), y, k)
        x = x + y
    return x


def convnet(sd, x, k, n_stacks, n_blocks):
    """ConvNet.forward (modules/convnet.py:115-119)."""
    p = (k - 1) // 2
    x = F.conv1d(x, sd("first_layer.weight"), sd("first_layer.bias"), padding=p)
    x = residual_stack(sd.sub("conv_stack"), x, k, n_stacks, n_blocks)
    return F.conv1d(x, sd("last_layer.weight"), sd("last_layer.bias"), padding=p)


def convnet_double(sd, x, k, n_layers, n_stacks, n_blocks, middle):
    """ConvNetDouble.forward (modules/convnet.py:202-210): every layer consumes the SAME
    first_layer output; layer outputs are summed; then last_layer."""
    p = (k - 1) // 2
    h = F.conv1d(x, sd("first_layer.weight"), sd("first_layer.bias"), padding=p)
    acc = None
    for l in range(n_layers):
        ls = sd.sub(f"layers.{l}")
        y = residual_stack(ls.sub("conv_stack1"), h, k, n_stacks, n_blocks)
        y = middle(ls, y)
        y = residual_stack(ls.sub("conv_stack2"), y, k, n_stacks, n_blocks)
        acc = y if acc is None else acc + y
    return F.conv1d(acc, sd("last_layer.weight"), sd("last_layer.bias"), padding=p)


def sine_pe_table(n_pos, dim):
    """SinePositionalEmbedding.extend_pe (modules/embedding.py:66-92), fp32 on host."""
    pe = torch.zeros(n_pos, dim)
    position = torch.arange(0, n_pos, dtype=torch.float32).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, dim, 2, dtype=torch.float32) * -(math.log(10000.0) / dim))
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe


def sine_pe_add(x, alpha):
    """SinePositionalEmbedding.forward (modules/embedding.py:94-98), x_scale = 1."""
    T, D = x.shape[1], x.shape[2]
    return x * 1.0 + alpha * sine_pe_table(max(T, 1), D)[None, :T]


def mha(sd, q_in, n_heads, kv_in=None, mask=None):
    """MultiHeadAttention.forward (modules/transformer.py:35-57)."""
    kv_in = q_in if kv_in is None else kv_in
    B, Tq, D = q_in.shape
    Tk = kv_in.shape[1]
    dh = D // n_heads
    q = F.linear(q_in, sd("w_q.weight"), sd("w_q.bias")).view(B, Tq, n_heads, dh).transpose(1, 2)
    k = F.linear(kv_in, sd("w_k.weight"), sd("w_k.bias")).view(B, Tk, n_heads, dh).transpose(1, 2)
    v = F.linear(kv_in, sd("w_v.weight"), sd("w_v.bias")).view(B, Tk, n_heads, dh).transpose(1, 2)
    s = torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(dh)
    if mask is not None:
        s = s + mask
    a = torch.matmul(torch.softmax(s, dim=-1), v)
    a = a.transpose(1, 2).reshape(B, Tq, D)
    return F.linear(a, sd("out_proj.0.weight"), sd("out_proj.0.bias"))


def encoder_layer(sd, x, n_heads, conv_ff, mask=None):
    """TransformerEncoderLayer.forward (modules/transformer.py:88-102).  NB the conv-FF
    branch replaces x by norm2(x) BEFORE the residual (:96-98)."""
    h = layer_norm(x, sd("norm1.weight"), sd("norm1.bias"))
    x = x + mha(sd.sub("attn"), h, n_heads, mask=mask)
    if conv_ff:
        x = layer_norm(x, sd("norm2.weight"), sd("norm2.bias"))
        y = x.transpose(1, 2)
        y = F.conv1d(y, sd("ff.0.weight"), sd("ff.0.bias"), padding=2)
        y = F.conv1d(F.relu(y), sd("ff.2.weight"), sd("ff.2.bias"), padding=2)
        return x + y.transpose(1, 2)
    h = layer_norm(x, sd("norm2.weight"), sd("norm2.bias"))
    h = F.relu(F.linear(h, sd("ff.0.weight"), sd("ff.0.bias")))
    return x + F.linear(h, sd("ff.3.weight"), sd("ff.3.bias"))


def attn_mask(lens, n_heads, causal):
    """make_attn_mask (utils/utils.py:21-39): additive float mask (B,H,T|1,T)."""
    T = int(lens.max())
    pad = torch.arange(T)[None, :] >= lens[:, None].to(torch.int64)            # (B,T)
    pad = pad[:, None, None, :].expand(-1, n_heads, -1, -1)
    if causal:
        cm = torch.triu(torch.ones(T, T, dtype=torch.bool), diagonal=1)[None, None]
        m = cm | pad
    else:
        m = pad
    return torch.zeros(m.shape).masked_fill(m, float("-inf"))


def encoder(sd, x, n_layers, n_heads, conv_ff, lens=None, causal=False):
    """TransformerEncoder.forward (modules/transformer.py:119-133); norm=None everywhere
    on the hot path."""
    mask = attn_mask(lens, n_heads, causal) if lens is not None else None
    for i in range(n_layers):
        x = encoder_layer(sd.sub(f"layers.{i}"), x, n_heads, conv_ff, mask)
    return x


# ----------------------------------------------------------------------------- a4
def vq_quantize(x, embed):
    """EuclideanCodebook.quantize (modules/quantization/core_vq.py:175-183):
    argmax_k -(|x|^2 - 2 x.e_k + |e_k|^2); ties -> first index (torch.max on CPU).
    x (N,D), embed (K,D) -> (N,) int64."""
    e = embed.t()
    dist = -(x.pow(2).sum(1, keepdim=True) - 2 * x @ e + e.pow(2).sum(0, keepdim=True))
    return dist.max(dim=-1).indices


def vq_decode(codes, embed):
    """ResidualVectorQuantizer.decode (vq.py:109-113) -> ResidualVectorQuantization.decode
    (core_vq.py:360-367) -> VectorQuantization.decode (:288-292), n_q = 1:
    codes (1,B,N) int64 -> (B,D,N)."""
    out = torch.tensor(0.0)
    for q in codes:
        out = out + F.embedding(q, embed).transpose(1, 2)
    return out


# ----------------------------------------------------------------------------- a2
def vqpe_forward(sd, mel, cfg):
    """VQProsodyEncoder.forward (modules/vqpe.py:50-62) in eval mode.
    mel (B,T,80) -> zq (B,T,256), commit_loss (1,1), vq_loss (), codes (1,B,ceil(T/8))."""
    T = mel.shape[1]
    S = cfg["vq_stride"]
    x = mel[..., :cfg["vq_mel_bins"]].transpose(1, 2)
    ze = convnet_double(
        sd.sub("convnet"), x, cfg["vq_kernel"], cfg["vq_n_layer"], cfg["vq_n_stack"], cfg["vq_n_block"],
        middle=lambda ls, y: F.max_pool1d(y, S, ceil_mode=True))
    embed = sd("vq.vq.layers.0._codebook.embed")
    B, D, N = ze.shape
    flat = ze.transpose(1, 2).reshape(-1, D)
    idx = vq_quantize(flat, embed).view(B, N)
    zq = F.embedding(idx, embed).transpose(1, 2)                     # (B,D,N)
    commit_loss = torch.zeros(1, 1)                                  # eval: loss tensor [0.] stacked (core_vq.py:302-346)
    vq_loss = F.mse_loss(ze, zq)
    up = zq.transpose(1, 2).unsqueeze(2).expand(-1, -1, S, -1).reshape(B, N * S, D)[:, :T]
    return up, commit_loss, vq_loss, idx.unsqueeze(0), ze


# ----------------------------------------------------------------------------- a5
def mrte_tc_latent(sd, phone, mel, cfg):
    """MRTE.tc_latent (modules/mrte.py:154-171).  phone (B,Tp) int64, mel (B,Tm,80)."""
    emb = F.embedding(phone, sd("phone_embedding.word_embeddings.weight"))
    x = sine_pe_add(emb, sd("phone_pos_embedding.alpha"))
    stride = cfg["mel_stride"]

    def middle(ls, y):
        return F.conv1d(y, sd("mel_encoder_middle_layer.weight"), sd("mel_encoder_middle_layer.bias"),
                        stride=stride, padding=stride // 2)
    ctx = convnet_double(sd.sub("mel_encoder"), mel.transpose(1, 2), cfg["mel_kernel"], cfg["mel_n_layer"],
                         cfg["mel_n_stack"], cfg["mel_n_block"], middle).transpose(1, 2)
    px = encoder(sd.sub("phone_encoder"), x, cfg["content_layers"], cfg["content_heads"], True)
    y = mha(sd.sub("mha"), px, 1, kv_in=ctx)
    y = layer_norm(y, sd("norm.weight"), sd("norm.bias"))
    return F.relu(y), ctx, px


# ----------------------------------------------------------------------------- a8
def length_regulate(x, durations):
    """LengthRegulator.forward + create_alignment (modules/mrte.py:23-31, 42-60):
    out[b, sum_{j<i} d_j + k] = x[b,i] for k < d_i; rows past sum(d_b) are zero;
    length = max_b sum(d_b).  (The reference builds a one-hot matrix and multiplies;
    for finite x that is exactly a gather.)"""
    B, Tp, D = x.shape
    d = durations.to(torch.int64)
    tot = d.sum(-1)
    L = int(tot.max())
    out = torch.zeros(B, L, D, dtype=x.dtype)
    for b in range(B):
        idx = torch.repeat_interleave(torch.arange(Tp), d[b])
        out[b, : idx.numel()] = x[b, idx]
    return out


# ----------------------------------------------------------------------------- a9
def adm_infer(sd, tc_latent, cfg, return_raw=False):
    """MegaADM.infer (models/megatts2.py:257-275): batch-1 AR regression with NON-causal
    full recompute each step; feeds back RAW float predictions; final
    (p + 0.5) -> int32 -> clamp(1,128).  Generalised to B independent rows."""
    B, T, _ = tc_latent.shape
    p = torch.zeros(B, 1, 1)
    w_dt, w_tc, w_out = sd("dt_linear_emb.weight"), sd("tc_linear_emb.weight"), sd("predict_layer.weight")
    for t in range(T):
        dt_emb = F.linear(p, w_dt)
        tc_emb = F.linear(tc_latent[:, : t + 1], w_tc)
        x = sine_pe_add(torch.cat([tc_emb, dt_emb], -1), sd("pos_emb.alpha"))
        x = encoder(sd.sub("adm"), x, cfg["n_layers"], cfg["n_heads"], False)
        y = F.linear(x, w_out)[:, -1:, :]
        p = torch.cat([p, y], 1)
    raw = p[:, 1:, :]
    dur = (raw + 0.5).to(torch.int32).clamp(1, 128)
    return (dur, raw) if return_raw else dur


# ----------------------------------------------------------------------------- a10
def plm_step_logits(sd, tc_latent, codes_in, cfg):
    """One body of the MegaPLM.infer loop (models/megatts2.py:172-178): codes_in (B,t+1)
    (BOS first) with tc_latent[:, :t+1] -> logits of the LAST position (B,1024)."""
    t1 = codes_in.shape[1]
    pc = F.embedding(codes_in, sd("pc_embedding.weight"))
    x = sine_pe_add(torch.cat([tc_latent[:, :t1], pc], -1), sd("pos.alpha"))
    x = encoder(sd.sub("plm"), x, cfg["n_layers"], cfg["n_heads"], False)
    return F.linear(x[:, -1], sd("predict_layer.weight"))


def plm_infer(sd, tc_latent, cfg, return_logits=False):
    """MegaPLM.infer (models/megatts2.py:165-181): greedy, BOS=1024, exactly T steps,
    non-causal full recompute.  Generalised to B independent rows."""
    B, T, _ = tc_latent.shape
    codes = torch.full((B, 1), cfg["vq_bins"], dtype=torch.int64)
    all_logits = []
    for t in range(T):
        lg = plm_step_logits(sd, tc_latent, codes, cfg)
        all_logits.append(lg)
        codes = torch.cat([codes, lg.argmax(-1, keepdim=True)], 1)
    out = codes[:, 1:]
    return (out, torch.stack(all_logits, 1)) if return_logits else out


def plm_forward(sd, tc_latent, p_codes, lens, cfg):
    """MegaPLM.forward (models/megatts2.py:148-163): teacher-forced, causal + padding mask."""
    pc = F.embedding(p_codes[:, :-1], sd("pc_embedding.weight"))
    x = sine_pe_add(torch.cat([tc_latent, pc], -1), sd("pos.alpha"))
    x = encoder(sd.sub("plm"), x, cfg["n_layers"], cfg["n_heads"], False, lens=lens, causal=True)
    return F.linear(x, sd("predict_layer.weight")), p_codes[:, 1:]


def adm_forward(sd, tc_latents, duration_tokens, lens, cfg):
    """MegaADM.forward (models/megatts2.py:233-255)."""
    dt_emb = F.linear(duration_tokens[:, :-1], sd("dt_linear_emb.weight"))
    tc_emb = F.linear(tc_latents, sd("tc_linear_emb.weight"))
    x = sine_pe_add(torch.cat([tc_emb, dt_emb], -1), sd("pos_emb.alpha"))
    x = encoder(sd.sub("adm"), x, cfg["n_layers"], cfg["n_heads"], False, lens=lens, causal=True)
    return F.linear(x, sd("predict_layer.weight"))[..., 0], duration_tokens[:, 1:, 0]


# ----------------------------------------------------------------------------- 8f-1 (next row)
def plm_infer_causal(sd, tc_latent, cfg, return_logits=False):
    """Checker for the OPT-IN causal decode: greedy loop whose step t takes the last row of the
    teacher-forced CAUSAL forward (MegaPLM.forward, models/megatts2.py:148-163) over the prefix generated so far.
    With a causal mask rows < t do not depend on later rows, so this equals a KV-cache decode by construction."""
    B, T, _ = tc_latent.shape
    codes = torch.full((B, 1), cfg["vq_bins"], dtype=torch.int64)
    all_logits = []
    for t in range(T):
        pcodes = torch.cat([codes, codes[:, :1]], 1)                      # forward() drops the last column
        lens = torch.full((B,), t + 1, dtype=torch.int32)
        lg = plm_forward(sd, tc_latent[:, : t + 1], pcodes, lens, cfg)[0][:, -1]
        all_logits.append(lg)
        codes = torch.cat([codes, lg.argmax(-1, keepdim=True)], 1)
    out = codes[:, 1:]
    return (out, torch.stack(all_logits, 1)) if return_logits else out


def adm_infer_causal(sd, tc_latent, cfg, return_raw=False):
    """Checker for the OPT-IN causal duration decode: MegaADM.forward's causal stack (models/megatts2.py:233-255)
    driven with infer()'s raw-float feedback and final rounding (models/megatts2.py:262-275)."""
    B, T, _ = tc_latent.shape
    p = torch.zeros(B, 1, 1)
    for t in range(T):
        dtok = torch.cat([p, p[:, :1]], 1)                                # forward() drops the last row
        lens = torch.full((B,), t + 1, dtype=torch.int32)
        y = adm_forward(sd, tc_latent[:, : t + 1], dtok, lens, cfg)[0][:, -1]
        p = torch.cat([p, y.reshape(B, 1, 1)], 1)
    raw = p[:, 1:, :]
    dur = (raw + 0.5).to(torch.int32).clamp(1, 128)
    return (dur, raw) if return_raw else dur


# ----------------------------------------------------------------------------- a12
def mel_decode(gsd, tc_latent_expand, p_codes, cfg):
    """Glue + MegaG.decoder of Megatts.forward (models/megatts2.py:361-368)."""
    embed = gsd("vqpe.vq.vq.layers.0._codebook.embed")
    zq = vq_decode(p_codes.unsqueeze(0), embed)                        # (B,256,T8)
    B, D, T8 = zq.shape
    zq = zq.transpose(1, 2).unsqueeze(2).expand(-1, -1, 8, -1).reshape(B, T8 * 8, D)
    L = tc_latent_expand.shape[1]
    x = torch.cat([tc_latent_expand, zq[:, :L]], -1).transpose(1, 2)
    return convnet(gsd.sub("decoder"), x, cfg["dec_kernel"], cfg["dec_n_stack"], cfg["dec_n_block"])


# ----------------------------------------------------------------------------- a13
def _reflect_same_conv(x, w, b, dilation=1):
    """speechbrain.nnet.CNN.Conv1d(padding='same', padding_mode='reflect') [memory]."""
    k = w.shape[-1]
    p = dilation * (k - 1) // 2
    return F.conv1d(F.pad(x, (p, p), mode="reflect"), w, b, dilation=dilation)


def hifigan_generator(sd, mel, cfg):
    """speechbrain HifiganGenerator.inference [memory; SURVEY.md §2.4 K13, §8c]:
    replicate-pad 5 frames each side, conv_pre(k7), 4x [lrelu(0.1) -> ConvTranspose1d ->
    mean of 3 ResBlock1], lrelu(0.01) -> conv_post(k7) -> tanh.
    mel (B,80,T) -> (B,1,256*(T+10))."""
    sd = sd if isinstance(sd, SD) else SD(sd)
    pad = cfg["inference_padding"]
    o = F.pad(mel, (pad, pad), mode="replicate")
    o = _reflect_same_conv(o, sd("conv_pre.weight"), sd("conv_pre.bias"))
    nk = len(cfg["resblock_kernel_sizes"])
    for i, (u, k) in enumerate(zip(cfg["upsample_factors"], cfg["upsample_kernel_sizes"])):
        o = F.leaky_relu(o, 0.1)
        o = F.conv_transpose1d(o, sd(f"ups.{i}.weight"), sd(f"ups.{i}.bias"), stride=u, padding=(k - u) // 2)
        z = None
        for j in range(nk):
            rb = sd.sub(f"resblocks.{i * nk + j}")
            x = o
            for m, dil in enumerate(cfg["resblock_dilation_sizes"][j]):
                xt = F.leaky_relu(x, 0.1)
                xt = _reflect_same_conv(xt, rb(f"convs1.{m}.weight"), rb(f"convs1.{m}.bias"), dil)
                xt = F.leaky_relu(xt, 0.1)
                xt = _reflect_same_conv(xt, rb(f"convs2.{m}.weight"), rb(f"convs2.{m}.bias"), 1)
                x = xt + x
            z = x if z is None else z + x
        o = z / nk
    o = F.leaky_relu(o)                                               # default slope 0.01
    o = _reflect_same_conv(o, sd("conv_post.weight"), sd("conv_post.bias"))
    return torch.tanh(o)


def hifigan_decode_batch(sd, mel):
    """speechbrain HIFIGAN.decode_batch(mel (B,80,T)) -> (B,1,samples) [memory]."""
    from . import weights
    with torch.no_grad():
        return hifigan_generator(sd, mel, weights.HIFIGAN_CFG)


# ----------------------------------------------------------------------------- a14
def synthesize(gsd, plmsd, admsd, hsd, phone, mel_prompt, cfgs, forced_durations=None):
    """Tensor-level body of Megatts.forward (models/megatts2.py:353-373) for B independent
    utterances: tc_latent -> adm.infer -> length regulate -> max-pool 8 -> plm.infer ->
    vq.decode + mel decoder -> HiFi-GAN.  ``forced_durations`` (B,Tp) replaces the ADM
    prediction for shape control (the ADM is still run and returned)."""
    gcfg, pcfg, acfg, hcfg = cfgs
    g = SD(gsd)
    with torch.no_grad():
        tc, _, _ = mrte_tc_latent(g.sub("mrte"), phone, mel_prompt, gcfg)
        dt = adm_infer(SD(admsd), tc, acfg)[..., 0]
        d_used = dt if forced_durations is None else forced_durations
        tc_exp = length_regulate(tc, d_used)
        tc8 = F.max_pool1d(tc_exp.transpose(1, 2), 8, ceil_mode=True).transpose(1, 2)
        codes = plm_infer(SD(plmsd), tc8, pcfg)
        mel = mel_decode(g, tc_exp, codes, gcfg)
        wav = hifigan_generator(SD(hsd), mel, hcfg)
    return dict(tc_latent=tc, dt=dt, tc_latent_expand=tc_exp, tc8=tc8, p_codes=codes, mel=mel, wav=wav)
