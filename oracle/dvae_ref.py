"""ORACLE (test infrastructure, not product): fp32 restatement of the reference dVAE tokenizer.

Functional PyTorch over a ``state_dict`` with the reference's key names.  Follows
``eventvae/vae/vae_model.py``: encoder stack construction :77-104 (L x [Conv2d(4, stride 2, pad 1) +
ReLU], R x ResBlock :29-41, Conv2d 1x1 to ``num_tokens``), ``norm`` :133-141,
``forward(return_logits=True)`` :182-189 and ``get_codebook_indices`` :153-158
(``logits.argmax(dim=1).flatten(1)``).

Pinned by ``tests/golden/dvae_tiny.npz`` (logits and indices of the UNMODIFIED reference class,
made by ``oracle/make_golden.py``) in tests/test_oracle_dvae.py.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

TINY_A = dict(input_H=32, input_W=32, num_tokens=64, codebook_dim=16, num_layers=2, num_resnet_blocks=1, hidden_dim=32,
              channels=2)
TINY_B = dict(input_H=64, input_W=96, num_tokens=256, codebook_dim=16, num_layers=4, num_resnet_blocks=3, hidden_dim=64,
              channels=3, normalization=((0.5, 0.0, 0.5), (0.5, 1.0, 0.25)))
TINY_C = dict(input_H=16, input_W=16, num_tokens=96, codebook_dim=8, num_layers=1, num_resnet_blocks=0, hidden_dim=32,
              channels=2)


def encoder_logits(img, sd, num_layers, num_resnet_blocks, normalization=None):
    x = img
    if normalization is not None:
        mean, std = (torch.as_tensor(t).to(img).view(1, -1, 1, 1) for t in normalization)
        x = (x - mean) / std
    for i in range(num_layers):
        x = F.relu(F.conv2d(x, sd[f"encoder.{i}.0.weight"], sd[f"encoder.{i}.0.bias"], stride=2, padding=1))
    for j in range(num_resnet_blocks):
        p = f"encoder.{num_layers + j}.net."
        y = F.relu(F.conv2d(x, sd[p + "0.weight"], sd[p + "0.bias"], padding=1))
        y = F.relu(F.conv2d(y, sd[p + "2.weight"], sd[p + "2.bias"], padding=1))
        x = F.conv2d(y, sd[p + "4.weight"], sd[p + "4.bias"]) + x
    k = num_layers + num_resnet_blocks
    return F.conv2d(x, sd[f"encoder.{k}.weight"], sd[f"encoder.{k}.bias"])


def codebook_indices(img, sd, num_layers, num_resnet_blocks, normalization=None):
    return encoder_logits(img, sd, num_layers, num_resnet_blocks, normalization).argmax(dim=1).flatten(1)


def synth_state_dict(template_sd, seed, head_gain=1.0):
    """Deterministic weights (sorted key order, one generator): conv weights ~ N(0, 1/fan_in) so that
    activations keep unit scale through the stack; ``head_gain`` widens the logit spread."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    keys = sorted(template_sd)
    last = max(int(k.split(".")[1]) for k in keys if k.startswith("encoder."))
    for k in keys:
        t = template_sd[k]
        r = torch.randn(t.shape, generator=g, dtype=torch.float32)
        if k.endswith(".weight") and t.dim() == 4:
            fan_in = t.shape[1] * t.shape[2] * t.shape[3]
            r = r * (2.0 / fan_in) ** 0.5
            if k == f"encoder.{last}.weight":
                r = r * head_gain
        elif k.endswith(".bias"):
            r = r * 0.05
        out[k] = r
    return out


def synth_images(B, C, H, W, seed):
    """Sparse non-negative images shaped like normalised event histograms (values in [0, 1])."""
    g = torch.Generator().manual_seed(seed)
    img = torch.rand(B, C, H, W, generator=g)
    return torch.where(img > 0.8, (img - 0.8) / 0.2, torch.zeros_like(img))


# ---------------------------------------------------------------------------------------------
# Training path and decoder: vae_model.py:160-213
# ---------------------------------------------------------------------------------------------
TRAIN_A = dict(TINY_A, kl_div_loss_weight=0.01, temperature=0.9)
TRAIN_B = dict(input_H=32, input_W=48, num_tokens=128, codebook_dim=24, num_layers=3, num_resnet_blocks=2, hidden_dim=32,
               channels=3, loss="smooth_l1", straight_through=True, kl_div_loss_weight=0.05,
               normalization=((0.5, 0.0, 0.5), (0.5, 1.0, 0.25)))
TRAIN_C = dict(TINY_C, kl_div_loss_weight=0.0)           # no residual blocks: the decoder starts with a ConvTranspose2d


def decoder_images(z, sd, num_layers, num_resnet_blocks):
    """``self.decoder`` (vae_model.py:84-103): [Conv2d 1x1 codebook_dim -> hidden, R x ResBlock] if R > 0, then
    L x [ConvTranspose2d(4, stride 2, pad 1) + ReLU], Conv2d 1x1 to the image channels."""
    i = 0
    x = z
    if num_resnet_blocks > 0:
        x = F.conv2d(x, sd["decoder.0.weight"], sd["decoder.0.bias"])
        i = 1
        for _ in range(num_resnet_blocks):
            p = f"decoder.{i}.net."
            y = F.relu(F.conv2d(x, sd[p + "0.weight"], sd[p + "0.bias"], padding=1))
            y = F.relu(F.conv2d(y, sd[p + "2.weight"], sd[p + "2.bias"], padding=1))
            x = F.conv2d(y, sd[p + "4.weight"], sd[p + "4.bias"]) + x
            i += 1
    for _ in range(num_layers):
        x = F.relu(F.conv_transpose2d(x, sd[f"decoder.{i}.0.weight"], sd[f"decoder.{i}.0.bias"], stride=2, padding=1))
        i += 1
    return F.conv2d(x, sd[f"decoder.{i}.weight"], sd[f"decoder.{i}.bias"])


def gumbel_noise(shape, seed):
    """The Gumbel sample ``F.gumbel_softmax`` draws for logits of ``shape`` right after ``torch.manual_seed(seed)``
    (torch/nn/functional.py: ``-torch.empty_like(logits).exponential_().log()``)."""
    torch.manual_seed(seed)
    return -torch.empty(shape, dtype=torch.float32).exponential_().log()


def train_loss(img, sd, cfg, noise, temp=None):
    """(loss, recons) of ``DiscreteVAE.forward(img, return_loss=True, return_recons=True, temp)`` (vae_model.py:173-213)
    with the Gumbel sample given explicitly."""
    L, R = cfg["num_layers"], cfg["num_resnet_blocks"]
    norm = cfg.get("normalization")
    temp = cfg.get("temperature", 0.9) if temp is None else temp
    x = img
    if norm is not None:
        mean, std = (torch.as_tensor(t).to(img).view(1, -1, 1, 1) for t in norm)
        x = (x - mean) / std
    logits = encoder_logits(img, sd, L, R, norm)
    soft = F.softmax((logits + noise) / temp, dim=1)
    if cfg.get("straight_through", False):
        index = soft.max(1, keepdim=True)[1]
        hard = torch.zeros_like(logits).scatter_(1, index, 1.0)
        soft = hard - soft.detach() + soft
    sampled = torch.einsum("b n h w, n d -> b d h w", soft, sd["codebook.weight"])
    out = decoder_images(sampled, sd, L, R)
    loss_fn = {"mse": F.mse_loss, "smooth_l1": F.smooth_l1_loss}[cfg.get("loss", "mse")]
    recon = loss_fn(x, out)
    lq = F.log_softmax(logits.permute(0, 2, 3, 1).reshape(logits.shape[0], -1, logits.shape[1]), dim=-1)
    log_uniform = torch.log(torch.tensor([1.0 / cfg["num_tokens"]]))
    kl = F.kl_div(log_uniform, lq, None, None, "batchmean", log_target=True)
    return recon + kl * cfg.get("kl_div_loss_weight", 0.0), out


def decode(img_seq, sd, cfg):
    h, w = cfg["input_H"] >> cfg["num_layers"], cfg["input_W"] >> cfg["num_layers"]
    z = sd["codebook.weight"][img_seq]                               # [B, n, d]
    z = z.reshape(z.shape[0], h, w, -1).permute(0, 3, 1, 2)
    return decoder_images(z, sd, cfg["num_layers"], cfg["num_resnet_blocks"])


def synth_train_state_dict(template_sd, seed):
    """``synth_state_dict`` plus a unit-scale codebook and decoder weights scaled for ConvTranspose fan-in."""
    sd = synth_state_dict(template_sd, seed, head_gain=3.0)
    g = torch.Generator().manual_seed(seed + 1)
    sd["codebook.weight"] = torch.randn(template_sd["codebook.weight"].shape, generator=g)
    return sd
