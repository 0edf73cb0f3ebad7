"""GPU: the dVAE training step and decoder through libmemb against the reference goldens
(tests/golden/dvae_train.npz: loss, reconstruction, every parameter gradient and decode() of the UNMODIFIED reference
``DiscreteVAE`` in fp32 with a seeded Gumbel sample; oracle/make_golden.py ``golden_dvae_train``).

Tolerance: this path computes in bf16 (tensor-core GEMMs with fp32 accumulation), so -- as for the ViT -- every tensor is
bounded by TOL_MULT x the error of the reference's own bf16-autocast run against its fp32 run, stored per tensor in the
same golden file (floored at the case's median relative gradient error)."""
import os

import numpy as np
import pytest
import torch

from mem_b200.vae_model import DiscreteVAE
from oracle import dvae_ref

pytestmark = pytest.mark.gpu
TOL_MULT = 2.0
CASES = [("a", dvae_ref.TRAIN_A, 3, 61, 0.8), ("b", dvae_ref.TRAIN_B, 2, 62, None), ("c", dvae_ref.TRAIN_C, 4, 63, 1.0)]


def rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def _setup(name, cfg, B, seed):
    torch.manual_seed(0)
    vae = DiscreteVAE(**cfg)
    sd = dvae_ref.synth_train_state_dict(vae.state_dict(), seed)
    vae.load_state_dict(sd)
    img = dvae_ref.synth_images(B, cfg["channels"], cfg["input_H"], cfg["input_W"], seed + 100)
    return vae.cuda().train(), sd, img


@pytest.mark.parametrize("name,cfg,B,seed,temp", CASES)
def test_training_step_matches_reference(golden_dir, name, cfg, B, seed, temp):
    gold = np.load(os.path.join(golden_dir, "dvae_train.npz"))
    vae, sd, img = _setup(name, cfg, B, seed)
    noise = torch.from_numpy(gold[f"{name}/noise"]).cuda()
    loss, recons = vae(img.cuda(), return_loss=True, return_recons=True, temp=temp, gumbel_noise=noise)
    loss.backward()
    l32, l16 = (float(v) for v in gold[f"{name}/loss"])
    cal = {k[len(name) + 5:]: float(gold[k]) / max(float(gold[f"{name}/ref/" + k[len(name) + 5:]]), 1e-30)
           for k in gold.files if k.startswith(f"{name}/err/") and not k.endswith("/decode")}
    grads = {k: v for k, v in cal.items() if k != "recons"}
    med = float(np.median(list(grads.values())))
    assert recons.shape == img.shape
    assert rel(recons, torch.from_numpy(gold[f"{name}/recons"]).cuda()) <= TOL_MULT * max(cal["recons"], med)
    assert abs(loss.item() - l32) <= TOL_MULT * max(abs(l16 - l32), 0.1 * max(cal["recons"], med) * abs(l32)), (loss.item(), l32, l16)
    bad, worst = [], 0.0
    for n, p in vae.named_parameters():
        assert p.grad is not None and p.grad.shape == p.shape, n
        want = torch.from_numpy(gold[f"{name}/grad/{n}"]).cuda()
        bound = TOL_MULT * max(grads[n], med)
        e = rel(p.grad, want)
        worst = max(worst, e / bound)
        if e > bound:
            bad.append((n, f"{e:.3e}", f"reference bf16 {grads[n]:.3e}"))
    print(f"[dvae {name}] worst gradient tensor at {worst * TOL_MULT:.2f}x the reference's bf16 error")
    assert not bad, bad[:8]
    # inference forms of the same call
    with torch.no_grad():
        r2 = vae(img.cuda(), temp=temp, gumbel_noise=noise)
        l2 = vae(img.cuda(), return_loss=True, temp=temp, gumbel_noise=noise)
    assert torch.equal(r2, recons) and l2.item() == pytest.approx(loss.item(), rel=1e-5)      # (atomic summation order)


@pytest.mark.parametrize("name,cfg,B,seed,temp", CASES)
def test_decode_matches_reference(golden_dir, name, cfg, B, seed, temp):
    gold = np.load(os.path.join(golden_dir, "dvae_train.npz"))
    vae, sd, img = _setup(name, cfg, B, seed)
    seq = torch.from_numpy(gold[f"{name}/seq"]).cuda()
    want = torch.from_numpy(gold[f"{name}/decode"]).cuda()
    got = vae.decode(seq)
    assert got.shape == want.shape and got.dtype == torch.float32
    bound = TOL_MULT * max(float(gold[f"{name}/err/decode"]) / want.double().norm().item(), 4e-3)
    assert rel(got, want) <= bound
    seq[0, 0] = cfg["num_tokens"]
    with pytest.raises(IndexError):
        vae.decode(seq)


def test_gumbel_sample_follows_the_torch_generator_like_the_reference():
    """F.gumbel_softmax draws ``-empty_like(logits).exponential_().log()`` from the global CUDA generator; the drop-in draws
    the same sample, so a seeded reference run and a seeded run of this build see the same noise."""
    name, cfg, B, seed, temp = CASES[0]
    vae, _, img = _setup(name, cfg, B, seed)
    h, w = cfg["input_H"] >> cfg["num_layers"], cfg["input_W"] >> cfg["num_layers"]
    torch.manual_seed(77)
    noise = -torch.empty(B, cfg["num_tokens"], h, w, device="cuda").exponential_().log()
    with torch.no_grad():
        want = vae(img.cuda(), return_loss=True, temp=temp, gumbel_noise=noise).item()
        torch.manual_seed(77)
        got = vae(img.cuda(), return_loss=True, temp=temp).item()
    assert got == pytest.approx(want, rel=1e-5)


def test_training_loop_of_the_reference_runs_and_learns():
    """eventvae/train_vae.py:304-392 in miniature: Adam + clip on ``vae.parameters()``, loss falls."""
    name, cfg, B, seed, temp = CASES[0]
    vae, _, img = _setup(name, dict(cfg, kl_div_loss_weight=0.0), 8, seed)
    opt = torch.optim.Adam(vae.parameters(), lr=3e-3)
    torch.manual_seed(5)
    losses = []
    for _ in range(40):
        loss, recons = vae(img.cuda(), return_loss=True, return_recons=True, temp=0.9)
        opt.zero_grad()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(vae.parameters(), 1.0)
        opt.step()
        losses.append(loss.item())
    assert np.isfinite(losses).all() and np.mean(losses[-5:]) < 0.7 * np.mean(losses[:5]), losses[::8]
    # the tokenizer path sees the trained weights (same module, same parameters)
    idx = vae.get_codebook_indices(img.cuda())
    assert idx.shape == (8, (cfg["input_H"] >> cfg["num_layers"]) * (cfg["input_W"] >> cfg["num_layers"]))
