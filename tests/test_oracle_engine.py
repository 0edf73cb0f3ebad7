"""CPU: pin oracle/engine_ref.py (fp32 restatement of one MEM step: tokenise, masked CE, clip, AdamW with the
reference's parameter groups) to three steps of the unmodified reference train_one_epoch
(tests/golden/engine_tiny.npz); host-side checks of utils / optim_factory API."""
import os

import pytest

import numpy as np
import torch

from mem_b200 import optim_factory, registry, utils
from mem_b200 import modeling_pretrain  # noqa: F401
from mem_b200.vae_model import DiscreteVAE
from oracle import dvae_ref, engine_ref, vit_ref


def test_engine_oracle_matches_reference_golden(golden_dir):
    gold = np.load(os.path.join(golden_dir, "engine_tiny.npz"))
    model = registry.create_model("pt_vit", **vit_ref.TINY)
    vit_sd = vit_ref.synth_state_dict(model.state_dict(), seed=31)
    vae_sd = dvae_ref.synth_state_dict(DiscreteVAE(**engine_ref.TINY_VAE).state_dict(), seed=32, head_gain=4.0)
    steps, final = engine_ref.run_steps(vit_sd, vae_sd, engine_ref.synth_batches())
    for it, st in enumerate(steps):
        assert abs(st["loss"] - float(gold[f"step{it}/loss"])) < 2e-5, it
        assert abs(st["grad_norm"] - float(gold[f"step{it}/grad_norm"])) < 2e-4, it
        assert abs(st["mlm_acc"] - float(gold[f"step{it}/mlm_acc"])) < 1e-6, it
    for k in gold.files:
        if k.startswith("final/"):
            np.testing.assert_allclose(final[k[6:]].numpy().reshape(-1)[:512], gold[k], rtol=1e-4, atol=1e-6)
    assert sorted(gold["keys"].tolist()) == ["grad_norm", "loss", "loss_scale", "lr", "min_lr", "mlm_acc", "weight_decay"]


def test_finetune_oracle_matches_reference_golden(golden_dir):
    """oracle/engine_ref.run_finetune vs the unmodified reference finetuning loop (tests/golden/engine_ft_tiny.npz)."""
    from mem_b200 import modeling_finetune  # noqa: F401
    gold = np.load(os.path.join(golden_dir, "engine_ft_tiny.npz"))
    model = registry.create_model("ft_vit", **vit_ref.TINY_FT)
    sd = vit_ref.synth_state_dict(model.state_dict(), seed=41)
    stats, final = engine_ref.run_finetune(sd, engine_ref.synth_class_batches())
    assert abs(stats["loss"] - float(gold["stat/loss"])) < 2e-5
    assert abs(stats["class_acc"] - float(gold["stat/class_acc"])) < 1e-6
    assert abs(stats["grad_norm"] - float(gold["stat/grad_norm"])) < 5e-4
    for k in gold.files:
        if k.startswith("final/"):
            np.testing.assert_allclose(final[k[6:]].numpy().reshape(-1)[:512], gold[k], rtol=1e-4, atol=2e-6)
    with torch.no_grad():
        logits = vit_ref.classify_logits(engine_ref.synth_class_batches()[0][0], final, 2, 16).numpy()
    np.testing.assert_allclose(logits, gold["eval_logits"], rtol=1e-4, atol=1e-5)
    assert sorted(gold["keys"].tolist()) == ["class_acc", "grad_norm", "loss", "loss_scale", "lr", "min_lr", "weight_decay"]


def test_finetune_losses_and_accuracy_helpers():
    from mem_b200.engine_for_finetuning import LabelSmoothingCrossEntropy, accuracy
    g = torch.Generator().manual_seed(0)
    x = torch.randn(7, 5, generator=g)
    t = torch.randint(0, 5, (7,), generator=g)
    # the oracle's restatement of timm's criteria against torch's own label smoothing (same definition)
    want = torch.nn.functional.cross_entropy(x, t, label_smoothing=0.1)
    assert abs(engine_ref.label_smoothing_ce(x, t, 0.1).item() - want.item()) < 1e-6
    soft = torch.nn.functional.one_hot(t, 5).float()
    assert abs(engine_ref.soft_target_ce(x, soft).item() - torch.nn.functional.cross_entropy(x, t).item()) < 1e-6
    with pytest.raises(RuntimeError):        # the product criteria are CUDA kernels: no CPU path
        LabelSmoothingCrossEntropy(0.1)(x, t)
    a1, a3 = accuracy(x, t, topk=(1, 3))
    top3 = x.topk(3, dim=1).indices
    assert abs(a1.item() - 100.0 * (x.argmax(1) == t).float().mean().item()) < 1e-4
    assert abs(a3.item() - 100.0 * (top3 == t[:, None]).any(1).float().mean().item()) < 1e-4


def test_cosine_scheduler_and_meters():
    import math
    s = utils.cosine_scheduler(1e-3, 1e-5, epochs=4, niter_per_ep=10, warmup_epochs=1, start_warmup_value=1e-6)
    assert len(s) == 40 and abs(s[0] - 1e-6) < 1e-12 and abs(s[9] - 1e-3) < 1e-12
    assert abs(s[10] - 1e-3) < 1e-12 and abs(s[25] - (1e-5 + 0.5 * (1e-3 - 1e-5) * (1 + math.cos(math.pi * 15 / 30)))) < 1e-12
    s2 = utils.cosine_scheduler(0.05, 0.05, epochs=2, niter_per_ep=3)
    assert np.allclose(s2, 0.05)
    m = utils.SmoothedValue(window_size=3)
    for v in (1.0, 2.0, 6.0, 3.0):
        m.update(v)
    assert m.global_avg == 3.0 and m.median == 3.0 and m.max == 6.0 and m.value == 3.0
    ml = utils.MetricLogger()
    ml.update(loss=2.0, skipped=None)
    ml.update(loss=torch.tensor(4.0))
    assert ml.loss.global_avg == 3.0 and "skipped" not in ml.meters
    sc = utils.NativeScalerWithGradNormCount()
    assert sc.state_dict()["scale"] == 1.0


def test_parameter_groups_follow_reference_rules(capsys):
    model = registry.create_model("pt_vit", **vit_ref.TINY)
    groups = optim_factory.get_parameter_groups(model, 0.05, model.no_weight_decay())
    capsys.readouterr()
    assert len(groups) == 2
    nd = next(g for g in groups if g["weight_decay"] == 0.0)
    dc = next(g for g in groups if g["weight_decay"] == 0.05)
    names = {id(p): n for n, p in model.named_parameters()}
    nd_names = {names[id(p)] for p in nd["params"]}
    dc_names = {names[id(p)] for p in dc["params"]}
    assert "cls_token" in nd_names and "blocks.0.gamma_1" in nd_names and "lm_head.bias" in nd_names and "blocks.1.attn.q_bias" in nd_names
    assert "mask_token" in dc_names and "lm_head.weight" in dc_names and "rel_pos_bias.relative_position_bias_table" in dc_names
    assert all(g["lr_scale"] == 1.0 for g in groups)
    # same split as the oracle's statement of optim_factory.py:56-95
    ref = engine_ref.param_groups(list(model.named_parameters()), 0.05)
    assert {names[id(p)] for p in ref[0]["params"]} == nd_names and {names[id(p)] for p in ref[1]["params"]} == dc_names
