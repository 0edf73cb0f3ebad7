"""GPU: mem_b200.engine_for_finetuning against the UNMODIFIED reference finetuning loop
(tests/golden/engine_ft_tiny.npz, CPU fp32: four micro-batches, update_freq 2, two AdamW steps).
Tolerance: bf16 tensor-core GEMMs vs fp32 -- loss within 2 %, grad norm within 6 %, accuracy within one sample,
weights after 2 AdamW steps: mean |diff| < 4e-4, max < 6e-3 (= sum of 2*lr)."""
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from mem_b200 import engine_for_finetuning as eft, optim_factory, registry, utils
from mem_b200 import modeling_finetune  # noqa: F401
from oracle import engine_ref, vit_ref

pytestmark = pytest.mark.gpu


def _build():
    model = registry.create_model("ft_vit", **vit_ref.TINY_FT)
    model.load_state_dict(vit_ref.synth_state_dict(model.state_dict(), seed=41))
    model.cuda()
    args = SimpleNamespace(opt="adamw", weight_decay=engine_ref.FT_WD[0], lr=engine_ref.FT_LR[0], opt_eps=1e-8, opt_betas=None)
    return model, optim_factory.create_optimizer(args, model)


def test_train_one_epoch_matches_reference_golden(golden_dir, capsys):
    gold = np.load(os.path.join(golden_dir, "engine_ft_tiny.npz"))
    model, opt = _build()
    stats = eft.train_one_epoch(SimpleNamespace(), model, torch.nn.CrossEntropyLoss(), engine_ref.synth_class_batches(), opt,
                                torch.device("cuda"), 0, utils.NativeScalerWithGradNormCount(), engine_ref.MAX_NORM,
                                start_steps=0, lr_schedule_values=engine_ref.FT_LR, wd_schedule_values=engine_ref.FT_WD,
                                num_training_steps_per_epoch=2, update_freq=engine_ref.FT_UPDATE_FREQ)
    assert sorted(stats) == sorted(gold["keys"].tolist())
    g = {k: float(gold["stat/" + k]) for k in stats}
    assert abs(stats["loss"] - g["loss"]) < 2e-2 * g["loss"], (stats, g)
    assert abs(stats["grad_norm"] - g["grad_norm"]) < 6e-2 * g["grad_norm"], (stats, g)
    assert abs(stats["class_acc"] - g["class_acc"]) <= 1.0 / 24 + 1e-6, (stats, g)
    assert stats["lr"] == pytest.approx(g["lr"]) and stats["min_lr"] == pytest.approx(g["min_lr"])
    assert stats["weight_decay"] == pytest.approx(g["weight_decay"]) and stats["loss_scale"] == 1.0
    sd = model.state_dict()
    for k in gold.files:
        if k.startswith("final/"):
            d = np.abs(sd[k[6:]].detach().float().cpu().numpy().reshape(-1)[:512] - gold[k])
            assert d.mean() < 4e-4 and d.max() < 6e-3, (k, float(d.mean()), float(d.max()))
    # evaluation logits of the trained model vs the reference's
    model.eval()
    with torch.no_grad():
        logits = model(engine_ref.synth_class_batches()[0][0].cuda()).float().cpu().numpy()
    assert np.abs(logits - gold["eval_logits"]).max() < 0.05 * max(1.0, np.abs(gold["eval_logits"]).max())
    capsys.readouterr()


def test_accumulation_equals_one_big_batch(capsys):
    """update_freq = 2 over two micro-batches gives the gradient (hence the step) of their concatenation."""
    batches = engine_ref.synth_class_batches()[:2]
    scaler = utils.NativeScalerWithGradNormCount()
    m1, o1 = _build()
    s1 = eft.train_one_epoch(None, m1, torch.nn.CrossEntropyLoss(), batches, o1, "cuda", 0, scaler, 1.0, update_freq=2)
    m2, o2 = _build()
    big = [(torch.cat([b[0] for b in batches]), torch.cat([b[1] for b in batches]))]
    s2 = eft.train_one_epoch(None, m2, torch.nn.CrossEntropyLoss(), big, o2, "cuda", 0, scaler, 1.0, update_freq=1)
    assert abs(s1["grad_norm"] - s2["grad_norm"]) < 1e-2 * s2["grad_norm"]
    assert abs(s1["loss"] - s2["loss"]) < 1e-3 * s2["loss"]
    for (n, a), (_, b) in zip(m1.state_dict().items(), m2.state_dict().items()):
        if a.is_floating_point():
            assert (a - b).abs().mean().item() < 2e-4, n
    capsys.readouterr()


def test_evaluate_ema_mixup_and_soft_targets(capsys):
    model, opt = _build()
    batches = engine_ref.synth_class_batches()
    ev = eft.evaluate([(b[0], b[1]) for b in batches[:2]], model, "cuda")
    assert set(ev) == {"loss", "acc1", "acc5"}
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    want_loss, want_acc, n = 0.0, 0.0, 0
    for x, t in batches[:2]:
        logits = vit_ref.classify_logits(x, sd, 2, 16)
        want_loss += torch.nn.functional.cross_entropy(logits, t).item()
        want_acc += 100.0 * (logits.argmax(1) == t).float().sum().item()
        n += len(t)
    assert abs(ev["loss"] - want_loss / 2) < 2e-2 * want_loss / 2
    assert abs(ev["acc1"] - want_acc / n) <= 100.0 / n + 1e-4 and ev["acc5"] == pytest.approx(100.0)
    # soft targets through a mixup callable + EMA update every optimizer step
    ema = eft.ModelEma(model, decay=0.5)
    before = {k: v.clone() for k, v in ema.state_dict().items()}

    def mixup(x, t):
        soft = torch.nn.functional.one_hot(t, 2).float()
        return 0.7 * x + 0.3 * x.flip(0), 0.7 * soft + 0.3 * soft.flip(0)

    st = eft.train_one_epoch(None, model, eft.SoftTargetCrossEntropy(), batches[:2], opt, "cuda", 0,
                             utils.NativeScalerWithGradNormCount(), 1.0, model_ema=ema, mixup_fn=mixup, update_freq=1)
    assert "class_acc" not in st and np.isfinite(st["loss"]) and st["grad_norm"] > 0
    after, cur = ema.state_dict(), model.state_dict()
    k = "head.weight"
    assert not torch.equal(after[k], before[k])
    # two updates with decay 0.5 from shadow s0 and weights w1, w2: s2 = 0.25 s0 + 0.25 w1 + 0.5 w2 -> between s0 and w2
    assert (after[k] - cur[k]).abs().max() <= (before[k] - cur[k]).abs().max() + 1e-6
    with pytest.raises(NotImplementedError):
        eft.train_one_epoch(None, model, torch.nn.CrossEntropyLoss(), batches[:1], opt, "cuda", 0, None)
    capsys.readouterr()


def test_criteria_kernel_matches_timm_formulas():
    """memb_soft_ce behind LabelSmoothingCrossEntropy / SoftTargetCrossEntropy: loss and gradient vs the oracle's
    restatement of timm's formulas (fp32), ragged class counts (C not a multiple of the warp) included."""
    from oracle import engine_ref
    g = torch.Generator().manual_seed(3)
    for B, C in ((7, 5), (128, 2), (33, 101), (4, 1000)):
        x = torch.randn(B, C, generator=g) * 3
        t = torch.randint(0, C, (B,), generator=g)
        soft = torch.softmax(torch.randn(B, C, generator=g), -1) * 0.9      # rows need not sum to one
        for name, ours, ref in (
                ("smooth", lambda z: eft.LabelSmoothingCrossEntropy(0.1)(z, t.cuda()), lambda z: engine_ref.label_smoothing_ce(z, t, 0.1)),
                ("soft", lambda z: eft.SoftTargetCrossEntropy()(z, soft.cuda()), lambda z: engine_ref.soft_target_ce(z, soft))):
            xc = x.clone().cuda().requires_grad_(True)
            xr = x.clone().requires_grad_(True)
            lo, lr = ours(xc), ref(xr)
            (lo * 1.7).backward()
            (lr * 1.7).backward()
            assert abs(lo.item() - lr.item()) < 2e-6 * max(1.0, abs(lr.item())), (name, B, C)
            assert torch.allclose(xc.grad.cpu(), xr.grad, rtol=1e-5, atol=1e-7), (name, B, C)
    with torch.no_grad():                                                   # no gradient requested: loss only
        assert eft.LabelSmoothingCrossEntropy(0.0)(x.cuda(), t.cuda()).item() == pytest.approx(
            torch.nn.functional.cross_entropy(x, t).item(), rel=1e-6)


def test_checkpoint_round_trip_keeps_model_ema(tmp_path):
    from types import SimpleNamespace
    from mem_b200 import optim_factory
    model = registry.create_model("ft_vit", **vit_ref.TINY_FT).cuda()
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        opt = optim_factory.create_optimizer(SimpleNamespace(opt="adamw", weight_decay=0.05, lr=1e-3, opt_eps=1e-8), model)
    ema = eft.ModelEma(model, decay=0.9)
    with torch.no_grad():
        model.head.weight.add_(1.0)
    ema.update(model)
    want = {k: v.clone() for k, v in ema.state_dict().items()}
    args = SimpleNamespace(output_dir=str(tmp_path), resume="", auto_resume=True, model_ema=True, epochs=10, start_epoch=0)
    utils.save_model(args, "best", model, model, opt, utils.NativeScalerWithGradNormCount(), model_ema=ema)
    ema2 = eft.ModelEma(model, decay=0.9)                    # starts from the current weights, not the averaged ones
    assert not torch.equal(ema2.state_dict()["head.weight"], want["head.weight"])
    args.resume = str(tmp_path / "checkpoint-best.pth")
    with contextlib.redirect_stdout(io.StringIO()):
        utils.auto_load_model(args, model, model, opt, utils.NativeScalerWithGradNormCount(), model_ema=ema2)
    assert args.start_epoch == 11                            # "best" resumes after the last epoch (mem/utils.py:519)
    for k, v in want.items():
        assert torch.equal(ema2.state_dict()[k], v), k
