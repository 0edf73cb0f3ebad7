"""CPU: the pretrain -> finetune seam against goldens the UNMODIFIED reference produced
(oracle/make_golden.py ``golden_finetune_remap``): ``utils.finetune`` / ``load_state_dict`` (mem/utils.py:302-348,
:613-732) and ``LayerDecayValueAssigner`` / ``get_num_layer_for_vit`` (mem/optim_factory.py:31-53)."""
import contextlib
import io
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from mem_b200 import modeling_finetune, modeling_pretrain, optim_factory, registry, utils  # noqa: F401
from oracle import vit_ref


def _check_against_golden(gold, case, sd, atol=0.0):
    seen = 0
    for key in gold.files:
        if not key.startswith(case + "/") or key.endswith("/log"):
            continue
        _, kind, name = key.split("/", 2)
        got = sd[name].detach().cpu().numpy()
        if kind == "full":
            assert got.shape == gold[key].shape, name
            assert np.allclose(got, gold[key], rtol=0.0, atol=atol), (name, np.abs(got - gold[key]).max())
        else:
            f = got.reshape(-1).astype(np.float64)
            digest = np.concatenate([[f.sum(), np.sqrt((f ** 2).sum())], f[:64]])
            assert np.allclose(digest, gold[key], rtol=1e-12, atol=atol), name
        seen += 1
    assert seen == len(sd), (seen, len(sd))


@pytest.mark.parametrize("case,pt_over,ft_over,atol", [
    ("same", dict(in_chans=3), dict(), 0.0),
    # spline evaluation goes through the same FITPACK routine as the golden's: identical up to the last bits
    ("interp", dict(in_chans=3, use_abs_pos_emb=True), dict(img_size=(160, 160), use_abs_pos_emb=True), 1e-6),
])
def test_finetune_checkpoint_remap_matches_reference(golden_dir, tmp_path, case, pt_over, ft_over, atol):
    gold = np.load(os.path.join(golden_dir, "finetune_remap.npz"))
    torch.manual_seed(0)
    pt = registry.create_model("pt_vit", **dict(vit_ref.TINY, **pt_over))
    pt.load_state_dict(vit_ref.synth_state_dict(pt.state_dict(), seed=51))
    path = str(tmp_path / (case + ".pth"))
    torch.save({"model": pt.state_dict(), "epoch": 3}, path)
    ft = registry.create_model("ft_vit", **dict(vit_ref.TINY_FT, **ft_over))
    ft.load_state_dict(vit_ref.synth_state_dict(ft.state_dict(), seed=52))
    log = io.StringIO()
    with contextlib.redirect_stdout(log):
        utils.finetune(SimpleNamespace(finetune=path, model_key="model|module", model_prefix=""), ft)
    _check_against_golden(gold, case, ft.state_dict(), atol)
    # the report lines a user of the reference greps for
    ref_log = str(gold[case + "/log"])
    for line in ("Load state_dict by model_key = model", "Expand the shared relative position embedding to each transformer block.",
                 "Weights from pretrained model not used in VisionTransformer"):
        assert (line in ref_log) == (line in log.getvalue()), line
    assert log.getvalue().count("Position interpolate") == ref_log.count("Position interpolate")
    # every parameter is still a view of the model's flat buffer after the load (when one exists)
    from mem_b200.vit_engine import engine_of
    assert engine_of(ft).flat().valid()


def test_layer_decay_assigner_matches_reference(golden_dir):
    gold = np.load(os.path.join(golden_dir, "finetune_remap.npz"))
    num_layers, decay = 12, 0.65
    assigner = optim_factory.LayerDecayValueAssigner([decay ** (num_layers + 1 - i) for i in range(num_layers + 2)])
    names = [str(n) for n in gold["layer_decay/names"]]
    assert [assigner.get_layer_id(n) for n in names] == gold["layer_decay/ids"].tolist()
    assert np.array_equal(np.array([assigner.get_scale(assigner.get_layer_id(n)) for n in names]), gold["layer_decay/scales"])
    # lr_scale reaches the parameter groups (run_class_finetuning.py:527-545)
    ft = registry.create_model("ft_vit", **vit_ref.TINY_FT)
    small = optim_factory.LayerDecayValueAssigner([decay ** (2 + 1 - i) for i in range(2 + 2)])
    with contextlib.redirect_stdout(io.StringIO()):
        groups = optim_factory.get_parameter_groups(ft, 0.05, ft.no_weight_decay(), small.get_layer_id, small.get_scale)
    scales = sorted({g["lr_scale"] for g in groups})
    assert scales == sorted(small.values)
