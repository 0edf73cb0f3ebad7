"""CPU: pin oracle/vit_ref.py (fp32 restatement) to the golden outputs of the unmodified reference
classes (tests/golden/vit_tiny.npz, made by oracle/make_golden.py), and check that the mem_b200 model
containers expose the reference's state_dict keys / shapes."""
import os

import numpy as np
import pytest
import torch

from mem_b200 import modeling_finetune, modeling_pretrain, registry  # noqa: F401
from oracle import vit_ref


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "vit_tiny.npz"))


def _check_grads(gold, prefix, grads, rtol=2e-4, atol=2e-6):
    seen = 0
    for key in gold.files:
        if not key.startswith(prefix + "grad_norm/"):
            continue
        name = key[len(prefix + "grad_norm/"):]
        g = grads[name].detach().reshape(-1).double().numpy()
        want_norm, want_sum = gold[key]
        assert abs(np.sqrt((g ** 2).sum()) - want_norm) <= rtol * want_norm + atol, name
        full, head = prefix + "grad_full/" + name, prefix + "grad_head/" + name
        if full in gold.files:
            np.testing.assert_allclose(g, gold[full], rtol=rtol, atol=atol, err_msg=name)
        else:
            np.testing.assert_allclose(g[:256], gold[head], rtol=rtol, atol=atol, err_msg=name)
        seen += 1
    assert seen == len(grads)


def test_pt_vit_oracle_matches_reference_golden(gold):
    model = registry.create_model("pt_vit", **vit_ref.TINY)
    sd = vit_ref.synth_state_dict(model.state_dict(), seed=11)
    sd = {k: (v.requires_grad_(True) if v.is_floating_point() else v) for k, v in sd.items()}
    img, mask, tokens = vit_ref.synth_inputs(3, 2, 112, 112, 49, 512, seed=5, n_mask=20)
    loss, acc, logits = vit_ref.mem_loss(img, mask, tokens, sd, heads=2, patch=16)
    np.testing.assert_allclose(logits.detach().numpy(), gold["pt/logits"], rtol=1e-4, atol=1e-5)
    assert abs(loss.item() - float(gold["pt/loss"])) < 1e-5
    assert abs(acc.item() - float(gold["pt/acc"])) < 1e-7
    loss.backward()
    names = [n for n, _ in model.named_parameters()]
    _check_grads(gold, "pt/", {n: sd[n].grad for n in names})
    with torch.no_grad():
        allt = vit_ref.masked_logits(img, mask, sd, 2, 16, return_all_tokens=True)[0]
    np.testing.assert_allclose(allt.numpy(), gold["pt/all_tokens_logits_b0"], rtol=1e-4, atol=1e-5)


def test_ft_vit_oracle_matches_reference_golden(gold):
    model = registry.create_model("ft_vit", **vit_ref.TINY_FT)
    sd = vit_ref.synth_state_dict(model.state_dict(), seed=12)
    sd = {k: (v.requires_grad_(True) if v.is_floating_point() else v) for k, v in sd.items()}
    img, _, _ = vit_ref.synth_inputs(4, 3, 112, 112, 49, 2, seed=6, n_mask=1)
    logits = vit_ref.classify_logits(img, sd, heads=2, patch=16)
    np.testing.assert_allclose(logits.detach().numpy(), gold["ft/logits"], rtol=1e-4, atol=1e-6)
    loss = torch.nn.functional.cross_entropy(logits, torch.tensor([0, 1, 1, 0]))
    assert abs(loss.item() - float(gold["ft/loss"])) < 1e-6
    loss.backward()
    _check_grads(gold, "ft/", {n: sd[n].grad for n, _ in model.named_parameters()})


def test_ft_vit_cls_token_head_oracle_matches_reference_golden(golden_dir):
    """use_mean_pooling=False (modeling_finetune.py:286-287, :349-352): final norm, cls token, head."""
    import os
    gold = np.load(os.path.join(golden_dir, "finetune_remap.npz"))
    model = registry.create_model("ft_vit", **dict(vit_ref.TINY_FT, use_mean_pooling=False))
    assert model.fc_norm is None and isinstance(model.norm, torch.nn.LayerNorm)
    sd = vit_ref.synth_state_dict(model.state_dict(), seed=53)
    sd = {k: (v.requires_grad_(True) if v.is_floating_point() else v) for k, v in sd.items()}
    img, _, _ = vit_ref.synth_inputs(4, 3, 112, 112, 49, 2, seed=16, n_mask=1)
    logits = vit_ref.classify_logits(img, sd, heads=2, patch=16)
    np.testing.assert_allclose(logits.detach().numpy(), gold["cls/logits"], rtol=1e-4, atol=1e-6)
    loss = torch.nn.functional.cross_entropy(logits, torch.tensor([1, 0, 1, 1]))
    assert abs(loss.item() - float(gold["cls/loss"])) < 1e-6
    loss.backward()
    _check_grads(gold, "cls/", {n: sd[n].grad for n, _ in model.named_parameters()})


def test_registered_names_and_state_dict_layout():
    assert {"pt_vit", "ft_vit", "beit_base_patch16_224_8k_vocab", "beit_large_patch16_224_8k_vocab"} <= set(registry.list_models())
    m = registry.create_model("beit_base_patch16_224_8k_vocab", pretrained=False, drop_path_rate=0.1, drop_block_rate=None,
                              use_shared_rel_pos_bias=True, use_abs_pos_emb=False, init_values=0.1, in_chans=2)
    sd = m.state_dict()
    n_params = sum(p.numel() for p in m.parameters())
    assert n_params == 91_769_168 and len(list(m.parameters())) == 189
    assert sd["rel_pos_bias.relative_position_bias_table"].shape == (732, 12)
    assert sd["blocks.11.attn.qkv.weight"].shape == (2304, 768) and "blocks.0.attn.q_bias" in sd
    assert sd["lm_head.weight"].shape == (8192, 768) and sd["patch_embed.proj.weight"].shape == (768, 2, 16, 16)
    assert m.get_num_layers() == 12 and m.no_weight_decay() == {"pos_embed", "cls_token"}
    assert m.patch_embed.patch_shape == (14, 14) and m.patch_embed.patch_size == (16, 16)
    assert m.blocks[0].drop_path.drop_prob == 0.0 and abs(m.blocks[11].drop_path.drop_prob - 0.1) < 1e-7


@pytest.mark.reference
def test_state_dict_keys_equal_reference():
    from oracle import ref_shims
    ref = ref_shims.ref_create_model("pt_vit", **vit_ref.TINY)
    ours = registry.create_model("pt_vit", **vit_ref.TINY)
    a, b = ref.state_dict(), ours.state_dict()
    assert list(a.keys()) == list(b.keys())
    for k in a:
        assert a[k].shape == b[k].shape, k
        if not a[k].is_floating_point():
            assert torch.equal(a[k], b[k]), k
    reff = ref_shims.ref_create_model("ft_vit", **vit_ref.TINY_FT)
    oursf = registry.create_model("ft_vit", **vit_ref.TINY_FT)
    a, b = reff.state_dict(), oursf.state_dict()
    assert list(a.keys()) == list(b.keys())
    for k in a:
        assert a[k].shape == b[k].shape, k
        if not a[k].is_floating_point():
            assert torch.equal(a[k], b[k]), k


def test_cpu_call_fails_loudly():
    m = registry.create_model("pt_vit", **vit_ref.TINY)
    img, mask, _ = vit_ref.synth_inputs(1, 2, 112, 112, 49, 512, seed=1, n_mask=5)
    with pytest.raises(RuntimeError):
        m(img, mask)
