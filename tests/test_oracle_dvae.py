"""CPU: pin oracle/dvae_ref.py to the golden outputs of the unmodified reference DiscreteVAE
(tests/golden/dvae_tiny.npz) and check the mem_b200 container's state_dict layout."""
import os

import numpy as np
import pytest
import torch

from mem_b200.vae_model import DiscreteVAE
from oracle import dvae_ref

CASES = (("a", dvae_ref.TINY_A, 3, 21, 1.0), ("b", dvae_ref.TINY_B, 2, 22, 4.0), ("c", dvae_ref.TINY_C, 5, 23, 1.0))


@pytest.mark.parametrize("name,cfg,B,seed,gain", CASES)
def test_dvae_oracle_matches_reference_golden(golden_dir, name, cfg, B, seed, gain):
    gold = np.load(os.path.join(golden_dir, "dvae_tiny.npz"))
    vae = DiscreteVAE(**cfg)           # container only: supplies keys / shapes
    sd = dvae_ref.synth_state_dict(vae.state_dict(), seed, gain)
    img = dvae_ref.synth_images(B, cfg["channels"], cfg["input_H"], cfg["input_W"], seed + 100)
    with torch.no_grad():
        logits = dvae_ref.encoder_logits(img, sd, cfg["num_layers"], cfg["num_resnet_blocks"], cfg.get("normalization"))
        idx = dvae_ref.codebook_indices(img, sd, cfg["num_layers"], cfg["num_resnet_blocks"], cfg.get("normalization"))
    np.testing.assert_allclose(logits.numpy(), gold[f"{name}/logits"], rtol=0, atol=2e-6)
    assert np.array_equal(idx.numpy(), gold[f"{name}/indices"])
    assert idx.dtype == torch.int64 and idx.shape == (B, (cfg["input_H"] >> cfg["num_layers"]) * (cfg["input_W"] >> cfg["num_layers"]))


def test_container_attributes():
    v = DiscreteVAE(input_H=224, input_W=224, num_tokens=8192, codebook_dim=32, num_layers=4, num_resnet_blocks=3,
                    hidden_dim=384, channels=2)
    assert v.input_size == (224, 224) and v.num_layers == 4 and v.num_tokens == 8192
    sd = v.state_dict()
    assert sd["encoder.0.0.weight"].shape == (384, 2, 4, 4) and sd["encoder.7.weight"].shape == (8192, 384, 1, 1)
    assert sd["encoder.4.net.2.weight"].shape == (384, 384, 3, 3) and sd["codebook.weight"].shape == (8192, 32)
    with pytest.raises(RuntimeError):
        v.get_codebook_indices(torch.zeros(1, 2, 224, 224))          # CPU tensor: no fallback
    with pytest.raises(RuntimeError):
        v(torch.zeros(1, 2, 224, 224), return_loss=True)             # the training path is CUDA-only as well
    with pytest.raises(RuntimeError):
        v.decode(torch.zeros(1, 196, dtype=torch.long))


@pytest.mark.reference
@pytest.mark.parametrize("cfg", [dvae_ref.TINY_A, dvae_ref.TINY_B, dvae_ref.TINY_C])
def test_same_seed_same_random_init_as_reference(cfg):
    from oracle import ref_shims
    vm = ref_shims.ref_module("vae.vae_model")
    torch.manual_seed(7)
    ref = vm.DiscreteVAE(**cfg)
    torch.manual_seed(7)
    ours = DiscreteVAE(**cfg)
    a, b = ref.state_dict(), ours.state_dict()
    assert list(a.keys()) == list(b.keys())
    for k in a:
        assert torch.equal(a[k], b[k]), k


def test_training_path_oracle_matches_reference_golden(golden_dir):
    """oracle/dvae_ref.py train_loss / decode (restating vae_model.py:160-213) vs the UNMODIFIED reference's loss,
    reconstruction, gradients and decode() stored in tests/golden/dvae_train.npz."""
    import os
    gold = np.load(os.path.join(golden_dir, "dvae_train.npz"))
    from mem_b200.vae_model import DiscreteVAE
    for name, cfg, B, seed, temp in (("a", dvae_ref.TRAIN_A, 3, 61, 0.8), ("b", dvae_ref.TRAIN_B, 2, 62, None),
                                     ("c", dvae_ref.TRAIN_C, 4, 63, 1.0)):
        torch.manual_seed(0)
        sd = dvae_ref.synth_train_state_dict(DiscreteVAE(**cfg).state_dict(), seed)
        sd = {k: v.requires_grad_(True) for k, v in sd.items()}
        img = dvae_ref.synth_images(B, cfg["channels"], cfg["input_H"], cfg["input_W"], seed + 100)
        loss, recons = dvae_ref.train_loss(img, sd, cfg, torch.from_numpy(gold[f"{name}/noise"]), temp)
        loss.backward()
        assert abs(loss.item() - float(gold[f"{name}/loss"][0])) < 1e-5 * max(1.0, abs(loss.item()))
        np.testing.assert_allclose(recons.detach().numpy(), gold[f"{name}/recons"], rtol=1e-4, atol=1e-5)
        for k in gold.files:
            if k.startswith(f"{name}/grad/"):
                np.testing.assert_allclose(sd[k[len(name) + 6:]].grad.numpy(), gold[k], rtol=2e-4, atol=2e-6, err_msg=k)
        with torch.no_grad():
            dec = dvae_ref.decode(torch.from_numpy(gold[f"{name}/seq"]), sd, cfg)
        np.testing.assert_allclose(dec.numpy(), gold[f"{name}/decode"], rtol=1e-4, atol=1e-5)
        # the seeded sample is what F.gumbel_softmax draws
        h, w = cfg["input_H"] >> cfg["num_layers"], cfg["input_W"] >> cfg["num_layers"]
        assert torch.equal(dvae_ref.gumbel_noise((B, cfg["num_tokens"], h, w), 1000 + seed), torch.from_numpy(gold[f"{name}/noise"]))
