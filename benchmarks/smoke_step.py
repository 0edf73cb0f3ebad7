"""smoke(): one tiny MEM pretraining step on cuda:0 (dVAE tokens -> masked ViT -> CE -> backward -> clip + AdamW
through engine_for_pretraining.train_one_epoch), checked against the fp32 oracle (oracle/engine_ref.py)."""
from __future__ import annotations

import contextlib
import io
from types import SimpleNamespace


def run():
    import torch
    from mem_b200 import _lib, engine_for_pretraining, optim_factory, registry, utils
    from mem_b200 import modeling_pretrain  # noqa: F401
    from mem_b200.vae_model import DiscreteVAE
    from oracle import dvae_ref, engine_ref, vit_ref
    model = registry.create_model("pt_vit", **vit_ref.TINY)
    vit_sd = vit_ref.synth_state_dict(model.state_dict(), seed=31)
    model.load_state_dict(vit_sd)
    vae = DiscreteVAE(**engine_ref.TINY_VAE)
    vae_sd = dvae_ref.synth_state_dict(vae.state_dict(), seed=32, head_gain=4.0)
    vae.load_state_dict(vae_sd)
    model.cuda(), vae.cuda()
    batches = engine_ref.synth_batches(steps=1)
    l0 = _lib.launch_count()
    with contextlib.redirect_stdout(io.StringIO()):
        opt = optim_factory.create_optimizer(SimpleNamespace(opt="adamw", weight_decay=engine_ref.WD[0], lr=engine_ref.LR[0], opt_eps=1e-8), model)
        stats = engine_for_pretraining.train_one_epoch(model, vae, [(batches[0], None)], opt, torch.device("cuda"), 0,
                                                       utils.NativeScalerWithGradNormCount(), engine_ref.MAX_NORM, start_steps=0,
                                                       lr_schedule_values=engine_ref.LR, wd_schedule_values=engine_ref.WD)
    want, _ = engine_ref.run_steps(vit_sd, vae_sd, batches)
    tok = vae.get_codebook_indices(batches[0][1].cuda()).cpu()
    tok_ref = dvae_ref.codebook_indices(batches[0][1], vae_sd, 4, 1)
    n_diff = int((tok != tok_ref).sum())
    assert n_diff <= 1, f"dVAE tokens differ from the oracle at {n_diff} positions"
    assert abs(stats["loss"] - want[0]["loss"]) < 2e-2 * want[0]["loss"], (stats, want[0])
    assert abs(stats["grad_norm"] - want[0]["grad_norm"]) < 6e-2 * want[0]["grad_norm"], (stats, want[0])
    print(f"smoke: MEM step loss {stats['loss']:.4f} (oracle {want[0]['loss']:.4f}), grad_norm {stats['grad_norm']:.3f} "
          f"(oracle {want[0]['grad_norm']:.3f}), dVAE token diffs {n_diff}/{tok.numel()}; libmemb launches {_lib.launch_count() - l0}")
