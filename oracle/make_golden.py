"""Oracle tooling (test infrastructure, not product): freeze golden vectors.

Runs the UNMODIFIED reference (``/root/reference`` through ``oracle/ref_shims``)
on small seeded inputs and stores inputs + outputs under ``tests/golden/``.
The reference has no tests or fixtures of its own, so these vectors are what
pins ``oracle/*_ref.py`` (and through it the CUDA path) to the reference.

    python -m oracle.make_golden            # in the build container only

The fixtures are committed; the GPU box never runs this script.
"""
from __future__ import annotations

import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from oracle import ref_shims  # noqa: E402


def synth_events(rng, n, H, W, kind="uniform", polarity=(-1.0, 1.0), frac=False):
    """Synthetic (n,4) float64 stream of SURVEY.md section 8(d) shapes."""
    if kind == "uniform":
        x = rng.integers(0, W, n).astype(np.float64)
        y = rng.integers(0, H, n).astype(np.float64)
    elif kind == "hot":
        x = rng.integers(0, W, n).astype(np.float64)
        y = rng.integers(0, H, n).astype(np.float64)
        hot = rng.random(n) < 0.5
        hx, hy = rng.integers(0, W, 4), rng.integers(0, H, 4)
        pick = rng.integers(0, 4, n)
        x[hot], y[hot] = hx[pick[hot]], hy[pick[hot]]
    elif kind == "edge":
        seg = rng.integers(0, 8, n)
        x0, y0 = rng.uniform(0, W, 8), rng.uniform(0, H, 8)
        dx, dy = rng.uniform(-1, 1, 8), rng.uniform(-1, 1, 8)
        s = rng.uniform(0, min(H, W) / 2, n)
        x = np.clip(x0[seg] + dx[seg] * s + rng.normal(0, 1, n), 0, W - 1e-3)
        y = np.clip(y0[seg] + dy[seg] * s + rng.normal(0, 1, n), 0, H - 1e-3)
    else:
        raise ValueError(kind)
    if frac:
        x = np.clip(x + rng.uniform(-0.9, 0.9, n), -0.99, W - 1e-3)
        y = np.clip(y + rng.uniform(0.0, 0.9, n), 0, H - 1e-3)
    t = np.sort(rng.uniform(0, 3e5, n))
    p = rng.choice(np.asarray(polarity, dtype=np.float64), n)
    return np.stack([x, y, t, p], axis=1)


def golden_histogram():
    ds = ref_shims.ref_module("datasets")
    rng = np.random.default_rng(20240117)
    cases = {}

    def add(name, ev, H, W, tss):
        img = ds.EventArrToImg(H, W, tss)(ev.copy())
        cases[name + "_ev"] = ev
        cases[name + "_img"] = np.ascontiguousarray(img)
        cases[name + "_cfg"] = np.array([-1 if H is None else H, -1 if W is None else W, int(tss)])

    add("uniform_180x240", synth_events(rng, 6000, 180, 240), 180, 240, False)
    add("wrap_100x100", synth_events(rng, 40000, 100, 100, "hot"), 100, 100, False)       # >255 per pixel
    add("edge_frac_120x160", synth_events(rng, 5000, 120, 160, "edge", frac=True), 120, 160, False)
    add("ncars_polarity01", synth_events(rng, 3000, 100, 120, polarity=(0.0, 1.0)), 100, 120, False)
    add("infer_hw", synth_events(rng, 2000, 130, 170), None, None, False)
    add("tss_180x240", synth_events(rng, 4000, 180, 240), 180, 240, True)
    add("tss_infer_hot", synth_events(rng, 3000, 100, 110, "hot"), None, None, True)
    neg = synth_events(rng, 500, 100, 100)
    neg[::7, 1] = 0.0
    neg[::7, 0] = -neg[::7, 0] - 1.0           # negative flat index -> numpy wraps once
    add("negative_wrap", neg, 100, 100, False)
    single = np.array([[5.0, 7.0, 10.0, 1.0]])
    add("single_event", single, 100, 100, False)
    add("empty_fixed", np.zeros((0, 4)), 100, 100, False)
    np.savez_compressed(os.path.join(GOLD, "histogram.npz"), **cases)
    print("histogram.npz:", len(cases) // 3, "cases")


def golden_masks():
    mg = ref_shims.ref_module("masking_generator")
    out = {}
    cfgs = [((14, 14), 75, 16), ((14, 14), 98, 16), ((14, 14), 75, 4), ((16, 16), 120, 8), ((7, 9), 20, 2)]
    for ci, (hw, nmask, mn) in enumerate(cfgs):
        masks = []
        for seed in range(32):
            random.seed(seed)
            masks.append(mg.MaskingGenerator(hw, nmask, min_num_patches=mn)())
        out[f"block_{ci}_cfg"] = np.array([hw[0], hw[1], nmask, mn])
        out[f"block_{ci}_masks"] = np.stack(masks).astype(np.uint8)
    # a stream of consecutive draws from one seed (what a DataLoader worker does)
    random.seed(1234)
    gen = mg.MaskingGenerator((14, 14), 75, min_num_patches=16)
    out["stream_masks"] = np.stack([gen() for _ in range(64)]).astype(np.uint8)
    masks = []
    for seed in range(16):
        random.seed(seed)
        masks.append(mg.MaskingGeneratorRandomLocation((14, 14), 75)())
    out["randloc_masks"] = np.stack(masks).astype(np.uint8)
    np.savez_compressed(os.path.join(GOLD, "masks.npz"), **out)
    print("masks.npz written")


def _grad_digest(named_grads):
    """Full gradient for small tensors, (norm, sum, first 256 values) for large ones."""
    out = {}
    for k, g in named_grads.items():
        g = g.detach().float().reshape(-1).numpy()
        if g.size <= 4096:
            out["grad_full/" + k] = g
        else:
            out["grad_head/" + k] = g[:256].copy()
        out["grad_norm/" + k] = np.array([np.sqrt((g.astype(np.float64) ** 2).sum()), g.astype(np.float64).sum()])
    return out


def golden_vit():
    """Loss / logits / gradients of the UNMODIFIED reference pt_vit and ft_vit on the seeded tiny cases."""
    import torch
    from oracle import vit_ref
    torch.manual_seed(0)
    out = {}
    # ---- pretraining model (mem/modeling_pretrain.py) through the reference's own CrossEntropyLoss step
    model = ref_shims.ref_create_model("pt_vit", **vit_ref.TINY)
    sd = vit_ref.synth_state_dict(model.state_dict(), seed=11)
    model.load_state_dict(sd)
    model.train()
    P = 49
    img, mask, tokens = vit_ref.synth_inputs(3, 2, 112, 112, P, vit_ref.TINY["vocab_size"], seed=5, n_mask=20)
    logits = model(img, bool_masked_pos=mask, return_all_tokens=False)
    labels = tokens[mask]
    loss = torch.nn.CrossEntropyLoss()(input=logits, target=labels)
    loss.backward()
    out["pt/loss"] = np.array(loss.item())
    out["pt/logits"] = logits.detach().numpy()
    out["pt/acc"] = np.array((logits.max(-1)[1] == labels).float().mean().item())
    out.update({"pt/" + k: v for k, v in _grad_digest({n: p.grad for n, p in model.named_parameters()}).items()})
    with torch.no_grad():
        model.eval()
        out["pt/all_tokens_logits_b0"] = model(img, bool_masked_pos=mask, return_all_tokens=True)[0].numpy()
    # ---- finetune model (mem/modeling_finetune.py ft_vit), plain CE on 2 classes
    ft = ref_shims.ref_create_model("ft_vit", **vit_ref.TINY_FT)
    sdf = vit_ref.synth_state_dict(ft.state_dict(), seed=12)
    ft.load_state_dict(sdf)
    ft.train()
    img3, _, _ = vit_ref.synth_inputs(4, 3, 112, 112, P, 2, seed=6, n_mask=1)
    target = torch.tensor([0, 1, 1, 0])
    lg = ft(img3)
    lossf = torch.nn.CrossEntropyLoss()(lg, target)
    lossf.backward()
    out["ft/loss"] = np.array(lossf.item())
    out["ft/logits"] = lg.detach().numpy()
    out.update({"ft/" + k: v for k, v in _grad_digest({n: p.grad for n, p in ft.named_parameters()}).items()})
    np.savez_compressed(os.path.join(GOLD, "vit_tiny.npz"), **out)
    print("vit_tiny.npz:", len(out), "arrays, pt loss", loss.item(), "ft loss", lossf.item())


def _cuda_autocast_policy_on_cpu():
    """Context: bf16 autocast on the CPU with CUDA autocast's fp32 list restored (layer_norm, softmax, cross_entropy run in
    fp32 on CUDA; the CPU autocast leaves them in the incoming dtype).  This is how the reference's own mixed-precision
    step (engine_for_pretraining.py:147, ``torch.cuda.amp.autocast``) can be executed in the build container: the
    reference files are untouched, only torch functions are wrapped for the duration of the context."""
    import contextlib
    import torch
    import torch.nn.functional as F

    @contextlib.contextmanager
    def ctx():
        saved = (F.layer_norm, torch.Tensor.softmax, F.softmax, F.cross_entropy)

        def layer_norm(x, shape, weight=None, bias=None, eps=1e-5):
            return saved[0](x.float(), shape, None if weight is None else weight.float(), None if bias is None else bias.float(), eps)

        def t_softmax(self, *a, **k):
            return saved[1](self.float(), *a, **k)

        def f_softmax(x, *a, **k):
            return saved[2](x.float(), *a, **k)

        def cross_entropy(x, *a, **k):
            return saved[3](x.float(), *a, **k)
        F.layer_norm, torch.Tensor.softmax, F.softmax, F.cross_entropy = layer_norm, t_softmax, f_softmax, cross_entropy
        try:
            with torch.autocast("cpu", dtype=torch.bfloat16):
                yield
        finally:
            F.layer_norm, torch.Tensor.softmax, F.softmax, F.cross_entropy = saved
    return ctx()


def golden_vit_bf16():
    """Calibration of the ViT parity tolerances (north_star: "within the tolerance of a bf16 run of the reference vs its
    fp32 run"): the UNMODIFIED reference models are run twice on the exact weights / inputs of the GPU parity tests,
    once in fp32 and once under bf16 autocast, and the error of the loss, the logits and EVERY parameter gradient is
    stored per tensor (``<case>/err/<name>`` = L2 norm of the difference, ``<case>/ref/<name>`` = L2 norm of the fp32
    tensor).  tests/test_vit_model_gpu.py allows a stated multiple of these."""
    import torch
    from oracle import vit_ref
    out = {}

    def run(case, model, fwd_loss):
        model.train()
        res = {}
        for mode in ("fp32", "bf16"):
            for p in model.parameters():
                p.grad = None
            if mode == "fp32":
                loss, logits = fwd_loss(model)
            else:
                with _cuda_autocast_policy_on_cpu():
                    loss, logits = fwd_loss(model)
            loss.backward()
            res[mode] = (loss.item(), logits.detach().float().clone(), {n: p.grad.detach().clone() for n, p in model.named_parameters()})
        l32, lg32, g32 = res["fp32"]
        l16, lg16, g16 = res["bf16"]
        out[f"{case}/loss"] = np.array([l32, l16])
        out[f"{case}/err/logits"] = np.array((lg16 - lg32).double().norm().item())
        out[f"{case}/ref/logits"] = np.array(lg32.double().norm().item())
        worst = (None, 0.0)
        for n in g32:
            e, r = (g16[n] - g32[n]).double().norm().item(), g32[n].double().norm().item()
            out[f"{case}/err/{n}"] = np.array(e)
            out[f"{case}/ref/{n}"] = np.array(r)
            if r > 0 and e / r > worst[1]:
                worst = (n, e / r)
        rels = sorted(out[f"{case}/err/{n}"] / max(out[f"{case}/ref/{n}"], 1e-30) for n in g32)
        print(f"{case}: loss {l32:.5f} -> {l16:.5f} ({abs(l16 - l32) / l32:.2e}), logits rel "
              f"{out[f'{case}/err/logits'] / out[f'{case}/ref/logits']:.2e}, grad rel median {rels[len(rels) // 2]:.2e} "
              f"max {worst[1]:.2e} ({worst[0]})")

    def mem_loss(img, mask, tokens):
        def f(model):
            logits = model(img, bool_masked_pos=mask, return_all_tokens=False)
            return torch.nn.CrossEntropyLoss()(input=logits, target=tokens[mask]), logits
        return f

    def cls_loss(img, target):
        def f(model):
            logits = model(img)
            return torch.nn.CrossEntropyLoss()(logits, target), logits
        return f

    def scaled(sd, factor, skip=("relative_position",), gamma=None):
        for k in sd:
            if sd[k].is_floating_point() and sd[k].dim() >= 2 and not any(s in k for s in skip):
                sd[k] = sd[k] * factor
            if gamma is not None and "gamma_" in k:
                sd[k] = sd[k] * 0.0 + gamma
        return sd

    torch.manual_seed(0)
    # tiny pt_vit / ft_vit: test_tiny_pt_vit_matches_golden_and_oracle, test_tiny_ft_vit_matches_golden_and_oracle
    m = ref_shims.ref_create_model("pt_vit", **vit_ref.TINY)
    m.load_state_dict(vit_ref.synth_state_dict(m.state_dict(), seed=11))
    run("tiny_pt", m, mem_loss(*vit_ref.synth_inputs(3, 2, 112, 112, 49, vit_ref.TINY["vocab_size"], seed=5, n_mask=20)))
    m = ref_shims.ref_create_model("ft_vit", **vit_ref.TINY_FT)
    m.load_state_dict(vit_ref.synth_state_dict(m.state_dict(), seed=12))
    run("tiny_ft", m, cls_loss(vit_ref.synth_inputs(4, 3, 112, 112, 49, 2, seed=6, n_mask=1)[0], torch.tensor([0, 1, 1, 0])))
    # cls-token head (use_mean_pooling=False): test_tiny_ft_vit_cls_token_head
    m = ref_shims.ref_create_model("ft_vit", **dict(vit_ref.TINY_FT, use_mean_pooling=False))
    m.load_state_dict(vit_ref.synth_state_dict(m.state_dict(), seed=53))
    run("tiny_ft_cls", m, cls_loss(vit_ref.synth_inputs(4, 3, 112, 112, 49, 2, seed=16, n_mask=1)[0], torch.tensor([1, 0, 1, 1])))
    # ViT-B/16 at batch 4: test_vit_base_step_vs_oracle_with_droppath (DropPath off here: it is a per-sample mask)
    base = dict(img_size=(224, 224), patch_size=(16, 16), in_chans=2, vocab_size=8192, embed_dim=768, depth=12, num_heads=12,
                mlp_ratio=4, init_values=0.1, use_shared_rel_pos_bias=True, use_abs_pos_emb=False, drop_path_rate=0.0)
    m = ref_shims.ref_create_model("pt_vit", **base)
    m.load_state_dict(scaled(vit_ref.synth_state_dict(m.state_dict(), seed=3), 0.4))
    run("base_pt_b4", m, mem_loss(*vit_ref.synth_inputs(4, 2, 224, 224, 196, 8192, seed=7, n_mask=75)))
    # ViT-B/16 ft_vit at batch 4: test_ft_vit_base_forward_backward_vs_oracle
    ftb = dict(img_size=(224, 224), patch_size=(16, 16), in_chans=3, num_classes=2, embed_dim=768, depth=12, num_heads=12,
               mlp_ratio=4, init_values=0.1, use_rel_pos_bias=True, use_abs_pos_emb=False, use_mean_pooling=True, drop_path_rate=0.0)
    m = ref_shims.ref_create_model("ft_vit", **ftb)
    m.load_state_dict(scaled(vit_ref.synth_state_dict(m.state_dict(), seed=4), 0.4, skip=("relative_position", "head")))
    run("base_ft_b4", m, cls_loss(vit_ref.synth_inputs(4, 3, 224, 224, 196, 2, seed=8, n_mask=1)[0], torch.tensor([0, 1, 1, 0])))
    # ViT-L/16 at batch 2: test_vit_large_step_vs_oracle
    large = dict(base, embed_dim=1024, depth=24, num_heads=16, init_values=1e-5)
    m = ref_shims.ref_create_model("pt_vit", **large)
    m.load_state_dict(scaled(vit_ref.synth_state_dict(m.state_dict(), seed=5), 0.3, gamma=0.05))
    run("large_pt_b2", m, mem_loss(*vit_ref.synth_inputs(2, 2, 224, 224, 196, 8192, seed=9, n_mask=75)))
    np.savez_compressed(os.path.join(GOLD, "vit_bf16_calibration.npz"), **out)
    print("vit_bf16_calibration.npz:", len(out), "arrays")


def golden_dvae():
    """Logits / indices of the UNMODIFIED reference DiscreteVAE (eventvae/vae/vae_model.py) on seeded tiny cases."""
    import torch
    from oracle import dvae_ref
    vm = ref_shims.ref_module("vae.vae_model")
    out = {}
    for name, cfg, B, seed, gain in (("a", dvae_ref.TINY_A, 3, 21, 1.0), ("b", dvae_ref.TINY_B, 2, 22, 4.0),
                                     ("c", dvae_ref.TINY_C, 5, 23, 1.0)):
        torch.manual_seed(0)
        vae = vm.DiscreteVAE(**cfg)
        vae.load_state_dict(dvae_ref.synth_state_dict(vae.state_dict(), seed, gain))
        vae.train()   # get_codebook_indices must switch to eval and back (eval_decorator)
        img = dvae_ref.synth_images(B, cfg["channels"], cfg["input_H"], cfg["input_W"], seed + 100)
        idx = vae.get_codebook_indices(img)
        assert vae.training
        with torch.no_grad():
            logits = vae(img, return_logits=True)
        out[f"{name}/indices"] = idx.numpy()
        out[f"{name}/logits"] = logits.numpy()
        print(name, "indices", tuple(idx.shape), "distinct tokens", idx.unique().numel())
    np.savez_compressed(os.path.join(GOLD, "dvae_tiny.npz"), **out)


def golden_engine():
    """Three steps of the UNMODIFIED reference train_one_epoch (mem/engine_for_pretraining.py:108-287) on CPU fp32:
    tiny pt_vit + tiny dVAE, AdamW with the reference's parameter groups, per-step lr / wd schedule, clip 1.0."""
    import torch
    from oracle import dvae_ref, engine_ref, vit_ref
    eng = ref_shims.ref_module("engine_for_pretraining")
    rutils = ref_shims.ref_module("utils")
    optf = ref_shims.ref_module("optim_factory")
    vm = ref_shims.ref_module("vae.vae_model")
    torch.cuda.synchronize = lambda *a, **k: None       # CPU run of a CUDA-only loop
    rutils.is_main_process = lambda: False               # skips the wandb / make_grid branch

    class Scaler(rutils.NativeScalerWithGradNormCount):  # GradScaler disables itself on CPU -> state_dict() == {}
        def state_dict(self):
            return {"scale": 1.0}

    torch.manual_seed(0)
    model = ref_shims.ref_create_model("pt_vit", **vit_ref.TINY)
    model.load_state_dict(vit_ref.synth_state_dict(model.state_dict(), seed=31))
    vae = vm.DiscreteVAE(**engine_ref.TINY_VAE)
    vae.load_state_dict(dvae_ref.synth_state_dict(vae.state_dict(), seed=32, head_gain=4.0))
    groups = optf.get_parameter_groups(model, engine_ref.WD[0], model.no_weight_decay())
    opt = torch.optim.AdamW(groups, lr=engine_ref.LR[0], betas=(0.9, 0.95), eps=1e-8)
    scaler = Scaler()
    out = {}
    for it, batch in enumerate(engine_ref.synth_batches()):
        stats = eng.train_one_epoch(model, vae, [(batch, None)], opt, torch.device("cpu"), 0, scaler, engine_ref.MAX_NORM,
                                    start_steps=it, lr_schedule_values=engine_ref.LR, wd_schedule_values=engine_ref.WD)
        for k, v in stats.items():
            out[f"step{it}/{k}"] = np.array(v)
        print(it, {k: round(float(v), 6) for k, v in stats.items()})
    out["keys"] = np.array(sorted(stats.keys()))
    sd = model.state_dict()
    for k in ("lm_head.weight", "blocks.0.attn.qkv.weight", "cls_token", "rel_pos_bias.relative_position_bias_table", "blocks.1.gamma_2"):
        out["final/" + k] = sd[k].detach().numpy().reshape(-1)[:512]
    np.savez_compressed(os.path.join(GOLD, "engine_tiny.npz"), **out)


def golden_event_pipeline():
    """Reference build_transformNPY (mem/datasets.py:611-660) on seeded synthetic N-ImageNet-shaped streams.

    Inputs are regenerated from the stored seeds (``synth_events``), outputs are stored: the (3,h,w) float32 tensor
    the reference chain returns, plus the RNG seeds it ran under."""
    import torch
    from types import SimpleNamespace
    ds = ref_shims.ref_module("datasets")
    out = {}
    cases = [  # name, is_train, n_events, kind, normalize, seed
        ("train_a", True, 45000, "edge", 1, 11), ("train_b", True, 45000, "hot", 1, 12), ("train_c", True, 20000, "uniform", 0, 13),
        ("train_d", True, 45000, "edge", 0, 14), ("eval_a", False, 45000, "edge", 1, 15), ("eval_b", False, 12000, "hot", 0, 16),
    ]
    for name, is_train, n, kind, norm, seed in cases:
        args = SimpleNamespace(data_path="/data/N_imagenet", input_H=224, input_W=224, slice_max_evs=30000,
                               max_random_shift_evs=15, timesurface=0, hotpixfilter=1, hotpix_num_stds=10, logtrafo=0,
                               gammatrafo=0, gamma=0.5, normalize_events=norm, rand_aug=0)
        import contextlib, io
        with contextlib.redirect_stdout(io.StringIO()):
            tf = ds.build_transformNPY(is_train, args)
        ev = synth_events(np.random.default_rng(seed), n, 480, 640, kind, frac=(kind == "edge"))
        random.seed(seed); np.random.seed(seed); torch.manual_seed(seed)
        res = tf(ev.copy())
        out[name + "_out"] = res.numpy()
        out[name + "_meta"] = np.array([int(is_train), n, norm, seed], dtype=np.int64)
        out[name + "_kind"] = np.array(kind)
        print(f"event_pipeline {name}: out {tuple(res.shape)} nnz {int((res != 0).sum())} max {float(res.max()):.4f}")
    np.savez_compressed(os.path.join(GOLD, "event_pipeline.npz"), **out)



def golden_decode():
    """Reference process_data/process_dataset.py decoders (ncaltech101 :24-63, ncars :66-103) run on synthetic
    recordings written to a temporary dataset tree; stores the raw file bytes and the .npy arrays they produced."""
    import importlib.util
    import tempfile
    import contextlib, io
    from types import SimpleNamespace
    from oracle import decode_ref
    ref_shims.install()
    pd_dir = os.path.join(ref_shims.REFERENCE_ROOT, "process_data")
    saved = sys.modules.pop("utils", None)          # process_dataset.py does `from utils import *` (its own utils.py)
    sys.path.insert(0, pd_dir)
    try:
        spec = importlib.util.spec_from_file_location("ref_process_dataset", os.path.join(pd_dir, "process_dataset.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        sys.path.remove(pd_dir)
        sys.modules.pop("utils", None)
        if saved is not None:
            sys.modules["utils"] = saved
    rng = np.random.default_rng(21)
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        # N-Caltech101: <input>/<class>/<file>.bin + a split file naming train / val members
        src, dst = os.path.join(tmp, "ncal_in"), os.path.join(tmp, "ncal_out")
        os.makedirs(os.path.join(src, "airplanes"))
        raws = {"image_0001": decode_ref.synth_ncaltech101(rng, 1500), "image_0002": decode_ref.synth_ncaltech101(rng, 1)}
        for k, v in raws.items():
            open(os.path.join(src, "airplanes", k + ".bin"), "wb").write(v)
        split = os.path.join(tmp, "split.txt")
        # (the reference's split parser keeps names up to ".bin\\n" and, its filter iterator being consumed by the
        # "val" pass, only ever finds "val" members -- both recordings are listed as such)
        open(split, "w").write("val/airplanes/image_0001.bin\nval/airplanes/image_0002.bin\n")
        with contextlib.redirect_stdout(io.StringIO()):
            mod.ncaltech101("airplanes", SimpleNamespace(input=src, output=dst, split=split))
        for k, part in (("image_0001", "val"), ("image_0002", "val")):
            out["ncaltech101_" + k + "_raw"] = np.frombuffer(raws[k], dtype=np.uint8)
            out["ncaltech101_" + k + "_npy"] = np.load(os.path.join(dst, part, "airplanes", k + ".npy"))
        # N-Cars: <input>/n-cars_train/<class>/<file>.dat and n-cars_test/...
        src, dst = os.path.join(tmp, "ncars_in"), os.path.join(tmp, "ncars_out")
        raws = {"obj_000001_td": decode_ref.synth_ncars(rng, 1200), "obj_000002_td": decode_ref.synth_ncars(rng, 3)}
        for sub, k in (("n-cars_train", "obj_000001_td"), ("n-cars_test", "obj_000002_td")):
            os.makedirs(os.path.join(src, sub, "cars"))
            open(os.path.join(src, sub, "cars", k + ".dat"), "wb").write(raws[k])
        with contextlib.redirect_stdout(io.StringIO()):
            mod.ncars("cars", SimpleNamespace(input=src, output=dst))
        for k, part in (("obj_000001_td", "train"), ("obj_000002_td", "val")):
            out["ncars_" + k + "_raw"] = np.frombuffer(raws[k], dtype=np.uint8)
            out["ncars_" + k + "_npy"] = np.load(os.path.join(dst, part, "cars", k + ".npy"))
    for k, v in out.items():
        print("decode", k, v.shape, v.dtype)
    np.savez_compressed(os.path.join(GOLD, "decode.npz"), **out)


def golden_engine_ft():
    """The UNMODIFIED reference finetuning loop (mem/engine_for_finetuning.py:42-205) on CPU fp32: tiny ft_vit, four
    micro-batches with update_freq 2 (two AdamW steps, per-step lr / wd schedule, clip 1.0), CrossEntropyLoss."""
    import torch
    from types import SimpleNamespace
    from oracle import engine_ref, vit_ref
    eng = ref_shims.ref_module("engine_for_finetuning")
    rutils = ref_shims.ref_module("utils")
    optf = ref_shims.ref_module("optim_factory")
    torch.cuda.synchronize = lambda *a, **k: None
    rutils.is_main_process = lambda: False

    class Scaler(rutils.NativeScalerWithGradNormCount):
        def state_dict(self):
            return {"scale": 1.0}

    torch.manual_seed(0)
    model = ref_shims.ref_create_model("ft_vit", **vit_ref.TINY_FT)
    model.load_state_dict(vit_ref.synth_state_dict(model.state_dict(), seed=41))
    groups = optf.get_parameter_groups(model, engine_ref.FT_WD[0], model.no_weight_decay())
    opt = torch.optim.AdamW(groups, lr=engine_ref.FT_LR[0], betas=(0.9, 0.95), eps=1e-8)
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        stats = eng.train_one_epoch(SimpleNamespace(), model, torch.nn.CrossEntropyLoss(), engine_ref.synth_class_batches(), opt,
                                    torch.device("cpu"), 0, Scaler(), engine_ref.MAX_NORM, start_steps=0,
                                    lr_schedule_values=engine_ref.FT_LR, wd_schedule_values=engine_ref.FT_WD,
                                    num_training_steps_per_epoch=2, update_freq=engine_ref.FT_UPDATE_FREQ)
    out = {"stat/" + k: np.array(v) for k, v in stats.items()}
    out["keys"] = np.array(sorted(stats.keys()))
    print({k: round(float(v), 6) for k, v in stats.items()})
    sd = model.state_dict()
    for k in ("head.weight", "head.bias", "fc_norm.weight", "blocks.0.attn.qkv.weight", "blocks.1.attn.relative_position_bias_table",
              "cls_token", "patch_embed.proj.bias", "blocks.1.gamma_2", "blocks.0.mlp.fc1.bias"):
        out["final/" + k] = sd[k].detach().numpy().reshape(-1)[:512]
    model.eval()
    with torch.no_grad():
        out["eval_logits"] = model(engine_ref.synth_class_batches()[0][0]).numpy()
    np.savez_compressed(os.path.join(GOLD, "engine_ft_tiny.npz"), **out)


def _sd_digest(sd, full=("relative_position_bias_table", "pos_embed", "head.")):
    out = {}
    for k, v in sd.items():
        a = v.detach().cpu().numpy()
        if any(f in k for f in full) or a.size <= 4096:
            out["full/" + k] = a
        else:
            f = a.reshape(-1).astype(np.float64)
            out["digest/" + k] = np.concatenate([[f.sum(), np.sqrt((f ** 2).sum())], f[:64]])
    return out


def golden_finetune_remap():
    """The UNMODIFIED reference ``utils.finetune`` (mem/utils.py:613-732) turning a pretraining checkpoint into a
    finetuning model, and its layer-decay helper.  Case "same": equal patch grid (shared rel-pos table expanded to one
    per block, lm_head / mask_token dropped, head kept at init).  Case "interp": 7x7 -> 10x10 patch grid with absolute
    position embeddings (geometric-progression bicubic spline for every rel-pos table -- through the interp2d
    replacement of ref_shims, scipy having removed the original -- and F.interpolate for pos_embed).  Also the
    cls-token head (use_mean_pooling=False) forward / gradients of the reference ft_vit."""
    import contextlib
    import io
    import tempfile
    import torch
    from types import SimpleNamespace
    from oracle import vit_ref
    rutils = ref_shims.ref_module("utils")
    optf = ref_shims.ref_module("optim_factory")
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        for case, pt_kw, ft_kw in (
                ("same", dict(vit_ref.TINY, in_chans=3), dict(vit_ref.TINY_FT)),
                ("interp", dict(vit_ref.TINY, in_chans=3, use_abs_pos_emb=True),
                 dict(vit_ref.TINY_FT, img_size=(160, 160), use_abs_pos_emb=True))):
            torch.manual_seed(0)
            pt = ref_shims.ref_create_model("pt_vit", **pt_kw)
            pt.load_state_dict(vit_ref.synth_state_dict(pt.state_dict(), seed=51))
            path = os.path.join(tmp, case + ".pth")
            torch.save({"model": pt.state_dict(), "epoch": 3}, path)
            torch.manual_seed(1)
            ft = ref_shims.ref_create_model("ft_vit", **ft_kw)
            ft.load_state_dict(vit_ref.synth_state_dict(ft.state_dict(), seed=52))
            log = io.StringIO()
            with contextlib.redirect_stdout(log):
                rutils.finetune(SimpleNamespace(finetune=path, model_key="model|module", model_prefix=""), ft)
            out.update({f"{case}/{k}": v for k, v in _sd_digest(ft.state_dict()).items()})
            out[f"{case}/log"] = np.array(log.getvalue())
            print(case, "->", len(ft.state_dict()), "tensors;", log.getvalue().count("Position interpolate"), "interpolations")
    # layer-wise lr decay ids / scales (optim_factory.py:31-53, run_class_finetuning.py:527-529)
    names = ["cls_token", "mask_token", "pos_embed", "patch_embed.proj.weight", "rel_pos_bias.relative_position_bias_table",
             "blocks.0.norm1.weight", "blocks.7.attn.qkv.weight", "blocks.11.mlp.fc2.bias", "fc_norm.weight", "head.weight", "norm.bias"]
    num_layers, decay = 12, 0.65
    assigner = optf.LayerDecayValueAssigner([decay ** (num_layers + 1 - i) for i in range(num_layers + 2)])
    out["layer_decay/names"] = np.array(names)
    out["layer_decay/ids"] = np.array([assigner.get_layer_id(n) for n in names])
    out["layer_decay/scales"] = np.array([assigner.get_scale(assigner.get_layer_id(n)) for n in names])
    # cls-token head of ft_vit (use_mean_pooling=False, modeling_finetune.py:286-287, :349-352)
    torch.manual_seed(0)
    kw = dict(vit_ref.TINY_FT, use_mean_pooling=False)
    m = ref_shims.ref_create_model("ft_vit", **kw)
    m.load_state_dict(vit_ref.synth_state_dict(m.state_dict(), seed=53))
    m.train()
    img, _, _ = vit_ref.synth_inputs(4, 3, 112, 112, 49, 2, seed=16, n_mask=1)
    target = torch.tensor([1, 0, 1, 1])
    logits = m(img)
    loss = torch.nn.CrossEntropyLoss()(logits, target)
    loss.backward()
    out["cls/logits"] = logits.detach().numpy()
    out["cls/loss"] = np.array(loss.item())
    out.update({"cls/" + k: v for k, v in _grad_digest({n: p.grad for n, p in m.named_parameters()}).items()})
    np.savez_compressed(os.path.join(GOLD, "finetune_remap.npz"), **out)
    print("finetune_remap.npz:", len(out), "arrays; cls loss", loss.item())


def golden_dvae_train():
    """The UNMODIFIED reference ``DiscreteVAE`` on its training path (eventvae/vae/vae_model.py:160-213): loss,
    reconstruction and every parameter gradient of ``vae(img, return_loss=True, return_recons=True, temp)`` in fp32 with
    a seeded Gumbel sample, the same under bf16 autocast (the calibration of the GPU test's tolerances, as for the ViT),
    and ``decode(img_seq)``.  The oracle restatement (oracle/dvae_ref.py ``train_loss`` / ``decode``) is checked against
    the reference here as well."""
    import torch
    from oracle import dvae_ref
    vm = ref_shims.ref_module("vae.vae_model")
    out = {}
    for name, cfg, B, seed, temp in (("a", dvae_ref.TRAIN_A, 3, 61, 0.8), ("b", dvae_ref.TRAIN_B, 2, 62, None),
                                     ("c", dvae_ref.TRAIN_C, 4, 63, 1.0)):
        torch.manual_seed(0)
        vae = vm.DiscreteVAE(**cfg)
        sd = dvae_ref.synth_train_state_dict(vae.state_dict(), seed)
        vae.load_state_dict(sd)
        vae.train()
        img = dvae_ref.synth_images(B, cfg["channels"], cfg["input_H"], cfg["input_W"], seed + 100)
        h, w = cfg["input_H"] >> cfg["num_layers"], cfg["input_W"] >> cfg["num_layers"]
        res = {}
        for mode in ("fp32", "bf16"):
            vae.zero_grad()
            torch.manual_seed(1000 + seed)                   # F.gumbel_softmax draws from the global generator
            if mode == "fp32":
                loss, recons = vae(img, return_loss=True, return_recons=True, temp=temp)
            else:
                with _cuda_autocast_policy_on_cpu():
                    loss, recons = vae(img, return_loss=True, return_recons=True, temp=temp)
            loss.backward()
            res[mode] = (loss.item(), recons.detach().float().clone(), {n: p.grad.detach().clone() for n, p in vae.named_parameters()})
        l32, r32, g32 = res["fp32"]
        l16, r16, g16 = res["bf16"]
        noise = dvae_ref.gumbel_noise((B, cfg["num_tokens"], h, w), 1000 + seed)
        # the restatement with the explicit sample reproduces the reference
        sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        lo, ro = dvae_ref.train_loss(img, sdr, cfg, noise, temp)
        lo.backward()
        assert abs(lo.item() - l32) < 1e-6 * max(1.0, abs(l32)) and torch.allclose(ro, r32, atol=1e-5), (name, lo.item(), l32)
        for n in g32:
            assert torch.allclose(sdr[n].grad, g32[n], rtol=1e-4, atol=1e-6), (name, n)
        out[f"{name}/noise"] = noise.numpy()
        out[f"{name}/loss"] = np.array([l32, l16])
        out[f"{name}/recons"] = r32.numpy()
        out[f"{name}/err/recons"] = np.array((r16 - r32).double().norm().item())
        out[f"{name}/ref/recons"] = np.array(r32.double().norm().item())
        for n in g32:
            out[f"{name}/grad/{n}"] = g32[n].numpy()
            out[f"{name}/err/{n}"] = np.array((g16[n] - g32[n]).double().norm().item())
            out[f"{name}/ref/{n}"] = np.array(g32[n].double().norm().item())
        rels = sorted(float(out[f"{name}/err/{n}"]) / max(float(out[f"{name}/ref/{n}"]), 1e-30) for n in g32)
        # decode (vae_model.py:160-171)
        seq = torch.randint(0, cfg["num_tokens"], (B, h * w), generator=torch.Generator().manual_seed(seed))
        with torch.no_grad():
            dec = vae.decode(seq)
            with _cuda_autocast_policy_on_cpu():
                dec16 = vae.decode(seq).float()
        assert torch.allclose(dvae_ref.decode(seq, sd, cfg), dec, atol=1e-5)
        out[f"{name}/seq"] = seq.numpy()
        out[f"{name}/decode"] = dec.numpy()
        out[f"{name}/err/decode"] = np.array((dec16 - dec).double().norm().item())
        print(f"{name}: loss {l32:.6f} (bf16 {l16:.6f}), recons rel err {float(out[f'{name}/err/recons']) / float(out[f'{name}/ref/recons']):.2e}, "
              f"grad rel median {rels[len(rels) // 2]:.2e} max {rels[-1]:.2e}, decode {tuple(dec.shape)}")
    np.savez_compressed(os.path.join(GOLD, "dvae_train.npz"), **out)
    print("dvae_train.npz:", len(out), "arrays")


def golden_event_pipeline_var():
    """Reference build_transformNPY on its VARIABLE-sensor branch (data_path containing "Caltech" / "ncars": H = W = None,
    per-sample inferred sizes, bilinear antialiased Resize to the input size; mem/datasets.py:611-660) on seeded synthetic
    N-Caltech101-shaped (240x180, p = +-1) and N-Cars-shaped (120x100, p in {0,1}) streams."""
    import torch
    from types import SimpleNamespace
    ds = ref_shims.ref_module("datasets")
    out = {}
    cases = [  # name, data_path, (H, W), polarity, is_train, n_events, kind, normalize, seed
        ("cal_train_a", "/data/N-Caltech101", (180, 240), (-1.0, 1.0), True, 45000, "edge", 1, 21),
        ("cal_train_b", "/data/N-Caltech101", (180, 240), (-1.0, 1.0), True, 20000, "hot", 0, 22),
        ("cal_train_c", "/data/N-Caltech101", (172, 233), (-1.0, 1.0), True, 38000, "uniform", 1, 23),
        ("cal_eval_a", "/data/N-Caltech101", (180, 240), (-1.0, 1.0), False, 45000, "edge", 1, 24),
        ("cars_train_a", "/data/ncars", (100, 120), (0.0, 1.0), True, 9000, "edge", 1, 25),
        ("cars_eval_a", "/data/ncars", (100, 120), (0.0, 1.0), False, 6000, "uniform", 0, 26),
    ]
    for name, path, (H, W), pol, is_train, n, kind, norm, seed in cases:
        args = SimpleNamespace(data_path=path, input_H=224, input_W=224, slice_max_evs=30000, max_random_shift_evs=15,
                               timesurface=0, hotpixfilter=1, hotpix_num_stds=10, logtrafo=0, gammatrafo=0, gamma=0.5,
                               normalize_events=norm, rand_aug=0)
        import contextlib, io
        with contextlib.redirect_stdout(io.StringIO()):
            tf = ds.build_transformNPY(is_train, args)
        ev = np.floor(synth_events(np.random.default_rng(seed), n, H, W, kind, polarity=pol))     # recordings hold integer pixels
        random.seed(seed); np.random.seed(seed); torch.manual_seed(seed)
        res = tf(ev.copy())
        out[name + "_out"] = res.numpy()
        out[name + "_meta"] = np.array([int(is_train), n, norm, seed, H, W, int(pol[0] == 0.0)], dtype=np.int64)
        out[name + "_kind"] = np.array(kind)
        print(f"event_pipeline_var {name}: out {tuple(res.shape)} nnz {int((res != 0).sum())} max {float(res.max()):.4f}")
    np.savez_compressed(os.path.join(GOLD, "event_pipeline_var.npz"), **out)


def synth_event_image(rng, H, W):
    """uint8 [3, H, W] image shaped like the pipeline's output after ToUnit8: sparse polarity counts on a few edges plus
    noise, an empty middle channel, a few saturated pixels."""
    img = np.zeros((3, H, W), dtype=np.uint8)
    for c in (0, 2):
        n = H * W // 6
        y, x = rng.integers(0, H, n), rng.integers(0, W, n)
        np.add.at(img[c], (y, x), rng.integers(1, 40, n).astype(np.uint8))
        for _ in range(5):
            y0, x0 = rng.integers(0, H), rng.integers(0, W)
            t = np.arange(0, min(H, W) // 2)
            yy = np.clip(y0 + (t * rng.uniform(-1, 1)).astype(int), 0, H - 1)
            xx = np.clip(x0 + (t * rng.uniform(-1, 1)).astype(int), 0, W - 1)
            img[c, yy, xx] = rng.integers(80, 256, len(t)).astype(np.uint8)
    return img


def golden_randaug():
    """The reference's EventRandAugment (mem/transforms.py:351-471) on torchvision (not vendored; the image's 0.26.0):
    every operation of its augmentation space at several magnitude bins and both signs through ``_apply_op``, and the whole
    module under fixed torch seeds (pins the order of its three draws per operation)."""
    import contextlib, io
    import torch
    from torchvision.transforms import InterpolationMode
    tr = ref_shims.ref_module("transforms")
    from oracle import randaug_ref as R
    out = {}
    rng = np.random.default_rng(77)
    imgs = {"a": synth_event_image(rng, 48, 64), "b": synth_event_image(rng, 224, 224)}
    bins_of = {"a": (0, 3, 11, 20), "b": (7, 20)}
    for k, v in imgs.items():
        out[f"img_{k}"] = v
    with contextlib.redirect_stdout(io.StringIO()):
        aug = tr.EventRandAugment(small=False, magnitude=20)
    cases = []
    for key, img in imgs.items():
        H, W = img.shape[1:]
        space = aug._augmentation_space(aug.num_magnitude_bins, (H, W))
        for name, (mags, signed) in space.items():
            bins = bins_of[key] if mags.ndim > 0 else (0,)
            for i0 in bins:
                for neg in ((0, 1) if signed else (0,)):
                    mag = float(mags[i0].item()) if mags.ndim > 0 else 0.0
                    if signed and neg:
                        mag *= -1.0
                    res = tr._apply_op(torch.from_numpy(img.copy()), name, mag, interpolation=InterpolationMode.BILINEAR, fill=None)
                    tag = f"op_{key}_{name}_{i0}_{neg}"
                    out[tag] = res.numpy()
                    out[tag + "_mag"] = np.array(mag, dtype=np.float64)
                    cases.append(tag)
    out["cases"] = np.array(cases)
    for seed in range(12):
        for key, img in imgs.items():
            if key == "b" and seed >= 6:
                continue
            torch.manual_seed(1000 + seed)
            out[f"full_{key}_{seed}"] = aug(torch.from_numpy(img.copy())).numpy()
    # ToUnit8 / ToFloat32 around it (transforms.py:333-349) on a float image of counts / 255 and of counts / max
    x = torch.from_numpy(imgs["a"].astype(np.float32) / np.float32(255))
    out["tou8_in"], out["tou8_out"] = x.numpy(), tr.ToUnit8()(x.clone()).numpy()
    out["tof32_out"] = tr.ToFloat32()(torch.from_numpy(imgs["a"].copy())).numpy()
    y = x / x.max()
    out["tou8n_in"], out["tou8n_out"] = y.numpy(), tr.ToUnit8()(y.clone()).numpy()
    np.savez_compressed(os.path.join(GOLD, "randaug.npz"), **out)
    print(f"randaug: {len(cases)} single-operation cases, 24 seeded runs of the module")


def golden_event_pipeline_randaug():
    """Reference build_transformNPY WITH its default ``rand_aug=1`` tail (ToUnit8 -> EventRandAugment(magnitude=20) ->
    ToFloat32, mem/datasets.py:655-658) on the fixed-sensor and the variable-sensor branch: pins the order in which one
    sample's crop and RandAugment draws consume the torch generator."""
    import contextlib, io
    import torch
    from types import SimpleNamespace
    ds = ref_shims.ref_module("datasets")
    out = {}
    cases = [  # name, data_path, (H, W), polarity, n_events, kind, normalize, seed
        ("im_a", "/data/N_imagenet", (480, 640), (-1.0, 1.0), 45000, "edge", 1, 31),
        ("im_b", "/data/N_imagenet", (480, 640), (-1.0, 1.0), 45000, "hot", 0, 32),
        ("im_c", "/data/N_imagenet", (480, 640), (-1.0, 1.0), 20000, "uniform", 1, 33),
        ("im_d", "/data/N_imagenet", (480, 640), (-1.0, 1.0), 45000, "edge", 0, 34),
        ("cal_a", "/data/N-Caltech101", (180, 240), (-1.0, 1.0), 45000, "edge", 1, 35),
        ("cars_a", "/data/ncars", (100, 120), (0.0, 1.0), 9000, "edge", 1, 36),
    ]
    for name, path, (H, W), pol, n, kind, norm, seed in cases:
        args = SimpleNamespace(data_path=path, input_H=224, input_W=224, slice_max_evs=30000, max_random_shift_evs=15,
                               timesurface=0, hotpixfilter=1, hotpix_num_stds=10, logtrafo=0, gammatrafo=0, gamma=0.5,
                               normalize_events=norm, rand_aug=1)
        with contextlib.redirect_stdout(io.StringIO()):
            tf = ds.build_transformNPY(True, args)
        fixed = "imagenet" in path
        ev = synth_events(np.random.default_rng(seed), n, H, W, kind, polarity=pol, frac=(fixed and kind == "edge"))
        if not fixed:
            ev = np.floor(ev)
        random.seed(seed); np.random.seed(seed); torch.manual_seed(seed)
        res = tf(ev.copy())
        out[name + "_out"] = (res.numpy() * 255).round().astype(np.uint8)        # ToFloat32 output is exactly k / 255
        assert np.array_equal(out[name + "_out"].astype(np.float32) / np.float32(255), res.numpy())
        out[name + "_meta"] = np.array([n, norm, seed, H, W, int(pol[0] == 0.0), int(fixed)], dtype=np.int64)
        out[name + "_kind"] = np.array(kind)
        print(f"event_pipeline_randaug {name}: nnz {int((res != 0).sum())} max {float(res.max()):.4f}")
    np.savez_compressed(os.path.join(GOLD, "event_pipeline_randaug.npz"), **out)


def golden_event_pipeline_loggamma():
    """Reference build_transformNPY with LogTransform / GammaTransform switched on (args.logtrafo / args.gammatrafo,
    mem/datasets.py:647-650), fixed-sensor branch, with and without NormalizeEvent."""
    import contextlib, io
    import torch
    from types import SimpleNamespace
    ds = ref_shims.ref_module("datasets")
    out = {}
    cases = [  # name, is_train, n_events, kind, normalize, log, gamma on, gamma, seed
        ("log_a", True, 45000, "edge", 1, 1, 0, 0.5, 41), ("gam_a", True, 45000, "hot", 0, 0, 1, 0.5, 42),
        ("both_a", True, 20000, "uniform", 1, 1, 1, 0.7, 43), ("both_eval", False, 45000, "edge", 0, 1, 1, 0.5, 44),
    ]
    for name, is_train, n, kind, norm, lg, gm, gamma, seed in cases:
        args = SimpleNamespace(data_path="/data/N_imagenet", input_H=224, input_W=224, slice_max_evs=30000,
                               max_random_shift_evs=15, timesurface=0, hotpixfilter=1, hotpix_num_stds=10, logtrafo=lg,
                               gammatrafo=gm, gamma=gamma, normalize_events=norm, rand_aug=0)
        with contextlib.redirect_stdout(io.StringIO()):
            tf = ds.build_transformNPY(is_train, args)
        ev = synth_events(np.random.default_rng(seed), n, 480, 640, kind, frac=(kind == "edge"))
        random.seed(seed); np.random.seed(seed); torch.manual_seed(seed)
        res = tf(ev.copy())
        out[name + "_out"] = res.numpy()
        out[name + "_meta"] = np.array([int(is_train), n, norm, lg, gm, seed], dtype=np.int64)
        out[name + "_gamma"] = np.array(gamma, dtype=np.float64)
        out[name + "_kind"] = np.array(kind)
        print(f"event_pipeline_loggamma {name}: nnz {int((res != 0).sum())} max {float(res.max()):.4f}")
    np.savez_compressed(os.path.join(GOLD, "event_pipeline_loggamma.npz"), **out)


def golden_event_pipeline_tss():
    """Reference build_transformNPY with args.timesurface=1 (EventArrToImg(timeSurface=True) after the event-space
    augmentations, the middle channel kept), fixed-sensor branch: seeds chosen so that both time-flip outcomes occur."""
    import contextlib, io
    import torch
    from types import SimpleNamespace
    ds = ref_shims.ref_module("datasets")
    out = {}
    cases = [  # name, is_train, n_events, kind, normalize, seed
        ("tss_a", True, 45000, "edge", 1, 51), ("tss_b", True, 45000, "hot", 0, 52), ("tss_c", True, 20000, "uniform", 1, 53),
        ("tss_d", True, 45000, "edge", 0, 54), ("tss_e", True, 33000, "uniform", 0, 55), ("tss_eval", False, 45000, "edge", 1, 56),
    ]
    flips = []
    for name, is_train, n, kind, norm, seed in cases:
        args = SimpleNamespace(data_path="/data/N_imagenet", input_H=224, input_W=224, slice_max_evs=30000,
                               max_random_shift_evs=15, timesurface=1, hotpixfilter=1, hotpix_num_stds=10, logtrafo=0,
                               gammatrafo=0, gamma=0.5, normalize_events=norm, rand_aug=0)
        with contextlib.redirect_stdout(io.StringIO()):
            tf = ds.build_transformNPY(is_train, args)
        ev = synth_events(np.random.default_rng(seed), n, 480, 640, kind, frac=(kind == "edge"))
        random.seed(seed); np.random.seed(seed); torch.manual_seed(seed)
        res = tf(ev.copy())
        np.random.seed(seed)
        flips.append(bool(is_train and np.random.random() < 0.5))
        out[name + "_out"] = res.numpy()
        out[name + "_meta"] = np.array([int(is_train), n, norm, seed], dtype=np.int64)
        out[name + "_kind"] = np.array(kind)
        print(f"event_pipeline_tss {name}: nnz {int((res != 0).sum())} tss nnz {int((res[1] != 0).sum())} time flip {flips[-1]}")
    assert any(flips) and not all(flips[:5])
    np.savez_compressed(os.path.join(GOLD, "event_pipeline_tss.npz"), **out)


def golden_event_pipeline_tss_loggamma():
    """Reference build_transformNPY, fixed-sensor branch, args.timesurface=1 together with LogTransform / GammaTransform (the
    maps touch the polarity planes only, mem/transforms.py:205-208, :219-221)."""
    import contextlib, io
    import torch
    from types import SimpleNamespace
    ds = ref_shims.ref_module("datasets")
    out = {}
    cases = [  # name, is_train, n_events, kind, normalize, log, gamma on, gamma, seed
        ("tl_a", True, 45000, "edge", 1, 1, 0, 0.5, 81), ("tl_b", True, 20000, "hot", 0, 1, 1, 0.5, 82),
        ("tl_c", True, 38000, "uniform", 1, 0, 1, 0.7, 83), ("tl_eval", False, 45000, "edge", 1, 1, 1, 0.5, 84),
    ]
    for name, is_train, n, kind, norm, lg, gm, gamma, seed in cases:
        args = SimpleNamespace(data_path="/data/N_imagenet", input_H=224, input_W=224, slice_max_evs=30000,
                               max_random_shift_evs=15, timesurface=1, hotpixfilter=1, hotpix_num_stds=10, logtrafo=lg,
                               gammatrafo=gm, gamma=gamma, normalize_events=norm, rand_aug=0)
        with contextlib.redirect_stdout(io.StringIO()):
            tf = ds.build_transformNPY(is_train, args)
        ev = synth_events(np.random.default_rng(seed), n, 480, 640, kind, frac=(kind == "edge"))
        random.seed(seed); np.random.seed(seed); torch.manual_seed(seed)
        res = tf(ev.copy())
        out[name + "_out"] = res.numpy()
        out[name + "_meta"] = np.array([int(is_train), n, norm, lg, gm, seed], dtype=np.int64)
        out[name + "_gamma"] = np.array(gamma, dtype=np.float64)
        out[name + "_kind"] = np.array(kind)
        print(f"event_pipeline_tss_loggamma {name}: nnz {int((res != 0).sum())} tss nnz {int((res[1] != 0).sum())} max {float(res.max()):.4f}")
    np.savez_compressed(os.path.join(GOLD, "event_pipeline_tss_loggamma.npz"), **out)


def golden_event_pipeline_var_loggamma():
    """Reference build_transformNPY on its variable-sensor branch with LogTransform / GammaTransform on: there they act on
    the float32 image after Resize (mem/datasets.py:639, :648-651), not on integer counts."""
    import contextlib, io
    import torch
    from types import SimpleNamespace
    ds = ref_shims.ref_module("datasets")
    out = {}
    cases = [  # name, data_path, (H, W), polarity, is_train, n_events, kind, normalize, log, gamma on, gamma, seed
        ("cal_log", "/data/N-Caltech101", (180, 240), (-1.0, 1.0), True, 45000, "edge", 1, 1, 0, 0.5, 61),
        ("cal_sqrt", "/data/N-Caltech101", (180, 240), (-1.0, 1.0), True, 20000, "hot", 0, 0, 1, 0.5, 62),
        ("cal_both", "/data/N-Caltech101", (172, 233), (-1.0, 1.0), True, 38000, "uniform", 1, 1, 1, 0.7, 63),
        ("cars_both_eval", "/data/ncars", (100, 120), (0.0, 1.0), False, 6000, "edge", 0, 1, 1, 0.5, 64),
        ("cars_pow", "/data/ncars", (100, 120), (0.0, 1.0), True, 9000, "uniform", 1, 0, 1, 1.6, 65),
    ]
    for name, path, (H, W), pol, is_train, n, kind, norm, lg, gm, gamma, seed in cases:
        args = SimpleNamespace(data_path=path, input_H=224, input_W=224, slice_max_evs=30000, max_random_shift_evs=15,
                               timesurface=0, hotpixfilter=1, hotpix_num_stds=10, logtrafo=lg, gammatrafo=gm, gamma=gamma,
                               normalize_events=norm, rand_aug=0)
        with contextlib.redirect_stdout(io.StringIO()):
            tf = ds.build_transformNPY(is_train, args)
        ev = np.floor(synth_events(np.random.default_rng(seed), n, H, W, kind, polarity=pol))
        random.seed(seed); np.random.seed(seed); torch.manual_seed(seed)
        res = tf(ev.copy())
        out[name + "_out"] = res.numpy()
        out[name + "_meta"] = np.array([int(is_train), n, norm, lg, gm, seed, H, W, int(pol[0] == 0.0)], dtype=np.int64)
        out[name + "_gamma"] = np.array(gamma, dtype=np.float64)
        out[name + "_kind"] = np.array(kind)
        print(f"event_pipeline_var_loggamma {name}: nnz {int((res != 0).sum())} max {float(res.max()):.4f}")
    np.savez_compressed(os.path.join(GOLD, "event_pipeline_var_loggamma.npz"), **out)


def golden_event_pipeline_var_tss():
    """Reference build_transformNPY on its variable-sensor branch with args.timesurface=1: EventArrToImg(None, None, True)
    after the event-space augmentations, the time-surface plane resized with the polarity planes and kept."""
    import contextlib, io
    import torch
    from types import SimpleNamespace
    ds = ref_shims.ref_module("datasets")
    out = {}
    cases = [  # name, data_path, (H, W), polarity, is_train, n_events, kind, normalize, log, seed
        ("cal_tss_a", "/data/N-Caltech101", (180, 240), (-1.0, 1.0), True, 45000, "edge", 1, 0, 71),
        ("cal_tss_b", "/data/N-Caltech101", (180, 240), (-1.0, 1.0), True, 20000, "hot", 0, 0, 72),
        ("cal_tss_c", "/data/N-Caltech101", (172, 233), (-1.0, 1.0), True, 38000, "uniform", 1, 1, 73),
        ("cal_tss_d", "/data/N-Caltech101", (180, 240), (-1.0, 1.0), True, 26000, "edge", 0, 0, 74),
        ("cars_tss_eval", "/data/ncars", (100, 120), (0.0, 1.0), False, 6000, "edge", 0, 0, 75),
        ("cars_tss_a", "/data/ncars", (100, 120), (0.0, 1.0), True, 9000, "uniform", 1, 0, 76),
    ]
    flips = []
    for name, path, (H, W), pol, is_train, n, kind, norm, lg, seed in cases:
        args = SimpleNamespace(data_path=path, input_H=224, input_W=224, slice_max_evs=30000, max_random_shift_evs=15,
                               timesurface=1, hotpixfilter=1, hotpix_num_stds=10, logtrafo=lg, gammatrafo=0, gamma=0.5,
                               normalize_events=norm, rand_aug=0)
        with contextlib.redirect_stdout(io.StringIO()):
            tf = ds.build_transformNPY(is_train, args)
        ev = np.floor(synth_events(np.random.default_rng(seed), n, H, W, kind, polarity=pol))
        random.seed(seed); np.random.seed(seed); torch.manual_seed(seed)
        res = tf(ev.copy())
        random.seed(seed); np.random.seed(seed)
        if n > 30000:
            random.choice(range(n - 30000 + 1))
        flips.append(bool(is_train and np.random.random() < 0.5))
        out[name + "_out"] = res.numpy()
        out[name + "_meta"] = np.array([int(is_train), n, norm, lg, seed, H, W, int(pol[0] == 0.0)], dtype=np.int64)
        out[name + "_kind"] = np.array(kind)
        print(f"event_pipeline_var_tss {name}: nnz {int((res != 0).sum())} tss nnz {int((res[1] != 0).sum())} time flip {flips[-1]}")
    assert any(flips) and not all(flips[:4])
    np.savez_compressed(os.path.join(GOLD, "event_pipeline_var_tss.npz"), **out)


SECTIONS = {"histogram": golden_histogram, "masks": golden_masks, "vit": golden_vit, "dvae": golden_dvae,
            "engine": golden_engine, "event_pipeline": golden_event_pipeline, "decode": golden_decode, "engine_ft": golden_engine_ft,
            "vit_bf16": golden_vit_bf16, "finetune_remap": golden_finetune_remap, "dvae_train": golden_dvae_train, "event_pipeline_var": golden_event_pipeline_var, "randaug": golden_randaug,
            "event_pipeline_randaug": golden_event_pipeline_randaug, "event_pipeline_loggamma": golden_event_pipeline_loggamma,
            "event_pipeline_tss": golden_event_pipeline_tss, "event_pipeline_var_loggamma": golden_event_pipeline_var_loggamma,
            "event_pipeline_var_tss": golden_event_pipeline_var_tss, "event_pipeline_tss_loggamma": golden_event_pipeline_tss_loggamma}


def main(argv):
    os.makedirs(GOLD, exist_ok=True)
    names = argv or list(SECTIONS)
    for n in names:
        SECTIONS[n]()


if __name__ == "__main__":
    main(sys.argv[1:])
