"""Probe (GPU box): dVAE tokenizer accuracy vs an fp64 oracle and time per image for several
tensor-core accumulation segment lengths.  Writes gpurun_out/dvae_probe.json."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mem_b200.vae_model import DiscreteVAE  # noqa: E402
from oracle import dvae_ref  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
torch.manual_seed(0)
cfg = dict(input_H=224, input_W=224, num_tokens=8192, codebook_dim=32, num_layers=4, num_resnet_blocks=3, hidden_dim=384, channels=2)
vae = DiscreteVAE(**cfg).cuda()
B = 8
img = dvae_ref.synth_images(B, 2, 224, 224, seed=5).cuda()
sd = {k: v.detach().double() for k, v in vae.state_dict().items()}
with torch.no_grad():
    ref = dvae_ref.encoder_logits(img.double(), sd, 4, 3)
    ref32 = dvae_ref.encoder_logits(img, {k: v.float() for k, v in sd.items()}, 4, 3)
scale = ref.abs().max().item()
flat = ref.reshape(B, 8192, -1).transpose(1, 2)
top2 = flat.topk(2, -1).values
margin = (top2[..., 0] - top2[..., 1])
out = {"scale": scale, "oracle_margin_min": margin.min().item(), "oracle_margin_median": margin.median().item(),
       "torch_fp32_cuda_max_err_vs_fp64": (ref32.double() - ref).abs().max().item(),
       "torch_fp32_cuda_token_diffs": int((ref32.flatten(2).argmax(1) != ref.flatten(2).argmax(1)).sum()), "runs": []}
big = dvae_ref.synth_images(128, 2, 224, 224, seed=6).cuda()
for precision, segs in (("f16x2", (1, 2, 4, 100000)), ("tf32x3", (2, 4))):
    vae.tokenizer_precision = precision
    object.__setattr__(vae, "_tok", None)
    tok = vae._tokenizer()
    for seg in segs:
        tok.seg_kblocks = seg
        logits = vae(img, return_logits=True)
        idx = vae.get_codebook_indices(img)
        err = (logits.double() - ref).abs().max().item()
        diffs = int((idx != ref.flatten(2).argmax(1)).sum())
        vae.get_codebook_indices(big)
        torch.cuda.synchronize()
        t0 = time.time()
        for _ in range(3):
            vae.get_codebook_indices(big)
        torch.cuda.synchronize()
        ms = (time.time() - t0) / 3 * 1e3
        vae.verify_range()
        out["runs"].append({"precision": precision, "seg_kblocks": seg, "max_logit_err_vs_fp64": err, "token_diffs_of": [diffs, idx.numel()],
                            "ms_per_128_images": ms, "tflops_algorithmic": 24.26e9 * 128 / (ms * 1e-3) / 1e12,
                            "exps": dict(tok.exps) if tok.exps else None})
        print(out["runs"][-1], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/dvae_probe.json", "w"), indent=1)
print(json.dumps({k: v for k, v in out.items() if k != "runs"}))
