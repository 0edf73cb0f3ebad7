"""Run the tokenizer on N images a few times (for ncu: per-layer conv launch list / --set full of one launch).
The last pass sits between cudaProfilerStart/Stop: with `ncu --profile-from-start off`, launch k of that pass is
im2col (0), act0..act3 (1-4), res blocks (5-13), head (14), argmax_decode (15)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mem_b200.vae_model import DiscreteVAE
from oracle import dvae_ref
cfg = dict(input_H=224, input_W=224, num_tokens=8192, codebook_dim=32, num_layers=4, num_resnet_blocks=3, hidden_dim=384, channels=2)
torch.manual_seed(0)
vae = DiscreteVAE(**cfg).cuda()
img = dvae_ref.synth_images(int(sys.argv[1]) if len(sys.argv) > 1 else 64, 2, 224, 224, seed=6).cuda()
for _ in range(2):
    vae.get_codebook_indices(img)
torch.cuda.synchronize()
torch.cuda.profiler.start()
vae.get_codebook_indices(img)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
