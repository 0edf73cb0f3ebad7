"""Margin of tests/test_engine_ft_gpu.py::test_accumulation_equals_one_big_batch over 5 repetitions."""
import os, sys, io, contextlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_engine_ft_gpu import _build
from mem_b200 import engine_for_finetuning as eft, utils
from oracle import engine_ref
batches = engine_ref.synth_class_batches()[:2]
big = [(torch.cat([b[0] for b in batches]), torch.cat([b[1] for b in batches]))]
for rep in range(5):
    with contextlib.redirect_stdout(io.StringIO()):
        m1, o1 = _build(); s1 = eft.train_one_epoch(None, m1, torch.nn.CrossEntropyLoss(), batches, o1, "cuda", 0, utils.NativeScalerWithGradNormCount(), 1.0, update_freq=2)
        m2, o2 = _build(); s2 = eft.train_one_epoch(None, m2, torch.nn.CrossEntropyLoss(), big, o2, "cuda", 0, utils.NativeScalerWithGradNormCount(), 1.0, update_freq=1)
    worst = max((a - b).abs().mean().item() for (n, a), (_, b) in zip(m1.state_dict().items(), m2.state_dict().items()) if a.is_floating_point())
    print(f"rep {rep}: grad_norm rel diff {abs(s1['grad_norm']-s2['grad_norm'])/s2['grad_norm']:.2e}, loss rel diff {abs(s1['loss']-s2['loss'])/s2['loss']:.2e}, worst mean |dw| {worst:.2e}")
