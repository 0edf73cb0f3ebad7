"""Does the tokenizer of batch k+1 hide under the ViT step of batch k?  Sequential vs two-stream software pipeline
(same work per step: one tokenizer pass + one ViT fwd/bwd/AdamW), CUDA-event timed."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from benchmarks import pretrain as P
from mem_b200.vit_engine import pretrain_step

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
model, vae, opt = P.build(torch, dev)
B, nb = 128, 4
batches = [P.synth_batch(torch, B, i, dev) for i in range(nb)]
batches = [(s, im, mk.to(dev)) for s, im, mk in batches]
cap = B * P.CFG["num_mask"]
opt.grad_divisor = 1.0


def vit(samples, masks, tokens):
    opt.zero_grad()
    pretrain_step(model, samples, masks.flatten(1), tokens, cap=cap)
    opt.step(max_norm=P.CFG["max_norm"])


def sequential(n):
    for i in range(n):
        s, im, mk = batches[i % nb]
        vit(s, mk, vae.get_codebook_indices(im))


side = torch.cuda.Stream(device=dev, priority=int(os.environ.get("SIDE_PRIO", "0")))


def pipelined(n):
    cur = torch.cuda.current_stream(dev)
    s, im, mk = batches[0]
    tokens = vae.get_codebook_indices(im)
    for i in range(n):
        s, im, mk = batches[i % nb]
        nxt = batches[(i + 1) % nb]
        start = torch.cuda.Event(); start.record(cur)
        with torch.cuda.stream(side):
            side.wait_event(start)
            nxt_tokens = vae.get_codebook_indices(nxt[1])
            done = torch.cuda.Event(); done.record(side)
        vit(s, mk, tokens)
        cur.wait_event(done)
        tokens = nxt_tokens


def timed(fn, n=12, warm=4):
    fn(warm)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(n); e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


print("sequential ms/step", round(timed(sequential), 3))
print("pipelined  ms/step", round(timed(pipelined), 3))
print("sequential ms/step", round(timed(sequential), 3))
