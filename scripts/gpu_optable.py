"""torch.profiler table of one eager training step (which ops launch the many tiny kernels?)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("HF_HUB_OFFLINE", "1")
import torch
import bench
import bayeformers_b200 as bf
from bayeformers_b200 import parallel

dev = torch.device("cuda", 0)
class A: layers = 0
model, cfg = bench.build_bert(0)
bf.manual_seed(1234)
bm = bf.to_bayesian(model, delta=0.05, freeze=True, gemm_dtype="bf16", kl_grad=True)
bf.accelerate_host_(bm)
bm = bm.to(dev).train()
bf.enable_presample(bm)
bf.cast_frequentist_(bm, torch.bfloat16)
params = [p for p in bm.parameters() if p.requires_grad]
optim = torch.optim.AdamW(params, lr=2e-5, eps=1e-8, fused=True)
bf.enable_device_step(dev)
B, T, S = 64, 128, 4
ids = torch.randint(0, cfg.vocab_size, (B, T), device=dev)
labels = torch.randint(0, 2, (B,), device=dev)

def step():
    bf.advance_step()
    optim.zero_grad(set_to_none=True)
    with bf.mc_samples(S):
        logits = bm(input_ids=ids.repeat(S, 1)).logits
    raw = logits.float().view(S, B, -1)
    nll = torch.nn.functional.cross_entropy(raw.mean(0), labels)
    lp, lq = bm.log_prior().mean(), bm.log_variational_posterior().mean()
    loss = (lq - lp) / 1000 + nll
    loss.backward()
    torch.nn.utils.clip_grad_norm_(params, 1.0)
    optim.step()

for _ in range(3): step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=60, max_name_column_width=70))
