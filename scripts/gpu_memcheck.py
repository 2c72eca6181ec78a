"""Debug aid: allocated device memory after each eager training step of the bench workload (leak check)."""
import os, sys, gc
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("HF_HUB_OFFLINE", "1")
import torch
import bayeformers_b200 as bf
from transformers import BertConfig, BertForSequenceClassification

links = int(os.environ.get("LINKS", "1"))
bf.runtime.enable_gelu_links(bool(links))
dev = torch.device("cuda", 0)
torch.manual_seed(0)
cfg = BertConfig(num_labels=2, num_hidden_layers=4)
model = BertForSequenceClassification(cfg)
bf.manual_seed(1)
bm = bf.to_bayesian(model, delta=0.05, freeze=True, gemm_dtype="bf16", kl_grad=True)
bf.accelerate_host_(bm, fuse_residual=True, grad_sinks=True)
bm = bm.to(dev).train()
bf.enable_presample(bm)
bf.cast_frequentist_(bm, torch.bfloat16)
params = [p for p in bm.parameters() if p.requires_grad]
opt = bf.optim.ClipAdamW(params, lr=2e-5, max_grad_norm=1.0)
bf.enable_device_step(dev)
S, B, T = 4, 64, 128
ids = torch.randint(0, cfg.vocab_size, (B, T), device=dev)
labels = torch.randint(0, 2, (B,), device=dev)
for it in range(6):
    bf.advance_step()
    opt.zero_grad(set_to_none=True)
    with bf.mc_samples(S):
        logits = bm(input_ids=ids.repeat(S, 1)).logits
    loss = torch.nn.functional.cross_entropy(logits.float().view(S, B, -1).mean(0), labels)
    loss = loss + (bm.log_variational_posterior().mean() - bm.log_prior().mean()) / 1000
    loss.backward()
    opt.step()
    del logits, loss
    torch.cuda.synchronize()
    print(f"links={links} step {it}: allocated {torch.cuda.memory_allocated() / 2**30:.2f} GiB  "
          f"peak {torch.cuda.max_memory_allocated() / 2**30:.2f} GiB  gc objects {len(gc.get_objects())}", flush=True)

# ---- who holds the survivors?
before = torch.cuda.memory_allocated()
n = gc.collect()
print(f"gc.collect() freed {n} objects, allocated {before / 2**30:.2f} -> {torch.cuda.memory_allocated() / 2**30:.2f} GiB", flush=True)
import collections
big = [o for o in gc.get_objects() if isinstance(o, torch.Tensor) and o.is_cuda and o.numel() * o.element_size() >= 32 << 20]
print("large tensors alive:", collections.Counter((tuple(t.shape), str(t.dtype)) for t in big).most_common(12))
for t in big[:3]:
    refs = [r for r in gc.get_referrers(t) if r is not big]
    print(tuple(t.shape), "referrers:", [type(r).__name__ + (":" + ",".join(list(r.keys())[:6]) if isinstance(r, dict) else "") for r in refs][:6])
    for r in refs[:3]:
        rr = gc.get_referrers(r)
        print("    <-", [type(x).__name__ for x in rr][:6])
