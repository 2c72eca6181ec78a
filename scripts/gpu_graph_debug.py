"""Which autograd nodes does the BERT step contain in eager mode vs under CUDA-graph capture?"""
import collections, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("HF_HUB_OFFLINE", "1")
import torch
from transformers import BertConfig, BertForSequenceClassification

dev = "cuda"
cfg = BertConfig(num_labels=2, num_hidden_layers=1)
m = BertForSequenceClassification(cfg).to(dev).bfloat16().train()
ids = torch.randint(0, 1000, (32, 128), device=dev)

def nodes(t):
    seen, cnt = set(), collections.Counter()
    def walk(fn):
        if fn is None or fn in seen: return
        seen.add(fn); cnt[type(fn).__name__] += 1
        for nf, _ in fn.next_functions: walk(nf)
    walk(t.grad_fn)
    return {k: v for k, v in cnt.items() if "Bmm" in k or "Attention" in k or "Softmax" in k}

out = m(input_ids=ids).logits
print("eager    :", nodes(out.sum()), flush=True)
out.sum().backward()
torch.cuda.synchronize()
s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(2):
        m(input_ids=ids).logits.sum().backward()
torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
try:
    with torch.cuda.graph(g):
        out = m(input_ids=ids).logits
        print("capturing:", nodes(out.sum()), flush=True)
        from torch.nn.attention import SDPBackend
        q = torch.randn(32, 12, 128, 64, device=dev, dtype=torch.bfloat16)
        print("capturing, plain sdpa dropout:", nodes(torch.nn.functional.scaled_dot_product_attention(q.requires_grad_(), q, q, dropout_p=0.1).sum()), flush=True)
        out.sum().backward()
    print("capture OK")
except Exception as e:
    print("capture FAIL", type(e).__name__, str(e)[:300])
