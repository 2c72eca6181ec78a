"""Which SDPA backends run (and capture into a CUDA graph) with dropout on this box, and how fast."""
import torch, time
import torch.nn.functional as F
from torch.nn.attention import sdpa_kernel, SDPBackend

dev = "cuda"
B, H, T, D = 256, 12, 128, 64
q, k, v = (torch.randn(B, H, T, D, device=dev, dtype=torch.bfloat16, requires_grad=True) for _ in range(3))

def step(p):
    o = F.scaled_dot_product_attention(q, k, v, dropout_p=p)
    o.sum().backward()

for name, be in [("cudnn", SDPBackend.CUDNN_ATTENTION), ("flash", SDPBackend.FLASH_ATTENTION),
                 ("efficient", SDPBackend.EFFICIENT_ATTENTION), ("math", SDPBackend.MATH)]:
    for p in (0.1, 0.0):
        try:
            with sdpa_kernel([be]):
                for _ in range(3): step(p)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(10): step(p)
                e1.record(); torch.cuda.synchronize()
                eager = e0.elapsed_time(e1) / 10
        except Exception as e:
            print(name, p, "eager FAIL", type(e).__name__, str(e)[:120]); continue
        try:
            with sdpa_kernel([be]):
                s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(s):
                    for _ in range(3): step(p)
                torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    step(p)
                g.replay(); torch.cuda.synchronize()
                e0.record()
                for _ in range(10): g.replay()
                e1.record(); torch.cuda.synchronize()
                print(name, p, f"eager {eager:.3f} ms  graph OK {e0.elapsed_time(e1)/10:.3f} ms", flush=True)
        except Exception as e:
            print(name, p, f"eager {eager:.3f} ms  graph FAIL", type(e).__name__, str(e)[:160], flush=True)
            torch.cuda.synchronize()
