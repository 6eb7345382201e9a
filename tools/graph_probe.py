"""Does CUDA-graph replay of one 420M micro-step (fwd+loss+bwd) beat eager launches? (launch-gap probe)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from collections import namedtuple
from plainlm_b200.models import construct_model

c = dict(vocab_size=50280, d_model=1024, n_layers=24, n_heads=16, seq_len=2048, expand='8/3', mlp_class='glu',
         tie_embeddings=False, model='transformer')
model, _ = construct_model(namedtuple('C', c.keys())(**c))
model = model.to('cuda')
rt = model.runtime()
ids = torch.randint(0, 50280, (8, 2049), device='cuda')
x, y = ids[:, :2048].contiguous(), ids[:, 1:].contiguous()

def step():
  return rt.loss_and_backward(x, y, None, grad_scale=0.25)

def timeit(fn, n=8):
  for _ in range(3):
    fn()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(n):
    fn()
  e1.record()
  torch.cuda.synchronize()
  return e0.elapsed_time(e1) / n

t_eager = timeit(step)
t0 = time.perf_counter()
for _ in range(4):
  step()
host_ms = (time.perf_counter() - t0) / 4 * 1e3
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
  step()
  torch.cuda.synchronize()
  with torch.cuda.graph(g, stream=s):
    step()
torch.cuda.synchronize()
t_graph = timeit(g.replay)
print(f'eager {t_eager:.3f} ms/micro-step (host enqueue {host_ms:.2f} ms), graph replay {t_graph:.3f} ms -> {100 * (t_eager - t_graph) / t_eager:.1f} % faster')
