"""Times the policy network alone (4096 x 227 -> 1024 -> 512 -> 28): dm_mlp_forward (tcgen05 kernels) against the fp32 torch actor, CUDA events."""
import os, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch
from deepmimic_b200.capi import TensorCoreMLP
rows, din, h0, h1, dout = int(os.environ.get("MLP_ROWS", "4096")), 227, 1024, 512, 28
rng = np.random.default_rng(0)
w0 = (rng.standard_normal((din, h0)) / np.sqrt(din)).astype(np.float32); w1 = (rng.standard_normal((h0, h1)) / np.sqrt(h0)).astype(np.float32); w2 = (rng.standard_normal((h1, dout)) / np.sqrt(h1)).astype(np.float32)
b0, b1, b2 = (rng.standard_normal(n).astype(np.float32) * 0.1 for n in (h0, h1, dout))
mean, std = rng.standard_normal(din).astype(np.float32), (0.5 + rng.random(din)).astype(np.float32)
mlp = TensorCoreMLP(w0, b0, w1, b1, w2, b2, in_mean=mean, in_std=std, in_clip=5.0, out_mean=np.zeros(dout, np.float32), out_std=np.ones(dout, np.float32), max_rows=rows)
x = torch.randn(rows, din, device="cuda"); out = torch.zeros(rows, dout, device="cuda"); noise = torch.zeros(rows, dout, device="cuda")
st = torch.cuda.current_stream()
tw = [torch.tensor(a, device="cuda") for a in (w0, b0, w1, b1, w2, b2, mean, std)]
def torch_actor():
    h = torch.clamp((x - tw[6]) / tw[7], -5, 5)
    h = torch.relu(h @ tw[0] + tw[1]); h = torch.relu(h @ tw[2] + tw[3]); return h @ tw[4] + tw[5]
def timeit(f, n=50):
    for _ in range(5): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1000.0
t_tc = timeit(lambda: mlp.forward(x, out, noise=noise, stream=st.cuda_stream))
t_th = timeit(torch_actor)
ref = torch_actor(); torch.cuda.synchronize()
print("policy network, %d rows: tcgen05 kernels %.1f us per forward, fp32 torch actor %.1f us; max |diff| %.2e" % (rows, t_tc, t_th, (out - ref).abs().max().item()))
