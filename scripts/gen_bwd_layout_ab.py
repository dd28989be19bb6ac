"""A/B of the generator's stock-op recompute+backward (config 3: B=16, 80x32 mel) under different activation layouts:
NCL (what the drop-in used so far: cuDNN converts to NHWC and back around every TF32 kernel) vs 4-D [N, C, 1, L] channels_last
tensors through conv2d (cuDNN's NHWC kernels then run without the conversions).  Measured: 5.66 -> 4.95 ms eager; NOT adopted
(about 0.2 ms under the backward's CUDA graph, and different TF32 kernels: see models.Generator._torch_forward)."""
import sys
import time

import torch

sys.path.insert(0, ".")
sys.path.insert(0, "scripts")
from melgan_multi_b200 import models, synth
from train_step_time import generator_stock_forward as fwd3

torch.manual_seed(0)
g = models.Generator().cuda()
x = torch.from_numpy(synth.mel_input(16, 32, 0)).cuda()
params = []
for m in models._layer_modules(g):
    params += [m.weight_v, m.weight_g, m.bias]




def fwd4(x, leaves):
    import torch.nn.functional as F
    ws = [torch._weight_norm(leaves[3 * i], leaves[3 * i + 1], 0).unsqueeze(2) for i in range(30)]
    bs = [leaves[3 * i + 2] for i in range(30)]
    x = x.unsqueeze(2).contiguous(memory_format=torch.channels_last)
    x = F.conv2d(x, ws[0], bs[0], padding=(0, 3))
    for i in range(4):
        k = ws[1 + i].shape[3]
        x = F.conv_transpose2d(F.leaky_relu(x), ws[1 + i], bs[1 + i], stride=(1, k // 2), padding=(0, k // 4))
        for j, d in enumerate((1, 3, 9)):
            a, b = 5 + 6 * i + j, 5 + 6 * i + 3 + j
            h = F.conv2d(F.leaky_relu(x), ws[a], bs[a], padding=(0, d), dilation=(1, d))
            x = F.conv2d(F.leaky_relu(h), ws[b], bs[b], padding=(0, 1)) + x
    return torch.tanh(F.conv2d(F.leaky_relu(x), ws[29], bs[29], padding=(0, 3))).squeeze(2)


def run(mode, n=20):
    leaves = [p.detach().clone().requires_grad_(True) for p in params]
    mel = x.clone()
    def step():
        y = fwd4(mel, leaves) if mode == "nlc" else fwd3(mel, leaves)
        gr = torch.autograd.grad(y, leaves, torch.ones_like(y))
        return y, gr
    y, gr = step()
    torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        step()
    t0.record()
    for _ in range(n):
        step()
    t1.record(); torch.cuda.synchronize()
    return t0.elapsed_time(t1) / n, y, gr


for bench in (False, True):
    torch.backends.cudnn.benchmark = bench
    a, ya, ga = run("ncl")
    b, yb, gb = run("nlc")
    err = max(float((p - q).abs().max() / (p.abs().max() + 1e-30)) for p, q in zip(ga, gb))
    print(f"cudnn.benchmark={bench}: NCL {a:.3f} ms  NLC {b:.3f} ms   y diff {float((ya - yb).abs().max()):.2e}  grad rel diff {err:.2e}"
          f"  y strides {tuple(yb.stride())}")
