"""One generator forward (config 2) and one multi-scale-discriminator forward (B=16+16, L=8192) after warm-up, for ncu."""
import sys

import torch

sys.path.insert(0, ".")
from melgan_multi_b200 import models, synth

g = models.Generator()
g.load_state_dict({k: torch.from_numpy(v) for k, v in synth.generator_state(1234).items()})
g = g.cuda().eval()
d = models.MultiScaleDiscriminator()
d.load_state_dict({k: torch.from_numpy(v) for k, v in synth.discriminator_state(4321).items()})
d = d.cuda().eval()
x = torch.from_numpy(synth.mel_input(64, 32, 0)).cuda()
y = torch.from_numpy(synth.audio_input(16, 8192, 0)).cuda()
yh = torch.from_numpy(synth.audio_input(16, 8192, 1)).cuda()
with torch.no_grad():
    for _ in range(3):  # warm-up (weights packed once; kernels configured)
        g(x); d(y, yh)
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_push("measured")
    g(x); d(y, yh)
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_pop()
