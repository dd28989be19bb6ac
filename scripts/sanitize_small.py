"""Tiny generator + discriminator forwards for compute-sanitizer (memcheck / racecheck / synccheck)."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from melgan_multi_b200 import engine, models, synth

g = models.Generator()
g.load_state_dict({k: torch.from_numpy(v) for k, v in synth.generator_state(1234).items()})
g = g.cuda().eval()
d = models.MultiScaleDiscriminator()
d.load_state_dict({k: torch.from_numpy(v) for k, v in synth.discriminator_state(4321).items()})
d = d.cuda().eval()
with torch.no_grad():
    y = g(torch.from_numpy(synth.mel_input(1, 3, 5)).cuda())
    g._dev.check_status(1, 3)
    out = d(y, torch.from_numpy(synth.audio_input(1, 768, 1)).cuda())
    d._dev.check_status()
torch.cuda.synchronize()
print("ok", float(y.abs().sum()), float(out[0][0].abs().sum()))
