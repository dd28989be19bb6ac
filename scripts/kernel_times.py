"""Per-kernel device times (mg_gen_forward_timed, single chain, L2 flushed) for a list of batch sizes at T frames.
usage: python scripts/kernel_times.py [T] [B ...]"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from melgan_multi_b200 import engine, synth

T = int(sys.argv[1]) if len(sys.argv) > 1 else 32
Bs = [int(a) for a in sys.argv[2:]] or [8, 16, 32, 64]
state = synth.generator_state(1234)
gd = engine.GeneratorDevice("cuda:0")
order = [n for n, *_ in synth.GENERATOR_LAYERS]
to = lambda a: torch.from_numpy(a).cuda()
gd.pack([to(state[n + ".weight_v"]) for n in order], [to(state[n + ".weight_g"]) for n in order],
        [to(state[n + ".bias"]) for n in order])
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for B in Bs:
    mel = to(synth.mel_input(B, T, 0))
    out = torch.empty((B, 1, 256 * T), device="cuda")
    acc, names = None, None
    for it in range(13):
        flush.zero_()
        t = gd.forward_timed(mel, out)
        if it >= 3:
            v = np.array([x for _, x in t])
            acc = v if acc is None else acc + v
            names = [n for n, _ in t]
    acc /= 10
    print("B=%3d T=%d: total %.1f us | " % (B, T, 1e3 * acc.sum()) + " ".join("%s %.1f" % (n, 1e3 * v) for n, v in zip(names, acc)))
