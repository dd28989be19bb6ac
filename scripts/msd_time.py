"""Times the multi-scale discriminator forward (B real + B generated, length L) through the drop-in module.
MG_DISC_GROUP=simt selects the fp32 SIMT grouped convs.  CUDA events, L2 flushed between steps."""
import json
import os
import sys

import torch

sys.path.insert(0, ".")
from melgan_multi_b200 import models, synth


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    L = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    d = models.MultiScaleDiscriminator()
    d.load_state_dict({k: torch.from_numpy(v) for k, v in synth.discriminator_state(4321).items()})
    d = d.cuda().eval()
    y = torch.from_numpy(synth.audio_input(B, L, 0)).cuda()
    yh = torch.from_numpy(synth.audio_input(B, L, 1)).cuda()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    with torch.no_grad():
        for _ in range(5):
            d(y, yh)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        torch.cuda.synchronize()
        for a, b in ev:
            flush.zero_()
            a.record(); d(y, yh); b.record()
        torch.cuda.synchronize()
    d._dev.check_status()
    ms = sorted(a.elapsed_time(b) for a, b in ev)
    print(json.dumps({"B": B, "L": L, "group_path": os.environ.get("MG_DISC_GROUP", "tc"), "msd_forward_ms_median": ms[len(ms) // 2],
                      "min": ms[0]}))


if __name__ == "__main__":
    main()
