"""Where the time of one training step (BASELINE config 3) goes on the GPU: torch.profiler kernel table of the drop-in
modules' step (native forwards + torch-recompute backward)."""
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, ".")
sys.path.insert(0, "scripts")
from melgan_multi_b200 import models, synth
from train_step_time import build, run

x = torch.from_numpy(synth.mel_input(16, 32, 0)).cuda()
y = torch.from_numpy(synth.audio_input(16, 8192, 0)).cuda()
g, d = build()
losses = (models.feature_loss, models.generator_loss, models.discriminator_loss)
import os
adam = torch.optim.Adam
if os.environ.get("MG_ADAM") == "1":  # the package's multi-tensor Adam (one launch per optimizer step)
    from melgan_multi_b200.optim import Adam as adam
run(g, d, losses, x, y, 2, 2, adam)
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    run(g, d, losses, x, y, 3, 0, adam)
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=90))
