mkdir -p gpurun_out; cd "${GRAFT_REPO_ROOT:-.}"
timeout 600 python scripts/compare_stock_pytorch.py > gpurun_out/compare_stock.json 2> gpurun_out/compare_stock.err; tail -3 gpurun_out/compare_stock.err; cat gpurun_out/compare_stock.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_msd.csv python -c "
import sys; sys.path.insert(0,'.')
import torch
from melgan_multi_b200 import models, synth
d = models.MultiScaleDiscriminator(); d.load_state_dict({k: torch.from_numpy(v) for k, v in synth.discriminator_state(4321).items()}); d = d.cuda().eval()
y = torch.from_numpy(synth.audio_input(16, 8192, 0)).cuda(); yh = torch.from_numpy(synth.audio_input(16, 8192, 1)).cuda()
with torch.no_grad():
    for _ in range(3): d(y, yh)
torch.cuda.synchronize()
" > gpurun_out/ncu_msd.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches_msd.csv
timeout 900 python -m pytest tests/test_disc_gpu.py -m gpu -q 2>&1 | tail -3
