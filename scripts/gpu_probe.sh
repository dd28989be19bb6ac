mkdir -p gpurun_out; cd "${GRAFT_REPO_ROOT:-.}"
timeout 120 ./melgan_multi_b200/csrc/probe/tc_probe > gpurun_out/tc_probe.log 2>&1; echo "rc=$?" >> gpurun_out/tc_probe.log
cat gpurun_out/tc_probe.log
{ echo "nproc=$(nproc)"; cat /sys/fs/cgroup/cpu.max 2>/dev/null; python - <<'PY'
import os, time, torch, sys
sys.path.insert(0, '.')
print("affinity", len(os.sched_getaffinity(0)), "cpu_count", os.cpu_count(), "torch threads", torch.get_num_threads())
from melgan_multi_b200 import synth
from oracle import torch_port
ws, bs = torch_port.fold_state(synth.generator_state(1234))
x = torch.from_numpy(synth.mel_input(4, 32, 0))
for n in (4, 8, 16, 32, 64, 128):
    torch.set_num_threads(n)
    torch_port.generator_forward(ws, bs, x)
    t0 = time.perf_counter(); torch_port.generator_forward(ws, bs, x); dt = time.perf_counter() - t0
    print("threads", n, "B=4 forward %.3f s" % dt, flush=True)
    if dt > 20: break
PY
} > gpurun_out/cpu_diag.log 2>&1
cat gpurun_out/cpu_diag.log
