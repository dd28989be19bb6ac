"""BASELINE config 5 (B=1, 80x1000 mel = 11.6 s of audio) cut along TIME over the ranks of a torchrun job:
each rank generates its frames +- 8 (no data-path collective), optionally followed by one all_gather of the audio.
Prints, from rank 0: device ms of the local generate (max over ranks), with the gather, and the difference to the
single-GPU whole-utterance result.   torchrun --nproc-per-node N scripts/utterance_shard_time.py"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from melgan_multi_b200 import distributed as mgd
from melgan_multi_b200 import models, synth


def timed(fn, n=30):
    for _ in range(5):
        fn()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    torch.cuda.synchronize()
    for a, b in ev:
        dist.barrier()
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in ev)
    t = torch.tensor([ms[len(ms) // 2]], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    real_out = os.fdopen(os.dup(1), "w")  # the JSON line goes here; fd 1 (NCCL's banner, warnings) is sent to stderr
    os.dup2(2, 1)
    dist.init_process_group("nccl")
    g = models.Generator()
    g.load_state_dict({k: torch.from_numpy(v) for k, v in synth.generator_state(1234).items()})
    g = g.cuda().eval()
    mel = torch.from_numpy(synth.mel_input(1, 1000, 0)).cuda()
    with torch.no_grad():
        whole = g(mel)
        full = mgd.generate_sharded(g, mel)
    out = {"world": world, "frames": 1000,
           "whole_utterance_one_gpu_ms": timed(lambda: g(mel)) if True else None,
           "sharded_local_ms": timed(lambda: mgd.generate_sharded(g, mel, gather=False)),
           "sharded_with_all_gather_ms": timed(lambda: mgd.generate_sharded(g, mel)),
           "max_abs_diff_vs_whole": float((full - whole).abs().max())}
    if rank == 0:
        real_out.write(json.dumps(out) + "\n")
        real_out.flush()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
