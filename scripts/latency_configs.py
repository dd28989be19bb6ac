"""Latency of the generator forward at the BASELINE configs that are not bench.py's headline (context numbers):
config 1 (B=1, T=32), config 5 (B=1, T=1000: 11.6 s of audio), and B=16 (the training batch).  Device time (CUDA events,
median) through models.Generator, and end to end through the host engine (pinned buffers)."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from melgan_multi_b200 import engine, models, synth


def main():
    g = models.Generator()
    state = synth.generator_state(1234)
    g.load_state_dict({k: torch.from_numpy(v) for k, v in state.items()})
    g = g.cuda().eval()
    out = {}
    for name, B, T in (("config1_B1_T32", 1, 32), ("config5_B1_T1000", 1, 1000), ("B16_T32", 16, 32)):
        x = torch.from_numpy(synth.mel_input(B, T, 0)).cuda()
        host = engine.GeneratorHost(B, T)
        host.load_state(state)
        mel_h = synth.mel_input(B, T, 0)
        with torch.no_grad():
            for _ in range(10):
                g(x)
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(50)]
            torch.cuda.synchronize()
            for a, b in ev:
                a.record(); g(x); b.record()
            torch.cuda.synchronize()
        ms = sorted(a.elapsed_time(b) for a, b in ev)
        # the same forward captured once in a CUDA graph and replayed (tests/test_generator_gpu.py checks equality)
        with torch.no_grad():
            sx = x.clone()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                g(sx)
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                g(sx)
            for _ in range(5):
                graph.replay()
            gev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(50)]
            torch.cuda.synchronize()
            for a, b in gev:
                a.record(); graph.replay(); b.record()
            torch.cuda.synchronize()
        gms = sorted(a.elapsed_time(b) for a, b in gev)
        for _ in range(5):
            host.forward(mel_h)
        ts = []
        for _ in range(30):
            t0 = time.perf_counter(); host.forward(mel_h); ts.append(time.perf_counter() - t0)
        ts.sort()
        out[name] = {"device_ms_median": ms[len(ms) // 2], "device_ms_min": ms[0], "cuda_graph_replay_ms_median": gms[len(gms) // 2], "e2e_ms_median": 1e3 * ts[len(ts) // 2],
                     "audio_seconds": B * T * 256 / 22050.0, "slices": engine.lib().mg_gen_forward_slices(B, T)}
        host.close()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
