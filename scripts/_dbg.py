import sys, os, numpy as np, torch
sys.path.insert(0, ".")
from melgan_multi_b200 import engine, synth
state = synth.generator_state(1234)
gd = engine.GeneratorDevice("cuda:0")
order = [n for n, *_ in synth.GENERATOR_LAYERS]
to = lambda a: torch.from_numpy(a).cuda()
gd.pack([to(state[n + ".weight_v"]) for n in order], [to(state[n + ".weight_g"]) for n in order], [to(state[n + ".bias"]) for n in order])
for stage, B, L in [(2,1,64),(2,1,256),(2,2,500),(1,2,300),(0,3,256)]:
    x = torch.randn(B, 256 >> stage, L, device="cuda")
    try:
        y = gd.resup(stage, x); torch.cuda.synchronize(); print("resup", stage, B, L, "ok", float(y.abs().mean()))
    except Exception as e:
        print("resup", stage, B, L, "FAIL", str(e)[:100]); break
