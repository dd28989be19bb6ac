"""BASELINE config 3 context number (NOT bench.py's metric): one full training step (train.py:108-129: G step through the
discriminators, then D step on the detached audio; Adam on both) at batch 16 x 8192 samples, fp32.
  ours  : drop-in modules -- native forwards (generator, discriminators, fused losses), backward by recomputation
          through stock PyTorch ops (the open row of DESIGN.md section 8)
  ours_with_multi_tensor_adam : + melgan_multi_b200.optim.Adam (one launch per optimizer step) instead of torch.optim.Adam
  stock : the same modules' stock-PyTorch restatement (_torch_forward) end to end, cuDNN default (TF32) and strict fp32
Prints step times and the relative difference of the losses of the first step."""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from melgan_multi_b200 import models, synth


def build():
    g = models.Generator()
    g.load_state_dict({k: torch.from_numpy(v) for k, v in synth.generator_state(1234).items()})
    d = models.MultiScaleDiscriminator()
    d.load_state_dict({k: torch.from_numpy(v) for k, v in synth.discriminator_state(4321).items()})
    return g.cuda().train(), d.cuda().train()


def generator_stock_forward(x, leaves):
    """The reference's own formulation (nn.Conv1d / nn.ConvTranspose1d on NCL tensors, models.py:61-71 of the reference) --
    the drop-in's recompute (models.Generator._torch_forward) runs the same graph on channels_last 4-D tensors instead."""
    import torch.nn.functional as F
    ws = [torch._weight_norm(leaves[3 * i], leaves[3 * i + 1], 0) for i in range(30)]
    bs = [leaves[3 * i + 2] for i in range(30)]
    x = F.conv1d(x, ws[0], bs[0], padding=3)
    for i in range(4):
        k = ws[1 + i].shape[2]
        x = F.conv_transpose1d(F.leaky_relu(x), ws[1 + i], bs[1 + i], stride=k // 2, padding=k // 4)
        for j, d in enumerate((1, 3, 9)):
            a, b = 5 + 6 * i + j, 5 + 6 * i + 3 + j
            h = F.conv1d(F.leaky_relu(x), ws[a], bs[a], padding=d, dilation=d)
            x = F.conv1d(F.leaky_relu(h), ws[b], bs[b], padding=1) + x
    return torch.tanh(F.conv1d(F.leaky_relu(x), ws[29], bs[29], padding=3))


class Stock(torch.nn.Module):
    """Routes forward() through the module's stock-PyTorch restatement (autograd records the usual cuDNN graph)."""

    def __init__(self, m, msd):
        super().__init__()
        self.m, self.msd = m, msd

    def forward(self, *a):
        vs, gs, bs = self.m._param_triplets()
        leaves = [t for trip in zip(vs, gs, bs) for t in trip]
        if not self.msd:
            return generator_stock_forward(a[0], leaves)
        y, y_hat = a
        B = y.shape[0]
        flat = self.m._torch_forward(torch.cat([y, y_hat]), leaves)
        fm = [flat[7 * s:7 * s + 7] for s in range(3)]
        rs = [f[-1][:B].flatten(1) for f in fm]
        gs_ = [f[-1][B:].flatten(1) for f in fm]
        return rs, gs_, [[t[:B] for t in f] for f in fm], [[t[B:] for t in f] for f in fm]


def torch_losses():
    def feature_loss(fr, fg):
        return sum((r - g).abs().mean() for a, b in zip(fr, fg) for r, g in zip(a, b)) * 10

    def generator_loss(dg):
        return sum(((1 - g) ** 2).mean() for g in dg)

    def discriminator_loss(dr, dg):
        r = [((1 - x) ** 2).mean() for x in dr]
        g = [(x ** 2).mean() for x in dg]
        return sum(r) + sum(g), [v.item() for v in r], [v.item() for v in g]
    return feature_loss, generator_loss, discriminator_loss


def run(gen, disc, losses, x, y, steps, warmup, adam=torch.optim.Adam):
    feature_loss, generator_loss, discriminator_loss = losses
    params_g = [p for p in gen.parameters()]
    params_d = [p for p in disc.parameters()]
    og = adam(params_g, 1e-4, betas=(0.5, 0.9))
    od = adam(params_d, 1e-4, betas=(0.5, 0.9))
    first = None
    times = []
    for it in range(warmup + steps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        og.zero_grad()
        y_hat = gen(x)
        dr, dg, fr, fg = disc(y, y_hat)
        loss_gen = generator_loss(dg) + feature_loss(fr, fg)
        lg = loss_gen.item()
        loss_gen.backward()
        og.step()
        od.zero_grad()
        dr, dg, _, _ = disc(y, y_hat.detach())
        loss_disc, _, _ = discriminator_loss(dr, dg)
        ld = loss_disc.item()
        loss_disc.backward()
        od.step()
        torch.cuda.synchronize()
        if it >= warmup:
            times.append(time.perf_counter() - t0)
        if first is None:
            first = (lg, ld)
    times.sort()
    return 1e3 * times[len(times) // 2], first


def main():
    steps, warmup = 10, 3
    x = torch.from_numpy(synth.mel_input(16, 32, 0)).cuda()
    y = torch.from_numpy(synth.audio_input(16, 8192, 0)).cuda()
    out = {"config": "BASELINE config 3: B=16, 80x32 mel / 8192-sample segments, fp32, G step + D step + Adam"}
    g, d = build()
    ours_losses = (models.feature_loss, models.generator_loss, models.discriminator_loss)
    ms, first = run(g, d, ours_losses, x, y, steps, warmup)
    out["ours_ms"], out["ours_first_losses"] = ms, first
    from melgan_multi_b200.optim import Adam
    g, d = build()
    out["ours_with_multi_tensor_adam_ms"] = run(g, d, ours_losses, x, y, steps, warmup, adam=Adam)[0]
    for prec in ("tf32", "ieee"):
        torch.backends.cudnn.conv.fp32_precision = prec
        g, d = build()
        ms, lf = run(Stock(g, False), Stock(d, True), torch_losses(), x, y, steps, warmup)
        out["stock_%s_ms" % prec], out["stock_%s_first_losses" % prec] = ms, lf
    ref = out["stock_ieee_first_losses"]
    out["ours_vs_stock_ieee_loss_rel"] = [abs(a / b - 1) for a, b in zip(out["ours_first_losses"], ref)]
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
