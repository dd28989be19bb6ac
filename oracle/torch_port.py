"""Functional PyTorch-CPU restatement of the reference's Generator.forward -- TEST INFRASTRUCTURE.

This is the "port" CPU baseline of bench.py: the same arithmetic the reference performs on a CPU
(torch -> oneDNN conv kernels, all host threads), written against folded weights so it needs
neither /root/reference nor the nn.Module tree.  It is pinned against the reference's golden
outputs by tests/test_oracle.py.  Never imported by the product path.

Follows /root/reference/models.py:61-71 (Generator.forward) and :32-40 (ResBlock.forward);
weight-norm fold per SURVEY 0.3 (melgan_multi_b200.synth.fold_weight_norm).
"""
import torch
import torch.nn.functional as F

_DIL = (1, 3, 9)


def fold_state(state):
    """state: name -> ndarray (synth.generator_state).  Returns (weights[30], biases[30]) torch CPU."""
    from melgan_multi_b200.synth import GENERATOR_LAYERS, fold_weight_norm
    ws, bs = [], []
    for name, *_ in GENERATOR_LAYERS:
        ws.append(torch.from_numpy(fold_weight_norm(state[name + ".weight_g"], state[name + ".weight_v"])))
        bs.append(torch.from_numpy(state[name + ".bias"]))
    return ws, bs


@torch.no_grad()
def generator_forward(ws, bs, mel):
    x = F.conv1d(mel, ws[0], bs[0], padding=3)
    for i in range(4):
        k = ws[1 + i].shape[2]
        x = F.conv_transpose1d(F.leaky_relu(x), ws[1 + i], bs[1 + i], stride=k // 2, padding=k // 4)
        for j, d in enumerate(_DIL):
            a, b = 5 + 6 * i + j, 5 + 6 * i + 3 + j
            h = F.conv1d(F.leaky_relu(x), ws[a], bs[a], padding=d, dilation=d)
            x = F.conv1d(F.leaky_relu(h), ws[b], bs[b], padding=1) + x
    return torch.tanh(F.conv1d(F.leaky_relu(x), ws[29], bs[29], padding=3))


def reference_state(state):
    """state: name -> ndarray.  Returns [(weight_v, weight_g, bias)] * 30 as torch CPU tensors: what the reference's modules
    hold (old-style weight_norm keeps g and v and recomputes w = g * v / ||v|| in a pre-forward hook on EVERY forward)."""
    from melgan_multi_b200.synth import GENERATOR_LAYERS
    return [(torch.from_numpy(state[n + ".weight_v"]), torch.from_numpy(state[n + ".weight_g"]), torch.from_numpy(state[n + ".bias"]))
            for n, *_ in GENERATOR_LAYERS]


@torch.no_grad()
def generator_forward_reference(params, mel):
    """Exactly the work of the reference's Generator.forward on a CPU (models.py:61-71 + the 30 weight_norm pre-forward
    hooks, torch._weight_norm(v, g, 0) as torch.nn.utils.weight_norm computes it): this is what bench.py's reference arm and
    cpu_baseline time."""
    ws = [torch._weight_norm(v, g, 0) for v, g, _ in params]
    return generator_forward(ws, [b for _, _, b in params], mel)
