"""Deterministic synthetic weights and inputs for the MelGAN hot path.

There is no checkpoint in the reference and no network on the GPU box, so every test,
golden fixture and benchmark in this repo uses weights generated here from a seed.  The
generator is numpy's frozen legacy ``RandomState`` (MT19937), whose streams are stable
across numpy versions and machines, so the 18 MB of weights never have to be committed:
only the seed travels.

Shapes, names and registration order follow the reference's ``state_dict`` exactly
(/root/reference/models.py:44-59 for the Generator, :75-85 and :109-113 for the
discriminators; old-style ``weight_norm`` registers ``bias, weight_g, weight_v`` per layer).
Initialisation mimics PyTorch's default (uniform in +-1/sqrt(fan_in)) and then scales
``weight_g`` by U(0.5, 1.5) so that g != ||v|| and a wrong weight-norm fold is visible
(SURVEY.md section 7, step 0).
"""
from collections import OrderedDict

import numpy as np

# (name, kind, C_in, C_out, K) in reference registration order.  kind: "conv" | "convT".
GENERATOR_LAYERS = [("conv_pre", "conv", 80, 512, 7)]
GENERATOR_LAYERS += [
    ("ups.0", "convT", 512, 256, 16),
    ("ups.1", "convT", 256, 128, 16),
    ("ups.2", "convT", 128, 64, 4),
    ("ups.3", "convT", 64, 32, 4),
]
for _i, _c in enumerate((256, 128, 64, 32)):
    for _grp in ("convs1", "convs2"):
        for _j in range(3):
            GENERATOR_LAYERS.append(("resblocks.%d.%s.%d" % (_i, _grp, _j), "conv", _c, _c, 3))
GENERATOR_LAYERS.append(("conv_post", "conv", 32, 1, 7))

# (name, C_in, C_out, K, stride, groups, padding) for one Discriminator (models.py:77-85).
DISCRIMINATOR_LAYERS = [
    ("conv_pre", 1, 16, 15, 1, 1, 7),
    ("grouped_convs.0", 16, 64, 41, 4, 4, 20),
    ("grouped_convs.1", 64, 256, 41, 4, 16, 20),
    ("grouped_convs.2", 256, 1024, 41, 4, 64, 20),
    ("grouped_convs.3", 1024, 1024, 41, 1, 256, 20),
    ("conv_post1", 1024, 1024, 5, 1, 1, 2),
    ("conv_post2", 1024, 1, 3, 1, 1, 1),
]


def _layer_params(rs, shape_v, fan_in, n_bias, norm_axes):
    bound = 1.0 / np.sqrt(fan_in)
    v = rs.uniform(-bound, bound, size=shape_v).astype(np.float32)
    norm = np.sqrt((v.astype(np.float64) ** 2).sum(axis=norm_axes, keepdims=True))
    g = (norm * rs.uniform(0.5, 1.5, size=norm.shape)).astype(np.float32)
    b = rs.uniform(-bound, bound, size=(n_bias,)).astype(np.float32)
    return b, g, v


def generator_state(seed=1234):
    """OrderedDict name -> float32 ndarray with the 90 Generator tensors, reference order."""
    rs = np.random.RandomState(seed)
    sd = OrderedDict()
    for name, kind, cin, cout, k in GENERATOR_LAYERS:
        if kind == "conv":  # Conv1d weight [C_out, C_in, K]; weight_norm dim=0 -> per C_out
            b, g, v = _layer_params(rs, (cout, cin, k), cin * k, cout, (1, 2))
        else:  # ConvTranspose1d weight [C_in, C_out, K]; dim=0 -> per C_in
            b, g, v = _layer_params(rs, (cin, cout, k), cout * k, cout, (1, 2))
        sd[name + ".bias"] = b
        sd[name + ".weight_g"] = g
        sd[name + ".weight_v"] = v
    return sd


def discriminator_state(seed=4321):
    """OrderedDict with the 63 MultiScaleDiscriminator tensors, reference order."""
    rs = np.random.RandomState(seed)
    sd = OrderedDict()
    for d in range(3):
        for name, cin, cout, k, _stride, groups, _pad in DISCRIMINATOR_LAYERS:
            b, g, v = _layer_params(rs, (cout, cin // groups, k), (cin // groups) * k, cout, (1, 2))
            base = "discriminators.%d.%s" % (d, name)
            sd[base + ".bias"] = b
            sd[base + ".weight_g"] = g
            sd[base + ".weight_v"] = v
    return sd


def mel_input(batch, frames, seed=0, realistic=False):
    """Synthetic mel batch [B, 80, T] float32.  ``realistic`` draws from the range of real
    log-mels, U(-11.5, 2) (meldataset.py:22 clips at log(1e-5)); default is N(0, 1)."""
    rs = np.random.RandomState(seed)
    if realistic:
        return rs.uniform(-11.5, 2.0, size=(batch, 80, frames)).astype(np.float32)
    return rs.standard_normal(size=(batch, 80, frames)).astype(np.float32)


def audio_input(batch, samples, seed=0):
    """Synthetic audio segments [B, 1, L] in U(-1, 1)."""
    rs = np.random.RandomState(seed + 1000)
    return rs.uniform(-1.0, 1.0, size=(batch, 1, samples)).astype(np.float32)


def fold_weight_norm(g, v):
    """w = g * v / ||v|| with the norm over every axis but 0 (weight_norm dim=0), float64
    internally.  Valid for both Conv1d ([C_out, ...]) and ConvTranspose1d ([C_in, ...])."""
    v64 = v.astype(np.float64)
    norm = np.sqrt((v64 ** 2).sum(axis=tuple(range(1, v.ndim)), keepdims=True))
    return (g.astype(np.float64) * v64 / norm).astype(np.float32)
