"""GPU parity of the tensor-core (tcgen05, split-bf16) path against the C oracle / reference goldens.
Tolerance: north_star allows 1e-3 relative; the 3-pass split keeps us near 1e-5, asserted at 1e-4."""
import os

import numpy as np
import pytest
import torch

import cases
from conftest import rel_errors
from melgan_multi_b200 import engine, synth
from oracle import cport

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.fixture(scope="module")
def state():
    return synth.generator_state(1234)


@pytest.fixture(scope="module")
def dev(state):
    gd = engine.GeneratorDevice("cuda:0")
    order = [n for n, *_ in synth.GENERATOR_LAYERS]
    to = lambda a: torch.from_numpy(a).cuda()
    gd.pack([to(state[n + ".weight_v"]) for n in order], [to(state[n + ".weight_g"]) for n in order],
            [to(state[n + ".bias"]) for n in order])
    return gd


@pytest.fixture()
def tc_path():
    old = os.environ.get("MG_GEN_PATH")
    os.environ["MG_GEN_PATH"] = "tc"
    yield
    if old is None:
        os.environ.pop("MG_GEN_PATH", None)
    else:
        os.environ["MG_GEN_PATH"] = old


def oracle_resblock(state, stage, x):
    """ResBlock.forward (models.py:32-40) with the oracle's primitives."""
    lr = lambda a: np.where(a > 0, a, a * np.float32(0.01)).astype(np.float32)
    for j, d in enumerate((1, 3, 9)):
        n1, n2 = "resblocks.%d.convs1.%d" % (stage, j), "resblocks.%d.convs2.%d" % (stage, j)
        w1 = cport.fold_weight_norm(state[n1 + ".weight_g"], state[n1 + ".weight_v"])
        w2 = cport.fold_weight_norm(state[n2 + ".weight_g"], state[n2 + ".weight_v"])
        h = cport.conv1d(lr(x), w1, state[n1 + ".bias"], 1, d, d, 1)
        x = cport.conv1d(lr(h), w2, state[n2 + ".bias"], 1, 1, 1, 1) + x
    return x


@pytest.mark.parametrize("stage,B,L", [(3, 2, 2100), (2, 2, 1000), (1, 2, 500), (0, 2, 200), (3, 1, 5), (0, 1, 8),
                                       (1, 1, 224), (1, 1, 225)])
def test_resblock_tc_matches_oracle(state, dev, stage, B, L):
    C = 256 >> stage
    rs = np.random.RandomState(stage * 100 + L)
    x = rs.standard_normal((B, C, L)).astype(np.float32)
    ref = oracle_resblock(state, stage, x)
    y = dev.resblock(stage, torch.from_numpy(x).cuda()).cpu().numpy()
    m, l2 = rel_errors(y, ref)
    assert m < TOL and l2 < TOL, (stage, B, L, m, l2)


@pytest.mark.parametrize("stage,B,L", [(0, 2, 32), (0, 64, 32), (0, 1, 1), (0, 3, 130), (1, 2, 256), (1, 1, 5),
                                       (2, 2, 700), (2, 1, 1), (3, 2, 1500), (3, 1, 511), (3, 1, 512), (3, 1, 513)])
def test_convt_tc_matches_oracle(state, dev, stage, B, L):
    cin = 512 >> stage
    S, pad = (8, 4) if stage < 2 else (2, 1)
    rs = np.random.RandomState(stage * 1000 + L + B)
    x = rs.standard_normal((B, cin, L)).astype(np.float32)
    name = "ups.%d" % stage
    w = cport.fold_weight_norm(state[name + ".weight_g"], state[name + ".weight_v"])
    lr = np.where(x > 0, x, x * np.float32(0.01)).astype(np.float32)
    ref = cport.conv_transpose1d(lr, w, state[name + ".bias"], S, pad)
    y = dev.convt(stage, torch.from_numpy(x).cuda()).cpu().numpy()
    assert y.shape == ref.shape
    m, l2 = rel_errors(y, ref)
    assert m < TOL and l2 < TOL, (stage, B, L, m, l2)


@pytest.mark.parametrize("stage,B,Lin", [(2, 2, 500), (2, 1, 1), (2, 1, 64), (2, 3, 129), (2, 1, 2048), (3, 2, 1050), (3, 1, 1),
                                         (3, 1, 128), (3, 1, 255), (3, 2, 256), (3, 1, 4096)])
def test_fused_convt_resblock_matches_oracle(state, dev, stage, B, Lin):
    """Stage 2 / 3 as ONE kernel (LeakyReLU -> ConvT k4 s2 -> ResBlock; what the pipeline runs) against the oracle's
    conv_transpose1d + ResBlock; odd and tiny lengths cover the pair de-interleave at both sequence ends, long ones the
    tile borders."""
    cin = 512 >> stage
    rs = np.random.RandomState(stage * 777 + Lin + B)
    x = rs.standard_normal((B, cin, Lin)).astype(np.float32)
    name = "ups.%d" % stage
    w = cport.fold_weight_norm(state[name + ".weight_g"], state[name + ".weight_v"])
    lr = np.where(x > 0, x, x * np.float32(0.01)).astype(np.float32)
    ref = oracle_resblock(state, stage, cport.conv_transpose1d(lr, w, state[name + ".bias"], 2, 1))
    y = dev.upres(stage, torch.from_numpy(x).cuda()).cpu().numpy()
    assert y.shape == ref.shape
    m, l2 = rel_errors(y, ref)
    assert m < TOL and l2 < TOL, (stage, B, Lin, m, l2)


@pytest.mark.parametrize("case", cases.GEN_CASES)
def test_tc_pipeline_matches_golden(golden, state, tc_path, case):
    B, T, seed, realistic = case
    eng = engine.GeneratorHost(B, T)
    eng.load_state(state)
    y = eng.forward(synth.mel_input(B, T, seed, realistic))
    eng.close()
    m, l2 = rel_errors(y, golden[cases.gen_key(*case)])
    assert m < TOL and l2 < TOL, (case, m, l2)


def test_tc_pipeline_config2_vs_simt(state, tc_path):
    """Config 2 (B=64, T=32): the two independent implementations (fp32 SIMT, split-bf16 tcgen05) agree."""
    x = synth.mel_input(64, 32, 0)
    eng = engine.GeneratorHost(64, 32)
    eng.load_state(state)
    y_tc = eng.forward(x)
    os.environ["MG_GEN_PATH"] = "simt"
    y_simt = eng.forward(x)
    eng.close()
    m, l2 = rel_errors(y_tc, y_simt)
    assert m < TOL and l2 < TOL, (m, l2)
