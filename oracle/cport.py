"""ctypes loader for oracle/melgan_oracle.c (plain-C restatement of the reference path).

TEST INFRASTRUCTURE: the checker, never the thing measured as the product or shipped.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libmelgan_oracle.so")
_lib = None

_f32p = ctypes.POINTER(ctypes.c_float)
_f32pp = ctypes.POINTER(_f32p)


def build(force=False):
    src = os.path.join(_HERE, "melgan_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.mgo_generator_forward.restype = ctypes.c_int
        _lib.mgo_conv1d_out_len.restype = ctypes.c_int
        _lib.mgo_avgpool1d_out_len.restype = ctypes.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(_f32p)


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def fold_weight_norm(g, v):
    v = _c(v)
    g = _c(g).reshape(-1)
    w = np.empty_like(v)
    lib().mgo_fold_weight_norm(_p(v), _p(g), _p(w), ctypes.c_int(v.shape[0]),
                               ctypes.c_int(int(np.prod(v.shape[1:]))))
    return w


def conv1d(x, w, b, stride=1, pad=0, dil=1, groups=1):
    x, w = _c(x), _c(w)
    B, cin, lin = x.shape
    cout, _, k = w.shape
    lout = lib().mgo_conv1d_out_len(lin, k, stride, pad, dil)
    y = np.empty((B, cout, lout), np.float32)
    bb = _c(b) if b is not None else None
    lib().mgo_conv1d(_p(x), _p(w), _p(bb) if bb is not None else None, _p(y), B, cin, lin, cout, k,
                     stride, pad, dil, groups)
    return y


def conv_transpose1d(x, w, b, stride, pad):
    x, w = _c(x), _c(w)
    B, cin, lin = x.shape
    _, cout, k = w.shape
    lout = (lin - 1) * stride - 2 * pad + k
    y = np.empty((B, cout, lout), np.float32)
    bb = _c(b) if b is not None else None
    lib().mgo_conv_transpose1d(_p(x), _p(w), _p(bb) if bb is not None else None, _p(y), B, cin, lin,
                               cout, k, stride, pad)
    return y


def avgpool1d(x, k, stride, pad):
    x = _c(x)
    B, c, lin = x.shape
    lout = lib().mgo_avgpool1d_out_len(lin, k, stride, pad)
    y = np.empty((B, c, lout), np.float32)
    lib().mgo_avgpool1d(_p(x), _p(y), B * c, lin, k, stride, pad)
    return y


def fold_generator(state):
    """state: name -> ndarray (synth.generator_state layout).  Returns (w[30], b[30]) folded,
    in reference registration order."""
    from melgan_multi_b200.synth import GENERATOR_LAYERS
    ws, bs = [], []
    for name, _kind, _cin, _cout, _k in GENERATOR_LAYERS:
        ws.append(fold_weight_norm(state[name + ".weight_g"], state[name + ".weight_v"]))
        bs.append(_c(state[name + ".bias"]))
    return ws, bs


def generator_forward(ws, bs, mel, want_stages=False):
    """ws/bs: folded weights/biases (fold_generator).  mel [B,80,T] -> audio [B,1,256T].
    With want_stages also returns [conv_pre out, resblock0..3 out, pre-tanh] as NCL arrays."""
    mel = _c(mel)
    B, _, T = mel.shape
    audio = np.empty((B, 1, 256 * T), np.float32)
    wp = (_f32p * 30)(*[_p(w) for w in ws])
    bp = (_f32p * 30)(*[_p(b) for b in bs])
    stages = None
    sp = None
    if want_stages:
        shapes = [(B, 512, T), (B, 256, 8 * T), (B, 128, 64 * T), (B, 64, 128 * T), (B, 32, 256 * T),
                  (B, 1, 256 * T)]
        stages = [np.empty(s, np.float32) for s in shapes]
        sp = (_f32p * 6)(*[_p(s) for s in stages])
    rc = lib().mgo_generator_forward(wp, bp, _p(mel), _p(audio), B, T, sp)
    if rc != 0:
        raise MemoryError("oracle allocation failed")
    return (audio, stages) if want_stages else audio


_D_SPECS = [(1, 16, 15, 1, 1, 7), (16, 64, 41, 4, 4, 20), (64, 256, 41, 4, 16, 20),
            (256, 1024, 41, 4, 64, 20), (1024, 1024, 41, 1, 256, 20), (1024, 1024, 5, 1, 1, 2),
            (1024, 1, 3, 1, 1, 1)]


def fold_discriminators(state):
    """Returns list of 3 (w[7], b[7]) tuples from synth.discriminator_state layout."""
    from melgan_multi_b200.synth import DISCRIMINATOR_LAYERS
    out = []
    for d in range(3):
        ws, bs = [], []
        for name, *_ in DISCRIMINATOR_LAYERS:
            base = "discriminators.%d.%s" % (d, name)
            ws.append(fold_weight_norm(state[base + ".weight_g"], state[base + ".weight_v"]))
            bs.append(_c(state[base + ".bias"]))
        out.append((ws, bs))
    return out


def discriminator_forward(ws, bs, y):
    """One Discriminator: y [B,1,L] -> (logits [B,l], fmaps list of 7)."""
    y = _c(y)
    B, _, L = y.shape
    lens = (ctypes.c_int * 7)()
    lib().mgo_discriminator_lengths(L, lens)
    fmaps = [np.empty((B, _D_SPECS[l][1], lens[l]), np.float32) for l in range(7)]
    wp = (_f32p * 7)(*[_p(w) for w in ws])
    bp = (_f32p * 7)(*[_p(b) for b in bs])
    fp = (_f32p * 7)(*[_p(f) for f in fmaps])
    lib().mgo_discriminator_forward(wp, bp, _p(y), B, L, fp)
    return fmaps[6].reshape(B, -1), fmaps


def msd_forward(folded, y, y_hat):
    """MultiScaleDiscriminator.forward restated (models.py:119-135)."""
    pools = [(4, 2, 2), (4, 4, 2)]
    y_d_rs, y_d_gs, fmap_rs, fmap_gs = [], [], [], []
    for i, (ws, bs) in enumerate(folded):
        if i != 0:
            y = avgpool1d(y, *pools[i - 1])
            y_hat = avgpool1d(y_hat, *pools[i - 1])
        r, fr = discriminator_forward(ws, bs, y)
        g, fg = discriminator_forward(ws, bs, y_hat)
        y_d_rs.append(r); fmap_rs.append(fr); y_d_gs.append(g); fmap_gs.append(fg)
    return y_d_rs, y_d_gs, fmap_rs, fmap_gs
