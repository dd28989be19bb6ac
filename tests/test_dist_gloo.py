"""world_size-2 gloo tests (CPU) of the N>1 host logic: the data-parallel gradient wrapper that replaces the
reference's distributed.py functions, and batch sharding of generator inference (no data-path collective)."""
import os
import socket
import warnings

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

warnings.filterwarnings("ignore")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class _TinyDisc(torch.nn.Module):
    """A small weight-normed conv stack with the parameter structure the wrapper meets in practice (weight_g /
    weight_v / bias leaves, grouped conv); plain PyTorch so that it runs on CPU under gloo."""

    def __init__(self):
        super().__init__()
        wn = torch.nn.utils.weight_norm
        self.pre = wn(torch.nn.Conv1d(1, 16, 15, padding=7))
        self.grp = wn(torch.nn.Conv1d(16, 64, 41, 4, groups=4, padding=20))
        self.post = wn(torch.nn.Conv1d(64, 1, 3, padding=1))

    def forward(self, x):
        fm = [torch.nn.functional.leaky_relu(self.pre(x))]
        fm.append(torch.nn.functional.leaky_relu(self.grp(fm[-1])))
        out = self.post(fm[-1])
        return out.flatten(1), fm


def _make_model(seed):
    torch.manual_seed(seed)
    return _TinyDisc()


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from melgan_multi_b200 import distributed as mgd
    torch.set_num_threads(1)
    # this test has no optimizer, so every gradient is "never consumed": switch the learned skipping of discarded gradients
    # off and check the plain reference semantics (it has its own tests below)
    os.environ["MG_DDP_DEDUP"] = "0"
    model = _make_model(100 + rank)  # ranks start from DIFFERENT weights; wrap must broadcast rank 0's
    mgd.apply_gradient_allreduce(model)
    params_after_bcast = torch.cat([p.detach().reshape(-1) for p in model.parameters()]).clone()

    torch.manual_seed(7 + rank)
    x = torch.randn(2, 1, 256)
    logits, _ = model(x)
    loss = (logits ** 2).mean()
    red = mgd.reduce_tensor(loss.detach(), world)
    loss.backward()
    g1 = torch.cat([p.grad.reshape(-1) for p in model.parameters()]).clone()

    # second step after zero_grad(set_to_none=True): grads are re-adopted into the flat buffer
    for p in model.parameters():
        p.grad = None
    logits, _ = model(x * 0.5)
    (logits ** 2).mean().backward()
    g2 = torch.cat([p.grad.reshape(-1) for p in model.parameters()]).clone()

    # skipped reduction keeps the local gradient
    for p in model.parameters():
        p.grad = None
    mgd.skip_next_reduction(model)
    logits, _ = model(x)
    (logits ** 2).mean().backward()
    g3 = torch.cat([p.grad.reshape(-1) for p in model.parameters()]).clone()
    torch.save({"params": params_after_bcast, "g1": g1, "g2": g2, "g3": g3, "loss": loss.detach(), "red": red},
               os.path.join(out_dir, "r%d.pt" % rank))
    dist.destroy_process_group()


def _single_rank_grads(rank, scale=1.0):
    """What rank `rank` would compute alone, starting from rank 0's weights."""
    model = _make_model(100)
    torch.manual_seed(7 + rank)
    x = torch.randn(2, 1, 256) * scale
    logits, _ = model(x)
    loss = (logits ** 2).mean()
    loss.backward()
    return torch.cat([p.grad.reshape(-1) for p in model.parameters()]), loss.detach()


def test_gradient_allreduce_wrapper_two_ranks(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r = [torch.load(os.path.join(tmp_path, "r%d.pt" % i)) for i in range(world)]
    ref_params = torch.cat([p.detach().reshape(-1) for p in _make_model(100).parameters()])
    for i in range(world):
        assert torch.equal(r[i]["params"], ref_params)  # broadcast from rank 0
    singles = [_single_rank_grads(i) for i in range(world)]
    mean_g = (singles[0][0] + singles[1][0]) / 2
    mean_loss = (singles[0][1] + singles[1][1]) / 2
    for i in range(world):
        assert torch.allclose(r[i]["g1"], mean_g, rtol=1e-5, atol=1e-7)
        assert torch.allclose(r[i]["red"], mean_loss, rtol=1e-6)
    assert torch.equal(r[0]["g1"], r[1]["g1"]) and torch.equal(r[0]["g2"], r[1]["g2"])
    halves = [_single_rank_grads(i, 0.5)[0] for i in range(world)]
    assert torch.allclose(r[0]["g2"], (halves[0] + halves[1]) / 2, rtol=1e-5, atol=1e-7)
    for i in range(world):  # skipped reduction: local gradient untouched
        assert torch.allclose(r[i]["g3"], singles[i][0], rtol=1e-5, atol=1e-7)
    assert not torch.allclose(r[0]["g3"], r[1]["g3"])


class _TinyGen(torch.nn.Module):
    def __init__(self):
        super().__init__()
        wn = torch.nn.utils.weight_norm
        self.pre = wn(torch.nn.Conv1d(4, 8, 7, padding=3))
        self.up = wn(torch.nn.ConvTranspose1d(8, 4, 16, 8, padding=4))
        self.post = wn(torch.nn.Conv1d(4, 1, 7, padding=3))

    def forward(self, x):
        return torch.tanh(self.post(torch.nn.functional.leaky_relu(self.up(torch.nn.functional.leaky_relu(self.pre(x))))))


def _naive_allreduce(module, world):
    """The reference's semantics in their plainest form: after a backward, every gradient becomes the mean over ranks."""
    for p in module.parameters():
        if p.grad is not None:
            dist.all_reduce(p.grad)
            p.grad /= world


def _train_flow(rank, world, wrapped, steps=3):
    """train.py:108-129 with stand-in modules: generator step (backward through D into G, only g_optim steps), then
    d_optim.zero_grad() and the discriminator step on the detached audio."""
    from melgan_multi_b200 import distributed as mgd
    torch.manual_seed(50)
    gen, disc = _TinyGen(), _TinyDisc()
    if wrapped:
        mgd.apply_gradient_allreduce(gen)
        mgd.apply_gradient_allreduce(disc)
    g_opt = torch.optim.Adam(gen.parameters(), 2e-4, betas=(0.5, 0.9))
    d_opt = torch.optim.Adam(disc.parameters(), 2e-4, betas=(0.5, 0.9))
    torch.manual_seed(900 + rank)  # every rank trains on its own shard
    losses = []
    for _ in range(steps):
        x, y = torch.randn(2, 4, 32), torch.rand(2, 1, 256) * 2 - 1
        g_opt.zero_grad()
        y_hat = gen(x)
        lr_, fr = disc(y)
        lg, fg = disc(y_hat)
        loss_gen = ((1 - lg) ** 2).mean() + sum((a - b).abs().mean() for a, b in zip(fr, fg))
        loss_gen.backward()
        if not wrapped:
            _naive_allreduce(disc, world)  # the reference reduces these too (and then discards them)
            _naive_allreduce(gen, world)
        g_opt.step()
        d_opt.zero_grad()
        lr_, _ = disc(y)
        lg, _ = disc(y_hat.detach())
        loss_disc = ((1 - lr_) ** 2).mean() + (lg ** 2).mean()
        loss_disc.backward()
        if not wrapped:
            _naive_allreduce(disc, world)
        d_opt.step()
        losses.append((loss_gen.item(), loss_disc.item()))
    flat = torch.cat([p.detach().reshape(-1) for p in list(gen.parameters()) + list(disc.parameters())])
    stats = (dict(gen._grad_reducer.stats), dict(disc._grad_reducer.stats)) if wrapped else None
    return flat, losses, stats


def _flow_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    ref_params, ref_losses, _ = _train_flow(rank, world, wrapped=False)
    our_params, our_losses, stats = _train_flow(rank, world, wrapped=True)
    torch.save({"ref": ref_params, "ours": our_params, "ref_losses": ref_losses, "our_losses": our_losses, "stats": stats},
               os.path.join(out_dir, "flow%d.pt" % rank))
    dist.destroy_process_group()


def test_train_flow_matches_reference_semantics_and_drops_the_wasted_allreduce(tmp_path):
    """Three train.py-shaped steps on 2 ranks: parameters after training equal those of the plain 'all-reduce everything after
    every backward' semantics of the reference, while the discriminator gradients of the generator step -- which
    d_optim.zero_grad() discards -- stop being all-reduced from the second step on (learned, no train.py edit)."""
    world, port = 2, _free_port()
    mp.spawn(_flow_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r = [torch.load(os.path.join(tmp_path, "flow%d.pt" % i)) for i in range(world)]
    assert torch.equal(r[0]["ours"], r[1]["ours"])  # replicas stay in sync
    for i in range(world):
        assert torch.allclose(r[i]["ours"], r[i]["ref"], rtol=1e-5, atol=1e-7), (r[i]["ours"] - r[i]["ref"]).abs().max()
        assert np.allclose(r[i]["our_losses"], r[i]["ref_losses"], rtol=1e-5)
        gs, ds = r[i]["stats"]
        assert gs["passes"] == 3 and gs["allreduce_calls"] == 3 and gs["skipped_bytes"] == 0  # G: one bucket, every step
        # D: 6 backward passes; the first generator step is reduced (nothing learned yet), the two later ones are not
        assert ds["passes"] == 6 and ds["allreduce_calls"] == 4 and ds["lazy_flushes"] == 0
        assert ds["skipped_bytes"] == 2 * ds["allreduce_bytes"] // 4


def _lazy_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from melgan_multi_b200 import distributed as mgd
    torch.set_num_threads(1)
    torch.manual_seed(3)
    model = _TinyDisc()
    mgd.apply_gradient_allreduce(model)
    opt = torch.optim.SGD(model.parameters(), 0.1)
    torch.manual_seed(40 + rank)
    x = torch.randn(2, 1, 256)

    def backward():
        opt.zero_grad()
        (model(x)[0] ** 2).mean().backward()

    backward()            # pass 1: reduced eagerly (nothing known), then ...
    model(x)              # ... dropped: the next forward arrives before any optimizer step -> key learned as "discarded"
    backward()            # pass 2, same key: kept local (lazy) ...
    local = torch.cat([p.grad.reshape(-1) for p in model.parameters()]).clone()
    opt.step()            # ... but an optimizer DOES want it: reduced here, before the update reads it
    flushed = model._grad_reducer.stats["lazy_flushes"]
    reduced = torch.cat([p.grad.reshape(-1) for p in model.parameters()]).clone()
    torch.save({"local": local, "reduced": reduced, "flushed": flushed}, os.path.join(out_dir, "lazy%d.pt" % rank))
    dist.destroy_process_group()


def test_mispredicted_discard_is_reduced_before_the_optimizer_reads_it(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_lazy_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r = [torch.load(os.path.join(tmp_path, "lazy%d.pt" % i)) for i in range(world)]
    assert r[0]["flushed"] == r[1]["flushed"] == 1
    assert not torch.allclose(r[0]["local"], r[1]["local"])
    mean = (r[0]["local"] + r[1]["local"]) / 2
    for i in range(world):
        assert torch.allclose(r[i]["reduced"], mean, rtol=1e-6, atol=1e-8)


def _shard_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from melgan_multi_b200 import synth
    from oracle import torch_port
    torch.set_num_threads(2)
    ws, bs = torch_port.fold_state(synth.generator_state(1234))
    mel = synth.mel_input(4, 6, 11)  # every rank can build the whole batch from the seed ...
    mine = torch.from_numpy(mel[rank::world])  # ... and owns the items rank, rank+world, ...
    y = torch_port.generator_forward(ws, bs, mine)
    gathered = [torch.empty_like(y) for _ in range(world)]
    dist.all_gather(gathered, y)  # test-only gather; the serving path keeps outputs on their rank
    if rank == 0:
        full = torch.empty(4, 1, 6 * 256)
        for r_, g in enumerate(gathered):
            full[r_::world] = g
        np.save(os.path.join(out_dir, "sharded.npy"), full.numpy())
    dist.destroy_process_group()


def test_batch_sharded_inference_equals_unsharded(tmp_path):
    """Generator inference shards by batch with no data-path collective: rank r computes items r::N and the
    union equals the unsharded result bit for bit (batch items are independent)."""
    world, port = 2, _free_port()
    mp.spawn(_shard_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    from melgan_multi_b200 import synth
    from oracle import torch_port
    ws, bs = torch_port.fold_state(synth.generator_state(1234))
    whole = torch_port.generator_forward(ws, bs, torch.from_numpy(synth.mel_input(4, 6, 11))).numpy()
    sharded = np.load(os.path.join(tmp_path, "sharded.npy"))
    np.testing.assert_allclose(sharded, whole, rtol=0, atol=1e-6)


def test_utterance_shard_partitions_time_axis():
    from melgan_multi_b200.distributed import HALO_FRAMES, utterance_shard
    for T in (1, 2, 7, 1000, 1001):
        for world in (1, 2, 3, 8):
            cover = []
            for r in range(world):
                lo, hi, a, b = utterance_shard(T, world, r)
                assert 0 <= a <= lo <= hi <= b <= T
                assert (lo - a == HALO_FRAMES or a == 0) and (b - hi == HALO_FRAMES or b == T)
                cover += list(range(lo, hi))
            assert cover == list(range(T))
            sizes = [utterance_shard(T, world, r)[1] - utterance_shard(T, world, r)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def _utt_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from melgan_multi_b200 import distributed as mgd
    from melgan_multi_b200 import synth
    from oracle import torch_port
    torch.set_num_threads(2)
    ws, bs = torch_port.fold_state(synth.generator_state(1234))
    mel = torch.from_numpy(synth.mel_input(1, 37, 5))  # replicated input: every rank builds it from the seed
    audio = mgd.generate_sharded(lambda m: torch_port.generator_forward(ws, bs, m), mel)
    np.save(os.path.join(out_dir, "utt%d.npy" % rank), audio.numpy())
    dist.destroy_process_group()


def test_time_sharded_utterance_equals_whole(tmp_path):
    """One long utterance split along time over 2 ranks (8-frame halo, no data-path collective, one all_gather of the
    audio): every rank ends up with the whole-utterance result."""
    world, port = 2, _free_port()
    mp.spawn(_utt_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    from melgan_multi_b200 import synth
    from oracle import torch_port
    ws, bs = torch_port.fold_state(synth.generator_state(1234))
    whole = torch_port.generator_forward(ws, bs, torch.from_numpy(synth.mel_input(1, 37, 5))).numpy()
    for r in range(world):
        got = np.load(os.path.join(tmp_path, "utt%d.npy" % r))
        assert got.shape == whole.shape
        assert np.abs(got - whole).max() <= 2e-6 * np.abs(whole).max()
