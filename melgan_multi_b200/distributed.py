"""Drop-in for the three functions train.py uses from the reference's distributed.py
(`init_distributed`, `apply_gradient_allreduce`, `reduce_tensor`; /root/reference/distributed.py:37-142).

Same observable behaviour -- parameters broadcast from rank 0 at wrap time, every parameter's gradient replaced by the
across-rank mean before an optimizer consumes it, NCCL underneath through torch.distributed -- with the data path
reorganised for NVLink-class fabrics, where launch count and exposed latency, not link bandwidth, are the cost:

  * one persistent flat fp32 gradient buffer per module: every ``param.grad`` is a *view* into it, so the all-reduce runs
    in place with no ``torch.cat`` and no copy-back (the reference flattens and un-flattens 18 MB / 68 MB per call,
    distributed.py:125-129).  Gradients that autograd allocated elsewhere (``zero_grad(set_to_none=True)``, the default
    of torch 2.x) are adopted bucket by bucket with one multi-tensor copy;
  * **overlap with backward**: the buffer is cut into buckets (reverse registration order, <= 24 MiB: one per
    Discriminator, one for the Generator); a bucket's all-reduce is launched asynchronously from the gradient hooks the
    moment its last gradient exists, while autograd is still computing the other buckets (the reference reduces
    everything after backward has finished, distributed.py:131-135); the end-of-backward callback only waits;
  * **no wasted collective, without touching train.py**: in the generator step ``loss_gen.backward()`` also produces
    discriminator gradients that ``d_optim.zero_grad()`` throws away (train.py:117,120), yet the reference all-reduces
    them (67.7 MB of the 153.5 MB per step).  The wrapper watches what happens to each backward's gradients -- consumed
    by an ``optimizer.step`` (global optimizer pre-hook) or dropped by the module's next forward -- keyed by the set of
    wrapped modules that ran forward before that backward ({G, D} in the generator step, {D} in the discriminator step).
    Once a key has been seen to end in a discard, that key's gradients are kept local and *lazily* reduced only if an
    optimizer does ask for them after all, so results never differ from the reference's (MG_DDP_DEDUP=0 turns it off);
  * one flat buffer for the start-up broadcast instead of 90 / 63 tiny broadcasts (distributed.py:100-103).

Batches shard naturally across ranks (DistributedSampler, train.py:71); there is no other collective.
"""
import os

import torch
import torch.distributed as dist
from torch.autograd import Variable

BUCKET_BYTES = 24 << 20


def reduce_tensor(tensor, num_gpus):
    """Mean of a (scalar) tensor across ranks, for logging (reference distributed.py:37-41)."""
    rt = tensor.detach().clone()
    dist.all_reduce(rt, op=dist.ReduceOp.SUM)
    rt /= num_gpus
    return rt


def init_distributed(rank, num_gpus, group_name, dist_backend, dist_url):
    """One process per GPU, TCP rendezvous (reference distributed.py:43-53)."""
    assert torch.cuda.is_available(), "Distributed mode requires CUDA."
    torch.cuda.set_device(rank % torch.cuda.device_count())
    dist.init_process_group(dist_backend, init_method=dist_url, world_size=num_gpus, rank=rank,
                            group_name=group_name)


# ---- process-wide bookkeeping of the wrapped modules -------------------------------------------------------------------
_REDUCERS = []          # every _GradReducer of this process
_FORWARD_SET = set()    # ids of the reducers whose module ran forward since the last backward pass ended
_PASS = {"key": None, "open": []}  # key of the backward pass in flight (frozenset of reducer ids), None between passes; "open":
                                   # [reducer, bytes] of buckets launched in this pass after which no gradient has appeared yet
_OPT_HOOK = {"handle": None}


def _end_pass():
    _PASS["key"] = None
    _PASS["open"] = []
    _FORWARD_SET.clear()


def _optimizer_pre_step(optimizer, args, kwargs):
    """Global optimizer hook: the gradients of every wrapped module this optimizer owns are about to be consumed."""
    owned = getattr(optimizer, "_mg_reducers", None)
    if owned is None or owned[0] != len(_REDUCERS):
        ids = {id(p) for g in optimizer.param_groups for p in g["params"]}
        owned = (len(_REDUCERS), [r for r in _REDUCERS if any(id(p) in ids for p in r.params)])
        optimizer._mg_reducers = owned
    for r in owned[1]:
        r.consume()


class _GradReducer:
    """Flat gradient buffer, buckets and reduction state of one wrapped module."""

    def __init__(self, module):
        self.module = module
        self.world = dist.get_world_size()
        self.params = [p for p in module.parameters() if p.requires_grad]
        self.dedup = os.environ.get("MG_DDP_DEDUP", "1") != "0"
        self.stats = {"allreduce_calls": 0, "allreduce_bytes": 0, "skipped_bytes": 0, "lazy_flushes": 0, "passes": 0,
                      "bytes_launched_with_backward_left": 0}
        self.history = {}            # pass key -> "consumed" | "discarded"
        self.unconsumed_key = None   # key of the last backward whose gradients nobody has consumed or dropped yet
        self.pending = False         # ... and those gradients are still local (lazy mode): reduce them if consumed
        self.state = 0               # 0: no backward pass in flight, 1: eager pass, 2: lazy pass
        self.dry = False             # measurement aid (bench.py): do everything except the collectives themselves
        self.skip_once = False
        if not self.params:
            self.flat = None
            return
        ref = self.params[0]
        if any(p.dtype != ref.dtype or p.device != ref.device for p in self.params):
            raise ValueError("apply_gradient_allreduce: parameters of one module must share dtype and device")
        n = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(n, dtype=ref.dtype, device=ref.device)
        # the flat buffer is laid out in REVERSE registration order: gradients appear roughly in that order during backward,
        # so bucket 0 (the last layers) fills -- and starts its all-reduce -- first
        self.views, self.bucket_of, self.buckets = [None] * len(self.params), [0] * len(self.params), []
        off, b_start, b_members = 0, 0, []
        for i in reversed(range(len(self.params))):
            p = self.params[i]
            nbytes = p.numel() * p.element_size()
            if b_members and (off - b_start) * ref.element_size() + nbytes > BUCKET_BYTES:
                self.buckets.append((b_start, off, b_members))
                b_start, b_members = off, []
            self.views[i] = self.flat.narrow(0, off, p.numel()).view_as(p)
            self.bucket_of[i] = len(self.buckets)
            b_members.append(i)
            off += p.numel()
        self.buckets.append((b_start, off, b_members))
        self.bucket_len = [len(m) for _, _, m in self.buckets]
        self._reset_pass()

    def _reset_pass(self):
        nb = len(self.buckets)
        self.ready, self.launched = [0] * nb, [False] * nb
        self.works = []

    # -- forward / backward events ------------------------------------------------------------------------------------
    def on_forward(self):
        self.module.needs_reduction = True
        _FORWARD_SET.add(id(self))
        if self.unconsumed_key is not None:  # the previous backward's gradients were never asked for: remember that
            self.history[self.unconsumed_key] = "discarded"
            if self.pending:
                self.stats["skipped_bytes"] += self.flat.numel() * self.flat.element_size()
            self.unconsumed_key, self.pending = None, False

    def _begin_pass(self):
        if _PASS["key"] is None:
            _PASS["key"] = frozenset(_FORWARD_SET)
            Variable._execution_engine.queue_callback(_end_pass)
        self.key = _PASS["key"]
        self.stats["passes"] += 1
        self._reset_pass()
        lazy = self.skip_once or (self.dedup and self.history.get(self.key) == "discarded")
        self.state = 2 if lazy else 1
        Variable._execution_engine.queue_callback(self._finish_pass)

    def on_grad(self, i):
        """post-accumulate-grad hook of parameter i (runs ~150 times per backward on a host-bound training step: keep it
        to a few bytecodes).  state 0: no pass in flight, 1: eager pass (count, launch full buckets), 2: lazy pass."""
        st = self.state
        if _PASS["open"]:  # backward produced another gradient after those buckets were launched: their all-reduce had
            for r, n in _PASS["open"]:  # compute to overlap with (statistics only)
                r.stats["bytes_launched_with_backward_left"] += n
            _PASS["open"] = []
        if st == 2:
            return
        if st == 0:
            if not self.module.needs_reduction:
                return  # (reference: nothing is reduced unless the module ran forward since the last reduction)
            self._begin_pass()
            if self.state == 2:
                return
        b = self.bucket_of[i]
        n = self.ready[b] + 1
        self.ready[b] = n
        if n == self.bucket_len[b]:
            self._launch(b)

    def _adopt(self, members):
        """Make the gradients of parameters `members` views of the flat buffer (one multi-tensor copy for those autograd
        allocated elsewhere: zero_grad(set_to_none=True) makes that all of them, every step)."""
        dsts, srcs = [], []
        for i in members:
            p, v = self.params[i], self.views[i]
            g = p.grad
            if g is not None and g.data_ptr() != v.data_ptr():
                dsts.append(v)
                srcs.append(g.detach())
                p.grad = v
        if dsts:
            torch._foreach_copy_(dsts, srcs)

    def _launch(self, b):
        if self.launched[b]:
            return
        self.launched[b] = True
        start, end, members = self.buckets[b]
        self._adopt(members)
        chunk = self.flat.narrow(0, start, end - start)
        if not self.dry:
            self.works.append(dist.all_reduce(chunk, async_op=True))
        self.stats["allreduce_calls"] += 1
        self.stats["allreduce_bytes"] += chunk.numel() * chunk.element_size()
        if self.state == 1:
            _PASS["open"].append((self, chunk.numel() * chunk.element_size()))

    def _finish_pass(self):
        """End-of-backward callback (the reference does ALL its work here, distributed.py:105-129): buckets whose
        parameters did not all receive a gradient are launched now, then the in-flight all-reduces are awaited."""
        st, self.state = self.state, 0
        if st == 0:
            return
        # (buckets launched from here on are launched by the end-of-backward callback: state is already 0)
        self.module.needs_reduction = False
        if st == 2:  # gradients stay where autograd put them, un-reduced
            if self.skip_once:  # skip_next_reduction(): the caller promised to throw these gradients away
                self.skip_once = False
                self.unconsumed_key, self.pending = None, False
            else:
                self.unconsumed_key, self.pending = self.key, True
            return
        for b in range(len(self.buckets)):
            self._launch(b)
        for w in self.works:
            w.wait()
        self.works = []
        self.flat.div_(self.world)
        self.unconsumed_key, self.pending = self.key, False

    def consume(self):
        """An optimizer is about to read this module's gradients."""
        if self.unconsumed_key is None:
            return
        if self.pending:  # predicted "discarded", but they are wanted after all: reduce now (blocking, exact)
            self._adopt(range(len(self.params)))
            if not self.dry:
                dist.all_reduce(self.flat)
            self.flat.div_(self.world)
            self.stats["lazy_flushes"] += 1
            self.stats["allreduce_calls"] += 1
            self.stats["allreduce_bytes"] += self.flat.numel() * self.flat.element_size()
        self.history[self.unconsumed_key] = "consumed"
        self.unconsumed_key, self.pending = None, False


def skip_next_reduction(module):
    """Explicit form of what the wrapper learns by itself: the next backward's gradients of `module` will be thrown
    away by the caller (train.py:117 followed by d_optim.zero_grad() at :120), do not all-reduce them."""
    module._grad_reducer.skip_once = True


def apply_gradient_allreduce(module):
    """Same contract as the reference's apply_gradient_allreduce (distributed.py:90-142)."""
    # start-up sync: one flat broadcast per dtype instead of one per tensor
    tensors = [t for t in module.state_dict().values() if torch.is_tensor(t)]
    by_dtype = {}
    for t in tensors:
        by_dtype.setdefault(t.dtype, []).append(t)
    for dtype, ts in by_dtype.items():
        flat = torch.cat([t.detach().reshape(-1) for t in ts])
        dist.broadcast(flat, 0)
        off = 0
        for t in ts:
            t.detach().copy_(flat.narrow(0, off, t.numel()).view_as(t))
            off += t.numel()

    red = _GradReducer(module)
    _REDUCERS.append(red)
    module._grad_reducer = red
    module._flat_grads = red  # (name kept from round 1)
    module.needs_reduction = False
    if _OPT_HOOK["handle"] is None:
        from torch.optim.optimizer import register_optimizer_step_pre_hook
        _OPT_HOOK["handle"] = register_optimizer_step_pre_hook(_optimizer_pre_step)

    on_grad = red.on_grad
    for i, p in enumerate(red.params):
        p.register_post_accumulate_grad_hook(lambda _p, i=i, f=on_grad: f(i))

    module.register_forward_hook(lambda mod, inputs, output: red.on_forward())
    return module


# ------------------------------------------------------------------------------------------------------------------
# Long-utterance inference sharded over ranks along TIME (BASELINE config 5).  Not in the reference:
# its inference script runs one utterance on one device.  The generator's receptive field is +-7 mel frames, so a rank
# that owns frames [lo, hi) reads [lo - 8, hi + 8) (clipped to the utterance) and keeps the middle: no data-path
# collective, the mel (320 B/frame) is simply replicated; only the optional gather of the audio talks to other ranks.
HALO_FRAMES = 8


def utterance_shard(T, world_size, rank, halo=HALO_FRAMES):
    """Frames rank `rank` owns and reads: (lo, hi, a, b) with owned [lo, hi) and read window [a, b).  Contiguous, balanced
    to one frame, empty (lo == hi) for ranks beyond T."""
    if T < 1 or world_size < 1 or not 0 <= rank < world_size:
        raise ValueError("utterance_shard(T=%r, world_size=%r, rank=%r)" % (T, world_size, rank))
    base, extra = divmod(T, world_size)
    lo = rank * base + min(rank, extra)
    hi = lo + base + (rank < extra)
    return lo, hi, max(0, lo - halo), min(T, hi + halo)


def generate_sharded(generator, mel, rank=None, world_size=None, gather=True, group=None):
    """mel [1, 80, T] (replicated on every rank) -> this rank's audio slice [1, 1, 256 * (hi - lo)]; with gather=True every
    rank returns the whole utterance [1, 1, 256 * T] (one all_gather of equal-size padded slices over NCCL)."""
    if rank is None:
        rank = dist.get_rank(group)
    if world_size is None:
        world_size = dist.get_world_size(group)
    T = mel.shape[2]
    lo, hi, a, b = utterance_shard(T, world_size, rank)
    if hi > lo:
        with torch.no_grad():
            part = generator(mel[:, :, a:b].contiguous())[:, :, (lo - a) * 256:(hi - a) * 256]
    else:
        part = mel.new_zeros((1, 1, 0))
    if not gather or world_size == 1:
        return part
    width = 256 * ((T + world_size - 1) // world_size)
    padded = mel.new_zeros((1, 1, width))
    padded[:, :, :part.shape[2]] = part
    parts = [torch.empty_like(padded) for _ in range(world_size)]
    dist.all_gather(parts, padded, group=group)
    sizes = [utterance_shard(T, world_size, r)[1] - utterance_shard(T, world_size, r)[0] for r in range(world_size)]
    return torch.cat([p[:, :, :256 * n] for p, n in zip(parts, sizes)], dim=2)
