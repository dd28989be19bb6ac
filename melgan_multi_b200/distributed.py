"""Drop-in for the three functions train.py uses from the reference's distributed.py
(`init_distributed`, `apply_gradient_allreduce`, `reduce_tensor`; /root/reference/distributed.py:37-142).

Same observable behaviour -- parameters broadcast from rank 0 at wrap time, every parameter's gradient
replaced by the across-rank mean after each backward, NCCL underneath through torch.distributed -- with the
data path reorganised for NVLink-class fabrics where launch count, not link bandwidth, is the cost:

  * one persistent flat fp32 gradient buffer per module: every ``param.grad`` is a *view* into it, so the
    all-reduce runs in place on one tensor with no ``torch.cat`` and no copy-back (the reference flattens and
    un-flattens 18 MB / 68 MB per call, distributed.py:125-129);
  * one flat buffer for the start-up broadcast instead of 90 / 63 tiny broadcasts (distributed.py:100-103);
  * the all-reduce is skipped for a module whose gradients were produced by a backward that its owner is
    about to discard: ``skip_next_reduction(module)`` lets the training loop drop the wasted 67.7 MB
    discriminator all-reduce of the generator step without changing results.

Batches shard naturally across ranks (DistributedSampler, train.py:71); there is no other collective.
"""
import torch
import torch.distributed as dist
from torch.autograd import Variable


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


class _FlatGrads:
    """Owns the flat gradient buffer of one module and re-points ``param.grad`` at views of it."""

    def __init__(self, module):
        self.params = [p for p in module.parameters() if p.requires_grad]
        if not self.params:
            self.flat = None
            return
        ref = self.params[0]
        n = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(n, dtype=ref.dtype, device=ref.device)
        self.views = []
        off = 0
        for p in self.params:
            self.views.append(self.flat.narrow(0, off, p.numel()).view_as(p))
            off += p.numel()

    def adopt(self):
        """Make every existing .grad a view of the flat buffer (copying a foreign grad in once).  Returns
        False if some parameter has no gradient yet (then nothing is reduced for it, like the reference)."""
        complete = True
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                complete = False
                v.zero_()
            elif p.grad.data_ptr() != v.data_ptr():
                v.copy_(p.grad)
                p.grad = v
        return complete


def skip_next_reduction(module):
    """The next backward's gradients of `module` will be thrown away by the caller (train.py:117 followed by
    d_optim.zero_grad() at :120): do not all-reduce them."""
    module._skip_reduction_once = True


def apply_gradient_allreduce(module):
    """Same contract as the reference's apply_gradient_allreduce (distributed.py:90-142)."""
    world = dist.get_world_size()

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

    state = _FlatGrads(module)
    module._flat_grads = state
    module.needs_reduction = False
    module._skip_reduction_once = False

    def allreduce_params():
        if not module.needs_reduction:
            return
        module.needs_reduction = False
        if module._skip_reduction_once:
            module._skip_reduction_once = False
            return
        if state.flat is None:
            return
        state.adopt()
        dist.all_reduce(state.flat)
        state.flat /= world

    def allreduce_hook(*unused):
        Variable._execution_engine.queue_callback(allreduce_params)

    for p in state.params:
        p.register_hook(allreduce_hook)

    def set_needs_reduction(self, inputs, output):
        self.needs_reduction = True

    module.register_forward_hook(set_needs_reduction)
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
