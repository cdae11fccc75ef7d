"""Multi-GPU plumbing for the roko hot path: one process per GPU, torch.distributed.

Windows are independent (no state crosses the batch dimension in roko/rnn_model.py:46-59), so the
path shards with NO data-path collective: every rank takes a contiguous range of window indices
(contiguous keeps labels aligned with the (contig, position) metadata the stitcher needs --
roko/inference.py:119-124).  Two collectives bracket the job, replacing the reference's dormant
``nn.DataParallel`` (roko/inference.py:12,96-97):

  * ``broadcast_weights``  rank 0's 31 tensors -> every rank, one flat 4.4 MB NCCL broadcast
  * ``gather_labels``      each rank's uint8 labels (90 B / window) -> rank 0, in rank order

Works with the ``nccl`` backend on GPUs and ``gloo`` on CPU tensors (used by the CPU tests).
"""
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world_size):
    """Contiguous, balanced [lo, hi) of rank: the first ``n % world`` ranks get one extra item."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank / world_size")
    base, extra = divmod(int(n_items), world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(n_items, world_size):
    return [shard_range(n_items, r, world_size)[1] - shard_range(n_items, r, world_size)[0]
            for r in range(world_size)]


def broadcast_weights(module, src=0, group=None):
    """Make every rank's parameters equal to ``src``'s with ONE flat broadcast."""
    params = [p for _, p in sorted(module.named_parameters(), key=lambda kv: kv[0])]
    with torch.no_grad():
        flat = torch.cat([p.detach().reshape(-1) for p in params])
        dist.broadcast(flat, src=src, group=group)
        off = 0
        for p in params:
            n = p.numel()
            p.copy_(flat[off:off + n].view_as(p))
            off += n
    return flat.numel() * flat.element_size()


def gather_labels(local_labels, n_total, group=None, dst=0):
    """Concatenate per-rank ``(n_r, 90)`` uint8 labels in rank order.

    Shards may be ragged; ranks pad to the largest shard so a single ``all_gather_into_tensor``
    (NVLink/NVSwitch under NCCL) moves everything, then rank ``dst`` strips the padding.
    Returns the ``(n_total, 90)`` tensor on rank ``dst`` and ``None`` elsewhere.
    """
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = shard_sizes(n_total, world)
    if local_labels.shape[0] != sizes[rank]:
        raise ValueError(f"rank {rank} holds {local_labels.shape[0]} windows, expected {sizes[rank]}")
    width = local_labels.shape[1]
    cap = max(sizes)
    send = local_labels
    if send.shape[0] != cap:
        send = torch.zeros((cap, width), dtype=local_labels.dtype, device=local_labels.device)
        send[:sizes[rank]] = local_labels
    recv = torch.empty((world * cap, width), dtype=local_labels.dtype, device=local_labels.device)
    dist.all_gather_into_tensor(recv, send.contiguous(), group=group)
    if rank != dst:
        return None
    if all(s == cap for s in sizes):
        return recv
    return torch.cat([recv[r * cap:r * cap + sizes[r]] for r in range(world)])


def sum_gradients(module, group=None):
    """Like ``average_gradients`` without the division: for losses already normalised by the GLOBAL batch
    (sum over local positions / global position count), the all-reduced SUM is the gradient of the global mean --
    also when shards are ragged or empty.  Returns the number of bytes reduced."""
    return average_gradients(module, group=group, divide=False)


def average_gradients(module, group=None, divide=True):
    """Data-parallel training step, reference roko/train.py:46-53 over several ranks: every rank ran
    forward/backward on its share of the batch; one flat all-reduce (4.4 MB for this network, NCCL
    over NVLink on GPUs) leaves the mean gradient in every ``param.grad``.  Parameters without a
    gradient on this rank contribute zeros.  Returns the number of bytes reduced."""
    params = [p for _, p in sorted(module.named_parameters(), key=lambda kv: kv[0]) if p.requires_grad]
    if not params:
        return 0
    world = dist.get_world_size(group)
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).detach().reshape(-1).to(torch.float32)
                      for p in params])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if divide:
        flat.div_(world)
    off = 0
    for p in params:
        n = p.numel()
        g = flat[off:off + n].view_as(p).to(p.dtype)
        if p.grad is None:
            p.grad = g.clone()
        else:
            p.grad.copy_(g)
        off += n
    return flat.numel() * flat.element_size()
