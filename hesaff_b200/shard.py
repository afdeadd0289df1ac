"""Image-parallel sharding of a batch over ranks: images are independent, so each rank runs
the whole path on a contiguous block of the batch; the only exchange is ONE all-gather of the per-image
{detected, described} counts, from which every rank derives the global keypoint offsets.  Host logic only
(works with the gloo backend on CPU and nccl on GPU); no pixel or descriptor data crosses NVLink."""
import numpy as np


def partition(n_images, world_size):
    """Contiguous block partition: rank r owns images [start[r], start[r+1])."""
    base, rem = divmod(n_images, world_size)
    sizes = np.array([base + (1 if r < rem else 0) for r in range(world_size)], np.int64)
    starts = np.concatenate([[0], np.cumsum(sizes)])
    return starts


def all_gather_counts(dist, torch, local_counts, device):
    """local_counts: int32 [n_local, 2] (detected, described). Returns the [n_global, 2] array in image order.
    Blocks of unequal size are padded to the largest block for the collective."""
    world = dist.get_world_size()
    n_local = torch.tensor([local_counts.shape[0]], dtype=torch.int64, device=device)
    sizes = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(sizes, n_local)
    sizes = [int(s.item()) for s in sizes]
    m = max(sizes) if sizes else 0
    buf = torch.zeros((m, 2), dtype=torch.int32, device=device)
    if local_counts.shape[0]:
        buf[:local_counts.shape[0]] = torch.as_tensor(np.ascontiguousarray(local_counts, np.int32)).to(device)
    out = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    return np.concatenate([o[:s].cpu().numpy() for o, s in zip(out, sizes)], 0)


def global_offsets(global_counts):
    """Exclusive prefix sum of the described counts: image i's Keypoint records start at offsets[i] in the
    concatenated output of all ranks."""
    d = np.asarray(global_counts)[:, 1].astype(np.int64)
    return np.concatenate([[0], np.cumsum(d)])
