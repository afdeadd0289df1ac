"""Image-parallel sharding of a batch over ranks: images are independent, so each rank runs
the whole path on a contiguous block of the batch; the only exchange is ONE all-gather of the per-image
{detected, described} counts, from which every rank derives the global keypoint offsets.  Host logic only
(works with the gloo backend on CPU and nccl on GPU); no pixel or descriptor data crosses NVLink on the path itself.

`all_gather_keypoints` is the step AFTER the path: a variable-size all-gather of the 164-byte
Keypoint records, laid out by the offsets the counts give, for a consumer that wants every rank to hold the whole
batch's keypoints (retrieval / matching).  With nccl and device tensors the records go GPU to GPU over NVLink."""
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


RECORD_BYTES = 164   # sizeof(hesaff_keypoint), hesaff.cpp:41-48


class _DevicePtr:
    """Zero-copy view of device memory for torch.as_tensor (CUDA array interface)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False), "version": 3}


def device_records(torch, det, device):
    """The detector's result records as a [n, 164] uint8 CUDA tensor that aliases the library's buffer (valid until
    the next detect call on this detector)."""
    n = det.total()
    if n == 0:
        return torch.empty((0, RECORD_BYTES), dtype=torch.uint8, device=device)
    return torch.as_tensor(_DevicePtr(det.keys_device_ptr(), n * RECORD_BYTES), device=device).view(n, RECORD_BYTES)


def all_gather_keypoints(dist, torch, local_records, global_counts, starts, device):
    """Variable-size all-gather of Keypoint records.

    local_records : [n_local_described, 164] uint8 tensor on `device` (or a numpy KEYPOINT_DTYPE / uint8 array), the
                    records of this rank's images in image order
    global_counts : [n_images, 2] from all_gather_counts;  starts : partition(n_images, world)
    Returns a [n_global_described, 164] uint8 tensor on `device` in global image order: image i's records are rows
    offsets[i]:offsets[i+1] with offsets = global_offsets(global_counts).  One collective; blocks are padded to the
    largest rank block (blocks of a batch of same-size images differ by a few percent)."""
    world, rank = dist.get_world_size(), dist.get_rank()
    off = global_offsets(global_counts)
    per_rank = [int(off[starts[r + 1]] - off[starts[r]]) for r in range(world)]
    if not hasattr(local_records, "data_ptr"):
        a = np.ascontiguousarray(local_records)
        local_records = torch.from_numpy(a.view(np.uint8).reshape(-1, RECORD_BYTES)).to(device)
    local_records = local_records.reshape(-1, RECORD_BYTES)
    if local_records.shape[0] != per_rank[rank]:
        raise ValueError("rank %d holds %d records, the gathered counts say %d" % (rank, local_records.shape[0], per_rank[rank]))
    m = max(per_rank) if per_rank else 0
    out = torch.empty((int(off[-1]), RECORD_BYTES), dtype=torch.uint8, device=device)
    if m == 0:
        return out
    send = torch.zeros((m, RECORD_BYTES), dtype=torch.uint8, device=device)
    send[:per_rank[rank]] = local_records
    recv = torch.empty((world, m, RECORD_BYTES), dtype=torch.uint8, device=device)
    dist.all_gather_into_tensor(recv.view(-1), send.view(-1))
    pos = 0
    for r in range(world):
        out[pos:pos + per_rank[r]] = recv[r, :per_rank[r]]
        pos += per_rank[r]
    return out


def match_descriptors(torch, query, database, device):
    """The consumer of the gathered records (README:49-53): nearest neighbour of every query record among the database
    records by squared L2 distance over the 128 descriptor bytes, on the GPU (hesaff_match_descriptors).
    query, database : [n, 164] uint8 CUDA tensors (e.g. device_records(...) and all_gather_keypoints(...)).
    Returns (best_index int32 [nq], best_dist2 int32 [nq], second_dist2 int32 [nq]); -1 everywhere if the database is empty."""
    import ctypes as C
    from . import api
    q = query.reshape(-1, RECORD_BYTES).contiguous()
    d = database.reshape(-1, RECORD_BYTES).contiguous()
    nq, nd = q.shape[0], d.shape[0]
    idx = torch.empty(nq, dtype=torch.int32, device=device)
    d1 = torch.empty(nq, dtype=torch.int32, device=device)
    d2 = torch.empty(nq, dtype=torch.int32, device=device)
    if nq == 0:
        return idx, d1, d2
    dev = torch.device(device)
    torch.cuda.current_stream(dev).synchronize()      # the records may still be in flight on torch's stream
    f = api.lib().hesaff_match_descriptors
    f.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    api._check(f(dev.index or 0, q.data_ptr(), nq, d.data_ptr() if nd else None, nd, idx.data_ptr(), d1.data_ptr(), d2.data_ptr(), None))
    return idx, d1, d2
