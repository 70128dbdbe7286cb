"""Scan-parallel sharding of the batched IESKF update across ranks.

Units (scan pair + prior) are independent, so the only exchange of the path is the pose gather at the end:
each rank owns a contiguous range of scan indices and contributes its fixed 64-byte ``lins_scan_result`` records
to one ``all_gather``.  ``torch.distributed`` is plumbing only (NCCL on GPUs, gloo in the CPU tests).
"""
import numpy as np

from .ctypes_defs import SCAN_RESULT_DTYPE


def shard_range(n_total, rank, world):
    """Contiguous, balanced ranges: the first n_total % world ranks get one extra unit."""
    base, extra = divmod(n_total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_records(local_records, n_total, world, dist=None, device=None):
    """all_gather of the per-scan records.  `local_records`: numpy SCAN_RESULT_DTYPE array (this rank's units, with
    GLOBAL scan ids) or a uint8 torch tensor of 64 * n_local bytes already on `device`.  Returns the n_total
    records ordered by scan id (numpy)."""
    import torch

    if dist is None or world == 1:
        rec = local_records if isinstance(local_records, np.ndarray) else np.frombuffer(local_records.cpu().numpy().tobytes(), dtype=SCAN_RESULT_DTYPE)
        return np.sort(rec, order="scan_id")
    per = max(shard_range(n_total, r, world)[1] - shard_range(n_total, r, world)[0] for r in range(world))
    if isinstance(local_records, np.ndarray):
        t = torch.from_numpy(np.frombuffer(local_records.tobytes(), dtype=np.uint8).copy())
        if device is not None:
            t = t.to(device)
    else:
        t = local_records
    pad = torch.full((per * 64,), 255, dtype=torch.uint8, device=t.device)  # scan_id = -1 marks padding
    pad[: t.numel()] = t
    out = torch.empty(world * per * 64, dtype=torch.uint8, device=t.device)
    dist.all_gather_into_tensor(out, pad)
    rec = np.frombuffer(out.cpu().numpy().tobytes(), dtype=SCAN_RESULT_DTYPE)
    rec = rec[rec["scan_id"] >= 0]
    rec = np.sort(rec, order="scan_id")
    assert len(rec) == n_total, (len(rec), n_total)
    return rec


def run_sharded(batch, rank, world, process, dist=None, device=None):
    """Process this rank's contiguous share of `batch` with `process(sub_batch) -> SCAN_RESULT_DTYPE records`
    (local ids 0..n_local-1) and gather everyone's poses.  Returns the full, id-ordered record array."""
    lo, hi = shard_range(batch.n, rank, world)
    sub = batch.subset(range(lo, hi))
    rec = process(sub).copy() if hi > lo else np.zeros(0, dtype=SCAN_RESULT_DTYPE)
    rec["scan_id"] += lo
    return gather_records(rec, batch.n, world, dist, device)
