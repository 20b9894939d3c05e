"""The 4-stream ensemble (joint / bone / joint-motion / bone-motion) as one sharded, batched inference call.

Reference behaviour this replaces:
  * ``inference_pipeline.py:284-309`` (``derive_modalities``), ``data_gen/gen_bone_data.py:5-58`` and
    ``data_gen/gen_motion_data.py:18-34``: the bone and motion streams are numpy loops on the host, and every stream is
    uploaded separately.  Here only the joint batch crosses PCIe; each rank derives its stream on the device
    (``ops.input_stream`` -> ``sgcn_input_stream``), fused with the model's input BatchNorm and the layout change.
  * ``inference_pipeline.py:342-366`` (``run_ensemble_inference``) / ``ensemble.py:18-27``: batch-1 loops with a
    ``.cpu()`` sync per window and stream, logits summed as ``sum_k alpha_k * logits_k`` on the host, softmax after the
    sum.  Here all windows form one batch, every rank scales its logits by its stream weight, and ONE all-reduce (sum)
    of the ``(N, num_class)`` logits over NVLink *is* the ensemble; softmax / argmax run on the reduced logits.

Placement: ``world`` ranks, 4 streams.  world = 1: all four streams on the one GPU;
world = 2: two streams per rank; world = 4: one stream per rank; world = 8: two ranks per stream, each taking half of
the batch.  The models are independent -- there is no exchange inside them, only the final logits reduction.
"""
import torch

from . import ops

MODALITIES = ("joint", "bone", "joint_motion", "bone_motion")        # inference_pipeline.py:25
ENSEMBLE_WEIGHTS_DEFAULT = (0.6, 0.6, 0.4, 0.4)                      # ensemble.py:18, inference_pipeline.py:24

# (joint, parent) pairs, 1-based, data_gen/gen_bone_data.py:5-31 (identical for all four NTU benchmarks)
_NTU_PAIRS_1BASED = ((1, 2), (2, 21), (3, 21), (4, 3), (5, 21), (6, 5), (7, 6), (8, 7), (9, 21), (10, 9), (11, 10),
                     (12, 11), (13, 1), (14, 13), (15, 14), (16, 15), (17, 1), (18, 17), (19, 18), (20, 19), (22, 23),
                     (21, 21), (23, 8), (24, 25), (25, 12))
# parent of landmark v, 0-based, inference_pipeline.py:16-22 (root = nose, self-referencing)
_MEDIAPIPE_PARENTS = (0, 0, 1, 2, 0, 4, 5, 3, 6, 0, 9, 0, 11, 11, 12, 13, 14, 15, 16, 15, 16, 15, 16, 11, 12, 23, 24, 25,
                      26, 27, 28, 27, 28)


def bone_parents(num_point):
    """0-based parent joint of every joint for the skeletons the reference ships (25 = NTU, 33 = MediaPipe)."""
    if num_point == 25:
        parents = [0] * 25
        for v, p in _NTU_PAIRS_1BASED:
            parents[v - 1] = p - 1
        return tuple(parents)
    if num_point == 33:
        return _MEDIAPIPE_PARENTS
    raise ValueError(f"no bone table for {num_point} joints (the reference defines 25 = NTU and 33 = MediaPipe)")


def stream_flags(modality):
    """(bone step?, motion step?) of a stream name"""
    if modality not in MODALITIES:
        raise ValueError(f"unknown modality {modality!r}; expected one of {MODALITIES}")
    return modality.startswith("bone"), modality.endswith("motion")


def derive_modality(joint, modality, parents=None):
    """One stream of a joint batch ``(N, C, T, V, M)`` (CUDA fp32), same layout; 'joint' returns the input itself."""
    bone, motion = stream_flags(modality)
    if not bone and not motion:
        return joint
    par = None
    if bone:
        par = torch.as_tensor(parents if parents is not None else bone_parents(joint.shape[3]), dtype=torch.int32,
                              device=joint.device)
    return ops.input_stream(joint.contiguous(), parent=par, motion=motion)


def derive_modalities(joint, parents=None):
    """dict of all four streams (the reference's ``derive_modalities``, batched and on the device)"""
    return {m: derive_modality(joint, m, parents) for m in MODALITIES}


def placement(world_size, rank):
    """Which (stream index, batch shard, number of shards) pairs this rank computes.

    world 1 -> four streams, whole batch; 2 -> two streams each; 4 -> one stream each; 8 (any multiple of 4) -> ranks
    {2k, 2k+1, ...} share stream k and split the batch evenly.  Other world sizes are rejected: the ensemble has
    exactly four independent models."""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError("bad world size / rank")
    if world_size in (1, 2, 4):
        per = 4 // world_size
        return [(rank * per + k, 0, 1) for k in range(per)]
    if world_size % 4 == 0:
        shards = world_size // 4
        return [(rank // shards, rank % shards, shards)]
    raise ValueError(f"4 streams cannot be placed on {world_size} ranks (use 1, 2, 4 or a multiple of 4)")


def shard_bounds(n, shard, shards):
    """rows [lo, hi) of a batch of n samples taken by `shard` of `shards` (first shards get the remainder)"""
    base, rem = divmod(n, shards)
    lo = shard * base + min(shard, rem)
    return lo, lo + base + (1 if shard < rem else 0)


class StreamEnsemble:
    """Sharded 4-stream ensemble inference.

    models: dict stream name -> model for the streams THIS rank owns (``placement``); each model maps a stream batch
    ``(n, C, T, V, M)`` to logits ``(n, num_class)``.  ``stream_fn(joint, modality)`` derives a stream on the rank's
    device (default: the CUDA kernel; tests of the host logic inject a CPU function).  ``group``: the
    ``torch.distributed`` process group, or None for a single process.
    """

    def __init__(self, models, num_class, weights=ENSEMBLE_WEIGHTS_DEFAULT, world_size=1, rank=0, group=None,
                 stream_fn=None):
        if len(weights) != 4:
            raise ValueError("the ensemble has four stream weights")
        self.plan = placement(world_size, rank)
        missing = [MODALITIES[k] for k, _, _ in self.plan if MODALITIES[k] not in models]
        if missing:
            raise ValueError(f"rank {rank} of {world_size} needs models for {missing}")
        self.models, self.num_class, self.weights = models, num_class, tuple(float(w) for w in weights)
        self.world_size, self.rank, self.group = world_size, rank, group
        self.stream_fn = stream_fn if stream_fn is not None else derive_modality

    @torch.no_grad()
    def logits(self, joint):
        """joint: the FULL batch (N, C, T, V, M) on this rank's device (every rank receives the same joint batch).
        Returns the ensemble logits (N, num_class) = sum_k alpha_k * model_k(stream_k), identical on every rank."""
        N = joint.shape[0]
        total = torch.zeros((N, self.num_class), device=joint.device, dtype=torch.float32)
        for k, shard, shards in self.plan:
            lo, hi = shard_bounds(N, shard, shards)
            if hi <= lo:
                continue
            name = MODALITIES[k]
            out = self.models[name](self.stream_fn(joint[lo:hi], name))
            total[lo:hi].add_(out.float(), alpha=self.weights[k])
        if self.world_size > 1:
            import torch.distributed as dist
            dist.all_reduce(total, op=dist.ReduceOp.SUM, group=self.group)
        return total

    def scores(self, joint):
        """softmax over the ensemble logits (inference_pipeline.py:358-360 takes class 1 of this as the fall score)"""
        return torch.softmax(self.logits(joint), dim=1)
