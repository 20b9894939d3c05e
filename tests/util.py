"""Shared helpers of the parity tests."""
import numpy as np
import torch

from oracle import model_ref, shift_torch


def rel_err(a, b):
    """max |a - b| / max |b|   (both moved to CPU fp64)"""
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    denom = max(b.abs().max().item(), 1e-30)
    return (a - b).abs().max().item() / denom


def rel_l2(a, b):
    """||a - b||_2 / ||b||_2 -- the metric for TF32-vs-exact GRADIENTS: TF32 rounding flips a ~1e-4 fraction of
    ReLU masks, which changes those gradient entries by O(1) (max-norm meaningless) but the vector by O(1e-2)."""
    a = torch.as_tensor(a).detach().double().cpu().reshape(-1)
    b = torch.as_tensor(b).detach().double().cpu().reshape(-1)
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def fill_pair(cuda_module, ref_module, prefix=""):
    """identical deterministic non-degenerate weights in the CUDA module (fp32) and the oracle (fp64)"""
    model_ref.fill_module_(ref_module, prefix)
    model_ref.fill_module_(cuda_module, prefix)
    return cuda_module, ref_module


class raw_pos_log:
    """context manager: collect the oracle's raw (pre-K5) position sums, keyed by parameter name"""

    def __init__(self, ref_module):
        self.ref = ref_module
        self.raw = {}

    def __enter__(self):
        shift_torch.RAW_POS_LOG = {}
        return self

    def __exit__(self, *exc):
        log = shift_torch.RAW_POS_LOG
        shift_torch.RAW_POS_LOG = None
        for name, p in self.ref.named_parameters():
            if name.endswith("xpos") and id(p) in log:
                self.raw[name[:-4] + "ypos"] = log[id(p)][1]
        return False


def check_ypos_grad(name, got, want, raw, floor=1e-9, decisive=1e-3, min_sure=0.5):
    """K5 output is +-0.01 by the SIGN of a reduced sum: only comparable where the sum is clearly non-zero
    (|raw| > decisive * max|raw|; against the exact oracle the TF32 noise of the sums needs a wider margin)."""
    got = got.detach().double().cpu()
    want = want.detach().double().cpu()
    scale = raw.abs().max().item() if raw is not None else 0.0
    sure = raw.abs() > max(decisive * scale, floor) if raw is not None else torch.ones_like(want, dtype=torch.bool)
    assert sure.float().mean().item() > min_sure, f"{name}: too few channels with a decisive sign"
    assert torch.equal(got[sure], want[sure].to(got.dtype)) or (got[sure] - want[sure]).abs().max().item() < 1e-9, \
        f"{name}: K5-constrained gradient differs on channels with decisive raw sums"
    assert set(np.round(got.abs().numpy(), 6).tolist()) <= {0.01, 0.0001}, f"{name}: values outside {{0.01, 1e-4}}"
