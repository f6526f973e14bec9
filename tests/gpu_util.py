"""Helpers shared by the GPU parity tests."""
import torch


def tf32_round(x):
    """cvt.rna.tf32.f32: round to nearest (ties away), keep 10 mantissa bits."""
    i = x.contiguous().view(torch.int32)
    i = (i + 0x1000) & ~0x1FFF
    return i.view(torch.float32)


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def nchw(x):
    return x.permute(0, 3, 1, 2).contiguous()


def rel_err(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def max_err(a, b):
    return float((a.double() - b.double()).abs().max())


def principal_angles_deg(A, B):
    """Principal angles (degrees) between the row spaces of A and B."""
    qa, _ = torch.linalg.qr(A.double().T.cpu())
    qb, _ = torch.linalg.qr(B.double().T.cpu())
    sv = torch.linalg.svdvals(qa.T @ qb).clamp(max=1.0)
    return torch.rad2deg(torch.acos(sv))
