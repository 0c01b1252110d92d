"""Shared helpers for the parity tests."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def golden(name):
    return np.load(os.path.join(GOLDEN, name + '.npz'))


def rel_l2(a, b):
    a = torch.as_tensor(a).double().cpu(); b = torch.as_tensor(b).double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def rel_max(a, b):
    """max|a-b| / max|b| -- relative metric with a floor, since single entries cross zero."""
    a = torch.as_tensor(a).double().cpu(); b = torch.as_tensor(b).double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def assert_close(a, b, tol, what=''):
    r2, rm = rel_l2(a, b), rel_max(a, b)
    assert r2 <= tol and rm <= tol, f'{what}: rel_l2={r2:.3e} rel_max={rm:.3e} > tol {tol:.1e}'
