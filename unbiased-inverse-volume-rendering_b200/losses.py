"""Image losses of the optimisation loop -- host mirror of python/losses.py.

Same names, argument order (`img, ref_img, shape=None`, extra keywords as in the reference) and
the same arithmetic, on torch tensors so that `loss.backward()` reaches `render` / `render_batch`
(the reference: `dr.backward(loss_value)`, optimize.py:349-350).  Every mean divides by
`dr.width(img)` = the number of entries of the image tensor, i.e. `img.numel()`.

Two quirks of the reference are kept on purpose (they decide which gradients the path receives):
  * `huber` compares the SIGNED residual with delta (losses.py:18-21), so every negative residual takes
    the quadratic branch whatever its magnitude;
  * `psnr` is a plain expression of the MSE (losses.py:49-51) without a guard for mse == 0.
"""
from __future__ import annotations

import math

import torch


def _width(img: torch.Tensor) -> int:
    return img.numel()


def average(img, ref_img=None, shape=None):
    """losses.py:4-5."""
    return img.sum() / _width(img)


def l1(img, ref_img, shape=None):
    """losses.py:7-8."""
    return (img - ref_img).abs().sum() / _width(img)


def l2(img, ref_img, shape=None):
    """losses.py:10-11."""
    return ((img - ref_img) ** 2).sum() / _width(img)


def root_mean_squared_error(*args, **kwargs):
    """losses.py:13-14."""
    return torch.sqrt(l2(*args, **kwargs))


def huber(img, ref_img, shape=None, delta=1.0):
    """losses.py:16-22 (signed comparison, see the module docstring)."""
    residual = img - ref_img
    loss = torch.where(residual < delta, 0.5 * residual ** 2, delta * residual.abs() - 0.5 * delta)
    return loss.sum() / _width(img)


def mean_relative_absolute_error(img, ref_img, shape=None, epsilon=1e-2):
    """losses.py:24-26."""
    errors = (img - ref_img).abs() / (ref_img.abs() + epsilon)
    return errors.sum() / _width(img)


def mean_relative_squared_error(img, ref_img, shape=None, epsilon=1e-2):
    """losses.py:28-30."""
    errors = (img - ref_img) ** 2 / (ref_img ** 2 + epsilon)
    return errors.sum() / _width(img)


def root_mean_relative_squared_error(*args, **kwargs):
    """losses.py:32-33."""
    return torch.sqrt(mean_relative_squared_error(*args, **kwargs))


def psnr(img, ref_img, max_value=1.0, shape=None):
    """losses.py:36-52: 20 log10(max_value) - 10 log10(mse)."""
    mse = ((img - ref_img) ** 2).sum() / _width(img)
    return 20.0 * (math.log(max_value) / math.log(10.0)) - (10.0 / math.log(10.0)) * torch.log(mse)


LOSSES = {f.__name__: f for f in (average, l1, l2, root_mean_squared_error, huber, mean_relative_absolute_error,
                                  mean_relative_squared_error, root_mean_relative_squared_error, psnr)}
