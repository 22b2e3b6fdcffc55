"""Oracle restatement of the inference aggregators' horizontal reductions (torch CPU).  TEST INFRASTRUCTURE.

Follows /root/reference:
  fme/core/metrics.py:35-60    weighted_sum
  fme/core/metrics.py:63-90    weighted_mean
  fme/core/metrics.py:118-143  weighted_std
  fme/core/metrics.py:146-168  weighted_mean_bias
  fme/core/metrics.py:171-197  root_mean_squared_error
  fme/core/metrics.py:388-408  spherical_power_spectrum
  fme/core/distributed/non_distributed.py  zonal_mean = data.mean(dim=-1)
Pinned against the reference's own functions (executed from /root/reference/fme/core/metrics.py by
``oracle/refload.py:load_metrics``) in tests/test_oracle_metrics.py.
"""
import torch

DIMS = (-2, -1)


def weighted_sum(tensor, weights, keepdim=False):
    w = weights.expand(tensor.shape)
    tensor = tensor.where(w != 0.0, 0.0)  # drop "expected NaNs" under zero weight
    return (tensor * w).sum(dim=DIMS, keepdim=keepdim)


def weighted_mean(tensor, weights, keepdim=False):
    w = weights.expand(tensor.shape)
    tensor = tensor.where(w != 0.0, 0.0)
    return (tensor * w).sum(dim=DIMS, keepdim=keepdim) / w.sum(dim=DIMS, keepdim=keepdim)


def weighted_std(tensor, weights):
    mean = weighted_mean(tensor, weights, keepdim=True)
    return torch.sqrt(weighted_mean((tensor - mean) ** 2, weights))


def weighted_mean_bias(truth, predicted, weights):
    return weighted_mean(predicted - truth, weights)


def root_mean_squared_error(truth, predicted, weights):
    return weighted_mean(torch.square(predicted - truth), weights).sqrt()


def zonal_mean(data):
    return data.mean(dim=-1)


def spherical_power_spectrum(field, sht):
    return torch.sum(abs(sht(field)) ** 2, dim=-1)
