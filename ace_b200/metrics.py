"""Device reductions of the inference aggregators, executed by libace_b200.

Mirrors, for lat-lon grids on a CUDA device:

* ``/root/reference/fme/core/gridded_ops.py:284-360`` ``LatLonOperations`` -- ``area_weighted_sum`` / ``_mean`` /
  ``_mean_bias`` / ``_rmse`` / ``_std`` over the last two (horizontal) dims and ``zonal_mean`` (mean over longitude,
  ``fme/core/distributed/non_distributed.py``); same argument names, ``keepdim`` semantics and validation
  (area weights must be longitudinally uniform, ``:305-311``);
* ``/root/reference/fme/core/metrics.py:35-197`` ``weighted_sum / weighted_mean / weighted_std / weighted_mean_bias /
  root_mean_squared_error`` -- the arithmetic, incl. "points with zero weight contribute nothing even if NaN";
* ``/root/reference/fme/core/metrics.py:388-408`` ``spherical_power_spectrum(field, sht)``.

All statistics of one call come from ONE pass over the data (``ace_weighted_moments``: fp64 accumulation of
sum w x, sum w x^2, sum w (x-t), sum w (x-t)^2, sum w), where the reference launches 4-8 elementwise / reduction kernels
per metric.  No CPU path: CPU tensors raise ``AceError``.
"""
import ctypes

import torch

from . import _lib


def _check(x, what):
    if not x.is_cuda:
        raise _lib.AceError(f"{what}: input must be a CUDA tensor (ace_b200 has no CPU path)")


def _moments(x, t, w):
    """x, t [..., H, W] fp32 CUDA, w [H, W] -> float64 [..., 5]."""
    lead = x.shape[:-2]
    hw = x.shape[-2] * x.shape[-1]
    nf = int(torch.tensor(lead).prod().item()) if len(lead) else 1
    out = torch.empty(*lead, 5, dtype=torch.float64, device=x.device)
    if nf == 0:
        return out
    x = x.float().contiguous()
    if t is not None:
        t = t.float().contiguous()
    with torch.cuda.device(x.device):
        for f0 in range(0, nf, 65535):  # grid.y limit
            n = min(65535, nf - f0)
            _lib.check(_lib.load().ace_weighted_moments(
                ctypes.c_void_p(x.data_ptr() + f0 * hw * 4), ctypes.c_void_p(t.data_ptr() + f0 * hw * 4) if t is not None else None,
                ctypes.c_void_p(w.data_ptr()), n, hw, ctypes.c_void_p(out.data_ptr() + f0 * 5 * 8), _lib.current_stream_ptr()))
    return out


class LatLonOperations:
    HORIZONTAL_DIMS = (-2, -1)

    def __init__(self, area_weights: torch.Tensor, grid: str = "legendre-gauss"):
        if area_weights.dim() != 2:
            raise ValueError(f"area_weights must be [n_lat, n_lon], got {tuple(area_weights.shape)}")
        if not torch.allclose(area_weights, area_weights[..., :1]):
            raise ValueError("Area weights must be longitudinally uniform, as assumed for zonal mean.")
        self._cpu_area_global = area_weights.to("cpu", copy=True)
        self._grid = grid
        self._device_area = {}

    def _weights(self, data):
        _check(data, "LatLonOperations")
        if tuple(data.shape[-2:]) != tuple(self._cpu_area_global.shape):
            raise ValueError(f"data horizontal shape {tuple(data.shape[-2:])} != area weights {tuple(self._cpu_area_global.shape)}")
        w = self._device_area.get(data.device)
        if w is None:
            w = self._cpu_area_global.to(data.device, dtype=torch.float32).contiguous()
            self._device_area[data.device] = w
        return w

    @staticmethod
    def _shape(v, data, keepdim):
        v = v.to(torch.float32 if data.dtype != torch.float64 else torch.float64)
        return v[..., None, None] if keepdim else v

    def area_weighted_sum(self, data, keepdim: bool = False, name=None):
        m = _moments(data, None, self._weights(data))
        return self._shape(m[..., 0], data, keepdim)

    def area_weighted_mean(self, data, keepdim: bool = False, name=None):
        m = _moments(data, None, self._weights(data))
        return self._shape(m[..., 0] / m[..., 4], data, keepdim)

    def area_weighted_mean_bias(self, truth, predicted, name=None):
        assert truth.shape == predicted.shape, "Truth and predicted should have the same shape."
        m = _moments(predicted, truth, self._weights(predicted))
        return self._shape(m[..., 2] / m[..., 4], predicted, False)

    def area_weighted_rmse(self, truth, predicted, name=None):
        assert truth.shape == predicted.shape, "Truth and predicted should have the same shape."
        m = _moments(predicted, truth, self._weights(predicted))
        return self._shape((m[..., 3] / m[..., 4]).sqrt(), predicted, False)

    def area_weighted_std(self, data, keepdim: bool = False, name=None):
        m = _moments(data, None, self._weights(data))
        mean = m[..., 0] / m[..., 4]
        var = (m[..., 1] / m[..., 4] - mean * mean).clamp_min(0.0)
        return self._shape(var.sqrt(), data, keepdim)

    def area_weighted_statistics(self, predicted, truth=None):
        """Everything the one pass yields: dict of mean, std (and bias, rmse when ``truth`` is given)."""
        m = _moments(predicted, truth, self._weights(predicted))
        mean = m[..., 0] / m[..., 4]
        out = {"mean": mean.float(), "std": (m[..., 1] / m[..., 4] - mean * mean).clamp_min(0.0).sqrt().float()}
        if truth is not None:
            out["bias"] = (m[..., 2] / m[..., 4]).float()
            out["rmse"] = (m[..., 3] / m[..., 4]).sqrt().float()
        return out

    def zonal_mean(self, data):
        _check(data, "zonal_mean")
        x = data.float().contiguous()
        lead, (h, w) = x.shape[:-2], x.shape[-2:]
        out = torch.empty(*lead, h, dtype=torch.float32, device=x.device)
        nf = out.numel() // h if h else 0
        if nf:
            with torch.cuda.device(x.device):
                _lib.check(_lib.load().ace_zonal_mean(ctypes.c_void_p(x.data_ptr()), nf, h, w, ctypes.c_void_p(out.data_ptr()),
                                                      _lib.current_stream_ptr()))
        return out

    def get_real_sht(self):
        from .sht import RealSHT

        return RealSHT(*self._cpu_area_global.shape, grid=self._grid)

    def get_real_isht(self):
        from .sht import InverseRealSHT

        return InverseRealSHT(*self._cpu_area_global.shape, grid=self._grid)

    def get_initialization_kwargs(self):
        return {"area_weights": self._cpu_area_global}


def spherical_power_spectrum(field: torch.Tensor, sht) -> torch.Tensor:
    """sum over m of |sht(field)|^2 per total wavenumber l (fme/core/metrics.py:388-408)."""
    _check(field, "spherical_power_spectrum")
    c = sht(field)  # complex64 [..., L, M]
    c = c.contiguous()
    lead, (L, M) = c.shape[:-2], c.shape[-2:]
    out = torch.empty(*lead, L, dtype=torch.float32, device=c.device)
    nf = out.numel() // L if L else 0
    if nf:
        with torch.cuda.device(c.device):
            _lib.check(_lib.load().ace_power_spectrum(ctypes.c_void_p(c.data_ptr()), nf, L, M, ctypes.c_void_p(out.data_ptr()),
                                                      _lib.current_stream_ptr()))
    return out


# ---------------------------------------------------------------------------------------------- window aggregators (row f3)
def _reduce_mean(t: torch.Tensor) -> torch.Tensor:
    """``Distributed.reduce_mean`` (fme/core/distributed/torch_distributed.py:130-136): mean over the data-parallel ranks."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        t = t.clone()
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        t /= dist.get_world_size()
    return t


class TimeMeanAggregator:
    """Time-mean maps of the generated fields -- the device work of ``TimeMeanAggregator.record_batch`` / ``get_data``
    (fme/ace/aggregator/inference/time_mean.py:103-160) on PACKED windows.

    ``record_batch(prediction [B, T, F, H, W], i_time_start)`` adds the window's sum over (sample, time) to a running fp32 map
    ``[F, H, W]`` in ONE kernel (``ace_time_sum``; the reference launches two reductions + one add per variable); the first time
    of the first window (the initial condition) is ignored like ``ignore_initial`` (:130-135).  ``get_data()`` divides by the
    number of recorded time steps and samples and averages over the data-parallel ranks (:151-160)."""

    def __init__(self, names):
        self.names = list(names)
        self._sum = None
        self._n_timesteps = 0
        self._n_samples = None

    @torch.no_grad()
    def record_batch(self, prediction: torch.Tensor, i_time_start: int = 0):
        _check(prediction, "TimeMeanAggregator")
        if prediction.dim() != 5 or prediction.shape[2] != len(self.names):
            raise ValueError(f"expected [sample, time, {len(self.names)}, H, W], got {tuple(prediction.shape)}")
        x = prediction.float().contiguous()
        B, T, F, H, W = x.shape
        ignore_initial = i_time_start == 0
        t0 = 1 if ignore_initial else 0
        if self._sum is None:
            self._sum = torch.zeros(F, H, W, dtype=torch.float32, device=x.device)
        if T - t0 > 0:
            n = F * H * W
            with torch.cuda.device(x.device):
                _lib.check(_lib.load().ace_time_sum(ctypes.c_void_p(x.data_ptr() + t0 * n * 4), B, T * n, T - t0, n, n,
                                                    ctypes.c_void_p(self._sum.data_ptr()), _lib.current_stream_ptr()))
        if self._n_samples is None:
            self._n_samples = B
        self._n_timesteps = (T - 1) if ignore_initial else self._n_timesteps + T

    def get_data(self):
        if self._n_timesteps == 0 or self._sum is None:
            raise ValueError("No data recorded.")
        gen = _reduce_mean(self._sum / self._n_timesteps / self._n_samples)
        order = sorted(range(len(self.names)), key=lambda i: self.names[i])  # sorted for rank-consistent order (:156)
        return {self.names[i]: gen[i] for i in order}


class MeanAggregator:
    """Per-forecast-step area-weighted series -- the device work of ``MeanAggregator`` / ``AreaWeightedReducedMetric``
    (fme/ace/aggregator/inference/reduced.py:160-300): ``weighted_rmse``, ``weighted_bias``, ``weighted_mean_gen``,
    ``weighted_mean_target``, ``weighted_std_gen`` of every variable at every forecast step, each the mean over the batch,
    accumulated over recorded batches and divided by the number of batches per step in ``get()`` (:200-212).

    All five statistics of a window come from ONE pass over the generated data (and one over the target for its mean):
    ``ace_weighted_moments`` on ``[B * T * F]`` fields (the reference runs five dict-of-tensor metric functions, each several
    elementwise + reduction kernels per variable).  ``weighted_grad_mag_percent_diff`` is not built."""

    METRICS = ("weighted_rmse", "weighted_bias", "weighted_mean_gen", "weighted_mean_target", "weighted_std_gen")

    def __init__(self, gridded_operations: LatLonOperations, names, n_timesteps: int):
        self._ops = gridded_operations
        self.names = list(names)
        self._n_timesteps = int(n_timesteps)
        self._total = None       # [metric, F, n_timesteps] float32
        self._n_batches = None   # [n_timesteps] int32

    @torch.no_grad()
    def record_batch(self, target: torch.Tensor, gen: torch.Tensor, i_time_start: int = 0):
        """target, gen: [B, T, F, H, W] (denormalised or normalised, as the reference's ``target=`` option decides)."""
        if target.shape != gen.shape:
            raise RuntimeError(f"Tensors in target and gen must have the same shape, but got {tuple(target.shape)} and {tuple(gen.shape)}")
        _check(gen, "MeanAggregator")
        B, T, F = gen.shape[:3]
        if F != len(self.names):
            raise ValueError(f"expected {len(self.names)} fields, got {F}")
        w = self._ops._weights(gen)
        mg = _moments(gen, target, w)    # [B, T, F, 5]: sum w g, sum w g^2, sum w (g - t), sum w (g - t)^2, sum w
        mt = _moments(target, None, w)
        sw = mg[..., 4]
        mean_g = mg[..., 0] / sw
        vals = torch.stack([
            (mg[..., 3] / sw).sqrt(),                                              # weighted_rmse (metrics.py:171-197)
            mg[..., 2] / sw,                                                       # weighted_bias
            mean_g,                                                                # weighted_mean_gen
            mt[..., 0] / mt[..., 4],                                               # weighted_mean_target
            (mg[..., 1] / sw - mean_g * mean_g).clamp_min(0.0).sqrt(),             # weighted_std_gen
        ]).float()                                                                 # [5, B, T, F]
        new = vals.mean(dim=1).permute(0, 2, 1)                                    # batch mean (:199) -> [5, F, T]
        if self._total is None:
            self._total = torch.zeros(len(self.METRICS), F, self._n_timesteps, dtype=torch.float32, device=gen.device)
            self._n_batches = torch.zeros(self._n_timesteps, dtype=torch.int32, device=gen.device)
        sl = slice(i_time_start, i_time_start + T)
        self._total[:, :, sl] += new
        self._n_batches[sl] += 1

    def get(self):
        """{metric: {name: series [n_timesteps]}}: totals / batches per step, averaged over the data-parallel ranks
        (reduced.py:36-55 ``get_series_data`` -> ``dist.reduce_mean``)."""
        if self._total is None:
            raise ValueError("No batches have been recorded.")
        series = _reduce_mean(self._total / self._n_batches)
        return {m: {n: series[i, j] for j, n in sorted(enumerate(self.names), key=lambda kv: kv[1])} for i, m in enumerate(self.METRICS)}
