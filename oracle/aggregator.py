"""Oracle restatement of the inference aggregators' accumulation (torch CPU).  TEST INFRASTRUCTURE.

Follows /root/reference:
  fme/ace/aggregator/inference/time_mean.py:103-124  TimeMeanAggregator._add_or_initialize_time_mean
  fme/ace/aggregator/inference/time_mean.py:126-160  record_batch (ignore_initial / n_timesteps bookkeeping), get_data
  fme/ace/aggregator/inference/reduced.py:160-212    AreaWeightedReducedMetric.record / get
  fme/ace/aggregator/inference/reduced.py:215-300    MeanAggregator: the metric set
Pinned in tests/test_oracle_aggregator.py against the reference's own ``_add_or_initialize_time_mean`` and
``AreaWeightedReducedMetric`` class bodies, executed from the reference files (``reference_snippets``).
"""
import ast
import os
import textwrap

import torch

from . import metrics as om

REFERENCE_ROOT = "/root/reference"


class TimeMean:
    def __init__(self):
        self._data = None
        self._n_timesteps = 0
        self._n_samples = None

    @staticmethod
    def _add_or_initialize_time_mean(maybe_dict, new_data, ignore_initial=False):
        time_slice = slice(1, None) if ignore_initial else slice(0, None)
        if maybe_dict is None:
            return {name: t[:, time_slice].sum(dim=1).sum(dim=0) for name, t in new_data.items()}
        d = dict(maybe_dict)
        for name, t in new_data.items():
            d[name] += t[:, time_slice].sum(dim=1).sum(dim=0)
        return d

    def record_batch(self, prediction, i_time_start):
        ignore_initial = i_time_start == 0
        self._data = self._add_or_initialize_time_mean(self._data, prediction, ignore_initial)
        first = prediction[list(prediction)[0]]
        if self._n_samples is None:
            self._n_samples = first.size(0)
        if ignore_initial:
            self._n_timesteps = first.size(1) - 1
        else:
            self._n_timesteps += first.size(1)

    def get_data(self):
        return {name: self._data[name] / self._n_timesteps / self._n_samples for name in sorted(self._data)}


class ReducedMetric:
    """AreaWeightedReducedMetric: per-step totals of the batch-mean metric, divided by the batches recorded per step."""

    def __init__(self, compute_metric, n_timesteps):
        self._compute_metric = compute_metric
        self._total = {}
        self._n_batches = torch.zeros(n_timesteps, dtype=torch.int32)
        self._n_timesteps = n_timesteps

    def record(self, target, gen, i_time_start):
        T = next(iter(gen.values())).shape[1]
        sl = slice(i_time_start, i_time_start + T)
        for name, tensor in self._compute_metric(truth=target, predicted=gen).items():
            if name not in self._total:
                self._total[name] = torch.zeros([self._n_timesteps], dtype=tensor.dtype)
            self._total[name][sl] += tensor.mean(dim=0)
        self._n_batches[sl] += 1

    def get(self):
        return {name: t / self._n_batches for name, t in self._total.items()}


def mean_aggregator_metrics(weights):
    """The metric functions MeanAggregator wires up (reduced.py:238-292), on dicts of [B, T, H, W] tensors."""
    def per_var(fn):
        return lambda truth, predicted: {n: fn(truth[n], predicted[n]) for n in predicted}
    return {
        "weighted_rmse": per_var(lambda t, p: om.root_mean_squared_error(t, p, weights)),
        "weighted_bias": per_var(lambda t, p: om.weighted_mean_bias(t, p, weights)),
        "weighted_mean_gen": per_var(lambda t, p: om.weighted_mean(p, weights)),
        "weighted_mean_target": per_var(lambda t, p: om.weighted_mean(t, weights)),
        "weighted_std_gen": per_var(lambda t, p: om.weighted_std(p, weights)),
    }


def reference_snippets():
    """(``_add_or_initialize_time_mean``, ``AreaWeightedReducedMetric``) compiled from the reference's own source text: the two
    class members are cut out of their files by AST position (the files themselves import xarray / wandb and cannot be imported
    in this image) and executed with torch and permissive type names in scope."""
    ns = {"torch": torch, "TensorDict": dict, "TensorMapping": dict, "AreaWeightedFunction": object,
          "get_device": lambda: torch.device("cpu"), "defaultdict": __import__("collections").defaultdict}

    def cut(path, cls_name, member=None):
        with open(os.path.join(REFERENCE_ROOT, path)) as f:
            src = f.read()
        tree = ast.parse(src)
        for node in tree.body:
            if isinstance(node, ast.ClassDef) and node.name == cls_name:
                if member is None:
                    return ast.get_source_segment(src, node)
                for sub in node.body:
                    if isinstance(sub, ast.FunctionDef) and sub.name == member:
                        seg = "\n".join(src.splitlines()[sub.lineno - 1 - len(sub.decorator_list):sub.end_lineno])
                        return textwrap.dedent(seg)
        raise KeyError((path, cls_name, member))

    exec(compile("from __future__ import annotations\n" + cut("fme/ace/aggregator/inference/time_mean.py", "TimeMeanAggregator",
                                                              "_add_or_initialize_time_mean"), "time_mean.py", "exec"), ns)
    exec(compile("from __future__ import annotations\n" + cut("fme/ace/aggregator/inference/reduced.py", "AreaWeightedReducedMetric"),
                 "reduced.py", "exec"), ns)
    add = ns["_add_or_initialize_time_mean"]
    add = add.__func__ if isinstance(add, staticmethod) else add
    return add, ns["AreaWeightedReducedMetric"]
