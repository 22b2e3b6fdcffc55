"""CPU: the library's operator scopes drive the reference's hierarchical Timer protocol (fme/core/benchmark/timer.py:48-51) the way
the conditional SFNO block nests its children (fme/core/models/conditional_sfno/sfnonet.py:388-437, s2convolutions.py:367-433).
``ace_debug_scope`` opens / closes a scope without launching anything, so the hook is exercised without a GPU."""
import pytest

import ace_b200
from ace_b200 import _lib, timing


class RecordingTimer:
    """The Timer protocol with the reference CUDATimer's rules (a child may only be taken from an entered timer, no re-entry)."""

    def __init__(self, log, path=()):
        self.log, self.path, self.children, self.entered, self.count = log, path, {}, False, 0

    def child(self, name):
        if not self.entered:
            raise RuntimeError("child before enter")
        return self.children.setdefault(name, RecordingTimer(self.log, self.path + (name,)))

    def __enter__(self):
        if self.entered:
            raise RuntimeError("already entered")
        self.entered, self.count = True, self.count + 1
        self.log.append(("enter", "/".join(self.path)))
        return self

    def __exit__(self, *exc):
        self.entered = False
        self.log.append(("exit", "/".join(self.path)))
        return False


def _scope(name):
    _lib.check(_lib.load().ace_debug_scope(name.encode()))


BLOCK = ["cln_vector_terms", "cond_layer_norm", "sht.dft_fwd", "sht.legendre_fwd", "dhconv", "sht.legendre_inv", "sht.dft_inv", "inner_skip",
         "cln_vector_terms", "cond_layer_norm", "mlp.fc1", "mlp.fc2"]


def test_block_operators_nest_like_the_reference_block():
    log = []
    root = RecordingTimer(log)
    with root, timing.timer_scopes(root):
        for name in ["norm_split", "encoder.0", "encoder.2"] + BLOCK + BLOCK + ["decoder.0", "decoder.2"]:
            _scope(name)
    assert not any(t.entered for t in root.children.values())
    assert set(root.children) == {"norm_split", "encoder.0", "encoder.2", "norm0", "filter", "inner_skip", "norm1", "mlp", "decoder.0", "decoder.2"}
    assert set(root.children["filter"].children) == {"forward_transform", "dhconv", "inverse_transform"}
    # one entry per block for every child of the block, two blocks
    for name in ("norm0", "filter", "inner_skip", "norm1", "mlp"):
        assert root.children[name].count == 2, name
    assert all(c.count == 2 for c in root.children["filter"].children.values())
    # filter stays open around its three parts; the two GEMMs of a transform share one entry
    i0 = log.index(("enter", "filter"))
    assert log[i0:i0 + 8] == [("enter", "filter"), ("enter", "filter/forward_transform"), ("exit", "filter/forward_transform"),
                              ("enter", "filter/dhconv"), ("exit", "filter/dhconv"), ("enter", "filter/inverse_transform"),
                              ("exit", "filter/inverse_transform"), ("exit", "filter")]
    assert log[0] == ("enter", "") and log[-1] == ("exit", "")


def test_hook_is_removed_and_errors_surface_on_exit():
    log = []
    root = RecordingTimer(log)
    with root, timing.timer_scopes(root):
        _scope("dhconv")
    n = len(log)
    _scope("dhconv")  # no hook any more
    assert len(log) == n
    # a timer that was never entered refuses child(): the error is held back until the context exits (never through the C frames)
    with pytest.raises(RuntimeError, match="child before enter"):
        with timing.timer_scopes(RecordingTimer([])):
            _scope("dhconv")
    _scope("dhconv")


def test_custom_path_function_and_nvtx_option():
    log = []
    root = RecordingTimer(log)
    with root, timing.timer_scopes(root, path_of=lambda name, state: ("ops", name)):
        _scope("a")
        _scope("b")
    assert [e for e in log if e[0] == "enter"] == [("enter", ""), ("enter", "ops"), ("enter", "ops/a"), ("enter", "ops/b")]
    ace_b200.set_option("nvtx", 1)  # header-only NVTX without an attached tool: ranges are no-ops
    try:
        _scope("dhconv")
    finally:
        ace_b200.set_option("nvtx", 0)
