"""The reference's hierarchical ``Timer`` protocol driven by the library's operator scopes.

Reference: ``fme/core/benchmark/timer.py:48-51`` (``Timer``: ``child(name)`` returns a context manager; ``CUDATimer`` records a CUDA
event pair on the current stream per entry) and its use inside the conditional SFNO block
(``fme/core/models/conditional_sfno/sfnonet.py:388-437``: ``norm0`` / ``filter`` / ``inner_skip`` / ``activation`` / ``norm1`` /
``mlp`` / ``outer_skip``; ``s2convolutions.py:367-433``: ``forward_transform`` / ``dhconv`` / ``inverse_transform`` / ``add_bias`` under
``filter``).  A network of this package runs as ONE library call, so there is no Python frame to wrap each part in; instead the
library reports the begin and end of every operator it enqueues (``ace_set_scope_callback``) and ``timer_scopes`` enters / exits the
matching children -- on the stream the kernels are launched on, which is what ``CUDATimer`` records its events on.

    timer = fme.core.benchmark.timer.CUDATimer()
    with timer, ace_b200.timing.timer_scopes(timer):
        net(x, context)
    timer.result.children["filter"].children["dhconv"].avg_time

Fused parts have no separate time: the GELU of the block rides in the ``inner_skip`` epilogue, the spectral bias and the outer skip in
the inverse transform / ``mlp`` epilogues, so ``activation`` / ``add_bias`` / ``outer_skip`` do not appear.  Operators outside a block
(encoder, decoder, input split, context preparation) appear under their library names.
"""
import ctypes
from typing import Callable, List, Optional

from . import _lib

_CALLBACK = ctypes.CFUNCTYPE(None, ctypes.c_char_p, ctypes.c_int, ctypes.c_void_p)

# library operator name -> path of reference child names
_BLOCK_PATHS = {
    "sht.dft_fwd": ("filter", "forward_transform"),
    "sht.legendre_fwd": ("filter", "forward_transform"),
    "dhconv": ("filter", "dhconv"),
    "diagonal_contract": ("filter", "dhconv"),
    "sht.legendre_inv": ("filter", "inverse_transform"),
    "sht.dft_inv": ("filter", "inverse_transform"),
    "sht.legendre_inv_res": ("filter", "round_trip_residual"),
    "inner_skip": ("inner_skip",),
    "mlp.fc1": ("mlp",),
    "mlp.fc2": ("mlp",),
}
_NORM_OPS = ("cond_layer_norm", "cln_stats", "cln_vector_terms", "prep_norm_conv")


def reference_block_path(name: str, state: dict) -> tuple:
    """Reference child path of the library operator ``name``; ``state`` remembers which of a block's two norms comes next."""
    if name in _NORM_OPS:
        return ("norm1" if state.get("after_filter") else "norm0",)
    path = _BLOCK_PATHS.get(name)
    if path is None:
        return (name,)
    if path == ("inner_skip",):
        state["after_filter"] = True
    elif name == "mlp.fc2":
        state["after_filter"] = False  # the next norm opens the next block
    return path


class timer_scopes:
    """Context manager: while active, every operator the library enqueues is timed under ``timer`` (any object with the reference's
    ``Timer`` protocol, already entered by the caller) at the path ``path_of(name, state)`` gives it.  Consecutive operators that share a
    prefix of their paths keep the shared children open, so ``filter`` is entered once around its three parts like in the reference."""

    def __init__(self, timer, path_of: Callable[[str, dict], tuple] = reference_block_path):
        self.timer, self.path_of = timer, path_of
        self._open: List[tuple] = []  # [(name, entered child)] from the root down
        self._state: dict = {}
        self._pending: Optional[tuple] = None
        self._cb = _CALLBACK(self._on_scope)
        self._error: Optional[BaseException] = None

    def _close_to(self, depth: int):
        while len(self._open) > depth:
            _, child = self._open.pop()
            child.__exit__(None, None, None)

    def _on_scope(self, name, begin, _user):
        if not begin or self._error is not None:
            return
        try:
            path = self.path_of(name.decode(), self._state)
            keep = 0
            while keep < len(self._open) and keep < len(path) and self._open[keep][0] == path[keep]:
                keep += 1
            if keep == len(path) and keep == len(self._open):
                return  # same leaf as the previous operator (e.g. the two GEMMs of a transform): one entry spans both
            self._close_to(keep)
            parent = self._open[-1][1] if self._open else self.timer
            for part in path[keep:]:
                child = parent.child(part)
                child.__enter__()
                self._open.append((part, child))
                parent = child
        except BaseException as e:  # noqa: BLE001 -- never unwind through the C frames; re-raised on exit
            self._error = e

    def __enter__(self):
        _lib.check(_lib.load().ace_set_scope_callback(ctypes.cast(self._cb, ctypes.c_void_p), None))
        return self

    def __exit__(self, exc_type, exc, tb):
        _lib.load().ace_set_scope_callback(None, None)
        self._close_to(0)
        if self._error is not None and exc_type is None:
            raise self._error
        return False
