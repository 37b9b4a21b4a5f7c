"""Process-wide default Engine (one pc_handle per device) for the reference-shaped classes
(LHMM / Clustering / AcousticModel), which have no device argument in the reference."""
from __future__ import annotations

_engines = {}


def get_engine(device=None):
    import torch

    from .engine import Engine

    if device is None:
        device = torch.cuda.current_device() if torch.cuda.is_available() else 0
    if device not in _engines:
        _engines[device] = Engine(device)  # raises without a GPU: there is no CPU path
    return _engines[device]


class NullLog:
    """Duck type of the reference's LogPrint.Log (`note(msg, cls=, show_console=)`, LogPrint.py:64)."""

    def note(self, *a, **k):
        pass

    def close(self):
        pass
