"""TEST INFRASTRUCTURE ONLY - minimal stand-in for the third-party `nncore==0.4.2` package
(reference requirements.txt:7), which is not installable here (no network).

Only the surface the reference's hot path touches is provided, restated
from the published behaviour of nncore: a class registry, `Config.from_file`, `swap_element` and
`ops.temporal_iou`.  Parity at this boundary is UNPINNED by any reference test; the IoU formula is
anchored indirectly on the docstring known answers of the in-tree FlashVTG/span_utils.py:53-59.
"""
import runpy
import types

from . import nn, ops  # noqa: F401


class _AttrDict(dict):
    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError as e:
            raise AttributeError(k) from e
        return _wrap(v)

    def get(self, k, default=None):
        return _wrap(super().get(k, default))


def _wrap(v):
    if isinstance(v, dict) and not isinstance(v, _AttrDict):
        return _AttrDict(v)
    return v


class Config(_AttrDict):
    @classmethod
    def from_file(cls, path):
        ns = runpy.run_path(path)
        return cls({k: v for k, v in ns.items()
                    if not k.startswith("_") and not isinstance(v, types.ModuleType)})


def swap_element(t, i, j):
    """Exchange rows i and j of a 2-D tensor and return it (used by inference.py:42)."""
    i, j = int(i), int(j)
    if i != j:
        tmp = t[i].clone()
        t[i] = t[j]
        t[j] = tmp
    return t
