"""
Host helpers on the sampling path.

`modulo_with_wrapped_range` keeps the name, argument meaning and numerics of
/root/reference/foldingdiff/utils.py:87-121 (shift into [0, span), floor-mod,
shift back) for Python numbers, numpy arrays and torch tensors.  The device
version of the same arithmetic is `fd::wrap_pi` (csrc/common.cuh).
"""
from __future__ import annotations

from typing import Any, Dict

import numpy as np


def modulo_with_wrapped_range(vals, range_min: float = -np.pi, range_max: float = np.pi):
    """
    Wrap `vals` into the half-open range [range_min, range_max).

    >>> modulo_with_wrapped_range(3, -2, 2)
    -1
    """
    assert range_min <= 0.0
    assert range_min < range_max
    span = range_max - range_min
    return ((vals - range_min) % span) + range_min


def update_dict_nonnull(d: Dict[str, Any], vals: Dict[str, Any]) -> Dict[str, Any]:
    """Overlay the non-None entries of `vals` onto `d` (reference utils.py:124-137)."""
    for k, v in vals.items():
        if v is not None:
            d[k] = v
    return d


def seq_to_groups(seq, size: int):
    """Split a sequence into consecutive chunks of at most `size` items."""
    return [seq[i:i + size] for i in range(0, len(seq), size)]
