"""Typed trajectory result, interface-compatible with the reference's
``TrajectoryResult`` (reference: src/scpn_quantum_control/phase/results.py:21-68): frozen,
read-only float64 arrays, Mapping over {"times", "R"}, read-only metadata."""

from __future__ import annotations

from collections.abc import Iterator, Mapping
from dataclasses import dataclass, field
from types import MappingProxyType
from typing import Any

import numpy as np


def _checked_1d(name: str, values: Any) -> np.ndarray:
    arr = np.asarray(values, dtype=np.float64)
    if arr.ndim != 1:
        raise ValueError(f"{name} must be one-dimensional, got shape {arr.shape}")
    return arr


@dataclass(frozen=True)
class TrajectoryResult(Mapping):
    """Immutable Kuramoto trajectory with legacy mapping compatibility."""

    times: np.ndarray
    R: np.ndarray
    metadata: Mapping[str, Any] = field(default_factory=dict)

    def __post_init__(self) -> None:
        t = _checked_1d("times", self.times)
        r = _checked_1d("R", self.R)
        if t.shape != r.shape:
            raise ValueError(f"times and R must have identical shape, got {t.shape} and {r.shape}")
        for name, arr in (("times", t), ("R", r)):
            if not np.all(np.isfinite(arr)):
                raise ValueError(f"{name} must contain only finite values")
        frozen = []
        for arr in (t, r):
            c = np.array(arr, dtype=np.float64, copy=True)
            c.setflags(write=False)
            frozen.append(c)
        object.__setattr__(self, "times", frozen[0])
        object.__setattr__(self, "R", frozen[1])
        object.__setattr__(self, "metadata", MappingProxyType(dict(self.metadata)))

    def __getitem__(self, key: str) -> np.ndarray:
        if key == "times":
            return self.times
        if key == "R":
            return self.R
        raise KeyError(key)

    def __iter__(self) -> Iterator[str]:
        return iter(("times", "R"))

    def __len__(self) -> int:
        return 2

    def to_dict(self) -> dict[str, np.ndarray]:
        """Return the legacy dictionary representation."""
        return {"times": self.times, "R": self.R}
