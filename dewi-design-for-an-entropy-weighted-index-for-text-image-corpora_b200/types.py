"""Payload / Weights / Signals records of the DEWI API.

`Payload` and `Weights` mirror the reference's dataclasses (src/dewi/types.py:8-51): same field
names, order, defaults and JSON codecs, so objects and files interchange with the reference.
`Signals` is named by the reference's README (README.md:67,83-92,109) but missing from its code
(SURVEY.md section 0.4); it is defined here so the quick-start runs: a dataclass of the seven signal
fields that also speaks the mapping protocol `DewiScorer.fit_stats/score` index rows with
(src/dewi/scorer.py:20-21,53-57).
"""

from __future__ import annotations

import json
from dataclasses import asdict, dataclass, fields
from typing import Dict, Iterator

PAYLOAD_FIELDS = ("dewi", "ht_mean", "ht_q90", "hi_mean", "hi_q90", "I_hat", "redundancy", "noise")
SIGNAL_FIELDS = PAYLOAD_FIELDS[1:]


@dataclass
class Payload:
    """Per-document scores carried by the index (types.py:8-39)."""

    dewi: float = 0.0
    ht_mean: float = 0.0
    ht_q90: float = 0.0
    hi_mean: float = 0.0
    hi_q90: float = 0.0
    I_hat: float = 0.0
    redundancy: float = 0.0
    noise: float = 0.0

    def to_dict(self) -> Dict[str, float]:
        return asdict(self)

    @classmethod
    def from_dict(cls, data: Dict[str, float]) -> "Payload":
        names = {f.name for f in fields(cls)}
        return cls(**{k: float(v) for k, v in data.items() if k in names})  # extras ignored (types.py:28-30)

    def to_bytes(self) -> bytes:
        return json.dumps(self.to_dict()).encode("utf-8")

    @classmethod
    def from_bytes(cls, data: bytes) -> "Payload":
        return cls.from_dict(json.loads(data.decode("utf-8")))


@dataclass
class Weights:
    """Mixing weights of the DEWI score (types.py:42-51)."""

    alpha_t: float = 1.0
    alpha_i: float = 1.0
    alpha_m: float = 1.0
    alpha_r: float = 1.0
    alpha_n: float = 1.0
    delta: float = 3.0


@dataclass
class Signals:
    """The seven per-document signals, usable as a dataclass (`.__dict__`, `__annotations__`) and as
    the mapping `DewiScorer` expects (`keys()`, `sig["ht_mean"]`)."""

    ht_mean: float = 0.0
    ht_q90: float = 0.0
    hi_mean: float = 0.0
    hi_q90: float = 0.0
    I_hat: float = 0.0
    redundancy: float = 0.0
    noise: float = 0.0

    def keys(self):
        return self.__dict__.keys()

    def values(self):
        return self.__dict__.values()

    def items(self):
        return self.__dict__.items()

    def __getitem__(self, key: str) -> float:
        try:
            return self.__dict__[key]
        except KeyError:
            raise KeyError(key) from None

    def __iter__(self) -> Iterator[str]:
        return iter(self.__dict__)

    def __len__(self) -> int:
        return len(self.__dict__)

    def __contains__(self, key) -> bool:
        return key in self.__dict__
