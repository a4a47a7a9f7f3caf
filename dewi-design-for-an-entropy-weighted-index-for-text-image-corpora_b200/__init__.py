"""dewi_b200 -- B200-native backend for DEWI's retrieval hot path.

Public names mirror the reference package (`dewi`, src/dewi/__init__.py:5-15 plus the modules its
README imports from): `DewiIndex`, `Payload`, `Weights`, `Signals`, `DewiScorer`, `RobustStats`.
Importing the package never touches the GPU; constructing an index or scoring does, and raises
ImportError when libdewi_b200.so or a B200 is missing (no CPU fallback).
"""

__version__ = "0.1.0"

from .types import Payload, Signals, Weights  # noqa: F401
from .backends import BaseIndex, CudaIndex, IndexBackend  # noqa: F401
from .index import DewiIndex  # noqa: F401
from .scorer import DewiScorer, RobustStats  # noqa: F401
from .redundancy import (combine_range_stats, cross_modal_similarity, redundancy_join, self_join_range,  # noqa: F401
                         sharded_self_join)
from .sharded import ShardedDewiIndex, shard_range  # noqa: F401
from .robust import (PayloadRobustStats, cluster_coverage, cluster_pairs, clusters_from_labels, duplicate_rate,  # noqa: F401
                     local_weights_from_surprisal)

__all__ = [
    "__version__", "DewiIndex", "BaseIndex", "CudaIndex", "IndexBackend", "Payload", "Signals", "Weights",
    "DewiScorer", "RobustStats", "cross_modal_similarity", "redundancy_join", "self_join_range", "combine_range_stats", "sharded_self_join", "ShardedDewiIndex", "shard_range", "PayloadRobustStats", "local_weights_from_surprisal",
    "cluster_pairs", "clusters_from_labels", "duplicate_rate", "cluster_coverage",
]
