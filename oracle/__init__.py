"""CPU oracle for the DEWI retrieval hot path -- TEST INFRASTRUCTURE ONLY.

This package is a numpy restatement of the reference's algorithm for the path
named in BASELINE.json (`ExactIndex.search`, `DewiScorer.fit_stats/score`, the
redundancy similarity).  It is the *checker*: only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu-baseline / `--impl reference`
legs may import it.  Nothing under the product package imports it, and the
product has no CPU fallback.

Parity pin: the restatement is validated against the reference implementation
itself (imported from /root/reference/src in the build container) by
`oracle/make_golden.py`, which also writes the committed fixtures under
`tests/golden/`.  The reference's own tests hold no numeric golden vectors for
this path, so the pin is "outputs of the reference run
here", bit-for-bit for search/scorer.  The thresholded redundancy *join* has no
reference definition beyond `normalize(T) @ normalize(I).T`
(src/dewi/signals/redundancy.py:36-38): its row statistics are "parity
unpinned" and are checked only against this restatement.
"""

from .search import OracleExactIndex, exact_search, exact_search_batch  # noqa: F401
from .scorer import OracleScorer, robust_fit, score_rows  # noqa: F401
from .redundancy import cross_modal_similarity, join_rowstats  # noqa: F401
