"""CPU oracle for the Lite-MKD episodic matching + D2M distillation hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package (`lite-mkd_b200/`) may import this
module.  The only legitimate importers are `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` (as the checker or the timed CPU arm,
never as the thing shipped).

What it is: a torch-CPU restatement (fp32 or fp64, autograd for the backward) of the
reference's algorithm for the path named in BASELINE.json `north_star`.  Each function cites
the reference file:line it follows (paths relative to the reference repo root).

Parity pin: the reference ships no tests or golden vectors for this path, so
the oracle is pinned against outputs of the reference code itself, imported and run in the
authoring container by `tests/golden/make_golden.py`; the resulting fixtures are committed
under `tests/golden/*.npz` and `tests/test_oracle_golden.py` checks the oracle against them.
"""
from .matching import (  # noqa: F401
    cos_sim, otam_cum_dist, otam_cum_dist_stable, otam_logits, otam_pair_dists,
    positional_encoding_table, frame_tuples, trx_logits, trx_branch_logits,
    trx_class_prototypes, trx_sup_outputs, support_dk, e_dist_logits, strm_distance_logits,
)
from .heads import frame_pool, feature_heads  # noqa: F401
from .loader import (  # noqa: F401
    write_feature, scan_teacher_tree, load_teacher_feature, episode_teacher_features,
)
from .losses import (  # noqa: F401
    kd_loss, inter_class_relation, cross_entropy, mse, Recipes, aggregate_accuracy,
)
