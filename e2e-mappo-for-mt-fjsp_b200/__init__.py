"""B200-native batched MT-FJSP disjunctive-graph environment (drop-in for the reference's
trainer/parallel_env.py hot path).  See DESIGN.md."""
from . import instances  # noqa: F401
