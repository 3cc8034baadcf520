"""Stage 3 of the hot path: batched per-event CUSUM+ level segmentation (GPU).

The reference has no implementation (it consumes `level_current_pA`, `level_duration_us`,
`blockages_pA`, `stdev_pA`, `n_levels`: readevents.py:843-846,1297-1306); definition of
record: oracle/events_oracle.py (cusum_event / level_stats)."""
from __future__ import annotations
