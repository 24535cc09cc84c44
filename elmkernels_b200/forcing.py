"""Host side of the per-step input producers: the time bookkeeping that the
reference's data managers do on the host before they launch their per-column functors.

  forcing_time_index / forcing_time_weights   AtmDataManager::forc_t_idx_check_bounds, forcing_time_weights
                                              (reference src/data/atm_data_impl.hh:147-199)
  month_frac / first_month_idx / monthly_data_weights
                                              reference src/data/monthly_data.cc:7-61
The per-column work itself runs on the device (elmk_atm_forcing, elmk_phenology)."""
from __future__ import annotations

import math
from typing import Tuple

NDAYPM = (31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31)   # no leap years, as the reference


def forcing_time_index(days_since_data_start: float, forc_dt_days: float) -> int:
    """Index of the forcing record that starts the interval containing the (step-centred) model time
    (forc_t_idx_check_bounds, atm_data_impl.hh:147-169): truncation, except within 1e-8 days of a record boundary,
    where the reference rounds to the nearest record (forc_t_idx_aligned :122-137)."""
    if abs(math.remainder(days_since_data_start, forc_dt_days)) < 1.0e-8:
        return int(round(days_since_data_start / forc_dt_days))
    return int(days_since_data_start / forc_dt_days)


def forcing_time_weights(days_since_data_start: float, forc_dt_days: float) -> Tuple[int, float, float]:
    """(t_idx, wt1, wt2): linear interpolation weights of records t_idx and t_idx + 1."""
    t_idx = forcing_time_index(days_since_data_start, forc_dt_days)
    elapsed = (days_since_data_start - t_idx * forc_dt_days) / forc_dt_days
    if not (-1.0e-8 <= elapsed <= 1.0):
        raise ValueError("model time outside the forcing interval")
    elapsed = max(elapsed, 0.0)   # a model time a hair before the record it was aligned to sits on that record
    return t_idx, 1.0 - elapsed, elapsed


def month_frac(month: int, day: int, sec: float) -> float:
    """Elapsed fraction of the current month (month 1..12, day 1..31, seconds into the day)."""
    return ((day - 1) + sec / 86400.0) / NDAYPM[month - 1]


def first_month_idx(month: int, day: int, sec: float) -> int:
    t1 = 0 if month_frac(month, day, sec) < 0.5 else 1
    m1 = month + t1 - 2
    return 11 if m1 < 0 else m1


def monthly_data_weights(month: int, day: int, sec: float) -> Tuple[float, float]:
    t = month_frac(month, day, sec)
    t1 = 0 if t < 0.5 else 1
    wt1 = (t1 + 0.5) - t
    return wt1, 1.0 - wt1
