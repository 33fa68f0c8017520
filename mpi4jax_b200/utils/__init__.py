"""Measurement utilities: device-side timing (max over ranks), clock sampling, L2 flush."""

from .timing import ClockSampler, device_time_ms, flush_l2, max_over_ranks  # noqa: F401
