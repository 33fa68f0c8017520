"""Clock sampling of the benchmark harness (mpi4jax_b200/utils/timing.py): works, and never raises, on a
box without a GPU / NVML / nvidia-smi; sample parsing and the summary follow the driver's contract
(`clocks: {sm_mhz, sm_max_mhz, reasons}`)."""

import time

from mpi4jax_b200.utils import ClockSampler


def test_sampler_without_a_gpu_reports_nothing_and_does_not_raise():
    with ClockSampler(0, period_s=0.01, first_delay_s=0.0) as c:
        time.sleep(0.05)
    s = c.summary()
    assert set(s) >= {"sm_mhz", "sm_max_mhz", "reasons", "samples", "source"}
    assert s["reasons"] == [] or isinstance(s["reasons"], list)


def test_summary_of_recorded_samples():
    c = ClockSampler(0)
    c.samples = [["1965", "1965", "700.0", "Not Active", "Not Active", "Not Active", "Not Active"],
                 ["1650", "1965", "990.1", "Not Active", "Not Active", "Not Active", "Active"],
                 ["1700", "1965", "985.0", "Not Active", "Not Active", "Not Active", "Active"]]
    s = c.summary()
    assert s["sm_mhz"] == 1700.0 and s["sm_max_mhz"] == 1965.0
    assert s["reasons"] == ["sw_power_cap"] and s["samples"] == 3
