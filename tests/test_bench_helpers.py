"""CPU checks of bench.py's reporting helpers (no GPU, no timing)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_build_phase_roofline_uses_survey_bytes():
    import bench

    n, stored, peak = 500_000_000, 499_000_000, 6561.0
    prof = {"radix_pass_kernel<W,false,ByteDigit<W>>": {"n": 3, "ms": 9.0}, "seg_sort_kernel<W,true>": {"n": 1, "ms": 6.0},
            "merge_apply_kernel<W,Suf,OP>": {"n": 1, "ms": 3.0}, "dir_fill_kernel": {"n": 1, "ms": 0.05}}
    r = bench.build_phase_roofline(prof, n, stored, peak)
    assert set(r) == {"radix_pass_kernel", "seg_sort_kernel", "merge_apply_kernel", "_aggregate"}
    agg = r["_aggregate"]
    assert agg["algorithmic_bytes"] == 3 * n * 16 + n * 16 + (n * 8 + stored * 4 + 3 * (1 << 24) * 4)
    assert abs(agg["kernel_ms"] - 18.05) < 1e-9
    assert r["radix_pass_kernel"]["algorithmic_bytes_per_launch"] == n * 16          # read + scatter of 8-byte words
    assert abs(r["radix_pass_kernel"]["achieved_GBps"] - 3 * n * 16 / 9.0e-3 / 1e9) < 1e-6
    assert r["merge_apply_kernel"]["algorithmic_bytes_per_launch"] == n * 8 + stored * 4 + 3 * (1 << 24) * 4
    assert 0 < r["seg_sort_kernel"]["frac_of_hbm_peak"] < 1
    assert bench.build_phase_roofline(None, n, stored, peak) is None


def test_external_clock_sampler_without_a_gpu():
    """The child-process sampler of N > 1 runs (bench.ExternalClockSampler) must not take the bench down when NVML is
    unusable (no driver here): it reports no samples instead."""
    import time

    import bench

    smp = bench.sampler_at(0, 1, None, 2, 0, "")
    assert isinstance(smp, bench.ExternalClockSampler)
    smp = bench.sampler_at(1, 1, smp, 2, 0, "")
    time.sleep(0.2)
    out = smp.stop()
    assert set(out) >= {"sm_mhz", "sm_max_mhz", "reasons"} and not out.get("samples")
