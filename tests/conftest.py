import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


# GPU tests written after this round's GPU minutes were spent: they have run only through their CPU counterparts (host
# builds of the device code, the oracle-backed stand-in library).  They are collected LAST so that, under `-x`, a surprise
# in one of them cannot hide the tests that have already been green on a B200.  Empty this list once they have run.
NOT_YET_RUN_ON_HARDWARE = (
    "test_host_maps_forcing_records_onto_simulation_steps", "test_daily_history_files_hold_the_single_file_run",
    "test_restart_files_written_during_the_run", "test_host_feeds_lake_evaporation_and_precipitation",
    "test_hype_reservoirs_and_their_calendar_survive_a_restart", "test_exact_restart_of_the_euler_schemes",
    "test_host_run_matches_oracle[345", "test_host_routes_gridded_forcing_like_the_oracle",
    "test_decomposed_euler_schemes_equal_single_domain", "test_host_reads_the_water_management_file",
    "test_host_reads_the_gauge_files_for_direct_insertion", "test_restart_under_data_assimilation_carries_the_discharge_error",
    "test_history_volume_inflow_and_instantaneous_runoff",
    "test_device_ingest_gives_the_same_history_as_host_built_rows", "test_history_at_gauges_only",
)
# ran on a B200 only through scripts/check_unverified_gpu.py (profiles/r1_unverified_gpu_check.jsonl), not under pytest: collected
# after the verified tests and before the ones above
RUN_ON_HARDWARE_OUTSIDE_PYTEST = (
    "test_schemes_gpu.py::test_lake_evaporation", "test_schemes_gpu.py::test_hype", "test_schemes_gpu.py::test_hanasaki",
    "test_schemes_gpu.py::test_water_management", "test_schemes_gpu.py::test_direct_insertion",
    "test_cuda_reproduces_option_golden", "test_device_ingest_feeds_routing_like_host_built_rows", "test_degenerate_networks_all_six_methods",
)


def pytest_collection_modifyitems(config, items):
    mid = [it for it in items if "gpu" in it.keywords and any(k in it.nodeid for k in RUN_ON_HARDWARE_OUTSIDE_PYTEST)]
    late = [it for it in items if "gpu" in it.keywords and any(k in it.nodeid for k in NOT_YET_RUN_ON_HARDWARE)]
    if mid or late:
        ids = {id(it) for it in mid + late}
        items[:] = [it for it in items if id(it) not in ids] + mid + late
    if _has_cuda():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)
