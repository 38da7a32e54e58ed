"""Torch-free runner for the GPU tests of tests/test_schemes_gpu.py that have not yet run on hardware (lake forcing, HYPE,
Hanasaki, water management, data assimilation): calls the test functions directly, most recent feature first, one JSON line
per test (also into gpurun_out/<tag>_unverified.jsonl), and stops starting new tests after --budget seconds.  Written for a
GPU slot of well under a minute; `pytest -m gpu` remains the real gate."""
import json
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    budget = float(sys.argv[sys.argv.index("--budget") + 1]) if "--budget" in sys.argv else 25.0
    tag = sys.argv[sys.argv.index("--tag") + 1] if "--tag" in sys.argv else "r1s"
    t_start = time.time()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    log = open(os.path.join(ROOT, "gpurun_out", tag + "_unverified.jsonl"), "a")

    def emit(**kw):
        line = json.dumps(kw)
        print(line, flush=True)
        log.write(line + "\n"); log.flush()

    from tests import test_golden as G
    from tests import test_remap as R
    from tests import test_schemes_gpu as T
    plan = [
        # written after the last GPU second of round 1
        ("option_golden[da]", G.test_cuda_reproduces_option_golden, ("conus300_hourly_0145_da",)),
        ("option_golden[wm]", G.test_cuda_reproduces_option_golden, ("conus300_hourly_135_wm",)),
        ("option_golden[345]", G.test_cuda_reproduces_option_golden, ("tree60_hourly_345",)),
        ("device_ingest[1:1]", R.test_device_ingest_feeds_routing_like_host_built_rows, (3600.0, 3600.0, 12, 12)),
        ("device_ingest[3 records per step]", R.test_device_ingest_feeds_routing_like_host_built_rows, (10800.0, 3600.0, 36, 12)),
        ("device_ingest[ragged]", R.test_device_ingest_feeds_routing_like_host_built_rows, (7200.0, 10800.0, 8, 12)),
        ("tiny[one_reach-1]", T.test_degenerate_networks_all_six_methods, ("one_reach", 1)),
        ("tiny[isolated-7]", T.test_degenerate_networks_all_six_methods, ("isolated_reaches", 7)),
        ("tiny[chain-1]", T.test_degenerate_networks_all_six_methods, ("chain_of_two", 1)),
        ("tiny[no_hru-7]", T.test_degenerate_networks_all_six_methods, ("middle_reach_without_hru", 7)),
        ("tiny[star-7]", T.test_degenerate_networks_all_six_methods, ("star_of_five", 7)),
        # passed on a B200 at the end of round 1 (profiles/r1_unverified_gpu_check.jsonl)
        ("direct_insertion[1]", T.test_direct_insertion, (1,)),
        ("water_management[1-0]", T.test_water_management, ("1", 0)),
        ("water_management[134-9]", T.test_water_management, ("134", 9)),
        ("lake_forcing[0-14]", T.test_lake_evaporation_and_precipitation_forcing, (0, "14")),
        ("hype[standard-13]", T.test_hype_reservoirs, ("standard", (2000, 2, 25, 0.0), "13")),
        ("hanasaki[0]", T.test_hanasaki_reservoirs, None),
        ("direct_insertion[3]", T.test_direct_insertion, (3,)),
        ("water_management[5-0]", T.test_water_management, ("5", 0)),
        ("lake_forcing[2-14]", T.test_lake_evaporation_and_precipitation_forcing, (2, "14")),
        ("hype[noleap-1]", T.test_hype_reservoirs, ("noleap", (2001, 12, 28, 43200.0), "1")),
        ("hype_errors", T.test_hype_without_calendar_or_parameters_is_an_error, ()),
        ("water_management_in_kwt", T.test_water_management_in_kwt, ()),
    ]
    h06 = [m.args[1] for m in T.test_hanasaki_reservoirs.pytestmark if m.name == "parametrize"][0]
    emit(item="start", import_s=round(time.time() - t_start, 2))
    for name, fn, args in plan:
        if time.time() - t_start > budget:
            emit(item=name, status="not started (budget)")
            continue
        if args is None:
            args = tuple(h06[0])
        t0 = time.time()
        try:
            fn(*args)
            emit(item=name, status="passed", seconds=round(time.time() - t0, 2))
        except Exception as e:                                    # noqa: BLE001 -- report and go on
            emit(item=name, status="FAILED", seconds=round(time.time() - t0, 2), error=repr(e)[:300], where=traceback.format_exc().splitlines()[-3:])
    emit(item="done", seconds=round(time.time() - t_start, 2))


if __name__ == "__main__":
    main()
