"""Runoff remapping (<is_remap> T, remap_1D_runoff, process_remap.f90:164-262): oracle restatement vs an independent
numpy evaluation on the CPU; on the GPU the device-side remap (mr_set_remap / k_remap) feeding the routing must equal
routing the oracle-remapped runoff, and the stand-alone host must read a reference-style mapping file."""
import json
import subprocess

import numpy as np
import pytest

from tests.test_host import BACKENDS, _routing_host
from tests.util import case


def make_mapping(net, n_forcing, seed=0):
    """Random ragged mapping: every network HRU but a few overlaps 1-4 forcing polygons; some mapping HRUs are not in
    the network, some polygons have no forcing, some weight rows do not sum to one."""
    rng = np.random.default_rng(seed)
    map_ids = list(net.hruId) + [10**6 + k for k in range(5)]            # 5 mapping HRUs outside the network
    rng.shuffle(map_ids)
    drop = set(rng.choice(net.hruId, size=max(1, net.nHRU // 20), replace=False).tolist())
    map_ids = [i for i in map_ids if i not in drop]                       # a few network HRUs are not in the mapping
    forcing_ids = np.arange(1, n_forcing + 1) * 3
    num, qid, w = [], [], []
    for _ in map_ids:
        k = int(rng.integers(1, 5))
        ids = rng.choice(np.concatenate([forcing_ids, [999999]]), size=k, replace=False)      # 999999: polygon without forcing
        ww = rng.random(k); ww /= ww.sum()
        if rng.random() < 0.3:
            ww *= rng.uniform(0.5, 0.9)                                   # weights that do not sum to one -> renormalised
        num.append(k); qid += ids.tolist(); w += ww.tolist()
    return np.array(map_ids), np.array(num), np.array(qid), np.array(w), forcing_ids


def indices(net, map_ids, qid, forcing_ids):
    pos = {int(h): i for i, h in enumerate(net.hruId)}
    fpos = {int(h): i for i, h in enumerate(forcing_ids)}
    return (np.array([pos.get(int(m), -1) for m in map_ids], dtype=np.int32), np.array([fpos.get(int(q), -1) for q in qid], dtype=np.int32))


def numpy_remap(hru_ix, num, qix, w, sim, n_hru):
    out = np.zeros(n_hru); o = 0
    for i, j in enumerate(hru_ix):
        sl = slice(o, o + num[i]); o += num[i]
        if j < 0:
            continue
        ok = (qix[sl] >= 0)
        ok[ok] &= sim[qix[sl][ok]] > -1e-6
        ws, vs = w[sl][ok], sim[qix[sl][ok]]
        acc, sw = 0.0, 0.0
        for a, b in zip(ws, vs):
            sw = sw + a; acc = acc + a * b
        if sw > 1e-6 and abs(1.0 - sw) > 1e-6:
            acc = acc / sw
        out[j] = acc
    return out


def test_oracle_remap_matches_numpy():
    from oracle import oracle as orc
    net, params, opts, ro = case("random", n=120, seed=4, dt=86400.0, route_opt="1", steps=3)
    map_ids, num, qid, w, fids = make_mapping(net, 90, seed=2)
    hru_ix, qix = indices(net, map_ids, qid, fids)
    rng = np.random.default_rng(5)
    sim = rng.lognormal(np.log(2e-5), 1.0, 90); sim[::11] = -9999.0; sim[5] = -5e-7
    got = orc.remap_1d(hru_ix, num, qix, w, sim, net.nHRU)
    assert np.array_equal(got, numpy_remap(hru_ix, num, qix, w, sim, net.nHRU))
    assert (got[np.isin(net.hruId, map_ids, invert=True)] == 0).all()


@pytest.mark.gpu
def test_device_remap_feeds_routing_like_the_oracle():
    from mizuroute_b200.route import Router
    from oracle import oracle as orc
    from oracle.oracle import Oracle
    net, params, opts, _ = case("conus", n=800, seed=4, dt=3600.0, route_opt="012", steps=1)
    nF, K = 500, 20
    map_ids, num, qid, w, fids = make_mapping(net, nF, seed=7)
    hru_ix, qix = indices(net, map_ids, qid, fids)
    rng = np.random.default_rng(9)
    forcing = rng.lognormal(np.log(2e-5), 1.0, size=(K, nF)); forcing[:, ::13] = -9999.0
    ro_net = np.stack([orc.remap_1d(hru_ix, num, qix, w, forcing[t], net.nHRU) for t in range(K)])
    qo = Oracle(net, params, opts).run(ro_net)
    r = Router(net, params, opts, max_batch=8)
    r.set_remap(nF, hru_ix, num, qix, w)
    qg = np.concatenate([r.route_batch(np.ascontiguousarray(forcing[s:s + 8])) for s in range(0, K, 8)], axis=1)
    assert np.array_equal(qg[0], qo[0]) and np.array_equal(qg[1], qo[1])       # remap + SUM / IRF: bit-identical
    assert np.max(np.abs(qg[2] - qo[2]) / np.maximum(np.abs(qo[2]), 1e-300)) <= 1e-4


@pytest.mark.parametrize("backend", BACKENDS)
def test_host_reads_reference_style_mapping_file(tmp_path, backend):
    from mizuroute_b200 import build as mrbuild, casefiles
    from oracle import oracle as orc
    from oracle.oracle import Oracle
    net, params, opts, _ = case("conus", n=400, seed=5, dt=86400.0, route_opt="12", steps=1)
    nF, K = 300, 12
    map_ids, num, qid, w, fids = make_mapping(net, nF, seed=3)
    hru_ix, qix = indices(net, map_ids, qid, fids)
    forcing = np.random.default_rng(1).lognormal(np.log(2e-5), 1.0, size=(K, nF))
    ctl = casefiles.write_case(str(tmp_path), net, params, opts, forcing, case_name="remap", remap=(map_ids, num, qid, w, fids))
    r = subprocess.run([_routing_host(backend), ctl, "--batch", "5"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = casefiles.read_history(json.loads(r.stdout.strip().splitlines()[-1])["history"])
    ro_net = np.stack([orc.remap_1d(hru_ix, num, qix, w, forcing[t], net.nHRU) for t in range(K)])
    qo = Oracle(net, params, opts).run(ro_net)
    np.testing.assert_allclose(out["IRFroutedRunoff"], qo[0].astype(np.float32), rtol=2e-6, atol=1e-30)
    np.testing.assert_allclose(out["KWTroutedRunoff"], qo[1].astype(np.float32), rtol=1e-4, atol=1e-30)


def test_host_passes_polygon_forcing_through_for_the_device_remap(tmp_path):
    """<is_remap> T: the rows the host hands to the library are the forcing polygons' records untouched (the device
    remaps them), with the fill value turned into realMissing (read_runoff.f90:324) -- checked without a GPU."""
    from mizuroute_b200 import build as mrbuild, casefiles
    net, params, opts, _ = case("random", n=50, seed=5, dt=86400.0, route_opt="1", steps=1)
    nF, K = 40, 5
    map_ids, num, qid, w, fids = make_mapping(net, nF, seed=3)
    forcing = np.random.default_rng(1).lognormal(np.log(2e-5), 1.0, size=(K, nF))
    forcing[2, 7] = -9999.0
    ctl = casefiles.write_case(str(tmp_path), net, params, opts, forcing, case_name="remap", remap=(map_ids, num, qid, w, fids))
    path = str(tmp_path / "rows.f64")
    r = subprocess.run([mrbuild.build_host(), ctl, "--dry-run", "--dump-forcing", path], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert np.array_equal(np.fromfile(path, dtype=np.float64).reshape(K, nF), forcing)


def make_grid_mapping(net, n_lat, n_lon, seed=0):
    """River-network HRUs overlapping 1-4 cells of an [n_lat, n_lon] grid; a few cells lie outside the grid (skipped by
    remap_2D_runoff, process_remap.f90:110-126), weights do not always sum to one."""
    rng = np.random.default_rng(seed)
    num = rng.integers(1, 5, size=net.nHRU).astype(np.int32)
    m = int(num.sum())
    ii = rng.integers(1, n_lon + 1, size=m).astype(np.int32)       # x / lon, 1-based
    jj = rng.integers(1, n_lat + 1, size=m).astype(np.int32)       # y / lat, 1-based
    out = rng.random(m) < 0.03
    ii[out] = n_lon + 1 + rng.integers(0, 3, size=int(out.sum()))
    w = rng.random(m) + 0.05
    ptr = np.concatenate([[0], np.cumsum(num)])
    for h in range(net.nHRU):
        if h % 3:
            w[ptr[h]:ptr[h + 1]] /= w[ptr[h]:ptr[h + 1]].sum()
    order = rng.permutation(net.nHRU)
    return net.hruId[order].astype(np.int32), num, ii, jj, w, order


def numpy_remap_2d(order, num, ii, jj, w, grid, n_hru):
    """Independent restatement of remap_2D_runoff (process_remap.f90:59-162) on grid[lat, lon]."""
    out = np.zeros(n_hru); k = 0
    for h, n in zip(order, num):
        sw = acc = 0.0
        for _ in range(n):
            i, j = ii[k] - 1, jj[k] - 1
            if 0 <= i < grid.shape[1] and 0 <= j < grid.shape[0] and grid[j, i] > -1e-6:
                sw += w[k]; acc += w[k] * grid[j, i]
            k += 1
        if sw > 1e-6 and abs(1.0 - sw) > 1e-6:
            acc /= sw
        out[h] = acc
    return out


def _grid_case(tmp_path, n=60, n_lat=7, n_lon=9, K=5):
    from mizuroute_b200 import casefiles
    net, params, opts, _ = case("random", n=n, seed=6, dt=86400.0, route_opt="12", steps=1)
    map_ids, num, ii, jj, w, order = make_grid_mapping(net, n_lat, n_lon, seed=2)
    grid = np.random.default_rng(4).lognormal(np.log(2e-5), 1.0, size=(K, n_lat, n_lon))
    grid[1, 2, 3] = -9999.0
    ctl = casefiles.write_case(str(tmp_path), net, params, opts, grid, case_name="grid", remap=("grid", map_ids, num, ii, jj, w))
    return net, params, opts, grid, (order, num, ii, jj, w), ctl


def test_host_flattens_gridded_forcing_for_the_device_remap(tmp_path):
    """[time, lat, lon] runoff + i_index/j_index mapping: the host hands the library flattened grid records and flat cell
    indices; the 1-D weighted sum over them (oracle remap_1d) is remap_2D_runoff on the grid."""
    from mizuroute_b200 import build as mrbuild
    from oracle import oracle as orc
    net, params, opts, grid, (order, num, ii, jj, w), ctl = _grid_case(tmp_path)
    rows_path, map_path = str(tmp_path / "rows.f64"), str(tmp_path / "map.bin")
    r = subprocess.run([mrbuild.build_host(), ctl, "--dry-run", "--dump-forcing", rows_path, "--dump-remap", map_path], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert json.loads(r.stdout.strip().splitlines()[0])["nHRU_forcing"] == grid.shape[1] * grid.shape[2]
    rows = np.fromfile(rows_path, dtype=np.float64).reshape(grid.shape[0], -1)
    assert np.array_equal(rows, grid.reshape(grid.shape[0], -1))
    raw = open(map_path, "rb").read()
    n, m = np.frombuffer(raw, dtype=np.int32, count=2)
    hru_ix = np.frombuffer(raw, dtype=np.int32, count=n, offset=8)
    num_q = np.frombuffer(raw, dtype=np.int32, count=n, offset=8 + 4 * n)
    q_ix = np.frombuffer(raw, dtype=np.int32, count=m, offset=8 + 8 * n)
    wgt = np.frombuffer(raw, dtype=np.float64, count=m, offset=8 + 8 * n + 4 * m)
    assert np.array_equal(hru_ix, order) and np.array_equal(num_q, num) and np.array_equal(wgt, w) and (q_ix < 0).sum() == (ii > grid.shape[2]).sum()
    for t in range(grid.shape[0]):
        want = numpy_remap_2d(order, num, ii, jj, w, grid[t], net.nHRU)
        got = orc.remap_1d(hru_ix, num_q, q_ix, wgt, rows[t], net.nHRU)
        assert np.array_equal(got, want)


@pytest.mark.parametrize("backend", BACKENDS)
def test_host_routes_gridded_forcing_like_the_oracle(tmp_path, backend):
    from mizuroute_b200 import build as mrbuild, casefiles
    from oracle.oracle import Oracle
    net, params, opts, grid, (order, num, ii, jj, w), ctl = _grid_case(tmp_path, n=300, n_lat=12, n_lon=15, K=10)
    r = subprocess.run([_routing_host(backend), ctl, "--batch", "4"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = casefiles.read_history(json.loads(r.stdout.strip().splitlines()[-1])["history"])
    ro_net = np.stack([numpy_remap_2d(order, num, ii, jj, w, grid[t], net.nHRU) for t in range(grid.shape[0])])
    qo = Oracle(net, params, opts).run(ro_net)
    np.testing.assert_allclose(out["IRFroutedRunoff"], qo[0].astype(np.float32), rtol=2e-6, atol=1e-30)
    np.testing.assert_allclose(out["KWTroutedRunoff"], qo[1].astype(np.float32), rtol=1e-4, atol=1e-30)


@pytest.mark.gpu
@pytest.mark.parametrize("dt,forcing_dt,records,sim_steps", [(3600.0, 3600.0, 12, 12), (10800.0, 3600.0, 36, 12), (7200.0, 10800.0, 8, 12)])
def test_device_ingest_feeds_routing_like_host_built_rows(dt, forcing_dt, records, sim_steps):
    """mr_set_ingest / mr_ingest_records (k_ingest): raw forcing records in shuffled forcing-HRU order, with fill values, a
    negative value, scale and offset -> resident runoff rows -> routing, against the oracle fed the rows the host build of
    the same source makes (tests/test_ingest_emul.py pins those to the stand-alone host's rows bit for bit)."""
    from mizuroute_b200.route import Router
    from oracle.oracle import Oracle
    from tests.test_ingest_emul import _emul_rows
    from tests.util import case, rel_err
    from tests.util_ingest import time_map
    net, params, opts, ro = case("conus", n=400, seed=4, dt=dt, route_opt="01", steps=records)
    rng = np.random.default_rng(9)
    nIn = net.nHRU + 7
    col = rng.permutation(nIn)[:net.nHRU].astype(np.int32)          # forcing column of every river-network HRU
    col[5] = -1                                                      # an HRU the forcing does not know
    rec = rng.lognormal(-11.0, 1.0, (records, nIn))
    rec[:, col[col >= 0]] = ro[:, col >= 0]
    rec[1, col[3]] = -2.0; rec[2, col[8]] = -9999.0
    ptr, idx, frac = time_map(sim_steps, dt, records, forcing_dt)
    rows = _emul_rows(rec, col, ptr, idx, frac, scale=1.25, offset=1e-10)
    qo = Oracle(net, params, opts).run(rows)
    r = Router(net, params, opts, max_batch=8)
    r.set_ingest(nIn, col, scale=1.25, offset=1e-10)
    parts = []
    for s in range(0, sim_steps, 5):
        e = min(s + 5, sim_steps)
        j0, j1 = ptr[s], ptr[e]
        lo, hi = idx[j0:j1].min(), idx[j0:j1].max() + 1             # only the records this batch needs travel
        K = r.ingest_records(rec[lo:hi], ptr[s:e + 1] - j0, idx[j0:j1] - lo, frac[j0:j1])
        r.route_resident(K)
        parts.append(r.download_q(K))
    qg = np.concatenate(parts, axis=1)
    assert np.array_equal(qg[0], qo[0]) and np.array_equal(qg[1], qo[1])          # SUM / IRF: bit-identical rows -> bit-identical flow
