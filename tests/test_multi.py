"""Multi-domain routing: decomposition invariants (CPU), the hand-off plan over torch.distributed
with world_size 2 on gloo (CPU), and -- on the GPU -- a decomposed run that must equal the single-domain run."""
import os

import numpy as np
import pytest

from mizuroute_b200 import partition, synth
from mizuroute_b200.synth import _down_index
from tests.util import case


@pytest.mark.parametrize("nparts", [2, 3, 8])
def test_decomposition_invariants(nparts):
    net = synth.conus_like(20000, seed=3)
    dec = partition.decompose(net, nparts)
    down = _down_index(net)
    size = partition.upstream_size(net)
    # brute-force check of upstream_size on a few reaches
    ups = [[] for _ in range(net.nRch)]
    for i, dn in enumerate(down):
        if dn >= 0:
            ups[dn].append(i)
    for r in np.random.default_rng(0).integers(0, net.nRch, 30):
        stack, cnt = [int(r)], 0
        while stack:
            x = stack.pop(); cnt += 1; stack.extend(ups[x])
        assert cnt == size[r]
    allr = np.concatenate(dec.trib + [dec.mainstem])
    assert np.array_equal(np.sort(allr), np.arange(net.nRch))                 # every reach in exactly one domain
    assert np.all(size[dec.mainstem] > net.nRch // nparts) and np.all(size[np.concatenate(dec.trib)] <= net.nRch // nparts)
    is_main = np.zeros(net.nRch, bool); is_main[dec.mainstem] = True
    assert np.all(is_main[down[dec.mainstem][down[dec.mainstem] >= 0]])       # mainstem is closed downstream
    owner = np.full(net.nRch, -1); [owner.__setitem__(t, k) for k, t in enumerate(dec.trib)]
    t_all = np.concatenate(dec.trib)
    inner = t_all[(down[t_all] >= 0) & ~is_main[np.where(down[t_all] >= 0, down[t_all], 0)]]
    assert np.array_equal(owner[inner], owner[down[inner]])                    # a tributary reach and its downstream share a rank
    assert np.all(is_main[down[dec.outlets]]) and np.array_equal(owner[dec.outlets], dec.outlet_owner)
    assert np.all(np.diff(dec.outlet_owner) >= 0)
    for k in range(nparts):
        lo, hi = dec.slot_range(k)
        assert np.array_equal(dec.outlets[lo:hi], dec.outlets_of(k))
    if dec.mainstem.size:
        ms = partition.mainstem_network(net, dec)
        assert ms.nRch == dec.mainstem.size + dec.outlets.size and ms.meta["ghost_mask"].sum() == dec.outlets.size
        assert not np.isin(ms.hruSegId, net.segId[dec.outlets]).any()


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from mizuroute_b200.multi import exchange_rows
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # world 2: one binary tree (the root is the mainstem, its two subtrees the tributary domains);
    # world 3: a forest whose largest basin exceeds nRch/3
    net = synth.binary_tree(2047, seed=7) if world == 2 else synth.conus_like(6000, seed=7)
    dec = partition.decompose(net, world, mainstem_cost=1.0)
    mine = dec.outlets_of(rank)
    L = 5
    local = torch.tensor(mine[:, None] * 1000.0 + np.arange(L)[None, :], dtype=torch.float64).reshape(len(mine), L)
    gathered = torch.full((dec.outlets.size, L), -1.0, dtype=torch.float64)
    exchange_rows(local, gathered, dec, rank, world)
    if rank == 0:
        want = dec.outlets[:, None] * 1000.0 + np.arange(L)[None, :]
        q.put((dec.outlets.size, bool(np.array_equal(gathered.numpy(), want))))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_hand_off_plan_gloo(world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500) + world
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    n_out, ok = q.get(timeout=120)
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert n_out > 0 and ok


@pytest.mark.gpu
@pytest.mark.parametrize("kind,n,nparts,dt,route", [("random", 300, 3, 3600.0, "012"), ("conus", 6000, 4, 86400.0, "12"),
                                                    ("conus", 6000, 8, 3600.0, "2")])
def test_decomposed_equals_single_domain(kind, n, nparts, dt, route):
    from mizuroute_b200.multi import route_decomposed_local
    from mizuroute_b200.route import Router
    net, params, opts, ro = case(kind, n=n, seed=7, dt=dt, route_opt=route, steps=20)
    single = Router(net, params, opts, max_batch=20).route_batch(ro)
    q, dec = route_decomposed_local(net, params, opts, ro, nparts, batch=7)
    assert dec.mainstem.size > 0 and dec.outlets.size > 0
    assert not np.isnan(q).any()
    assert np.array_equal(q, single), "decomposed routing must reproduce the single-domain run bit for bit"


@pytest.mark.gpu
def test_decomposed_euler_schemes_equal_single_domain():
    """The tributary -> mainstem hand-off carries REACH_Q of every active method, so the Euler schemes (route_opt 3/4/5)
    decompose like SUM / IRF: bit-identical to the single-domain run."""
    from mizuroute_b200.multi import route_decomposed_local
    from mizuroute_b200.route import Router
    net, params, opts, ro = case("conus", n=3000, seed=7, dt=3600.0, route_opt="345", steps=20)
    single = Router(net, params, opts, max_batch=20).route_batch(ro)
    q, dec = route_decomposed_local(net, params, opts, ro, 4, batch=7)
    assert dec.mainstem.size > 0 and dec.outlets.size > 0
    assert np.array_equal(q, single)


def test_subnetworks_carry_the_named_lake_parameters():
    from mizuroute_b200 import synth
    from mizuroute_b200.partition import decompose, mainstem_network, subnetwork
    net, params, opts, ro = case("conus", n=600, seed=4, dt=86400.0, route_opt="1", steps=1, lakes=8)
    synth.make_hype_lakes(net, np.random.default_rng(5), frac=0.7)
    dec = decompose(net, 3)
    subs = [subnetwork(net, r) for r in dec.trib if len(r)] + [mainstem_network(net, dec)]
    for sub in subs:
        pos = {int(s): i for i, s in enumerate(net.segId)}
        glob = np.array([pos[int(s)] for s in sub.segId])
        assert set(sub.lake_params) == set(net.lake_params)
        for k, v in sub.lake_params.items():
            assert v.shape == (sub.nRch,) and np.array_equal(v, net.lake_params[k][glob])

@pytest.mark.parametrize("div", [1.0, 4.0, 16.0])
def test_lower_mainstem_threshold_keeps_the_decomposition_valid(div):
    """decompose(mainstem_div): the threshold nRch / (nparts * div) only moves reaches between the mainstem and the tributary
    domains -- every reach is routed exactly once, the mainstem is closed under 'downstream of', tributary domains are closed
    under 'upstream of', every outlet drains into the mainstem, and a lower threshold gives shallower tributaries."""
    from mizuroute_b200 import partition, synth
    from mizuroute_b200.synth import _down_index
    net = synth.conus_like(20000, seed=3)
    dec = partition.decompose(net, 4, mainstem_div=div)
    down = _down_index(net)
    owner = np.full(net.nRch, -2)
    owner[dec.mainstem] = -1
    for k, t in enumerate(dec.trib):
        assert np.all(owner[t] == -2)
        owner[t] = k
    assert np.all(owner != -2)
    main = owner == -1
    has_down = down >= 0
    assert np.all(main[down[main & has_down]])                                  # downstream of a mainstem reach is mainstem
    t = ~main & has_down
    same = owner[down[t]] == owner[t]
    assert np.all(same | main[down[t]])                                         # a tributary reach drains within its domain or into the mainstem
    assert np.array_equal(np.sort(dec.outlets), np.flatnonzero(t & main[np.where(has_down, down, 0)]))
    if div > 1:
        ref = partition.decompose(net, 4, mainstem_div=1.0)
        assert dec.mainstem.size >= ref.mainstem.size and np.all(np.isin(ref.mainstem, dec.mainstem))
