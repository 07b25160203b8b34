"""High-density exciton devices ("Path XD", SURVEY.md §8 f4, BASELINE configs[3]): occupancy blocking and
exciton-exciton annihilation.  The reference has neither excitons nor annihilation (parity UNPINNED):
these tests check the CUDA kernel against this repository's own specification
(oracle/exciton_dense_oracle.c), bit for bit, plus properties the model must have."""
import numpy as np
import pytest

from oracle import exciton_oracle as xo
from oracle import frm_oracle as fo
from tests.emu import emu


def _oracle(dims, seed, params=None, props=(0.6, 0.3, 0.1), **kw):
    lat = fo.OracleLattice.generate(dims, props, seed).arrays()
    ws1, wt1 = xo.weights(lat["s1"], lat["t1"], lat["species"])
    return xo.DenseExcitonOracle(dims, lat["species"], ws1, wt1, params or xo.default_params(), **kw), lat


def _check_state(res, dims):
    """Occupancy mask == the exciton list, one exciton per site."""
    assert len(set(res["site"].tolist())) == len(res["site"])
    if "occupancy" in res:
        occ = res["occupancy"]
        assert int((occ != 0).sum()) == len(res["site"])
        assert np.array_equal(occ[res["site"]], res["spin"].astype(np.uint8))


def test_oracle_cache_is_exact_and_model_properties():
    """The specification's own cached totals equal a from-scratch evaluation after every event, and
    the physics switches do what they say."""
    dims = (10, 9, 6)
    o, _ = _oracle(dims, 3, dict(xo.default_params(), cutoff=2.0))
    r = o.run_device(11, 0, 40, 600, verify=True)
    assert r["cache_mismatches"] == 0
    t = r["tallies"]
    assert t["events"] == 600 and t["ann_ss"] + t["ann_st"] + t["ann_ts"] + t["ann_tt"] > 0     # crowded: 40 of 360 interior sites
    lost = sum(t[k] for k in t if k.startswith(("rad_", "nonrad_", "ann_")))
    assert t["generated"] == 40 + lost
    assert t["tta_up"] <= t["ann_tt"]
    # pure blocking: annihilation factors 0 => no annihilation events at all
    ob, _ = _oracle(dims, 3, dict(xo.default_params(), cutoff=2.0), annihilation=((0, 0), (0, 0)))
    tb = ob.run_device(11, 0, 40, 600, verify=True)
    assert tb["cache_mismatches"] == 0
    assert all(tb["tallies"][k] == 0 for k in ("ann_ss", "ann_st", "ann_ts", "ann_tt", "tta_up"))
    # a single exciton per device never meets anybody
    o1 = o.run_device(11, 5, 1, 300, verify=True)
    assert o1["cache_mismatches"] == 0 and all(o1["tallies"][k] == 0 for k in ("ann_ss", "ann_st", "ann_ts", "ann_tt"))


@pytest.mark.parametrize("dims,cutoff,seed,n,ann,p_tta", [
    ((10, 9, 6), 2.0, 4, 40, ((1.0, 1.0), (1.0, 1.0)), 0.25),      # crowded: many annihilations, wrap on every site
    ((12, 11, 7), 3.0, 5, 33, ((0.5, 2.0), (0.0, 1.0)), 0.6),       # two slots of excitons, mixed factors, TS blocked
    ((9, 9, 5), 1.5, 6, 7, ((0.0, 0.0), (0.0, 0.0)), 0.0),          # pure blocking
    ((22, 20, 6), 4.0, 7, 64, ((1.0, 1.0), (1.0, 1.0)), 0.25),      # the cutoff-4 table, wrap-free sites exist
    ((11, 10, 6), 3.0, 8, 96, ((1.0, 1.0), (1.0, 3.0)), 1.0)])      # very crowded, every TT annihilation promotes the survivor
def test_emulated_dense_kernel_equals_specification(dims, cutoff, seed, n, ann, p_tta):
    p = dict(xo.default_params(), cutoff=cutoff)
    o, _ = _oracle(dims, seed, p, annihilation=ann, p_tta=p_tta)
    n_events = 300
    dev = emu.xd_run(o, seed * 100 + 3, 17, n, n_events // 2, n_calls=2)
    ref = o.run_device(seed * 100 + 3, 17, n, n_events, verify=True)
    assert ref["cache_mismatches"] == 0
    tr = dev["trace"]
    assert np.array_equal(tr["kind"], ref["trace"]["kind"])
    assert np.array_equal(tr["initial"], ref["trace"]["initial"]) and np.array_equal(tr["final"], ref["trace"]["final"])
    assert np.array_equal((tr["spins"] & 0xffff).astype(np.int32), ref["trace"]["channel"])
    assert np.array_equal((tr["spins"] >> 16).astype(np.int32), ref["trace"]["exciton"])
    assert np.array_equal(tr["pad"][:, 0], ref["trace"]["spin"])
    assert np.array_equal(tr["tau"], ref["trace"]["dt"])                         # same libm under emulation
    assert np.array_equal(dev["site"], ref["site"]) and np.array_equal(dev["spin"], ref["spin"])
    assert np.array_equal(dev["total_rate"], ref["total_rate"])
    assert (dev["time"], dev["lifetime_sum"], dev["draws"]) == (ref["time"], ref["lifetime_sum"], ref["draws"])
    names = [k for k in xo.XDT_NAMES]
    for i, k in enumerate(names):
        if not k.startswith("_"):
            assert int(dev["totals"][i]) == ref["tallies"][k], k
    _check_state(dev, dims)
    if any(v > 0 for row in ann for v in row):
        assert sum(ref["tallies"][k] for k in ("ann_ss", "ann_st", "ann_ts", "ann_tt")) > 0


@pytest.mark.gpu
def test_gpu_dense_devices_equal_specification():
    """Through the C-ABI: 96 devices x 48 excitons on 2 realisations; traces of 8 devices, every
    final device state and the per-realisation tallies equal the specification run on the device's own
    Boltzmann weights."""
    from hyperfluorescence_b200 import _native as nat
    dims, props, seed, B, R, n, n_events = (20, 18, 8), (0.6, 0.3, 0.1), 91, 96, 2, 48, 400
    sim = nat.Sim(dims, props, 0.1, 1, 1, n_trajectories=B, n_realisations=R, seed=seed, first_trajectory=500)
    try:
        sim.generate_lattice()
        sim.x_configure(nat.x_params_from_dict(dict(cutoff=3.0)))
        ann = ((1.0, 0.5), (2.0, 1.0))
        sim.xd_configure(n, ann, 0.3)
        sim.set_trace(0, 8, n_events)
        sim.xd_spawn()
        sim.xd_run(n_events // 4)
        sim.xd_run(n_events - n_events // 4)
        sim.sync()
        tab = sim.x_tables()
        params = dict(nat.x_params_to_dict(nat.x_default_params()), cutoff=3.0)
        per_real = sim.xd_realisation_tallies()
        for r in range(R):
            lat = sim.download_lattice(r)
            ws1, wt1 = sim.x_weights(r)
            o = xo.DenseExcitonOracle(dims, lat["species"], ws1, wt1, params, tab, annihilation=ann, p_tta=0.3)
            lo, hi = r * (B // R), (r + 1) * (B // R)
            acc = {}
            for d in range(lo, hi):
                ref = o.run_device(seed, 500 + d, n, n_events, verify=(d % 16 == 0))
                assert ref["cache_mismatches"] == 0
                got = sim.xd_device(d)
                assert np.array_equal(got["site"], ref["site"]) and np.array_equal(got["spin"], ref["spin"]), d
                assert got["draws"] == ref["draws"]
                np.testing.assert_allclose(got["time"], ref["time"], rtol=1e-12)
                np.testing.assert_allclose(got["lifetime_sum"], ref["lifetime_sum"], rtol=1e-11)
                assert np.array_equal(got["total_rate"], ref["total_rate"])             # no transcendental on the rate path
                occ = sim.xd_occupancy(d)
                assert int((occ != 0).sum()) == n and np.array_equal(occ[got["site"]], got["spin"].astype(np.uint8))
                if d < 8:
                    tr = sim.x_trace(d, n_events)
                    assert np.array_equal(tr["kind"], ref["trace"]["kind"]) and np.array_equal(tr["final"], ref["trace"]["final"])
                    assert np.array_equal((tr["spins"] & 0xffff).astype(np.int32), ref["trace"]["channel"])
                    assert np.array_equal((tr["spins"] >> 16).astype(np.int32), ref["trace"]["exciton"])
                    np.testing.assert_allclose(tr["tau"], ref["trace"]["dt"], rtol=1e-12)
                for k, v in ref["tallies"].items():
                    acc[k] = acc.get(k, 0) + v
            for k, v in acc.items():
                assert int(per_real[k][r]) == v, (r, k)
        tot = sim.xd_tallies()
        assert tot["events"] == B * n_events
        assert tot["ann_ss"] + tot["ann_st"] + tot["ann_ts"] + tot["ann_tt"] > 0
    finally:
        sim.close()


@pytest.mark.gpu
def test_gpu_dense_ensemble_api():
    from hyperfluorescence_b200.excitons import DenseExcitonEnsemble
    ens = DenseExcitonEnsemble((32, 32, 16), devices=256, excitons_per_device=128, seed=9)
    ens.run(2000)
    s = ens.summary()
    assert s["events"] == 256 * 2000
    assert s["annihilated"] > 0 and 0.0 < s["annihilation_fraction"] < 1.0
    assert abs(sum(s["channels"].values()) - 1.0) < 1e-12
    # the same devices with annihilation switched off lose nothing to it and emit more per exciton
    blk = DenseExcitonEnsemble((32, 32, 16), devices=256, excitons_per_device=128, seed=9, annihilation=((0, 0), (0, 0)))
    blk.run(2000)
    sb = blk.summary()
    assert sb["annihilated"] == 0 and sb["iqe"] > s["iqe"]
    with pytest.raises(ValueError):
        DenseExcitonEnsemble((8, 8, 4), devices=1, excitons_per_device=2000)
    ens.close(); blk.close()


@pytest.mark.gpu
def test_gpu_dense_full_size_properties():
    """BASELINE configs[3] at size (128^3, 2048 devices x 256 excitons): conservation laws, mask == list on
    sampled devices, and two sampled devices against the specification (with its cache verification)."""
    from hyperfluorescence_b200 import _native as nat
    dims, B, n, n_events = (128, 128, 128), 2048, 256, 150
    sim = nat.Sim(dims, (0.84, 0.15, 0.01), 0.1, 1, 1, n_trajectories=B, seed=13)
    try:
        sim.generate_lattice()
        sim.x_configure()
        sim.xd_configure(n)
        sim.xd_spawn()
        sim.xd_run(n_events)
        sim.sync()
        t = sim.xd_tallies()
        lost = sum(v for k, v in t.items() if k.startswith(("rad_", "nonrad_", "ann_")))
        assert t["events"] == B * n_events and t["generated"] == B * n + lost
        assert t["forster"] + t["dexter"] + t["isc"] + t["risc"] + lost == t["events"]
        assert t["tta_up"] <= t["ann_tt"]
        lat = sim.download_lattice(0)
        ws1, wt1 = sim.x_weights(0)
        o = xo.DenseExcitonOracle(dims, lat["species"], ws1, wt1, nat.x_params_to_dict(nat.x_default_params()), sim.x_tables())
        for d in (0, B - 1):
            got = sim.xd_device(d)
            occ = sim.xd_occupancy(d)
            assert int((occ != 0).sum()) == n and np.array_equal(occ[got["site"]], got["spin"].astype(np.uint8))
            ref = o.run_device(13, d, n, n_events, verify=(d == 0))
            assert ref["cache_mismatches"] == 0
            assert np.array_equal(got["site"], ref["site"]) and np.array_equal(got["spin"], ref["spin"])
            assert np.array_equal(got["total_rate"], ref["total_rate"]) and got["draws"] == ref["draws"]
    finally:
        sim.close()


@pytest.mark.gpu
@pytest.mark.parametrize("dims,cutoff,n,B,R", [((12, 12, 5), 2.0, 1, 5, 1),         # one exciton per device: never meets anybody
                                                 ((40, 40, 12), 3.0, 1024, 3, 1),     # the largest device (32 slots of excitons)
                                                 ((14, 13, 6), 4.0, 33, 9, 3),        # cutoff 4 with wrap everywhere, 3 devices per realisation
                                                 ((26, 24, 7), 4.0, 200, 6, 2),       # cutoff 4 with wrap-free sites: the static fast path, crowded
                                                 ((16, 16, 8), 2.5, 100, 7, 7)])      # one device per realisation
def test_gpu_dense_edge_shapes(dims, cutoff, n, B, R):
    """Extreme shapes of the high-density devices against the specification (with its cache verification)."""
    from hyperfluorescence_b200 import _native as nat
    n_events = 150
    sim = nat.Sim(dims, (0.7, 0.2, 0.1), 0.1, 1, 1, n_trajectories=B, n_realisations=R, seed=23, first_trajectory=11)
    try:
        sim.generate_lattice()
        params = dict(nat.x_params_to_dict(nat.x_default_params()), cutoff=cutoff)
        sim.x_configure(nat.x_params_from_dict(dict(cutoff=cutoff)))
        ann = ((1.0, 2.0), (0.5, 1.0))
        sim.xd_configure(n, ann, 0.4)
        sim.xd_spawn()
        sim.xd_run(n_events)
        sim.sync()
        tab = sim.x_tables()
        per_real = sim.xd_realisation_tallies()
        for r in range(R):
            lat = sim.download_lattice(r)
            ws1, wt1 = sim.x_weights(r)
            o = xo.DenseExcitonOracle(dims, lat["species"], ws1, wt1, params, tab, annihilation=ann, p_tta=0.4)
            acc = {}
            for d in range(r * (B // R), (r + 1) * (B // R)):
                ref = o.run_device(23, 11 + d, n, n_events, verify=(n <= 100))
                assert ref["cache_mismatches"] == 0
                got = sim.xd_device(d)
                assert np.array_equal(got["site"], ref["site"]) and np.array_equal(got["spin"], ref["spin"]), d
                assert np.array_equal(got["total_rate"], ref["total_rate"]) and got["draws"] == ref["draws"]
                for k, v in ref["tallies"].items():
                    acc[k] = acc.get(k, 0) + v
            for k, v in acc.items():
                assert int(per_real[k][r]) == v, (r, k)
        if n == 1:
            t = sim.xd_tallies()
            assert t["ann_ss"] + t["ann_st"] + t["ann_ts"] + t["ann_tt"] == 0
    finally:
        sim.close()
