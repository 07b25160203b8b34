"""GPU parity suite (`-m gpu`): every test drives the CUDA path through the C-ABI
(`hyperfluorescence_b200._native.Sim` = ctypes over libhfkmc.so, or the `Lattice` façade on top
of it) and checks it against (a) the golden vectors frozen from the UNMODIFIED reference and
(b) the C oracle (oracle/frm_oracle.c, itself pinned bit-exactly to those vectors).

Bars (BASELINE.json north_star): event-index sequences bit-exact; rates and tau within 1e-12
relative (CUDA exp/log vs glibc); tallies exact; statistics within 3 sigma.
"""
import numpy as np
import pytest

from tests.conftest import golden_names, load_golden

pytestmark = pytest.mark.gpu

TAU_RTOL = 1e-12
STATUS = {"": 0, "ZeroDivisionError": 2, "ValueError": 3}


@pytest.fixture(scope="module")
def nat():
    from hyperfluorescence_b200 import _native
    lib = _native.load()
    if lib.hf_device_count() < 1:
        pytest.fail("no sm_100 device visible: the GPU suite must run on a B200")
    return _native


def _sim_for(nat, g, generated=False, n_trajectories=1):
    sim = nat.Sim(g["dims"], tuple(float(p) for p in g["props"]), g["field_f"], g["charges"], g["distance"],
                  n_trajectories=n_trajectories, n_realisations=1, seed=g["seed_int"],
                  first_trajectory=g["trajectory"], first_realisation=g["realisation"])
    if generated:
        sim.generate_lattice()
    else:
        sim.upload_lattice(0, g["species"], g["homo"], g["lumo"], g["s1"], g["t1"])
    return sim


def _assert_tau(a, b, clock=0.0):
    """Freshly drawn waiting times (and rates) must agree to 1e-12 relative.  A *residual* tau
    (reseau.py:564-578 subtracts every fired tau from every pending one) inherits the absolute
    error of the chain of subtractions, which is bounded by the last-bit differences between
    CUDA's and glibc's exp/log times the simulation clock, so residuals are compared to 1e-12 of
    the clock (sum of fired taus so far) on top of 1e-12 relative."""
    a, b = np.asarray(a, float), np.asarray(b, float)
    assert a.shape == b.shape
    tol = TAU_RTOL * np.abs(b) + TAU_RTOL * np.asarray(clock, float)
    bad = np.abs(a - b) > tol
    assert not bad.any(), (np.flatnonzero(bad)[:5], a[bad][:5], b[bad][:5])


def _check_run(nat, g, sim):
    if "replay_traj" in g:
        sim.set_replay(g["replay_traj"])
    sim.inject()
    sim.sync()
    # state right after Lattice.__init__
    info = sim.info(0)
    assert info.status == 0
    assert np.array_equal(sim.positions(0, 0), g["init_e_pos"])
    assert np.array_equal(sim.positions(0, 1), g["init_h_pos"])
    for which, k in ((0, "me"), (1, "mh")):
        ini, fin, tau = sim.move_events(0, which)
        assert np.array_equal(ini, g[f"init_{k}_init"])
        assert np.array_equal(fin, g[f"init_{k}_final"])
        _assert_tau(tau, g[f"init_{k}_tau"])
    assert info.draws == int(g["init_draws"][0])
    assert info.injection == 2 * g["charges"]
    # the step loop, in two launches to exercise state save / restore
    n = g["n_steps"]
    sim.set_trace(0, 1, n)
    first = n // 3
    sim.run(first)
    info_mid = sim.info(0)
    if info_mid.status == 0:
        sim.run(n - first)
    sim.sync()
    tr = sim.trace(0, n)
    total = len(g["trace_kind"])
    assert len(tr) == total
    assert np.array_equal(tr["kind"], g["trace_kind"])
    assert np.array_equal(tr["particule"], g["trace_particule"])
    assert np.array_equal(tr["initial"], g["trace_initial"])
    assert np.array_equal(tr["final"], g["trace_final"])
    assert np.array_equal(tr["draws"].astype(np.int64), g["trace_draws"])
    assert np.array_equal(tr["spins"].astype(np.int64), g["trace_spins"])
    clock = np.cumsum(g["trace_tau"])
    _assert_tau(tr["tau"], g["trace_tau"], clock)
    # final state
    info = sim.info(0)
    assert info.status == STATUS[g["error_s"]]
    assert np.array_equal(sim.positions(0, 0), g["final_e_pos"])
    assert np.array_equal(sim.positions(0, 1), g["final_h_pos"])
    assert np.array_equal(sim.positions(0, 2), g["final_x_pos"])
    for which, k in ((0, "me"), (1, "mh")):
        ini, fin, tau = sim.move_events(0, which)
        assert np.array_equal(ini, g[f"final_{k}_init"])
        assert np.array_equal(fin, g[f"final_{k}_final"])
        _assert_tau(tau, g[f"final_{k}_tau"], clock[-1] if total else 0.0)
    assert [info.emission, info.recombination, info.injection, info.steps] == g["final_tallies"].tolist()
    assert [info.draws, info.spins] == g["final_draws"][:2].tolist()
    t = sim.tallies()
    assert [t["emission"], t["recombination"], t["injection"], t["steps"]] == g["final_tallies"].tolist()
    assert t["fired"] == total
    assert t["move_e"] + t["move_h"] + t["bound"] + t["decay"] + t["capture_e"] + t["capture_h"] == total
    assert t["move_e"] == int(np.sum((g["trace_kind"] == 1) & (g["trace_particule"] == 1)))
    assert t["decay"] == t["recombination"] == t["recomb_host"] + t["recomb_tadf"] + t["recomb_fluo"]
    assert t["emit_tadf"] + t["emit_fluo"] == t["emission"]
    # Lattice._cache: the last ten fired events (the reference also pushes the event of a step that raised)
    last = sim.last_events(0)
    if not g["error_s"]:
        k = min(10, total)
        assert np.array_equal(last["initial"][-k:], g["trace_initial"][-k:])
        assert np.array_equal(last["kind"][-k:], g["trace_kind"][-k:])


def test_golden_uploaded_lattice(nat, golden):
    """Reference-built lattice uploaded through hf_upload_lattice: bit-identical energies, so only
    exp/log separate the device from the reference."""
    sim = _sim_for(nat, golden, generated=False)
    try:
        _check_run(nat, golden, sim)
    finally:
        sim.close()


def test_golden_generated_lattice(nat, golden):
    """Device-side lattice generation (species shuffle + Gaussian energies) from the Philox
    streams, then the same trajectory."""
    g = golden
    sim = _sim_for(nat, g, generated=True)
    try:
        a = sim.download_lattice(0)
        assert np.array_equal(a["species"], g["species"])
        for k in ("homo", "lumo", "s1", "t1"):
            np.testing.assert_allclose(a[k], g[k], rtol=0, atol=2e-15)
        _check_run(nat, g, sim)
    finally:
        sim.close()


def test_hop_rates_match_oracle(nat):
    """Known-answer check of _time_move_electron / _time_move_hole on hand-built hops
    (SURVEY.md §4): every neighbour hop of every charge of a 10+10-charge configuration, rates
    and tau within 1e-12 of the oracle, ZeroDivisionError flagged alike."""
    from oracle import frm_oracle as fo
    from oracle import philox as px
    g = load_golden("c10_s1")
    sim = _sim_for(nat, g)
    sim.inject()
    lat = fo.OracleLattice.from_arrays(g["dims"], g["species"], g["homo"], g["lumo"])
    base = fo.OracleTrajectory(lat, g["charges"], g["seed_int"], g["trajectory"], g["field_f"])
    n0 = base.tallies()["draws"]
    construction = np.array([px.uniform(g["seed_int"], px.STREAM_TRAJ, g["trajectory"], d) for d in range(n0)])
    rng = np.random.default_rng(5)
    try:
        for particule, which in ((1, "electrons"), (2, "holes")):
            ini, fin = [], []
            for s in base.positions(which):
                for f in lat.neighbourhood(int(s)):
                    ini.append(int(s)); fin.append(int(f))
            u = rng.random(len(ini))
            u[::17] = 0.0                                           # rng = 1.0 -> tau = -0.0
            tau, rate, zd = sim.eval_hops(0, particule, ini, fin, u)
            for i in range(len(ini)):
                # the oracle draws the hop's uniform from the trajectory stream: replay the
                # construction draws, then u[i]
                o = fo.OracleTrajectory(lat, g["charges"], g["seed_int"], g["trajectory"], g["field_f"],
                                        replay=np.append(construction, u[i]))
                try:
                    expect = o.time_move(ini[i], fin[i], particule)
                except ZeroDivisionError:
                    assert zd[i]
                    continue
                assert not zd[i]
                np.testing.assert_allclose(tau[i], expect, rtol=TAU_RTOL, atol=0)
                assert 0.0 < rate[i] <= 1e13
    finally:
        sim.close()


def test_batch_matches_oracle_per_trajectory(nat):
    """64 trajectories sharing one realisation in one launch: every trajectory equals the oracle's
    run of the same global trajectory id (tallies, draw counts, final positions)."""
    from oracle import frm_oracle as fo
    g = load_golden("d0_s2")
    B, steps = 64, 1200
    sim = _sim_for(nat, g, n_trajectories=B)
    lat = fo.OracleLattice.from_arrays(g["dims"], g["species"], g["homo"], g["lumo"])
    try:
        sim.inject()
        sim.set_trace(0, B, steps)
        sim.run(steps)
        sim.sync()
        tot = dict(emission=0, recombination=0, injection=0, steps=0, fired=0)
        for t in range(B):
            o = fo.OracleTrajectory(lat, g["charges"], g["seed_int"], t, g["field_f"])
            n, otr = o.run(steps)
            info = sim.info(t)
            ot = o.tallies()
            assert info.fired == n
            assert [info.emission, info.recombination, info.injection, info.steps, info.draws, info.spins] == \
                [ot["emission"], ot["recombination"], ot["injection"], ot["step"], ot["draws"], ot["spins"]]
            assert info.status == {"ok": 0, "no_events": 1, "ZeroDivisionError": 2, "ValueError": 3}[o.status]
            tr = sim.trace(t, steps)
            assert np.array_equal(tr["initial"], otr["initial"]) and np.array_equal(tr["final"], otr["final"])
            assert np.array_equal(tr["kind"], otr["kind"])
            _assert_tau(tr["tau"], otr["tau"], np.cumsum(otr["tau"]))
            assert np.array_equal(sim.positions(t, 0), o.positions("electrons"))
            assert np.array_equal(sim.positions(t, 1), o.positions("holes"))
            flags = o.flags()
            assert np.array_equal(np.sort(sim.flags(t, 0)), np.flatnonzero(flags & 1))
            assert np.array_equal(np.sort(sim.flags(t, 1)), np.flatnonzero(flags & 2))
            for k in tot:
                tot[k] += {"emission": ot["emission"], "recombination": ot["recombination"],
                           "injection": ot["injection"], "steps": ot["step"], "fired": n}[k]
        got = sim.tallies()
        for k, v in tot.items():
            assert got[k] == v, k
    finally:
        sim.close()


@pytest.mark.parametrize("dims,charges,steps", [((16, 16, 8), 40, 400), ((12, 10, 6), 24, 600), ((20, 20, 20), 100, 150)])
def test_many_charges_match_oracle(nat, dims, charges, steps):
    """charges > 12 run on the warp kernel, whose candidate loop screens in fp32 and shares the O(charges) Coulomb
    sum of the few candidates it verifies among the lanes: every trajectory must still equal the C restatement of
    the reference (pinned to the golden vectors) event for event, also where the reference's quirks stop it."""
    from oracle import frm_oracle as fo
    props, seed, B = (0.84, 0.15, 0.01), 4242, 24
    olat = fo.OracleLattice.generate(dims, props, seed)
    sim = nat.Sim(dims, props, 0.1, charges, 1, n_trajectories=B, seed=seed, first_trajectory=7)
    try:
        sim.generate_lattice()
        sim.inject()
        sim.set_trace(0, B, steps)
        sim.run(steps)
        sim.sync()
        stopped = 0
        for t in range(B):
            o = fo.OracleTrajectory(olat, charges, seed, 7 + t, 0.1)
            n, otr = o.run(steps)
            info = sim.info(t)
            ot = o.tallies()
            assert info.fired == n, t
            assert [info.emission, info.recombination, info.injection, info.steps, info.draws, info.spins] == \
                [ot["emission"], ot["recombination"], ot["injection"], ot["step"], ot["draws"], ot["spins"]], t
            assert info.status == {"ok": 0, "no_events": 1, "ZeroDivisionError": 2, "ValueError": 3}[o.status]
            stopped += o.status != "ok"
            tr = sim.trace(t, steps)
            assert np.array_equal(tr["initial"], otr["initial"]) and np.array_equal(tr["final"], otr["final"])
            assert np.array_equal(tr["kind"], otr["kind"])
            _assert_tau(tr["tau"], otr["tau"], np.cumsum(otr["tau"]))
            assert np.array_equal(sim.positions(t, 0), o.positions("electrons"))
            assert np.array_equal(sim.positions(t, 1), o.positions("holes"))
    finally:
        sim.close()


@pytest.mark.parametrize("name,steps,charges", [("tiny333_s5", 1500, 1), ("c1_s1", 2000, 1),
                                                ("d0_s1", 1500, 4), ("d0_s2", 1200, 2), ("rag456_s6", 1000, 8),
                                                ("dense_s8", 1500, 6), ("tiny333_s5", 1500, 3),
                                                ("c10_s1", 600, 10), ("dense_s8", 800, 12)])
def test_thread_kernel_equals_warp_kernel(nat, name, steps, charges):
    """charges <= 12 has two step kernels (thread-per-trajectory by default: hf_frm_c1.cuh for one charge,
    hf_frm_small.cuh for 2..12; warp-per-trajectory with HF_FLAG_WARP_KERNEL): 4096 trajectories must end in
    exactly the same state under both — including the ones the reference's quirks stop (ZeroDivisionError,
    ValueError) on crowded little lattices."""
    g = load_golden(name)
    B = 4096
    def run(flags):
        sim = nat.Sim(g["dims"], tuple(float(p) for p in g["props"]), g["field_f"], charges, 1, n_trajectories=B,
                      seed=977, first_trajectory=10, flags=flags)
        sim.upload_lattice(0, g["species"], g["homo"], g["lumo"])
        sim.inject()
        sim.set_trace(0, 8, steps)
        sim.run(steps // 2)
        sim.run(steps - steps // 2)
        sim.sync()
        out = dict(tallies=sim.tallies(), traces=[sim.trace(i, steps) for i in range(8)],
                   info=[(lambda i: (i.draws, i.spins, i.emission, i.recombination, i.injection, i.steps, i.fired,
                                     i.status, i.special_kind, i.special_site, i.exciton_state, i.n_eflags, i.n_hflags))(sim.info(t))
                         for t in range(0, B, 37)],
                   pos=[(sim.positions(t, 0).tolist(), sim.positions(t, 1).tolist(), sim.flags(t, 0).tolist(),
                         sim.flags(t, 1).tolist(), [a.tolist() for a in sim.move_events(t, 0)],
                         [a.tolist() for a in sim.move_events(t, 1)],
                         [sim.last_events(t)[k].tolist() for k in ("initial", "final", "tau", "kind", "particule")])
                        for t in range(0, B, 211)])
        sim.close()
        return out
    thread, warp = run(0), run(1)
    assert thread["tallies"] == warp["tallies"]
    assert thread["info"] == warp["info"]
    assert thread["pos"] == warp["pos"]
    for a, b in zip(thread["traces"], warp["traces"]):
        assert np.array_equal(a, b)                      # same device libm: bit-identical, tau included
    if name == "tiny333_s5" and charges == 1:
        assert thread["tallies"]["recombination"] > 100_000 and thread["tallies"]["capture_e"] > 1000
    if charges > 1 and name != "c10_s1":                  # (20^3 with 10 charges: hardly any meeting in 600 steps)
        assert thread["tallies"]["recombination"] > 1000


@pytest.mark.parametrize("charges,dims", [(2, (10, 10, 5)), (4, (10, 10, 5)), (12, (12, 9, 6)), (5, (6, 5, 4))])
def test_thread_inject_equals_warp_inject(nat, charges, dims, monkeypatch):
    """hf_inject for 2..12 charges runs one thread per trajectory (hf_inject_small_kernel) when random.sample takes its
    set variant; HF_INJECT_WARP=1 keeps the warp kernel: positions, flags, initial events (taus bit for bit), draw
    counts and status of 2048 trajectories must be identical (reseau.py:207-276)."""
    out = []
    for warp in ("1", None):
        if warp: monkeypatch.setenv("HF_INJECT_WARP", warp)
        else: monkeypatch.delenv("HF_INJECT_WARP", raising=False)
        sim = nat.Sim(dims, (0.84, 0.15, 0.01), 0.1, charges, 1, n_trajectories=2048, seed=4242, first_trajectory=3)
        sim.generate_lattice(); sim.inject(); sim.sync()
        out.append([(sim.positions(t, 0).tolist(), sim.positions(t, 1).tolist(), sorted(sim.flags(t, 0).tolist()),
                     sorted(sim.flags(t, 1).tolist()), [a.tolist() for a in sim.move_events(t, 0)],
                     [a.tolist() for a in sim.move_events(t, 1)],
                     (lambda i: (i.draws, i.injection, i.status))(sim.info(t))) for t in range(0, 2048, 13)])
        sim.close()
    assert out[0] == out[1]
    assert any(len(o[4][0]) == charges for o in out[0])


def test_sharding_is_invisible(nat):
    """Two handles holding trajectories [0,96) and [96,192) give the same box totals as one handle
    holding [0,192): Philox ids are global, so results do not depend on how ranks split the work."""
    g = load_golden("rag456_s6")
    def totals(first, n):
        sim = nat.Sim(g["dims"], tuple(g["props"]), g["field_f"], g["charges"], 1, n_trajectories=n,
                      seed=g["seed_int"], first_trajectory=first)
        try:
            sim.upload_lattice(0, g["species"], g["homo"], g["lumo"])
            sim.inject()
            sim.run(700)
            return sim.tallies()
        finally:
            sim.close()
    whole, a, b = totals(0, 192), totals(0, 96), totals(96, 96)
    for k in whole:
        assert whole[k] == a[k] + b[k], k
    assert whole["recombination"] > 100


def test_facade_matches_oracle_and_oled_driver(nat, capsys):
    """The drop-in surface: Lattice(...) / operations / get_IQE / _emission / _injection / _cache /
    get_particules_positions, and the OLED.py driver, against the oracle for the same seed."""
    from hyperfluorescence_b200 import OLED, Lattice
    from hyperfluorescence_b200.event import Event, Point
    from hyperfluorescence_b200.molecule import Fluorescent, Host, TADF
    from oracle import frm_oracle as fo
    dims, props, seed = (10, 10, 5), (0.84, 0.15, 0.01), 1234
    lat = Lattice(dims, props, charges=4, seed=seed)
    lat.operations(100)
    olat = fo.OracleLattice.generate(dims, props, seed)
    otr = fo.OracleTrajectory(olat, 4, seed, 0, 0.1)
    n, tr = otr.run(100)
    ot = otr.tallies()
    assert lat.get_IQE() == ot["IQE"]
    assert (lat._emission, lat._recombination, lat._injection, lat._step) == \
        (ot["emission"], ot["recombination"], ot["injection"], ot["step"])
    cache = lat._cache
    assert len(cache) == 10 and all(isinstance(e, Event) for e in cache)
    assert [lat._site(e.final) for e in cache] == tr["final"][-10:].tolist()
    e, h, x = lat.get_particules_positions()
    assert [lat._site(p) for p, _ in e] == otr.positions("electrons").tolist()
    assert all(k in (Host, TADF, Fluorescent) for _, k in e + h + x)
    mol = lat._grid[0][0][0]
    assert isinstance(mol, Host) and mol.position == Point(0, 0, 0)
    assert [(p.x, p.y, p.z) for p in mol.neighbourhood[:6]] == [(0, 0, 1), (0, 1, 0), (0, 1, 1), (0, 9, 0), (0, 9, 1), (1, 0, 0)]
    a = olat.arrays()
    np.testing.assert_allclose(mol.homo_energy, a["homo"][0], atol=2e-15)
    # reference-style exceptions before any device work
    with pytest.raises(TypeError):
        Lattice([10, 10, 5], props)
    with pytest.raises(IndexError):
        Lattice((10, 10), props)
    with pytest.raises(ValueError):
        Lattice((10, 10, 2), props)
    with pytest.raises(ValueError):
        Lattice((3, 3, 3), props)
    with pytest.raises(ValueError):
        Lattice((4, 4, 8), (0.5, 0.3, 0.2), charges=17)
    lat.close()
    test = OLED.main(seed=7)
    out = capsys.readouterr().out
    assert "IQE :" in out and "emissions :" in out and "injections :" in out
    assert test._injection >= 8
    test.close()


@pytest.mark.parametrize("dims,props,seed", [((3, 3, 3), (0.5, 0.3, 0.2), 777), ((4, 4, 3), (0.5, 0.3, 0.2), 778)])
def test_adjacent_charges_screening_is_exact(nat, dims, props, seed):
    """The fp32 screening pass of hf_c1_gen neglects the Coulomb term of the opposite charge (hf_frm_c1.cuh: at most
    C_e / kT = 1.9e-5 relative, most with the two carriers on neighbouring sites).  On these lattices the carriers
    are neighbours most of the time: > 10^6 events regenerated next to the opposite charge, every trajectory compared
    with the C restatement, which evaluates all candidates exactly (reseau.py:283-345, 380-396)."""
    from oracle import frm_oracle as fo
    B, steps = 8192, 800
    sim = nat.Sim(dims, props, 0.1, 1, 1, n_trajectories=B, seed=seed)
    try:
        sim.generate_lattice()
        sim.inject()
        sim.run(steps)
        sim.sync()
        lat = sim.download_lattice(0)
        infos = sim.infos(0, B)
        pos = [(sim.positions(t, 0).tolist(), sim.positions(t, 1).tolist(), sim.move_events(t, 0), sim.move_events(t, 1))
               for t in range(0, B, 64)]
        tallies = sim.tallies()
    finally:
        sim.close()
    olat = fo.OracleLattice.from_arrays(dims, lat["species"], lat["homo"], lat["lumo"])
    X, Y = dims[0], dims[1]
    near = moves = 0
    for t in range(B):
        o = fo.OracleTrajectory(olat, 1, seed, t, 0.1)
        if t % 64 == 0 and t < 2048:
            # how often the regenerated charge sits next to the opposite one (sampled on 32 trajectories)
            for _ in range(steps):
                n, tr = o.run(1)
                if n and tr["kind"][0] == 1:
                    e, h = o.positions("electrons"), o.positions("holes")
                    if len(e) and len(h):
                        moves += 1
                        de = [abs(int(e[0]) % X - int(h[0]) % X), abs(int(e[0]) // X % Y - int(h[0]) // X % Y),
                              abs(int(e[0]) // (X * Y) - int(h[0]) // (X * Y))]
                        near += max(de) <= 1
        else:
            o.run(steps, trace=False)
        ot, info = o.tallies(), infos[t]
        assert [int(info[k]) for k in ("emission", "recombination", "injection", "steps", "draws", "spins", "status")] == \
            [ot["emission"], ot["recombination"], ot["injection"], ot["step"], ot["draws"], ot["spins"], 0], t
        if t % 64 == 0:
            e, h, me, mh = pos[t // 64]
            assert e == o.positions("electrons").tolist() and h == o.positions("holes").tolist()
            for got, name in ((me, "move_e"), (mh, "move_h")):
                ev = o.events(name)
                assert np.array_equal(got[1], ev["final"])
                _assert_tau(got[2], ev["tau"], 1e-9)
    regenerated_near = near / max(moves, 1) * (tallies["move_e"] + tallies["move_h"])
    assert regenerated_near > 1e6, (near, moves, regenerated_near)


def test_channel_fractions_over_1e5_trajectories(nat):
    """The north_star's statistical gate: EQE (here IQE, the only efficiency the reference defines, reseau.py:595)
    and CHANNEL FRACTIONS over 10^5 trajectories within 3 sigma of the unmodified reference —
    tests/golden/stats_d0_1e5.npz holds 10^5 unmodified runs of OLED.py's own device (OLED.py:6-11: 10x10x5,
    charges = 4, operations(100)) with the species every exciton decayed on and whether it emitted
    (make_golden.py --stats-d0-1e5).  Device: 10^5 trajectories, one lattice realisation each, like the reference."""
    import os
    from tests.conftest import GOLDEN_DIR
    ref = np.load(os.path.join(GOLDEN_DIR, "stats_d0_1e5.npz"))
    dims = tuple(int(v) for v in ref["dims"])
    props = tuple(float(v) for v in ref["props"])
    charges, steps = int(ref["charges"]), int(ref["steps"])
    B = 100_000
    sim = nat.Sim(dims, props, 0.1, charges, 1, n_trajectories=B, n_realisations=B, seed=20261020)
    try:
        sim.generate_lattice()
        sim.inject()
        sim.run(steps)
        sim.sync()
        infos = sim.infos(0, B)
        t = sim.tallies()
    finally:
        sim.close()
    em, rec, inj = (infos[k].astype(float) for k in ("emission", "recombination", "injection"))
    n_ref = ref["emission"].size
    assert n_ref >= 100_000
    # per-run means: recombinations, emissions, IQE (a swallowed ZeroDivisionError leaves IQE at 0: reseau.py:585-595)
    stopped = infos["status"] != 0
    iqe = np.where(stopped, 0.0, 200.0 * em / inj)
    for name, ours, theirs in (("recombination", rec, ref["recombination"].astype(float)),
                               ("emission", em, ref["emission"].astype(float)),
                               ("IQE", iqe, ref["iqe"].astype(float))):
        sigma = np.sqrt(ours.var(ddof=1) / ours.size + theirs.var(ddof=1) / theirs.size)
        assert abs(ours.mean() - theirs.mean()) < 3 * sigma, (name, ours.mean(), theirs.mean(), sigma)
    # channel fractions: where excitons decay and where photons come from (binomial proportions, pooled sigma)
    ref_rec = float(ref["recombination"].sum())
    ref_em = float(ref["emission"].sum())
    for name, k_ours, n_ours, k_ref, n_r in (
            ("recomb_host", t["recomb_host"], t["recombination"], float(ref["recomb_host"].sum()), ref_rec),
            ("recomb_tadf", t["recomb_tadf"], t["recombination"], float(ref["recomb_tadf"].sum()), ref_rec),
            ("recomb_fluo", t["recomb_fluo"], t["recombination"], float(ref["recomb_fluo"].sum()), ref_rec),
            ("emit_tadf", t["emit_tadf"], t["emission"], float(ref["emit_tadf"].sum()), ref_em),
            ("emit_fluo", t["emit_fluo"], t["emission"], float(ref["emit_fluo"].sum()), ref_em),
            ("photons_per_recombination", t["emission"], t["recombination"], ref_em, ref_rec)):
        p1, p2 = k_ours / n_ours, k_ref / n_r
        pool = (k_ours + k_ref) / (n_ours + n_r)
        sigma = np.sqrt(pool * (1 - pool) * (1 / n_ours + 1 / n_r))
        assert abs(p1 - p2) < 3 * sigma + 1e-12, (name, p1, p2, sigma)
    assert float(ref["emit_host"].sum()) == 0 and t["recombination"] > 50_000


@pytest.mark.parametrize("charges", [1, 4, 16])
def test_trace_armed_after_earlier_runs(nat, charges):
    """hf_set_trace after hf_run: the records start at the first event fired AFTER arming (all three step kernels),
    and equal the tail of a trace armed from the start."""
    dims, props, seed = (10, 10, 5), (0.84, 0.15, 0.01), 77
    def run(arm_late):
        sim = nat.Sim(dims, props, 0.1, charges, 1, n_trajectories=3, seed=seed)
        try:
            sim.generate_lattice(); sim.inject()
            if not arm_late:
                sim.set_trace(0, 3, 500)
            sim.run(200)
            if arm_late:
                sim.set_trace(0, 3, 300)
            sim.run(300)
            sim.sync()
            return [sim.trace(t, 500) for t in range(3)]
        finally:
            sim.close()
    full, late = run(False), run(True)
    for t in range(3):
        assert len(full[t]) == 500 and len(late[t]) == 300
        for k in ("kind", "particule", "initial", "final", "tau", "draws"):
            assert np.array_equal(late[t][k], full[t][k][200:]), (t, k)


def test_exciton_rates_need_s1_t1(nat):
    """A lattice uploaded without S1 / T1 energies reads zeros through the facade and cannot feed the exciton rates."""
    from hyperfluorescence_b200._native import HfError
    g = load_golden("d0_s1")
    sim = nat.Sim(g["dims"], tuple(float(p) for p in g["props"]), 0.1, 4, 1, n_trajectories=1, seed=1)
    try:
        sim.upload_lattice(0, g["species"], g["homo"], g["lumo"])
        a = sim.download_lattice(0)
        assert not a["s1"].any() and not a["t1"].any() and np.array_equal(a["homo"], g["homo"])
        with pytest.raises(HfError):
            sim.x_configure()
        sim.upload_lattice(0, g["species"], g["homo"], g["lumo"], g["s1"], g["t1"])
        sim.x_configure(nat.x_params_from_dict(dict(cutoff=2.0)))
    finally:
        sim.close()


REFERENCE_STYLE_DRIVER = """
from time import time
from reseau import Lattice
from plot import plot
from event import EVENTS, Point
from molecule import Host, TADF, Fluorescent
import constants

started = time()
device = Lattice((10, 10, 5), (0.84, 0.15, 0.01), charges = 4)
device.operations(10**2)
electrons, holes, excitons = device.get_particules_positions()
assert len(electrons) + len(excitons) <= 4 and all(kind in (Host, TADF, Fluorescent) for _, kind in electrons + holes)
assert callable(plot) and EVENTS["move"] == 1 and isinstance(electrons[0][0], Point)
print(f"IQE : {device.get_IQE()}")
print(f"emissions : {device._emission}")
print(f"injections : {device._injection}")
print(f"{time() - started} s")
"""


def test_dropin_shims_run_a_reference_style_script(nat, tmp_path):
    """The zero-touch path of INTEGRATION.md section 1: a script written against the reference's flat modules
    (`from reseau import Lattice; from plot import plot`, the statements of OLED.py:1-22) runs unmodified in a
    fresh interpreter with only PYTHONPATH=hyperfluorescence_b200/dropin, and prints what the oracle computes for
    the same seed (the shim has no seed argument: $HF_SEED)."""
    import os
    import subprocess
    import sys
    from oracle import frm_oracle as fo
    from tests.conftest import ROOT
    script = tmp_path / "driver.py"
    script.write_text(REFERENCE_STYLE_DRIVER)
    seed = 424242
    env = dict(os.environ, PYTHONPATH=os.path.join(ROOT, "hyperfluorescence_b200", "dropin"), HF_SEED=str(seed))
    res = subprocess.run([sys.executable, str(script)], cwd=str(tmp_path), env=env, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = dict(l.split(" : ") for l in res.stdout.splitlines() if " : " in l)
    olat = fo.OracleLattice.generate((10, 10, 5), (0.84, 0.15, 0.01), seed)
    otr = fo.OracleTrajectory(olat, 4, seed, 0, 0.1)
    otr.run(100)
    t = otr.tallies()
    assert float(lines["IQE"]) == t["IQE"]
    assert (int(lines["emissions"]), int(lines["injections"])) == (t["emission"], t["injection"])


def test_zero_division_is_swallowed_like_the_reference(nat, capsys):
    """dense_s8 hits ZeroDivisionError at step 111 (two electrons on one site): operations() prints
    the cache and returns without refreshing IQE; a later call continues (reseau.py:585-589)."""
    from hyperfluorescence_b200 import Lattice
    g = load_golden("dense_s8")
    lat = Lattice(g["dims"], tuple(float(p) for p in g["props"]), charges=g["charges"], seed=g["seed_int"],
                  lattice=dict(species=g["species"], homo=g["homo"], lumo=g["lumo"]))
    lat.operations(2000)
    out = capsys.readouterr().out
    assert out.startswith("deque([")
    assert lat.get_IQE() == 0.0
    assert lat._step == int(g["final_tallies"][3])
    lat.operations(5)             # continues after the swallowed exception
    assert lat._step > int(g["final_tallies"][3])
    lat.close()


def test_full_size_properties(nat):
    """BASELINE config 2 at full size (64^3 lattice, 10^5 trajectories): size-independent properties.
    Event accounting closes, the run is deterministic, and a sample of trajectories equals the oracle."""
    from oracle import frm_oracle as fo
    dims, props, seed, B, steps = (64, 64, 64), (0.84, 0.15, 0.01), 20261019, 100_000, 300
    def run():
        sim = nat.Sim(dims, props, 0.1, 1, 1, n_trajectories=B, seed=seed)
        sim.generate_lattice()
        sim.inject()
        sim.run(steps)
        t = sim.tallies()
        sample = {i: (sim.info(i).draws, sim.positions(i, 0).tolist(), sim.positions(i, 1).tolist())
                  for i in (0, 1, 77, 31337, B - 1)}
        lat = sim.download_lattice(0)
        sim.close()
        return t, sample, lat
    t1, s1, lat1 = run()
    t2, s2, _ = run()
    assert t1 == t2 and s1 == s2                                   # determinism
    assert t1["steps"] == B * steps and t1["fired"] == B * steps and t1["stopped"] == 0
    assert t1["move_e"] + t1["move_h"] + t1["bound"] + t1["decay"] + t1["capture_e"] + t1["capture_h"] == t1["fired"]
    assert t1["injection"] == 2 * B + 2 * t1["decay"] + t1["capture_e"] + t1["capture_h"]
    olat = fo.OracleLattice.generate(dims, props, seed)
    a = olat.arrays()
    assert np.array_equal(a["species"], lat1["species"])
    np.testing.assert_allclose(lat1["lumo"], a["lumo"], rtol=0, atol=2e-15)
    for i, (draws, e, h) in s1.items():
        o = fo.OracleTrajectory(olat, 1, seed, i, 0.1)
        o.run(steps, trace=False)
        assert o.tallies()["draws"] == draws
        assert o.positions("electrons").tolist() == e and o.positions("holes").tolist() == h


@pytest.mark.parametrize("stats", ["stats_tiny", "stats_d0"])
def test_statistics_within_three_sigma_of_reference(nat, stats):
    """Photons per recombination, recombinations per run and IQE over 20 000 device trajectories
    against the UNMODIFIED reference run with its own Mersenne-Twister streams
    (tests/golden/stats_tiny.npz: 400 runs on 3x3x3; stats_d0.npz: 150 runs of the reference driver's own
    device, OLED.py:6-10, 2000 operations; make_golden.py): within 3 sigma."""
    import os
    from tests.conftest import GOLDEN_DIR
    ref = np.load(os.path.join(GOLDEN_DIR, stats + ".npz"))
    dims = tuple(int(v) for v in ref["dims"])
    props = tuple(float(v) for v in ref["props"])
    charges, steps = int(ref["charges"]), int(ref["steps"])
    B = R = 20_000                       # one trajectory per lattice realisation, like the reference runs
    sim = nat.Sim(dims, props, 0.1, charges, 1, n_trajectories=B, n_realisations=R, seed=4242)
    try:
        sim.generate_lattice()
        sim.inject()
        sim.run(steps)
        sim.sync()
        infos = [sim.info(i) for i in range(B)]
        em = np.array([i.emission for i in infos], float)
        rec = np.array([i.recombination for i in infos], float)
        inj = np.array([i.injection for i in infos], float)
    finally:
        sim.close()
    for name, ours, theirs in (("recombination", rec, ref["recombination"].astype(float)),
                               ("emission", em, ref["emission"].astype(float)),
                               ("IQE", 200.0 * em / inj, 200.0 * ref["emission"] / ref["injection"])):
        sigma = np.sqrt(ours.var(ddof=1) / ours.size + theirs.var(ddof=1) / theirs.size)
        assert abs(ours.mean() - theirs.mean()) < 3 * sigma, (name, ours.mean(), theirs.mean(), sigma)
