"""Integrated exciton dynamics: the exciton formed by a `bound` event lives on in the SAME First Reaction Method
as the carriers, in the slots the reference reserves and leaves empty (`_move_exciton_events` / `_isc_events` /
`_decay_events` reseau.py:233-235, the `...` branches 526-550, the routing comment 533).

Own model beyond the gate (specification: the "integrated excitons" block of oracle/frm_oracle.c).  What IS pinned to
the reference: with the exciton set to "decay at once" (mode 1) the new slots, lists, selection and dispatch must
reproduce every golden trajectory of the UNMODIFIED reference bit for bit — in the specification, in the kernel
source under emulation, and on the GPU through the C-ABI.  Mode 2 (dynamics) is checked against the specification
event for event, plus the reference's three exciton facts in the limits where nothing else acts."""
import numpy as np
import pytest

from oracle import exciton_oracle as xo
from oracle import frm_oracle as fo
from tests.conftest import golden_names, load_golden
from tests.emu import emu

FIELDS = ("kind", "particule", "initial", "final", "tau")
STATUS = {"": "ok", "ZeroDivisionError": "ZeroDivisionError", "ValueError": "ValueError"}
FAST = dict(k_rad=np.array([[0.0, 0.0], [2e11, 0.0], [4e11, 0.0]]),       # decays on the carriers' own time scale, so that
            k_nonrad=np.array([[4e11, 1e11], [1e11, 2e11], [1e11, 3e11]]),  # a few thousand steps see hundreds of excitons
            k_isc=np.array([1e10, 4e11, 1e10]), k_risc=np.array([0.0, 4e11, 0.0]),
            dexter_prefactor=np.full((3, 3), 3e13))


def fast_params(cutoff):
    return dict(xo.default_params(), cutoff=cutoff, **FAST)


def model_for(a, params):
    tab = xo.tables(params)
    ws1, wt1 = xo.weights(a["s1"], a["t1"], a["species"])
    return fo.ExcitonModel(2, params, tab, ws1, wt1), tab, ws1, wt1


def test_specification_mode1_reproduces_every_golden_trajectory(golden):
    g = golden
    lat = fo.OracleLattice.from_arrays(g["dims"], g["species"], g["homo"], g["lumo"], g["s1"], g["t1"])
    tr = fo.OracleTrajectory(lat, g["charges"], g["seed_int"], g["trajectory"], g["field_f"], g["distance"], replay=g.get("replay_traj"))
    tr.set_excitons(fo.ExcitonModel(1))
    n, t = tr.run(g["n_steps"])
    assert n == len(g["trace_kind"]) and tr.status == STATUS[g["error_s"]]
    for k in FIELDS + ("draws", "spins"):
        assert np.array_equal(t[k], g["trace_" + k]), k
    tl = tr.tallies()
    assert [tl["emission"], tl["recombination"], tl["injection"], tl["step"]] == g["final_tallies"].tolist()
    assert np.array_equal(tr.positions("electrons"), g["final_e_pos"]) and np.array_equal(tr.positions("holes"), g["final_h_pos"])
    assert np.array_equal(tr.positions("excitons"), g["final_x_pos"])
    assert tr.channels()["xdraws"] == 0                                   # the gate consumes nothing on the exciton stream


@pytest.mark.parametrize("name,small,cap", [("d0_s1", False, 700), ("d0_s2", True, 3000), ("c1_tiny333_s21", True, 3000),
                                            ("c1_rag443_s22", False, 900), ("tiny333_s5", True, 1500), ("tiny333_s5", False, 700),
                                            ("dense_s8", False, 900), ("dense_s8", True, 2000), ("rag456_s6", True, 2000),
                                            ("dist2_thin_s32", False, 500), ("replay_s99", True, 1500)])
def test_emulated_kernels_mode1_reproduce_golden_trajectories(name, small, cap):
    """The kernel SOURCE (warp kernel / thread kernel) with the exciton routed through the new slots."""
    g = load_golden(name)
    stop = min(g["n_steps"], cap)
    out = emu.run(g["dims"], g["field_f"], g["charges"], g["seed_int"], g["trajectory"], g["species"], g["homo"], g["lumo"],
                  stop=stop, small_kernel=small, distance=g["distance"], replay=g.get("replay_traj"), excitons=emu.xmodel(1))
    tr = out["trace"]
    n = min(stop, len(g["trace_kind"]))
    assert len(tr) == n
    for k in FIELDS:
        assert np.array_equal(tr[k], g["trace_" + k][:n]), k
    assert np.array_equal(tr["draws"].astype(np.int64), g["trace_draws"][:n])
    assert np.array_equal(tr["spins"].astype(np.int64), g["trace_spins"][:n])
    assert out["xdraws"] == 0
    if n == len(g["trace_kind"]):
        assert [out["emission"], out["recombination"], out["injection"], out["steps"]] == g["final_tallies"].tolist()


CASES = [  # dims, props, charges, cutoff, seed, steps (emulated warp kernel), steps (thread kernel)
    ((6, 6, 4), (0.6, 0.3, 0.1), 3, 2.0, 11, 1500, 6000),
    ((3, 3, 3), (0.5, 0.3, 0.2), 2, 1.0, 12, 1500, 6000),
    ((4, 4, 3), (0.5, 0.3, 0.2), 1, 1.0, 13, 1500, 6000),
    ((9, 9, 4), (0.84, 0.15, 0.01), 6, 3.0, 14, 600, 2500),
]


def _spec_run(dims, props, charges, cutoff, seed, steps, trajectory=0):
    lat = fo.OracleLattice.generate(dims, props, seed)
    a = lat.arrays()
    params = fast_params(cutoff)
    m, tab, ws1, wt1 = model_for(a, params)
    o = fo.OracleTrajectory(lat, charges, seed, trajectory, 0.1)
    o.set_excitons(m)
    n, t = o.run(steps)
    return a, params, tab, ws1, wt1, o, n, t


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("small", [True, False])
def test_emulated_kernels_mode2_equal_the_specification(case, small):
    dims, props, charges, cutoff, seed, steps_warp, steps_thread = case
    steps = steps_thread if small else steps_warp
    a, params, tab, ws1, wt1, o, n, t = _spec_run(dims, props, charges, cutoff, seed, steps)
    out = emu.run(dims, 0.1, charges, seed, 0, a["species"], a["homo"], a["lumo"], stop=steps, small_kernel=small,
                  excitons=emu.xmodel(2, dims, params, tab, ws1, wt1))
    tr = out["trace"]
    assert len(tr) == n and out["status"] == {"ok": 0, "no_events": 1, "ZeroDivisionError": 2, "ValueError": 3}[o.status]
    for k in FIELDS:
        assert np.array_equal(tr[k], t[k]), k
    assert np.array_equal(tr["draws"].astype(np.int64), t["draws"]) and np.array_equal(tr["spins"].astype(np.int64), t["spins"])
    ch, ox, tl = o.channels(), o.excitons(), o.tallies()
    assert out["xdraws"] == ch["xdraws"] and out["clock"] == ch["clock"]
    for k in ("site", "spin", "kind", "final", "tau"):
        assert np.array_equal(out["x_" + k], ox[k]), k
    assert np.array_equal(out["x_rad"], ox["rad"]) and np.array_equal(out["x_onsite"], ox["onflags"])
    tot = out["totals"]
    assert [int(v) for v in tot[17:21]] == [ch["forster"], ch["dexter"], ch["isc"], ch["risc"]]
    assert [int(v) for v in tot[11:16]] == [ch["recomb_host"], ch["recomb_tadf"], ch["recomb_fluo"], ch["emit_tadf"], ch["emit_fluo"]]
    assert [out["emission"], out["recombination"], out["injection"]] == [tl["emission"], tl["recombination"], tl["injection"]]
    assert np.array_equal(out["e_pos"], o.positions("electrons")) and np.array_equal(out["h_pos"], o.positions("holes"))
    flags = o.flags()
    assert np.array_equal(np.sort(out["e_flags"]), np.flatnonzero(flags & 1))
    assert np.array_equal(np.sort(out["h_flags"]), np.flatnonzero(flags & 2))
    if small:                                                              # the long runs do exercise every channel
        assert ch["forster"] + ch["dexter"] > 20 and tl["recombination"] > 5


def test_specification_model_properties():
    """Closed-form properties of the integrated model (own model): the clock is the sum of the taus that advanced
    it, every exciton event carries particule = exciton, photons come only from radiative decays, and carriers
    are conserved (charges = electrons + excitons = holes + excitons at all times)."""
    dims, props, charges, cutoff, seed = (6, 6, 4), (0.6, 0.3, 0.1), 3, 2.0, 21
    a, params, tab, ws1, wt1, o, n, t = _spec_run(dims, props, charges, cutoff, seed, 0)
    for _ in range(400):
        n, t = o.run(25)
        c = o.counts()
        assert c["electrons"] + c["excitons"] == charges == c["holes"] + c["excitons"] or o.status != "ok"
        if o.status != "ok":
            break
    ch, tl = o.channels(), o.tallies()
    assert ch["emit_host"] == 0 and ch["emit_tadf"] + ch["emit_fluo"] == tl["emission"]
    assert ch["recomb_host"] + ch["recomb_tadf"] + ch["recomb_fluo"] == tl["recombination"] > 20
    assert ch["xdraws"] % 2 == 0 and ch["clock"] > 0


# ---- GPU: through the C-ABI ------------------------------------------------------------------------------------

@pytest.fixture(scope="module")
def nat():
    from hyperfluorescence_b200 import _native
    if _native.load().hf_device_count() < 1:
        pytest.fail("no sm_100 device visible: the GPU suite must run on a B200")
    return _native


@pytest.mark.gpu
@pytest.mark.parametrize("warp", [False, True])
def test_gpu_mode1_reproduces_every_golden_trajectory(nat, golden, warp):
    """The gate on the device: mode 1 through hf_p_excitons(1), thread kernels and warp kernel."""
    g = golden
    sim = nat.Sim(g["dims"], tuple(float(p) for p in g["props"]), g["field_f"], g["charges"], g["distance"], n_trajectories=1,
                  seed=g["seed_int"], first_trajectory=g["trajectory"], first_realisation=g["realisation"],
                  flags=nat.HF_FLAG_WARP_KERNEL if warp else 0)
    try:
        sim.upload_lattice(0, g["species"], g["homo"], g["lumo"], g["s1"], g["t1"])
        if "replay_traj" in g:
            sim.set_replay(g["replay_traj"])
        sim.p_excitons(1)
        sim.inject()
        n = g["n_steps"]
        sim.set_trace(0, 1, n)
        sim.run(n // 2)
        if sim.info(0).status == 0:
            sim.run(n - n // 2)
        sim.sync()
        tr = sim.trace(0, n)
        assert len(tr) == len(g["trace_kind"])
        for k in ("kind", "particule", "initial", "final"):
            assert np.array_equal(tr[k], g["trace_" + k]), k
        assert np.array_equal(tr["draws"].astype(np.int64), g["trace_draws"])
        clock = np.cumsum(g["trace_tau"])
        assert np.all(np.abs(tr["tau"] - g["trace_tau"]) <= 1e-12 * (np.abs(g["trace_tau"]) + clock))
        info = sim.info(0)
        assert [info.emission, info.recombination, info.injection, info.steps] == g["final_tallies"].tolist()
        assert np.array_equal(sim.positions(0, 2), g["final_x_pos"])
        assert sim.excitons(0)["draws"] == 0
    finally:
        sim.close()


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("warp", [False, True])
def test_gpu_mode2_equals_the_specification(nat, case, warp):
    """64 trajectories per case: traces of 8, final slots / lists / tallies of all, against the specification run
    on the device's own tables and Boltzmann weights."""
    dims, props, charges, cutoff, seed, _, steps = case
    B = 64
    sim = nat.Sim(dims, props, 0.1, charges, 1, n_trajectories=B, seed=seed, flags=nat.HF_FLAG_WARP_KERNEL if warp else 0)
    try:
        sim.generate_lattice()
        params = fast_params(cutoff)
        sim.x_configure(nat.x_params_from_dict(params))
        sim.p_excitons(2)
        sim.inject()
        sim.set_trace(0, 8, steps)
        sim.run(steps // 3)
        sim.run(steps - steps // 3)
        sim.sync()
        a = sim.download_lattice(0)
        ws1, wt1 = sim.x_weights(0)
        tab = sim.x_tables()
        lat = fo.OracleLattice.from_arrays(dims, a["species"], a["homo"], a["lumo"], a["s1"], a["t1"])
        m = fo.ExcitonModel(2, params, tab, ws1, wt1)
        tot = dict(emission=0, recombination=0, injection=0, x_forster=0, x_dexter=0, x_isc=0, x_risc=0, emit_tadf=0, recomb_tadf=0)
        for b in range(B):
            o = fo.OracleTrajectory(lat, charges, seed, b, 0.1)
            o.set_excitons(m)
            n, t = o.run(steps)
            info, ch, tl, ox = sim.info(b), o.channels(), o.tallies(), o.excitons()
            assert info.fired == n and info.status == {"ok": 0, "no_events": 1, "ZeroDivisionError": 2, "ValueError": 3}[o.status]
            assert [info.emission, info.recombination, info.injection, info.draws, info.spins] == \
                [tl["emission"], tl["recombination"], tl["injection"], tl["draws"], tl["spins"]], b
            x = sim.excitons(b)
            assert x["draws"] == ch["xdraws"]
            for k in ("site", "spin", "kind", "final"):
                assert np.array_equal(x[k], ox[k]), (b, k)
            assert np.array_equal(x["radiative"], ox["rad"])
            np.testing.assert_allclose(x["tau"], ox["tau"], rtol=1e-9, atol=1e-12 * ch["clock"])
            assert abs(x["clock"] - ch["clock"]) <= 1e-11 * ch["clock"]
            assert np.array_equal(sim.positions(b, 0), o.positions("electrons")) and np.array_equal(sim.positions(b, 1), o.positions("holes"))
            assert np.array_equal(sim.positions(b, 2), o.positions("excitons"))
            if b < 8:
                tr = sim.trace(b, steps)
                for k in ("kind", "particule", "initial", "final"):
                    assert np.array_equal(tr[k], t[k]), (b, k)
                assert np.all(np.abs(tr["tau"] - t["tau"]) <= 1e-12 * (np.abs(t["tau"]) + np.cumsum(t["tau"])))
            for k, v in (("emission", tl["emission"]), ("recombination", tl["recombination"]), ("injection", tl["injection"]),
                         ("x_forster", ch["forster"]), ("x_dexter", ch["dexter"]), ("x_isc", ch["isc"]), ("x_risc", ch["risc"]),
                         ("emit_tadf", ch["emit_tadf"]), ("recomb_tadf", ch["recomb_tadf"])):
                tot[k] += v
        got = sim.tallies()
        for k, v in tot.items():
            assert got[k] == v, k
        assert got["x_forster"] + got["x_dexter"] > 100 and got["recombination"] > 50
    finally:
        sim.close()


@pytest.mark.gpu
def test_gpu_facade_with_excitons(nat):
    """Lattice(..., excitons=params).operations(n): the reference's surface with the reserved lists filled."""
    from hyperfluorescence_b200 import Lattice
    from hyperfluorescence_b200.event import EVENTS, PARTICULES
    dims, props, seed = (10, 10, 5), (0.84, 0.15, 0.01), 99
    params = fast_params(2.0)
    lat = Lattice(dims, props, charges=4, seed=seed, excitons=params)
    lat.operations(4000)
    olat = fo.OracleLattice.generate(dims, props, seed)
    a = lat._sim.download_lattice(0)
    ws1, wt1 = lat._sim.x_weights(0)
    o = fo.OracleTrajectory(olat, 4, seed, 0, 0.1)
    o.set_excitons(fo.ExcitonModel(2, params, lat._sim.x_tables(), ws1, wt1))
    o.run(4000)
    tl = o.tallies()
    assert (lat._emission, lat._recombination, lat._injection, lat._step) == (tl["emission"], tl["recombination"], tl["injection"], tl["step"])
    assert lat.get_IQE() == tl["IQE"] and lat.get_EQE() == 0.2 * tl["IQE"]
    ox = o.excitons()
    events = lat._move_exciton_events + lat._isc_events + lat._decay_events
    assert len(events) == len(ox["site"]) == len(lat._excitons_locations)
    assert all(e.particule == PARTICULES["exciton"] for e in events)
    assert sorted(lat._site(e.initial) for e in events) == sorted(ox["site"].tolist())
    assert all(e.kind in (EVENTS["move"], EVENTS["Forster"]) for e in lat._move_exciton_events)
    assert len(lat._get_all_events()) == len(lat._move_electron_events) + len(lat._move_hole_events) + len(events) + \
        len(lat._binding_events) + len(lat._capture_events)
    lat.close()
    # "instant": the reference's behaviour through the new lists equals the default path
    l0 = Lattice(dims, props, charges=4, seed=seed)
    l1 = Lattice(dims, props, charges=4, seed=seed, excitons="instant")
    l0.operations(3000); l1.operations(3000)
    assert (l0._emission, l0._recombination, l0._injection, l0.get_IQE()) == (l1._emission, l1._recombination, l1._injection, l1.get_IQE())
    assert [lat._site(p) for p in l0._electrons_locations] == [lat._site(p) for p in l1._electrons_locations]
    l0.close(); l1.close()


@pytest.mark.gpu
def test_gpu_integrated_degenerate_limit_equals_the_reference_facts(nat):
    """Dynamics switched on but with nothing except decay able to act (transfer / ISC / RISC rates 0): photons per
    recombination on each species must be the UNMODIFIED reference's (tests/golden/stats_d0_1e5.npz) within 3 sigma,
    on OLED.py's own device (OLED.py:6-11).  20 000 trajectories x 2000 operations."""
    from tests.test_exciton_reference_limits import degenerate_params, reference_yields, two_sample
    dims, props, B = (10, 10, 5), (0.84, 0.15, 0.01), 20_000
    sim = nat.Sim(dims, props, 0.1, 4, 1, n_trajectories=B, n_realisations=B, seed=515)
    try:
        sim.generate_lattice()
        p = degenerate_params(cutoff=2.0)
        p["k_rad"] = p["k_rad"] * 1e6; p["k_nonrad"] = p["k_nonrad"] * 1e6        # decays within a carrier hop or two
        sim.x_configure(nat.x_params_from_dict(p))
        sim.p_excitons(2)
        sim.inject()
        sim.run(2000)
        sim.sync()
        t = sim.tallies()
    finally:
        sim.close()
    ref = reference_yields()
    assert t["x_forster"] == t["x_dexter"] == t["x_isc"] == t["x_risc"] == 0
    assert t["recomb_tadf"] > 50_000
    for sp in ("tadf", "fluo"):
        z = two_sample(t["emit_" + sp], t["recomb_" + sp], *ref[sp])
        assert z < 3, (sp, t["emit_" + sp] / t["recomb_" + sp], ref[sp][0] / ref[sp][1], z)
    assert t["emission"] == t["emit_tadf"] + t["emit_fluo"]
