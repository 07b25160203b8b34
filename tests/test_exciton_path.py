"""Exciton path ("Path X").  The reference has no exciton dynamics, so there is no
golden vector from it: parity is UNPINNED and these tests check the CUDA kernel against this
repository's own specification (oracle/exciton_oracle.c) — bit for bit on event sequences and
tallies — plus closed-form properties of the model."""
import numpy as np
import pytest

from oracle import exciton_oracle as xo
from oracle import frm_oracle as fo
from tests.emu import emu


def _oracle(dims, seed, params=None, props=(0.84, 0.15, 0.01)):
    lat = fo.OracleLattice.generate(dims, props, seed).arrays()
    ws1, wt1 = xo.weights(lat["s1"], lat["t1"], lat["species"])
    return xo.ExcitonOracle(dims, lat["species"], ws1, wt1, params or xo.default_params()), lat


def test_tables():
    for rc, m in ((1.0, 6), (1.5, 18), (3.0, 122), (4.0, 256), (5.0, 514)):      # SURVEY.md §8d candidate counts
        p = dict(xo.default_params(), cutoff=rc)
        t = xo.tables(p)
        assert len(t["g6"]) == m
        off = t["offsets"].astype(int)
        d2 = (off ** 2).sum(1)
        assert np.allclose(t["g6"], d2.astype(float) ** -3) and np.allclose(t["gd"], np.exp(-2 * np.sqrt(d2) / 0.5))
        keys = [tuple(o[::-1]) for o in off]                                      # dz-major, dx fastest
        assert keys == sorted(keys)
    t = xo.tables(xo.default_params())
    assert t["kf"][1, 2] == (1e7 + 1e6) * (4.0 ** 3) ** 2


def test_oracle_closed_forms():
    """Sanity of the specification itself on cases with known answers."""
    dims = (12, 12, 6)
    p = xo.default_params()
    # no transfer, no spin flip: every exciton decays on its birth site; photons only from singlets
    # on TADF / fluorescent sites with probability k_rad / (k_rad + k_nonrad)
    p0 = dict(p, forster_radius=np.zeros((3, 3)), dexter_prefactor=np.zeros((3, 3)), k_isc=np.zeros(3), k_risc=np.zeros(3))
    o, lat = _oracle(dims, 3, p0, props=(0.5, 0.3, 0.2))
    b = o.run_batch(11, 0, 400, 200, n_threads=4)
    assert b["forster"] == b["dexter"] == b["isc"] == b["risc"] == 0
    assert b["events"] == 400 * 200 and b["generated"] == 400 + b["events"]       # every event is a decay
    n = b["events"]
    interior = lat["species"].reshape(6, 12, 12)[1:-1].ravel()
    f_tadf, f_fluo = (interior == 1).mean(), (interior == 2).mean()
    expect = 0.25 * (f_tadf * 1e7 / (1e7 + 1e6) + f_fluo * 2e8 / (2e8 + 1e7))
    got = (b["rad_tadf"] + b["rad_fluo"]) / n
    assert abs(got - expect) < 4 * np.sqrt(expect * (1 - expect) / n)
    assert b["rad_host"] == 0
    triplets = b["nonrad_t_host"] + b["nonrad_t_tadf"] + b["nonrad_t_fluo"]
    assert abs(triplets / n - 0.75) < 4 * np.sqrt(0.75 * 0.25 / n)               # molecule.py:93


@pytest.mark.parametrize("dims,cutoff,seed,B,generic", [
    ((12, 12, 6), 4.0, 5, 32, False),      # register-table variant, 8 full slots; every site needs the periodic wrap
    ((24, 22, 7), 4.0, 8, 27, False),      # ... with wrap-free sites too, and a ragged last block of trajectories
    ((9, 10, 5), 3.0, 6, 32, False),       # 3 full slots + a tail slot that mixes transfer and on-site channels
    ((9, 10, 5), 3.0, 6, 5, True),         # the generic shared-memory variant on the same model
    ((8, 8, 8), 1.5, 7, 32, False),        # M = 18 < 32: generic variant
    ((26, 24, 7), 5.0, 9, 19, False),      # cutoff 5: 16 full slots gathered in two halves, wrap-free sites exist
    ((12, 13, 6), 2.5, 10, 32, False),     # M = 80 at run time: 2 register slots + a tail of 16 transfer and 3 on-site channels
    ((14, 15, 6), 3.5, 11, 30, False),     # M = 178: 5 register slots
    ((12, 12, 5), 4.5, 12, 9, False),      # M = 388: 12 slots in two halves of 6
    ((14, 13, 5), 6.0, 13, 6, False),      # M = 924: the first 16 of 28 slots prefetched, 12 through shared memory
    ((12, 12, 6), 4.0, 5, 32, "brick"),    # weights in padded 4 x 4 x 1 bricks: every site sees a periodic image ...
    ((25, 22, 7), 4.0, 8, 27, "brick"),    # ... dimensions that are not multiples of 4 (extra padding)
    ((9, 10, 5), 3.0, 6, 32, "brick"),     # cutoff 3 with the tail slot, odd dimensions
    ((28, 24, 7), 5.0, 9, 19, "brick")])   # cutoff 5: two halves
def test_emulated_exciton_kernel_equals_specification(dims, cutoff, seed, B, generic):
    """hfx_spawn_kernel + hfx_run_kernel run lane by lane on the host (tests/emu): identical event
    sequences, channels, spins, waiting times, final states and tallies for 32 trajectories."""
    p = dict(xo.default_params(), cutoff=cutoff)
    o, _ = _oracle(dims, seed, p, props=(0.6, 0.3, 0.1))
    n = 250
    dev = emu.x_run(o, seed * 1000 + 1, 40, B, n // 2, n_calls=2, generic=generic is True, brick=generic == "brick")
    tot = np.zeros(16, np.int64)
    for b in range(B):
        ref = o.run(seed * 1000 + 1, 40 + b, n)
        tr = dev["trace"][b]
        assert np.array_equal(tr["kind"], ref["trace"]["kind"])
        assert np.array_equal(tr["initial"], ref["trace"]["initial"]) and np.array_equal(tr["final"], ref["trace"]["final"])
        assert np.array_equal(tr["spins"].astype(np.int32), ref["trace"]["channel"])
        assert np.array_equal(tr["pad"][:, 0], ref["trace"]["spin"])
        assert np.array_equal(tr["tau"], ref["trace"]["dt"])                     # same libm under emulation
        assert (dev["site"][b], dev["spin"][b], int(dev["draws"][b])) == (ref["site"], ref["spin"], ref["draws"])
        assert dev["lifetime_sum"][b] == ref["lifetime_sum"]
        tot += np.array([ref["tallies"][k] for k in xo.XT_NAMES[:15]] + [0])
    assert np.array_equal(dev["totals"].astype(np.int64), tot)
    assert tot[2] > 0 and tot[3] > 0 and tot[6:15].sum() > 0                           # Forster, Dexter and decays occurred
    if cutoff <= 4.0 and B >= 27:
        assert tot[4] + tot[5] > 0                                                     # ... and spin flips (rare next to >= 388 transfer channels)


@pytest.mark.parametrize("bX,bY,r", [(16, 12, 3), (24, 20, 4), (136, 136, 4), (40, 36, 5)])
def test_brick_offsets_are_exact(bX, bY, r):
    """hfx_brick_lin packs, per neighbour offset, the part of the brick-address difference that does not depend on the
    site and two carry flags per residue; the gather adds 12 / 4 X' - 16 when they are set.  Exhaustively equal to the
    difference of the brick addresses for every offset within the cutoff and every site of the padded plane."""
    assert emu.brick_selftest(bX, bY, r) == 0


@pytest.mark.gpu
@pytest.mark.parametrize("dims,cutoff", [((24, 24, 12), 4.0), ((21, 27, 9), 3.0), ((32, 26, 11), 5.0)])
def test_gpu_brick_layout_changes_nothing(dims, cutoff, monkeypatch):
    """The 4 x 4 x 1 brick layout of the weights (used when they exceed the L2; forced here with HF_X_BRICK=1) is a
    layout only: traces, final states, tallies and the downloaded (row-major) weights equal the row-major run's."""
    from hyperfluorescence_b200 import _native as nat
    out = {}
    for brick in (0, 1):
        monkeypatch.setenv("HF_X_BRICK", str(brick))
        sim = nat.Sim(dims, (0.84, 0.15, 0.01), 0.1, 1, 1, n_trajectories=3000, n_realisations=3, seed=91, first_trajectory=7)
        try:
            sim.generate_lattice()
            p = nat.x_default_params()
            p.cutoff = cutoff
            sim.x_configure(p)
            assert sim.x_launch_info()["brick"] is bool(brick)
            sim.set_trace(0, 8, 200)
            sim.x_spawn()
            sim.x_run(120); sim.x_run(80)
            sim.sync()
            st = sim.x_state()
            out[brick] = (st, sim.x_realisation_tallies(), [sim.x_trace(b, 200) for b in range(8)], sim.x_weights(2))
        finally:
            sim.close()
    a, b = out[0], out[1]
    for k in a[0]:
        assert np.array_equal(a[0][k], b[0][k]), k
    for k in a[1]:
        assert np.array_equal(a[1][k], b[1][k]), k
    for ta, tb in zip(a[2], b[2]):
        assert np.array_equal(ta, tb)
    assert np.array_equal(a[3][0], b[3][0]) and np.array_equal(a[3][1], b[3][1])
    # cutoffs without a compiled brick variant keep the row-major layout
    monkeypatch.setenv("HF_X_BRICK", "1")
    sim = nat.Sim((22, 24, 8), (0.84, 0.15, 0.01), 0.1, 1, 1, n_trajectories=64, seed=1)
    try:
        sim.generate_lattice()
        p = nat.x_default_params()
        p.cutoff = 2.5
        sim.x_configure(p)
        assert sim.x_launch_info()["brick"] is False
    finally:
        sim.close()


@pytest.mark.gpu
def test_gpu_exciton_path_equals_specification():
    """Through the C-ABI on the device: 2048 trajectories on 2 realisations; traces of 16 of them,
    all final states and the per-realisation tallies equal the oracle run on the device's own
    Boltzmann weights and tables (identical inputs => identical selection, bit for bit)."""
    from hyperfluorescence_b200 import _native as nat
    dims, props, seed, B, R, n = (24, 24, 12), (0.84, 0.15, 0.01), 77, 2048, 2, 300
    sim = nat.Sim(dims, props, 0.1, 1, 1, n_trajectories=B, n_realisations=R, seed=seed, first_trajectory=1000)
    try:
        sim.generate_lattice()
        sim.x_configure()
        tab = sim.x_tables()
        ref_tab = xo.tables(xo.default_params())
        assert np.array_equal(tab["offsets"], ref_tab["offsets"]) and np.array_equal(tab["g6"], ref_tab["g6"])
        np.testing.assert_allclose(tab["gd"], ref_tab["gd"], rtol=1e-15)
        assert np.array_equal(tab["kf"], ref_tab["kf"]) and np.array_equal(tab["kd"], ref_tab["kd"])
        sim.set_trace(0, 16, n)
        sim.x_spawn()
        sim.x_run(n // 3)
        sim.x_run(n - n // 3)
        sim.sync()
        st = sim.x_state()
        per_real = sim.x_realisation_tallies()
        params = nat.x_params_to_dict(nat.x_default_params())
        for r in range(R):
            lat = sim.download_lattice(r)
            ws1, wt1 = sim.x_weights(r)
            np.testing.assert_allclose(ws1, np.exp(-lat["s1"] / xo.KT), rtol=1e-13)
            assert np.array_equal(ws1.view(np.uint64) & np.uint64(3), lat["species"].astype(np.uint64))     # the species tag
            o = xo.ExcitonOracle(dims, lat["species"], ws1, wt1, params, tab)
            lo, hi = r * (B // R), (r + 1) * (B // R)
            sample = list(range(lo, lo + 16 if r == 0 else lo + 4)) + list(range(hi - 40, hi))
            for b in sample:
                ref = o.run(seed, 1000 + b, n)
                assert (st["site"][b], st["spin"][b], int(st["draws"][b])) == (ref["site"], ref["spin"], ref["draws"])
                np.testing.assert_allclose(st["lifetime_sum"][b], ref["lifetime_sum"], rtol=1e-12)
                if b < 16:
                    tr = sim.x_trace(b, n)
                    assert np.array_equal(tr["kind"], ref["trace"]["kind"]) and np.array_equal(tr["final"], ref["trace"]["final"])
                    assert np.array_equal(tr["spins"].astype(np.int32), ref["trace"]["channel"])
                    np.testing.assert_allclose(tr["tau"], ref["trace"]["dt"], rtol=1e-12)
            whole = o.run_batch(seed, 1000 + lo, B // R, n, n_threads=8)
            for k in xo.XT_NAMES[:15]:
                assert int(per_real[k][r]) == whole[k], (r, k)
        tot = sim.x_tallies()
        assert tot["events"] == B * n and tot["generated"] == B + sum(tot[k] for k in tot if k.startswith(("rad_", "nonrad_")))
    finally:
        sim.close()


@pytest.mark.gpu
def test_gpu_exciton_results_do_not_depend_on_the_warp_batch(monkeypatch):
    """The number of trajectories a warp owns at a time (32, or fewer when the weights exceed the L2:
    hf_x_configure) is a scheduling choice: states and tallies must not change with it."""
    from hyperfluorescence_b200 import _native as nat
    out = []
    for block in ("32", "3", "7"):
        monkeypatch.setenv("HF_X_BLOCK", block)
        sim = nat.Sim((20, 20, 10), (0.84, 0.15, 0.01), 0.1, 1, 1, n_trajectories=1000, n_realisations=2, seed=3)
        sim.generate_lattice()
        sim.x_configure()
        sim.x_spawn()
        sim.x_run(150)
        sim.sync()
        st = sim.x_state()
        out.append((st["site"].copy(), st["spin"].copy(), st["draws"].copy(), st["lifetime_sum"].copy(), sim.x_realisation_tallies()))
        sim.close()
    for o in out[1:]:
        for a, b in zip(out[0][:4], o[:4]):
            assert np.array_equal(a, b)
        for k in out[0][4]:
            assert np.array_equal(out[0][4][k], o[4][k]), k


@pytest.mark.gpu
def test_gpu_exciton_ensemble_api():
    from hyperfluorescence_b200.excitons import ExcitonEnsemble
    ens = ExcitonEnsemble((20, 20, 20), trajectories=4096, seed=5)
    ens.run(2000)
    s = ens.summary()
    assert s["events"] == 4096 * 2000 and s["excitons_decayed"] > 10_000
    assert 0.3 < s["iqe"] < 0.8 and abs(s["eqe"] - 0.2 * s["iqe"]) < 1e-15
    assert abs(sum(s["channels"].values()) - 1.0) < 1e-12
    assert s["channels"]["rad_host"] == 0.0                          # hosts never emit (molecule.py:396-400)
    assert s["channels"]["rad_fluo"] > 0.02                          # hyperfluorescence: TADF -> emitter transfer
    with pytest.raises(ValueError):
        ExcitonEnsemble((20, 20, 2))
    ens.close()


@pytest.mark.gpu
def test_gpu_exciton_full_size_properties():
    """The north_star's own size (128^3, 10^6 trajectories, cutoff 4), where the oracle cannot follow:
    conservation laws of the tallies, the draw accounting of every trajectory and a sampled comparison
    of 24 trajectories with the specification."""
    from hyperfluorescence_b200 import _native as nat
    B, n = 1_000_000, 60
    sim = nat.Sim((128, 128, 128), (0.84, 0.15, 0.01), 0.1, 1, 1, n_trajectories=B, seed=11)
    try:
        sim.generate_lattice()
        sim.x_configure()
        sim.x_spawn()
        sim.x_run(n // 2)
        sim.x_run(n - n // 2)
        sim.sync()
        t = sim.x_tallies()
        decays = sum(v for k, v in t.items() if k.startswith(("rad_", "nonrad_")))
        assert t["events"] == B * n
        assert t["generated"] == B + decays
        assert t["forster"] + t["dexter"] + t["isc"] + t["risc"] + decays == t["events"]
        assert t["rad_host"] == 0                                                  # hosts never emit (molecule.py:396-400)
        st = sim.x_state()
        assert int(st["draws"].sum()) == 2 * (B + t["events"] + decays)            # spawn + event + respawn blocks
        assert st["site"].min() >= 0 and st["site"].max() < 128 ** 3
        assert set(np.unique(st["spin"])) <= {1, 3}
        assert np.all(st["time"] >= 0) and np.all(np.isfinite(st["time"]))
        lat = sim.download_lattice(0)
        ws1, wt1 = sim.x_weights(0)
        o = xo.ExcitonOracle((128, 128, 128), lat["species"], ws1, wt1, nat.x_params_to_dict(nat.x_default_params()), sim.x_tables())
        for b in list(range(0, 8)) + list(range(499_990, 499_998)) + list(range(B - 8, B)):
            ref = o.run(11, b, n)
            assert (st["site"][b], st["spin"][b], int(st["draws"][b])) == (ref["site"], ref["spin"], ref["draws"]), b
    finally:
        sim.close()


@pytest.mark.gpu
def test_gpu_exciton_api_errors():
    """Misuse of the hf_x_* / hf_xd_* entry points fails with a status and a message, never with a crash."""
    from hyperfluorescence_b200 import _native as nat
    sim = nat.Sim((12, 12, 6), (0.84, 0.15, 0.01), 0.1, 1, 1, n_trajectories=64, seed=1)
    try:
        with pytest.raises(nat.HfError, match="before a lattice"):
            sim.x_configure()
        sim.generate_lattice()
        with pytest.raises(nat.HfError, match="before hf_x_configure"):
            sim.x_spawn()
        with pytest.raises(nat.HfError, match="before hf_x_configure"):
            sim.xd_configure(8)
        for bad in (dict(cutoff=0.5), dict(cutoff=7.0), dict(dexter_length=0.0), dict(singlet_fraction=1.5),
                    dict(k_nonrad=np.zeros((3, 2)), k_rad=np.zeros((3, 2))), dict(k_isc=np.array([-1.0, 0, 0]))):
            with pytest.raises(nat.HfError):
                sim.x_configure(nat.x_params_from_dict(bad))
        with pytest.raises(nat.HfError, match="unique periodic images"):
            sim.x_configure(nat.x_params_from_dict(dict(cutoff=6.0)))              # 12 < 2 * 6 + 1
        sim.x_configure()
        with pytest.raises(nat.HfError, match="before hf_x_spawn"):
            sim.x_run(10)
        with pytest.raises(nat.HfError, match="before hf_xd_spawn"):
            sim.xd_configure(8); sim.xd_run(10)
        for bad_n in (0, 2000, 400):                                               # 400 > half of the 576 interior sites
            with pytest.raises(nat.HfError):
                sim.xd_configure(bad_n)
        with pytest.raises(nat.HfError):
            sim.xd_configure(8, ((1.0, -1.0), (1.0, 1.0)))
        with pytest.raises(nat.HfError):
            sim.xd_configure(8, p_tta=1.5)
        sim.x_spawn(); sim.x_run(0); sim.x_run(5); sim.sync()
        assert sim.x_tallies()["events"] == 64 * 5
        with pytest.raises(nat.HfError):
            sim.x_run(-1)
        sim.x_configure(nat.x_params_from_dict(dict(cutoff=2.0)))                  # reconfiguring drops the dense devices
        with pytest.raises(nat.HfError, match="before hf_xd_configure"):
            sim.xd_spawn()
    finally:
        sim.close()


@pytest.mark.gpu
@pytest.mark.parametrize("dims,cutoff,B,R", [((13, 13, 3), 1.0, 1, 1),       # one interior plane, 6 channels, one trajectory
                                               ((13, 14, 4), 6.0, 33, 1),      # the largest cutoff (M = 924): generic kernel, ragged batch
                                               ((16, 16, 9), 2.5, 70, 7),      # realisation boundaries inside a warp's batch
                                               ((11, 12, 6), 5.0, 40, 2)])     # cutoff 5 (17 slots of channels)
def test_gpu_exciton_edge_shapes(dims, cutoff, B, R):
    """Extreme shapes of the exciton path against the specification: final state of every trajectory and the
    per-realisation tallies."""
    from hyperfluorescence_b200 import _native as nat
    n = 120
    sim = nat.Sim(dims, (0.7, 0.2, 0.1), 0.1, 1, 1, n_trajectories=B, n_realisations=R, seed=31, first_trajectory=3)
    try:
        sim.generate_lattice()
        params = dict(nat.x_params_to_dict(nat.x_default_params()), cutoff=cutoff)
        sim.x_configure(nat.x_params_from_dict(dict(cutoff=cutoff)))
        sim.x_spawn()
        sim.x_run(n)
        sim.sync()
        st = sim.x_state()
        per_real = sim.x_realisation_tallies()
        tab = sim.x_tables()
        for r in range(R):
            lat = sim.download_lattice(r)
            ws1, wt1 = sim.x_weights(r)
            o = xo.ExcitonOracle(dims, lat["species"], ws1, wt1, params, tab)
            lo, hi = r * (B // R), (r + 1) * (B // R)
            acc = {}
            for b in range(lo, hi):
                ref = o.run(31, 3 + b, n)
                assert (st["site"][b], st["spin"][b], int(st["draws"][b])) == (ref["site"], ref["spin"], ref["draws"]), b
                for k, v in ref["tallies"].items():
                    acc[k] = acc.get(k, 0) + v
            for k, v in acc.items():
                assert int(per_real[k][r]) == v, (r, k)
    finally:
        sim.close()


@pytest.mark.gpu
def test_gpu_exciton_batch_tuning_leaves_no_trace(monkeypatch):
    """Weights larger than half the L2 (128^3 x 4 realisations): the first hf_x_run measures a few candidate
    warp batches on a copy of the state.  States and tallies must equal those of an untuned run."""
    from hyperfluorescence_b200 import _native as nat
    out = []
    for block in (None, "32"):
        if block is None:
            monkeypatch.delenv("HF_X_BLOCK", raising=False)
        else:
            monkeypatch.setenv("HF_X_BLOCK", block)
        sim = nat.Sim((128, 128, 128), (0.84, 0.15, 0.01), 0.1, 1, 1, n_trajectories=20_000, n_realisations=4, seed=17)
        sim.generate_lattice()
        sim.x_configure()
        sim.x_spawn()
        sim.x_run(40)
        sim.x_run(25)
        sim.sync()
        st = sim.x_state()
        out.append((st["site"].copy(), st["spin"].copy(), st["draws"].copy(), st["time"].copy(), st["lifetime_sum"].copy(),
                    sim.x_realisation_tallies()))
        sim.close()
    for a, b in zip(out[0][:5], out[1][:5]):
        assert np.array_equal(a, b)
    for k in out[0][5]:
        assert np.array_equal(out[0][5][k], out[1][5][k]), k
    assert int(out[0][5]["events"].sum()) == 20_000 * 65


@pytest.mark.gpu
def test_gpu_exciton_ensemble_on_uploaded_lattices():
    """ExcitonEnsemble on lattices supplied by the caller (one dict per realisation) equals the ensemble that
    generated the same lattices on the device."""
    from hyperfluorescence_b200.excitons import ExcitonEnsemble
    dims = (18, 16, 9)
    gen = ExcitonEnsemble(dims, trajectories=600, realisations=3, seed=44, parameters=dict(cutoff=3.0))
    lats = [gen._sim.download_lattice(r) for r in range(3)]
    up = ExcitonEnsemble(dims, trajectories=600, realisations=3, seed=44, parameters=dict(cutoff=3.0), lattice=lats)
    gen.run(180); up.run(180)
    assert gen.tallies() == up.tallies()
    a, b = gen._sim.x_state(), up._sim.x_state()
    for k in ("site", "spin", "draws", "time", "lifetime_sum"):
        assert np.array_equal(a[k], b[k]), k
    s = up.summary()
    assert s["events"] == 600 * 180 and abs(sum(s["channels"].values()) - 1.0) < 1e-12
    gen.close(); up.close()
