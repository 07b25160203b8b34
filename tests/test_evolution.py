"""Composition GA: host logic on the CPU, fitness kernel against the oracle on the GPU."""
import numpy as np
import pytest

from hyperfluorescence_b200.evolution import CompositionGA, Compositions, DeviceFitness, GAConfig, next_generation


def test_compositions_round_trip_and_constraints():
    comps = Compositions((10, 10, 5))
    rng = np.random.default_rng(0)
    pop = comps.random(rng, 500)
    assert pop.shape == (500, 2) and pop.min() >= 1 and (pop.sum(1) <= comps.max_dopants).all()
    props = comps.proportions(pop)
    assert np.allclose(props.sum(1), 1.0)
    # the reference's own reduction of proportions to counts (reseau.py:159-160)
    assert np.array_equal(np.array([int(comps.N * p) for p in props[:, 1]]), pop[:, 0])
    assert np.array_equal(np.array([int(comps.N * p) for p in props[:, 2]]), pop[:, 1])
    assert (props.min(1) * comps.N >= 1).all()                     # reseau.py:137
    wild = comps.clip(np.array([[-5.0, 1e9], [0.2, 0.2], [3.6, 7.4]]))
    assert wild.min() >= 1 and (wild.sum(1) <= comps.max_dopants).all() and wild[2].tolist() == [4, 7]


def test_ga_converges_on_a_synthetic_landscape():
    dims = (10, 10, 5)
    comps = Compositions(dims)
    target = np.array([60.0, 12.0])
    fitness = lambda pop: -np.sum((np.asarray(pop, float) - target) ** 2, axis=1)
    ga = CompositionGA(dims, fitness, GAConfig(population=64, elite=4), seed=3)
    first = ga.step().max()
    out = ga.run(40)
    assert out["best_fitness"] > first and out["best_fitness"] > -9.0
    assert abs(out["best_counts"][0] - 60) <= 3 and abs(out["best_counts"][1] - 12) <= 3
    assert all(h2["best_fitness"] >= h1["best_fitness"] for h1, h2 in zip(ga.history, ga.history[1:]))   # elitism
    assert abs(sum(out["best_proportions"]) - 1.0) < 1e-12


def test_next_generation_is_deterministic_and_valid():
    comps, cfg = Compositions((8, 8, 6)), GAConfig(population=32, elite=2)
    pop = comps.random(np.random.default_rng(1), 32)
    fit = np.arange(32, dtype=float)
    a = next_generation(pop, fit, comps, cfg, np.random.default_rng(9))
    b = next_generation(pop, fit, comps, cfg, np.random.default_rng(9))
    assert np.array_equal(a, b) and a.shape == pop.shape
    assert np.array_equal(a[0], pop[31]) and np.array_equal(a[1], pop[30])       # elites first
    assert a.min() >= 1 and (a.sum(1) <= comps.max_dopants).all()


@pytest.mark.gpu
def test_device_fitness_equals_oracle_per_candidate():
    """Each candidate's emission / injection / recombination sums on the GPU equal the oracle's
    run of the same composition, realisation id and trajectory ids; IQE follows reseau.py:595."""
    from oracle import frm_oracle as fo
    dims, charges, seed = (10, 10, 5), 4, 99
    cfg = GAConfig(population=12, trajectories_per_candidate=40, steps=600)
    comps = Compositions(dims)
    pop = comps.random(np.random.default_rng(2), cfg.population)
    fit_fn = DeviceFitness(dims, charges=charges, cfg=cfg, seed=seed)
    fit = fit_fn(pop)
    t = fit_fn.last_tallies
    props = comps.proportions(pop)
    for r in range(cfg.population):
        lat = fo.OracleLattice.generate(dims, tuple(props[r]), seed, realisation=r, resolved=True)
        o = lat.run_batch(charges, 0.1, seed, r * cfg.trajectories_per_candidate, cfg.trajectories_per_candidate, cfg.steps)
        assert (int(t["emission"][r]), int(t["recombination"][r]), int(t["injection"][r]), int(t["fired"][r])) == \
            (o["emission"], o["recombination"], o["injection"], o["events"])
        assert fit[r] == 100.0 * 2.0 * o["emission"] / o["injection"]
    assert t["recombination"].sum() > 50
    ga = CompositionGA(dims, fit_fn, cfg, seed=5)
    out = ga.run(3)
    assert len(out["history"]) == 3 and out["best_fitness"] >= 0.0
    fit_fn.close()


@pytest.mark.gpu
def test_exciton_fitness_ranks_compositions_and_equals_specification():
    """EQE fitness from the exciton path: per-candidate photons / decayed excitons equal the specification run
    on each candidate's own lattice."""
    from hyperfluorescence_b200 import _native as nat
    from hyperfluorescence_b200.evolution import ExcitonFitness, GAConfig
    from oracle import exciton_oracle as xo
    dims = (16, 16, 8)
    cfg = GAConfig(population=4, trajectories_per_candidate=256, steps=400)
    fit = ExcitonFitness(dims, parameters=dict(cutoff=3.0), cfg=cfg, seed=21)
    counts = fit.comps.clip(np.array([[1, 1], [300, 20], [600, 60], [150, 150]]))
    f = fit(counts)
    assert f.shape == (4,) and np.all(f >= 0) and np.all(f <= 20.0)                 # outcoupling 0.2
    assert len(set(np.round(f, 9))) == 4                                             # four different devices
    sim = fit._handle()
    per = fit.last_tallies
    params = dict(nat.x_params_to_dict(nat.x_default_params()), cutoff=3.0)
    tab = sim.x_tables()
    for r in (0, 2):
        lat = sim.download_lattice(r)
        ws1, wt1 = sim.x_weights(r)
        o = xo.ExcitonOracle(dims, lat["species"], ws1, wt1, params, tab)
        ref = o.run_batch(21, r * 256, 256, 400, n_threads=4)
        for k in xo.XT_NAMES[:15]:
            assert int(per[k][r]) == ref[k], (r, k)
        rad = ref["rad_host"] + ref["rad_tadf"] + ref["rad_fluo"]
        lost = rad + sum(v for k, v in ref.items() if k.startswith("nonrad_"))
        assert abs(f[r] - 100.0 * 0.2 * rad / lost) < 1e-12
    fit.close()


@pytest.mark.gpu
def test_device_eqe_fitness_equals_specification_per_candidate():
    """EQE fitness from the INTEGRATED device (carriers and excitons in one event loop): the per-candidate tallies equal
    the specification (the integrated block of the Path P restatement) run on each candidate's own lattice."""
    from hyperfluorescence_b200 import _native as nat
    from hyperfluorescence_b200.evolution import DeviceEQEFitness, GAConfig
    from oracle import frm_oracle as fo
    from tests.test_integrated_excitons import fast_params
    dims, charges, T, steps = (8, 8, 5), 3, 64, 1500
    params = fast_params(2.0)
    cfg = GAConfig(population=4, trajectories_per_candidate=T, steps=steps)
    fit = DeviceEQEFitness(dims, charges=charges, parameters=params, cfg=cfg, seed=33)
    counts = fit.comps.clip(np.array([[4, 4], [60, 10], [100, 30], [30, 60]]))
    f = fit(counts)
    assert f.shape == (4,) and np.all(f >= 0) and np.all(f <= 20.0) and f.max() > 0
    sim = fit._handle()
    per = fit.last_tallies
    tab = sim.x_tables()
    for r in (1, 3):
        lat = sim.download_lattice(r)
        ws1, wt1 = sim.x_weights(r)
        olat = fo.OracleLattice.from_arrays(dims, lat["species"], lat["homo"], lat["lumo"], lat["s1"], lat["t1"])
        ref = olat.run_batch_x(charges, 0.1, 33, r * T, T, steps, fo.ExcitonModel(2, params, tab, ws1, wt1), n_threads=4)
        assert (int(per["emission"][r]), int(per["recombination"][r]), int(per["injection"][r])) == \
            (ref["emission"], ref["recombination"], ref["injection"]), r
        assert abs(f[r] - 0.2 * 200.0 * ref["emission"] / ref["injection"]) < 1e-12
    fit.close()
