"""Pin oracle/frm_oracle.c to the golden vectors frozen from the UNMODIFIED reference
(tests/golden/make_golden.py).  Everything here is bit-exact: the oracle runs on the same glibc
libm as CPython, so even tau must match to the last bit."""
import numpy as np
import pytest

from oracle import frm_oracle as fo
from oracle import philox as px


def test_philox_known_answers():
    # Random123 kat_vectors for philox4x32-10
    assert px.philox4x32_10((0, 0, 0, 0), (0, 0)) == (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)
    assert px.philox4x32_10((0xffffffff,) * 4, (0xffffffff,) * 2) == (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)
    assert px.philox4x32_10((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)) == \
        (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)


def test_c_uniform_matches_python_contract():
    for seed, stream, ident, d in [(0, 0, 0, 0), (1, 0, 0, 5), (0x48464B4D43, 3, 77, 2 ** 33 + 5), (2 ** 63 + 9, 1, 2 ** 31, 123456)]:
        assert fo.contract_uniform(seed, stream, ident, d) == px.uniform(seed, stream, ident, d)


def test_lattice_generation_matches_reference(golden):
    g = golden
    lat = fo.OracleLattice.generate(g["dims"], tuple(g["props"]), g["seed_int"], g["realisation"])
    a = lat.arrays()
    assert np.array_equal(a["species"], g["species"])
    for k in ("homo", "lumo", "s1", "t1"):
        assert np.array_equal(a[k], g[k]), k            # bit-exact (same libm)
    assert lat.species_draws == int(g["init_draws"][2])


@pytest.mark.parametrize("generated", [False, True])
def test_trajectory_matches_reference(golden, generated):
    g = golden
    if generated:
        lat = fo.OracleLattice.generate(g["dims"], tuple(g["props"]), g["seed_int"], g["realisation"])
    else:
        lat = fo.OracleLattice.from_arrays(g["dims"], g["species"], g["homo"], g["lumo"], g["s1"], g["t1"])
    replay = g.get("replay_traj")
    tr = fo.OracleTrajectory(lat, g["charges"], g["seed_int"], g["trajectory"], g["field_f"], g["distance"], replay=replay)
    assert tr.status == "ok"
    # state right after Lattice.__init__
    assert np.array_equal(tr.positions("electrons"), g["init_e_pos"])
    assert np.array_equal(tr.positions("holes"), g["init_h_pos"])
    for name in ("me", "mh"):
        ev = tr.events({"me": "move_e", "mh": "move_h"}[name])
        assert np.array_equal(ev["initial"], g[f"init_{name}_init"])
        assert np.array_equal(ev["final"], g[f"init_{name}_final"])
        assert np.array_equal(ev["tau"], g[f"init_{name}_tau"])
    t = tr.tallies()
    assert t["draws"] == int(g["init_draws"][0])
    # the step loop
    n, trace = tr.run(g["n_steps"])
    assert n == len(g["trace_kind"])
    assert np.array_equal(trace["kind"], g["trace_kind"])
    assert np.array_equal(trace["particule"], g["trace_particule"])
    assert np.array_equal(trace["initial"], g["trace_initial"])
    assert np.array_equal(trace["final"], g["trace_final"])
    assert np.array_equal(trace["tau"], g["trace_tau"])
    assert np.array_equal(trace["draws"], g["trace_draws"])
    assert np.array_equal(trace["spins"], g["trace_spins"])
    assert tr.status == ({"": "ok"}.get(g["error_s"], g["error_s"]) if n == g["n_steps"] or g["error_s"] else "no_events")
    # final state
    assert np.array_equal(tr.positions("electrons"), g["final_e_pos"])
    assert np.array_equal(tr.positions("holes"), g["final_h_pos"])
    assert np.array_equal(tr.positions("excitons"), g["final_x_pos"])
    for name in ("me", "mh"):
        ev = tr.events({"me": "move_e", "mh": "move_h"}[name])
        assert np.array_equal(ev["initial"], g[f"final_{name}_init"])
        assert np.array_equal(ev["final"], g[f"final_{name}_final"])
        assert np.array_equal(ev["tau"], g[f"final_{name}_tau"])
    t = tr.tallies()
    assert [t["emission"], t["recombination"], t["injection"], t["step"]] == g["final_tallies"].tolist()
    assert [t["draws"], t["spins"]] == g["final_draws"][:2].tolist()
