"""Run the CUDA kernel SOURCE (hyperfluorescence_b200/csrc/hf_frm.cuh) lane by lane on the host
under tests/emu/warp_emu.h and pin it to the golden vectors of the unmodified reference.

This is a logic check that works without a GPU (the `-m gpu` suite repeats it on the device
through the C-ABI).  Under emulation exp/log/sin/cos come from glibc — the same libm CPython
uses — so here even tau is compared bit for bit."""
import numpy as np
import pytest

from tests.conftest import load_golden
from tests.emu import emu

STATUS = {"": 0, "ZeroDivisionError": 2, "ValueError": 3}

# the emulator is ~1e5x slower than the device: cap the steps per case
DEFAULT_CAP = 3000


def _step_cap(g):
    if g["charges"] >= 10:
        return 1000
    return DEFAULT_CAP


def _check_prefix(g, out, n):
    assert out["init_status"] == 0
    assert np.array_equal(out["init_e_pos"], g["init_e_pos"])
    assert np.array_equal(out["init_h_pos"], g["init_h_pos"])
    for k in ("me", "mh"):
        assert np.array_equal(out[f"init_{k}_init"], g[f"init_{k}_init"])
        assert np.array_equal(out[f"init_{k}_final"], g[f"init_{k}_final"])
        assert np.array_equal(out[f"init_{k}_tau"], g[f"init_{k}_tau"])
    assert out["init_draws"] == int(g["init_draws"][0])
    tr = out["trace"]
    assert len(tr) == n
    assert np.array_equal(tr["kind"], g["trace_kind"][:n])
    assert np.array_equal(tr["particule"], g["trace_particule"][:n])
    assert np.array_equal(tr["initial"], g["trace_initial"][:n])
    assert np.array_equal(tr["final"], g["trace_final"][:n])
    assert np.array_equal(tr["tau"], g["trace_tau"][:n])
    assert np.array_equal(tr["draws"].astype(np.int64), g["trace_draws"][:n])
    assert np.array_equal(tr["spins"].astype(np.int64), g["trace_spins"][:n])


@pytest.mark.parametrize("small_kernel", [False, True])
def test_emulated_kernel_matches_reference(golden, small_kernel):
    """small_kernel: the thread-per-trajectory kernel for 2 <= charges <= 12 (hf_frm_small.cuh): hf_step<1>
    and the screened candidate loop must reproduce the same golden trajectories."""
    g = golden
    if small_kernel and (not 2 <= g["charges"] <= 12 or g["distance"] != 1):
        pytest.skip("the thread-per-trajectory kernel of hf_frm_small.cuh covers 2 <= charges <= 12")
    total = len(g["trace_kind"])
    complete = total <= _step_cap(g) or bool(g["error_s"])
    stop = g["n_steps"] if complete else _step_cap(g)
    out = emu.run(g["dims"], g["field_f"], g["charges"], g["seed_int"], g["trajectory"], g["species"], g["homo"],
                  g["lumo"], stop=stop, replay=g.get("replay_traj"), small_kernel=small_kernel, distance=g["distance"])
    _check_prefix(g, out, total if complete else stop)
    if complete:
        # complete run: final lists, events, tallies and the stop reason must match too
        assert np.array_equal(out["e_pos"], g["final_e_pos"])
        assert np.array_equal(out["h_pos"], g["final_h_pos"])
        for k in ("me", "mh"):
            assert np.array_equal(out[f"{k}_init"], g[f"final_{k}_init"])
            assert np.array_equal(out[f"{k}_final"], g[f"final_{k}_final"])
            assert np.array_equal(out[f"{k}_tau"], g[f"final_{k}_tau"])
        assert [out["emission"], out["recombination"], out["injection"], out["steps"]] == g["final_tallies"].tolist()
        assert [out["draws"], out["spins"]] == g["final_draws"][:2].tolist()
        assert out["status"] == STATUS[g["error_s"]]


def test_emulated_split_runs_equal_one_run():
    """operations(a); operations(b) must equal operations(a+b): state save/restore across launches."""
    g = load_golden("d0_s4")
    args = (g["dims"], g["field_f"], g["charges"], g["seed_int"], 0, g["species"], g["homo"], g["lumo"])
    one = emu.run(*args, stop=600)
    split = emu.run(*args, stop=200, n_calls=3)
    for k in ("kind", "particule", "initial", "final", "tau", "draws"):
        assert np.array_equal(one["trace"][k], split["trace"][k])
    for k in ("e_pos", "h_pos", "me_tau", "mh_tau"):
        assert np.array_equal(one[k], split[k])
    for k in ("e_flags", "h_flags"):
        assert np.array_equal(np.sort(one[k]), np.sort(split[k]))
    assert np.array_equal(one["totals"], split["totals"])
    assert one["cache_n"] == split["cache_n"] == 10


def test_emulated_lattice_generation(golden):
    g = golden
    if g["dims"][0] * g["dims"][1] * g["dims"][2] > 1000:
        pytest.skip("large lattice: covered by the oracle test and by the GPU suite")
    a = emu.generate(g["dims"], tuple(float(p) for p in g["props"]), g["seed_int"], g["realisation"])
    assert np.array_equal(a["species"], g["species"])
    for k in ("homo", "lumo", "s1", "t1"):
        np.testing.assert_allclose(a[k], g[k], rtol=0, atol=4e-16)


def _c1_compare(a, b, B):
    """Everything the two step kernels leave behind must be identical."""
    for k in ("kind", "particule", "initial", "final", "tau", "draws", "spins"):
        assert np.array_equal(a["trace"][k], b["trace"][k]), k
    assert np.array_equal(a["totals"], b["totals"])
    assert np.array_equal(a["cache_i"], b["cache_i"]) and np.array_equal(a["cache_tau"], b["cache_tau"])
    for name in a["sc"].dtype.names:
        assert np.array_equal(a["sc"][name], b["sc"][name]), name
    for t in range(2):
        for i in range(B):
            if a["sc"]["n_pos"][i, t]:
                assert a["pos"][t, i] == b["pos"][t, i]
            if a["sc"]["n_ev"][i, t]:
                assert (a["ev_init"][t, i], a["ev_final"][t, i], a["ev_tau"][t, i]) == \
                    (b["ev_init"][t, i], b["ev_final"][t, i], b["ev_tau"][t, i])
            if a["sc"]["n_flag"][i, t]:
                assert a["flag"][t, i, 0] == b["flag"][t, i, 0]


@pytest.mark.parametrize("name,first", [("c1_s1", 0), ("c1_s2", 0)])
def test_emulated_c1_thread_kernel_matches_reference_and_warp_kernel(name, first):
    """The thread-per-trajectory kernel (charges == 1): trajectory 0 (and 7 for seed 1) against the
    golden vectors, and all 32 lanes against the warp kernel, bit for bit."""
    g = load_golden(name)
    B, stop = 32, 600
    args = (g["dims"], g["field_f"], g["seed_int"], first, B)
    lat = (g["species"], g["homo"], g["lumo"])
    c1 = emu.run_batch_c1(*args, True, *lat, stop=stop // 3, n_calls=3)
    wk = emu.run_batch_c1(*args, False, *lat, stop=stop)
    _c1_compare(c1, wk, B)
    for traj, gold in ((0, g),) + (((7, load_golden("c1_s1_t7")),) if name == "c1_s1" else ()):
        tr = c1["trace"][traj]
        n = min(stop, len(gold["trace_kind"]))
        assert np.array_equal(tr["kind"][:n], gold["trace_kind"][:n])
        assert np.array_equal(tr["initial"][:n], gold["trace_initial"][:n])
        assert np.array_equal(tr["final"][:n], gold["trace_final"][:n])
        assert np.array_equal(tr["tau"][:n], gold["trace_tau"][:n])
        assert np.array_equal(tr["draws"][:n].astype(np.int64), gold["trace_draws"][:n])


@pytest.mark.parametrize("name", ["c1_tiny333_s21", "c1_rag443_s22", "c1_rag554_s23"])
def test_emulated_c1_thread_kernel_recombines_like_the_reference(name):
    """charges == 1 on lattices so small that the carriers meet all the time: the thread kernel's bound / decay /
    capture / re-injection branches (hf_frm_c1.cuh) against the UNMODIFIED reference's own trajectory
    (reseau.py:501-508, 529-561), every event, tau and draw count, plus the final tallies."""
    g = load_golden(name)
    traj, stop = g["trajectory"], len(g["trace_kind"])
    c1 = emu.run_batch_c1(g["dims"], g["field_f"], g["seed_int"], 0, 32, True, g["species"], g["homo"], g["lumo"],
                          stop=stop // 3, n_calls=3)
    tr = c1["trace"][traj]
    for k in ("kind", "particule", "initial", "final", "tau"):
        assert np.array_equal(tr[k][:stop], g["trace_" + k]), k
    assert np.array_equal(tr["draws"][:stop].astype(np.int64), g["trace_draws"])
    assert np.array_equal(tr["spins"][:stop].astype(np.int64), g["trace_spins"])
    sc = c1["sc"]
    assert [int(sc["emission"][traj]), int(sc["recombination"][traj]), int(sc["injection"][traj]), int(sc["steps"][traj])] == \
        g["final_tallies"].tolist()
    assert int(g["final_tallies"][1]) >= 100                             # the case does recombine


def test_emulated_c1_thread_kernel_recombination_paths():
    """A tiny lattice where bound / decay / capture / re-injection fire constantly: the thread kernel
    must follow the warp kernel (itself pinned to the reference) through every rare branch."""
    g = load_golden("tiny333_s5")
    B, stop = 32, 600
    args = (g["dims"], g["field_f"], 4321, 100, B)
    lat = (g["species"], g["homo"], g["lumo"])
    c1 = emu.run_batch_c1(*args, True, *lat, stop=stop // 4, n_calls=4)
    wk = emu.run_batch_c1(*args, False, *lat, stop=stop)
    _c1_compare(c1, wk, B)
    t = c1["totals"]
    assert t[7] > 250 and t[8] > 250 and t[9] + t[10] > 50          # bound, decay, captures all exercised
    assert int(c1["sc"]["recombination"].sum()) == int(t[8])
