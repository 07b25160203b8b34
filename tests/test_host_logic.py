"""CPU-side checks (`-m "not gpu"`): the C-ABI library loads and exports every symbol
include/hfkmc.h declares, the product fails loudly without a GPU (no CPU fallback), the Python
data model mirrors the reference's event.py / molecule.py / constants.py, and the multi-rank
host logic (sharding + tally all-reduce) works over gloo with world_size 2."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "hfkmc.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hf_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from hyperfluorescence_b200 import _native, build
    build.build()
    lib = _native.load()
    declared = _declared_symbols()
    assert len(declared) >= 20
    assert sorted(_native.SIGNATURES) == declared          # the binding covers the whole header
    for name in declared:
        assert hasattr(lib, name), name
    out = subprocess.run(["nm", "-D", "--defined-only", _native.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\bT (hf_[a-z_0-9]+)\b", out))
    assert set(declared) <= exported
    m = re.search(r"#define HF_ABI_VERSION (\d+)", open(os.path.join(ROOT, "include", "hfkmc.h")).read())
    assert lib.hf_abi_version() == int(m.group(1)) == 2


def test_struct_layouts_match_the_header():
    from hyperfluorescence_b200 import _native
    assert ctypes.sizeof(_native.HfEvent) == 32
    assert ctypes.sizeof(_native.HfTrajInfo) == 12 * 4 + 7 * 8
    assert ctypes.sizeof(_native.HfConfig) == 3 * 4 + 4 + 3 * 8 + 8 + 4 + 4 + 8 + 4 + 4 + 4 + 4 + 8 + 4 + 4


@pytest.mark.skipif(os.path.exists("/dev/nvidiactl"), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback():
    """Without a device the product path raises; it never routes through the oracle."""
    from hyperfluorescence_b200 import Lattice, _native
    assert _native.load().hf_device_count() == 0
    with pytest.raises(_native.HfError) as exc:
        Lattice((10, 10, 5), (0.84, 0.15, 0.01), charges=4)
    assert exc.value.code == _native.HF_ERR_NO_DEVICE
    # argument validation happens before any device work and mirrors reseau.py:126-139, 212-213
    for args, err in ((([10, 10, 5], (0.84, 0.15, 0.01)), TypeError), (((10, 10), (0.84, 0.15, 0.01)), IndexError),
                      (((10, 10, 2), (0.84, 0.15, 0.01)), ValueError), (((3, 3, 3), (0.84, 0.15, 0.01)), ValueError),
                      (((10, 10, 5), (0.84, 0.15)), IndexError)):
        with pytest.raises(err):
            Lattice(*args)
    with pytest.raises(ValueError):
        Lattice((4, 4, 8), (0.5, 0.3, 0.2), charges=17)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "hyperfluorescence_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "frm_oracle" not in text and "libfrm_oracle" not in text, f


def test_constants_and_codes():
    from hyperfluorescence_b200 import constants as c
    from hyperfluorescence_b200.event import EVENTS, PARTICULES
    from hyperfluorescence_b200.molecule import EXCITON
    assert c.ELECTROSTATIC == 4.799881784501343e-07            # SURVEY.md §2 row 4 [measured]
    assert EVENTS == {"move": 1, "bound": 2, "ISC": 3, "Forster": 4, "decay": 5, "unbound": 6, "capture": 7}
    assert PARTICULES == {"electron": 1, "hole": 2, "exciton": 3}
    assert EXCITON == {"none": 0, "singlet": 1, "doublet": 2, "triplet": 3}


def test_event_semantics():
    from hyperfluorescence_b200.event import EVENTS, PARTICULES, Event, Point, Vector
    a, b = Point(1, 2, 3), Point(4, 6, 3)
    d = b - a
    assert isinstance(d, Vector) and (d.x, d.y, d.z) == (3, 4, 0) and d.norm() == 5.0
    assert (a + 1) == Vector(2, 3, 4) and d * d == 25 and (d * 2.).x == 6. and (-1. * Vector(0, 0, 0.1)).z == -0.1
    with pytest.raises(TypeError):
        a + "x"
    m = EVENTS["move"]
    e1 = Event(a, b, 2.0, m, PARTICULES["electron"])
    e2 = Event(Point(9, 9, 9), b, 1.0, m, PARTICULES["electron"])     # same final -> "equal" (collision rule)
    e3 = Event(Point(9, 9, 9), b, 1.0, m, PARTICULES["hole"])
    assert e1 == e2 and e1 != e3 and e2 < e1 and e1 >= e2 and min([e1, e2]) is e2
    c1 = Event(a, a, 0., EVENTS["capture"], 1)
    assert c1 == Event(a, a, 5., EVENTS["capture"], 1) and c1 != Event(b, b, 0., EVENTS["capture"], 1)
    with pytest.raises(TypeError):
        e1 == 3


def test_molecule_semantics():
    from hyperfluorescence_b200.event import Point
    from hyperfluorescence_b200.molecule import EXCITON, Fluorescent, Host, Proportion, TADF, Molecule
    p = Point(0, 0, 0)
    for cls, means, emits in ((Host, (6.0, -2.0, 3.5, 3.0), False), (TADF, (5.8, -2.6, 2.55, 2.52), True),
                              (Fluorescent, (5.25, -1.84, 2.69, 1.43), True)):
        mols = [cls(p, []) for _ in range(400)]
        got = np.array([[m.homo_energy, m.lumo_energy, m.s1_energy, m.t1_energy] for m in mols])
        assert np.all(np.abs(got.mean(0) - np.array(means)) < 0.03) and np.all(np.abs(got.std(0) - 0.1) < 0.03)
        m = mols[0]
        assert m.empty() and m.exciton_decay() is None
        m.switch_electron(); m.generate_exciton()
        assert m.exciton == 0                                   # needs both carriers
        m.switch_hole(); m.generate_exciton()
        assert m.exciton in (EXCITON["singlet"], EXCITON["triplet"])
        state = m.exciton
        assert m.exciton_decay() == (emits and state == EXCITON["singlet"])
        assert not m.electron and not m.hole and m.exciton == 0
    t = TADF(p, [])
    t.exciton = EXCITON["singlet"]; t.intersystem_crossing()
    assert t.exciton == EXCITON["triplet"]
    t.intersystem_crossing()
    assert t.exciton == EXCITON["singlet"]
    singlets = 0
    for _ in range(4000):
        m = Host(p, []); m.electron = m.hole = True; m.generate_exciton(); singlets += m.exciton == 1
    assert abs(singlets / 4000 - 0.25) < 0.03                    # molecule.py:93
    with pytest.raises(NotImplementedError):
        Molecule(p, []).exciton_decay()
    assert Proportion(0.84, 0.15, 0.01).tadf == 0.15


def test_data_model_equals_reference_when_mounted():
    """Where /root/reference is mounted, compare the façade's data model with the real modules."""
    from oracle import ref_harness as rh
    if not rh.reference_available():
        pytest.skip("reference tree not mounted")
    _, rmol, rev, rcst = rh.import_reference()
    from hyperfluorescence_b200 import constants, event, molecule
    assert (constants.BOLTZMANN, constants.ELECTROSTATIC) == (rcst.BOLTZMANN, rcst.ELECTROSTATIC)
    assert event.EVENTS == rev.EVENTS and event.PARTICULES == rev.PARTICULES and molecule.EXCITON == rmol.EXCITON
    for name in ("Host", "TADF", "Fluorescent"):
        import inspect
        ours = inspect.signature(getattr(molecule, name).__init__)
        theirs = inspect.signature(getattr(rmol, name).__init__)
        assert [(p.name, p.default) for p in ours.parameters.values()] == [(p.name, p.default) for p in theirs.parameters.values()]
    pts = [(0, 0, 0), (9, 4, 2), (3, 9, 4)]
    for xyz in pts:
        for other in pts:
            a, b = event.Point(*xyz) - event.Point(*other), rev.Point(*xyz) - rev.Point(*other)
            assert (a.x, a.y, a.z, a.norm()) == (b.x, b.y, b.z, b.norm())


def test_neighbour_table_matches_oracle_and_reference_order():
    from hyperfluorescence_b200.reseau import born_von_karman
    from oracle import frm_oracle as fo
    assert born_von_karman(0, 1, 10, "x") == [0, 1, 9] and born_von_karman(9, 1, 10, "y") == [8, 9, 0]
    assert born_von_karman(0, 1, 5, "z") == [0, 1] and born_von_karman(4, 1, 5, "z") == [3, 4]
    assert born_von_karman(1, 2, 10, "x") == [0, 1, 2, 8, 9] and born_von_karman(8, 2, 10, "x") == [6, 7, 8, 9, 0, 1]   # Q3
    dims = (10, 10, 5)
    lat = fo.OracleLattice(dims)
    site = lambda x, y, z: (z * dims[1] + y) * dims[0] + x
    nb = lat.neighbourhood(site(0, 0, 0))
    assert len(nb) == 17
    assert nb[:6].tolist() == [site(0, 0, 1), site(0, 1, 0), site(0, 1, 1), site(0, 9, 0), site(0, 9, 1), site(1, 0, 0)]
    assert len(lat.neighbourhood(site(5, 5, 2))) == 26
    for (x, y, z) in ((0, 0, 0), (9, 9, 4), (5, 0, 2), (0, 5, 4)):
        xs, ys, zs = born_von_karman(x, 1, 10, "x"), born_von_karman(y, 1, 10, "y"), born_von_karman(z, 1, 5, "z")
        expect = [site(a, b, c) for a in xs for b in ys for c in zs if (a, b, c) != (x, y, z)]
        assert lat.neighbourhood(site(x, y, z)).tolist() == expect


def test_shard_partition():
    from hyperfluorescence_b200.distributed import shard
    seen = []
    for r in range(8):
        s = shard(64 * 1000, 64, r, 8)
        assert s.n_realisations == 8 and s.n_trajectories == 8000 and s.first_realisation == 8 * r
        seen += list(range(s.first_trajectory, s.first_trajectory + s.n_trajectories, 1000))
    assert seen == list(range(0, 64000, 1000))
    s = shard(100_000, 1, 3, 4)
    assert (s.first_trajectory, s.n_trajectories, s.first_realisation, s.n_realisations) == (75_000, 25_000, 0, 1)
    with pytest.raises(ValueError):
        shard(10, 3, 0, 2)


_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np
from hyperfluorescence_b200 import distributed as hd
from oracle import frm_oracle as fo
rank, _, world = hd.init_process_group("gloo")
dims, props, seed, total = (6, 5, 4), (0.6, 0.3, 0.1), 77, 24
sh = hd.shard(total, 1, rank, world)
lat = fo.OracleLattice.generate(dims, props, seed)
# each rank steps ITS shard of the global trajectory ids with the oracle (the CPU stand-in for
# the per-GPU kernel in this host-logic test) ...
loc = lat.run_batch(3, 0.1, seed, sh.first_trajectory, sh.n_trajectories, 400)
vec = np.array([loc["emission"], loc["recombination"], loc["injection"], loc["events"]])
# ... and the tallies are summed with the product's all-reduce
tot = hd.all_reduce_vector(vec)
whole = lat.run_batch(3, 0.1, seed, 0, total, 400)
assert tot.tolist() == [whole["emission"], whole["recombination"], whole["injection"], whole["events"]], (tot, whole)
assert whole["recombination"] > 0
# the exciton extension: each rank runs its shard of global trajectory ids with the specification oracle,
# the named tallies are summed with the product's all-reduce and equal the unsharded run
from oracle import exciton_oracle as xo
arr = lat.arrays()
ws1, wt1 = xo.weights(arr["s1"], arr["t1"], arr["species"])
o = xo.ExcitonOracle(dims, arr["species"], ws1, wt1, dict(xo.default_params(), cutoff=1.5))
mine = o.run_batch(seed, sh.first_trajectory, sh.n_trajectories, 120, n_threads=2)
summed = hd.all_reduce_named(mine)
assert summed == o.run_batch(seed, 0, total, 120, n_threads=2), summed
assert summed["events"] == total * 120
print("rank", rank, "ok", tot.tolist())
'''


def test_tally_all_reduce_gloo_world2(tmp_path):
    """world_size-2 gloo run of the N>1 host path: shard by global id, reduce, compare with the
    unsharded run."""
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29531", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script), ROOT], env=dict(env, RANK=str(r), LOCAL_RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
        assert "ok" in o
