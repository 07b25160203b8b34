"""Species of the reference's ``molecule.py``: ``Host`` (DPEPO), ``TADF`` (ACRSA),
``Fluorescent`` (TBPe), the ``EXCITON`` codes and ``Proportion``.

In the reference every lattice site *is* one of these objects.  Here the lattice lives on the
device (uint8 species + fp64 energies + per-trajectory flag sets); ``Lattice._grid[z][y][x]``
materialises one of these objects on demand as a snapshot of the device state.  The classes
also work stand-alone exactly like the reference's (own RNG, Gaussian-disordered levels,
flag toggles, spin statistics) so user code that builds molecules by hand keeps working.
"""
from __future__ import annotations

from dataclasses import dataclass
from random import Random

from .event import Point

# molecule.py:16-21 (HF_X_* in the C-ABI)
EXCITON: dict[str, int] = {"none": 0, "singlet": 1, "doublet": 2, "triplet": 3}

# (homo, lumo, s1, t1) means in eV and the species code used on the device (HF_SP_*)
_LEVELS = {
    "Host": (6.0, -2.0, 3.50, 3.00),          # molecule.py:358-361
    "TADF": (5.8, -2.6, 2.55, 2.52),          # molecule.py:252-255
    "Fluorescent": (5.25, -1.84, 2.69, 1.43),  # molecule.py:156-159
}
STANDARD_DEVIATION = 0.1


class Molecule:
    """Generic site: position, neighbour list, carrier flags, exciton state, private RNG
    (``molecule.py:23-105``)."""

    def __init__(self, position: Point, neighbours: list[Point]):
        self.position: Point = position
        self.neighbourhood: list[Point] = neighbours
        self.electron: bool = False
        self.hole: bool = False
        self.exciton: int = 0
        self.seed: Random = Random()

    def empty(self) -> bool:
        return not (self.electron or self.hole or self.exciton)

    def switch_electron(self) -> None:
        self.electron = not self.electron

    def switch_hole(self) -> None:
        self.hole = not self.hole

    def generate_exciton(self) -> None:
        """25 % singlet / 75 % triplet, only when both carriers are flagged (``molecule.py:88-96``)."""
        if self.electron and self.hole:
            self.exciton = EXCITON["singlet"] if 0 <= self.seed.random() < 0.25 else EXCITON["triplet"]

    def unbound_exciton(self) -> None:
        if self.exciton:
            self.exciton = EXCITON["none"]

    def exciton_decay(self) -> bool:
        raise NotImplementedError(f"{self.exciton_decay.__name__} is not implemented.")


class _Disordered(Molecule):
    """Shared constructor: four ``gauss`` draws (homo, lumo, s1, t1) on the molecule's own RNG."""
    _emissive = True

    def __init__(self, position, voisins, homo_energy, lumo_energy, s1_energy, t1_energy, standard_deviation):
        super().__init__(position, voisins)
        self.homo_energy: float = self.seed.gauss(homo_energy, standard_deviation)
        self.lumo_energy: float = self.seed.gauss(lumo_energy, standard_deviation)
        self.s1_energy: float = self.seed.gauss(s1_energy, standard_deviation)
        self.t1_energy: float = self.seed.gauss(t1_energy, standard_deviation)

    def exciton_decay(self) -> bool:
        """Clears the exciton and both carrier flags; True when a photon is emitted."""
        if self.exciton:
            state, self.exciton = self.exciton, EXCITON["none"]
            self.electron = False
            self.hole = False
            return self._emissive and state == EXCITON["singlet"]


class Fluorescent(_Disordered):
    """Blue fluorescent emitter TBPe (``molecule.py:108-199``): singlets emit."""

    def __init__(self, position: Point, voisins: list[Point], homo_energy: float = 5.25, lumo_energy: float = -1.84,
                 s1_energy: float = 2.69, t1_energy: float = 1.43, standard_deviation: float = 0.1) -> None:
        super().__init__(position, voisins, homo_energy, lumo_energy, s1_energy, t1_energy, standard_deviation)


class TADF(_Disordered):
    """TADF sensitizer ACRSA (``molecule.py:202-307``): singlets emit; ISC flips the spin."""

    def __init__(self, position: Point, voisins: list[Point], homo_energy: float = 5.8, lumo_energy: float = -2.6,
                 s1_energy: float = 2.55, t1_energy: float = 2.52, standard_deviation: float = 0.1) -> None:
        super().__init__(position, voisins, homo_energy, lumo_energy, s1_energy, t1_energy, standard_deviation)

    def intersystem_crossing(self) -> None:
        if self.exciton == EXCITON["singlet"]:
            self.exciton = EXCITON["triplet"]
        elif self.exciton == EXCITON["triplet"]:
            self.exciton = EXCITON["singlet"]


class Host(_Disordered):
    """Host matrix DPEPO (``molecule.py:310-400``): never emits in the visible."""
    _emissive = False

    def __init__(self, position: Point, voisins: list[Point], homo_energy: float = 6.0, lumo_energy: float = -2.0,
                 s1_energy: float = 3.50, t1_energy: float = 3.00, standard_deviation: float = 0.1) -> None:
        super().__init__(position, voisins, homo_energy, lumo_energy, s1_energy, t1_energy, standard_deviation)


SPECIES_CLASSES = (Host, TADF, Fluorescent)          # index = device species code


def molecule_from_device(code: int, position: Point, neighbours: list[Point], homo: float, lumo: float,
                         s1: float, t1: float, electron: bool, hole: bool, exciton: int):
    """A snapshot object for one site of a device-resident lattice (no RNG draws)."""
    cls = SPECIES_CLASSES[code]
    mol = cls.__new__(cls)
    Molecule.__init__(mol, position, neighbours)
    mol.homo_energy, mol.lumo_energy, mol.s1_energy, mol.t1_energy = homo, lumo, s1, t1
    mol.electron, mol.hole, mol.exciton = electron, hole, exciton
    return mol


@dataclass
class Proportion:
    """Fractions of each species in the lattice (``molecule.py:403-418``)."""
    host: float
    tadf: float
    fluo: float
