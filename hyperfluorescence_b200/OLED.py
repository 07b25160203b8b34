"""The reference's driver script (``OLED.py:6-22``) against the GPU backend:
``python -m hyperfluorescence_b200.OLED``.  Same device, same number of operations, same four output lines."""
import time

from .reseau import Lattice


def main(dimensions=(10, 10, 5), proportions=(0.84, 0.15, 0.01), OP=10**2, charges=4, **kwargs):
    t0 = time.perf_counter()
    lattice = Lattice(dimensions, proportions, charges=charges, **kwargs)
    lattice.operations(OP)
    elapsed = time.perf_counter() - t0
    for label, value in (("IQE", lattice.get_IQE()), ("emissions", lattice._emission), ("injections", lattice._injection)):
        print(f"{label} : {value}")
    print(f"{elapsed} s")
    return lattice


if __name__ == "__main__":
    main()
