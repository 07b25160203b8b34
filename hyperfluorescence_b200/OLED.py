"""The reference's driver script (``OLED.py:6-22``) against the GPU backend:
``python -m hyperfluorescence_b200.OLED``."""
from time import time

from .reseau import Lattice


def main(dimensions=(10, 10, 5), proportions=(0.84, 0.15, 0.01), OP=10**2, charges=4, **kwargs):
    start = time()
    test = Lattice(dimensions, proportions, charges=charges, **kwargs)
    test.operations(OP)
    end = time()
    print(f"IQE : {test.get_IQE()}")
    print(f"emissions : {test._emission}")
    print(f"injections : {test._injection}")
    print(f"{end - start} s")
    return test


if __name__ == "__main__":
    main()
