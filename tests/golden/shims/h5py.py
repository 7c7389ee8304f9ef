"""h5py.File stand-in over oqrl_b200.hdf5_subset (TEST INFRASTRUCTURE, see pennylane.py in this directory)."""
from oqrl_b200.hdf5_subset import File  # noqa: F401
