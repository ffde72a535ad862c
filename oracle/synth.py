"""Test-side alias of the package's synthetic data generator (geoa3_b200/synth.py: pure numpy, no kernels),
kept so that oracle-side scripts and tests read `from oracle import synth`."""
from geoa3_b200.synth import CLASS_IDS, lattice_cloud, make_batch, make_instance, make_offsets  # noqa: F401
