"""bart_rs_b200 — B200-native Particle-Gibbs BART sweep behind the bart-rs sampler surface.

Only the hot path of GStechschulte/bart-rs is here: `pymc_bart` mirrors the
reference's compiled module (PyBartSettings / PySampler), `pgbart` mirrors the PGBART step method's
parameter derivations and astep loop.  The compute lives in csrc/ (CUDA, sm_100a) behind the C ABI
of include/pgbart.h; importing this package does not load the library until a sampler is created.
"""
from .pgbart import PGBART, calculate_max_tree_depth  # noqa: F401
from .pymc_bart import DeviceLikelihood, PyBartSettings, PySampler  # noqa: F401

__all__ = ["DeviceLikelihood", "PyBartSettings", "PySampler", "PGBART", "calculate_max_tree_depth"]
