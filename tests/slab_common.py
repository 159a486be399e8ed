"""The slab-ring helpers live in the package (stochquant_b200/slabs.py); tests import them from here."""
from stochquant_b200.slabs import run_rank, split_slabs  # noqa: F401
