"""Import-path shim for the reference's ``src/vqgan.py`` -> paella_b200.vqgan (same class surface)."""
from paella_b200.vqgan import ResBlock, VectorQuantize, VQModel  # noqa: F401
