"""Import-path shim: ``from src.vqgan import VQModel`` (paella_inference.ipynb) resolves to paella_b200."""
