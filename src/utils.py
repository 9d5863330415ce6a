"""Import-path shim for the reference's ``src/utils.py``: ``sample`` with the reference signature."""
from paella_b200.utils import load_conditional_models, sample  # noqa: F401
