"""Import-path shim: ``from utils.modules import Paella`` / ``from utils.alter_attention import
replace_attention_layers`` (paella_inference.ipynb) resolve to paella_b200."""
