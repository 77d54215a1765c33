"""Shim: the plain-torch restatement of the reference forward lives in oracle/torch_ref.py (test infrastructure;
bench.py's reference arm times it, tests compare against it)."""
from oracle.torch_ref import *          # noqa: F401,F403
from oracle.torch_ref import unet_forward, grads_with, tf32_round   # noqa: F401
