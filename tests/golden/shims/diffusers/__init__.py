"""Stand-in for the parts of `diffusers==0.16.0` that the reference UNet imports.

TEST INFRASTRUCTURE ONLY (used by tests/golden/make_golden.py inside the build
container to import the unmodified reference from /root/reference).  Written
from the published diffusers 0.16.0 behaviour; the reference keeps an in-tree
mirror of FeedForward/GEGLU at vsr/models/diffusers_attention.py:734-822 and of
the sinusoidal embedding at base/models/utils.py:74-94.
"""
__version__ = "0.16.0"
