"""Minimal stand-in for the few `diffusers==0.31.0` base-class symbols the reference
schedulers import.  TEST INFRASTRUCTURE ONLY: it exists so that
`tests/golden/make_golden.py` can import the UNMODIFIED reference from /root/reference
in the build container and record golden vectors.  Nothing in the product imports it.
"""
__version__ = "0.31.0-standin"
