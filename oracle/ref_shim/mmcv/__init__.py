"""Minimal stand-in for the `mmcv` package (TEST INFRASTRUCTURE ONLY).

The reference (Originofamonia/DFormer) imports five symbols from mmcv/mmengine
that are not installed in this image.  This shim restates their documented
behaviour so that the *unmodified* reference modules under /root/reference can
be imported by `oracle/make_golden.py` to pin the oracle.  It never ships on the
product path.
"""
__version__ = "2.1.0"
