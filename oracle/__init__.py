"""ORACLE -- test infrastructure only (see ref_unet.py).  Never imported by the product package."""
