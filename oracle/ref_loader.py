"""Locate and import the UNMODIFIED reference (test / baseline infrastructure): `oracle/_ref/` (the snapshot made by
oracle/build_ref.py, present on the GPU box) or `/root/reference` (authoring container), with the import shims for the
diffusers / pytorch_lightning base classes first on sys.path."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def ref_root():
    for p in (os.path.join(HERE, "_ref"), "/root/reference"):
        if os.path.isdir(os.path.join(p, "lvdm")):
            return p
    return None


def enable():
    """Put the reference on sys.path; returns its root or None when no copy is available."""
    root = ref_root()
    if root is None:
        return None
    for p in (root, os.path.join(HERE, "shim")):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    return root
