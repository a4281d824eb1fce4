"""Minimal stand-in for timm, used ONLY when the real package is absent (add <repo>/compat/optional to PYTHONPATH)."""
