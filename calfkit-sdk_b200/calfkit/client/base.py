"""Declared in calfkit/client/_requests.py; re-exported under the reference's module path (reference calfkit/client/base.py:27-172)."""
from calfkit.client._requests import BaseClient  # noqa: F401

__all__ = ['BaseClient']
