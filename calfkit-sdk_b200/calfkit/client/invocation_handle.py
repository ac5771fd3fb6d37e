"""Declared in calfkit/client/_requests.py; re-exported under the reference's module path (reference calfkit/client/invocation_handle.py:13-41)."""
from calfkit.client._requests import InvocationHandle  # noqa: F401

__all__ = ['InvocationHandle']
