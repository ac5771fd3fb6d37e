"""Declared in calfkit/client/_requests.py; re-exported under the reference's module path (reference calfkit/client/node_result.py:11-32)."""
from calfkit.client._requests import NodeResult  # noqa: F401

__all__ = ['NodeResult']
