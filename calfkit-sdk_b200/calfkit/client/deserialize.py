"""Declared in calfkit/client/_requests.py; re-exported under the reference's module path (reference calfkit/client/deserialize.py:15-89)."""
from calfkit.client._requests import _UNSET, _extract_output, deserialize_to_node_result  # noqa: F401

__all__ = ['_UNSET', '_extract_output', 'deserialize_to_node_result']
