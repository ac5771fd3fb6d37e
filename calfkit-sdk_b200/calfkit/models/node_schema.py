"""Routing data of a node (reference calfkit/models/node_schema.py:6-21): declared in calfkit/models/wire.py, re-exported under the reference's module path."""
from calfkit.models.wire import BaseNodeSchema, BaseToolNodeSchema  # noqa: F401

__all__ = ['BaseNodeSchema', 'BaseToolNodeSchema']
