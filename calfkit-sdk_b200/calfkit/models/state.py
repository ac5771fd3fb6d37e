"""Agent state carried inside every Envelope (reference calfkit/models/state.py:18-141): declared in calfkit/models/wire.py, re-exported under the reference's module path."""
from calfkit.models.wire import BaseAgentActivityState, CoreMessageState, InFlightToolsState, OverridesState, PendingToolBatch, State  # noqa: F401

__all__ = ['BaseAgentActivityState', 'CoreMessageState', 'InFlightToolsState', 'OverridesState', 'PendingToolBatch', 'State']
