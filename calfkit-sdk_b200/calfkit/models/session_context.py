"""Call stack + session context (reference calfkit/models/session_context.py:13-91): declared in calfkit/models/wire.py, re-exported under the reference's module path."""
from calfkit.models.wire import BaseSessionRunContext, CallFrame, CallFrameStack, Deps, SessionRunContext, Stack, WorkflowState  # noqa: F401

__all__ = ['BaseSessionRunContext', 'CallFrame', 'CallFrameStack', 'Deps', 'SessionRunContext', 'Stack', 'WorkflowState']
