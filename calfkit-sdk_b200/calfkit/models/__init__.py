"""calfkit.models — same public names as the reference package (calfkit/models/__init__.py); the wire models live in
calfkit/models/wire.py, the result actions in calfkit/models/actions.py."""
from calfkit.models import actions as _actions
from calfkit.models import wire as _wire
from calfkit.models.tool_context import ToolContext

_EXPORTS = {
    _actions: ("Call", "NodeResult", "ReturnCall", "Silent", "TailCall", "_Call"),
    _wire: ("Envelope", "ContentPart", "DataPart", "FilePart", "TextPart", "ToolCallPart", "BaseSessionRunContext", "CallFrame",
            "CallFrameStack", "Deps", "SessionRunContext", "Stack", "WorkflowState", "BaseAgentActivityState", "CoreMessageState",
            "InFlightToolsState", "OverridesState", "State", "PendingToolBatch"),
}
__all__ = ["ToolContext"]
for _mod, _names in _EXPORTS.items():
    for _n in _names:
        globals()[_n] = getattr(_mod, _n)
        __all__.append(_n)
del _mod, _names, _n
