from calfkit.models.actions import (Call, Delegate, Emit, NodeResult, Parallel, Reply, ReturnCall,
                                    Sequential, Silent, TailCall, _Call)
from calfkit.models.envelope import Envelope
from calfkit.models.payload import ContentPart, DataPart, FilePart, TextPart, ToolCallPart
from calfkit.models.session_context import (BaseSessionRunContext, CallFrame, CallFrameStack, Deps,
                                            SessionRunContext, Stack, WorkflowState)
from calfkit.models.state import (BaseAgentActivityState, CoreMessageState, InFlightToolsState,
                                  OverridesState, PendingToolBatch, State)
from calfkit.models.tool_context import ToolContext

__all__ = [
    "Call", "Delegate", "Emit", "NodeResult", "Parallel", "Reply", "ReturnCall", "Sequential",
    "Silent", "TailCall", "_Call", "Envelope", "ContentPart", "DataPart", "FilePart", "TextPart",
    "ToolCallPart", "BaseSessionRunContext", "CallFrame", "CallFrameStack", "Deps",
    "SessionRunContext", "Stack", "WorkflowState", "BaseAgentActivityState", "CoreMessageState",
    "InFlightToolsState", "OverridesState", "State", "PendingToolBatch", "ToolContext",
]
