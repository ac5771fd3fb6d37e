"""The wire format (reference calfkit/models/envelope.py:9-17).  The byte contract of the whole
hot path is `Envelope.model_dump_json()`."""
from pydantic import BaseModel, Field

from calfkit.models.session_context import SessionRunContext, WorkflowState


class Envelope(BaseModel):
    context: SessionRunContext
    internal_workflow_state: WorkflowState = Field(description="framework-level workflow state")
