"""Final-output content parts (reference calfkit/models/payload.py:6-35): declared in calfkit/models/wire.py, re-exported under the reference's module path."""
from calfkit.models.wire import ContentPart, DataPart, FilePart, TextPart, ToolCallPart  # noqa: F401

__all__ = ['ContentPart', 'DataPart', 'FilePart', 'TextPart', 'ToolCallPart']
