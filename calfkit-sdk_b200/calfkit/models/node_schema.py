"""Routing data of a node: id, subscribe topics, publish topic
(reference calfkit/models/node_schema.py:6-21)."""
from dataclasses import KW_ONLY, dataclass

from calfkit.models.messages import ToolDefinition


@dataclass
class BaseNodeSchema:
    _: KW_ONLY
    node_id: str
    subscribe_topics: list[str]
    publish_topic: str | None

    def __post_init__(self) -> None:
        if not isinstance(self.subscribe_topics, (list, tuple)):
            self.subscribe_topics = [self.subscribe_topics]


@dataclass
class BaseToolNodeSchema(BaseNodeSchema):
    _: KW_ONLY
    tool_schema: ToolDefinition
