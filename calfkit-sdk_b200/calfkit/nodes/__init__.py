"""calfkit.nodes — node definitions of the hot path (reference calfkit/nodes/__init__.py) + the model-client shim."""
from calfkit.nodes.base import BaseNodeDef
from calfkit.nodes.node import NodeDef
from calfkit.nodes.tool import BaseToolNodeDef, ToolNodeDef, agent_tool
from calfkit.nodes.agent import Agent, BaseAgentNodeDef, FunctionModelClient  # noqa: I001  (agent builds on tool)

__all__ = sorted(n for n in dir() if not n.startswith("_") and n[0].isupper() or n == "agent_tool")
