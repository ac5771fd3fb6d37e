from calfkit.nodes.agent import Agent, BaseAgentNodeDef, FunctionModelClient
from calfkit.nodes.base import BaseNodeDef
from calfkit.nodes.node import NodeDef
from calfkit.nodes.tool import BaseToolNodeDef, ToolNodeDef, agent_tool

__all__ = ["Agent", "BaseAgentNodeDef", "BaseNodeDef", "BaseToolNodeDef", "FunctionModelClient", "NodeDef", "ToolNodeDef",
           "agent_tool"]
