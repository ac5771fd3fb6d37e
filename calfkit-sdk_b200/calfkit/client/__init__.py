from calfkit.client.base import BaseClient
from calfkit.client.client import Client
from calfkit.client.invocation_handle import InvocationHandle
from calfkit.client.node_result import NodeResult

__all__ = ["BaseClient", "Client", "InvocationHandle", "NodeResult"]
