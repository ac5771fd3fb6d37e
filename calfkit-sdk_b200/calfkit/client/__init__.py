"""calfkit.client — Client / BaseClient / InvocationHandle / NodeResult, as in the reference package."""
from calfkit.client._requests import BaseClient, InvocationHandle, NodeResult
from calfkit.client.client import Client

__all__ = sorted(["Client", "BaseClient", "InvocationHandle", "NodeResult"])
