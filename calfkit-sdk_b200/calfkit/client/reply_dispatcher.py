"""Declared in calfkit/client/_requests.py; re-exported under the reference's module path (reference calfkit/client/reply_dispatcher.py:15-53)."""
from calfkit.client._requests import _ReplyDispatcher  # noqa: F401

__all__ = ['_ReplyDispatcher']
