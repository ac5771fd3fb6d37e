"""The wire format (reference calfkit/models/envelope.py:9-17): declared in calfkit/models/wire.py, re-exported under the reference's module path."""
from calfkit.models.wire import Envelope  # noqa: F401

__all__ = ['Envelope']
