"""TEST INFRASTRUCTURE ONLY — loads the UNMODIFIED reference (calf-ai/calfkit-sdk) from
/root/reference so golden vectors can be generated from the reference's own code.

The reference cannot be `import calfkit`-ed in this image: faststream, aiokafka, uuid_utils,
pydantic_graph and genai_prices are not installed and there is no network.  The recipe below
 pre-registers *bare* package modules so the reference's `__init__.py` files do
not run, stubs the absent third-party modules with the minimum surface the hot-path modules
touch at import time, and then imports these reference modules unmodified:

    calfkit.models.*            (wire format: Envelope, State, CallFrame ...)
    calfkit.nodes.base          (BaseNodeDef.handler / prepare_context / _publish_action)
    calfkit.nodes.tool          (ToolNodeDef.run, agent_tool)

/root/reference exists only in the build container.  oracle/build_ref.py (run by __graft_entry__.build()) mirrors
the reference's package tree byte for byte into oracle/_ref/ (git-ignored build output that ships with the gpurun
snapshot), and this module loads the reference from there when /root/reference is absent — so bench.py's CPU arm
times the reference's own code on the GPU box's host cores.  Used by tests/golden/make_golden*.py (here) and by
bench.py's `--impl reference` / cpu_baseline legs.  Never imported by the product.
"""
from __future__ import annotations

import importlib
import itertools
import sys
import types

import os

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = "/root/reference" if os.path.isdir("/root/reference/calfkit") else os.path.join(_HERE, "_ref")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "calfkit", "nodes"))

_uuid_counter = itertools.count(1)
_uuid_hook = None  # callable() -> 32-hex string, installed by tests for deterministic frame ids


def set_uuid_source(fn) -> None:
    """Install a deterministic id source for `uuid_utils.uuid7().hex` (CallFrame.frame_id,
    reference calfkit/models/session_context.py:38)."""
    global _uuid_hook
    _uuid_hook = fn


class _FakeUUID:
    def __init__(self, hexstr: str):
        self.hex = hexstr

    def __str__(self) -> str:
        h = self.hex
        return f"{h[:8]}-{h[8:12]}-{h[12:16]}-{h[16:20]}-{h[20:]}"


def _uuid7():
    if _uuid_hook is not None:
        return _FakeUUID(_uuid_hook())
    return _FakeUUID(f"{next(_uuid_counter):032x}")


def _bare_package(name: str, path: str) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__path__ = [path]
    m.__package__ = name
    sys.modules[name] = m
    return m


def _stub(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


_loaded = None


def load_reference():
    """Returns a namespace with the reference's hot-path symbols.  Idempotent."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if "calfkit" in sys.modules and not getattr(sys.modules["calfkit"], "__ref_harness__", False):
        raise RuntimeError("a different `calfkit` is already imported in this process; "
                           "run the reference harness in its own interpreter")

    # --- bare packages (their __init__.py would import faststream / providers) -------------
    pk = _bare_package("calfkit", f"{REF_ROOT}/calfkit")
    pk.__ref_harness__ = True
    _bare_package("calfkit._vendor", f"{REF_ROOT}/calfkit/_vendor")
    pai = _bare_package("calfkit._vendor.pydantic_ai", f"{REF_ROOT}/calfkit/_vendor/pydantic_ai")
    _bare_package("calfkit.nodes", f"{REF_ROOT}/calfkit/nodes")

    # --- absent third-party modules ------------------------------------------------------
    _stub("uuid_utils", uuid7=_uuid7, uuid4=_uuid7)

    class AbstractSpan:  # pydantic_graph._utils.AbstractSpan (type annotation only)
        pass

    pg = _stub("pydantic_graph")
    pg._utils = _stub("pydantic_graph._utils", AbstractSpan=AbstractSpan)

    gp = _stub("genai_prices", calc_price=lambda *a, **k: None)
    gp.types = _stub("genai_prices.types", PriceCalculation=object, Usage=object)
    gp.data_snapshot = _stub("genai_prices.data_snapshot", get_snapshot=lambda: None)

    def Context(*a, **k):  # faststream.Context() marker used in Annotated[...] only
        return None

    class BaseMiddleware:
        pass

    class PublishCommand:
        pass

    class FastStream:
        def __init__(self, *a, **k):
            pass

    class KafkaBroker:
        pass

    fs = _stub("faststream", Context=Context, BaseMiddleware=BaseMiddleware,
               PublishCommand=PublishCommand, FastStream=FastStream)
    fs.kafka = _stub("faststream.kafka", KafkaBroker=KafkaBroker, TestKafkaBroker=KafkaBroker)
    fs.kafka.annotations = _stub("faststream.kafka.annotations", KafkaBroker=KafkaBroker)

    sys.path.insert(0, REF_ROOT)
    try:
        tools = importlib.import_module("calfkit._vendor.pydantic_ai.tools")
        pai.Tool = tools.Tool
        messages = importlib.import_module("calfkit._vendor.pydantic_ai.messages")
        models = importlib.import_module("calfkit.models")
        envelope = importlib.import_module("calfkit.models.envelope")
        session = importlib.import_module("calfkit.models.session_context")
        state = importlib.import_module("calfkit.models.state")
        node_schema = importlib.import_module("calfkit.models.node_schema")
        base = importlib.import_module("calfkit.nodes.base")
        tool = importlib.import_module("calfkit.nodes.tool")
    finally:
        sys.path.remove(REF_ROOT)

    _loaded = types.SimpleNamespace(
        tools=tools, messages=messages, models=models, envelope=envelope, session=session,
        state=state, node_schema=node_schema, base=base, tool=tool,
        Envelope=envelope.Envelope, State=state.State, agent_tool=tool.agent_tool,
        ToolNodeDef=tool.ToolNodeDef, BaseNodeDef=base.BaseNodeDef,
    )
    return _loaded


class CaptureBroker:
    """Capture-only stand-in for the FastStream KafkaBroker the reference publishes through
    (call sites: reference calfkit/nodes/base.py:82,99,113,130).  Serialises with the byte
    contract SURVEY.md §8c fixes: `Envelope.model_dump_json()`."""

    def __init__(self):
        self.published = []  # (topic, key bytes|None, correlation_id, payload bytes)

    async def publish(self, msg, topic, correlation_id=None, key=None, **kw):
        self.published.append((topic, key, correlation_id, msg.model_dump_json().encode()))
