"""More tool signatures for the schema-parity goldens (tests/golden/make_golden_schemas.py).  Imported by both the
reference harness and the product: plain functions only, no imports from either package."""
from typing import Literal, Optional


def with_defaults(city: str, units: str = "metric", days: int = 3) -> str:
    """Forecast for a city.

    Args:
        city: The city to look up.
        units: metric or imperial.
        days: How many days ahead.
    """
    return f"{city}/{units}/{days}"


def numeric(a: int, b: float, flag: bool) -> float:
    """Mixes scalar types."""
    return a + b if flag else a - b


def optional_and_lists(names: list[str], limit: Optional[int] = None, tags: list[str] | None = None) -> list[str]:
    """Filter names.

    Args:
        names: candidate names
        limit: optional cap
        tags: optional tags
    """
    return names[: limit or len(names)]


def nested(payload: dict[str, list[int]], mode: Literal["fast", "slow"] = "fast") -> dict:
    """Takes a nested mapping."""
    return {"mode": mode, "n": sum(len(v) for v in payload.values())}


def no_doc(x: str) -> str:
    return x


def contextual(ctx, query: str, top_k: int = 5) -> str:
    """Search with access to the tool context.

    Args:
        query: what to search for
        top_k: number of hits
    """
    return f"{query}:{top_k}"


def with_ctx(ctx: "ToolContext", q: str, n: int = 2) -> str:  # noqa: F821  (the importer sets tools_more.ToolContext to ITS context class)
    """Uses the injected tool context; the context parameter must not appear in the schema.

    Args:
        q: the query
        n: repetitions
    """
    return f"{ctx.deps.provided_deps.get('tenant')}:{q * n}"


def numpy_style(alpha: float, beta: int = 1) -> float:
    """Scale alpha.

    Longer explanation line.

    Parameters
    ----------
    alpha : float
        The value to scale,
        continued on a second line.
    beta : int
        The factor.

    Returns
    -------
    float
        The product.
    """
    return alpha * beta


def sphinx_style(path: str, recursive: bool = False) -> list[str]:
    """List files.

    :param path: where to look
    :param recursive: descend into
        sub-directories
    :returns: names
    """
    return [path] if recursive else []


def google_multiline(text: str, width: int = 80) -> str:
    """Wrap text.

    Second paragraph of the summary.

    Args:
        text: the text to wrap, which may be
            long and span lines.
        width (int): column limit.

    Returns:
        The wrapped text.

    Raises:
        ValueError: never.
    """
    return text[:width]


MORE = {f.__name__: f for f in (with_defaults, numeric, optional_and_lists, nested, no_doc, contextual, with_ctx, numpy_style, sphinx_style,
                                google_multiline)}


# ---- return-value variety (tests/golden/make_golden_returns.py): how a tool's Python return value lands in the envelope
import dataclasses as _dc
import datetime as _dt
import decimal as _dec
import enum as _enum
import uuid as _uuid

import pydantic as _pyd


class Reading(_pyd.BaseModel):
    city: str
    temp: float
    tags: list[str] = []


@_dc.dataclass
class Point:
    x: int
    y: float


class Color(_enum.Enum):
    RED = "red"


def returns_dict(k: str) -> dict:
    """d"""
    return {"k": k, "n": [1, 2.5, None, True], "nested": {"é": "ü\n\t\"q\""}}


def returns_list(n: int) -> list:
    """l"""
    return [i * 1.5 for i in range(n)]


def returns_none(x: str) -> None:
    """n"""
    return None


def returns_bool(x: str) -> bool:
    """b"""
    return x == "yes"


def returns_model(city: str) -> Reading:
    """m"""
    return Reading(city=city, temp=21.700000000000003, tags=["a"])


def returns_dataclass(x: int) -> Point:
    """dc"""
    return Point(x=x, y=0.1 + 0.2)


def returns_datetime(x: str) -> _dt.datetime:
    """dt"""
    return _dt.datetime(2026, 1, 2, 3, 4, 5, 600, tzinfo=_dt.timezone.utc)


def returns_tuple_set(x: str) -> tuple:
    """t"""
    return (x, 1, frozenset([3]))


def returns_misc(x: str) -> dict:
    """misc"""
    return {"dec": _dec.Decimal("1.50"), "uuid": _uuid.UUID(int=5), "enum": Color.RED, "bytes": b"hi", "date": _dt.date(2026, 1, 2),
            "big": 2 ** 70, "neg0": -0.0, "exp": 1e22, "small": 1e-7}


def inspect_ctx(ctx: "ToolContext", x: str) -> dict:  # noqa: F821
    """what a contextual tool can see"""
    return {"x": x, "tenant": ctx.deps.provided_deps.get("tenant"), "corr": ctx.deps.correlation_id, "agent": ctx.agent_name,
            "tool_call_id": ctx.tool_call_id, "tool": ctx.tool_name, "run_id": ctx.run_id, "n_messages": len(ctx.messages),
            "first_kind": ctx.messages[0].kind if ctx.messages else None}


RETURNS = {f.__name__: (f, a) for f, a in ((returns_dict, {"k": "v"}), (returns_list, {"n": 4}), (returns_none, {"x": "a"}), (returns_bool, {"x": "yes"}),
                                            (returns_model, {"city": "Kraków"}), (returns_dataclass, {"x": 3}), (returns_datetime, {"x": "a"}),
                                            (returns_tuple_set, {"x": "a"}), (returns_misc, {"x": "a"}), (inspect_ctx, {"x": "v"}))}
