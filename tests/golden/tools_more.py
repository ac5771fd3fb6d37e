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
