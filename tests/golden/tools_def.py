"""Pure tool functions shared by tests/golden/make_golden.py (run against the real reference)
and the parity tests (run against the oracle port and the CUDA path)."""


def get_weather(location: str) -> str:
    """Get the current weather at a location"""
    return f"It's sunny in {location}"


def get_temperature(location: str) -> dict:
    """Structured weather"""
    return {"location": location, "temp_c": 21, "humid": 0.5, "tags": ["a", "b"], "ok": True, "none": None}


def count_chars(location: str) -> int:
    """Length of the name"""
    return len(location)


def no_args() -> str:
    """Takes nothing"""
    return "pong\n\t\"quoted\" \\ back"


TOOLS = {f.__name__: f for f in (get_weather, get_temperature, count_chars, no_args)}
