"""``basix.utils`` (python/basix/utils.py)."""

from basix._basixcpp import index  # noqa: F401
