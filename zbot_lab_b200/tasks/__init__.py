"""Importing this package registers the task ids (reference: zbot/tasks/__init__.py:3-13)."""
from . import zbot6_direct  # noqa: F401
from . import zbot6b_direct  # noqa: F401
from . import zbotlab_manager  # noqa: F401
