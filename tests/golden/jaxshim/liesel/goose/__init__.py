"""Import-time placeholder for liesel.goose (type names only)."""
from . import types  # noqa: F401
