"""Import-time placeholder for liesel.contrib."""
from . import splines  # noqa: F401
