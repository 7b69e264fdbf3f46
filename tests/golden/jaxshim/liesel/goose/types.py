"""`ModelInterface` / `ModelState` are only used as annotations on the hot path."""
from typing import Any

ModelInterface = Any
ModelState = Any
Position = dict
