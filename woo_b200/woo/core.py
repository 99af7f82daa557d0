"""woo.core of the reference: Scene, Cell, Node, Engine."""
from .._woo import Scene, Cell, Node, Engine  # noqa: F401
