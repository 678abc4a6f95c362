"""Templates of the destriping map-maker (only ``Offset`` is on the hot path)."""

from .amplitudes import Amplitudes, AmplitudesMap  # noqa: F401
from .offset import Offset, Template  # noqa: F401
