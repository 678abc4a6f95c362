"""Host-side mirror of the ``toast.ops`` operators on the map-making hot path, with the
reference's operator names and trait (keyword) names."""

from .operator import Operator, Pipeline  # noqa: F401
from .pointing import (  # noqa: F401
    PixelsHealpix,
    PointingDetectorFP,
    PointingDetectorSimple,
    StokesWeights,
)
from .mapmaker_utils import (  # noqa: F401
    BinMap,
    BuildHitMap,
    BuildInverseCovariance,
    BuildNoiseWeighted,
    CovarianceAndHits,
    NoiseWeight,
    ScanMap,
    ScanMask,
)
from .mapmaker import MapMaker, TemplateMatrix  # noqa: F401
from .pixels_wcs import PixelsWCS  # noqa: F401
