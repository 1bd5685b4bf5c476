"""Drop-in mirror of the reference package ``hydromodel`` (cve-mohd/flow-sim, src/hydromodel): same class
names, constructor arguments and result attributes; ``PreissmannSolver.run`` executes on the GPU.

``install()`` registers this package under the names the reference's case scripts import
(``hydromodel.*`` and ``src.hydromodel.*``), so those scripts run unmodified on top of it.
"""
import sys

from . import boundary, channel, cross_section, hydraulics, hydrograph, lumped_storage, preissmann, rating_curve, solver, utility
from .boundary import Boundary
from .channel import Channel
from .cross_section import IrregularSection, TrapezoidalSection, interpolate_cross_section
from .hydrograph import Hydrograph
from .lumped_storage import LumpedStorage
from .preissmann import PreissmannSolver
from .rating_curve import RatingCurve

_SUBMODULES = ["boundary", "channel", "cross_section", "hydraulics", "hydrograph", "lumped_storage", "preissmann",
               "rating_curve", "solver", "utility"]


def install(names=("hydromodel", "src.hydromodel")) -> None:
    import types

    me = sys.modules[__name__]
    for name in names:
        if "." in name:
            parent = name.split(".")[0]
            sys.modules.setdefault(parent, types.ModuleType(parent))
        sys.modules[name] = me
        for sub in _SUBMODULES:
            sys.modules[f"{name}.{sub}"] = sys.modules[f"{__name__}.{sub}"]
        if "." in name:
            setattr(sys.modules[name.split(".")[0]], name.split(".")[1], me)
