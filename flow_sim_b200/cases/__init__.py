"""The three shipped cases of the reference (cases/example, cases/akbari_firoozi, cases/gerd_roseires) and a synthetic
polyline-section reach built on the mirror API; each ``build_*`` returns an un-run ``PreissmannSolver`` plus the ``run()`` keyword arguments."""
from .akbari_firoozi import build as build_akbari
from .example import build as build_example
from .gerd_roseires import build as build_gerd
from .irregular import build as build_irregular
from .irregular import build_mixed
