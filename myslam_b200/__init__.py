"""myslam_b200 -- B200-native (sm_100a) implementation of ESLAM's per-iteration render-and-optimise
hot path behind the reference's Python surface.  See DESIGN.md / INTEGRATION.md.

Importing the package never touches CUDA; the shared library is loaded on first use and there is
no CPU or eager-PyTorch fallback (calls raise if it is missing)."""
from .decoders import Decoders  # noqa: F401
from .renderer import Renderer, ReplayDraws, TorchDraws  # noqa: F401
from .field import FieldStore  # noqa: F401
from .tracker import TrackerStep, optimize_tracking, track_frame  # noqa: F401
from .mapper import MapperStep, map_window, optimize_mapping  # noqa: F401
from .mesher import eval_points, grid_axes, hull_planes, query_grid_sdf  # noqa: F401
from .ingest import ingest_frame  # noqa: F401
from .install import install  # noqa: F401

__version__ = "0.1.0"
