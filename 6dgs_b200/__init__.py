"""sixdgs-b200: the single-query pose-estimation hot path of 6DGS (mbortolon97/6dgs) rebuilt for
NVIDIA B200 (sm_100a): candidate-ray generation, ray key cache, softmax-over-rays scoring, top-k and
least-squares pose solve as hand-written CUDA behind a C ABI (include/sixdgs.h), exposed through the
reference's own Python entry points.

The directory is called ``6dgs_b200`` (not a valid identifier): import it with
``importlib.import_module("6dgs_b200")`` or through the alias module ``import sixdgs_b200``.
"""
import sys as _sys

from . import _lib, ops, synthetic  # noqa: F401
from ._lib import SixdgsError  # noqa: F401
from .eig3 import sym_eig_3x3  # noqa: F401
from .evaluate import test_pose_estimation  # noqa: F401
from .identification import IdentificationModule, MultiHeadAttention, RayKeyCache, RayPreprocessor  # noqa: F401
from .image_tokens import BackboneWrapper, DinoV2ViTS14, create_backbone  # noqa: F401
from .camera_up import CameraDirectionPredictor  # noqa: F401
from .losses import DistanceBasedScoreLoss, best_one_to_one_rays_selector  # noqa: F401
from .pose_solve import (compute_line_intersection_impl2, exclude_negatives, make_rotation_mat,  # noqa: F401
                         pose_from_topk)
from .raygen import generate_all_possible_rays, quadricell_cells  # noqa: F401
from .scene import GaussianScene  # noqa: F401
from .sharding import CudaBackend, ShardedPoseEstimator  # noqa: F401

test_pose_estimation.__test__ = False  # not a pytest test despite the reference's name

__version__ = "0.2.0"

# make `import sixdgs_b200.<sub>` resolve to the same module objects
for _name, _mod in list(_sys.modules.items()):
    if _name == __name__ or _name.startswith(__name__ + "."):
        _sys.modules.setdefault("sixdgs_b200" + _name[len(__name__):], _mod)
