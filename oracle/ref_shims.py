"""Import shims that let the UNMODIFIED reference (/root/reference) run on CPU in the build
container.  TEST INFRASTRUCTURE ONLY: used by oracle/gen_golden.py to produce tests/golden/*.npz.
Nothing on the GPU box imports this file (``/root/reference`` does not exist there).

Shims: stub ``plyfile`` and ``simple_knn._C`` (only needed at import of
scene/gaussian_model.py:23,25) and replace ``pose_estimation.backbone.create_backbone``
(backbone.py:6-22 calls torch.hub -> network) by a deterministic fake feature extractor.
"""
import sys
import types

import torch

REFERENCE_ROOT = "/root/reference"


class _PlyElement:
    """The two plyfile calls scene/gaussian_model.py:save_ply makes (:331-332), writing what plyfile writes for a
    structured float32 array: `element <name> <count>` + one `property float <field>` line per field, then the packed
    little-endian records.  (plyfile itself is not installed in the build container.)"""

    def __init__(self, data, name):
        self.data, self.name = data, name

    @staticmethod
    def describe(data, name):
        assert all(data.dtype[f].str in ("<f4", "=f4") for f in data.dtype.names), "save_ply only writes f4 fields"
        return _PlyElement(data, name)


class _PlyData:
    def __init__(self, elements):
        self.elements = list(elements)

    def write(self, path):
        with open(path, "wb") as fh:
            fh.write(b"ply\nformat binary_little_endian 1.0\n")
            for el in self.elements:
                fh.write(f"element {el.name} {len(el.data)}\n".encode("ascii"))
                for f in el.data.dtype.names:
                    fh.write(f"property float {f}\n".encode("ascii"))
            fh.write(b"end_header\n")
            for el in self.elements:
                fh.write(el.data.astype(el.data.dtype.newbyteorder("<")).tobytes())


def install(backbone_factory=None):
    """``backbone_factory()`` must return a module exposing ``forward_features`` (default: the
    product package's SyntheticBackbone so reference and product see the same fake features)."""
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    if "plyfile" not in sys.modules:
        m = types.ModuleType("plyfile")
        m.PlyData = _PlyData
        m.PlyElement = _PlyElement
        sys.modules["plyfile"] = m
    if "simple_knn" not in sys.modules:
        pkg = types.ModuleType("simple_knn")
        sub = types.ModuleType("simple_knn._C")
        sub.distCUDA2 = None
        pkg._C = sub
        sys.modules["simple_knn"] = pkg
        sys.modules["simple_knn._C"] = sub
    import pose_estimation.backbone as bb

    if backbone_factory is None:
        import importlib
        if "/root/repo" not in sys.path:
            sys.path.insert(0, "/root/repo")
        backbone_factory = importlib.import_module("6dgs_b200.synthetic").SyntheticBackbone

    def create_backbone(type="dino", **kwargs):
        return backbone_factory(), (16, 16), 384

    bb.create_backbone = create_backbone


def make_gaussian_model(xyz, scaling_raw, rotation, f_dc, f_rest, sh_degree=3):
    """Build a reference GaussianModel by attribute assignment (create_from_*/load_ply hard-code
    device='cuda', scene/gaussian_model.py:187,394)."""
    install()
    from scene.gaussian_model import GaussianModel

    gm = GaussianModel(sh_degree)
    gm._xyz = xyz
    gm._scaling = scaling_raw
    gm._rotation = rotation
    gm._features_dc = f_dc
    gm._features_rest = f_rest
    gm._opacity = torch.zeros(xyz.shape[0], 1)
    gm.active_sh_degree = sh_degree
    return gm
