"""Scene container mirroring the getters of the reference GaussianModel that the hot path reads
(reference scene/gaussian_model.py:125-158): ``get_xyz``, ``get_scaling``, ``get_rotation``,
``get_rotation_mat()``, ``get_features``, ``active_sh_degree``, ``max_sh_degree``.

HBM layout: four packed fp32 arrays, 232 B per Gaussian -- xyz[N,3], log-scale[N,3],
quaternion[N,4] (w,x,y,z, un-normalised, as stored in the PLY) and SH features[N,16,3]
(coefficient-major, the layout ``get_features`` returns) -- exactly what the ray-generation kernel
gathers per selected ellipsoid.  Training state (optimizer, densification) is out of scope.
"""
from __future__ import annotations

import numpy as np
import torch


class GaussianScene:
    def __init__(self, xyz, scaling, rotation, features_dc, features_rest, sh_degree: int = 3, device=None):
        dev = torch.device(device) if device is not None else xyz.device
        f = lambda t: torch.as_tensor(t).detach().to(device=dev, dtype=torch.float32).contiguous()  # noqa: E731
        self._xyz = f(xyz)
        self._scaling = f(scaling)
        self._rotation = f(rotation)
        # [N, (max_deg+1)^2, 3]: 16 coefficients for the usual degree-3 storage, fewer for low-degree models.  A model
        # may be evaluated below its stored degree (3DGS raises active_sh_degree during training, gaussian_model.py:121-123)
        self._features = torch.cat((f(features_dc), f(features_rest)), dim=1).contiguous()
        n = self._xyz.shape[0]
        assert self._scaling.shape == (n, 3) and self._rotation.shape == (n, 4)
        nc = self._features.shape[1]
        self.max_sh_degree = int(round(nc ** 0.5)) - 1
        if self._features.shape != (n, nc, 3) or (self.max_sh_degree + 1) ** 2 != nc:
            raise ValueError(f"SH features must be [N, (deg+1)^2, 3], got {tuple(self._features.shape)}")
        if not 0 <= sh_degree <= min(self.max_sh_degree, 3):
            raise ValueError(f"active sh_degree {sh_degree} outside 0..min(stored degree {self.max_sh_degree}, 3) "
                             "(eval_sh on the pose path goes up to degree 3, sampling.py:116-124)")
        self.active_sh_degree = sh_degree

    @classmethod
    def from_dict(cls, d, device=None):
        return cls(d["xyz"], d["scaling"], d["rotation"], d["features_dc"], d["features_rest"],
                   d.get("sh_degree", 3), device=device)

    @classmethod
    def from_gaussian_model(cls, gm, device=None):
        """Adopt a reference GaussianModel (attribute names of scene/gaussian_model.py:62-69)."""
        return cls(gm._xyz, gm._scaling, gm._rotation, gm._features_dc, gm._features_rest,
                   gm.active_sh_degree, device=device)

    @classmethod
    def load_ply(cls, path: str, device="cuda"):
        """3DGS point_cloud.ply (binary little endian, all float32 properties): x,y,z,nx,ny,nz,
        f_dc_0..2, f_rest_0..44 (channel-major), opacity, scale_0..2, rot_0..3
        (reference scene/gaussian_model.py:284-296,342-420)."""
        with open(path, "rb") as fh:
            names = []
            n = None
            while True:
                line = fh.readline().decode("ascii").strip()
                if line.startswith("format") and "binary_little_endian" not in line:
                    raise ValueError("only binary_little_endian PLY is supported")
                if line.startswith("element vertex"):
                    n = int(line.split()[-1])
                elif line.startswith("property"):
                    _, typ, name = line.split()
                    if typ not in ("float", "float32"):
                        raise ValueError(f"unsupported PLY property type {typ}")
                    names.append(name)
                elif line == "end_header":
                    break
            data = np.frombuffer(fh.read(n * 4 * len(names)), dtype="<f4").reshape(n, len(names))
        col = {k: i for i, k in enumerate(names)}
        take = lambda keys: torch.from_numpy(np.ascontiguousarray(data[:, [col[k] for k in keys]]))  # noqa: E731
        n_rest = len([k for k in names if k.startswith("f_rest_")])
        deg = int(round((n_rest // 3 + 1) ** 0.5)) - 1
        rest = take([f"f_rest_{i}" for i in range(n_rest)]).reshape(n, 3, n_rest // 3).transpose(1, 2)
        dc = take(["f_dc_0", "f_dc_1", "f_dc_2"]).reshape(n, 3, 1).transpose(1, 2)
        nsc = len([k for k in names if k.startswith("scale_")])
        nrot = len([k for k in names if k.startswith("rot")])
        return cls(take(["x", "y", "z"]), take([f"scale_{i}" for i in range(nsc)]),
                   take([f"rot_{i}" for i in range(nrot)]), dc, rest, deg, device=device)

    # --- getters with the reference names -------------------------------------------------------
    @property
    def get_xyz(self):
        return self._xyz

    @property
    def get_scaling(self):
        return torch.exp(self._scaling)

    @property
    def get_rotation(self):
        return torch.nn.functional.normalize(self._rotation)

    def get_rotation_mat(self):
        """[N,3,3] rotation of every Gaussian (reference gaussian_model.py:133-134 -> general_utils.py:103-126): the
        stored quaternion (w,x,y,z) is normalised by ``get_rotation`` and once more by the matrix builder, in that
        order -- the ray-generation kernel does the same two normalisations in registers (csrc/raygen.cu) and never
        calls this; it is here for callers of the reference getter."""
        q = self.get_rotation
        q = q / torch.sqrt(q[:, 0] * q[:, 0] + q[:, 1] * q[:, 1] + q[:, 2] * q[:, 2] + q[:, 3] * q[:, 3])[:, None]
        w, x, y, z = q.unbind(-1)
        rows = (1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
                2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
                2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y))
        return torch.stack(rows, dim=-1).reshape(-1, 3, 3)

    @property
    def get_features(self):
        return self._features

    @property
    def num_gaussians(self) -> int:
        return self._xyz.shape[0]

    @property
    def device(self):
        return self._xyz.device
