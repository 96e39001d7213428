"""Camera-up head (boundary only -- stays PyTorch/cuDNN): three valid 5x5 convs
(16->12->8->4), one 4x4 conv (4->1), then 384->256->3.  Parameter names match the reference
``CameraDirectionPredictor`` (pose_estimation/camera_direction_network.py:7-89) so ``id_module.th``
checkpoints load unchanged."""
import torch


def _reducer(channels: int, kernel: int, count: int) -> torch.nn.Sequential:
    layers = []
    for _ in range(count):
        layers += [torch.nn.Conv2d(channels, channels, kernel_size=kernel), torch.nn.ReLU(inplace=True)]
    return torch.nn.Sequential(*layers)


class CameraDirectionPredictor(torch.nn.Module):
    def __init__(self, image_feature_channel: int = 256, image_size=(16, 16), pospe: int = 8, featureC: int = 256,
                 fea_output: int = 3):
        """same positional order and defaults as the reference (camera_direction_network.py:8-15); ``pospe`` only
        sets two attributes there (``direction_input``, ``pospe``) that no forward path reads"""
        super().__init__()
        self.pospe = pospe
        self.direction_input = 2 * pospe * 3 + 3
        self.dim_reducer1 = _reducer(image_feature_channel, 5, 3)
        self.dim_reducer2 = _reducer(image_feature_channel, 4, 1)
        side = [s - 3 * 4 - 3 for s in image_size]
        self.in_mlpC = side[0] * side[1] * image_feature_channel
        self.mlp = torch.nn.Sequential(torch.nn.Linear(self.in_mlpC, featureC), torch.nn.ReLU(inplace=True),
                                       torch.nn.Linear(featureC, fea_output))

    @staticmethod
    def _conv_gemm(x: torch.Tensor, conv: torch.nn.Conv2d) -> torch.Tensor:
        """valid convolution as im2col + one fp32 GEMM.  cuDNN's heuristics pick Winograd/FFT style
        algorithms for these 384-channel 5x5 layers whose results differ from the fp32 direct sum by ~2e-4
        relative -- more than the 1e-4 pose tolerance once it reaches the rotation; a plain GEMM does not."""
        k = conv.kernel_size[0]
        cols = torch.nn.functional.unfold(x, k)  # [B, C*k*k, L]
        out = conv.weight.view(conv.out_channels, -1) @ cols + conv.bias[:, None]
        side = x.shape[-1] - k + 1
        return out.view(x.shape[0], conv.out_channels, side, side)

    def forward_batch(self, image_features: torch.Tensor) -> torch.Tensor:
        """[B,384,16,16] -> [B,3]"""
        y = image_features
        for conv in (self.dim_reducer1[0], self.dim_reducer1[2], self.dim_reducer1[4], self.dim_reducer2[0]):
            y = torch.relu(self._conv_gemm(y, conv))
        return self.mlp(y.reshape(y.shape[0], -1))

    def forward(self, image_features: torch.Tensor) -> torch.Tensor:
        return self.forward_batch(image_features[None])[0]
