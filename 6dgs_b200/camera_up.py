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
    def __init__(self, image_feature_channel: int = 384, image_size=(16, 16), featureC: int = 256, fea_output: int = 3):
        super().__init__()
        self.dim_reducer1 = _reducer(image_feature_channel, 5, 3)
        self.dim_reducer2 = _reducer(image_feature_channel, 4, 1)
        side = [s - 3 * 4 - 3 for s in image_size]
        self.in_mlpC = side[0] * side[1] * image_feature_channel
        self.mlp = torch.nn.Sequential(torch.nn.Linear(self.in_mlpC, featureC), torch.nn.ReLU(inplace=True),
                                       torch.nn.Linear(featureC, fea_output))

    def forward(self, image_features: torch.Tensor) -> torch.Tensor:
        y = self.dim_reducer2(self.dim_reducer1(image_features[None]))
        return self.mlp(y.view(y.shape[0], -1))[0]
