"""SimplePoseNet as the reference defines it (bpc/pose/models/simple_pose_net.py:7-38): a torchvision ResNet50
backbone + Linear(2048, {3, 4, 6}).  A third-party dense network, outside the hot path; provided so that
checkpoints trained with the reference load unchanged and the batched crop tensor can be fed straight in."""
from __future__ import annotations

import torch
import torch.nn as nn

OUT_DIM = {'euler': 3, 'quat': 4, '6d': 6}


class SimplePoseNet(nn.Module):
    def __init__(self, loss_type="euler", pretrained=True):
        super().__init__()
        import torchvision.models as tv_models
        if loss_type not in OUT_DIM:
            raise ValueError("loss_type must be one of 'euler', 'quat', or '6d'")
        backbone = tv_models.resnet50(weights=(tv_models.ResNet50_Weights.IMAGENET1K_V2 if pretrained else None))
        self.backbone = nn.Sequential(*list(backbone.children())[:-1])
        self.fc = nn.Linear(2048, OUT_DIM[loss_type])

    def forward(self, x):
        feats = self.backbone(x)
        return self.fc(feats.view(feats.size(0), -1))


def load_pose_model(pose_model_path, device='cuda:0', rotation_mode=None):
    """Checkpoint loader with rotation-mode auto-detection -- reference process_pose.py:44-74."""
    checkpoint = torch.load(pose_model_path, map_location=device)
    if "fc.weight" not in checkpoint:
        raise KeyError("The checkpoint does not contain 'fc.weight'.")
    output_dim = checkpoint["fc.weight"].shape[0]
    if rotation_mode is None:
        rotation_mode = {3: 'euler', 4: 'quat', 6: '6d'}.get(output_dim)
        if rotation_mode is None:
            raise ValueError(f"Unexpected output dimension: {output_dim}. Cannot determine rotation mode.")
    pose_model = SimplePoseNet(loss_type=rotation_mode, pretrained=False)
    pose_model.load_state_dict(checkpoint)
    pose_model.to(device).eval()
    return pose_model, rotation_mode
