"""Mirror of the parts of `models/decoders/decode_head.py` that are on the hot path: the constructor
(:55-109: `conv_seg`, `Dropout2d`, input selection), `_transform_inputs` (:156-181, 'multiple_select') and
`cls_seg` (:226-231).  The mmseg loss zoo / `losses()` of the reference is dead code on this path."""
import torch.nn as nn


class BaseDecodeHead(nn.Module):
    def __init__(self, in_channels, channels, *, num_classes, dropout_ratio=0.1, conv_cfg=None, norm_cfg=None,
                 act_cfg=dict(type="ReLU"), in_index=-1, input_transform=None, loss_decode=None, ignore_index=255,
                 sampler=None, align_corners=False, init_cfg=None, **kwargs):
        super().__init__()
        self._init_inputs(in_channels, in_index, input_transform)
        self.channels, self.num_classes, self.dropout_ratio = channels, num_classes, dropout_ratio
        self.conv_cfg, self.norm_cfg, self.act_cfg = conv_cfg, norm_cfg, act_cfg
        self.in_index, self.ignore_index, self.align_corners = in_index, ignore_index, align_corners
        self.init_cfg = init_cfg
        self.conv_seg = nn.Conv2d(channels, num_classes, kernel_size=1)
        self.dropout = nn.Dropout2d(dropout_ratio) if dropout_ratio > 0 else None
        self.fp16_enabled = False

    def _init_inputs(self, in_channels, in_index, input_transform):
        if input_transform is not None:
            assert input_transform in ("resize_concat", "multiple_select")
            assert isinstance(in_channels, (list, tuple)) and isinstance(in_index, (list, tuple))
            assert len(in_channels) == len(in_index)
        else:
            assert isinstance(in_channels, int) and isinstance(in_index, int)
        self.input_transform, self.in_channels = input_transform, in_channels

    def _transform_inputs(self, inputs):
        if self.input_transform == "multiple_select":
            return [inputs[i] for i in self.in_index]
        if self.input_transform is None:
            return inputs[self.in_index]
        raise NotImplementedError("'resize_concat' is not used by LightHamHead")
