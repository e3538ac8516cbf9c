"""Oracle restatement of the reference network forward (CPU, fp32, pure PyTorch).

TEST INFRASTRUCTURE ONLY - see oracle/__init__.py.

Module / parameter names equal the reference's so that one `state_dict` loads
into the reference modules, this oracle and the product engine alike
(SURVEY.md §8b "state_dict contract").  Citations are relative to
/root/reference/SipMask-mmdetection/ (MM/).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops as O


# ---------------------------------------------------------------- backbone ---
class Bottleneck(nn.Module):
    """MM/mmdet/models/backbones/resnet.py:84-239, style='caffe' (stride on conv1, :125-130)."""
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=None, dcn=False):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 1, stride=stride, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.with_dcn = dcn
        if dcn:
            # DeformConvPack, deformable_groups=1 (MM/mmdet/ops/dcn/deform_conv.py:258-296)
            self.conv2 = DeformConvPack(planes, planes)
        else:
            self.conv2 = nn.Conv2d(planes, planes, 3, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv3 = nn.Conv2d(planes, planes * 4, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * 4)
        self.downsample = downsample

    def forward(self, x):
        identity = x
        out = F.relu(self.bn1(self.conv1(x)))
        out = F.relu(self.bn2(self.conv2(out)))
        out = self.bn3(self.conv3(out))
        if self.downsample is not None:
            identity = self.downsample(x)
        return F.relu(out + identity)


class DeformConvPack(nn.Module):
    """MM/mmdet/ops/dcn/deform_conv.py:258-296: offset conv 3x3 (bias) + DCN, dg=1."""

    def __init__(self, cin, cout):
        super().__init__()
        self.weight = nn.Parameter(torch.zeros(cout, cin, 3, 3))
        self.conv_offset = nn.Conv2d(cin, 18, 3, padding=1, bias=True)

    def forward(self, x):
        return O.deform_conv(x, self.conv_offset(x), self.weight, 1, 1, 1, 1)


class ResNet(nn.Module):
    """MM/mmdet/models/backbones/resnet.py:312-521 (caffe style, eval-mode BN)."""
    arch = {50: (3, 4, 6, 3), 101: (3, 4, 23, 3)}

    def __init__(self, depth=50, stage_with_dcn=(False, False, False, False)):
        super().__init__()
        self.conv1 = nn.Conv2d(3, 64, 7, stride=2, padding=3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        inplanes = 64
        for i, nb in enumerate(self.arch[depth]):
            planes = 64 * 2 ** i
            stride = 1 if i == 0 else 2
            down = nn.Sequential(nn.Conv2d(inplanes, planes * 4, 1, stride=stride, bias=False),
                                 nn.BatchNorm2d(planes * 4))
            # make_res_layer: block 0 always gets `dcn`, later blocks only when i % 3 == 0 (:288-291)
            layers = [Bottleneck(inplanes, planes, stride, down, dcn=stage_with_dcn[i])]
            inplanes = planes * 4
            for j in range(1, nb):
                layers.append(Bottleneck(inplanes, planes, dcn=stage_with_dcn[i] and j % 3 == 0))
            setattr(self, 'layer%d' % (i + 1), nn.Sequential(*layers))
        self.eval()

    def forward(self, x):
        x = F.relu(self.bn1(self.conv1(x)))
        x = F.max_pool2d(x, 3, 2, 1)
        outs = []
        for i in range(4):
            x = getattr(self, 'layer%d' % (i + 1))(x)
            outs.append(x)
        return tuple(outs)


# -------------------------------------------------------------------- neck ---
class _Conv(nn.Module):
    """ConvModule without norm/act: parameter path `.conv.{weight,bias}` (conv_module.py:68-77)."""

    def __init__(self, cin, cout, k, stride=1, padding=0, bias=True):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, k, stride=stride, padding=padding, bias=bias)

    def forward(self, x):
        return self.conv(x)


class FPN(nn.Module):
    """MM/mmdet/models/necks/fpn.py:50-178 with the sipmask config
    (start_level=1, add_extra_convs, extra_convs_on_inputs=False, relu_before_extra_convs)."""

    def __init__(self, in_channels=(256, 512, 1024, 2048), out_channels=256, num_outs=5, start_level=1):
        super().__init__()
        self.start_level = start_level
        self.num_outs = num_outs
        self.lateral_convs = nn.ModuleList()
        self.fpn_convs = nn.ModuleList()
        for i in range(start_level, len(in_channels)):
            self.lateral_convs.append(_Conv(in_channels[i], out_channels, 1))
            self.fpn_convs.append(_Conv(out_channels, out_channels, 3, padding=1))
        for i in range(num_outs - (len(in_channels) - start_level)):
            self.fpn_convs.append(_Conv(out_channels, out_channels, 3, stride=2, padding=1))

    def forward(self, inputs):
        laterals = [l(inputs[i + self.start_level]) for i, l in enumerate(self.lateral_convs)]
        n = len(laterals)
        for i in range(n - 1, 0, -1):
            laterals[i - 1] = laterals[i - 1] + F.interpolate(
                laterals[i], size=laterals[i - 1].shape[2:], mode='nearest')       # fpn.py:149-152
        outs = [self.fpn_convs[i](laterals[i]) for i in range(n)]
        outs.append(self.fpn_convs[n](outs[-1]))                                  # fpn.py:171-172
        for i in range(n + 1, self.num_outs):
            outs.append(self.fpn_convs[i](F.relu(outs[-1])))                      # fpn.py:174-175
        return tuple(outs)


# -------------------------------------------------------------------- head ---
class ConvModule(nn.Module):
    """conv -> GN -> ReLU (MM/mmdet/ops/conv_module.py:124-132); norm attr name `gn` (norm.py:5-10)."""

    def __init__(self, cin, cout, k=3, stride=1, padding=1, gn=True, bias=None):
        super().__init__()
        if bias is None:
            bias = not gn
        self.conv = nn.Conv2d(cin, cout, k, stride=stride, padding=padding, bias=bias)
        if gn:
            self.gn = nn.GroupNorm(32, cout, eps=1e-5)
        self.with_gn = gn

    def forward(self, x):
        x = self.conv(x)
        if self.with_gn:
            x = self.gn(x)
        return F.relu(x)


class Scale(nn.Module):
    def __init__(self, scale=1.0):
        super().__init__()
        self.scale = nn.Parameter(torch.tensor(scale, dtype=torch.float))

    def forward(self, x):
        return x * self.scale


class _DeformConvW(nn.Module):
    def __init__(self, cin, cout, dg):
        super().__init__()
        self.weight = nn.Parameter(torch.zeros(cout, cin, 3, 3))
        self.dg = dg

    def forward(self, x, offset):
        return O.deform_conv(x, offset, self.weight, 1, 1, 1, self.dg)


class FeatureAlign(nn.Module):
    """MM/mmdet/models/anchor_heads/sipmask_head.py:21-55."""

    def __init__(self, cin, cout, deformable_groups=4, flag_norm=True):
        super().__init__()
        self.conv_offset = nn.Conv2d(4, deformable_groups * 18, 1, bias=False)
        self.conv_adaption = _DeformConvW(cin, cout, deformable_groups)
        self.norm = nn.GroupNorm(32, cin)
        self.flag_norm = flag_norm

    def forward(self, x, shape):
        offset = self.conv_offset(shape)
        x = self.conv_adaption(x, offset)
        if self.flag_norm:
            x = self.norm(x)
        return F.relu(x)


class SipMaskHead(nn.Module):
    """MM/mmdet/models/anchor_heads/sipmask_head.py:107-287 (forward only)."""

    def __init__(self, num_classes=81, in_channels=256, feat_channels=256, stacked_convs=4,
                 strides=(8, 16, 32, 64, 128), ssd_flag=False, rescoring_flag=False, gn=True):
        super().__init__()
        self.num_classes = num_classes
        self.cls_out_channels = num_classes - 1
        self.strides = strides
        self.ssd_flag = ssd_flag
        self.rescoring_flag = rescoring_flag
        self.cls_convs = nn.ModuleList()
        self.reg_convs = nn.ModuleList()
        for i in range(stacked_convs - 1):
            self.cls_convs.append(ConvModule(in_channels if i == 0 else feat_channels, feat_channels, gn=gn))
        for i in range(stacked_convs):
            self.reg_convs.append(ConvModule(in_channels if i == 0 else feat_channels, feat_channels, gn=gn))
        self.fcos_cls = nn.Conv2d(feat_channels, self.cls_out_channels, 3, padding=1)
        self.fcos_reg = nn.Conv2d(feat_channels, 4, 3, padding=1)
        self.fcos_centerness = nn.Conv2d(feat_channels, 1, 3, padding=1)
        self.scales = nn.ModuleList([Scale(1.0) for _ in strides])
        self.nc = 32
        self.feat_align = FeatureAlign(feat_channels, feat_channels, 4, flag_norm=gn)
        self.sip_cof = nn.Conv2d(feat_channels, self.nc * 4, 3, padding=1)
        self.sip_mask_lat = nn.Conv2d(512, self.nc, 3, padding=1)
        self.sip_mask_lat0 = nn.Conv2d(768, 512, 1, padding=0)
        if rescoring_flag:
            ch = [1, 16, 16, 16, 32, 64, 128]
            self.convs_scoring = nn.Sequential(*[
                ConvModule(ch[i], ch[i + 1], 3, stride=2, padding=0, gn=False, bias=True) for i in range(6)])
            self.mask_scoring = nn.Conv2d(128, num_classes - 1, 1)

    def forward(self, feats):
        cls_scores, bbox_preds, centernesses, cof_preds, feat_masks = [], [], [], [], []
        for count, (x, scale, stride) in enumerate(zip(feats, self.scales, self.strides)):
            cls_feat = x
            reg_feat = x
            for l in self.cls_convs:
                cls_feat = l(cls_feat)
            for l in self.reg_convs:
                reg_feat = l(reg_feat)
            bbox_pred = scale(self.fcos_reg(reg_feat))                    # no exp / relu (:261)
            cls_feat = self.feat_align(cls_feat, bbox_pred)               # (:263)
            cls_scores.append(self.fcos_cls(cls_feat))
            centernesses.append(self.fcos_centerness(reg_feat))           # on reg feature (:265)
            bbox_preds.append(bbox_pred.float() * stride)                 # (:268)
            cof_preds.append(self.sip_cof(cls_feat))                      # (:271)
            if count < 3:
                if count == 0:
                    feat_masks.append(reg_feat)
                else:
                    feat_masks.append(F.interpolate(reg_feat, scale_factor=(2 ** count),
                                                    mode='bilinear', align_corners=False))
        fm = torch.cat(feat_masks, dim=1)
        fm = F.relu(self.sip_mask_lat(F.relu(self.sip_mask_lat0(fm))))
        fm = F.interpolate(fm, scale_factor=4, mode='bilinear', align_corners=False)
        return cls_scores, bbox_preds, centernesses, cof_preds, fm


class SipMaskDetector(nn.Module):
    """backbone + neck + head (MM/mmdet/models/detectors/single_stage.py:44-49,75-93)."""

    def __init__(self, depth=50, stacked_convs=4, gn=True, ssd_flag=False, rescoring_flag=False,
                 num_classes=81, backbone_dcn=False):
        super().__init__()
        self.backbone = ResNet(depth, (False, backbone_dcn, backbone_dcn, backbone_dcn))
        self.neck = FPN()
        self.bbox_head = SipMaskHead(num_classes=num_classes, stacked_convs=stacked_convs, gn=gn,
                                     ssd_flag=ssd_flag, rescoring_flag=rescoring_flag)
        self.eval()

    def extract_feat(self, img):
        return self.neck(self.backbone(img))

    def forward(self, img):
        return self.bbox_head(self.extract_feat(img))


class SipMaskVISHead(SipMaskHead):
    """SipMask-VIS head (VIS/mmdet/models/anchor_heads/sipmask_head.py:160-317; paths below relative to
    /root/reference/SipMask-VIS/mmdet/): the image head plus a tracking branch - `track_convs` (stacked_convs - 1
    ConvModules, :274-286) applied to levels 0..2, bilinearly upsampled to P3 resolution (:306-309; level 0 is "interpolated"
    by 1), concatenated and reduced by `sipmask_track` 1x1 768 -> 512 (:287,:310-312).  Every level's reg feature is
    interpolated for the prototype branch, also level 0 with factor 1 (:301-303).  forward(feats, feats_x, flag_train) returns
    the 7-tuple of :315-317; at test time track_feats_ref is track_feats."""

    def __init__(self, num_classes=41, in_channels=256, feat_channels=256, stacked_convs=3, strides=(8, 16, 32, 64, 128), gn=True):
        super().__init__(num_classes=num_classes, in_channels=in_channels, feat_channels=feat_channels,
                         stacked_convs=stacked_convs, strides=strides, ssd_flag=False, rescoring_flag=False, gn=gn)
        self.track_convs = nn.ModuleList([ConvModule(in_channels if i == 0 else feat_channels, feat_channels, gn=gn)
                                          for i in range(stacked_convs - 1)])
        self.sipmask_track = nn.Conv2d(feat_channels * 3, 512, 1, padding=0)

    def forward(self, feats, feats_x=None, flag_train=False):
        cls_scores, bbox_preds, centernesses, cof_preds, feat_masks, track_feats = [], [], [], [], [], []
        for count, (x, scale, stride) in enumerate(zip(feats, self.scales, self.strides)):
            cls_feat = reg_feat = track_feat = x
            for l in self.cls_convs:
                cls_feat = l(cls_feat)
            for l in self.reg_convs:
                reg_feat = l(reg_feat)
            if count < 3:
                for l in self.track_convs:
                    track_feat = l(track_feat)
                track_feats.append(F.interpolate(track_feat, scale_factor=(2 ** count), mode='bilinear', align_corners=False))
            bbox_pred = scale(self.fcos_reg(reg_feat))
            cls_feat = self.feat_align(cls_feat, bbox_pred)
            cls_scores.append(self.fcos_cls(cls_feat))
            centernesses.append(self.fcos_centerness(reg_feat))
            bbox_preds.append(bbox_pred.float() * stride)
            cof_preds.append(self.sip_cof(cls_feat))
            if count < 3:
                feat_masks.append(F.interpolate(reg_feat, scale_factor=(2 ** count), mode='bilinear', align_corners=False))
        fm = torch.cat(feat_masks, dim=1)
        fm = F.relu(self.sip_mask_lat(F.relu(self.sip_mask_lat0(fm))))
        fm = F.interpolate(fm, scale_factor=4, mode='bilinear', align_corners=False)
        tf = self.sipmask_track(torch.cat(track_feats, dim=1))
        return cls_scores, bbox_preds, centernesses, cof_preds, fm, tf, tf
