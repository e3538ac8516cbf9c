"""Import shim that lets the *unmodified* reference python (SipMask-mmdetection)
be imported in the build container, where mmcv / pycocotools / the compiled
mmdet.ops extensions are absent.

Test infrastructure only (used by tests/golden/gen_golden.py to produce the
committed fixtures).  Nothing here is shipped or imported by the product.

What is stubbed (no arithmetic on the hot path lives in any of these):
  * mmcv            -> weight-init helpers restated from their documented
                       behaviour, dummy base classes (VGG, Hook, ...)
  * pycocotools, terminaltables, matplotlib, six -> inert mocks
  * compiled extension modules (deform_conv_cuda, crop_split_cuda, nms_cuda...)
                    -> inert mocks; nms_cpu is replaced by the reference's own
                       nms_cpu.cpp compiled into oracle/_ref when available.
"""
import importlib.abc
import importlib.machinery
import sys
import types
from unittest import mock

import torch
import torch.nn as nn

REF_MM = '/root/reference/SipMask-mmdetection'

_MOCK_ROOTS = ('pycocotools', 'terminaltables', 'matplotlib', 'six', 'imagecorruptions', 'albumentations')
_COMPILED = ('deform_conv_cuda', 'deform_pool_cuda', 'nms_cpu', 'nms_cuda', 'soft_nms_cpu',
             'crop_split_cuda', 'crop_split_gt_cuda', 'roi_align_cuda',
             'roi_pool_cuda', 'sigmoid_focal_loss_cuda', 'masked_conv2d_cuda',
             'carafe_cuda', 'carafe_naive_cuda', 'grid_sampler_cuda',
             'affine_grid_cuda', 'compiling_info')


class _MockLoader(importlib.abc.Loader):
    def create_module(self, spec):
        m = mock.MagicMock(name=spec.name)
        m.__name__ = spec.name
        m.__path__ = []
        m.__spec__ = spec
        m.__loader__ = self
        return m

    def exec_module(self, module):
        pass


class _MockFinder(importlib.abc.MetaPathFinder):
    def find_spec(self, name, path, target=None):
        root = name.split('.')[0]
        leaf = name.split('.')[-1]
        if root in _MOCK_ROOTS or (root == 'mmdet' and leaf in _COMPILED):
            return importlib.machinery.ModuleSpec(name, _MockLoader(), is_package=True)
        return None


def _constant_init(module, val, bias=0):
    if hasattr(module, 'weight') and module.weight is not None:
        nn.init.constant_(module.weight, val)
    if hasattr(module, 'bias') and module.bias is not None:
        nn.init.constant_(module.bias, bias)


def _xavier_init(module, gain=1, bias=0, distribution='normal'):
    if distribution == 'uniform':
        nn.init.xavier_uniform_(module.weight, gain=gain)
    else:
        nn.init.xavier_normal_(module.weight, gain=gain)
    if hasattr(module, 'bias') and module.bias is not None:
        nn.init.constant_(module.bias, bias)


def _normal_init(module, mean=0, std=1, bias=0):
    nn.init.normal_(module.weight, mean, std)
    if hasattr(module, 'bias') and module.bias is not None:
        nn.init.constant_(module.bias, bias)


def _kaiming_init(module, a=0, mode='fan_out', nonlinearity='relu', bias=0,
                  distribution='normal'):
    if distribution == 'uniform':
        nn.init.kaiming_uniform_(module.weight, a=a, mode=mode, nonlinearity=nonlinearity)
    else:
        nn.init.kaiming_normal_(module.weight, a=a, mode=mode, nonlinearity=nonlinearity)
    if hasattr(module, 'bias') and module.bias is not None:
        nn.init.constant_(module.bias, bias)


def _caffe2_xavier_init(module, bias=0):
    _kaiming_init(module, a=1, mode='fan_in', nonlinearity='leaky_relu',
                  distribution='uniform')


class _Dummy(object):
    def __init__(self, *a, **k):
        pass


class _DummyVGG(nn.Module):
    def __init__(self, *a, **k):
        super().__init__()


def _make_fake_mmcv():
    mmcv = types.ModuleType('mmcv')
    mmcv.__path__ = []
    mmcv.is_str = lambda x: isinstance(x, str)
    mmcv.is_list_of = lambda seq, t: isinstance(seq, list) and all(isinstance(s, t) for s in seq)
    mmcv.__getattr__ = lambda name: mock.MagicMock(name='mmcv.' + name)

    cnn = types.ModuleType('mmcv.cnn')
    cnn.__path__ = []
    wi = types.ModuleType('mmcv.cnn.weight_init')
    for mod in (cnn, wi):
        mod.constant_init = _constant_init
        mod.xavier_init = _xavier_init
        mod.normal_init = _normal_init
        mod.kaiming_init = _kaiming_init
        mod.caffe2_xavier_init = _caffe2_xavier_init
    cnn.VGG = _DummyVGG
    cnn.weight_init = wi

    runner = types.ModuleType('mmcv.runner')
    runner.load_checkpoint = mock.MagicMock(name='load_checkpoint')
    runner.get_dist_info = lambda: (0, 1)
    for n in ('OptimizerHook', 'Hook', 'DistSamplerSeedHook', 'Runner'):
        setattr(runner, n, type(n, (_Dummy,), {}))

    runner.__path__ = []
    ru = types.ModuleType('mmcv.runner.utils')          # VIS tree: `from mmcv.runner.utils import get_dist_info`
    ru.get_dist_info = runner.get_dist_info
    runner.utils = ru
    sys.modules['mmcv.runner.utils'] = ru

    parallel = types.ModuleType('mmcv.parallel')
    for n in ('MMDataParallel', 'MMDistributedDataParallel', 'DataContainer'):
        setattr(parallel, n, type(n, (_Dummy,), {}))
    parallel.collate = mock.MagicMock(name='collate')
    parallel.scatter = mock.MagicMock(name='scatter')

    mmcv.cnn, mmcv.runner, mmcv.parallel = cnn, runner, parallel
    sys.modules.update({'mmcv': mmcv, 'mmcv.cnn': cnn, 'mmcv.cnn.weight_init': wi,
                        'mmcv.runner': runner, 'mmcv.parallel': parallel})


_installed = False


def install(ref=REF_MM):
    """Make `import mmdet` resolve to the unmodified reference tree (`ref`: SipMask-mmdetection by default, or
    /root/reference/SipMask-VIS for the video head; one tree per process)."""
    global _installed
    if _installed:
        return
    _make_fake_mmcv()
    sys.meta_path.insert(0, _MockFinder())
    sys.path.insert(0, ref)
    _installed = True
