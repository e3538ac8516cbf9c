"""CPU-side checks of the C ABI: the library builds, loads, and exports every symbol include/sipmask_b200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from sipmask_b200 import _lib, build
    path = build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    header = open(os.path.join(ROOT, 'include', 'sipmask_b200.h')).read()
    declared = sorted(set(re.findall(r'\b(smb_[a-z0-9_]+)\s*\(', header)))
    assert declared, 'no declarations parsed'
    for name in declared:
        assert hasattr(lib, name), 'missing export %s' % name
    assert sorted(declared) == sorted(_lib.SYMBOLS)
    assert lib.smb_version() >= 100


def test_ops_fail_loudly_without_cuda():
    import torch
    from sipmask_b200 import ops, _lib
    if torch.cuda.is_available():
        pytest.skip('CUDA present')
    with pytest.raises(_lib.SmbError):
        ops.crop_split(torch.zeros(4, 4, 4, 1), torch.zeros(1, 4))
    with pytest.raises(_lib.SmbError):
        ops.nms(torch.zeros(3, 5), 0.5)
    with pytest.raises(_lib.SmbError):
        ops.mask_assemble(torch.zeros(32, 4, 4), torch.zeros(1, 128), torch.zeros(1, 4), 0.5)


def test_sass_contains_blackwell_instructions():
    """The conv kernel must be real tcgen05/TMA code (B200_PROFILING.md 'What proves a Blackwell-native kernel')."""
    import shutil
    import subprocess
    from sipmask_b200 import build
    if shutil.which('cuobjdump') is None:
        pytest.skip('cuobjdump not available')
    sass = subprocess.run(['cuobjdump', '-sass', build.LIB], capture_output=True, text=True).stdout
    for mnemonic in ('UTCHMMA', 'UTMALDG', 'LDTM'):
        assert mnemonic in sass, mnemonic
