"""Build recipe for the oracle's native pieces.  TEST INFRASTRUCTURE ONLY.

  * oracle/_ref/sipmask_ref_nms_cpu*.so  - the reference's own CPU NMS, compiled from where
    it lies (/root/reference/SipMask-mmdetection/mmdet/ops/nms/src/nms_cpu.cpp) with
    -DAT_CHECK=TORCH_CHECK (the macro was removed from torch).  Only built when
    /root/reference exists (the build container); the GPU box uses the prebuilt file.
  * oracle/_ref/liboracle_c.so           - plain-C restatement (oracle/csrc/oracle_c.c) of the
    NMS / CropSplit / mask-assembly arithmetic, used by tests for full-size cases.

Nothing from /root/reference is copied into the repository; outputs go to oracle/_ref/ only
(git-ignored, not gpurun-ignored).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_NMS = '/root/reference/SipMask-mmdetection/mmdet/ops/nms/src/nms_cpu.cpp'
OUT = os.path.join(HERE, '_ref')


def build_c():
    os.makedirs(OUT, exist_ok=True)
    src = os.path.join(HERE, 'csrc', 'oracle_c.c')
    dst = os.path.join(OUT, 'liboracle_c.so')
    if os.path.exists(dst) and os.path.getmtime(dst) >= os.path.getmtime(src):
        return dst
    subprocess.check_call(['gcc', '-O2', '-fPIC', '-shared', '-std=c99', '-ffp-contract=off',
                           '-o', dst, src, '-lm'])
    return dst


def build_ref_nms():
    os.makedirs(OUT, exist_ok=True)
    import glob
    have = glob.glob(os.path.join(OUT, 'sipmask_ref_nms_cpu*.so'))
    if have:
        return have[0]
    if not os.path.exists(REF_NMS):
        return None
    from torch.utils.cpp_extension import load
    load(name='sipmask_ref_nms_cpu', sources=[REF_NMS], extra_cflags=['-O2', '-DAT_CHECK=TORCH_CHECK', '-w'],
         build_directory=OUT, verbose=False)
    have = glob.glob(os.path.join(OUT, 'sipmask_ref_nms_cpu*.so'))
    return have[0] if have else None


def build_all():
    return build_c(), build_ref_nms()


if __name__ == '__main__':
    print(build_all())
