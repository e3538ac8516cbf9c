"""CPU oracle for the SipMask inference hot path.

TEST INFRASTRUCTURE ONLY.  This package restates, in plain PyTorch / numpy / C,
the algorithm of the reference (JialeCao001/SipMask @ bc63fa9,
`SipMask-mmdetection/`) for the path ResNet -> FPN -> SipMaskHead.forward ->
get_bboxes (decode, NMS, mask assembly).  Only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` may import it; the product package `sipmask_b200` never does.

Parity pin: the restatement is checked against
  * the reference's own known-answer NMS vectors
    (MM/tests/test_nms.py:17-41, MM/mmdet/ops/nms/nms_wrapper.py:25-34,
     BM/tests/test_nms.py:16-58), in tests/test_oracle_nms.py;
  * outputs of the *unmodified reference python* (`SipMaskHead.forward`,
    `get_bboxes`, `ResNet`, `FPN`, `multiclass_nms_idx`, `fast_nms`) imported in
    the build container through tests/golden/_ref_import.py, committed as
    fixtures under tests/golden/*.npz (generator: tests/golden/gen_golden.py);
  * the reference's `nms_cpu.cpp`, compiled from where it lies into
    oracle/_ref/ (recipe: oracle/build.py).
The two CUDA-only reference kernels (deformable im2col, CropSplit) have no
CPU implementation and no reference test vectors; for them the oracle's
restatement of the .cu source is the pin ("parity unpinned" by reference
vectors; cross-checked against torchvision.ops.deform_conv2d for DCN).
"""
