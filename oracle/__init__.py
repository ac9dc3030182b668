"""CPU oracle for the gaze+EEG fusion hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this package.  Nothing under ``eyegaze_multimodal_b200/`` imports it, and the product
path raises when its CUDA library is missing instead of falling back to anything here.

Contents
--------
``eeg.py``     functional fp32 PyTorch-CPU restatement of ``3_Models/backbones/dual_eeg_transformer.py``
               and the encoder pieces of ``3_Models/backbones/art.py`` (state_dict in, tensors out).
``fuzzy.py``   restatement of ``3_Models/fusion/fuzzy_gating_fusion.py``.
``vit.py``     restatement of the timm VisionTransformer arithmetic the gaze encoders call
               (timm is an un-vendored, un-pinned dependency of the reference; see the file header)
               plus the wrapper logic of ``early_fusion_vit.py`` / ``late_fusion_vit.py``.
``reference_loader.py``  imports the UNMODIFIED reference modules by file path, exactly as the
               reference's own training scripts do; usable only where ``/root/reference`` exists.
``make_golden.py``       runs the unmodified reference and writes ``tests/golden/*.npz``.

Pinning status
--------------
EEG branch and fuzzy fusion: PINNED -- the restatement is checked (tests/test_oracle_pinned.py) against
golden vectors produced by the unmodified reference in this container, and against the reference's own
self-test known answers (alpha = 0.5 / 0.7907 / 0.2102, T_img = 1.5, T_eeg = 1.0).
ViT gaze branch: PARITY UNPINNED by the reference (timm is absent, version unpinned, its only tests need
pretrained downloads).  The restatement is cross-checked against torchvision's VisionTransformer by
weight remapping and against the parameter count the reference documents (86,390,787).
"""
