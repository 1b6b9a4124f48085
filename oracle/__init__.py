"""CPU oracle for the UnFlow hot path -- TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU, the reference algorithm for every function
on the hot path.  It exists so the CUDA product under
``unflow_b200/`` can be checked against something that follows the reference
line by line.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it; the product never
does (tests/test_no_oracle_in_product.py enforces that).

oracle/_ref/libref_ops.so (oracle/ref_kernels.py, oracle/ref_ops/, oracle/tf_stub/) holds the
reference's own CUDA kernels, compiled where they lie against stand-in TensorFlow headers: the GPU
ground truth for the four ops (built here, not yet run on a GPU).

Pinning status (see DESIGN.md "Oracle"):
  * custom ops (oracle_ops.c): pinned by the reference's own known-answer
    tests (tests/golden/reference_kats.json, transcribed from
    src/e2eflow/test/ops/*.py) and by independent float64 brute-force
    definitions (oracle/brute.py) for the cases the reference leaves unpinned.
  * image_warp / smoothness deltas / outgoing mask / gradient loss: pinned by
    the reference KATs in src/e2eflow/test/test_image_warp.py and
    src/e2eflow/test/test_losses.py.
  * ternary loss, compute_losses (all occlusion modes, values and flow
    gradients), flownet (C, S, stacked: every output of every network),
    unsupervised_loss (value, output flows, variable gradients), the
    augmentation cores (spatial transformer, random_affine, random_photometric
    with recorded draws): pinned against
    the reference's OWN Python source, executed unmodified under a
    TensorFlow-API stand-in (tests/golden/tf_shim.py ->
    tests/golden/make_reference_run.py -> tests/golden/reference_run.npz,
    tests/test_oracle_vs_reference_run.py).  That pins the graph -- op order,
    constants, masks, weights, pyramid bookkeeping, variable scopes / names /
    shapes, stop_gradient placement.
  * STILL PARITY UNPINNED: the arithmetic inside the TensorFlow-supplied
    primitives the graph rests on (SAME padding rule, legacy resize_bilinear,
    rgb_to_grayscale weights, conv2d_transpose) -- TensorFlow 1.x cannot run
    here and the reference holds no golden values for them; the stand-in and
    the oracle both restate the documented TF1 behaviour.  A second,
    independently written float64 pixel-loop definition of the loss assembly
    (oracle/brute.py, tests/test_oracle_losses_brute.py) guards against slips.
"""
