"""CPU oracle for the UnFlow hot path -- TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU, the reference algorithm for every function
on the hot path.  It exists so the CUDA product under
``unflow_b200/`` can be checked against something that follows the reference
line by line.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it; the product never
does (tests/test_no_oracle_in_product.py enforces that).

Pinning status (see DESIGN.md "Oracle"):
  * custom ops (oracle_ops.c): pinned by the reference's own known-answer
    tests (tests/golden/reference_kats.json, transcribed from
    src/e2eflow/test/ops/*.py) and by independent float64 brute-force
    definitions (oracle/brute.py) for the cases the reference leaves unpinned.
  * image_warp / smoothness deltas / outgoing mask / gradient loss: pinned by
    the reference KATs in src/e2eflow/test/test_image_warp.py and
    src/e2eflow/test/test_losses.py.
  * ternary loss, compute_losses, flownet, unsupervised_loss and every
    TensorFlow-supplied primitive they rest on (SAME padding, legacy
    resize_bilinear, rgb_to_grayscale): PARITY UNPINNED -- the reference holds
    no usable golden values for them (its ternary test is dead code) and
    TensorFlow 1.x cannot run here.  The oracle restates the documented TF1
    behaviour.  The loss assembly is cross-checked against a second,
    independently written float64 pixel-loop definition (oracle/brute.py,
    tests/test_oracle_losses_brute.py); that is a self-consistency guard, not
    a reference pin.
"""
