"""B200-native view-synthesis loss for kieran514/baseboostdepth (the path SURVEY.md 8 scopes).

Layers, bottom up:

* ``csrc/`` + ``include/bbd_loss.h`` -- the sm_100a kernels behind a C ABI (``libbbd_loss.so``).
* ``_lib``     -- ctypes binding of that ABI; raises when the library is missing (no CPU fallback).
* ``layers``   -- tier A: the reference's ``layers.py`` classes/functions, kernel-backed.
* ``fused``    -- tier B: the whole loss of one step (forward + backward) as one autograd node.
* ``trainer``  -- ``FusedLossMixin``: the reference trainer's ``generate_images_pred`` /
  ``compute_losses`` signatures on top of ``fused``.
* ``plan`` / ``synthetic`` / ``staging`` -- candidate tables, bench workloads, host->device staging.

Nothing here imports ``oracle/``.
"""
__version__ = "0.1.0"
