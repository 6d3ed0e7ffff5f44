"""CPU oracle for the view-synthesis loss hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import it, and only as the checker / the timed CPU
baseline.  The product path (``baseboostdepth_b200``) never imports it and
fails loudly when its CUDA library is missing.

What it is: a restatement, in plain PyTorch tensor ops that run on the host,
of the reference's algorithm for this path (``layers.py`` and the loss methods
of ``trainer.py`` of kieran514/baseboostdepth), each function citing the
reference lines it follows.  The arithmetic of the reference lives in a
third-party dependency, PyTorch ATen (pinned ``pytorch=1.8.0`` in the
reference's ``environment.yml:162``; this image has torch 2.11): ``bmm``,
``grid_sampler_2d`` (bilinear / border / align_corners=True),
``reflection_pad2d``, ``avg_pool2d``, ``min``.  ``loss_path`` calls those same
ops at the reference's call sites; ``pixel_model`` restates their published
per-pixel algorithms (ATen ``GridSampler.h``) in float64 numpy loops for small
cases, including the analytic gradients.

Pinning: the reference has no tests, golden vectors or fixtures for this
path.  The oracle is therefore pinned against outputs of
the reference itself: ``tests/golden/make_golden.py`` imports the unmodified
reference from ``/root/reference`` (build container only), runs
``Trainer.generate_images_pred`` + ``compute_losses`` + ``backward`` on small
seeded inputs in plain / tri-min / tri-min+decomp / no-ssim modes and commits
inputs and outputs as ``tests/golden/*.npz``; ``tests/test_oracle_golden.py``
checks the oracle against every one of them.
"""
