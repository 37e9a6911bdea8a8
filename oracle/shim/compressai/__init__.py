"""ORACLE / TEST INFRASTRUCTURE ONLY -- not part of the product path.

Pure-torch restatement of the parts of ``compressai==1.2.1`` (pinned in the
reference's ``env.txt:12``) that the DeepSVC structure/texture layer imports:

    image_model.py:4-8   EntropyBottleneck, GaussianConditional, subpel_conv3x3,
                         conv3x3, conv, deconv, update_registered_buffers,
                         ste_round, BufferedRansEncoder, RansDecoder
    modules.py:9         compressai.models.utils.conv
    video_model.py:5     compressai.entropy_models.EntropyBottleneck

compressai is an un-vendored third-party dependency that is absent from
``/root/reference`` and from this image (no network).  The arithmetic below
restates its published algorithm (entropy_models/entropy_models.py,
ops/ops.py, ops/bound_ops.py, layers/layers.py, models/utils.py of release
1.2.1).  PARITY UNPINNED at this boundary: the reference ships no golden
vectors for these ops and the real package cannot be executed here; the
restatement is anchored on the reference's call sites only.

With ``oracle/shim`` first on ``sys.path`` the reference's own
``modules.py`` / ``image_model.py`` / ``video_model.py`` import unmodified.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this package.
"""

__version__ = "1.2.1+oracle-shim"
