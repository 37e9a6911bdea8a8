"""ORACLE / TEST INFRASTRUCTURE ONLY.  Import stub for ``pytorch_msssim``
(pinned 0.2.1 in the reference's env.txt:58; absent from this image) so that
the reference's ``video_model.py:10`` imports unmodified.  MS-SSIM is a
distortion metric outside the warp+entropy hot path and
is deliberately not implemented."""


def ms_ssim(*args, **kwargs):
    raise NotImplementedError(
        "pytorch_msssim is not available; MS-SSIM distortion is outside the hot path")


def ssim(*args, **kwargs):
    raise NotImplementedError("pytorch_msssim is not available")
