"""CPU oracle for the FrameINO denoise-step hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Plain-PyTorch (CPU, fp32 or bf16-emulating) restatement of the reference's transformer forward, written as pure
functions over a diffusers-layout state dict. Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this package, and only as the checker or the timed CPU
baseline. The product path (``frameino_b200``) never imports it.

Pinning status: the reference repo holds no golden vectors for this path. The oracle is pinned
instead against outputs of the reference's OWN source files executed in the build container
(``tests/golden/make_golden.py`` imports ``/root/reference/architecture/*.py`` with a minimal stand-in for the
absent ``diffusers`` package); the classes that only exist upstream in diffusers (FeedForward, FP32LayerNorm,
RMSNorm, AdaLayerNorm, CogVideoXLayerNormZero, the mixins) are restated from recalled upstream semantics in that
stand-in, so for those pieces parity is "unpinned" in the strict sense.
"""
